"""Synthetic inputs for tests and bench.py (no network: no COCO, no trained weights).

* he_weights(seed)           seeded He-normal conv weights in the Chainer .npz layout
                             ("<layer>/W" [Cout,Cin,k,k] f32, "<layer>/b" [Cout] f32).
* random_images(n, h, w)     uint8 BGR noise frames.
* eight_person_maps(...)     full-resolution heatmaps/PAFs of eight stick figures, built
                             with the reference's ground-truth recipe
                             (coco_data_loader.py:208-214 Gaussian joints,
                             :232-268 constant-vector PAF band, averaged on overlap).
"""
import numpy as np

from .models.CocoPoseNet import LAYERS

# unit-box template (x, y), y downwards, JointType order (entity.py:9-46)
_TEMPLATE = np.array([
    (0.50, 0.10), (0.50, 0.22), (0.36, 0.22), (0.30, 0.38), (0.27, 0.52), (0.64, 0.22), (0.70, 0.38),
    (0.73, 0.52), (0.42, 0.52), (0.41, 0.72), (0.40, 0.92), (0.58, 0.52), (0.59, 0.72), (0.60, 0.92),
    (0.46, 0.07), (0.54, 0.07), (0.41, 0.09), (0.59, 0.09)], np.float64)

_LIMBS = ((1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13), (1, 2), (2, 3), (3, 4), (2, 16),
          (1, 5), (5, 6), (6, 7), (5, 17), (1, 0), (0, 14), (0, 15), (14, 16), (15, 17))


def he_weights(seed=0, bias_scale=0.02, gain=2.0, layers=None):
    """W ~ N(0, sqrt(gain/fan_in)), b ~ N(0, bias_scale); one RandomState stream in the
    reference's layer declaration order (models/CocoPoseNet.py:26-129; `layers` = another net's table)."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, cin, cout, k in (LAYERS if layers is None else layers):
        fan_in = cin * k * k
        out[name + "/W"] = (rs.standard_normal((cout, cin, k, k)) * np.sqrt(gain / fan_in)).astype(np.float32)
        out[name + "/b"] = (rs.standard_normal(cout) * bias_scale).astype(np.float32)
    return out


def random_images(n, h=368, w=656, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


def person_layout(map_h=320, map_w=576, n_person=8, seed=0):
    """Integer joint coordinates [P,18,2] (x,y): two rows of stick figures."""
    rs = np.random.RandomState(seed)
    per_row = (n_person + 1) // 2
    cell_w = map_w / per_row
    cell_h = map_h / 2.0
    people = []
    for p in range(n_person):
        r, c = divmod(p, per_row)
        size = cell_h * rs.uniform(0.80, 0.92)
        x0 = c * cell_w + (cell_w - size) / 2 + rs.uniform(-6, 6)
        y0 = r * cell_h + (cell_h - size) / 2 + rs.uniform(-3, 3)
        pts = _TEMPLATE * size + np.array([x0, y0])
        pts += rs.uniform(-2, 2, pts.shape)
        pts = np.round(pts).astype(np.int64)
        pts[:, 0] = pts[:, 0].clip(1, map_w - 2)
        pts[:, 1] = pts[:, 1].clip(1, map_h - 2)
        people.append(pts)
    return np.stack(people)


def eight_person_maps(map_h=320, map_w=576, n_person=8, seed=0, sigma=3.0, band=3.0, noise=0.002):
    """Returns (pafs [38,H,W] f32, heatmaps [19,H,W] f32, joints [P,18,2] int)."""
    joints = person_layout(map_h, map_w, n_person, seed)
    yy, xx = np.mgrid[0:map_h, 0:map_w].astype(np.float64)
    heat = np.zeros((19, map_h, map_w), np.float64)
    for j in range(18):
        for p in range(len(joints)):
            jx, jy = joints[p, j]
            heat[j] = np.maximum(heat[j], np.exp(-((xx - jx) ** 2 + (yy - jy) ** 2) / (2 * sigma * sigma)))
    heat[18] = 1.0 - heat[:18].max(axis=0)
    paf = np.zeros((38, map_h, map_w), np.float64)
    for l, (ja, jb) in enumerate(_LIMBS):
        cnt = np.zeros((map_h, map_w), np.float64)
        for p in range(len(joints)):
            a = joints[p, ja].astype(np.float64)
            b = joints[p, jb].astype(np.float64)
            v = b - a
            n = np.hypot(v[0], v[1])
            if n == 0:
                continue
            u = v / n
            along = (xx - a[0]) * u[0] + (yy - a[1]) * u[1]
            perp = np.abs((xx - a[0]) * u[1] - (yy - a[1]) * u[0])
            m = (along >= 0) & (along <= n) & (perp <= band)
            paf[2 * l][m] += u[0]
            paf[2 * l + 1][m] += u[1]
            cnt[m] += 1
        nz = cnt > 0
        paf[2 * l][nz] /= cnt[nz]
        paf[2 * l + 1][nz] /= cnt[nz]
    rs = np.random.RandomState(seed + 12345)
    heat += rs.standard_normal(heat.shape) * noise
    paf += rs.standard_normal(paf.shape) * noise
    return paf.astype(np.float32), heat.astype(np.float32), joints


def procedural_image(h, w, seed=0):
    """A smooth, seeded BGR uint8 'scene' (low-frequency colour field + blobs + mild noise).
    Stand-in for the reference's data/*.png, which the licence does not allow copying into
    this repository (LICENSE:28) and which do not exist on the GPU box."""
    rs = np.random.RandomState(seed)
    gh, gw = max(h // 48, 2), max(w // 48, 2)
    coarse = rs.uniform(0, 255, (gh, gw, 3))
    # separable linear interpolation of the coarse grid (numpy only, deterministic)
    yi = np.linspace(0, gh - 1, h)
    xi = np.linspace(0, gw - 1, w)
    y0 = np.clip(np.floor(yi).astype(int), 0, gh - 2)
    x0 = np.clip(np.floor(xi).astype(int), 0, gw - 2)
    fy = (yi - y0)[:, None, None]
    fx = (xi - x0)[None, :, None]
    img = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0][:, x0 + 1] * (1 - fy) * fx +
           coarse[y0 + 1][:, x0] * fy * (1 - fx) + coarse[y0 + 1][:, x0 + 1] * fy * fx)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(12):
        cy, cx = rs.uniform(0, h), rs.uniform(0, w)
        r = rs.uniform(0.03, 0.12) * min(h, w)
        col = rs.uniform(0, 255, 3)
        m = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))[..., None]
        img = img * (1 - m) + col * m
    img += rs.standard_normal(img.shape) * 6.0
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def eight_person_lowres(h=46, w=82, n_person=8, seed=0):
    """Network-output-resolution (H/8 x W/8) maps of eight stick figures: what is injected as
    the CocoPoseNet output for the full-pipeline benchmark (random weights cannot produce
    people).  Joint sigma ~1 low-res pixel (training uses heatmap_sigma 7 at 368 px,
    entity.py:61), PAF band +-1 pixel.  Returns (paf [38,h,w] f32, heat [19,h,w] f32)."""
    paf, heat, _ = eight_person_maps(h, w, n_person, seed, sigma=1.0, band=1.0, noise=0.002)
    return paf, heat
