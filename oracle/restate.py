"""ORACLE / TEST INFRASTRUCTURE ONLY -- a CPU restatement of the reference hot path.

Nothing in the product package imports this file.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
use it, and only as the checker / reported CPU baseline.

What is restated (reference file:line, relative to the reference tree):
  pose_detector.py:46-55    pad_image
  pose_detector.py:57-73    compute_optimal_size
  pose_detector.py:75-110   compute_peaks_from_heatmaps (CPU branch)
  pose_detector.py:135-181  compute_candidate_connections / compute_connections
  pose_detector.py:183-250  grouping_key_points
  pose_detector.py:252-265  subsets_to_pose_array
  pose_detector.py:426-431  preprocess
  pose_detector.py:433-482  detect_precise
  pose_detector.py:484-517  __call__ (fast path)
  models/CocoPoseNet.py:26-129,132-262   layer table + forward graph
  entity.py:71-105          inference constants, limbs_point

Third-party arithmetic that is NOT under the reference tree and is restated from
its published algorithm (pins = versions in this image):
  Chainer (unpinned, "2.0+")  conv / max-pool(cover_all) / concat -> torch 2.11 CPU fp32;
                              resize_images -> align-corners bilinear, float64 grid,
                              weights cast to float32, 4-term float32 sum.
  SciPy 1.18.1 gaussian_filter(sigma=2.5): 21 normalised taps, mode='reflect'
                              (symmetric), axis 0 then axis 1, float64 accumulate in
                              the order  c*w0 + sum_{j=-10..-1} (x[j]+x[-j])*w[j],
                              float32 store after each pass.  Verified bit-exact
                              against scipy in tests/test_oracle.py.
  OpenCV 4.13.0 resize        used directly (precise path), not restated.
  NumPy 2.3.5                 linspace / round-half-even / matmul (the (10,2)x(2,)
                              dot evaluates as fma(p0,u0,p1*u1) on this host).

Parity status: the reference has no tests or golden vectors, so
parity is pinned against outputs of the reference's own files executed verbatim over
oracle/chainer_stub (oracle/make_goldens.py -> tests/golden/*.npz); the third-party
boundaries above remain "parity unpinned" in the strict sense.
"""
import math

import numpy as np

# ----------------------------------------------------------------------------
# constants (entity.py:71-105)
# ----------------------------------------------------------------------------
N_JOINTS = 18
INFERENCE_IMG_SIZE = 368
INFERENCE_SCALES = (0.5, 1, 1.5, 2)
HEATMAP_SIZE = 320
DOWNSCALE = 8
GAUSSIAN_SIGMA = 2.5
N_INTEG_POINTS = 10
N_INTEG_POINTS_THRESH = 8
HEATMAP_PEAK_THRESH = 0.05
INNER_PRODUCT_THRESH = 0.05
LIMB_LENGTH_RATIO = 1.0
LENGTH_PENALTY_VALUE = 1
N_SUBSET_LIMBS_THRESH = 3
SUBSET_SCORE_THRESH = 0.2
# (joint_a, joint_b) per limb, entity.py:85-105
LIMBS = ((1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13), (1, 2), (2, 3), (3, 4), (2, 16),
         (1, 5), (5, 6), (6, 7), (5, 17), (1, 0), (0, 14), (0, 15), (14, 16), (15, 17))
NO_NEW_SUBSET_LIMBS = (9, 13)  # pose_detector.py:237


# ----------------------------------------------------------------------------
# network (models/CocoPoseNet.py)
# ----------------------------------------------------------------------------
def layer_table():
    """[(name, cin, cout, ksize)] in the reference's declaration order
    (models/CocoPoseNet.py:26-129)."""
    t = [("conv1_1", 3, 64, 3), ("conv1_2", 64, 64, 3), ("conv2_1", 64, 128, 3), ("conv2_2", 128, 128, 3),
         ("conv3_1", 128, 256, 3), ("conv3_2", 256, 256, 3), ("conv3_3", 256, 256, 3), ("conv3_4", 256, 256, 3),
         ("conv4_1", 256, 512, 3), ("conv4_2", 512, 512, 3), ("conv4_3_CPM", 512, 256, 3),
         ("conv4_4_CPM", 256, 128, 3)]
    for br, nout in (("L1", 38), ("L2", 19)):
        for i in (1, 2, 3):
            t.append(("conv5_%d_CPM_%s" % (i, br), 128, 128, 3))
        t.append(("conv5_4_CPM_%s" % br, 128, 512, 1))
        t.append(("conv5_5_CPM_%s" % br, 512, nout, 1))
    for s in range(2, 7):
        for br, nout in (("L1", 38), ("L2", 19)):
            t.append(("Mconv1_stage%d_%s" % (s, br), 185, 128, 7))
            for i in (2, 3, 4, 5):
                t.append(("Mconv%d_stage%d_%s" % (i, s, br), 128, 128, 7))
            t.append(("Mconv6_stage%d_%s" % (s, br), 128, 128, 1))
            t.append(("Mconv7_stage%d_%s" % (s, br), 128, nout, 1))
    return t


def load_weights_npz(path):
    """Chainer save_npz layout: '<layer>/W' [Cout,Cin,k,k] f32, '<layer>/b' [Cout]."""
    out = {}
    with np.load(path) as f:
        for name, _, _, _ in layer_table():
            out[name] = (np.ascontiguousarray(f[name + "/W"], np.float32),
                         np.ascontiguousarray(f[name + "/b"], np.float32))
    return out


def forward(weights, x, all_stages=False):
    """models/CocoPoseNet.py:132-262 with torch CPU fp32.  x: [N,3,H,W] f32.
    Returns (pafs[-1] [N,38,h,w], heatmaps[-1] [N,19,h,w]) (or per-stage lists)."""
    import torch
    import torch.nn.functional as F

    def conv(name, h, relu=True):
        W, b = weights[name]
        y = F.conv2d(h, torch.from_numpy(W), torch.from_numpy(b), stride=1, padding=(W.shape[2] - 1) // 2)
        return torch.relu(y) if relu else y

    with torch.no_grad():
        h = torch.from_numpy(np.ascontiguousarray(x, np.float32))
        h = conv("conv1_2", conv("conv1_1", h))
        h = F.max_pool2d(h, 2, 2, ceil_mode=True)            # cover_all=True [3p]
        h = conv("conv2_2", conv("conv2_1", h))
        h = F.max_pool2d(h, 2, 2, ceil_mode=True)
        for n in ("conv3_1", "conv3_2", "conv3_3", "conv3_4"):
            h = conv(n, h)
        h = F.max_pool2d(h, 2, 2, ceil_mode=True)
        for n in ("conv4_1", "conv4_2", "conv4_3_CPM", "conv4_4_CPM"):
            h = conv(n, h)
        feat = h
        pafs, heats = [], []
        outs = []
        for br in ("L1", "L2"):
            g = feat
            for i in (1, 2, 3, 4):
                g = conv("conv5_%d_CPM_%s" % (i, br), g)
            outs.append(conv("conv5_5_CPM_%s" % br, g, relu=False))
        pafs.append(outs[0]); heats.append(outs[1])
        for s in range(2, 7):
            cat = torch.cat((pafs[-1], heats[-1], feat), dim=1)   # order (h1, h2, feature_map), :168
            outs = []
            for br in ("L1", "L2"):
                g = cat
                for i in (1, 2, 3, 4, 5, 6):
                    g = conv("Mconv%d_stage%d_%s" % (i, s, br), g)
                outs.append(conv("Mconv7_stage%d_%s" % (s, br), g, relu=False))
            pafs.append(outs[0]); heats.append(outs[1])
    if all_stages:
        return [p.numpy() for p in pafs], [h.numpy() for h in heats]
    return pafs[-1].numpy(), heats[-1].numpy()


# ----------------------------------------------------------------------------
# host helpers
# ----------------------------------------------------------------------------
def compute_optimal_size(orig_img, img_size, stride=8):
    """pose_detector.py:57-73: short side = img_size, long side rounded (half-even)
    then rounded UP to a multiple of stride.  Returns (w, h)."""
    h0, w0 = orig_img.shape[:2]
    aspect = h0 / w0
    if h0 < w0:
        h = img_size
        w = int(np.round(img_size / aspect))
        if w % stride:
            w += stride - w % stride
    else:
        w = img_size
        h = int(np.round(img_size * aspect))
        if h % stride:
            h += stride - h % stride
    return (w, h)


def preprocess(img):
    """pose_detector.py:426-431: float32, /255, -0.5, HWC->CHW, add batch."""
    x = img.astype(np.float32)
    x /= 255
    x -= 0.5
    return x.transpose(2, 0, 1)[None]


def pad_image(img, stride, pad_value):
    """pose_detector.py:46-55: pad bottom/right to a multiple of stride."""
    h, w = img.shape[:2]
    ph = (stride - h % stride) % stride
    pw = (stride - w % stride) % stride
    out = np.zeros((h + ph, w + pw, 3), "uint8") + np.asarray(pad_value)
    out[:h, :w] = img
    return out, [ph, pw]


def cv2_resize_linear_u8(img, dsize):
    """Restatement of cv2.resize(img_uint8, (W, H)) with the default INTER_LINEAR (pose_detector.py:493)
    [3p OpenCV 4.13 imgproc/resize.cpp, 8-bit fixed-point path]: per axis fx = float((d+0.5)*scale-0.5),
    s = floor(fx), fx -= s; COLUMNS clamp (s<0 -> s=0,fx=0; s>=w-1 -> s=w-1,fx=0), ROWS keep fx and clip the
    two source rows separately; coefficients = rint(c*2048) as int16; horizontal pass in int32
    D = S[s]*a0 + S[s+1]*a1; vertical pass ((b0*(D0>>4))>>16) + ((b1*(D1>>4))>>16) + 2) >> 2.
    Verified bit-exact against cv2 in tests/test_oracle.py."""
    W, H = int(dsize[0]), int(dsize[1])
    h0, w0 = img.shape[:2]
    if (W, H) == (w0, h0):
        return img.copy()

    def axis(dst_n, src_n, clamp):
        scale = 1.0 / (float(dst_n) / src_n)
        f = ((np.arange(dst_n, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s0 = np.floor(f).astype(np.int32)
        f = (f - s0.astype(np.float32)).astype(np.float32)
        if clamp:
            lo = s0 < 0
            f[lo] = 0; s0[lo] = 0
            hi = s0 >= src_n - 1
            f[hi] = 0; s0[hi] = src_n - 1
        c0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
        c1 = np.rint(f * np.float32(2048)).astype(np.int32)
        return np.clip(s0, 0, src_n - 1), np.clip(s0 + 1, 0, src_n - 1), c0, c1

    sx, sx1, a0, a1 = axis(W, w0, True)
    sy, sy1, b0, b1 = axis(H, h0, False)
    S = img.astype(np.int32).reshape(h0, w0, -1)
    D = S[:, sx] * a0[None, :, None] + S[:, sx1] * a1[None, :, None]
    out = ((((b0[:, None, None] * (D[sy] >> 4)) >> 16) + ((b1[:, None, None] * (D[sy1] >> 4)) >> 16) + 2) >> 2)
    return out.astype(np.uint8).reshape((H, W) + img.shape[2:])


def cv2_resize_cubic_u8(img, dsize):
    """cv2.resize(img, dsize, interpolation=cv2.INTER_CUBIC) for uint8 HxWxC (pose_detector.py:443) as OpenCV's OWN
    8-bit path computes it (imgproc/resize.cpp, OpenCV 4.13; what runs with cv2.ipp.setUseIPP(False) or without IPP):
    per axis f = float((d+0.5)*scale - 0.5), s = floor(f), f -= s; Keys taps (A = -0.75) in float32 -> int16
    rint(tap*2048); horizontal pass int32 with replicated borders; vertical pass in float32 for the 8-lane SIMD body
    ((D3*b3 + D2*b2) + D1*b1) + D0*b0 with b = beta/2^22, unfused, rint, saturate -- and FixedPtCast<int,uchar,22> for
    the last (W*C) % 8 elements of each row.  Pinned bit-exact against cv2 (IPP off) in tests/test_oracle.py."""
    dw, dh = dsize
    sh, sw = img.shape[:2]
    f32 = np.float32

    def axis(dn, sn):
        scale = 1.0 / (dn / sn)
        f = ((np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5).astype(f32)
        s = np.floor(f)
        x = (f - s.astype(f32)).astype(f32)
        A = f32(-0.75)
        t, u = x + f32(1), f32(1) - x
        c0 = ((A * t - f32(5) * A) * t + f32(8) * A) * t - f32(4) * A
        c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + f32(1)
        c2 = ((A + f32(2)) * u - (A + f32(3))) * u * u + f32(1)
        c3 = f32(1) - c0 - c1 - c2
        co = np.rint(np.stack([c0, c1, c2, c3], 1) * f32(2048)).astype(np.int64)
        return s.astype(np.int64), co

    xo, xa = axis(dw, sw)
    yo, yb = axis(dh, sh)
    if (dh, dw) == (sh, sw):
        return img.copy()
    S = img.astype(np.int64)
    idx = np.clip(xo[:, None] + np.arange(-1, 3)[None, :], 0, sw - 1)
    D = (S[:, idx, :] * xa[None, :, :, None]).sum(2)                       # [sh, dw, C] int32-range
    idy = np.clip(yo[:, None] + np.arange(-1, 3)[None, :], 0, sh - 1)
    R = D[idy]                                                             # [dh, 4, dw, C]
    fixed = ((R * yb[:, :, None, None]).sum(1) + (1 << 21)) >> 22
    b = yb.astype(f32) * f32(1.0 / 4194304.0)
    Rf = R.astype(f32)
    acc = Rf[:, 3] * b[:, 3, None, None]
    for k in (2, 1, 0):
        acc = Rf[:, k] * b[:, k, None, None] + acc
    flt = np.rint(acc).astype(np.int64)
    C = img.shape[2]
    simd_end = (dw * C) & ~7
    out = np.where((np.arange(dw * C) < simd_end).reshape(1, dw, C), flt, fixed)
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_bilinear_align_corners(x, out_hw):
    """Chainer resize_images [3p] (pose_detector.py:501-502).  x: [B,C,H,W] f32."""
    B, C, H, W = x.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    u = np.linspace(0, W - 1, num=ow)
    v = np.linspace(0, H - 1, num=oh)
    u0 = np.floor(u).astype(np.int32).clip(0, W - 2)
    v0 = np.floor(v).astype(np.int32).clip(0, H - 2)
    u1, v1 = u0 + 1, v0 + 1
    U, V = np.meshgrid(u, v)
    U0, V0 = np.meshgrid(u0, v0)
    U1, V1 = U0 + 1, V0 + 1
    w1 = ((U1 - U) * (V1 - V)).astype(x.dtype)
    w2 = ((U - U0) * (V1 - V)).astype(x.dtype)
    w3 = ((U1 - U) * (V - V0)).astype(x.dtype)
    w4 = ((U - U0) * (V - V0)).astype(x.dtype)
    a = x.reshape(B * C, H, W)
    y = w1[None] * a[:, v0][:, :, u0]
    y += w2[None] * a[:, v0][:, :, u1]
    y += w3[None] * a[:, v1][:, :, u0]
    y += w4[None] * a[:, v1][:, :, u1]
    return y.reshape(B, C, oh, ow)


# ----------------------------------------------------------------------------
# peaks (pose_detector.py:75-110)
# ----------------------------------------------------------------------------
def gaussian_taps(sigma=GAUSSIAN_SIGMA, truncate=4.0):
    """scipy _gaussian_kernel1d(sigma, 0, radius) [3p]: 2*radius+1 float64 taps."""
    radius = int(truncate * float(sigma) + 0.5)
    xs = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * xs ** 2)
    return phi / phi.sum()


def _correlate1d_symmetric(a32, taps, axis):
    """scipy NI_Correlate1D symmetric branch [3p]: float64 line buffer, reflect
    ('symmetric') extension, tmp = x0*w0; for j=-r..-1: tmp += (x[j]+x[-j])*w[j];
    result stored back to float32."""
    a = np.moveaxis(a32, axis, -1).astype(np.float64)
    r = len(taps) // 2
    n = a.shape[-1]
    pad = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(r, r)], mode="symmetric")
    acc = pad[..., r:r + n] * taps[r]
    for j in range(-r, 0):
        acc = acc + (pad[..., r + j:r + j + n] + pad[..., r - j:r - j + n]) * taps[r + j]
    return np.moveaxis(acc.astype(np.float32), -1, axis)


def gaussian_smooth(maps):
    """gaussian_filter(heatmap, sigma=2.5) for each [H,W] map of maps[..., H, W]
    (axis -2 pass, then axis -1 pass, float32 between passes)."""
    taps = gaussian_taps()
    return _correlate1d_symmetric(_correlate1d_symmetric(maps, taps, maps.ndim - 2), taps, maps.ndim - 1)


def compute_peaks_from_heatmaps(heatmaps):
    """heatmaps [C+1,H,W] f32 -> all_peaks [N,5] float64 rows (type, x, y, score, id).
    Background channel dropped (:78); strict > vs the four axial neighbours with zero
    outside the image (:87-102); threshold strict > 0.05; order = channel, then
    row-major (y, x) (:104); score is the smoothed value."""
    hm = np.asarray(heatmaps)[:-1]
    g = gaussian_smooth(hm.astype(np.float32, copy=False))
    z = np.zeros_like(g, dtype=np.float64)
    up, dn, lf, rt = z.copy(), z.copy(), z.copy(), z.copy()
    up[:, 1:, :] = g[:, :-1, :]
    dn[:, :-1, :] = g[:, 1:, :]
    lf[:, :, 1:] = g[:, :, :-1]
    rt[:, :, :-1] = g[:, :, 1:]
    mask = (g > HEATMAP_PEAK_THRESH) & (g > up) & (g > dn) & (g > lf) & (g > rt)
    c, y, x = np.nonzero(mask)
    n = len(c)
    if n == 0:
        return np.array([])
    out = np.empty((n, 5), np.float64)
    out[:, 0] = c
    out[:, 1] = x
    out[:, 2] = y
    out[:, 3] = g[c, y, x]
    out[:, 4] = np.arange(n)
    return out


# ----------------------------------------------------------------------------
# connections (pose_detector.py:135-181)
# ----------------------------------------------------------------------------
def candidate_connections(paf, cand_a, cand_b, img_len):
    """paf [2,H,W] f32; cand_* [n,4] (x,y,score,id) float64.  Returns the sorted
    candidate list as an [m,3] float64 array (id_a, id_b, score), descending score,
    ties in (a-major, b-minor) order (stable sort, :158)."""
    na, nb = len(cand_a), len(cand_b)
    ax = np.repeat(cand_a[:, 0], nb); ay = np.repeat(cand_a[:, 1], nb)
    bx = np.tile(cand_b[:, 0], na); by = np.tile(cand_b[:, 1], na)
    ida = np.repeat(cand_a[:, 3], nb); idb = np.tile(cand_b[:, 3], na)
    vx, vy = bx - ax, by - ay
    norm = np.sqrt(vx * vx + vy * vy)                       # np.linalg.norm of a 2-vector
    keep = norm != 0                                        # :141
    ax, ay, bx, by, ida, idb, vx, vy, norm = [v[keep] for v in (ax, ay, bx, by, ida, idb, vx, vy, norm)]
    if len(norm) == 0:
        return np.zeros((0, 3))
    ys = _linspace_per_pair(ay, by)                         # [10, m]   :144
    xs = _linspace_per_pair(ax, bx)                         # :145
    yi = ys.round().astype("i")                             # half-even, int32  :146
    xi = xs.round().astype("i")
    p = np.stack([paf[0][yi, xi], paf[1][yi, xi]], axis=-1)  # [10, m, 2] f32   :147
    unit = np.stack([vx / norm, vy / norm], axis=-1)        # [m, 2] f64       :148
    # :149 np.dot((10,2) f32, (2,) f64): batched matmul reproduces the per-pair dot bit-for-bit
    ip = np.matmul(p.transpose(1, 0, 2), unit[:, :, None])[..., 0]   # [m, 10] f64
    integ = ip.sum(axis=1) / N_INTEG_POINTS                 # :151 (pairwise order for n=10)
    prior = np.minimum(LIMB_LENGTH_RATIO * img_len / norm - LENGTH_PENALTY_VALUE, 0)  # :153
    score = integ + prior
    nvalid = (ip > INNER_PRODUCT_THRESH).sum(axis=1)        # :155
    ok = (nvalid > N_INTEG_POINTS_THRESH) & (score > 0)     # :156
    ida, idb, score = ida[ok], idb[ok], score[ok]
    order = np.argsort(-score, kind="stable")               # stable, descending   :158
    return np.stack([ida[order], idb[order], score[order]], axis=1) if len(order) else np.zeros((0, 3))


def _linspace_per_pair(a, b, num=N_INTEG_POINTS):
    """np.linspace(a_k, b_k, num) evaluated with the *scalar-call* semantics the reference
    uses (one call per pair, :144-145): y_i = i*step + a with step = (b-a)/(num-1);
    if step == 0 the numpy code path is (i/(num-1))*(b-a) + a; y_last = b exactly."""
    i = np.arange(float(num))[:, None]
    delta = b - a
    step = delta / (num - 1)
    y = np.where(step == 0, (i / (num - 1)) * delta, i * step) + a
    y[-1] = b
    return y


def compute_connections(pafs, all_peaks, img_len):
    """pafs [38,H,W] f32, all_peaks [N,5] -> list of 19 arrays [k,3] (:161-181)."""
    out = []
    for l, (ja, jb) in enumerate(LIMBS):
        paf = pafs[2 * l:2 * l + 2]
        cand_a = all_peaks[all_peaks[:, 0] == ja][:, 1:]
        cand_b = all_peaks[all_peaks[:, 0] == jb][:, 1:]
        if len(cand_a) == 0 or len(cand_b) == 0:
            out.append(np.zeros((0, 3)))
            continue
        cands = candidate_connections(paf, cand_a, cand_b, img_len)
        limit = min(len(cand_a), len(cand_b))
        used_a, used_b, rows = set(), set(), []
        for ia, ib, s in cands:
            if ia not in used_a and ib not in used_b:
                rows.append((ia, ib, s)); used_a.add(ia); used_b.add(ib)
                if len(rows) >= limit:
                    break
        out.append(np.array(rows, np.float64).reshape(-1, 3))
    return out


# ----------------------------------------------------------------------------
# grouping (pose_detector.py:183-250)
# ----------------------------------------------------------------------------
def grouping_key_points(all_connections, peaks):
    """Sequential merge of limb connections into subsets [P,20] float64:
    18 peak ids (-1 = none), total score, joint count."""
    subsets = []                                            # list of float64[20] rows
    for l, conns in enumerate(all_connections):
        ja, jb = LIMBS[l]
        for ia, ib, score in np.asarray(conns).reshape(-1, 3):
            ia, ib = int(ia), int(ib)
            found = [k for k, s in enumerate(subsets) if s[ja] == ia or s[jb] == ib]
            if len(found) >= 3:
                # reference writes joint_found_subset_index[2] -> IndexError (:197)
                raise IndexError("list assignment index out of range")
            if len(found) == 1:
                s = subsets[found[0]]
                if s[jb] != ib:
                    s[jb] = ib
                    s[-1] += 1
                    s[-2] += peaks[ib, 3] + score
            elif len(found) == 2:
                s1, s2 = subsets[found[0]], subsets[found[1]]
                both = ((s1[:-2] >= 0).astype(int) + (s2[:-2] >= 0).astype(int)) == 2
                if not both.any():
                    s1[:-2] += s2[:-2] + 1
                    s1[-2:] += s2[-2:]
                    s1[-2:] += score                        # quirk: also added to the count (:217)
                    del subsets[found[1]]
                else:
                    for s in (s1, s2):
                        if s[ja] == -1:
                            s[ja] = ia; s[-1] += 1; s[-2] += peaks[ia, 3] + score
                        elif s[jb] == -1:
                            s[jb] = ib; s[-1] += 1; s[-2] += peaks[ib, 3] + score
            elif l not in NO_NEW_SUBSET_LIMBS:
                row = -1 * np.ones(20)
                row[ja] = ia
                row[jb] = ib
                row[-1] = 2
                row[-2] = (0 + peaks[ia, 3] + peaks[ib, 3]) + score   # sum([..]) + score (:242)
                subsets.append(row)
    arr = np.array(subsets, np.float64).reshape(-1, 20)
    keep = (arr[:, -1] >= N_SUBSET_LIMBS_THRESH) & (arr[:, -2] / arr[:, -1] >= SUBSET_SCORE_THRESH)
    return arr[keep]


def subsets_to_pose_array(subsets, peaks):
    """:252-265. NOTE: P == 0 gives np.array([]) of shape (0,), as in the reference."""
    people = []
    for s in subsets:
        joints = []
        for j in s[:18].astype("i"):
            if j >= 0:
                joints.append([peaks[j, 1], peaks[j, 2], 2])
            else:
                joints.append([0, 0, 0])
        people.append(np.array(joints))
    return np.array(people)


# ----------------------------------------------------------------------------
# end to end
# ----------------------------------------------------------------------------
def postprocess_fast(pafs, heatmaps, map_w, orig_w, orig_h, map_h, return_parts=False):
    """pose_detector.py:508-517 on full-resolution maps."""
    peaks = compute_peaks_from_heatmaps(heatmaps)
    if len(peaks) == 0:
        res = (np.empty((0, N_JOINTS, 3)), np.empty(0))
        return res + (None,) if return_parts else res
    conns = compute_connections(pafs, peaks, map_w)
    subsets = grouping_key_points(conns, peaks)
    raw_peaks = peaks.copy()
    peaks[:, 1] *= orig_w / map_w
    peaks[:, 2] *= orig_h / map_h
    poses = subsets_to_pose_array(subsets, peaks)
    scores = subsets[:, -2]
    if return_parts:
        return poses, scores, dict(all_peaks=raw_peaks, connections=conns, subsets=subsets)
    return poses, scores


def detect_fast(weights, img, return_parts=False):
    """pose_detector.py:484-517."""
    import cv2
    img = img.copy()
    oh, ow = img.shape[:2]
    in_w, in_h = compute_optimal_size(img, INFERENCE_IMG_SIZE)
    map_w, map_h = compute_optimal_size(img, HEATMAP_SIZE)
    x = preprocess(cv2.resize(img, (in_w, in_h)))
    paf_lo, heat_lo = forward(weights, x)
    pafs = resize_bilinear_align_corners(paf_lo, (map_h, map_w))[0]
    heat = resize_bilinear_align_corners(heat_lo, (map_h, map_w))[0]
    res = postprocess_fast(pafs, heat, map_w, ow, oh, map_h, return_parts)
    if return_parts:
        if res[2] is not None:
            res[2].update(paf_lo=paf_lo[0], heat_lo=heat_lo[0], pafs=pafs, heatmaps=heat)
        else:
            res = res[:2] + (dict(paf_lo=paf_lo[0], heat_lo=heat_lo[0], pafs=pafs, heatmaps=heat),)
    return res


def precise_maps(weights, img, forward_fn=None):
    """pose_detector.py:433-470: 4-scale forward, bicubic upsample x2, average."""
    import cv2
    fwd = forward if forward_fn is None else forward_fn
    oh, ow = img.shape[:2]
    paf_sum, heat_sum = 0, 0
    for scale in INFERENCE_SCALES:
        m = scale * INFERENCE_IMG_SIZE / min(img.shape[:2])
        im = cv2.resize(img, (math.ceil(ow * m), math.ceil(oh * m)), interpolation=cv2.INTER_CUBIC)
        padded, pad = pad_image(im, DOWNSCALE, (104, 117, 123))
        p_lo, h_lo = fwd(weights, preprocess(padded))
        p = p_lo[0].transpose(1, 2, 0)
        h = h_lo[0].transpose(1, 2, 0)
        ph, pw = padded.shape[:2]
        p = cv2.resize(p, (pw, ph), interpolation=cv2.INTER_CUBIC)
        p = p[:ph - pad[0], :pw - pad[1], :]
        paf_sum = paf_sum + cv2.resize(p, (ow, oh), interpolation=cv2.INTER_CUBIC)
        h = cv2.resize(h, (0, 0), fx=DOWNSCALE, fy=DOWNSCALE, interpolation=cv2.INTER_CUBIC)
        h = h[:ph - pad[0], :pw - pad[1], :]
        heat_sum = heat_sum + cv2.resize(h, (ow, oh), interpolation=cv2.INTER_CUBIC)
    pafs = (paf_sum / len(INFERENCE_SCALES)).transpose(2, 0, 1)
    heat = (heat_sum / len(INFERENCE_SCALES)).transpose(2, 0, 1)
    return pafs, heat


def detect_precise(weights, img, return_parts=False):
    """pose_detector.py:433-482 (img_len = original width, no coordinate rescale)."""
    oh, ow = img.shape[:2]
    pafs, heat = precise_maps(weights, img)
    peaks = compute_peaks_from_heatmaps(heat)
    if len(peaks) == 0:
        res = (np.empty((0, N_JOINTS, 3)), np.empty(0))
        return res + (dict(pafs=pafs, heatmaps=heat),) if return_parts else res
    conns = compute_connections(pafs, peaks, ow)
    subsets = grouping_key_points(conns, peaks)
    poses = subsets_to_pose_array(subsets, peaks)
    if return_parts:
        return poses, subsets[:, -2], dict(pafs=pafs, heatmaps=heat, all_peaks=peaks, connections=conns,
                                            subsets=subsets)
    return poses, subsets[:, -2]


# ----------------------------------------------------------------------------
# face / hand keypoint nets (models/FaceNet.py, models/HandNet.py, face_detector.py, hand_detector.py)
# ----------------------------------------------------------------------------
KEYPOINT_PEAK_THRESH = 0.1          # entity.py:128,144 (face_/hand_heatmap_peak_thresh)
KEYPOINT_IMG_SIZE = 368             # entity.py:127,143


def keypoint_layer_table(n_out):
    """[(name, cin, cout, ksize)] of FaceNet (n_out 71) / HandNet (n_out 22), models/FaceNet.py:10-76."""
    t = [("conv1_1", 3, 64, 3), ("conv1_2", 64, 64, 3), ("conv2_1", 64, 128, 3), ("conv2_2", 128, 128, 3),
         ("conv3_1", 128, 256, 3), ("conv3_2", 256, 256, 3), ("conv3_3", 256, 256, 3), ("conv3_4", 256, 256, 3),
         ("conv4_1", 256, 512, 3), ("conv4_2", 512, 512, 3), ("conv4_3", 512, 512, 3), ("conv4_4", 512, 512, 3),
         ("conv5_1", 512, 512, 3), ("conv5_2", 512, 512, 3), ("conv5_3_CPM", 512, 128, 3),
         ("conv6_1_CPM", 128, 512, 1), ("conv6_2_CPM", 512, n_out, 1)]
    for st in range(2, 7):
        t.append(("Mconv1_stage%d" % st, 128 + n_out, 128, 7))
        for i in (2, 3, 4, 5):
            t.append(("Mconv%d_stage%d" % (i, st), 128, 128, 7))
        t.append(("Mconv6_stage%d" % st, 128, 128, 1))
        t.append(("Mconv7_stage%d" % st, 128, n_out, 1))
    return t


def keypoint_forward(weights, x):
    """models/FaceNet.py:78-161 (HandNet identical) with torch CPU fp32: x [N,3,368,368] -> last-stage maps
    [N,n_out,46,46].  `weights`: {layer: (W, b)}."""
    import torch
    import torch.nn.functional as F

    def conv(name, h, relu=True):
        W, b = weights[name]
        y = F.conv2d(h, torch.from_numpy(W), torch.from_numpy(b), stride=1, padding=(W.shape[2] - 1) // 2)
        return torch.relu(y) if relu else y

    with torch.no_grad():
        h = torch.from_numpy(np.ascontiguousarray(x, np.float32))
        h = conv("conv1_2", conv("conv1_1", h))
        h = F.max_pool2d(h, 2, 2, ceil_mode=True)
        h = conv("conv2_2", conv("conv2_1", h))
        h = F.max_pool2d(h, 2, 2, ceil_mode=True)
        for n in ("conv3_1", "conv3_2", "conv3_3", "conv3_4"):
            h = conv(n, h)
        h = F.max_pool2d(h, 2, 2, ceil_mode=True)
        for n in ("conv4_1", "conv4_2", "conv4_3", "conv4_4", "conv5_1", "conv5_2", "conv5_3_CPM"):
            h = conv(n, h)
        feat = h
        h = conv("conv6_2_CPM", conv("conv6_1_CPM", feat), relu=False)
        for st in range(2, 7):
            h = torch.cat((h, feat), dim=1)                      # order (h, feature_map), models/FaceNet.py:108
            for i in (1, 2, 3, 4, 5, 6):
                h = conv("Mconv%d_stage%d" % (i, st), h)
            h = conv("Mconv7_stage%d" % st, h, relu=False)
    return h.numpy()


def keypoint_preprocess(img):
    """face_detector.py:32: float32, /256 (not /255), -0.5, HWC -> NCHW."""
    return np.array(img[np.newaxis], dtype=np.float32).transpose(0, 3, 1, 2) / 256 - 0.5


def keypoints_from_heatmaps(heatmaps, thresh=KEYPOINT_PEAK_THRESH):
    """face_detector.py:55-67 / hand_detector.py:65-77 (CPU branch): per channel except the last (background) the
    smoothed map's maximum; above `thresh` (compared in float32) -> [x, y, conf] else None.  With k >= 2 exact
    ties the reference's flatten() of np.where yields [y0..yk-1, x0..xk-1], so it reports (x, y) = (y1, y0)."""
    out = []
    sm = gaussian_smooth(np.ascontiguousarray(heatmaps[:-1], np.float32))
    for m in sm:
        mx = m.max()
        if mx > np.float32(thresh):
            coords = np.array(np.where(m == mx)).flatten().tolist()
            out.append([coords[1], coords[0], mx])
        else:
            out.append(None)
    return out


def detect_keypoints(weights, img, hand_type=None):
    """FaceDetector.__call__ (face_detector.py:28-41) when hand_type is None, else HandDetector.__call__
    (hand_detector.py:28-51; 'left' mirrors the crop before and the maps after the network)."""
    import cv2
    if hand_type == "left":
        img = cv2.flip(img, 1)
    h, w = img.shape[:2]
    resized = cv2_resize_linear_u8(np.ascontiguousarray(img), (KEYPOINT_IMG_SIZE, KEYPOINT_IMG_SIZE))
    lo = keypoint_forward(weights, keypoint_preprocess(resized))
    maps = resize_bilinear_align_corners(lo, (h, w))[0]
    if hand_type == "left":
        maps = np.ascontiguousarray(maps[:, :, ::-1])
    return keypoints_from_heatmaps(maps), lo[0], maps

