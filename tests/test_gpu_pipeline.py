"""`-m gpu`: end-to-end parity of the device pipeline against the committed goldens (outputs of
the reference's own files run verbatim, oracle/make_goldens.py) on identical seeded weights.

Float maps: |device - oracle| <= 1e-3 (north_star tolerance; parity mode is expected ~3e-5).
Keypoints: the peak / person sets must be identical.  Because the conv output differs from
the fp32 oracle by ~1e-5, a peak whose decision margin in the oracle is itself below that
noise can legitimately flip; such near-ties are identified FROM THE ORACLE MAPS (margin <
TIE_EPS) and excluded from the identity check -- the count of excluded peaks is asserted to
be tiny.  Kernel-level bit-exactness on identical inputs is in test_gpu_postprocess.py."""
import numpy as np
import pytest

from conftest import load_golden, pkg, split_conns
from oracle import restate as R

pytestmark = pytest.mark.gpu

MAP_TOL = 1e-3
TIE_FACTOR = 4.0      # a peak may flip only if its oracle margin is below TIE_FACTOR x the measured map error


@pytest.fixture(scope="module")
def weights_model():
    m = pkg("models.CocoPoseNet").CocoPoseNet()
    m.load_npz(pkg("synthetic").he_weights(0))
    return m


@pytest.fixture(scope="module")
def det_parity(weights_model):
    return pkg("pose_detector").PoseDetector(model=weights_model, device=0, precision="parity",
                                             max_candidates=131072, max_persons=4096)


def _oracle_margins(heat):
    """Per-pixel decision margin of the oracle's peak test on the smoothed maps."""
    g = R.gaussian_smooth(heat[:-1].astype(np.float32)).astype(np.float64)
    pad = np.pad(g, ((0, 0), (1, 1), (1, 1)))
    nb = np.maximum.reduce([pad[:, :-2, 1:-1], pad[:, 2:, 1:-1], pad[:, 1:-1, :-2], pad[:, 1:-1, 2:]])
    return np.minimum(g - R.HEATMAP_PEAK_THRESH, g - nb)


def _peak_sets_match(got_peaks, ref_peaks, heat_oracle, tie_eps):
    margin = _oracle_margins(heat_oracle)
    key = lambda p: set(map(tuple, p[:, :3].astype(int))) if len(p) else set()
    G, Rf = key(got_peaks), key(ref_peaks)
    sym = G ^ Rf
    for (c, x, y) in sym:
        assert abs(margin[c, y, x]) < tie_eps, "peak (%d,%d,%d) differs with oracle margin %.3e (eps %.1e)" % (
            c, x, y, margin[c, y, x], tie_eps)
    return len(sym)


def test_forward_maps_parity_mode(det_parity):
    g = load_golden("fast_584_he0.npz")
    img = pkg("synthetic").procedural_image(584, 584, seed=1)
    import cv2
    x = det_parity.preprocess(cv2.resize(img, (368, 368)))
    paf, heat = det_parity.engine.forward(x)
    e1, e2 = np.abs(paf[0] - g["paf_lo_0"]).max(), np.abs(heat[0] - g["heat_lo_0"]).max()
    print("parity-mode max abs err: paf %.3e heat %.3e" % (e1, e2))
    assert e1 <= MAP_TOL and e2 <= MAP_TOL
    # uint8 entry (preprocess fused into conv1_1) gives the same maps
    paf_u8, heat_u8 = det_parity.engine.forward(cv2.resize(img, (368, 368))[None])
    assert np.abs(paf_u8 - paf).max() <= 1e-5 and np.abs(heat_u8 - heat).max() <= 1e-5


def test_fast_mode_uint8_entry_matches_float_entry(weights_model):
    """conv1_1 on tensor cores (uint8 frames, im2col built per thread) vs the CUDA-core conv1_1 fed the preprocessed
    float32 image: same network, so the maps may differ only by fp16 rounding of the first layer's operands (which the
    other 91 layers amplify to ~1e-2 on a white-noise frame whose maps are O(10); a misplaced tap would give O(1)).
    The bound is relative to the map magnitude."""
    det = pkg("pose_detector").PoseDetector(model=weights_model, device=0, precision="fast")
    imgs = pkg("synthetic").random_images(2, 368, 496, seed=4)
    imgs[1, :, :, :] = pkg("synthetic").procedural_image(368, 496, seed=9)
    paf_u8, heat_u8 = det.engine.forward(imgs)
    x = np.concatenate([det.preprocess(im) for im in imgs])
    paf_f, heat_f = det.engine.forward(x)
    for i in range(2):
        e = max(float(np.abs(paf_u8[i] - paf_f[i]).max()), float(np.abs(heat_u8[i] - heat_f[i]).max()))
        mag = max(float(np.abs(paf_f[i]).max()), float(np.abs(heat_f[i]).max()))
        print("fast mode, uint8 (tensor-core conv1_1) vs float32 entry, image %d: max abs diff %.3e (maps up to %.2f)"
              % (i, e, mag))
        assert e < 4e-3 * max(mag, 1.0)


def test_fused_1x1_pair_is_bit_identical_to_two_launches(weights_model, monkeypatch):
    """Mconv6 + Mconv7 fused into one kernel (csrc/conv_mlp2.cuh, the 128-channel intermediate stays in shared memory)
    vs the two separate 1x1 launches: same MMA shapes and accumulation order, so the maps must be bit-identical --
    batch 2 (throughput tile shapes), batch 1 (small-batch shapes) and HandNet (single branch)."""
    syn = pkg("synthetic")
    imgs = syn.random_images(2, 368, 496, seed=4)
    hn = pkg("models.HandNet")
    hand = hn.HandNet()
    hand.load_npz(syn.he_weights(0, layers=hn.LAYERS))
    crop = syn.procedural_image(368, 368, seed=5)[None]
    out = []
    for no_fuse in ("0", "1"):
        monkeypatch.setenv("OPB_NO_MLP2", no_fuse)
        det = pkg("pose_detector").PoseDetector(model=weights_model, device=0, precision="fast")
        hd = pkg("hand_detector").HandDetector(model=hand, device=0, precision="fast")
        out.append((det.engine.forward(imgs), det.engine.forward(imgs[:1]), hd.engine.forward_keypoint_maps(crop)))
        del det, hd
    (a2, a1, ah), (b2, b1, bh) = out
    for x, y in ((a2[0], b2[0]), (a2[1], b2[1]), (a1[0], b1[0]), (a1[1], b1[1]), (ah, bh)):
        assert np.isfinite(x).all() and np.array_equal(x, y), float(np.abs(x - y).max())


def test_forward_maps_fast_mode(weights_model):
    det = pkg("pose_detector").PoseDetector(model=weights_model, device=0, precision="fast")
    g = load_golden("fast_584_he0.npz")
    import cv2
    img = pkg("synthetic").procedural_image(584, 584, seed=1)
    paf, heat = det.engine.forward(cv2.resize(img, (368, 368))[None])
    e1, e2 = np.abs(paf[0] - g["paf_lo_0"]).max(), np.abs(heat[0] - g["heat_lo_0"]).max()
    print("fast-mode (fp16) max abs err: paf %.3e heat %.3e" % (e1, e2))
    assert e1 <= 5e-2 and e2 <= 5e-2   # reported, not a parity claim (~5e-3 expected)


def _check_call(det, name, img):
    g = load_golden(name)
    poses, scores = det(img)
    if g["all_peaks"].shape[0] == 0:
        assert poses.shape == (0, 18, 3) and scores.shape == (0,)
        return
    oh, ow = img.shape[:2]
    in_w, in_h = R.compute_optimal_size(img, 368)
    map_w, map_h = R.compute_optimal_size(img, 320)
    heat_or = R.resize_bilinear_align_corners(g["heat_lo_0"][None], (map_h, map_w))[0]
    peaks, conns, subsets = det.engine.image_detail(0)
    import cv2
    paf_lo, heat_lo = det.engine.forward(cv2.resize(img, (in_w, in_h))[None])
    map_err = max(np.abs(paf_lo[0] - g["paf_lo_0"]).max(), np.abs(heat_lo[0] - g["heat_lo_0"]).max())
    assert map_err <= MAP_TOL
    n_ties = _peak_sets_match(peaks, g["all_peaks"], heat_or, max(TIE_FACTOR * map_err, 1e-4))
    print(name, "map err %.2e" % map_err, "peaks", len(peaks), "ref", len(g["all_peaks"]), "near-tie flips", n_ties)
    assert n_ties <= max(2, len(g["all_peaks"]) // 500)
    if n_ties == 0:
        assert np.array_equal(peaks[:, :3], g["all_peaks"][:, :3])
        assert np.abs(peaks[:, 3] - g["all_peaks"][:, 3]).max() <= MAP_TOL
        ref_conns = split_conns(g["conn_lens"], g["conn_flat"])
        same_conn = all(a.shape == b.shape and np.array_equal(a[:, :2], b[:, :2]) for a, b in zip(conns, ref_conns))
        # connection acceptance thresholds (ip > 0.05, score > 0) have their own measure-zero ties
        if same_conn:
            assert subsets.shape == g["subsets"].shape
            assert np.array_equal(subsets[:, :18], g["subsets"][:, :18])
            assert np.abs(subsets[:, 18:] - g["subsets"][:, 18:]).max() <= 1e-2
            assert poses.shape == g["poses"].shape and np.array_equal(poses, g["poses"])      # OKS = 1.0
            assert np.abs(scores - g["scores"]).max() <= 1e-2
        else:
            n_diff = sum(0 if (a.shape == b.shape and np.array_equal(a[:, :2], b[:, :2])) else 1
                         for a, b in zip(conns, ref_conns))
            print("connection lists differ on", n_diff, "limbs (threshold near-ties)")
            assert n_diff <= 2


def test_call_fast_path_584(det_parity):
    _check_call(det_parity, "fast_584_he0.npz", pkg("synthetic").procedural_image(584, 584, seed=1))


def test_call_fast_path_webcam_shape(det_parity):
    _check_call(det_parity, "fast_480x640_he0.npz", pkg("synthetic").procedural_image(480, 640, seed=2))


def test_call_fast_path_dense_noise(det_parity):
    _check_call(det_parity, "fast_368x656_he0_img0.npz", pkg("synthetic").random_images(2, 368, 656, seed=0)[0])


def test_call_default_init_returns_empty():
    m = pkg("models.CocoPoseNet").CocoPoseNet()
    m.load_npz(pkg("synthetic").he_weights(0, bias_scale=0.0, gain=1.0))
    det = pkg("pose_detector").PoseDetector(model=m, device=0)
    poses, scores = det(pkg("synthetic").procedural_image(584, 584, seed=1))
    assert poses.shape == (0, 18, 3) and scores.shape == (0,)


def test_detect_batch_matches_single_calls(det_parity):
    imgs = pkg("synthetic").random_images(2, 368, 656, seed=0)
    res = det_parity.detect_batch(imgs)
    for i in range(2):
        g = load_golden("fast_368x656_he0_img%d.npz" % i)
        poses, scores = res[i]
        single = det_parity(imgs[i])
        assert poses.shape == single[0].shape and np.array_equal(poses, single[0])
        assert abs(len(poses) - len(g["poses"])) <= 3


def test_stream_mode_matches_single_calls(det_parity):
    """Pipelined submit/collect (two slots, copy stream, device resize) returns, in order, exactly what the
    synchronous __call__ returns for every frame -- frames of different sizes, single frames and batches."""
    syn = pkg("synthetic")
    frames = [syn.procedural_image(480, 640, seed=3), syn.procedural_image(584, 584, seed=1),
              syn.random_images(2, 368, 656, seed=0), syn.procedural_image(240, 320, seed=7),
              syn.procedural_image(480, 640, seed=4)]
    got = list(det_parity.detect_stream(iter(frames)))
    assert len(got) == len(frames)
    for f, g in zip(frames, got):
        if f.ndim == 4:
            for i in range(len(f)):
                p, sc = det_parity(f[i])
                assert g[i][0].shape == p.shape and np.array_equal(g[i][0], p) and np.array_equal(g[i][1], sc)
        else:
            p, sc = det_parity(f)
            assert g[0].shape == p.shape and np.array_equal(g[0], p) and np.array_equal(g[1], sc)
    # same shape submitted repeatedly: from the second submit per slot the launch sequence is captured into a CUDA
    # graph and replayed -- results must not change (8 frames = 4 per slot: eager, capture, replay, replay)
    same = [syn.procedural_image(480, 640, seed=30 + i) for i in range(8)]
    got = list(det_parity.detect_stream(iter(same)))
    for f, g in zip(same, got):
        p, sc = det_parity(f)
        assert g[0].shape == p.shape and np.array_equal(g[0], p) and np.array_equal(g[1], sc)
    # protocol errors are reported, not silently overwritten
    eng = det_parity.engine
    eng.stream_submit(frames[0], 368, 496, 320, 432, slot=0)
    with pytest.raises(RuntimeError):
        eng.stream_submit(frames[0], 368, 496, 320, 432, slot=0)
    eng.stream_collect(0)
    with pytest.raises(RuntimeError):
        eng.stream_collect(0)


def test_injected_synthetic_eight_person_maps(det_parity):
    """Config #3: synthetic 8-person maps injected as the network output (the conv chain still
    runs); 8 persons must come out for every image of the batch, identical to the oracle run
    on the same low-resolution maps."""
    import torch
    syn = pkg("synthetic")
    paf_lo, heat_lo = syn.eight_person_lowres(46, 82, seed=0)
    imgs = syn.random_images(2, 368, 656, seed=5)
    n = len(imgs)
    d_paf = torch.from_numpy(np.repeat(paf_lo[None], n, 0)).cuda()
    d_heat = torch.from_numpy(np.repeat(heat_lo[None], n, 0)).cuda()
    torch.cuda.synchronize()
    headers, persons = det_parity.engine.detect_batch(imgs, 320, 576, inject_paf=d_paf.data_ptr(),
                                                      inject_heat=d_heat.data_ptr())
    pafs = R.resize_bilinear_align_corners(paf_lo[None], (320, 576))[0]
    heat = R.resize_bilinear_align_corners(heat_lo[None], (320, 576))[0]
    ref_poses, ref_scores, parts = R.postprocess_fast(pafs, heat, 576, 576, 320, 320, return_parts=True)
    for i in range(n):
        assert headers[i]["status"] == 0 and headers[i]["n_persons"] == len(ref_scores) == 8
        peaks, conns, subsets = det_parity.engine.image_detail(i)
        assert np.array_equal(peaks, parts["all_peaks"])
        assert np.array_equal(subsets, parts["subsets"])


def test_precise_path_480(weights_model):
    det = pkg("pose_detector").PoseDetector(model=weights_model, device=0, precise=True, precision="parity",
                                            max_candidates=131072, max_persons=4096)
    g = load_golden("precise_480_he0.npz")
    img = pkg("synthetic").procedural_image(480, 480, seed=3)
    poses, scores = det(img)
    e1 = np.abs(det.pafs[:, ::7, ::7] - g["pafs_sample"]).max()
    e2 = np.abs(det.heatmaps[:, ::7, ::7] - g["heatmaps_sample"]).max()
    print("precise maps max abs err: paf %.3e heat %.3e; peaks %d ref %d" % (e1, e2, len(det.all_peaks),
                                                                            len(g["all_peaks"])))
    assert e1 <= MAP_TOL and e2 <= MAP_TOL
    G = set(map(tuple, det.all_peaks[:, :3].astype(int)))
    Rf = set(map(tuple, g["all_peaks"][:, :3].astype(int)))
    print("precise peak-set symmetric difference:", len(G ^ Rf))
    assert len(G ^ Rf) <= max(2, len(Rf) // 500)
    if G == Rf and poses.shape == g["poses"].shape:
        assert np.array_equal(poses, g["poses"])       # OKS = 1.0 vs the reference


def test_precise_path_padded_200x300(weights_model):
    """Precise path with padding: the scaled inputs 184x276 and 552x828 are padded to 280 / 832 columns
    (pad_image, :445) and the x8 maps cropped again (:462,:466)."""
    det = pkg("pose_detector").PoseDetector(model=weights_model, device=0, precise=True, precision="parity",
                                            max_candidates=131072, max_persons=4096)
    g = load_golden("precise_200x300_he0.npz")
    img = pkg("synthetic").procedural_image(200, 300, seed=4)
    poses, scores = det(img)
    e1 = np.abs(det.pafs[:, ::5, ::5] - g["pafs_sample"]).max()
    e2 = np.abs(det.heatmaps[:, ::5, ::5] - g["heatmaps_sample"]).max()
    G = set(map(tuple, det.all_peaks[:, :3].astype(int)))
    Rf = set(map(tuple, g["all_peaks"][:, :3].astype(int)))
    print("precise 200x300 max abs err: paf %.3e heat %.3e; peaks %d ref %d symdiff %d" % (
        e1, e2, len(G), len(Rf), len(G ^ Rf)))
    assert e1 <= MAP_TOL and e2 <= MAP_TOL
    assert len(G ^ Rf) <= max(2, len(Rf) // 500)
    if G == Rf and poses.shape == g["poses"].shape:
        assert np.array_equal(poses, g["poses"])
