"""`-m gpu`: FaceNet / HandNet path against the committed goldens -- outputs of the reference's
own face_detector.py / hand_detector.py run verbatim (oracle/make_goldens.py keypoint_goldens) on seeded weights.

(i) kernel level, bit-exact: the device smoothing + per-channel maximum fed with the ORACLE's upsampled maps must
    reproduce x, y and the float32 confidence exactly (incl. the mirrored left-hand case and exact ties).
(ii) end to end (device resize + conv chain + upsample + smooth + argmax): last-stage maps within 1e-3 of the oracle
    (parity mode ~2e-5); keypoints identical except channels whose decision is a near-tie IN THE ORACLE (threshold
    margin or runner-up margin below TIE_FACTOR x the measured map error)."""
import numpy as np
import pytest

from conftest import load_golden, pkg
from oracle import restate as R

pytestmark = pytest.mark.gpu

MAP_TOL = 1e-3
TIE_FACTOR = 4.0
CASES = [("face_150x170_he0.npz", "face", None), ("face_401x401_he0.npz", "face", None),
         ("hand_120x131_he0_right.npz", "hand", "right"), ("hand_120x131_he0_left.npz", "hand", "left")]


def _model(kind):
    mod = pkg("models.FaceNet" if kind == "face" else "models.HandNet")
    m = getattr(mod, "FaceNet" if kind == "face" else "HandNet")()
    m.load_npz(pkg("synthetic").he_weights(0, layers=mod.LAYERS))
    return m


@pytest.fixture(scope="module")
def detectors():
    return {"face": pkg("face_detector").FaceDetector(model=_model("face"), device=0, precision="parity"),
            "hand": pkg("hand_detector").HandDetector(model=_model("hand"), device=0, precision="parity")}


@pytest.fixture(scope="module")
def detectors_comp():
    """the default ("comp": fp16 + 8-bit-float rounding corrections) precision of the drop-in detectors"""
    return {"face": pkg("face_detector").FaceDetector(model=_model("face"), device=0),
            "hand": pkg("hand_detector").HandDetector(model=_model("hand"), device=0)}


def _golden_list(g):
    return [[int(x), int(y), c] if v else None for (x, y), c, v in zip(g["xy"], g["conf"], g["valid"])]


def _oracle_maps(g, hand_type):
    h, w, _ = (int(v) for v in g["img_hw_seed"])
    maps = R.resize_bilinear_align_corners(g["heat_lo"][None], (h, w))[0]
    return np.ascontiguousarray(maps[:, :, ::-1]) if hand_type == "left" else maps


@pytest.mark.parametrize("name,kind,hand_type", CASES)
def test_keypoints_from_oracle_maps_bit_exact(detectors, name, kind, hand_type):
    g = load_golden(name)
    maps = _oracle_maps(g, hand_type)                 # what compute_peaks_from_heatmaps receives in the reference
    got = detectors[kind].compute_peaks_from_heatmaps(maps)
    ref = _golden_list(g)
    assert len(got) == len(ref) == maps.shape[0] - 1
    for a, b in zip(got, ref):
        assert (a is None) == (b is None)
        if a is not None:
            assert a[0] == b[0] and a[1] == b[1] and np.float32(a[2]) == np.float32(b[2])
    if hand_type == "left":                           # mirror flag == flipping the maps first
        unflipped = np.ascontiguousarray(maps[:, :, ::-1])
        thr = pkg("entity").params["hand_heatmap_peak_thresh"]
        again = detectors[kind].engine.keypoints_from_heatmaps(unflipped[:-1], thr, mirror=True)
        assert [None if k is None else (k[0], k[1], float(k[2])) for k in again] == \
               [None if k is None else (k[0], k[1], float(k[2])) for k in got]


def test_keypoints_exact_ties_and_threshold(detectors):
    """np.where(g == max) with k >= 2 ties: the reference reads (x, y) = (y1, y0); strict `>` at the threshold."""
    eng = detectors["hand"].engine
    rs = np.random.RandomState(0)
    maps = np.zeros((6, 40, 52), np.float32)
    maps[0] = 0.5                                      # plateau: every pixel ties
    maps[1, 10:30, 8:44] = 0.4; maps[1, :, :] += 0    # symmetric block: centre ties after smoothing
    maps[2] = rs.uniform(0, 1, (40, 52)).astype(np.float32)
    maps[3] = 0.0999                                   # constant below the threshold
    maps[4] = np.float32(0.1)                          # exactly the float32 threshold: not strictly greater
    maps[5, 3, 50] = 9.0; maps[5, 36, 1] = 9.0         # two mirrored impulses
    full = np.concatenate([maps, np.zeros((1, 40, 52), np.float32)])
    ref = R.keypoints_from_heatmaps(full)
    got = eng.keypoints_from_heatmaps(maps, 0.1)
    for a, b in zip(got, ref):
        assert (a is None) == (b is None), (a, b)
        if a is not None:
            assert (a[0], a[1]) == (b[0], b[1]) and np.float32(a[2]) == np.float32(b[2]), (a, b)
    ref_m = R.keypoints_from_heatmaps(np.ascontiguousarray(full[:, :, ::-1]))
    got_m = eng.keypoints_from_heatmaps(maps, 0.1, mirror=True)
    for a, b in zip(got_m, ref_m):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a[0], a[1]) == (b[0], b[1]) and np.float32(a[2]) == np.float32(b[2]), (a, b)


@pytest.mark.parametrize("precision", ["parity", "comp"])
@pytest.mark.parametrize("name,kind,hand_type", CASES)
def test_detector_end_to_end_parity(detectors, detectors_comp, name, kind, hand_type, precision):
    import cv2
    g = load_golden(name)
    h, w, seed = (int(v) for v in g["img_hw_seed"])
    img = pkg("synthetic").procedural_image(h, w, seed=seed)
    det = (detectors if precision == "parity" else detectors_comp)[kind]
    # last-stage maps through the model object (the reference's `self.model(x_data)`), same preprocessing
    src = cv2.flip(img, 1) if hand_type == "left" else img
    x = R.keypoint_preprocess(cv2.resize(src, (368, 368)))
    lo = det.model(x)[-1][0]
    err = float(np.abs(lo - g["heat_lo"]).max())
    print("%s %s-precision max abs map err %.3e" % (name, precision, err))
    assert err <= MAP_TOL
    kw = {"hand_type": hand_type} if hand_type else {}
    got = det(img, **kw)
    ref = _golden_list(g)
    sm = R.gaussian_smooth(np.ascontiguousarray(_oracle_maps(g, hand_type)[:-1], np.float32))
    eps = TIE_FACTOR * max(err, 1e-6)
    flips = 0
    for c, (a, b) in enumerate(zip(got, ref)):
        mx = float(sm[c].max())
        if (a is None) != (b is None):
            assert abs(mx - np.float32(0.1)) < eps, "channel %d validity differs with margin %.3e" % (c, mx - 0.1)
            flips += 1
            continue
        if a is None:
            continue
        assert abs(float(a[2]) - float(b[2])) <= MAP_TOL
        if (a[0], a[1]) != (b[0], b[1]):
            assert mx - float(sm[c][a[1], a[0]]) < eps, "channel %d argmax differs beyond the near-tie margin" % c
            flips += 1
    print("%s near-tie flips %d of %d" % (name, flips, len(ref)))
    assert flips <= (3 if precision == "parity" else 6)


def test_fast_mode_runs_and_reports_error():
    g = load_golden("face_150x170_he0.npz")
    h, w, seed = (int(v) for v in g["img_hw_seed"])
    det = pkg("face_detector").FaceDetector(model=_model("face"), device=0, precision="fast")
    import cv2
    img = pkg("synthetic").procedural_image(h, w, seed=seed)
    lo = det.model(R.keypoint_preprocess(cv2.resize(img, (368, 368))))[-1][0]
    err = float(np.abs(lo - g["heat_lo"]).max())
    print("fast-mode (fp16) max abs map err %.3e" % err)
    assert err < 0.05
    got = det(img)
    same = sum((a is None) == (b is None) and (a is None or (a[0], a[1]) == (int(x), int(y)))
               for a, b, (x, y) in zip(got, g["valid"].tolist(), g["xy"]) if True)
    print("fast-mode keypoints identical to the oracle: %d of %d" % (same, len(got)))


def test_pose_api_rejected_on_keypoint_context(detectors):
    eng = detectors["face"].engine
    with pytest.raises(RuntimeError):
        eng.detect_batch(np.zeros((1, 368, 368, 3), np.uint8), 320, 320)
