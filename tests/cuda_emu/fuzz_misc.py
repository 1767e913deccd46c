"""TEST INFRASTRUCTURE ONLY: random-shape fuzzing of the remaining bit-exact entry points under emulation (no GPU):
opb_resize_linear_u8 vs cv2.resize, opb_upsample vs the oracle, opb_peaks vs the oracle (images down to 1x1),
opb_keypoints_from_heatmaps vs the oracle (mirrored or not).
    python tests/cuda_emu/fuzz_misc.py [seed]"""
import sys, os, time, ctypes as C, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
import numpy as np, cv2, build_emu
PKG = "chainer_realtime_multi-person_pose_estimation_b200"
native = importlib.import_module(PKG + "._native")
lib = C.CDLL(build_emu.build(contract=False))
for name, (res, args) in native._SIGNATURES.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
native._lib = lib
from oracle import restate as R
PD = importlib.import_module(PKG + ".pose_detector")
eng = native.Engine(0, PD.make_opb_params(max_peaks=16384, max_candidates=131072, max_persons=512))
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for i in range(120):   # uint8 resize vs cv2
    h0, w0, h, w = (int(v) for v in rs.randint(1, 90, 4))
    n = int(rs.choice([1, 1, 3]))
    img = rs.randint(0, 256, (n, h0, w0, 3)).astype(np.uint8)
    try:
        got = eng.resize_linear_u8(img, h, w)
    except native.OpbError as e:
        print("resize", (h0, w0, h, w), "refused", str(e)[:60]); continue
    for k in range(n):
        if not np.array_equal(got[k], cv2.resize(img[k], (w, h))):
            bad += 1; print("resize FAIL", (h0, w0), "->", (h, w)); break
for i in range(80):    # bilinear align-corners upsample vs oracle
    h, w = (int(v) for v in rs.randint(2, 30, 2)); H, W = (int(v) for v in rs.randint(1, 120, 2)); p = int(rs.randint(1, 25))
    x = (rs.standard_normal((p, h, w)) * rs.uniform(0.1, 3)).astype(np.float32)
    try:
        got = eng.upsample(x, H, W)
    except native.OpbError as e:
        print("upsample", (h, w, H, W), "refused", str(e)[:60]); continue
    if not np.array_equal(got, R.resize_bilinear_align_corners(x[None], (H, W))[0]):
        bad += 1; print("upsample FAIL", (p, h, w), "->", (H, W))
for i in range(60):    # peaks on awkward sizes
    h, w = (int(v) for v in rs.randint(1, 70, 2))
    heat = (rs.standard_normal((19, h, w)) * rs.choice([0.05, 0.3, 1.0])).astype(np.float32)
    try:
        got = eng.peaks(heat)
    except native.OpbError as e:
        print("peaks", (h, w), "refused", str(e)[:60]); continue
    ref = R.compute_peaks_from_heatmaps(heat).reshape(-1, 5)
    if got.shape != ref.shape or not np.array_equal(got, ref):
        bad += 1; print("peaks FAIL", (h, w), got.shape, ref.shape)
for i in range(40):    # keypoint arg-max
    c, h, w = int(rs.randint(1, 8)), int(rs.randint(1, 50)), int(rs.randint(1, 50))
    hm = (rs.standard_normal((c, h, w)) * 0.3).astype(np.float32)
    mirror = bool(rs.randint(0, 2))
    try:
        got = eng.keypoints_from_heatmaps(hm, 0.1, mirror=mirror)
    except native.OpbError as e:
        print("keypoints", (c, h, w), "refused", str(e)[:60]); continue
    full = np.concatenate([hm, np.zeros((1, h, w), np.float32)])
    ref = R.keypoints_from_heatmaps(np.ascontiguousarray(full[:, :, ::-1]) if mirror else full, 0.1)
    ok = all((a is None) == (b is None) and (a is None or ((a[0], a[1]) == (b[0], b[1]) and np.float32(a[2]) == np.float32(b[2]))) for a, b in zip(got, ref))
    if not ok:
        bad += 1; print("keypoints FAIL", (c, h, w), mirror)
print("misc fuzz done, failures:", bad)
