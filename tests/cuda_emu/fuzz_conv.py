"""TEST INFRASTRUCTURE ONLY: random-shape fuzzing of the tcgen05 conv kernels under emulation (no GPU).
    python tests/cuda_emu/fuzz_conv.py [seed] [n_cases]      (OPB_EMU_SMS=2..4 makes persistent CTAs wrap their pipelines)
Every case (kernel size, channels, image size, batch, ReLU, fused max-pool, precision drawn at random) goes through
opb_test_conv of the emulated library and is compared with a torch fp32 conv2d under the tolerances of
tests/test_gpu_conv.py.  270 cases over two seeds passed when this was written."""
import sys, os, time, ctypes as C, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
import build_emu
import numpy as np
os.environ["OPB_EMU_SMS"] = os.environ.get("OPB_EMU_SMS", "3")
PKG = "chainer_realtime_multi-person_pose_estimation_b200"
native = importlib.import_module(PKG + "._native")
lib = C.CDLL(build_emu.build(contract=False))
for name, (res, args) in native._SIGNATURES.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
native._lib = lib
import test_gpu_conv as T
eng = native.Engine(0, importlib.import_module(PKG + ".pose_detector").make_opb_params())
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = 0
t0 = time.time()
for i in range(ncase):
    ks = int(rs.choice([1, 3, 7]))
    cin = int(rs.choice([64, 128, 185, 192, 256, 512, 150, 199]))
    cout = int(rs.choice([19, 22, 38, 64, 71, 128, 256, 512]))
    n = int(rs.choice([1, 1, 2, 3]))
    h = int(rs.randint(16 + ks - 1, 52)); w = int(rs.randint(8, 70))
    relu = int(rs.randint(0, 2)); pool = bool(rs.randint(0, 4) == 0) and ks == 3 and h % 2 == 0 and w % 2 == 0
    mode = "fast" if rs.randint(0, 3) else "parity"
    if cin * ks * ks * cout * h * w * n > 3e10: continue
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    tag = "case %d: n%d %dx%d cin%d cout%d k%d relu%d pool%d %s" % (i, n, h, w, cin, cout, ks, relu, pool, mode)
    try:
        y = eng.test_conv(x, W, b, relu, native.PRECISION_FAST if mode == "fast" else native.PRECISION_PARITY, pool=pool)
    except native.OpbError as e:
        print(tag, "-> refused:", str(e)[:90]); continue
    ref = T._ref_conv(x, W, b, relu, quantize=(mode == "fast"))
    if pool:
        import torch
        ref = torch.nn.functional.max_pool2d(torch.from_numpy(ref).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).numpy()
    scale = np.abs(ref).max(); err = np.abs(y - ref).max()
    tol = (2e-3 if mode == "fast" else 1e-4) * scale
    if not (err <= tol):
        bad += 1
        print(tag, "FAIL err %.3e tol %.3e" % (err, tol), flush=True)
print("done %d cases, %d failures, %.0fs" % (ncase, bad, time.time() - t0))
