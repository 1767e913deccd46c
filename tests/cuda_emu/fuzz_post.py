"""TEST INFRASTRUCTURE ONLY: random-shape fuzzing of opb_postprocess_batch under emulation (no GPU), cycling through the
OPB_FUSED_PEAKS / OPB_PAF_LOWRES variants; every stage must reproduce the oracle bit for bit.
    python tests/cuda_emu/fuzz_post.py [seed] [n_cases]"""
import sys, os, time, ctypes as C, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
import build_emu
import numpy as np
PKG = "chainer_realtime_multi-person_pose_estimation_b200"
native = importlib.import_module(PKG + "._native")
lib = C.CDLL(build_emu.build(contract=False))
for name, (res, args) in native._SIGNATURES.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
native._lib = lib
from postprocess_batch_cases import check_batch_against_oracle
PD = importlib.import_module(PKG + ".pose_detector")
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 30
t0 = time.time(); bad = 0
for i in range(ncase):
    knobs = [(0, 0), (1, 1), (2, 1), (1, 0), (0, 1)][i % 5]
    os.environ["OPB_FUSED_PEAKS"], os.environ["OPB_PAF_LOWRES"] = str(knobs[0]), str(knobs[1])
    eng = native.Engine(0, PD.make_opb_params(max_peaks=8192, max_candidates=262144, max_persons=512))
    n = int(rs.choice([1, 1, 2])); h8 = int(rs.randint(2, 24)); w8 = int(rs.randint(2, 30))
    H = int(rs.randint(max(h8, 4), 8 * h8 + 1)); W = int(rs.randint(max(w8, 4), 8 * w8 + 1))
    amp_h = float(rs.choice([0.05, 0.12, 0.3])); amp_p = float(rs.choice([0.2, 0.6]))
    paf = (rs.standard_normal((n, 38, h8, w8)) * amp_p).astype(np.float32)
    heat = (rs.standard_normal((n, 19, h8, w8)) * amp_h).astype(np.float32)
    tag = "case %d knobs %s n%d %dx%d -> %dx%d amp %.2f" % (i, knobs, n, h8, w8, H, W, amp_h)
    try:
        hd = check_batch_against_oracle(eng, paf, heat, H, W)
        print(tag, "ok peaks", list(hd["n_peaks"]), "persons", list(hd["n_persons"]), flush=True)
    except IndexError as e:
        print(tag, "IndexError (reference raises too?)", e)
    except AssertionError as e:
        bad += 1; print(tag, "FAIL", str(e)[:200], flush=True)
    except Exception as e:
        print(tag, type(e).__name__, str(e)[:120], flush=True)
print("done %d cases, %d failures, %.0fs" % (ncase, bad, time.time() - t0))
