#!/bin/bash
# scratch: single-frame latency (fast + parity), with and without the small-batch tile heuristic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for SB in 0 1; do
OPB_SMALL_BATCH=$SB python - <<'PY'
import importlib, time, numpy as np, sys, os
sys.path.insert(0, '.')
P = "chainer_realtime_multi-person_pose_estimation_b200"
syn = importlib.import_module(P + ".synthetic")
m = importlib.import_module(P + ".models.CocoPoseNet").CocoPoseNet(); m.load_npz(syn.he_weights(0))
frame = syn.procedural_image(480, 640, seed=2)
for prec in ("fast", "parity"):
    det = importlib.import_module(P + ".pose_detector").PoseDetector(model=m, device=0, precision=prec)
    for _ in range(3): det(frame)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); det(frame); ts.append(time.perf_counter() - t0)
    x = det.preprocess(np.zeros((368, 496, 3), np.uint8))
    for _ in range(3): det.engine.forward(x)
    t0 = time.perf_counter()
    for _ in range(10): det.engine.forward(x)
    tf = (time.perf_counter() - t0) / 10
    print("SMALL_BATCH=%s %s: __call__ median %.3f ms, forward-only (incl. H2D/D2H) %.3f ms" % (os.environ["OPB_SMALL_BATCH"], prec, 1e3 * np.median(ts), 1e3 * tf))
    del det
PY
done
