#!/bin/bash
# scratch: streaming single-frame throughput with and without CUDA-graph replay
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for G in 0 1; do
OPB_GRAPH=$G python - <<'PY'
import importlib, time, numpy as np, sys, os
sys.path.insert(0, '.')
P = "chainer_realtime_multi-person_pose_estimation_b200"
syn = importlib.import_module(P + ".synthetic")
m = importlib.import_module(P + ".models.CocoPoseNet").CocoPoseNet(); m.load_npz(syn.he_weights(0))
frames = [syn.procedural_image(480, 640, seed=2 + (i % 4)) for i in range(204)]
for prec in ("fast",):
    det = importlib.import_module(P + ".pose_detector").PoseDetector(model=m, device=0, precision=prec)
    for _ in range(3): det(frames[0])
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); det(frames[0]); ts.append(time.perf_counter() - t0)
    it = det.detect_stream(iter(frames))
    for _ in range(4): next(it)
    t0 = time.perf_counter(); k = 0
    for _ in it: k += 1
    dt = (time.perf_counter() - t0) / k
    print("GRAPH=%s %s: __call__ median %.3f ms; detect_stream %.3f ms/frame (%.0f frames/s)" % (os.environ["OPB_GRAPH"], prec, 1e3 * np.median(ts), 1e3 * dt, 1 / dt))
PY
done
