#!/bin/bash
# scratch experiment: per-launch profile of ONE 640x480 frame (fast precision)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OPB_PROFILE=1 python - > gpurun_out/exp.log 2>&1 <<'PY'
import importlib, time, numpy as np, sys
sys.path.insert(0, '.')
P = "chainer_realtime_multi-person_pose_estimation_b200"
syn = importlib.import_module(P + ".synthetic")
m = importlib.import_module(P + ".models.CocoPoseNet").CocoPoseNet(); m.load_npz(syn.he_weights(0))
det = importlib.import_module(P + ".pose_detector").PoseDetector(model=m, device=0, precision="fast")
frame = syn.procedural_image(480, 640, seed=2)
import cv2
resized = cv2.resize(frame, (496, 368))
for i in range(4):
    t0 = time.perf_counter(); det.engine.detect_batch(resized[None], 320, 432); print("wall ms", 1e3 * (time.perf_counter() - t0))
PY
tail -n 45 gpurun_out/exp.log
