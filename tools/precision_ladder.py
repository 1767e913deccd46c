#!/usr/bin/env python
"""Precision ladder of the conv chain on a B200: for every precision mode, the map error against
the committed reference golden (fast_584_he0: reference files run verbatim, fp32), the peak-set symmetric difference on
that golden, and the conv-chain time / frames/s at the benchmark shape (batch 32, 368x656).

    python tools/precision_ladder.py [--batch 32] > profiles/rNN_precision_ladder.txt
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "chainer_realtime_multi-person_pose_estimation_b200"


def pkg(sub=None):
    return importlib.import_module(PKG + ("." + sub if sub else ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--modes", default="fast,comp,parity")
    args = ap.parse_args()
    import cv2
    native, syn = pkg("_native"), pkg("synthetic")
    model = pkg("models.CocoPoseNet").CocoPoseNet()
    model.load_npz(syn.he_weights(0))
    goldens = [("fast_584_he0.npz", syn.procedural_image(584, 584, seed=1)),
               ("fast_480x640_he0.npz", syn.procedural_image(480, 640, seed=2)),
               ("fast_368x656_he0_img0.npz", syn.random_images(2, 368, 656, seed=0)[0])]
    from oracle import restate as R   # (test infrastructure: the checker, not the thing measured)
    print("# precision ladder, batch %d x 368x656 (conv chain = 15.508 TFLOP useful per batch)" % args.batch)
    print("# mode | max abs map err vs reference golden (3 goldens) | peak-set symmetric difference / reference peaks | conv chain ms | frames/s (conv chain only)")
    for mode in args.modes.split(","):
        det = pkg("pose_detector").PoseDetector(model=model, device=0, precision=mode, max_candidates=131072, max_persons=4096)
        errs, syms, fits = [], [], []
        for name, img in goldens:
            g = np.load(os.path.join(ROOT, "tests", "golden", name))
            in_w, in_h = R.compute_optimal_size(img, 368)
            paf, heat = det.engine.forward(cv2.resize(img, (in_w, in_h))[None])
            errs.append(max(float(np.abs(paf[0] - g["paf_lo_0"]).max()), float(np.abs(heat[0] - g["heat_lo_0"]).max())))
            # systematic part of the error: least-squares scale s of device = s * reference, and what is left without it
            d = np.concatenate([paf[0].ravel(), heat[0].ravel()]).astype(np.float64)
            r = np.concatenate([g["paf_lo_0"].ravel(), g["heat_lo_0"].ravel()]).astype(np.float64)
            sfit = float(d @ r / (r @ r))
            fits.append("s-1=%+.2e resid=%.2e max|ref|=%.2f" % (sfit - 1.0, float(np.abs(d - sfit * r).max()), float(np.abs(r).max())))
            det(img)
            peaks = det.engine.image_detail(0)[0]
            key = lambda p: set(map(tuple, p[:, :3].astype(int))) if len(p) else set()
            syms.append("%d/%d" % (len(key(peaks) ^ key(g["all_peaks"])), len(g["all_peaks"])))
        imgs = syn.random_images(args.batch, 368, 656, seed=0)
        det.engine.forward(imgs)
        ms = det.engine.time_stage("conv_chain", reps=5)
        print("%-7s | %s | %s | %8.3f | %8.1f" % (mode, " ".join("%.2e" % e for e in errs), " ".join(syms), ms, args.batch / ms * 1e3), flush=True)
        print("#        scale fit: " + " ; ".join(fits), flush=True)
        del det


if __name__ == "__main__":
    main()
