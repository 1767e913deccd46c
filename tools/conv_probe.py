"""GPU bring-up diagnostics for the tcgen05 conv kernel (not a test; prints structured
evidence about layout / tap / padding mistakes).  Each case runs in its own subprocess so a
device trap in one case cannot poison the others.

    python tools/conv_probe.py            # run all cases
    python tools/conv_probe.py CASE_ID    # run one case in-process
"""
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "chainer_realtime_multi-person_pose_estimation_b200"

CASES = {
    # id: (n, h, w, cin, cout, ks, kind)
    "id1_onetile": (1, 16, 8, 64, 64, 1, "identity"),
    "id1_partial": (1, 22, 11, 64, 64, 1, "identity"),
    "id1_c128": (1, 16, 16, 128, 128, 1, "identity"),
    "d3_center": (1, 22, 16, 64, 64, 3, "delta:1:1"),
    "d3_00": (1, 22, 16, 64, 64, 3, "delta:0:0"),
    "d3_21": (1, 22, 16, 64, 64, 3, "delta:2:1"),
    "d7_center": (1, 24, 16, 64, 64, 7, "delta:3:3"),
    "d7_06": (1, 24, 16, 64, 64, 7, "delta:0:6"),
    "d7_52": (1, 24, 16, 64, 64, 7, "delta:5:2"),
    "r1": (1, 24, 24, 128, 128, 1, "random"),
    "r3": (2, 23, 31, 64, 64, 3, "random"),
    "r7": (1, 46, 82, 128, 128, 7, "random"),
    "r7_256": (1, 46, 46, 185, 256, 7, "random"),
    "r1_head": (1, 30, 17, 512, 38, 1, "random"),
    "r3_512": (1, 24, 24, 256, 512, 3, "random"),
}


def run_case(cid):
    import torch
    n, h, w, cin, cout, ks, kind = CASES[cid]
    native = importlib.import_module(PKG + "._native")
    pd = importlib.import_module(PKG + ".pose_detector")
    eng = native.Engine(0, pd.make_opb_params())
    rs = np.random.RandomState(1)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    x = x.astype(np.float16).astype(np.float32)
    W = np.zeros((cout, cin, ks, ks), np.float32)
    if kind == "identity":
        for o in range(min(cout, cin)):
            W[o, o, ks // 2, ks // 2] = 1.0
    elif kind.startswith("delta"):
        _, r0, s0 = kind.split(":")
        for o in range(min(cout, cin)):
            W[o, o, int(r0), int(s0)] = 1.0
    else:
        W = (rs.standard_normal(W.shape) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float16).astype(np.float32)
    b = np.zeros(cout, np.float32) if kind != "random" else (rs.standard_normal(cout) * 0.1).astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(W).double(),
                                     torch.from_numpy(b).double(), padding=ks // 2).permute(0, 2, 3, 1).numpy()
    for mode, prec in (("fast", native.PRECISION_FAST), ("parity", native.PRECISION_PARITY)):
        y = eng.test_conv(x, W, b, 0, prec)
        err = np.abs(y - ref)
        scale = max(np.abs(ref).max(), 1e-9)
        tol = (2e-3 if mode == "fast" else 1e-4) * scale
        bad = err > tol
        print("CASE %s mode %s: max_err %.3e tol %.3e scale %.3f bad %d/%d nan %d" % (
            cid, mode, err.max(), tol, scale, bad.sum(), bad.size, np.isnan(y).sum()))
        if bad.any():
            by_row = bad.reshape(n, h, w, cout).any(axis=(0, 2, 3))
            by_col = bad.reshape(n, h, w, cout).any(axis=(0, 1, 3))
            by_ch = bad.reshape(n, h, w, cout).any(axis=(0, 1, 2))
            print("  bad rows:", "".join("X" if v else "." for v in by_row))
            print("  bad cols:", "".join("X" if v else "." for v in by_col))
            print("  bad chans:", "".join("X" if v else "." for v in by_ch))
            if kind != "random":
                # where does each wrong output actually come from?
                shown = 0
                for (ni, yi, xi, ci) in zip(*np.nonzero(bad)):
                    val = y[ni, yi, xi, ci]
                    hits = np.argwhere(np.isclose(x[ni], val, rtol=0, atol=1e-6) & (np.abs(x[ni]) > 1e-3))
                    print("  out[n%d,y%d,x%d,c%d]=%.4f expected %.4f; equals x at (y,x,c)=%s" % (
                        ni, yi, xi, ci, val, ref[ni, yi, xi, ci], hits[:3].tolist()))
                    shown += 1
                    if shown >= 12:
                        break
    sys.stdout.flush()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for cid in CASES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), cid], stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True, timeout=300)
            out = r.stdout.strip().splitlines()
            print("\n".join(out[-40:]) if out else "(no output)")
            if r.returncode != 0:
                print("CASE %s exited with code %d" % (cid, r.returncode))
            sys.stdout.flush()
