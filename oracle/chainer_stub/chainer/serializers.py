"""ORACLE ONLY. serializers.load_npz/save_npz: keys "<link>/W", "<link>/b"
(pose_detector.py:26)."""
import numpy as np


def load_npz(path, model):
    with np.load(path) as f:
        for name, link in model.children_items():
            link.W.data = np.ascontiguousarray(f[name + "/W"], dtype=np.float32)
            link.b.data = np.ascontiguousarray(f[name + "/b"], dtype=np.float32)


def save_npz(path, model):
    d = {}
    for name, link in model.children_items():
        d[name + "/W"] = link.W.data
        d[name + "/b"] = link.b.data
    np.savez(path, **d)
