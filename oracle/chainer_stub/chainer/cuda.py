"""ORACLE ONLY. chainer.cuda subset: always selects the reference's CPU branch
(pose_detector.py:80-82)."""
import numpy as np


def get_array_module(*args):
    return np


def to_cpu(x):
    return x


def to_gpu(x, device=None):
    raise RuntimeError("oracle chainer stub is CPU only")


def get_device_from_id(i):
    raise RuntimeError("oracle chainer stub is CPU only")
