"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal CPU stand-in for the `chainer` package so that the reference's own
Python files (/root/reference/pose_detector.py, entity.py, models/CocoPoseNet.py)
execute verbatim in a container where Chainer cannot be installed (no network).

Third-party arithmetic restated here (Chainer itself is NOT under /root/reference;
the reference pins nothing beyond "Chainer 2.0+", README.md:36):

* links.Convolution2D   -> torch CPU fp32 conv2d (cross-correlation + bias), the
                           pinned conv restatement (torch 2.11.0, oneDNN).
* functions.max_pooling_2d -> 2x2/2 max with Chainer's default cover_all=True
                           (== ceil mode).
* functions.resize_images  -> restatement of Chainer's published v2..v5
                           ResizeImages.forward: align-corners bilinear, sample
                           grid from numpy.linspace in float64, the four tap
                           weights formed in float64 then cast to the input dtype,
                           y = w1*x00; y += w2*x01; y += w3*x10; y += w4*x11.
* serializers.load_npz  -> "<link>/W", "<link>/b" arrays copied into the links.

Call sites in the reference: models/CocoPoseNet.py:26-129,136-260,
pose_detector.py:26,80,501-502.
"""
import contextlib

import numpy as np

from . import functions, links, serializers, cuda  # noqa: F401
from .variable import Variable  # noqa: F401


class _Config(object):
    enable_backprop = False
    train = False


config = _Config()


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name, None)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)


class Link(object):
    def to_gpu(self, device=None):
        raise RuntimeError("oracle chainer stub is CPU only")

    def to_cpu(self):
        return self


class Chain(Link):
    """chainer.Chain(**links): children are registered as attributes
    (models/CocoPoseNet.py:24-130)."""

    def __init__(self, **lnks):
        self._children = []
        for name, link in lnks.items():
            setattr(self, name, link)
            self._children.append(name)

    def children_items(self):
        return [(n, getattr(self, n)) for n in self._children]
