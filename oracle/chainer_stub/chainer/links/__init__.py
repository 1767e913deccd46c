"""ORACLE ONLY. chainer.links subset: Convolution2D (models/CocoPoseNet.py:26-129)."""
import numpy as np
import torch

from ..variable import Variable, as_array
from . import caffe  # noqa: F401  (models/CocoPoseNet.py:4 imports it)


class _Param(object):
    def __init__(self, data):
        self.data = data

    @property
    def array(self):
        return self.data


class Convolution2D(object):
    """W [Cout,Cin,k,k] f32, b [Cout] f32; call = cross-correlation + bias,
    stride 1, zero padding `pad` (torch CPU fp32 conv2d restatement).
    Default init follows Chainer's LeCunNormal (sigma = sqrt(1/fan_in)), b = 0,
    drawn from a FIXED seed derived from the layer shape so oracle runs are
    reproducible (Chainer draws from numpy's global RNG; not reproducible)."""

    _init_counter = [0]

    def __init__(self, in_channels, out_channels, ksize, stride=1, pad=0):
        self.in_channels, self.out_channels = in_channels, out_channels
        self.ksize, self.stride, self.pad = ksize, stride, pad
        fan_in = in_channels * ksize * ksize
        rs = np.random.RandomState(1000 + Convolution2D._init_counter[0] % 92)
        Convolution2D._init_counter[0] += 1
        self.W = _Param((rs.standard_normal((out_channels, in_channels, ksize, ksize))
                         * np.sqrt(1.0 / fan_in)).astype(np.float32))
        self.b = _Param(np.zeros(out_channels, np.float32))

    def __call__(self, x):
        t = torch.from_numpy(np.ascontiguousarray(as_array(x), dtype=np.float32))
        y = torch.nn.functional.conv2d(t, torch.from_numpy(self.W.data), torch.from_numpy(self.b.data),
                                       stride=self.stride, padding=self.pad)
        return Variable(y.numpy())
