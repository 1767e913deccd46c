"""ORACLE ONLY. chainer.functions subset used by the reference hot path
(models/CocoPoseNet.py:136-260, pose_detector.py:501-502)."""
import numpy as np
import torch

from .variable import Variable, as_array


def relu(x):
    a = as_array(x)
    return Variable(np.maximum(a, np.float32(0)).astype(a.dtype, copy=False))


def max_pooling_2d(x, ksize, stride=None, pad=0, cover_all=True):
    # Chainer default cover_all=True == ceil mode [3p]
    stride = ksize if stride is None else stride
    t = torch.from_numpy(as_array(x))
    y = torch.nn.functional.max_pool2d(t, ksize, stride, pad, ceil_mode=bool(cover_all))
    return Variable(y.numpy())


def concat(xs, axis=1):
    return Variable(np.concatenate([as_array(v) for v in xs], axis=axis))


def resize_images(x, output_shape):
    """Restatement of Chainer v2..v5 ResizeImages.forward [3p]: align-corners
    bilinear; grid = linspace in float64; weights formed in float64 and cast to
    x.dtype; four-term sum in x.dtype, left to right."""
    a = as_array(x)
    B, C, H, W = a.shape
    out_H, out_W = int(output_shape[0]), int(output_shape[1])
    u_1d = np.linspace(0, W - 1, num=out_W)
    v_1d = np.linspace(0, H - 1, num=out_H)
    grid = np.meshgrid(u_1d, v_1d)
    u = grid[0].ravel()
    v = grid[1].ravel()
    u0 = np.floor(u).astype(np.int32).clip(0, W - 2)
    u1 = u0 + 1
    v0 = np.floor(v).astype(np.int32).clip(0, H - 2)
    v1 = v0 + 1
    w1 = ((u1 - u) * (v1 - v)).astype(a.dtype)
    w2 = ((u - u0) * (v1 - v)).astype(a.dtype)
    w3 = ((u1 - u) * (v - v0)).astype(a.dtype)
    w4 = ((u - u0) * (v - v0)).astype(a.dtype)
    a = a.reshape(B * C, H, W)
    y = w1[None, :] * a[:, v0, u0]
    y += w2[None, :] * a[:, v0, u1]
    y += w3[None, :] * a[:, v1, u0]
    y += w4[None, :] * a[:, v1, u1]
    return Variable(y.reshape(B, C, out_H, out_W))


def convolution_2d(x, W, b=None, stride=1, pad=0):
    t = torch.from_numpy(as_array(x))
    w = torch.from_numpy(as_array(W))
    bb = None if b is None else torch.from_numpy(as_array(b))
    return Variable(torch.nn.functional.conv2d(t, w, bb, stride=stride, padding=pad).numpy())
