"""CocoPoseNet layer table and weight container (host side).

Mirrors the *interface* of the reference class (models/CocoPoseNet.py:20-262): an object
constructed with no arguments that owns 92 named conv layers and is callable on an
[N,3,H,W] float32 array, returning (pafs, heatmaps) lists of per-stage outputs.  There is
no Chainer graph here: the forward runs on the B200 through the C-ABI library
(csrc/opb_api.cu), which executes the whole 92-conv chain as tcgen05 implicit-GEMM
kernels.  Only the last stage is materialised on the host fast path; `__call__` keeps the
six-entry list shape of the reference by returning the final stage in every slot that the
reference's callers actually read ([-1], pose_detector.py:453-454,501-502).
"""
import numpy as np

try:
    from ._container import NetContainer, _ConvParam  # noqa: F401
except ImportError:  # flat import, like the reference
    from models._container import NetContainer, _ConvParam  # noqa: F401


def _build_layer_table():
    t = [("conv1_1", 3, 64, 3), ("conv1_2", 64, 64, 3), ("conv2_1", 64, 128, 3), ("conv2_2", 128, 128, 3),
         ("conv3_1", 128, 256, 3)]
    t += [("conv3_%d" % i, 256, 256, 3) for i in (2, 3, 4)]
    t += [("conv4_1", 256, 512, 3), ("conv4_2", 512, 512, 3), ("conv4_3_CPM", 512, 256, 3),
          ("conv4_4_CPM", 256, 128, 3)]
    for branch, n_out in (("L1", 38), ("L2", 19)):
        t += [("conv5_%d_CPM_%s" % (i, branch), 128, 128, 3) for i in (1, 2, 3)]
        t += [("conv5_4_CPM_%s" % branch, 128, 512, 1), ("conv5_5_CPM_%s" % branch, 512, n_out, 1)]
    for stage in range(2, 7):
        for branch, n_out in (("L1", 38), ("L2", 19)):
            t.append(("Mconv1_stage%d_%s" % (stage, branch), 185, 128, 7))
            t += [("Mconv%d_stage%d_%s" % (i, stage, branch), 128, 128, 7) for i in (2, 3, 4, 5)]
            t.append(("Mconv6_stage%d_%s" % (stage, branch), 128, 128, 1))
            t.append(("Mconv7_stage%d_%s" % (stage, branch), 128, n_out, 1))
    return tuple(t)


#: (name, in_channels, out_channels, ksize) in the reference's declaration order.
LAYERS = _build_layer_table()
assert len(LAYERS) == 92


def conv_flops_per_image(h, w):
    """2*MAC FLOPs of the 92-conv chain for an h x w input (true channel counts)."""
    total = 0
    for name, cin, cout, k in LAYERS:
        if name.startswith("conv1"):
            s = 1
        elif name.startswith("conv2"):
            s = 2
        elif name.startswith("conv3"):
            s = 4
        else:
            s = 8
        total += 2 * cin * cout * k * k * (h // s) * (w // s)
    return total


class CocoPoseNet(NetContainer):
    LAYERS = LAYERS

    def __call__(self, x):
        paf, heat = self._bound_engine().forward(np.ascontiguousarray(x, np.float32))
        return [paf] * 6, [heat] * 6
