"""HandNet layer table and weight container (host side).

Mirrors the interface of the reference class (models/HandNet.py:5-161): constructed with no arguments, owns 52
named conv layers, callable on an [N,3,368,368] float32 array and returns the list of six per-stage heat maps
([N,22,46,46]; 21 keypoints + background).  VGG-19 front (conv1_1..conv5_2), conv5_3_CPM -> 128 features, stage 1 =
two 1x1 convs, stages 2-6 = five 7x7 + two 1x1 convs on concat(previous maps, features) (:150 input channels).
The forward runs on the B200 (csrc/opb_api.cu, build_chain_keypoint); only the last stage is materialised and is
returned in every slot (the reference's callers read hs[-1], hand_detector.py:43)."""
import numpy as np

try:
    from ._container import NetContainer
except ImportError:  # flat import, like the reference
    from models._container import NetContainer

N_OUT = 22


def _build_layer_table():
    t = [("conv1_1", 3, 64, 3), ("conv1_2", 64, 64, 3), ("conv2_1", 64, 128, 3), ("conv2_2", 128, 128, 3),
         ("conv3_1", 128, 256, 3)]
    t += [("conv3_%d" % i, 256, 256, 3) for i in (2, 3, 4)]
    t += [("conv4_1", 256, 512, 3)] + [("conv4_%d" % i, 512, 512, 3) for i in (2, 3, 4)]
    t += [("conv5_1", 512, 512, 3), ("conv5_2", 512, 512, 3), ("conv5_3_CPM", 512, 128, 3)]
    t += [("conv6_1_CPM", 128, 512, 1), ("conv6_2_CPM", 512, N_OUT, 1)]
    for stage in range(2, 7):
        t.append(("Mconv1_stage%d" % stage, 128 + N_OUT, 128, 7))
        t += [("Mconv%d_stage%d" % (i, stage), 128, 128, 7) for i in (2, 3, 4, 5)]
        t.append(("Mconv6_stage%d" % stage, 128, 128, 1))
        t.append(("Mconv7_stage%d" % stage, 128, N_OUT, 1))
    return tuple(t)


#: (name, in_channels, out_channels, ksize) in the reference's declaration order.
LAYERS = _build_layer_table()
assert len(LAYERS) == 52


class HandNet(NetContainer):
    LAYERS = LAYERS

    def __call__(self, x):
        heat = self._bound_engine().forward_keypoint_maps(np.ascontiguousarray(x, np.float32))
        return [heat] * 6
