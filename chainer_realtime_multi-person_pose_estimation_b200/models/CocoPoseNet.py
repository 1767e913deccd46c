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


class _ConvParam(object):
    """Stand-in for a Chainer link: `.W.data` / `.b.data` numpy arrays."""

    class _P(object):
        def __init__(self, a):
            self.data = a

        @property
        def array(self):
            return self.data

    def __init__(self, W, b):
        self.W = _ConvParam._P(W)
        self.b = _ConvParam._P(b)


class CocoPoseNet(object):
    insize = 368

    def __init__(self, seed=None):
        # Chainer's default is LeCunNormal from numpy's *global* RNG (not reproducible);
        # here: sigma = sqrt(1/fan_in), b = 0, from RandomState(seed or 0).
        rs = np.random.RandomState(0 if seed is None else seed)
        self._names = []
        for name, cin, cout, k in LAYERS:
            W = (rs.standard_normal((cout, cin, k, k)) * np.sqrt(1.0 / (cin * k * k))).astype(np.float32)
            setattr(self, name, _ConvParam(W, np.zeros(cout, np.float32)))
            self._names.append(name)
        self._engine = None       # set by PoseDetector (device context owning the packed weights)

    # -- chainer.Chain-like helpers -------------------------------------------------
    def children_items(self):
        return [(n, getattr(self, n)) for n in self._names]

    def load_npz(self, path_or_dict):
        """Chainer save_npz layout: '<layer>/W' [Cout,Cin,k,k] f32 and '<layer>/b' [Cout]."""
        f = np.load(path_or_dict) if isinstance(path_or_dict, str) else path_or_dict
        for name, cin, cout, k in LAYERS:
            W = np.ascontiguousarray(f[name + "/W"], np.float32)
            b = np.ascontiguousarray(f[name + "/b"], np.float32)
            if W.shape != (cout, cin, k, k) or b.shape != (cout,):
                raise ValueError("bad shape for layer %s: %s %s" % (name, W.shape, b.shape))
            link = getattr(self, name)
            link.W.data, link.b.data = W, b
        self._engine = None

    def state_dict(self):
        d = {}
        for n, l in self.children_items():
            d[n + "/W"] = l.W.data
            d[n + "/b"] = l.b.data
        return d

    def to_gpu(self, device=None):
        return self

    def to_cpu(self):
        return self

    def __call__(self, x):
        if self._engine is None:
            raise RuntimeError("CocoPoseNet is not bound to a device engine; construct a PoseDetector "
                               "with model=<this object> (there is no CPU forward in this package)")
        paf, heat = self._engine.forward(np.ascontiguousarray(x, np.float32))
        return [paf] * 6, [heat] * 6
