"""Host-side weight container shared by the three nets (CocoPoseNet / FaceNet / HandNet): an object with one
attribute per conv layer exposing `.W.data` / `.b.data` like a Chainer link, loadable from Chainer's save_npz
layout.  There is no host forward: `__call__` runs on the B200 through the bound device engine.

Model <-> engine binding.  A detector snapshots the weights into its own device context when it is constructed
(`Engine.load_model`); detectors built from one model object therefore stay independent of each other.  The model's own
`__call__` (the reference's `self.model(x)`) uses the engine of the detector constructed LAST with it.  Changing
`link.W.data` afterwards does not reach an existing engine -- construct a new detector (or call
`detector.engine.load_model(model)`); `load_npz` drops the binding so that a stale engine cannot be used by accident."""
import numpy as np


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


class NetContainer(object):
    insize = 368
    LAYERS = ()

    def __init__(self, seed=None):
        # Chainer's default is LeCunNormal from numpy's *global* RNG (not reproducible);
        # here: sigma = sqrt(1/fan_in), b = 0, from RandomState(seed or 0).
        rs = np.random.RandomState(0 if seed is None else seed)
        self._names = []
        for name, cin, cout, k in self.LAYERS:
            W = (rs.standard_normal((cout, cin, k, k)) * np.sqrt(1.0 / (cin * k * k))).astype(np.float32)
            setattr(self, name, _ConvParam(W, np.zeros(cout, np.float32)))
            self._names.append(name)
        self._engine = None       # set by the detector (device context owning the packed weights)

    # -- chainer.Chain-like helpers -------------------------------------------------
    def children_items(self):
        return [(n, getattr(self, n)) for n in self._names]

    def load_npz(self, path_or_dict):
        """Chainer save_npz layout: '<layer>/W' [Cout,Cin,k,k] f32 and '<layer>/b' [Cout]."""
        f = np.load(path_or_dict) if isinstance(path_or_dict, str) else path_or_dict
        for name, cin, cout, k in self.LAYERS:
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

    def save_npz(self, path):
        """Writes the layout chainer.serializers.save_npz produces (and load_npz reads)."""
        with open(path, "wb") as f:     # np.savez(path) would append ".npz"; chainer writes the exact file name
            np.savez(f, **self.state_dict())

    def to_gpu(self, device=None):
        return self

    def to_cpu(self):
        return self

    def _bound_engine(self):
        if self._engine is None:
            raise RuntimeError("%s is not bound to a device engine; construct its detector with model=<this object> "
                               "(there is no CPU forward in this package)" % type(self).__name__)
        return self._engine
