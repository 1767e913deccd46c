"""Weight tooling: Chainer save_npz layout round trip and the caffemodel importer, checked on
synthetic caffemodels written with a minimal protobuf encoder (new `layer` and legacy V1 `layers` records,
BlobShape and legacy num/channels/height/width dims, packed and unpacked float data)."""
import os
import struct

import numpy as np
import pytest

from conftest import pkg


def _vi(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fno, payload):
    return _vi((fno << 3) | 2) + _vi(len(payload)) + payload


def _blob(a, legacy, packed=True):
    a = np.ascontiguousarray(a, "<f4")
    dims = list(a.shape)
    if legacy:
        d4 = [1] * (4 - len(dims)) + dims
        head = b"".join(_vi((i + 1) << 3) + _vi(d) for i, d in enumerate(d4))
    else:
        head = _ld(7, _ld(1, b"".join(_vi(d) for d in dims)))
    if packed:
        body = _ld(5, a.tobytes())
    else:
        body = b"".join(_vi((5 << 3) | 5) + struct.pack("<f", v) for v in a.ravel())
    return head + body


def _caffemodel(weights, layers, v1):
    out = _ld(1, b"synthetic")                                         # NetParameter.name
    for i, (name, cin, cout, k) in enumerate(layers):
        W, b = weights[name + "/W"], weights[name + "/b"]
        blobs = _blob(W, legacy=v1, packed=(i % 3 != 1 or W.size > 4096)) + b""
        if v1:
            rec = _ld(4, name.encode()) + _vi((5 << 3)) + _vi(4) + _ld(6, _blob(W, True)) + _ld(6, _blob(b, True))
            out += _ld(2, rec)
            out += _ld(2, _ld(4, ("relu_" + name).encode()) + _vi(5 << 3) + _vi(18))      # a blob-less layer in between
        else:
            rec = _ld(1, name.encode()) + _ld(2, b"Convolution") + _ld(7, _blob(W, False, i % 3 != 1 or W.size > 4096)) + \
                _ld(7, _blob(b, False, packed=bool(i % 2)))
            out += _ld(100, rec)
            out += _ld(100, _ld(1, ("relu_" + name).encode()) + _ld(2, b"ReLU"))
    return out


@pytest.mark.parametrize("arch,v1", [("handnet", False), ("facenet", True), ("posenet", False)])
def test_caffemodel_importer_round_trip(tmp_path, arch, v1):
    cm = pkg("models.convert_model")
    cls = pkg("entity").params["archs"][arch]
    wd = pkg("synthetic").he_weights(3, layers=cls.LAYERS)
    path = os.path.join(tmp_path, arch + ".caffemodel")
    with open(path, "wb") as f:
        f.write(_caffemodel(wd, cls.LAYERS, v1))
    out = os.path.join(tmp_path, arch + ".npz")
    net = cm.convert(arch, path, out)
    with np.load(out) as z:
        assert sorted(z.files) == sorted(wd)                 # incl. conv5_5_CPM_L1, which the reference's list omits
        for k in wd:
            assert z[k].dtype == np.float32 and np.array_equal(z[k], wd[k]), k
    again = cls()
    again.load_npz(out)                                      # the layout chainer.serializers.load_npz reads
    assert all(np.array_equal(a.W.data, b.W.data) for (_, a), (_, b) in zip(net.children_items(), again.children_items()))


def test_caffemodel_importer_errors(tmp_path):
    cm = pkg("models.convert_model")
    cls = pkg("entity").params["archs"]["handnet"]
    wd = pkg("synthetic").he_weights(3, layers=cls.LAYERS)
    short = [l for l in cls.LAYERS if l[0] != "Mconv7_stage6"]
    p = os.path.join(tmp_path, "missing.caffemodel")
    open(p, "wb").write(_caffemodel(wd, short, False))
    with pytest.raises(KeyError):
        cm.convert("handnet", p)
    bad = dict(wd)
    bad["conv1_1/W"] = np.zeros((64, 3, 5, 5), np.float32)
    p = os.path.join(tmp_path, "shape.caffemodel")
    open(p, "wb").write(_caffemodel(bad, cls.LAYERS, False))
    with pytest.raises(ValueError):
        cm.convert("handnet", p)
