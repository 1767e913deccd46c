"""caffemodel -> Chainer-layout .npz (replaces the reference's models/convert_model.py:257-281).

    python convert_model.py {posenet,facenet,handnet} pose_iter_440000.caffemodel coco_posenet.npz

The reference goes through chainer.links.caffe.CaffeFunction (Chainer + protobuf bindings) and copies the layers of
a hand-written name list -- which omits `conv5_5_CPM_L1` (convert_model.py:25-33), so the stage-1 PAF head of a
converted posenet keeps its random init.  Here the binary NetParameter is read directly (protobuf wire format, no
caffe.proto needed): every layer of the target net's own table is looked up by name, shapes are checked, and a
missing or mismatching layer is an error instead of a printed warning.

Wire format used (caffe.proto): NetParameter.layer = 100 (LayerParameter: name = 1, blobs = 7) and the legacy
NetParameter.layers = 2 (V1LayerParameter: name = 4, blobs = 6); BlobProto: num/channels/height/width = 1..4,
data = 5 (packed or repeated float), shape = 7 (BlobShape.dim = 1, packed or repeated int64), double_data = 8."""
import argparse
import struct

import numpy as np


def _varint(buf, pos):
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _fields(buf):
    """Yields (field_number, wire_type, value) of one message; value is int (varint / fixed) or a memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _packed_varints(buf):
    out, pos = [], 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(v)
    return out


def _blob(buf):
    legacy, dims, chunks, singles = {}, [], [], []
    dtype = np.float32
    for fno, wt, v in _fields(buf):
        if fno in (1, 2, 3, 4) and wt == 0:
            legacy[fno] = v
        elif fno == 5:
            if wt == 2:
                chunks.append(np.frombuffer(v, "<f4"))
            else:
                singles.append(struct.unpack("<f", bytes(v))[0])
        elif fno == 8:
            dtype = np.float64
            if wt == 2:
                chunks.append(np.frombuffer(v, "<f8"))
            else:
                singles.append(struct.unpack("<d", bytes(v))[0])
        elif fno == 7 and wt == 2:
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    dims += _packed_varints(v2) if w2 == 2 else [v2]
    data = np.concatenate(chunks + ([np.asarray(singles, dtype)] if singles else [])) if (chunks or singles) \
        else np.zeros(0, dtype)
    if not dims and legacy:
        dims = [legacy.get(i, 1) for i in (1, 2, 3, 4)]
    return np.asarray(data, np.float32), [int(d) for d in dims]


def read_caffemodel(path):
    """{layer name: [(float32 data, dims), ...]} for every layer that carries blobs."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    layers = {}
    for fno, wt, v in _fields(buf):
        if wt != 2 or fno not in (2, 100):
            continue
        name_f, blob_f = (4, 6) if fno == 2 else (1, 7)
        name, blobs = None, []
        for f2, w2, v2 in _fields(v):
            if f2 == name_f and w2 == 2:
                name = bytes(v2).decode("utf-8")
            elif f2 == blob_f and w2 == 2:
                blobs.append(_blob(v2))
        if name is not None and blobs:
            layers[name] = blobs
    return layers


def convert(arch, caffe_file, chainer_file=None):
    """Fills a fresh `arch` net from the caffemodel and (optionally) writes the Chainer-layout .npz.
    Returns the net object.  Raises KeyError / ValueError for a missing layer or a shape mismatch."""
    try:
        from ..entity import params
    except ImportError:  # flat import, like the reference
        from entity import params
    net = params["archs"][arch]()
    caffe = read_caffemodel(caffe_file)
    for name, cin, cout, k in net.LAYERS:
        if name not in caffe:
            raise KeyError("layer %s not found in %s" % (name, caffe_file))
        blobs = caffe[name]
        if len(blobs) < 2:
            raise ValueError("layer %s has no bias blob" % name)
        (wd, wdims), (bd, _) = blobs[0], blobs[1]
        if wd.size != cout * cin * k * k or bd.size != cout or (len(wdims) == 4 and tuple(wdims) != (cout, cin, k, k)):
            raise ValueError("layer %s: caffemodel blob dims %s / %d values do not match %s" % (
                name, wdims, wd.size, (cout, cin, k, k)))
        link = getattr(net, name)
        link.W.data = np.ascontiguousarray(wd.reshape(cout, cin, k, k))
        link.b.data = np.ascontiguousarray(bd.reshape(cout))
    if chainer_file:
        net.save_npz(chainer_file)
    return net


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Convert caffemodel into chainermodel")
    ap.add_argument("arch", help="model architecture: ['posenet', 'facenet', 'handnet']")
    ap.add_argument("caffe_file", help="caffe weights file path")
    ap.add_argument("chainer_file", help="file path to save chainer weights file")
    a = ap.parse_args()
    print("Loading caffemodel file...")
    convert(a.arch, a.caffe_file, a.chainer_file)
    print("Saved weights file into '%s'." % a.chainer_file)
