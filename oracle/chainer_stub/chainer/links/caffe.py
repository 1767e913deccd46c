"""ORACLE ONLY. Import-only placeholder for chainer.links.caffe."""


class CaffeFunction(object):
    def __init__(self, *a, **k):
        raise RuntimeError("caffe import is not available in the oracle stub")
