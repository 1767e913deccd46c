"""ORACLE ONLY. Variable-like wrapper: `.data` (numpy), `var[i]`, shape
(used at pose_detector.py:453-454,501-502)."""
import numpy as np


class Variable(object):
    def __init__(self, data):
        self.data = np.ascontiguousarray(data)

    @property
    def array(self):
        return self.data

    @property
    def shape(self):
        return self.data.shape

    def __getitem__(self, idx):
        return Variable(self.data[idx])

    def __len__(self):
        return len(self.data)


def as_array(x):
    return x.data if isinstance(x, Variable) else np.asarray(x)
