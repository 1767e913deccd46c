"""Import shim, NOT Chainer: lets the reference's demo scripts (`camera_pose_demo.py:3,6`,
`demo.py:3,9`), which do `import chainer; chainer.using_config('enable_backprop', False)`,
import on a machine without Chainer.  Put `<package>/compat` on sys.path only when the real
Chainer is absent.  There is no compute here; the pose path runs in libopb (sm_100a CUDA)."""
import contextlib


class _Config(object):
    enable_backprop = False
    train = False


config = _Config()


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name, None)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)
