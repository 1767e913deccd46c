"""ORACLE / TEST INFRASTRUCTURE ONLY.

Imports the reference's own Python files *verbatim* from /root/reference over the
chainer stub in oracle/chainer_stub (Chainer itself cannot be installed offline).
Only usable in the build container: /root/reference does not exist on the GPU box.
It is used by oracle/make_goldens.py to produce the committed fixtures under
tests/golden/ and by the `not gpu` tests that pin oracle/restate.py against the
reference when /root/reference is present.

One in-memory source fix is applied (nothing is written to /root/reference):
pose_detector.py:147 indexes with a *list* of arrays, which NumPy >= 1.23 treats as
an array index (-> ValueError at :149). The old tuple semantics are restored:
    paf[0][np.hsplit(integ_points, 2)]  ->  paf[0][tuple(np.hsplit(integ_points, 2))]
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("OPB_REFERENCE_ROOT", "/root/reference")
_STUB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chainer_stub")

_OLD = "paf[0][np.hsplit(integ_points, 2)], paf[1][np.hsplit(integ_points, 2)]"
_NEW = "paf[0][tuple(np.hsplit(integ_points, 2))], paf[1][tuple(np.hsplit(integ_points, 2))]"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pose_detector.py"))


_REF_MODS = ("chainer", "entity", "models", "models.CocoPoseNet", "models.FaceNet", "models.HandNet", "pose_detector",
             "face_detector", "hand_detector")


def _exec_reference(filename, modname, fixes=()):
    """Executes one reference file verbatim (plus the listed in-memory source fixes) as module `modname`."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if modname in sys.modules:
        return sys.modules[modname]
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in _REF_MODS}
    for k in list(sys.modules):
        if k == "chainer" or k.startswith("chainer.") or k in saved_mods:
            del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _STUB_DIR)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with open(os.path.join(REFERENCE_ROOT, filename), "r") as f:
                src = f.read()
            for old, new in fixes:
                assert src.count(old) == 1, "reference source changed; fix does not apply"
                src = src.replace(old, new)
            mod = types.ModuleType(modname)
            mod.__file__ = os.path.join(REFERENCE_ROOT, filename)
            code = compile(src, mod.__file__, "exec")
            mod.__dict__["__name__"] = modname
            exec(code, mod.__dict__)
            mod.ref_entity = sys.modules["entity"]
            mod.ref_chainer = sys.modules["chainer"]
    finally:
        sys.path[:] = saved_path
        # leave the reference modules reachable only through `mod`
        for k in list(sys.modules):
            if k == "chainer" or k.startswith("chainer.") or k in _REF_MODS[1:]:
                del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
    sys.modules[modname] = mod
    return mod


def load():
    """Returns the reference `pose_detector` module object (executed verbatim + the numpy-2 fix)."""
    return _exec_reference("pose_detector.py", "ref_pose_detector", ((_OLD, _NEW),))


def load_face():
    """The reference `face_detector` module (face_detector.py), verbatim."""
    return _exec_reference("face_detector.py", "ref_face_detector")


def load_hand():
    """The reference `hand_detector` module (hand_detector.py), verbatim."""
    return _exec_reference("hand_detector.py", "ref_hand_detector")
