"""B200-native OpenPose inference hot path behind the API of the reference's pose_detector.py.

    from <package> import PoseDetector, draw_person_pose, params, JointType

or, flat like the reference: put this directory on sys.path and `import pose_detector`.
The CUDA library (libopb.so) is built by `__graft_entry__.build()`; importing this package
does not need a GPU, constructing a PoseDetector does."""
from .entity import JointType, params  # noqa: F401
from .pose_detector import PoseDetector, draw_person_pose, make_opb_params  # noqa: F401
from . import _native, synthetic  # noqa: F401
