"""Constants of the OpenPose inference path -- the `entity` module of the drop-in boundary
(`from entity import params, JointType`, reference demo.py:4, pose_detector.py:11).

Values restate entity.py:9-46 (JointType), :71-105 (inference params, limbs_point) and the
few other keys the pose path reads (:59 'downscale'), plus the face / hand inference keys (:126-152) used by
face_detector / hand_detector.  The training keys of the reference dict are outside the accelerated path and
are intentionally absent."""
from enum import IntEnum

try:
    from .models.CocoPoseNet import CocoPoseNet
    from .models.FaceNet import FaceNet
    from .models.HandNet import HandNet
except ImportError:  # flat import, like the reference
    from models.CocoPoseNet import CocoPoseNet
    from models.FaceNet import FaceNet
    from models.HandNet import HandNet

_JOINT_NAMES = ("Nose Neck RightShoulder RightElbow RightHand LeftShoulder LeftElbow LeftHand RightWaist "
                "RightKnee RightFoot LeftWaist LeftKnee LeftFoot RightEye LeftEye RightEar LeftEar").split()

JointType = IntEnum("JointType", [(n, i) for i, n in enumerate(_JOINT_NAMES)])

_J = JointType
_LIMB_NAMES = (("Neck", "RightWaist"), ("RightWaist", "RightKnee"), ("RightKnee", "RightFoot"), ("Neck", "LeftWaist"),
               ("LeftWaist", "LeftKnee"), ("LeftKnee", "LeftFoot"), ("Neck", "RightShoulder"),
               ("RightShoulder", "RightElbow"), ("RightElbow", "RightHand"), ("RightShoulder", "RightEar"),
               ("Neck", "LeftShoulder"), ("LeftShoulder", "LeftElbow"), ("LeftElbow", "LeftHand"),
               ("LeftShoulder", "LeftEar"), ("Neck", "Nose"), ("Nose", "RightEye"), ("Nose", "LeftEye"),
               ("RightEye", "RightEar"), ("LeftEye", "LeftEar"))



def _chain(first, last, closed=False):
    """[[first, first+1], ..., [last-1, last]] (+ [last, first] when the polyline is closed)."""
    pairs = [[i, i + 1] for i in range(first, last)]
    return pairs + ([[last, first]] if closed else [])


params = {
    "archs": {"posenet": CocoPoseNet, "facenet": FaceNet, "handnet": HandNet},
    "insize": 368,
    "downscale": 8,
    # inference (entity.py:71-84)
    "inference_img_size": 368,
    "inference_scales": [0.5, 1, 1.5, 2],
    "heatmap_size": 320,
    "gaussian_sigma": 2.5,
    "ksize": 17,
    "n_integ_points": 10,
    "n_integ_points_thresh": 8,
    "heatmap_peak_thresh": 0.05,
    "inner_product_thresh": 0.05,
    "limb_length_ratio": 1.0,
    "length_penalty_value": 1,
    "n_subset_limbs_thresh": 3,
    "subset_score_thresh": 0.2,
    "limbs_point": [[_J[a], _J[b]] for a, b in _LIMB_NAMES],
    # face (entity.py:126-140): 70-point layout -- jaw line, brows, nose bridge, nostrils, eyes, outer / inner lips
    "face_inference_img_size": 368,
    "face_heatmap_peak_thresh": 0.1,
    "face_crop_scale": 1.5,
    "face_line_indices": (_chain(0, 16) + _chain(17, 21) + _chain(22, 26) + _chain(27, 30) + _chain(31, 35) +
                          _chain(36, 41, True) + _chain(42, 47, True) + _chain(48, 59, True) + _chain(60, 67, True)),
    # hand (entity.py:142-152): wrist 0, four joints per finger
    "hand_inference_img_size": 368,
    "hand_heatmap_peak_thresh": 0.1,
    "fingers_indices": [[[0, 4 * f + 1]] + _chain(4 * f + 1, 4 * f + 4) for f in range(5)],
}
