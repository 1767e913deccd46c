"""Host-side helpers of the drop-in `pose_detector` module (get_unit_length, crop_face,
crop_hands, crop_image, crop_around_keypoint, draw_person_pose) against the reference's own methods, executed verbatim
(oracle/reference_loader.py), on random poses.  CPU only; skipped where /root/reference is absent (GPU box)."""
import numpy as np
import pytest

from conftest import pkg
from oracle import reference_loader

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")


def _dets():
    ref = reference_loader.load()
    PD = pkg("pose_detector").PoseDetector
    return ref, ref.PoseDetector.__new__(ref.PoseDetector), PD.__new__(PD)


def _random_pose(rs, h, w, p_missing):
    pose = np.zeros((18, 3))
    pose[:, 0] = rs.uniform(0, w, 18)
    pose[:, 1] = rs.uniform(0, h, 18)
    pose[:, 2] = 2
    pose[rs.uniform(size=18) < p_missing] = 0
    return pose


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, dict):
        return set(a) == set(b) and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (tuple, list)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(12))
def test_unit_length_and_crops_match_reference(seed):
    ref, rdet, det = _dets()
    rs = np.random.RandomState(seed)
    h, w = int(rs.randint(120, 500)), int(rs.randint(120, 700))
    img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    for p_missing in (0.0, 0.3, 0.7):
        pose = _random_pose(rs, h, w, p_missing)
        ul_ref = rdet.get_unit_length(pose.copy())
        ul = det.get_unit_length(pose.copy())
        assert (ul_ref is None and ul is None) or np.isclose(ul, ul_ref, rtol=0, atol=0) or ul == ul_ref
        if ul_ref is None or not np.isfinite(ul_ref) or ul_ref <= 0:
            continue
        assert _same(det.crop_face(img, pose.copy(), ul), rdet.crop_face(img, pose.copy(), ul_ref))
        assert _same(det.crop_hands(img, pose.copy(), ul), rdet.crop_hands(img, pose.copy(), ul_ref))
        bbox = (int(rs.randint(-50, w)), int(rs.randint(-50, h)), int(rs.randint(10, w + 80)), int(rs.randint(10, h + 80)))
        if bbox[2] > bbox[0] + 2 and bbox[3] > bbox[1] + 2:
            assert _same(det.crop_image(img, bbox), rdet.crop_image(img, bbox))
        kp = (float(rs.uniform(0, w)), float(rs.uniform(0, h)))
        half = float(rs.uniform(10, 120))
        assert _same(det.crop_around_keypoint(img, kp, half), rdet.crop_around_keypoint(img, kp, half))


@pytest.mark.parametrize("seed", range(6))
def test_draw_person_pose_matches_reference(seed):
    ref, _, _ = _dets()
    draw = pkg("pose_detector").draw_person_pose
    rs = np.random.RandomState(100 + seed)
    img = rs.randint(0, 256, (240, 320, 3)).astype(np.uint8)
    poses = np.stack([_random_pose(rs, 240, 320, 0.25) for _ in range(int(rs.randint(1, 5)))])
    assert np.array_equal(draw(img, poses.copy()), ref.draw_person_pose(img, poses.copy()))
    assert np.array_equal(draw(img, np.empty((0, 18, 3))), ref.draw_person_pose(img, np.empty((0, 18, 3))))


@pytest.mark.parametrize("seed", range(8))
def test_crop_person_matches_reference(seed):
    """pose_detector.py:311-352.  The reference method raises NameError as shipped (`sys` is never imported, :1-12);
    the comparison supplies the missing global to the verbatim module."""
    import sys
    ref, rdet, det = _dets()
    rs = np.random.RandomState(300 + seed)
    h, w = int(rs.randint(200, 500)), int(rs.randint(200, 700))
    img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    with pytest.raises(NameError):
        ref.__dict__.pop("sys", None)
        rdet.crop_person(img, _random_pose(rs, h, w, 0.0), 20.0)
    ref.sys = sys
    try:
        for p_missing in (0.0, 0.3, 0.6):
            pose = _random_pose(rs, h, w, p_missing)
            if not (pose[:, 2] > 0).any():
                continue
            ul = float(rs.uniform(5, 40))
            # degenerate poses (no joint right of the first one: right_pos stays the int 0) make both raise --
            # AttributeError (`.astype` on an int, :346-349) in the reference, ValueError (negative crop size) here
            out = []
            for d in (det, rdet):
                try:
                    out.append(d.crop_person(img, pose.copy(), ul))
                except Exception as e:
                    out.append(Exception)
            assert (out[0] is out[1]) if isinstance(out[0], type) or isinstance(out[1], type) else _same(out[0], out[1])
    finally:
        ref.__dict__.pop("sys", None)
