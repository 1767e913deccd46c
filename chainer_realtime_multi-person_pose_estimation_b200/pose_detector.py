"""B200-native drop-in for the reference module `pose_detector` (reference pose_detector.py).

Same public surface -- `PoseDetector(arch, weights_file, model, device, precise)`,
`__call__`, `compute_peaks_from_heatmaps`, `compute_connections`, `grouping_key_points`,
`subsets_to_pose_array`, `detect_precise`, the crop / unit-length helpers and
`draw_person_pose` -- so `demo.py` / `camera_pose_demo.py` run unchanged.  All numerics of
the hot path (CocoPoseNet forward, upsampling, Gaussian-smoothed peak extraction, PAF line
integrals, greedy limb assignment, person grouping) execute as hand-written sm_100a CUDA
behind the C ABI of include/opb.h (libopb.so, loaded with ctypes).  Host code only moves
NumPy arrays, resizes the uint8 input image with OpenCV exactly as the reference does
(pose_detector.py:443,493) and rebuilds the float64 pose array.

Differences from the reference, by design:
  * `device < 0` selects GPU 0 (the reference's CPU mode does not exist here; there is no
    CPU fallback anywhere in this package).
  * extra keyword `precision`: "parity" (split-fp16 operands, ~3e-5 from fp32; default) or
    "fast" (plain fp16 operands, fp32 accumulate).
"""
import math
import os

import cv2
import numpy as np

try:  # imported as part of the package ...
    from . import _native
    from .entity import JointType, params
    from .models.CocoPoseNet import CocoPoseNet
except ImportError:  # ... or flat, like the reference (package directory on sys.path)
    import _native
    from entity import JointType, params
    from models.CocoPoseNet import CocoPoseNet


def _gaussian_taps(sigma, truncate=4.0):
    """The float64 taps scipy.ndimage.gaussian_filter uses (pose_detector.py:86) -- formed on
    the host with NumPy exactly as scipy's _gaussian_kernel1d does, then handed to the device."""
    radius = int(truncate * float(sigma) + 0.5)
    xs = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * xs ** 2)
    return radius, phi / phi.sum()


def make_opb_params(p=params, max_peaks=None, max_candidates=None, max_persons=None):
    """entity.params -> the POD struct of the C ABI (include/opb.h: opb_params)."""
    s = _native.OpbParams()
    for i, (a, b) in enumerate(p["limbs_point"]):
        s.limbs[i][0], s.limbs[i][1] = int(a), int(b)
    s.heatmap_peak_thresh = p["heatmap_peak_thresh"]
    s.inner_product_thresh = p["inner_product_thresh"]
    s.limb_length_ratio = p["limb_length_ratio"]
    s.length_penalty_value = p["length_penalty_value"]
    s.n_subset_limbs_thresh = p["n_subset_limbs_thresh"]
    s.subset_score_thresh = p["subset_score_thresh"]
    s.n_integ_points = p["n_integ_points"]
    s.n_integ_points_thresh = p["n_integ_points_thresh"]
    radius, taps = _gaussian_taps(p["gaussian_sigma"])
    s.gauss_radius = radius
    for i, t in enumerate(taps):
        s.gauss_taps[i] = float(t)
    s.max_peaks = int(max_peaks or os.environ.get("OPB_MAX_PEAKS", 8192))
    s.max_candidates = int(max_candidates or os.environ.get("OPB_MAX_CANDIDATES", 32768))
    s.max_persons = int(max_persons or os.environ.get("OPB_MAX_PERSONS", 1024))
    return s


# "comp" (compensated: fp16 product + 8-bit-float rounding corrections, map error ~1e-4) is the fastest precision inside
# the 1e-3 map tolerance; "parity" (split fp16, ~2e-5) the most accurate; "fast" (plain fp16, ~3e-3) is outside it.
_PRECISIONS = {"parity": _native.PRECISION_PARITY, "fast": _native.PRECISION_FAST, "comp": _native.PRECISION_COMP,
               _native.PRECISION_PARITY: _native.PRECISION_PARITY, _native.PRECISION_FAST: _native.PRECISION_FAST,
               _native.PRECISION_COMP: _native.PRECISION_COMP}


class PoseDetector(object):
    def __init__(self, arch=None, weights_file=None, model=None, device=-1, precise=False, precision=None,
                 max_peaks=None, max_candidates=None, max_persons=None, device_cubic=None):
        self.arch = arch
        self.precise = precise
        # precise path: run the per-scale uint8 INTER_CUBIC resize (:443) on the device as well.  Off by default: the host
        # cv2 call is what the reference executes (IPP builds differ from OpenCV's own path by 1 LSB on 4-8 % of the pixels)
        self.device_cubic = bool(int(os.environ.get("OPB_DEVICE_CUBIC", "0"))) if device_cubic is None else bool(device_cubic)
        if model is not None:
            self.model = model
        else:
            print('Loading the model...')
            self.model = params['archs'][arch]()          # KeyError for an unknown arch, like the reference
            if weights_file:
                self.model.load_npz(weights_file)
        self.device = device
        precision = precision if precision is not None else os.environ.get("OPB_PRECISION", "comp")
        self.engine = _native.Engine(max(int(device), 0), make_opb_params(params, max_peaks, max_candidates,
                                                                           max_persons), _PRECISIONS[precision])
        self.engine.load_model(self.model)
        if isinstance(self.model, CocoPoseNet):
            self.model._engine = self.engine
        self.gaussian_kernel = self.create_gaussian_kernel(params['gaussian_sigma'], params['ksize'])[None, None]

    # ------------------------------------------------------------------ small host helpers
    def create_gaussian_kernel(self, sigma=1, ksize=5):
        """2-D Gaussian of the reference's GPU branch (pose_detector.py:38-44); kept for API
        compatibility -- peak extraction here follows the CPU branch semantics."""
        c = ksize // 2
        ax = np.arange(ksize) - c
        d2 = ax[None, :] ** 2 + ax[:, None] ** 2
        return (np.exp(-0.5 * d2 / sigma ** 2) / (2 * np.pi * sigma ** 2)).astype('f')

    def pad_image(self, img, stride, pad_value):
        h, w = img.shape[:2]
        pad = [(-h) % stride, (-w) % stride]                 # [down, right]
        canvas = np.zeros((h + pad[0], w + pad[1], 3), 'uint8') + pad_value
        canvas[:h, :w] = img
        return canvas, pad

    def compute_optimal_size(self, orig_img, img_size, stride=8):
        """Short side = img_size; long side rounded half-to-even, then up to a multiple of
        stride (pose_detector.py:57-73).  Returns (w, h)."""
        h0, w0 = orig_img.shape[:2]
        aspect = h0 / w0
        landscape = h0 < w0
        long_side = int(np.round(img_size / aspect)) if landscape else int(np.round(img_size * aspect))
        long_side += (-long_side) % stride
        return (long_side, img_size) if landscape else (img_size, long_side)

    def preprocess(self, img):
        x = img.astype('f')
        x /= 255
        x -= 0.5
        return x.transpose(2, 0, 1)[None]

    # ------------------------------------------------------------------ stage methods (device)
    def compute_peaks_from_heatmaps(self, heatmaps):
        """all_peaks: [N,5] float64 rows (jointtype, x, y, score, index); np.array([]) if none."""
        peaks = self.engine.peaks(np.asarray(heatmaps))
        return peaks if len(peaks) else np.array([])

    def compute_connections(self, pafs, all_peaks, img_len, params):
        return self.engine.connections(np.asarray(pafs), all_peaks, img_len)

    def compute_candidate_connections(self, paf, cand_a, cand_b, img_len, params):
        """Sorted candidate list of ONE limb, [[id_a, id_b, score], ...] (pose_detector.py:135-159);
        the PAF line integrals run on the device (opb_candidates)."""
        rows = self.engine.candidates(np.asarray(paf), cand_a, cand_b, img_len)
        return [[int(r[0]), int(r[1]), r[2]] for r in rows]

    def grouping_key_points(self, all_connections, candidate_peaks, params):
        return self.engine.group(all_connections, candidate_peaks)

    def subsets_to_pose_array(self, subsets, all_peaks):
        people = []
        for row in subsets:
            ids = row[:18].astype('i')
            people.append(np.array([[all_peaks[j][1], all_peaks[j][2], 2] if j >= 0 else [0, 0, 0] for j in ids]))
        return np.array(people)

    # ------------------------------------------------------------------ body-scale / crop helpers
    def compute_limbs_length(self, joints):
        limbs, lens = [], np.zeros(len(params["limbs_point"]))
        for i, (ja, jb) in enumerate(params["limbs_point"]):
            a, b = joints[ja], joints[jb]
            if a is not None and b is not None:
                limbs.append([a, b])
                lens[i] = np.linalg.norm(b[:-1] - a[:-1])
            else:
                limbs.append(None)
        return lens, limbs

    def compute_unit_length(self, limbs_len):
        # nose-neck, neck-Lwaist, neck-Rwaist, shoulder-Lear, shoulder-Rear take priority
        base = limbs_len[[14, 3, 0, 13, 9]]
        present = base > 0
        if present.any():
            ratio = np.array([0.85, 2.2, 2.2, 0.85, 0.85])
            return np.sum(base[present] / ratio[present]) / np.count_nonzero(present)
        ratio = np.array([2.2, 1.7, 1.7, 2.2, 1.7, 1.7, 0.6, 0.93, 0.65, 0.85, 0.6, 0.93, 0.65, 0.85, 1, 0.2, 0.2,
                          0.25, 0.25])
        present = limbs_len > 0
        return np.sum(limbs_len[present] / ratio[present]) / np.count_nonzero(present)

    def get_unit_length(self, person_pose):
        return self.compute_unit_length(self.compute_limbs_length(person_pose)[0])

    def crop_image(self, img, bbox):
        left, top, right, bottom = bbox
        ih, iw, ic = img.shape
        out = np.zeros((bottom - top, right - left, ic), dtype=np.uint8)
        l, t, r, b = max(0, left), max(0, top), min(iw, right), min(ih, bottom)
        ox, oy = max(0, -left), max(0, -top)
        out[oy:oy + (b - t), ox:ox + (r - l)] = img[t:b, l:r]
        return out

    def crop_around_keypoint(self, img, keypoint, crop_size):
        x, y = keypoint
        bbox = (int(x - crop_size), int(y - crop_size), int(x + crop_size), int(y + crop_size))
        return self.crop_image(img, bbox), bbox

    # joint -> rank tables of crop_person (pose_detector.py:312-313; lower rank = preferred anchor) and the padding,
    # in unit lengths, above / below the anchoring joint (:343-344)
    _TOP_RANK = (4, 5, 6, 12, 16, 7, 13, 17, 8, 10, 14, 9, 11, 15, 2, 3, 0, 1)
    _BOTTOM_RANK = (9, 6, 7, 14, 16, 8, 15, 17, 4, 2, 0, 5, 3, 1, 10, 11, 12, 13)
    _TOP_PAD = (0.9, 1.9, 1.9, 2.9, 3.7, 1.9, 2.9, 3.7, 4.0, 5.5, 7.0, 4.0, 5.5, 7.0, 0.7, 0.8, 0.7, 0.8)
    _BOTTOM_PAD = (6.9, 5.9, 5.9, 4.9, 4.1, 5.9, 4.9, 4.1, 3.8, 2.3, 0.8, 3.8, 2.3, 0.8, 7.1, 7.0, 7.1, 7.0)

    def crop_person(self, img, person_pose, unit_length):
        """pose_detector.py:311-352 (the reference's own version raises NameError: it uses `sys` without importing
        it, :1-12).  Same walk over the joints, including its either/or updates: a joint that improves the top anchor
        (or the top / left extent) is not considered for the bottom anchor (bottom / right extent) in the same step."""
        import sys
        top_rank, bottom_rank = sys.maxsize, sys.maxsize
        top_joint = bottom_joint = len(self._TOP_RANK)          # "none": index of the sentinel entry
        top_pos = left_pos = sys.maxsize
        bottom_pos = right_pos = 0
        for i, joint in enumerate(person_pose):
            if not joint[2] > 0:
                continue
            if self._TOP_RANK[i] < top_rank:
                top_rank, top_joint = self._TOP_RANK[i], i
            elif self._BOTTOM_RANK[i] < bottom_rank:
                bottom_rank, bottom_joint = self._BOTTOM_RANK[i], i
            if joint[1] < top_pos:
                top_pos = joint[1]
            elif joint[1] > bottom_pos:
                bottom_pos = joint[1]
            if joint[0] < left_pos:
                left_pos = joint[0]
            elif joint[0] > right_pos:
                right_pos = joint[0]
        # (IndexError, like the reference, when no anchor joint was found)
        bbox = (int(left_pos - 0.3 * unit_length), int(top_pos - self._TOP_PAD[top_joint] * unit_length),
                int(right_pos + 0.3 * unit_length), int(bottom_pos + self._BOTTOM_PAD[bottom_joint] * unit_length))
        return self.crop_image(img, bbox), bbox

    def crop_face(self, img, person_pose, unit_length):
        nose = person_pose[JointType.Nose]
        if not nose[2] > 0:
            return None, None
        nx, ny = nose[:2]
        bbox = (int(nx - unit_length), int(ny - unit_length * 1.2), int(nx + unit_length), int(ny + unit_length * 0.8))
        return self.crop_image(img, bbox), bbox

    def crop_hands(self, img, person_pose, unit_length):
        hands = {"left": None, "right": None}
        for side, hand_j, elbow_j in (("left", JointType.LeftHand, JointType.LeftElbow),
                                      ("right", JointType.RightHand, JointType.RightElbow)):
            if person_pose[hand_j][2] > 0:
                center = person_pose[hand_j][:-1]
                if person_pose[elbow_j][2] > 0:
                    direction = person_pose[hand_j][:-1] - person_pose[elbow_j][:-1]
                    center += (0.3 * direction).astype(center.dtype)
                hand_img, bbox = self.crop_around_keypoint(img, center, unit_length * 0.95)
                hands[side] = {"img": hand_img, "bbox": bbox}
        return hands

    # ------------------------------------------------------------------ result assembly
    def _poses_from_records(self, header, persons, sx, sy):
        """header/persons of ONE image -> (poses, scores) with the reference's shapes."""
        self.engine.raise_for_status(int(header["status"]))
        if int(header["n_peaks"]) == 0:
            return np.empty((0, len(JointType), 3)), np.empty(0)       # pose_detector.py:509-510
        n = int(header["n_persons"])
        rec = persons[:n]
        scores = rec["score"].astype(np.float64).copy()
        if n == 0:
            return np.array([]), scores                                 # shape (0,), pose_detector.py:264
        has = rec["peak_id"] >= 0
        poses = np.zeros((n, len(JointType), 3), np.float64)
        poses[..., 0] = np.where(has, rec["x"].astype(np.float64) * sx, 0.0)   # x *= orig_w/map_w  (:513)
        poses[..., 1] = np.where(has, rec["y"].astype(np.float64) * sy, 0.0)
        poses[..., 2] = np.where(has, 2.0, 0.0)
        return poses, scores

    def detect_precise(self, orig_img):
        """Multi-scale path (pose_detector.py:433-482): per scale the uint8 image is resized on the host with
        cv2 INTER_CUBIC exactly as the reference does (this OpenCV build dispatches 8-bit cubic to IPP, whose arithmetic
        is unpublished) -- or, with device_cubic=True / OPB_DEVICE_CUBIC=1, on the device with OpenCV's own
        8-bit cubic arithmetic (bit-exact with cv2 when IPP is off); padding, forward, both cubic map resizes, averaging
        and the whole post-process run on the device."""
        oh, ow = orig_img.shape[:2]
        scales = params['inference_scales']
        self.engine.precise_begin(oh, ow)
        for k, scale in enumerate(scales):
            m = scale * params['inference_img_size'] / min(orig_img.shape[:2])
            rw, rh = math.ceil(ow * m), math.ceil(oh * m)
            if self.device_cubic:
                # the uint8 INTER_CUBIC resize of :443 on the device too (OpenCV's own 8-bit path, csrc/ingest.cuh)
                self.engine.precise_add_scale_orig(orig_img, rh, rw, params['downscale'], (104, 117, 123), k, len(scales))
                continue
            img = cv2.resize(orig_img, (rw, rh), interpolation=cv2.INTER_CUBIC)
            # pad_image(img, 8, (104, 117, 123)) of :445 runs on the device (csrc/ingest.cuh)
            self.engine.precise_add_scale_unpadded(img, params['downscale'], (104, 117, 123), k, len(scales))
        header, persons = self.engine.precise_finish(ow)
        self.pafs, self.heatmaps = self.engine.download_maps(oh, ow)
        self.engine.raise_for_status(int(header[0]["status"]))
        self.all_peaks = self.engine.image_detail(0)[0] if header[0]["n_peaks"] else np.array([])
        self._last_scale = (1.0, 1.0)
        return self._poses_from_records(header[0], persons[0], 1.0, 1.0)

    def __call__(self, orig_img):
        orig_img = orig_img.copy()
        if self.precise:
            return self.detect_precise(orig_img)
        oh, ow = orig_img.shape[:2]
        in_w, in_h = self.compute_optimal_size(orig_img, params['inference_img_size'])
        map_w, map_h = self.compute_optimal_size(orig_img, params['heatmap_size'])
        # cv2.resize(orig_img, (input_w, input_h)) of the reference (:493) runs on the device, bit-exact with
        # OpenCV's 8-bit INTER_LINEAR (csrc/ingest.cuh); the frame is uploaded once at its original size
        headers, persons = self.engine.detect_image(orig_img, in_h, in_w, map_h, map_w, img_len=map_w)
        self._last_scale = (ow / map_w, oh / map_h)
        return self._poses_from_records(headers[0], persons[0], ow / map_w, oh / map_h)

    def draw_last_result(self, orig_img):
        """draw_person_pose(orig_img, poses) for the frame just passed to __call__, with the poses taken from the
        device-resident result records (the per-frame drawing of camera_pose_demo.py:27 without the host round trip)."""
        sx, sy = self._last_scale
        return self.engine.draw_last_result(orig_img, sx, sy)

    def detect_batch(self, imgs, orig_sizes=None):
        """Batched fast path for equally sized BGR frames [N,H,W,3] (no reference analogue: the
        reference is batch-1).  Returns a list of (poses, scores)."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        in_w, in_h = self.compute_optimal_size(imgs[0], params['inference_img_size'])
        map_w, map_h = self.compute_optimal_size(imgs[0], params['heatmap_size'])
        oh, ow = imgs.shape[1:3]
        # the per-frame cv2.resize of :493 runs on the device for the whole batch (csrc/ingest.cuh)
        self.engine.stream_submit(imgs, in_h, in_w, map_h, map_w, slot=0, img_len=map_w)
        headers, persons = self.engine.stream_collect(0)
        return [self._poses_from_records(headers[i], persons[i], ow / map_w, oh / map_h) for i in range(len(imgs))]

    def detect_stream(self, frames):
        """Pipelined camera loop (camera_pose_demo.py:20-31): `frames` is an iterable of BGR uint8 frames
        [H,W,3] or equally sized batches [N,H,W,3]; yields (poses, scores) per frame (a list of them per batch)
        in order, one item behind the input: the upload of item i+1 overlaps the kernels of item i."""
        pending = None                                    # (slot, single, ow, oh, map_w, map_h, n)
        slot = 0
        try:
            for item in frames:
                a = np.ascontiguousarray(item, np.uint8)
                single = a.ndim == 3
                if single:
                    a = a[None]
                in_w, in_h = self.compute_optimal_size(a[0], params['inference_img_size'])
                map_w, map_h = self.compute_optimal_size(a[0], params['heatmap_size'])
                self.engine.stream_submit(a, in_h, in_w, map_h, map_w, slot=slot, img_len=map_w)
                cur = (slot, single, a.shape[2], a.shape[1], map_w, map_h, len(a))
                prev, pending, slot = pending, cur, slot ^ 1
                if prev is not None:
                    yield self._collect_stream(prev)
            if pending is not None:
                last, pending = pending, None
                yield self._collect_stream(last)
        finally:
            # the consumer stopped early (break / exception / generator closed): a submitted batch must not stay
            # "busy" in the C context, or the next submit on that slot would fail for the lifetime of the engine
            if pending is not None:
                try:
                    self.engine.stream_collect(pending[0])
                except Exception:
                    pass

    def _collect_stream(self, pending):
        slot, single, ow, oh, map_w, map_h, n = pending
        headers, persons = self.engine.stream_collect(slot)
        out = [self._poses_from_records(headers[i], persons[i], ow / map_w, oh / map_h) for i in range(n)]
        return out[0] if single else out


_LIMB_COLORS = [
    [0, 255, 0], [0, 255, 85], [0, 255, 170], [0, 255, 255], [0, 170, 255], [0, 85, 255], [255, 0, 0],
    [255, 85, 0], [255, 170, 0], [255, 255, 0.], [255, 0, 85], [170, 255, 0], [85, 255, 0], [170, 0, 255.],
    [0, 0, 255], [0, 0, 255], [255, 0, 255], [170, 0, 255], [255, 0, 170]]
_JOINT_COLORS = [
    [255, 0, 0], [255, 85, 0], [255, 170, 0], [255, 255, 0], [170, 255, 0], [85, 255, 0], [0, 255, 0],
    [0, 255, 85], [0, 255, 170], [0, 255, 255], [0, 170, 255], [0, 85, 255], [0, 0, 255], [85, 0, 255],
    [170, 0, 255], [255, 0, 255], [255, 0, 170], [255, 0, 85]]


def draw_person_pose(orig_img, poses, engine=None):
    """Skeleton overlay (pose_detector.py:520-553): limbs first (ear-shoulder limbs 9 and 13 are
    not drawn), then joints.  With `engine` (a detector's .engine) the overlay is rasterised on the device
    (opb_draw_person_pose: OpenCV's thick-line / filled-circle arithmetic restated, pixel-identical; csrc/overlay.cuh)."""
    if len(poses) == 0:
        return orig_img
    if engine is not None:
        return engine.draw_person_pose(orig_img, poses)
    canvas = orig_img.copy()
    int_poses = poses.round().astype('i')
    for pose in int_poses:
        for i, ((ja, jb), color) in enumerate(zip(params['limbs_point'], _LIMB_COLORS)):
            if i in (9, 13):
                continue
            if pose[ja][2] != 0 and pose[jb][2] != 0:
                cv2.line(canvas, tuple(pose[ja][:2]), tuple(pose[jb][:2]), color, 2)
    for pose in int_poses:
        for (x, y, v), color in zip(pose, _JOINT_COLORS):
            if v != 0:
                cv2.circle(canvas, (x, y), 3, color, -1)
    return canvas


if __name__ == '__main__':
    import argparse
    parser = argparse.ArgumentParser(description='Pose detector')
    parser.add_argument('arch', choices=params['archs'].keys(), default='posenet', help='Model architecture')
    parser.add_argument('weights', help='weights file path')
    parser.add_argument('--img', '-i', default=None, help='image file path')
    parser.add_argument('--gpu', '-g', type=int, default=-1, help='GPU ID (negative selects GPU 0)')
    parser.add_argument('--precise', action='store_true', help='do precise inference')
    args = parser.parse_args()
    pose_detector = PoseDetector(args.arch, args.weights, device=args.gpu, precise=args.precise)
    img = cv2.imread(args.img)
    poses, _ = pose_detector(img)
    img = draw_person_pose(img, poses)
    print('Saving result into result.png...')
    cv2.imwrite('result.png', img)
