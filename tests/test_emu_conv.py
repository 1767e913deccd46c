"""CPU (`-m "not gpu"`): the tcgen05 / TMA convolution kernels themselves, run WITHOUT a GPU.

tests/cuda_emu/ptx_emu.cuh replaces the inline-PTX wrappers of csrc/ptx.cuh by a functional model of what the kernels
use of sm_100a -- mbarrier phases and transaction counts, TMA tiled loads with out-of-bounds zero fill and the
128-byte swizzle, UMMA shared-memory descriptors, tcgen05.mma into TMEM (queued until the issuing thread commits, i.e.
executed as late as the kernel's own synchronisation allows), tcgen05.ld lane quadrants -- so the unmodified kernel
sources (descriptor arithmetic, swizzled operand layouts, pipeline phase bookkeeping, role-swapped and fused variants,
epilogues, max-pool fusion, split-fp16 two-level accumulation) execute on the CPU and are compared with a torch fp32
conv2d / the oracle's forward pass under the SAME tolerances as tests/test_gpu_conv.py.

CTA pairs (cta_group::2: the fused 7x7 N=256 launch) run with both CTAs of a pair resident together: remote mbarrier
arrives, pair TMA crediting the leader's barriers, M=256 MMAs over both CTAs' operands / TMEM, multicast commits.
Not modelled: timing, bank conflicts, and the hardware's accumulation order inside one MMA (results agree to rounding).
Test infrastructure only: the package never loads the emulated library."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import pkg
from oracle import restate as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import build_emu  # noqa: E402
import test_gpu_conv as G  # noqa: E402  (plain helpers; the gpu mark belongs to that module only)


@pytest.fixture(scope="module")
def emu_lib():
    native = pkg("_native")
    try:
        lib = C.CDLL(build_emu.build(contract=False))
    except (RuntimeError, OSError) as e:     # no g++ / CUDA headers on this box: the harness, not the product, is missing
        pytest.skip("emulated build unavailable: %s" % str(e)[:200])
    for name, (res, args) in native._SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


@pytest.fixture()
def emu_native(emu_lib, monkeypatch):
    native = pkg("_native")
    monkeypatch.setattr(native, "_lib", emu_lib)
    # 3 "SMs": every persistent CTA loops over many tiles, so operand / accumulator stages wrap around and the phase
    # parities of all mbarrier pipelines flip repeatedly (with 148 SMs these small cases give each CTA one tile)
    monkeypatch.setenv("OPB_EMU_SMS", "3")
    return native


@pytest.fixture()
def engine(emu_native):
    return emu_native.Engine(0, pkg("pose_detector").make_opb_params())


CASES = [
    # n, h, w, cin, cout, ks, relu   (small versions of every kernel family of tests/test_gpu_conv.py)
    (1, 24, 30, 128, 128, 7, 1),     # role-swapped 7x7 kernel, W % 16 == 14
    (2, 23, 19, 128, 128, 7, 0),     # ... two images, last row tile trimmed (23 rows), 8-px edge tile
    (2, 24, 24, 185, 256, 7, 1),     # fused Mconv1 shape (Cin 185 -> 192, N = 256) on the single-CTA kernel
    (2, 23, 31, 64, 64, 3, 1),       # conv1_2-like, partial tiles in x and y
    (1, 24, 24, 256, 512, 3, 1),     # two N blocks
    (1, 24, 40, 128, 512, 1, 1),     # conv5_4-like 1x1
    (2, 30, 17, 512, 38, 1, 0),      # PAF head (N = 48 tile, 38 valid)
    (1, 24, 24, 128, 19, 1, 0),      # heat head
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_%dx%d_c%d_o%d_k%d" % c[:6])
@pytest.mark.parametrize("mode", ["fast", "parity", "comp"])
def test_conv_vs_torch(engine, emu_native, case, mode):
    n, h, w, cin, cout, ks, relu = case
    rs = np.random.RandomState(sum(case))
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    prec = {"fast": emu_native.PRECISION_FAST, "parity": emu_native.PRECISION_PARITY, "comp": emu_native.PRECISION_COMP}[mode]
    y = engine.test_conv(x, W, b, relu, prec)
    ref = G._ref_conv(x, W, b, relu, quantize=(mode == "fast"))
    scale = np.abs(ref).max()
    err = np.abs(y - ref).max()
    # comp: first-order rounding terms corrected with 3-significant-bit operands -> ~2^-3 of the fp16 error
    tol = {"fast": 2e-3, "parity": 1e-4, "comp": 3e-4}[mode] * scale
    assert err <= tol, "max abs err %.3e > tol %.3e (scale %.3f)" % (err, tol, scale)


@pytest.mark.parametrize("case", [(1, 24, 30, 128, 128, 7, 1), (1, 24, 24, 185, 256, 7, 1)], ids=lambda c: "n%d_%dx%d_c%d_o%d_k%d" % c[:6])
def test_comp_small_batch_two_level_accumulation(emu_native, monkeypatch, case):
    """compensated precision on a launch that fills less than half a wave (one camera frame): add_conv then picks the plain
    kernel's two-level accumulation variants (BN <= 128, MT = 1) with the 8-bit correction rows and the three-plane
    epilogue -- the path PoseDetector.__call__ takes for a single frame."""
    monkeypatch.setenv("OPB_EMU_SMS", "148")
    n, h, w, cin, cout, ks, relu = case
    rs = np.random.RandomState(sum(case) + 1)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params())
    y = eng.test_conv(x, W, b, relu, emu_native.PRECISION_COMP)
    ref = G._ref_conv(x, W, b, relu, quantize=False)
    assert np.abs(y - ref).max() <= 3e-4 * np.abs(ref).max()


@pytest.mark.parametrize("mode", ["fast", "comp"])
def test_lean_issue_swap7_kernel_bit_identical_to_swap_kernel(emu_native, monkeypatch, mode):
    """conv_tcgen05_swap7_kernel (7-stage weight ring = one filter column, compile-time stage addresses, elect.sync issue;
    the default for the 7x7 128->128 layers) issues the same MMAs in the same order as conv_tcgen05_swap_kernel
    (OPB_SWAP7=0): identical bits, incl. the 8-pixel edge tile, the trimmed last tile row and ring wrap-around."""
    case = (2, 23, 19, 128, 128, 7, 1)
    n, h, w, cin, cout, ks, relu = case
    rs = np.random.RandomState(7)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    prec = {"fast": emu_native.PRECISION_FAST, "comp": emu_native.PRECISION_COMP}[mode]
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("OPB_SWAP7", flag)
        eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params())
        out.append(eng.test_conv(x, W, b, relu, prec))
    assert np.array_equal(out[0], out[1])


def _emu_stats(lib):
    st = (C.c_longlong * 4)()
    lib.opb_emu_stats(st)
    return dict(launches=st[0], cluster_launches=st[1], mma=st[2], tma=st[3])


@pytest.mark.parametrize("case", [(2, 24, 24, 185, 256, 7, 1), (1, 25, 28, 128, 256, 7, 1), (1, 24, 24, 256, 256, 3, 1),
                                  (1, 24, 40, 128, 512, 1, 1)], ids=lambda c: "n%d_%dx%d_c%d_o%d_k%d" % c[:6])
def test_cta_pair_kernel(emu_native, emu_lib, monkeypatch, case):
    """conv_tcgen05_pair_kernel (cta_group::2) for every kernel family (OPB_PAIR=7; by default only the fused 7x7
    N=256 launch uses it), fast precision; the emulator's counters prove that the launch really ran as CTA pairs."""
    monkeypatch.setenv("OPB_PAIR", "7")
    n, h, w, cin, cout, ks, relu = case
    rs = np.random.RandomState(sum(case))
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params())
    before = _emu_stats(emu_lib)
    y = eng.test_conv(x, W, b, relu, emu_native.PRECISION_FAST)
    after = _emu_stats(emu_lib)
    assert after["cluster_launches"] == before["cluster_launches"] + 1 and after["mma"] > before["mma"] and after["tma"] > before["tma"]
    ref = G._ref_conv(x, W, b, relu, quantize=True)
    assert np.abs(y - ref).max() <= 2e-3 * np.abs(ref).max()


def test_conv_zero_padding_borders(engine):
    G.test_conv_zero_padding_borders(engine)


@pytest.mark.parametrize("mode", ["fast", "parity", "comp"])
def test_conv_fused_maxpool(engine, mode):
    G.test_conv_fused_maxpool(engine, (2, 24, 40, 64, 64, 3), mode)


@pytest.mark.parametrize("mode,tol", [("parity", 1e-4), ("comp", 5e-4)])   # fast precision: test_pose_detector_call_end_to_end
def test_whole_network_forward(emu_native, he_weights, mode, tol):
    """All 92 convolutions of CocoPoseNet (3 fused max-pools, concat-by-slice, fused 1x1 pairs, conv1_1 on tensor
    cores in fast precision / the split-fp16 DRAIN kernels in parity precision) on a 176x128 frame -- the smallest the
    7x7 TMA boxes admit -- against the oracle's torch-CPU fp32 forward.  north_star's tolerance is 1e-3 (parity)."""
    syn = pkg("synthetic")
    model = pkg("models.CocoPoseNet").CocoPoseNet()
    model.load_npz(syn.he_weights(0))
    img = syn.procedural_image(176, 128, seed=4)
    ref_paf, ref_heat = R.forward(he_weights, R.preprocess(img))
    eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params(),
                            {"fast": emu_native.PRECISION_FAST, "parity": emu_native.PRECISION_PARITY, "comp": emu_native.PRECISION_COMP}[mode])
    eng.load_model(model)
    paf, heat = eng.forward(img[None])
    err = max(float(np.abs(paf[0] - ref_paf[0]).max()), float(np.abs(heat[0] - ref_heat[0]).max()))
    assert err <= tol, err


def test_fused_1x1_pair_compensated_bit_identical_to_two_launches(emu_native, monkeypatch):
    """conv_mlp2_kernel<COMP> (Mconv6 + Mconv7 of both branches in one launch, the intermediate's fp16 values AND its 8-bit
    correction bytes written straight into the swizzled operand tile) vs the two separate 1x1 launches (OPB_NO_MLP2=1) in
    compensated precision: same MMAs in the same order on the same operand bits -> identical network outputs."""
    syn = pkg("synthetic")
    model = pkg("models.CocoPoseNet").CocoPoseNet()
    model.load_npz(syn.he_weights(0))
    img = syn.procedural_image(176, 128, seed=6)
    out = []
    monkeypatch.setenv("OPB_MLP2_COMP", "1")   # opt-in variant
    for no_fuse in ("1", "0"):
        monkeypatch.setenv("OPB_NO_MLP2", no_fuse)
        eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params(), emu_native.PRECISION_COMP)
        eng.load_model(model)
        out.append(eng.forward(img[None]))
        del eng
    assert np.isfinite(out[0][0]).all() and np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_pose_detector_call_end_to_end(emu_native, monkeypatch):
    """PoseDetector.__call__ (pose_detector.py:484-517) entirely under emulation: upload -> device cv2-exact resize ->
    92-conv chain (CTA-pair kernel included) -> upsample -> peaks -> PAF integrals -> assignment -> grouping -> records
    -> host rescale.  The network runs at a reduced inference size (176 instead of 368; the emulated tensor core does
    ~4 GMAC/s) and in fast precision; the reference result is the oracle's post-process applied to the maps the same
    engine returns from opb_forward for the same resized frame, so every stage after the network must agree bit for
    bit, whatever fp16 rounding did to the maps."""
    import cv2
    entity, PD, syn = pkg("entity"), pkg("pose_detector"), pkg("synthetic")
    monkeypatch.setitem(entity.params, "inference_img_size", 176)
    monkeypatch.setitem(entity.params, "heatmap_size", 160)
    monkeypatch.setenv("OPB_GRAPH", "0")
    model = pkg("models.CocoPoseNet").CocoPoseNet()
    model.load_npz(syn.he_weights(0))
    det = PD.PoseDetector(model=model, device=0, precision="fast")
    img = syn.procedural_image(150, 205, seed=9)                       # short side -> 176 (network input), 160 (maps)
    poses, scores = det(img)
    in_w, in_h = det.compute_optimal_size(img, 176)
    map_w, map_h = det.compute_optimal_size(img, 160)
    assert in_h == 176 and in_w % 8 == 0 and map_h == 160 and map_w % 8 == 0
    paf_lo, heat_lo = det.engine.forward(cv2.resize(img, (in_w, in_h))[None])
    pafs = R.resize_bilinear_align_corners(paf_lo, (map_h, map_w))[0]
    heat = R.resize_bilinear_align_corners(heat_lo, (map_h, map_w))[0]
    ref_poses, ref_scores = R.postprocess_fast(pafs, heat, map_w, img.shape[1], img.shape[0], map_h)
    assert len(scores) > 0, "random-weight maps at this size give some 'persons'; an empty result tests nothing"
    assert np.array_equal(scores, ref_scores) and np.array_equal(poses, ref_poses)
    # the overlay of the frame just processed, with the poses taken from the device-resident records
    # (opb_draw_last_result), against the reference's cv2 drawing loop on the returned poses
    assert np.array_equal(det.draw_last_result(img), PD.draw_person_pose(img, poses))


@pytest.mark.parametrize("modname,cls,n_kp", [("models.FaceNet", "FaceNet", 70)])   # HandNet: same chain builder, checked by hand
def test_face_and_hand_detectors_end_to_end(emu_native, modname, cls, n_kp):
    """opb_keypoints_detect (face_detector.py:28-67 / hand_detector.py:28-77) under emulation: device cv2-exact resize of
    the crop, the FaceNet / HandNet chain (7x7 (128+C)->128 concat layers, role-swapped 7x7, 1x1 pairs), F.resize_images to
    the crop size, exact Gaussian passes and arg-max.  Network size reduced to 176 (emulation speed), fast precision: the
    maps agree with the oracle's fp32 forward to fp16 accuracy, and the keypoints are exactly what the oracle's peak
    finder returns for the device's own maps."""
    syn = pkg("synthetic")
    nm = pkg(modname)
    wd = syn.he_weights(0, layers=nm.LAYERS)
    net = getattr(nm, cls)()
    net.load_npz(wd)
    eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params(), emu_native.PRECISION_FAST)
    eng.load_model(net)
    crop = syn.procedural_image(120, 140, seed=21)
    kps, maps = eng.keypoints_detect(crop, 176, 0.1, mirror=False, return_maps=True)
    assert len(kps) == n_kp and sum(k is not None for k in kps) > 0
    weights = {n: (wd[n + "/W"], wd[n + "/b"]) for n, _, _, _ in nm.LAYERS}
    lo = R.keypoint_forward(weights, R.keypoint_preprocess(R.cv2_resize_linear_u8(np.ascontiguousarray(crop), (176, 176))))
    ref_maps = R.resize_bilinear_align_corners(lo, crop.shape[:2])[0]
    assert np.abs(maps - ref_maps[:-1]).max() <= 1e-2
    full = np.concatenate([maps, np.zeros((1,) + maps.shape[1:], np.float32)])
    for a, b in zip(kps, R.keypoints_from_heatmaps(full, 0.1)):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a[0], a[1]) == (b[0], b[1]) and np.float32(a[2]) == np.float32(b[2])
