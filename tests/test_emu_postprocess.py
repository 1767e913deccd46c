"""CPU (`-m "not gpu"`): the library's own plain-CUDA kernels and C-ABI host code, run on a
box WITHOUT a GPU through tests/cuda_emu (CUDA threads as fibers, runtime calls on host
memory; test infrastructure only -- the package never loads it), checked bit-for-bit
against the oracle.

What this covers that the other CPU tests cannot: the arithmetic and control flow of the
device code itself (upsample, smooth + NMS + sort, PAF line integrals, limb assignment,
person grouping, uint8 resize, keypoint arg-max) and the host orchestration behind
opb_upsample / opb_peaks / opb_connections / opb_candidates / opb_group /
opb_resize_linear_u8 / opb_keypoints_from_heatmaps.  The emulator also aborts on
warp-synchronous code that diverges when lanes are not executed in lockstep (it runs the lanes
of a warp one after another between synchronisation points), i.e. on races that independent
thread scheduling is allowed to expose.

What it cannot cover: memory coalescing, occupancy, timing (the tcgen05 / TMA convolution kernels run under a
functional model of their PTX: tests/test_emu_conv.py).

The same test bodies as tests/test_gpu_postprocess.py are reused with the emulated engine;
every case runs against two builds of the emulated library: -ffp-contract=off and
-ffp-contract=fast -mfma (g++ then fuses any a*b+c not written as an explicit _rn intrinsic,
like nvcc does by default) -- results must be identical, which shows that no bit-exact
result depends on the compiler's contraction choices.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import load_golden, pkg
from oracle import restate as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import build_emu  # noqa: E402
import test_gpu_postprocess as G  # noqa: E402  (plain functions; the gpu mark belongs to that module only)
from postprocess_batch_cases import check_batch_against_oracle, run_batch_cases  # noqa: E402


def _load(contract):
    native = pkg("_native")
    try:
        lib = C.CDLL(build_emu.build(contract=contract))
    except (RuntimeError, OSError) as e:     # no g++ / CUDA headers on this box: the harness, not the product, is missing
        pytest.skip("emulated build unavailable: %s" % str(e)[:200])
    for name, (res, args) in native._SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


@pytest.fixture(scope="module", params=["nofma", "fma"])
def emu_lib(request):
    return _load(request.param == "fma")


# the FMA-contracting build repeats the arithmetic-heavy cases only (suite time); everything else runs once
_FMA_CASES = ("test_upsample_bilinear_bit_exact", "test_full_postprocess_synth8", "test_postprocess_batch_from_network_resolution_maps",
              "test_device_resize_linear_u8_bit_exact_vs_cv2", "test_device_resize_cubic_u8_bit_exact_vs_cv2", "test_device_overlay_pixel_identical_to_cv2",
              "test_keypoints_exact_ties_and_threshold",
              "test_upsample_bicubic_vs_cv2", "test_candidate_connections_single_limb", "test_emulated_library_is_not_the_product")


@pytest.fixture(autouse=True)
def _fma_build_only_where_it_matters(request):
    cs = getattr(request.node, "callspec", None)
    if cs is not None and cs.params.get("emu_lib") == "fma" and request.node.originalname not in _FMA_CASES:
        pytest.skip("one build is enough for this case (suite time)")


@pytest.fixture()
def emu_native(emu_lib, monkeypatch):
    """For the duration of one test, Engine() binds the emulated library instead of libopb.so."""
    native = pkg("_native")
    monkeypatch.setattr(native, "_lib", emu_lib)
    return native


@pytest.fixture()
def engine(emu_native):
    prm = pkg("pose_detector").make_opb_params(max_peaks=16384, max_candidates=131072, max_persons=4096)
    return emu_native.Engine(0, prm)


def test_emulated_library_is_not_the_product(emu_lib):
    native = pkg("_native")
    assert "cuda_emu" in build_emu.lib_path(False) and not native.LIB_PATH.startswith(build_emu.BUILD)
    assert emu_lib.opb_version() == 1


def test_upsample_bilinear_bit_exact(engine):
    G.test_upsample_bilinear_bit_exact(engine)


@pytest.mark.parametrize("seed", [0, 1])
def test_full_postprocess_synth8(engine, seed):
    """upsample -> peaks -> connections -> grouping on the synthetic 8-person maps, every stage
    against the golden produced by the reference's own code."""
    g = load_golden("synth8_post_seed%d.npz" % seed)
    paf, heat, _ = pkg("synthetic").eight_person_maps(seed=seed)
    peaks = engine.peaks(heat)
    assert np.array_equal(peaks, g["all_peaks"])
    conns = engine.connections(paf, peaks, 576)
    G._check_conns(conns, G.split_conns(g["conn_lens"], g["conn_flat"]))
    subsets = engine.group(conns, peaks)
    assert np.array_equal(subsets, g["subsets"]) and len(subsets) == 8


def test_peaks_random_maps_and_edges(engine):
    rs = np.random.RandomState(3)
    for (h, w) in ((64, 96), (37, 53), (16, 64), (17, 65), (9, 7), (130, 21)):
        heat = (rs.standard_normal((19, h, w)) * 0.3).astype(np.float32)
        G._peaks_case(engine, heat)
    heat = np.zeros((19, 40, 40), np.float32)
    assert len(engine.peaks(heat)) == 0
    heat[:] = 1.0
    G._peaks_case(engine, heat)
    heat = np.zeros((19, 40, 40), np.float32)
    heat[3, 20, 20] = 5.0
    got = G._peaks_case(engine, heat)
    assert got.shape == (1, 5) and tuple(got[0, :3]) == (3.0, 20.0, 20.0)
    heat[3, 0, 0] = 9.0
    G._peaks_case(engine, heat)


def test_peaks_on_oracle_upsampled_network_maps(engine):
    G.test_peaks_on_oracle_upsampled_network_maps(engine)


def test_connections_webcam_shape(engine):
    G.test_connections_webcam_shape(engine)


def test_connections_empty_and_degenerate(engine):
    G.test_connections_empty_and_degenerate(engine)


def test_grouping_quirks_random_graphs(engine):
    G.test_grouping_quirks_random_graphs(engine)


def test_grouping_third_match_raises_index_error(engine):
    G.test_grouping_third_match_raises_index_error(engine)


def test_public_stage_methods_chain(engine):
    G.test_public_stage_methods_chain(engine)


def test_candidate_connections_single_limb(engine):
    G.test_candidate_connections_single_limb(engine)


def test_capacity_errors_are_loud(emu_native):
    G.test_capacity_errors_are_loud()


def test_upsample_bicubic_vs_cv2(engine):
    import cv2
    native = pkg("_native")
    rs = np.random.RandomState(5)
    for (h, w, H, W) in ((23, 23, 184, 184), (31, 17, 100, 90)):
        x = rs.standard_normal((3, h, w)).astype(np.float32)
        got = engine.upsample(x, H, W, mode=native.UPSAMPLE_BICUBIC)
        ref = cv2.resize(np.ascontiguousarray(x.transpose(1, 2, 0)), (W, H), interpolation=cv2.INTER_CUBIC).transpose(2, 0, 1)
        assert np.abs(got - ref).max() <= 1e-5


def test_device_resize_linear_u8_bit_exact_vs_cv2(engine):
    import cv2
    rs = np.random.RandomState(0)
    shapes = [((120, 160), (124, 92)), ((146, 146), (92, 92)), ((100, 37), (92, 250)), ((92, 164), (164, 92)), ((83, 129), (144, 92))]
    for _ in range(4):
        shapes.append(((rs.randint(20, 150), rs.randint(20, 150)), (rs.randint(20, 130), rs.randint(20, 130))))
    for (h0, w0), (W, H) in shapes:
        img = rs.randint(0, 256, (h0, w0, 3)).astype(np.uint8)
        got = engine.resize_linear_u8(img, H, W)
        assert np.array_equal(got, cv2.resize(img, (W, H))), ((h0, w0), (W, H))
    batch = rs.randint(0, 256, (3, 60, 80, 3)).astype(np.uint8)
    got = engine.resize_linear_u8(batch, 48, 64)
    for i in range(3):
        assert np.array_equal(got[i], cv2.resize(batch[i], (64, 48)))


def test_device_resize_cubic_u8_bit_exact_vs_cv2(engine):
    """cv2.resize(..., INTER_CUBIC) on uint8 (pose_detector.py:443): OpenCV's own 8-bit path = cv2 with IPP dispatch off."""
    import cv2
    rs = np.random.RandomState(1)
    shapes = [((120, 120), (46, 46)), ((120, 120), (184, 184)), ((50, 75), (69, 46)), ((300, 18), (9, 150)), ((64, 10), (5, 32)),
              ((83, 129), (144, 92))]
    for _ in range(4):
        shapes.append(((rs.randint(10, 150), rs.randint(10, 150)), (rs.randint(5, 160), rs.randint(5, 160))))
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        for (h0, w0), (W, H) in shapes:
            img = rs.randint(0, 256, (h0, w0, 3)).astype(np.uint8)
            got = engine.resize_cubic_u8(img, H, W)
            assert np.array_equal(got, cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC)), ((h0, w0), (W, H))
            assert np.array_equal(got, R.cv2_resize_cubic_u8(img, (W, H)))
        batch = rs.randint(0, 256, (3, 60, 80, 3)).astype(np.uint8)
        got = engine.resize_cubic_u8(batch, 91, 123)
        for i in range(3):
            assert np.array_equal(got[i], cv2.resize(batch[i], (123, 91), interpolation=cv2.INTER_CUBIC))
    finally:
        cv2.ipp.setUseIPP(was)


def test_device_overlay_pixel_identical_to_cv2(engine):
    G.test_device_overlay_pixel_identical_to_cv2(engine)


def test_keypoints_exact_ties_and_threshold(engine):
    """face / hand compute_peaks_from_heatmaps (face_detector.py:55-67): exact Gaussian passes + arg-max with the
    np.where(...).flatten() tie quirk, the float32 threshold compare, mirrored hands."""
    rs = np.random.RandomState(0)
    maps = np.zeros((6, 40, 52), np.float32)
    maps[0] = 0.5
    maps[1, 10:30, 8:44] = 0.4
    maps[2] = rs.uniform(0, 1, (40, 52)).astype(np.float32)
    maps[3] = 0.0999
    maps[4] = np.float32(0.1)
    maps[5, 3, 50] = 9.0
    maps[5, 36, 1] = 9.0
    full = np.concatenate([maps, np.zeros((1, 40, 52), np.float32)])
    for mirror in (False, True):
        ref = R.keypoints_from_heatmaps(np.ascontiguousarray(full[:, :, ::-1]) if mirror else full)
        got = engine.keypoints_from_heatmaps(maps, 0.1, mirror=mirror)
        assert len(got) == len(ref) == 6
        for a, b in zip(got, ref):
            assert (a is None) == (b is None), (a, b)
            if a is not None:
                assert (a[0], a[1]) == (b[0], b[1]) and np.float32(a[2]) == np.float32(b[2]), (a, b)


@pytest.mark.parametrize("fused_peaks,paf_lowres", [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (2, 1), (3, 1)])
def test_postprocess_batch_from_network_resolution_maps(emu_native, monkeypatch, request, fused_peaks, paf_lowres):
    """opb_postprocess_batch = pose_detector.py:501-512 for a batch; with OPB_FUSED_PEAKS / OPB_PAF_LOWRES the peak
    kernel / the PAF line integrals interpolate from the low-resolution maps on demand -- every variant must give the
    oracle's peaks, connections, subsets and person records bit for bit (OPB_FUSED_PEAKS=2: materialised maps, tile-skip
    bound from the low-resolution maps)."""
    if request.node.callspec.params["emu_lib"] == "fma" and (fused_peaks, paf_lowres) not in ((0, 0), (1, 1), (2, 1), (3, 1)):
        pytest.skip("the contraction build repeats only the corner combinations (suite time)")
    monkeypatch.setenv("OPB_FUSED_PEAKS", str(fused_peaks))
    monkeypatch.setenv("OPB_PAF_LOWRES", str(paf_lowres))
    eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params(max_peaks=4096, max_candidates=65536, max_persons=128))
    run_batch_cases(eng)


def _nofma_only(request):
    if request.node.callspec.params["emu_lib"] == "fma":
        pytest.skip("one build is enough for this case (suite time)")


def test_results_do_not_depend_on_the_order_threads_run_in(emu_native, monkeypatch, request):
    """The emulator normally runs thread 0 first; OPB_EMU_ORDER=reverse runs the highest thread first.  Code that is
    correct under independent thread scheduling gives the same bits either way (and the convergence check must not
    fire): the whole post-process, default and low-resolution variants."""
    _nofma_only(request)
    monkeypatch.setenv("OPB_EMU_ORDER", "reverse")
    for knobs in ((0, 0), (1, 1)):
        monkeypatch.setenv("OPB_FUSED_PEAKS", str(knobs[0]))
        monkeypatch.setenv("OPB_PAF_LOWRES", str(knobs[1]))
        eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params(max_peaks=4096, max_candidates=65536, max_persons=128))
        run_batch_cases(eng)


def test_weight_loading_and_repack_host_code(emu_native, request):
    """opb_load_weights x92 + opb_finalize_weights (the K-major fp16 repack, hi/lo split, concat-order permutation and
    tensor-map construction are host code) for CocoPoseNet in both precisions and for FaceNet / HandNet, plus the error
    paths; run under the AddressSanitizer build (see tests/cuda_emu/build_emu.py) this is the memcheck of that code."""
    _nofma_only(request)
    syn, PD = pkg("synthetic"), pkg("pose_detector")
    prm = PD.make_opb_params(max_peaks=512, max_candidates=4096, max_persons=32)
    model = pkg("models.CocoPoseNet").CocoPoseNet()
    model.load_npz(syn.he_weights(0))
    for prec in (emu_native.PRECISION_FAST, emu_native.PRECISION_PARITY):
        eng = emu_native.Engine(0, prm, prec)
        eng.load_model(model)
        assert eng._weights_ready and eng.kp_channels == 0
    for modname, cls, n_out in (("models.FaceNet", "FaceNet", 71), ("models.HandNet", "HandNet", 22)):
        nm = pkg(modname)
        net = getattr(nm, cls)()
        net.load_npz(syn.he_weights(0, layers=nm.LAYERS))
        eng = emu_native.Engine(0, prm, emu_native.PRECISION_FAST)
        eng.load_model(net)
        assert eng.kp_channels == n_out
    eng = emu_native.Engine(0, prm)
    with pytest.raises(emu_native.OpbError):
        eng._check(eng.lib.opb_finalize_weights(eng.ctx, 0))          # nothing loaded yet
    with pytest.raises(emu_native.OpbError):
        eng._check(eng.lib.opb_load_weights(eng.ctx, b"conv1_1", None, (C.c_int64 * 4)(64, 3, 3, 3), None))


def test_peak_kernel_v2_same_bits(emu_native, monkeypatch, request):
    """OPB_PEAKS_V2=1 (with OPB_FUSED_PEAKS=2): the float32 smoothing passes of the peak kernel spread over all 256
    threads -- same sums in the same order, so every result must stay bit-identical."""
    _nofma_only(request)
    monkeypatch.setenv("OPB_FUSED_PEAKS", "2")
    monkeypatch.setenv("OPB_PAF_LOWRES", "1")
    monkeypatch.setenv("OPB_PEAKS_V2", "1")
    eng = emu_native.Engine(0, pkg("pose_detector").make_opb_params(max_peaks=4096, max_candidates=65536, max_persons=128))
    run_batch_cases(eng)
