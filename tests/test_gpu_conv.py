"""`-m gpu`: the tcgen05 implicit-GEMM conv kernel against a plain PyTorch fp32 conv2d of the
same op (floating-point kernel -> torch fp32 reference, tolerance stated per mode):
  fast   : operands rounded to fp16, fp32 accumulate, fp16 output  -> 2e-3 * scale
  parity : split-fp16 operands (hi+lo, 3 MMAs), hi+lo output       -> 1e-4 * scale
  comp   : fp16 product + 8-bit-float rounding corrections (kind::f8f6f4 MMAs into the same accumulator),
           hi + e5m2 lo output                                      -> 3e-4 * scale (vs the UNquantized reference)
where scale = max|reference|.  Covers every (ksize, Cout tile) kernel variant the chain uses,
partial tiles at the right/bottom edge, zero padding at all four borders and multi-image
batches."""
import numpy as np
import pytest
import torch

from conftest import pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    native = pkg("_native")
    return native.Engine(0, pkg("pose_detector").make_opb_params())


def _ref_conv(x_nhwc, W, b, relu, quantize):
    x = torch.from_numpy(x_nhwc).permute(0, 3, 1, 2).contiguous()
    w = torch.from_numpy(W)
    if quantize:
        x = x.half().float()
        w = w.half().float()
    y = torch.nn.functional.conv2d(x.double(), w.double(), torch.from_numpy(b).double(), padding=(W.shape[2] - 1) // 2)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


CASES = [
    # n, h, w, cin, cout, ks, relu
    (1, 46, 82, 128, 128, 7, 1),     # the dominant family (Mconv2..5)
    (2, 46, 46, 185, 256, 7, 1),     # fused Mconv1 (Cin 185 -> 192, N = 256)
    (2, 23, 31, 64, 64, 3, 1),       # conv1_2-like, partial tiles in x and y
    (1, 48, 40, 128, 128, 3, 0),
    (1, 32, 24, 256, 256, 3, 1),
    (1, 24, 24, 256, 512, 3, 1),     # two N blocks
    (1, 46, 82, 128, 512, 1, 1),     # conv5_4
    (2, 30, 17, 512, 38, 1, 0),      # PAF head (N = 48 tile, 38 valid)
    (1, 46, 46, 128, 19, 1, 0),      # heat head
    (1, 40, 40, 128, 128, 1, 1),     # Mconv6
    # tile-edge coverage of the weights-as-A (swap) 7x7 kernel: W % 16 == 3 / 0 / 8 / 12, two images
    (2, 23, 35, 128, 128, 7, 1),
    (1, 40, 16, 128, 128, 7, 0),
    (1, 30, 24, 128, 128, 7, 1),
    (1, 25, 28, 128, 256, 7, 1),     # Cout 256 in fast mode = the CTA-pair kernel, odd tile counts
]


_PREC = {"fast": "PRECISION_FAST", "parity": "PRECISION_PARITY", "comp": "PRECISION_COMP"}
_TOL = {"fast": 2e-3, "parity": 1e-4, "comp": 3e-4}


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_%dx%d_c%d_o%d_k%d" % c[:6])
@pytest.mark.parametrize("mode", ["fast", "parity", "comp"])
def test_conv_vs_torch(engine, case, mode):
    n, h, w, cin, cout, ks, relu = case
    native = pkg("_native")
    rs = np.random.RandomState(hash(case) % (2 ** 31))
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    prec = getattr(native, _PREC[mode])
    y = engine.test_conv(x, W, b, relu, prec)
    ref = _ref_conv(x, W, b, relu, quantize=(mode == "fast"))
    scale = np.abs(ref).max()
    err = np.abs(y - ref).max()
    tol = _TOL[mode] * scale
    assert err <= tol, "max abs err %.3e > tol %.3e (scale %.3f)" % (err, tol, scale)


def test_conv_zero_padding_borders(engine):
    """All-ones input/weights: the output counts the in-image taps, so border handling
    (TMA out-of-bounds zero fill) is checked exactly."""
    native = pkg("_native")
    for ks in (3, 7):
        x = np.ones((1, 30, 19, 64), np.float32)
        W = np.ones((64, 64, ks, ks), np.float32) / 64.0
        b = np.zeros(64, np.float32)
        y = engine.test_conv(x, W, b, 0, native.PRECISION_FAST)
        ref = _ref_conv(x, W, b, 0, quantize=True)
        assert np.array_equal(y, ref.astype(np.float32))


@pytest.mark.parametrize("mode", ["fast", "parity", "comp"])
@pytest.mark.parametrize("case", [(2, 24, 40, 64, 64, 3), (1, 32, 18, 128, 128, 3), (1, 46, 82, 256, 256, 3)],
                         ids=lambda c: "n%d_%dx%d_c%d_o%d_k%d" % c)
def test_conv_fused_maxpool(engine, case, mode):
    """conv + ReLU + F.max_pooling_2d(2,2) fused in the epilogue (conv1_2 / conv2_2 / conv3_4)."""
    n, h, w, cin, cout, ks = case
    native = pkg("_native")
    rs = np.random.RandomState(11)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    prec = getattr(native, _PREC[mode])
    y = engine.test_conv(x, W, b, 1, prec, pool=True)
    ref = _ref_conv(x, W, b, 1, quantize=(mode == "fast"))
    ref = torch.nn.functional.max_pool2d(torch.from_numpy(ref).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).numpy()
    scale = np.abs(ref).max()
    err = np.abs(y - ref).max()
    assert y.shape == ref.shape
    assert err <= _TOL[mode] * scale, "max abs err %.3e (scale %.3f)" % (err, scale)


@pytest.mark.parametrize("mode", ["fast", "comp"])
@pytest.mark.parametrize("knob,case", [
    ("OPB_SWAP7", (32, 46, 82, 128, 128, 7, 1)),        # lean-issue 7x7 kernel + LPT tile lists vs conv_tcgen05_swap_kernel
    ("OPB_SWAP7_LPT", (32, 46, 82, 128, 128, 7, 1)),    # LPT tile lists vs round-robin tiles
    ("OPB_PAIR_UNITS", (32, 46, 82, 185, 256, 7, 1)),   # CTA pairs on consecutive 8-column units vs 16-column pair tiles
], ids=["swap7", "lpt", "pair_units"])
def test_scheduling_variants_bit_identical_at_benchmark_size(monkeypatch, knob, case, mode):
    """BASELINE.json's full size (batch 32, 46x82 maps): the round-2 scheduling changes only reorder WHICH CTA computes
    which tile (or which kernel issues the same MMAs); every output must be bit-identical to the variant they replace --
    a size-independent property checked where the tile counts, LPT lists and ring wrap-arounds are the benchmark's."""
    native = pkg("_native")
    n, h, w, cin, cout, ks, relu = case
    rs = np.random.RandomState(11)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    W = (rs.standard_normal((cout, cin, ks, ks)) * np.sqrt(2.0 / (cin * ks * ks))).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.1).astype(np.float32)
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv(knob, flag)
        eng = native.Engine(0, pkg("pose_detector").make_opb_params())
        out.append(eng.test_conv(x, W, b, relu, getattr(native, _PREC[mode])))
        del eng
    assert np.isfinite(out[0]).all() and np.array_equal(out[0], out[1])
