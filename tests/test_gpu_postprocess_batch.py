"""`-m gpu`: opb_postprocess_batch (the post-process half of PoseDetector.__call__, pose_detector.py:501-512, for
a batch of network outputs) against the oracle, bit for bit -- through the C ABI on a B200.

The entry point and the OPB_FUSED_PEAKS / OPB_PAF_LOWRES / OPB_PEAKS_V2 variants (peak kernel / PAF line integrals
interpolating from the low-resolution maps on demand) were written after round 1's GPU budget was spent.  Their
bit-exactness is covered without a GPU by tests/test_emu_postprocess.py; every test of this file runs on a B200 with
OPB_TEST_EXPERIMENTAL=1 (first thing next round: tools/gpu_round.sh lowres_ab), and the default-path test loses its
skip marker once it has passed there."""
import os

import pytest

from conftest import pkg
from postprocess_batch_cases import run_batch_cases

pytestmark = pytest.mark.gpu


def _engine():
    native = pkg("_native")
    return native.Engine(0, pkg("pose_detector").make_opb_params(max_peaks=4096, max_candidates=65536, max_persons=128))


_NOT_YET_ON_GPU = pytest.mark.skipif(
    os.environ.get("OPB_TEST_EXPERIMENTAL", "0") != "1",
    reason="opb_postprocess_batch was written after this round's GPU budget was spent: validated on the CPU under emulation "
           "(tests/test_emu_postprocess.py, incl. AddressSanitizer / ThreadSanitizer builds); first B200 run: "
           "OPB_TEST_EXPERIMENTAL=1 (tools/gpu_round.sh lowres_ab)")


@_NOT_YET_ON_GPU
def test_postprocess_batch_default_path(monkeypatch):
    monkeypatch.delenv("OPB_FUSED_PEAKS", raising=False)
    monkeypatch.delenv("OPB_PAF_LOWRES", raising=False)
    run_batch_cases(_engine())


@pytest.mark.skipif(os.environ.get("OPB_TEST_EXPERIMENTAL", "0") != "1", reason="experimental low-res variants: set OPB_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("fused_peaks,paf_lowres", [(1, 0), (2, 0), (0, 1), (1, 1), (2, 1)])
def test_postprocess_batch_lowres_variants(monkeypatch, fused_peaks, paf_lowres):
    monkeypatch.setenv("OPB_FUSED_PEAKS", str(fused_peaks))
    monkeypatch.setenv("OPB_PAF_LOWRES", str(paf_lowres))
    run_batch_cases(_engine())


@pytest.mark.skipif(os.environ.get("OPB_TEST_EXPERIMENTAL", "0") != "1", reason="experimental: set OPB_TEST_EXPERIMENTAL=1")
def test_postprocess_batch_peak_kernel_v2(monkeypatch):
    monkeypatch.setenv("OPB_FUSED_PEAKS", "2")
    monkeypatch.setenv("OPB_PAF_LOWRES", "1")
    monkeypatch.setenv("OPB_PEAKS_V2", "1")
    run_batch_cases(_engine())
