"""`-m gpu`: opb_postprocess_batch (the post-process half of PoseDetector.__call__, pose_detector.py:501-512, for
a batch of network outputs) against the oracle, bit for bit -- through the C ABI on a B200.

Covers every OPB_FUSED_PEAKS / OPB_PAF_LOWRES / OPB_PEAKS_V2 variant (peak kernel / PAF line integrals interpolating
from the low-resolution maps on demand): all pass on a B200 since round 2 (profiles/r02_lowres_ab.txt holds the A/B
that made OPB_FUSED_PEAKS=1 + OPB_PAF_LOWRES=1 the default of the fused pipeline)."""
import pytest

from conftest import pkg
from postprocess_batch_cases import run_batch_cases

pytestmark = pytest.mark.gpu


def _engine():
    native = pkg("_native")
    return native.Engine(0, pkg("pose_detector").make_opb_params(max_peaks=4096, max_candidates=65536, max_persons=128))


def test_postprocess_batch_default_path(monkeypatch):
    monkeypatch.delenv("OPB_FUSED_PEAKS", raising=False)
    monkeypatch.delenv("OPB_PAF_LOWRES", raising=False)
    run_batch_cases(_engine())


def test_postprocess_batch_materialised_maps(monkeypatch):
    """the reference's data flow (pose_detector.py:501-502: full-resolution maps are materialised)"""
    monkeypatch.setenv("OPB_FUSED_PEAKS", "0")
    monkeypatch.setenv("OPB_PAF_LOWRES", "0")
    run_batch_cases(_engine())


@pytest.mark.parametrize("fused_peaks,paf_lowres", [(1, 0), (2, 0), (0, 1), (1, 1), (2, 1), (3, 0), (3, 1)])
def test_postprocess_batch_lowres_variants(monkeypatch, fused_peaks, paf_lowres):
    monkeypatch.setenv("OPB_FUSED_PEAKS", str(fused_peaks))
    monkeypatch.setenv("OPB_PAF_LOWRES", str(paf_lowres))
    run_batch_cases(_engine())


def test_postprocess_batch_peak_kernel_v2(monkeypatch):
    monkeypatch.setenv("OPB_FUSED_PEAKS", "2")
    monkeypatch.setenv("OPB_PAF_LOWRES", "1")
    monkeypatch.setenv("OPB_PEAKS_V2", "1")
    run_batch_cases(_engine())
