"""Shared cases for opb_postprocess_batch (pose_detector.py:501-512 for a batch of network outputs): used by the
no-GPU emulation tests (tests/test_emu_postprocess.py) and by the B200 tests (tests/test_gpu_postprocess_batch.py)."""
import numpy as np

from conftest import pkg
from oracle import restate as R
import test_gpu_postprocess as G


def check_batch_against_oracle(eng, paf_lo, heat_lo, map_h, map_w):
    headers, persons = eng.postprocess_batch(paf_lo, heat_lo, map_h, map_w)
    for i in range(len(paf_lo)):
        pafs = R.resize_bilinear_align_corners(paf_lo[i][None], (map_h, map_w))[0]
        heat = R.resize_bilinear_align_corners(heat_lo[i][None], (map_h, map_w))[0]
        _, scores, parts = R.postprocess_fast(pafs, heat, map_w, map_w, map_h, map_h, return_parts=True)
        pk, conns, subs = eng.image_detail(i)
        if parts is None:
            assert headers["n_peaks"][i] == 0 and headers["n_persons"][i] == 0
            continue
        assert headers["status"][i] == 0
        assert np.array_equal(pk, parts["all_peaks"])
        G._check_conns(conns, parts["connections"])
        assert np.array_equal(subs, parts["subsets"])
        n = int(headers["n_persons"][i])
        assert n == len(scores) and np.array_equal(persons["score"][i, :n], scores)
        ids = parts["subsets"][:, :18].astype(np.int64)
        assert np.array_equal(persons["peak_id"][i, :n], ids)
        xs = np.where(ids >= 0, parts["all_peaks"][np.maximum(ids, 0), 1], 0)
        assert np.array_equal(persons["x"][i, :n], xs.astype(np.int64))
    return headers



def run_batch_cases(eng):
    syn = pkg("synthetic")
    # two different synthetic 8-person frames at the benchmark shape (46x82 -> 320x576)
    lo = [syn.eight_person_lowres(46, 82, seed=s) for s in (0, 3)]
    paf_lo = np.stack([p for p, _ in lo])
    heat_lo = np.stack([h for _, h in lo])
    headers = check_batch_against_oracle(eng, paf_lo, heat_lo, 320, 576)
    assert (headers["n_persons"] == 8).all()
    # awkward sizes: non-multiple-of-tile maps, odd low-res shapes, noise (many near-threshold peaks)
    rs = np.random.RandomState(5)
    paf_lo = (rs.standard_normal((1, 38, 13, 17)) * 0.5).astype(np.float32)
    heat_lo = (rs.standard_normal((1, 19, 13, 17)) * 0.2).astype(np.float32)
    check_batch_against_oracle(eng, paf_lo, heat_lo, 75, 101)
    # nothing above the threshold
    headers = check_batch_against_oracle(eng, np.zeros((1, 38, 8, 8), np.float32), np.zeros((1, 19, 8, 8), np.float32), 40, 40)
    assert headers["n_peaks"][0] == 0
