"""CPU: the CUDA-on-CPU emulator checked against itself (tests/cuda_emu/selftest_kernels.cpp): warp / block
primitives give hand-computed values; the lockstep-divergence check aborts; and the ThreadSanitizer build (the
racecheck analogue) reports purpose-built block-level and warp-level races and stays silent on their synchronised
twins.  Test infrastructure only."""
import os
import subprocess
import sys

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu")
sys.path.insert(0, HERE)
import build_emu  # noqa: E402


def _build(tsan):
    os.makedirs(build_emu.BUILD, exist_ok=True)
    exe = os.path.join(build_emu.BUILD, "selftest_tsan" if tsan else "selftest")
    srcs = [os.path.join(HERE, f) for f in ("selftest_kernels.cpp", "emu_runtime.cpp")]
    deps = srcs + [os.path.join(HERE, "cuda_emu.h")]
    if not os.path.isfile(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-w", "-I", build_emu.CUDA_INC, "-I", HERE, "-include",
               os.path.join(HERE, "cuda_emu.h")] + (["-fsanitize=thread"] if tsan else []) + srcs + ["-o", exe]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            pytest.skip("emulator self-test build unavailable: %s" % r.stdout[-200:])
    return exe


def _run(exe, mode):
    env = dict(os.environ, TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0 exitcode=66")
    return subprocess.run([exe, mode], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=120)


def test_primitives_and_divergence_check():
    exe = _build(False)
    r = _run(exe, "func")
    assert r.returncode == 0 and r.stdout.startswith("func ok"), r.stdout + r.stderr
    r = _run(exe, "diverge")
    assert r.returncode != 0 and "divergent warp-synchronous code" in r.stderr
    for mode in ("race_block", "sync_block", "race_warp", "sync_warp"):     # without TSan all of them just run
        assert _run(exe, mode).returncode == 0


def test_racecheck_reports_races_and_only_races():
    exe = _build(True)
    probe = _run(exe, "sync_block")
    if "unexpected memory mapping" in probe.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow on this kernel configuration")
    for mode, racy in (("race_block", True), ("sync_block", False), ("race_warp", True), ("sync_warp", False)):
        r = _run(exe, mode)
        reported = "WARNING: ThreadSanitizer: data race" in r.stderr
        assert reported == racy, (mode, r.stderr[-600:])
