"""TEST INFRASTRUCTURE ONLY: builds tests/cuda_emu/_build/libopb_emu.so, the library's own
sources (csrc/opb_api.cu + kernels, unmodified) compiled by g++ on top of cuda_emu.h (CUDA
execution model as fibers) and emu_runtime.cpp (host-memory CUDA runtime stubs).

The package never loads this library: only tests/test_emu_*.py dlopen it, to run the
plain-CUDA kernels and the C-ABI host orchestration on a box without a GPU and compare them
bit-for-bit with the oracle.  Tensor-core kernels (inline PTX) trap under emulation.

A memcheck analogue without a GPU: every "device" allocation is a host heap block, so an AddressSanitizer build of
the emulated library reports out-of-bounds and use-after-free accesses of kernels and host code alike:
    OPB_EMU_SANITIZE=address LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
        ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_emu_postprocess.py -q

A racecheck analogue: a ThreadSanitizer build in which every CUDA thread is a TSan fiber and barriers / warp
primitives are the only happens-before edges (see cuda_emu.h); unsynchronised accesses of two CUDA threads to the
same shared or global location are reported with both source lines (log files appear only if something is found):
    OPB_EMU_SANITIZE=thread LD_PRELOAD=$(gcc -print-file-name=libtsan.so) OMP_NUM_THREADS=1 \
        TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0 log_path=/tmp/opb_tsan" \
        python -m pytest tests/test_emu_postprocess.py -q -k nofma

Source rewriting (textual, into _build/src; the originals are not touched):
  kernel<<<grid, block, smem, stream>>>(args)  ->  emu::Launcher(grid, block, smem, stream).run(kernel, args)
  extern __shared__ T name[];                  ->  T* name = reinterpret_cast<T*>(emu::dyn_smem());
  asm volatile(...);                           ->  emu::unsupported_ptx();
  csrc/ptx.cuh (inline-PTX wrappers)           ->  ptx_emu.cuh: functional model of mbarrier / TMA / tcgen05 + TMEM, so
                                                   the tensor-core conv kernels run too (CTA-pair kernels excepted)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "chainer_realtime_multi-person_pose_estimation_b200")
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(HERE, "_build")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")

_LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
_DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def _strip_asm(text):
    out, i = [], 0
    while True:
        m = re.compile(r"\basm\s+(?:volatile\s*)?\(").search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        out.append(text[i:m.start()])
        depth, j, in_str = 1, m.end(), False
        while depth:
            c = text[j]
            if in_str:
                if c == "\\":
                    j += 1
                elif c == '"':
                    in_str = False
            elif c == '"':
                in_str = True
            elif c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            j += 1
        while text[j] in " \t\r\n":
            j += 1
        assert text[j] == ";", text[m.start():j + 1]
        out.append("emu::unsupported_ptx();")
        i = j + 1


def rewrite(text):
    text = _strip_asm(text)
    text = text.replace('#include "../../include/opb.h"', '#include "opb.h"')
    text = _LAUNCH.sub(lambda m: "emu::Launcher(%s).run(%s, " % (m.group(2), m.group(1)), text)
    text = _DYN_SMEM.sub(lambda m: "%s* %s = reinterpret_cast<%s*>(emu::dyn_smem());" % (m.group(1), m.group(2), m.group(1)), text)
    return text


SANITIZE = os.environ.get("OPB_EMU_SANITIZE", "")   # "address": -fsanitize=address (run pytest under LD_PRELOAD=libasan.so)


def lib_path(contract):
    return os.path.join(BUILD, "libopb_emu_%s%s.so" % ("fma" if contract else "nofma", "_" + SANITIZE if SANITIZE else ""))


class EmuCompileError(Exception):
    """g++ ran and rejected the (rewritten) device sources: a defect of the sources under test, not a missing harness --
    the tests must FAIL on it (a skip would hide a library that no longer compiles)."""


def build(contract=False, force=False):
    """contract=True compiles with -ffp-contract=fast -mfma: g++ may then fuse every a*b+c that is
    not written with an explicit _rn intrinsic, as nvcc does by default -- results must not change."""
    src_dir = os.path.join(BUILD, "src")
    os.makedirs(src_dir, exist_ok=True)
    lib = lib_path(contract)
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if os.path.isfile(os.path.join(CSRC, f))] + [
        os.path.join(HERE, f) for f in ("cuda_emu.h", "ptx_emu.cuh", "emu_tmap.h", "emu_runtime.cpp", "build_emu.py")] + [
        os.path.join(ROOT, "include", "opb.h")]
    if not force and os.path.isfile(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    for f in sorted(os.listdir(CSRC)):
        if not os.path.isfile(os.path.join(CSRC, f)):
            continue
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        if f == "ptx.cuh":
            # the inline-PTX wrappers are replaced by the functional model in ptx_emu.cuh; the UMMA descriptor
            # encoders (plain C++) are taken from the original so that the kernels' own encodings are what runs
            desc = text[text.index("// K-major operand, 128-byte swizzle"):text.index("}  // namespace ptx")]
            with open(os.path.join(HERE, "ptx_emu.cuh")) as fh:
                text = fh.read().replace("// @@DESCRIPTORS@@", desc)
        text = rewrite(text)
        with open(os.path.join(src_dir, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as fh:
            fh.write(text)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-w", "-Wl,-Bsymbolic",   # own cuda* stubs win over a loaded libcudart
           "-ffp-contract=fast" if contract else "-ffp-contract=off"] + (["-mfma"] if contract else []) + (
           ["-fsanitize=" + SANITIZE, "-fno-omit-frame-pointer"] if SANITIZE else []) + [
           "-I", CUDA_INC, "-I", os.path.join(ROOT, "include"), "-I", HERE, "-include", os.path.join(HERE, "cuda_emu.h"),
           os.path.join(src_dir, "opb_api.cpp"), os.path.join(HERE, "emu_runtime.cpp"), "-o", lib]
    import shutil
    if shutil.which("g++") is None or not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")):
        raise RuntimeError("no g++ / CUDA headers on this box")          # harness unavailable -> the tests skip
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        open(os.path.join(BUILD, "build.log"), "w").write(r.stdout); sys.stderr.write(r.stdout[-3000:])
        raise EmuCompileError("g++ failed building the emulated library:\n" + r.stdout[-1500:])
    return lib


if __name__ == "__main__":
    print(build(contract="--fma" in sys.argv, force=True))
