"""Build the in-tree native library of the hider (sm_100a only).

    python -m aqsis_b200.build            # build if sources are newer than the library
    python -m aqsis_b200.build --force

The library is built IN-TREE (aqsis_b200/_lib/libaqsis_b200_hider.so) so that it travels with
the repository snapshot to the GPU box.  -fmad=false is not a tuning choice: the reference's
arithmetic has no fused multiply-add and bit parity depends on it.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libaqsis_b200_hider.so")
SOURCES = ["hider_kernels.cu", "hider_api.cpp", "hider_shard.cpp", "host_sampling.cpp", "host_filters.cpp"]
HEADERS = ["hider_device.h", "hider_internal.h", "host_sampling.h", os.path.join("..", "..", "include", "aqsis_b200_hider.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the hider has no CPU build")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False, phase_timing=False, defines=(), out=None):
    """phase_timing: a DEVELOPMENT build whose k_hide counts warp cycles per phase (-DAQH_PHASE_TIMING, printed on stderr
    after every frame); never the library that is benchmarked."""
    if not force and not phase_timing and not defines and not out and not needs_build():
        return LIB
    target = out or LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O3", "-cudart", "static", "-shared",
           "-o", target + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    for d in defines:                      # experimental variants (tools/gpu_ab.sh): python -m aqsis_b200.build --out X -DNAME
        cmd.insert(1, "-D" + d)
    if phase_timing:
        cmd.insert(1, "-DAQH_PHASE_TIMING")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    os.replace(target + ".tmp", target)
    return target


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phase_timing="--phase-timing" in sys.argv,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")], out=_out))
