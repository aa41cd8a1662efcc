"""Build libpfpp_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python puzzlefusion-plusplus_b200/build.py [--force]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 (nvcc defaults otherwise: no --use_fast_math;
-fmad=true, which is why every fp32 operation whose rounding must match the oracle is written with the
__fadd_rn / __fmul_rn intrinsics of common.cuh, which nvcc never contracts).  The .so is git-ignored but travels to
the GPU box with the gpurun snapshot.  The library is rebuilt whenever the hash of its sources differs from the
pfpp_build_id() the existing .so reports.
"""
import ctypes
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
DEPS = SRC + sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [os.path.join(HERE, "..", "include", "pfpp.h")]
OUT = os.path.join(HERE, "libpfpp_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def source_id():
    """FNV-1a 64 over the sources the library is built from (names + contents, in sorted order)."""
    h = 0xcbf29ce484222325
    for path in DEPS:
        for blob in (os.path.basename(path).encode(), open(path, "rb").read()):
            for byte in blob:
                h = ((h ^ byte) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def built_id():
    """pfpp_build_id() of the existing .so (None if it is missing or predates the symbol)."""
    if not os.path.exists(OUT):
        return None
    try:
        fn = ctypes.CDLL(OUT).pfpp_build_id
        fn.restype = ctypes.c_ulonglong
        return int(fn())
    except (OSError, AttributeError):
        return None


def up_to_date():
    return built_id() == source_id()


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    sid = source_id()
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SRC:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [NVCC] + FLAGS + ["-DPFPP_BUILD_ID=%dull" % sid, "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-lcudart_static", "-lrt", "-lpthread", "-ldl"])
    if verbose:
        print("built", OUT, "build id %016x" % sid)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
