"""In-tree build of the native library (explicit nvcc / g++ commands, no JIT cache).

  libmm2chain_b200.so = csrc/chain_kernels.cu + csrc/chain_api.cu + csrc/seed_kernels.cu (nvcc, sm_100a)
                        + host/chain_backend.cpp + host/map_backend.cpp (g++)
  mm2b-replay         = host/replay_main.cpp (g++) linked against the library: the batched caller for anchor dumps

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libmm2chain_b200.so")
LIB_DBG = os.path.join(HERE, "libmm2chain_b200_dbg.so")
REPLAY = os.path.join(HERE, "mm2b-replay")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]

CU_SRCS = ["csrc/chain_kernels.cu", "csrc/chain_api.cu", "csrc/seed_kernels.cu"]
CXX_SRCS = ["host/chain_backend.cpp", "host/map_backend.cpp"]
HEADERS = ["csrc/chain_kernels.cuh", "csrc/sort_replay.cuh", "csrc/seed_kernels.cuh", "csrc/shim_internal.h", "../include/mm2chain_b200.h", "../include/mm2seed_b200.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_all(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS]
    objs = []
    for src in CU_SRCS:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or _newer(o, [s] + hdrs):
            _run([NVCC] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + INC + ["-c", s, "-o", o], verbose)
        objs.append(o)
    for src in CXX_SRCS:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or _newer(o, [s] + hdrs):
            _run(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-I" + os.path.join(CUDA_HOME, "include")] + INC + ["-c", s, "-o", o], verbose)
        objs.append(o)
    if force or _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-lpthread"], verbose)
    # range-checking twin of the library (same sources, kernels compiled with -DMM2B_DEBUG_CHECKS); used only by tests
    dbg_obj = os.path.join(OBJ, "chain_kernels_dbg.cu.o")
    ksrc = os.path.join(HERE, "csrc/chain_kernels.cu")
    if force or _newer(dbg_obj, [ksrc] + hdrs):
        _run([NVCC] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-DMM2B_DEBUG_CHECKS"] + INC + ["-c", ksrc, "-o", dbg_obj], verbose)
    if force or _newer(LIB_DBG, [dbg_obj] + objs[1:]):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB_DBG, dbg_obj] + objs[1:] + ["-cudart", "static", "-lpthread"], verbose)
    # the batched caller for anchor dumps (include/mm2chain_dump.h); finds the library next to itself
    rsrc = os.path.join(HERE, "host/replay_main.cpp")
    if force or _newer(REPLAY, [rsrc, LIB, os.path.join(ROOT, "include/mm2chain_dump.h")] + hdrs):
        _run(["g++", "-O2", "-std=c++17", "-Wall"] + INC + [rsrc, "-o", REPLAY, "-L" + HERE, "-lmm2chain_b200", "-Wl,-rpath,$ORIGIN", "-lz"], verbose)
    return LIB


if __name__ == "__main__":
    build_all(verbose=True, force="--force" in sys.argv)
