"""Builds libjwasb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU), and libjwasio.so (host C,
gcc: genotype text files -> 2-bit image, include/jwas_io.h).

Environment: JWAS_B200_BUILD_SO=<path> writes another file (select it at run time with JWAS_B200_LIB=<path>, e.g.
for an A/B of two builds on one box); JWAS_B200_BUILD_FLAGS adds nvcc flags:
  -DJW_TIMERS  in-kernel phase timers for tools/phase_probe.py (cost the sweep ~10 %)
  -DJW_NEXT    round-2 candidates not yet measured on a GPU (arithmetic block metadata in the record replay,
               L1 prefetch of the next chunk, red instead of atomicAdd; profiles/r1_fused_kernel_stall_hotspots.md)"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libjwasb200.so")
SRC = os.path.join(HERE, "csrc", "jwas_b200.cu")


IO_SO = os.path.join(HERE, "libjwasio.so")
IO_SRC = os.path.join(HERE, "csrc", "io", "jw_io.c")
IO_SRCS = [IO_SRC, os.path.join(HERE, "csrc", "io", "jw_annot.c")]


def _newest_source_mtime():
    m = 0.0
    for d in (os.path.join(HERE, "csrc"), os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h")) and f != "jwas_io.h":
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build_io(force=False):
    """libjwasio.so: plain C + OpenMP, no CUDA."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "jwas_io.h")
    if not force and os.path.exists(IO_SO) and os.path.getmtime(IO_SO) >= max([os.path.getmtime(f) for f in IO_SRCS + [hdr]]):
        return IO_SO
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([gcc, "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wextra",
                           "-o", IO_SO] + IO_SRCS + ["-lm"])
    return IO_SO


def build(force=False, verbose=False):
    build_io(force)
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source_mtime():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    so = os.environ.get("JWAS_B200_BUILD_SO", SO)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
           "-o", so, SRC, "-lcublas"] + os.environ.get("JWAS_B200_BUILD_FLAGS", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)
