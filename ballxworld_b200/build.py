"""Builds ballxworld_b200/csrc/libbxw_cuda.so in-tree with nvcc for sm_100a (no other arch, no JIT cache)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libbxw_cuda.so")
SOURCES = ["bxw_api.cu", "mesh_kernel.cu", "emit_kernel.cu", "gen_kernel.cu", "aux_kernels.cu", "sched_kernels.cu", "tables.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-cudart", "static", "--fmad=false",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".hpp"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "bxw_cuda.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, phase_timing=False):
    """phase_timing: profiling variant csrc/libbxw_cuda_phase.so (-DBXW_PHASE_TIMING; select it with BXW_CUDA_LIB)."""
    out = LIB.replace(".so", "_phase.so") if phase_timing else LIB
    if not force and not phase_timing and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = ["-DBXW_PHASE_TIMING"] if phase_timing else []
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out] + srcs
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phase_timing="--phase-timing" in sys.argv))
