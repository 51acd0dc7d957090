"""Builds csrc/libpdeop.so: hand-written CUDA for sm_100a behind the C ABI of include/pdeop.h.

    python -m mech_nn_discovery_pde_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is built in-tree so it travels with the repository
snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libpdeop.so")
SOURCES = ["pdeop_cuda.cu", "pdeop_solver.cpp", "pdeop_coeff.cu"]
HEADERS = ["pdeop_common.h", "pdeop_elem.h", "pdeop_backend.h", "pdeop_lstsq.h", "pdeop_gs_fast.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def build_debug(out_path, extra_flags):
    """Debug/experiment build into another path (e.g. tools/): never loaded by the package."""
    cmd = [NVCC] + FLAGS + list(extra_flags) + ["-o", out_path] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out_path


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(os.path.dirname(HERE), "include", "pdeop.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpdeop.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
