"""Gauss-Seidel kernel A/B harness (profiling aid): one operator set-up, then 5-sweep smoothing calls per kernel
variant (per-plan switch gs_pipe), timed with CUDA events; the pack/unpack kernels of the stage call are measured
separately (count = 0) and subtracted; outputs are compared bit for bit with the one-launch-per-step kernel.
Usage: python tools/gs_bench.py N0 N1 N2 [modes...]   env: B, REPS, NGRID, DSF"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import _lib
from oracle.cases import IV_LISTS
from tests.helpers import StageRunner

nd = 2 if os.environ.get("D2") else 3
dims = tuple(int(a) for a in sys.argv[1:1 + nd])
modes = [int(a) for a in sys.argv[1 + nd:]] or [0, 4]
B = int(os.environ.get("B", "32"))
reps = int(os.environ.get("REPS", "10"))
n_grid = int(os.environ.get("NGRID", "2"))
lib = _lib.PdeopLibrary(os.environ['PDEOP_VARIANT_SO']) if os.environ.get('PDEOP_VARIANT_SO') else _lib.get_library()
G = int(np.prod(dims))
M = 1 + 2 * nd
g = torch.Generator().manual_seed(1)
coeffs = torch.zeros(B, G, M, dtype=torch.float64)
coeffs[..., 0] = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
coeffs[..., 1] = 1
coeffs[..., 1 + nd:] = -1
hs = (0.1, 0.3906, 0.3906) if nd == 3 else (0.025, 20.0 / 256)
steps = [np.full((B, n - 1), h) * (1 + 0.05 * np.random.default_rng(c).random((B, n - 1))) for c, (n, h) in enumerate(zip(dims, hs))]
sr = StageRunner(lib, "cuda:0", dims, IV_LISTS["gl" if nd == 3 else "burgers"], B, n_grid, os.environ.get("DSF", "0") == "1",
                 coeffs.numpy(), steps)
dev = torch.device("cuda:0")
n = B * G * M
b = torch.randn(n, dtype=torch.float64, device=dev)
x = torch.randn(n, dtype=torch.float64, device=dev)
out = torch.zeros(n, dtype=torch.float64, device=dev)


def call(count, variant=0):
    cfg = sr.plan.cfg(False)
    cfg.gs_variant = variant
    lib.check(lib.dll.pdeop_stage(sr.plan.handle, ctypes.byref(cfg), _lib.STAGE_GS, 0, count, _lib._ptr(b), _lib._ptr(x),
                                  _lib._ptr(out), _lib._ptr(sr.persist), _lib._ptr(sr.scratch),
                                  _lib.current_stream_ptr(dev)))


def timed(count):
    call(count)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call(count)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


call(5, variant=1)
torch.cuda.synchronize()
ref = out.clone()
base = timed(0)
alg = 4 * M * 8 * G * B * 5
for mode in modes:
    sr.plan.set_tuning("gs_pipe", mode)
    ms = timed(5) - base
    same = bool(torch.equal(out, ref))
    print(f"dims {dims} B {B} gs_pipe={mode}: {ms:.3f} ms per 5 sweeps  ({alg / ms / 1e6:.0f} GB/s algorithmic, "
          f"{alg / ms / 1e6 / 6555.2:.3f} of HBM peak)  bit-identical to step kernel: {same}", flush=True)
