"""Coarse-solve micro harness (profiling aid): set-up for the GL grid, then repeated coarsest-level solves."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import _lib
from tests.helpers import StageRunner
from oracle.cases import IV_LISTS
dims = (32, 64, 64); B = int(os.environ.get("B", "32")); reps = int(os.environ.get("REPS", "10"))
n_grid, dsf = int(os.environ.get("NGRID", "4")), bool(int(os.environ.get("DSF", "0")))
lib = _lib.get_library()
G = int(np.prod(dims)); M = 7
g = torch.Generator().manual_seed(1)
coeffs = torch.zeros(B, G, M, dtype=torch.float64); coeffs[..., 0] = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
coeffs[..., 1] = 1; coeffs[..., 5] = -1; coeffs[..., 6] = -1
steps = [np.full((B, n - 1), h) for n, h in zip(dims, (0.1, 0.3906, 0.3906))]
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sr = StageRunner(lib, "cuda:0", dims, IV_LISTS["gl"], B, n_grid, dsf, coeffs.numpy(), steps)
e1.record(); torch.cuda.synchronize()
print("setup+factor (incl. host work) ms", e0.elapsed_time(e1))
lc = n_grid - 1
nc = sr.level_n(lc)
torch.manual_seed(7)
b = torch.randn(B * nc, dtype=torch.float64, device=dev); out = torch.zeros_like(b)
cfg = sr.plan.cfg(False)
def call():
    lib.check(lib.dll.pdeop_stage(sr.plan.handle, ctypes.byref(cfg), _lib.STAGE_COARSE_SOLVE, lc, 0, _lib._ptr(b), None, _lib._ptr(out),
                                  _lib._ptr(sr.persist), _lib._ptr(sr.scratch), _lib.current_stream_ptr(dev)))
call(); torch.cuda.synchronize()
e0.record()
for _ in range(reps): call()
e1.record(); torch.cuda.synchronize()
print(f"coarse solve n={nc} B={B}: {e0.elapsed_time(e1)/reps:.3f} ms per solve (incl. pack/unpack)")
tag = os.environ.get("TAG")
if tag:
    os.makedirs("gpurun_out", exist_ok=True)
    np.save(f"gpurun_out/cs_out_{tag}.npy", out.cpu().numpy())
    print("checksum", float(out.abs().sum()), "finite", bool(torch.isfinite(out).all()))
