"""K-apply micro harness (profiling aid): times the fine-level y = K x kernel; PDEOP_VARIANT_FLAGS builds and uses a
variant library with extra nvcc flags (A/B testing)."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import _lib
from mech_nn_discovery_pde_b200.build import build_debug
from tests.helpers import StageRunner
from oracle.cases import IV_LISTS
flags = os.environ.get("PDEOP_VARIANT_FLAGS", "")
if flags:
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpdeop_variant.so")
    build_debug(so, flags.split())
    lib = _lib.PdeopLibrary(so)
else:
    lib = _lib.get_library()
dims, B, n_grid = (32, 64, 64), 32, 4
G = int(np.prod(dims)); M = 7
g = torch.Generator().manual_seed(1)
coeffs = torch.zeros(B, G, M, dtype=torch.float64); coeffs[..., 0] = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
coeffs[..., 1] = 1; coeffs[..., 5] = -1; coeffs[..., 6] = -1
steps = [np.full((B, n - 1), h) for n, h in zip(dims, (0.1, 0.3906, 0.3906))]
sr = StageRunner(lib, "cuda:0", dims, IV_LISTS["gl"], B, n_grid, False, coeffs.numpy(), steps)
dev = torch.device("cuda:0")
n = B * G * M
torch.manual_seed(3)
x = torch.randn(n, dtype=torch.float64, device=dev); out = torch.zeros(n, dtype=torch.float64, device=dev)
cfg = sr.plan.cfg(False)
def call():
    lib.check(lib.dll.pdeop_stage(sr.plan.handle, ctypes.byref(cfg), _lib.STAGE_APPLY_K, 0, 0, _lib._ptr(x), None, _lib._ptr(out),
                                  _lib._ptr(sr.persist), _lib._ptr(sr.scratch), _lib.current_stream_ptr(dev)))
call(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): call()
e1.record(); torch.cuda.synchronize()
print(f"flags [{flags}]: stage(APPLY_K) fine level incl. pack + unpack kernels: {e0.elapsed_time(e1)/10*1e3:.1f} us per call; checksum {float(out.abs().sum()):.9e}")
