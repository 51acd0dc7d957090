"""GS-only micro harness (profiling aid): operator set-up for the GL grid, then a few 5-sweep smoothing calls
timed with CUDA events.  Usage: python tools/gs_micro.py [dims...] e.g. 32 64 64"""
import os, sys, ctypes, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import _lib
from mech_nn_discovery_pde_b200.ops import PdePlan
from tests.helpers import StageRunner
from oracle.cases import IV_LISTS
dims = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 64, 64)
B = int(os.environ.get("B", "32")); reps = int(os.environ.get("REPS", "5"))
n_grid = int(os.environ.get("NGRID", "4"))
flags = os.environ.get('PDEOP_VARIANT_FLAGS', '')
if flags:   # A/B testing: build and use a variant library with extra nvcc flags
    from mech_nn_discovery_pde_b200.build import build_debug
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libpdeop_variant.so')
    build_debug(so, flags.split())
    lib = _lib.PdeopLibrary(so)
else:
    lib = _lib.get_library()
G = int(np.prod(dims)); M = 7
g = torch.Generator().manual_seed(1)
coeffs = torch.zeros(B, G, M, dtype=torch.float64); coeffs[..., 0] = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
coeffs[..., 1] = 1; coeffs[..., 5] = -1; coeffs[..., 6] = -1
steps = [np.full((B, n - 1), h) for n, h in zip(dims, (0.1, 0.3906, 0.3906))]
sr = StageRunner(lib, "cuda:0", dims, IV_LISTS["gl"], B, n_grid, False, coeffs.numpy(), steps)
dev = torch.device("cuda:0")
n = B * G * M
b = torch.randn(n, dtype=torch.float64, device=dev); x = torch.zeros(n, dtype=torch.float64, device=dev)
out = torch.zeros(n, dtype=torch.float64, device=dev)
cfg = sr.plan.cfg(False)
def call():
    lib.check(lib.dll.pdeop_stage(sr.plan.handle, ctypes.byref(cfg), _lib.STAGE_GS, 0, 5, _lib._ptr(b), _lib._ptr(x), _lib._ptr(out),
                                  _lib._ptr(sr.persist), _lib._ptr(sr.scratch), _lib.current_stream_ptr(dev)))
call(); torch.cuda.synchronize()
sr.plan.profile_enable(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"[{flags}] dims {dims} B {B}: stage(GS x5) incl. 2 pack + 1 unpack kernels: {ms:.3f} ms per call")
