"""Debug experiment: where does a chain-solver step spend its cycles (consumer warp 0 of CTA 0)?"""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import _lib
from mech_nn_discovery_pde_b200.build import build_debug
from tests.helpers import StageRunner
from oracle.cases import IV_LISTS
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpdeop_timing.so")
if not os.path.exists(so):
    build_debug(so, ["-DPDEOP_GS_TIMING"])
lib = _lib.PdeopLibrary(so)
lib.dll.pdeop_gs_dbg_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
dev = torch.device("cuda:0")
dims = (32, 64, 64); M = 7; G = int(np.prod(dims))
for B in (4, 32):
    g = torch.Generator().manual_seed(1)
    coeffs = torch.zeros(B, G, M, dtype=torch.float64); coeffs[..., 0] = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
    coeffs[..., 1] = 1; coeffs[..., 5] = -1; coeffs[..., 6] = -1
    steps = [np.full((B, n - 1), h) for n, h in zip(dims, (0.1, 0.3906, 0.3906))]
    sr = StageRunner(lib, "cuda:0", dims, IV_LISTS["gl"], B, 4, False, coeffs.numpy(), steps)
    nc = sr.level_n(3)
    b = torch.randn(B * nc, dtype=torch.float64, device=dev); out = torch.zeros_like(b)
    cfg = sr.plan.cfg(False)
    def call():
        lib.check(lib.dll.pdeop_stage(sr.plan.handle, ctypes.byref(cfg), _lib.STAGE_COARSE_SOLVE, 3, 0, _lib._ptr(b), None, _lib._ptr(out),
                                      _lib._ptr(sr.persist), _lib._ptr(sr.scratch), _lib.current_stream_ptr(dev)))
    call(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8)()
    lib.dll.pdeop_gs_dbg_read(buf)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    lib.dll.pdeop_gs_dbg_read(buf)
    st, tO, tR, tN, tF = [int(buf[i]) for i in range(5)]
    print(f"B={B}: {e0.elapsed_time(e1)/reps:.3f} ms/solve | per chain step (cycles): old phase {tO/st:.0f}, wait previous block {tR/st:.0f}, "
          f"wait new-block buffer {tN/st:.0f}, new phase + reduce + deliver {tF/st:.0f}")
    del sr
