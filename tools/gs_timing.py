"""Debug experiment: where does a GS step spend its cycles?  Builds a -DPDEOP_GS_TIMING copy of the library into
tools/, runs the GS micro harness with it and prints per-step cycle averages of thread 0 / CTA 0."""
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
for dims, n_grid in (((32, 64, 64), 4), ((32, 32, 32), 3), ((32, 16, 16), 2)):
    B = 32
    G = int(np.prod(dims)); M = 7
    g = torch.Generator().manual_seed(1)
    coeffs = torch.zeros(B, G, M, dtype=torch.float64); coeffs[..., 0] = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
    coeffs[..., 1] = 1; coeffs[..., 5] = -1; coeffs[..., 6] = -1
    steps = [np.full((B, n - 1), h) for n, h in zip(dims, (0.1, 0.3906, 0.3906))]
    sr = StageRunner(lib, "cuda:0", dims, IV_LISTS["gl"], B, n_grid, False, coeffs.numpy(), steps)
    n = B * G * M
    b = torch.randn(n, dtype=torch.float64, device=dev); x = torch.zeros(n, dtype=torch.float64, device=dev); out = torch.zeros_like(b)
    cfg = sr.plan.cfg(False)
    def call():
        lib.check(lib.dll.pdeop_stage(sr.plan.handle, ctypes.byref(cfg), _lib.STAGE_GS, 0, 5, _lib._ptr(b), _lib._ptr(x), _lib._ptr(out),
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
    st, cA, cB, cW, npt = [int(buf[i]) for i in range(5)]
    print(f"dims {dims}: {e0.elapsed_time(e1)/reps:.3f} ms/call | thread0 per step: A {cA/st:.0f}  B {cB/st:.0f}  wait {cW/st:.0f} cycles; "
          f"steps {st//reps}, point updates per step {npt/st:.2f}")
