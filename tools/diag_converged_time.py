"""Diagnostic: where a converged-mode PCG iteration spends its time (per-plan instrumentation)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import MultigridLayer, PDEConfig
from oracle import pde_oracle as O
from oracle.cases import IV_LISTS, make_inputs

dims, ivn, n_grid, dsf, B = (16, 16, 16), "gl", 2, True, 3
for iters in (50, 200):
    class Cfg(PDEConfig):
        solver_mode = "converged"
        mg_pcg_rtol = 1e-30
        mg_pcg_max_iter = iters
    iv = IV_LISTS[ivn]
    st = O.build_structure(dims, iv)
    inp = make_inputs(dims, B, st.n_init, seed=123)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=n_grid, downsample_first=dsf,
                           init_index_mi_list=iv, n_iv_steps=1)
    layer.config = Cfg
    args = (t(inp["coeffs"]), t(inp["rhs"]), t(inp["iv_rhs"]), [t(s) for s in inp["steps"]])
    layer(*args); torch.cuda.synchronize()
    plan = layer.mg_solver.plan
    plan.profile_enable(True)
    t0 = time.time()
    layer(args[0], args[1], args[2], list(args[3])); torch.cuda.synchronize()
    dt = time.time() - t0
    prof = plan.profile_collect(); plan.profile_enable(False)
    print("iters", iters, "wall s", round(dt, 3), {k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]}, flush=True)
