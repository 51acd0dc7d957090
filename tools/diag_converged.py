"""Diagnostic: converged mode (per-instance PCG) on a 3-D case under several smoother settings."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mech_nn_discovery_pde_b200 import MultigridLayer, PDEConfig, _lib
from oracle import pde_oracle as O
from oracle.cases import IV_LISTS, make_inputs

def run(dims, ivn, n_grid, dsf, **kw):
    class Cfg(PDEConfig):
        solver_mode = "converged"
        mg_pcg_rtol = 1e-8
        mg_pcg_max_iter = 4000
    for k, v in kw.items():
        setattr(Cfg, k, v)
    B = 3
    iv = IV_LISTS[ivn]
    st = O.build_structure(dims, iv)
    inp = make_inputs(dims, B, st.n_init, seed=123)
    inp["rhs"][1] *= 1e-3
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=n_grid, downsample_first=dsf,
                           init_index_mi_list=iv, n_iv_steps=1)
    layer.config = Cfg
    u0, u, _ = layer(t(inp["coeffs"]), t(inp["rhs"]), t(inp["iv_rhs"]), [t(s) for s in inp["steps"]])
    torch.cuda.synchronize()
    f, b = layer.solver_info()
    print(dims, ivn, n_grid, dsf, kw, "->", f, flush=True)

for kw in (dict(mg_power_iters=60), dict(mg_cheb_ratio=10.0, mg_power_iters=60), dict(mg_power_iters=60, mg_smoother_sweeps=4),
           dict(mg_power_iters=60, mg_smoother="jacobi")):
    run((8, 16, 16), "gl", 2, False, **kw)
run((16, 16, 16), "gl", 2, True, mg_power_iters=60)
run((16, 16), "burgers", 2, True, mg_power_iters=60)
run((16, 16), "burgers", 2, True)
