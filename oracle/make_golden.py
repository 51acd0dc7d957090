"""Generate golden fixtures from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place /root/reference exists):

    cd /tmp/scratch && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/oracle/make_golden.py /root/repo/tests/golden

The reference is imported with three stub modules (ipdb, cupy, cupyx.scipy.sparse[.linalg] ->
numpy/scipy; see oracle/stubs).  Inputs are seeded; inputs and reference outputs are stored as
compressed .npz so the tests (CPU here, GPU box later) never need the reference.

Run from a scratch cwd: importing the reference's QP modules creates ./logs/misc/<n>/
(extras/source.py:7-21).
"""
import os
import sys
import time

import numpy as np
import torch

OUT = sys.argv[1] if len(sys.argv) > 1 else "/root/repo/tests/golden"
os.makedirs(OUT, exist_ok=True)

from solver.lp_pde_central_diff import PDESYSLP, ConstraintType  # noqa: E402
from solver.pde_layer_dense import PDEDenseLayer  # noqa: E402
from solver.multigrid import MultigridLayer  # noqa: E402
import solver.fgmres as FG  # noqa: E402
from config import PDEConfig  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.cases import IV_LISTS, make_inputs  # noqa: E402


def t2n(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------
# 1. structure + per-call values
# ---------------------------------------------------------------------------------------------
def golden_structure(name, dims, iv_name, seed):
    iv_list = IV_LISTS[iv_name]
    pde = PDESYSLP(bs=2, coord_dims=dims, n_iv=1, init_index_mi_list=iv_list, n_auxiliary=0, n_equations=1,
                   step_size=0.01, order=2, evolution=False, dtype=torch.float64, n_iv_steps=1)
    g = torch.Generator().manual_seed(seed)
    steps = [(0.05 + 0.2 * torch.rand(2, n - 1, generator=g, dtype=torch.float64)) for n in dims]
    dv = pde.build_derivative_values(steps)
    hu = 0.125
    steps_u = [torch.full((2, n - 1), hu, dtype=torch.float64) for n in dims]
    dv_u = pde.build_derivative_values(steps_u)
    np.savez_compressed(
        os.path.join(OUT, f"struct_{name}.npz"),
        dims=np.array(dims), iv_name=iv_name,
        eq_row=np.array(pde.row_dict[ConstraintType.Equation]), eq_col=np.array(pde.col_dict[ConstraintType.Equation]),
        init_row=np.array(pde.row_dict[ConstraintType.Initial]), init_col=np.array(pde.col_dict[ConstraintType.Initial]),
        d_row=np.array(pde.row_dict[ConstraintType.Derivative]), d_col=np.array(pde.col_dict[ConstraintType.Derivative]),
        d_static_val=np.array(pde.value_dict[ConstraintType.Derivative], dtype=np.float64),
        static_step=np.array(0.01),
        n_eq=pde.num_added_equation_constraints, n_init=pde.num_added_initial_constraints,
        n_deriv=pde.num_added_derivative_constraints,
        steps=np.concatenate([t2n(s).reshape(2, -1) for s in steps], axis=1), dvals=t2n(dv),
        uniform_step=np.array(hu), dvals_uniform=t2n(dv_u),
    )
    print("struct", name, "rows", pde.num_constraints)


# ---------------------------------------------------------------------------------------------
# 2. dense layer fwd + bwd
# ---------------------------------------------------------------------------------------------
def run_layer(layer, inp, loss_w):
    coeffs = torch.tensor(inp["coeffs"]).requires_grad_(True)
    rhs = torch.tensor(inp["rhs"]).requires_grad_(True)
    iv = torch.tensor(inp["iv_rhs"]).requires_grad_(True)
    steps = [torch.tensor(s).requires_grad_(True) for s in inp["steps"]]
    infos = []
    orig = FG.fgmres_matvec

    def wrap(*a, **k):
        x, info = orig(*a, **k)
        infos.append((int(info[0]), float(info[1])))
        return x, info
    FG.fgmres_matvec = wrap
    try:
        u0, u, _ = layer(coeffs, rhs, iv, list(steps))
        loss = (u * torch.tensor(loss_w).reshape(u.shape)).sum()
        loss.backward()
    finally:
        FG.fgmres_matvec = orig
    out = dict(u=t2n(u), d_coeffs=t2n(coeffs.grad), d_rhs=t2n(rhs.grad), d_iv_rhs=t2n(iv.grad))
    for c, s in enumerate(steps):
        out[f"d_steps{c}"] = t2n(s.grad)
    if infos:
        out["info"] = np.array(infos, dtype=np.float64)
    return out


def golden_layer(name, kind, dims, iv_name, bs, seed, n_grid=2, dsf=True, uniform=False):
    iv_list = IV_LISTS[iv_name]
    t = time.time()
    if kind == "dense":
        layer = PDEDenseLayer(bs=bs, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, init_index_mi_list=iv_list,
                              n_iv_steps=1, double_ret=True, solver_dbl=True)
    else:
        layer = MultigridLayer(bs=bs, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=n_grid, evolution=False,
                               downsample_first=dsf, init_index_mi_list=iv_list, n_iv_steps=1, double_ret=True,
                               solver_dbl=True)
    n_init = layer.pde.num_added_initial_constraints
    inp = make_inputs(dims, bs, n_init, seed, uniform=uniform)
    # shipped behaviour: add_pad allocates a float32 buffer (lp_pde_central_diff.py:1634) -> d_rhs rounded to fp32
    out_q = run_layer(layer, inp, inp["loss_w"])
    # same run with the default dtype raised so that buffer is fp64: exact d_rhs
    torch.set_default_dtype(torch.float64)
    try:
        out = run_layer(layer, inp, inp["loss_w"])
    finally:
        torch.set_default_dtype(torch.float32)
    out["d_rhs_fp32quirk"] = out_q["d_rhs"]
    save = dict(kind=kind, dims=np.array(dims), iv_name=iv_name, bs=bs, seed=seed, n_grid=n_grid, dsf=dsf,
                coeffs=inp["coeffs"], rhs=inp["rhs"], iv_rhs=inp["iv_rhs"], loss_w=inp["loss_w"])
    for c, s in enumerate(inp["steps"]):
        save[f"steps{c}"] = s
    save.update(out)
    np.savez_compressed(os.path.join(OUT, f"layer_{name}.npz"), **save)
    print("layer", name, "done in %.1fs" % (time.time() - t), out.get("info"))
    return layer, inp


# ---------------------------------------------------------------------------------------------
# 3. multigrid stages: GS sweep, residual, restrict, prolong, V-cycle, on random vectors
# ---------------------------------------------------------------------------------------------
def golden_mg_stages(name, layer, inp, seed):
    mg = layer.mg_solver
    pde = layer.pde
    bs = layer.bs
    coeffs = torch.tensor(inp["coeffs"]).reshape(bs, layer.grid_size, layer.n_orders)
    rhs = torch.tensor(inp["rhs"]).reshape(bs, layer.grid_size)
    iv = torch.tensor(inp["iv_rhs"]).reshape(bs, -1)
    steps = [torch.tensor(s).reshape(bs, -1) for s in inp["steps"]]
    with torch.no_grad():
        dc = pde.build_derivative_tensor(steps)
        ec = pde.build_equation_tensor(coeffs)
        # qp_dual_sparse_multigrid_normal_kkt.py:28-47
        cA, cr = mg.fill_coarse_grids(coeffs, rhs, iv, steps)
        A, A_rhs = pde.fill_block_constraints_torch(ec, rhs, iv, dc)
        AtA, D, Atb, A_L, A_U = mg.make_AtA(pde, A, A_rhs)
        AtA_list, rhs_list, D_list, L_list, U_list = mg.make_coarse_AtA_matrices(cA, cr)
        AtA_list = [AtA] + AtA_list
        AL = [A_L] + L_list
        AU = [A_U] + U_list
        L = mg.factor_coarsest(AtA_list[-1].to_dense())
        n = AtA.shape[0]
        g = torch.Generator().manual_seed(seed)
        v = torch.randn(n, generator=g, dtype=torch.float64)
        x0 = torch.randn(n, generator=g, dtype=torch.float64)
        Kv = torch.mm(AtA, v.unsqueeze(1)).squeeze(1)
        gs1 = mg.smooth_gs(AL[0], AU[0], v.numpy(), x0.numpy(), nsteps=1)
        gs3 = mg.smooth_gs(AL[0], AU[0], v.numpy(), x0.numpy(), nsteps=3)
        rst = mg.restrict(0, v.numpy())
        nc = rst.shape[0]
        vc = torch.randn(nc, generator=g, dtype=torch.float64)
        pro = mg.prolong(1, vc.numpy())
        vcyc = mg.v_cycle_gs_start(AtA_list, v, AL, AU, L, n_step=1, back=False)
        save = dict(v=t2n(v), x0=t2n(x0), vc=t2n(vc), Atb=t2n(Atb), Kv=t2n(Kv), gs1=np.asarray(gs1), gs3=np.asarray(gs3),
                    restrict=np.asarray(rst), prolong=np.asarray(pro), vcycle=t2n(vcyc))
        if len(AtA_list) > 2:
            K1 = AtA_list[1]
            v1 = torch.randn(K1.shape[0], generator=g, dtype=torch.float64)
            save["v1"] = t2n(v1)
            save["K1v1"] = t2n(torch.mm(K1, v1.unsqueeze(1)).squeeze(1))
        Kc = AtA_list[-1].to_dense()
        vcs = torch.randn(Kc.shape[0], Kc.shape[1], generator=g, dtype=torch.float64)
        save["vcs"] = t2n(vcs)
        save["Kc_vcs"] = t2n(torch.bmm(Kc, vcs.unsqueeze(2)).squeeze(2))
    np.savez_compressed(os.path.join(OUT, f"stages_{name}.npz"), **save)
    print("stages", name)


if __name__ == "__main__":
    torch.manual_seed(0)
    golden_structure("1d_24", (24,), "kamani", 11)
    golden_structure("2d_8x10", (8, 10), "burgers", 12)
    golden_structure("2d_9x8_sine", (9, 8), "sine", 13)
    golden_structure("3d_8x9x10", (8, 9, 10), "gl", 14)

    golden_layer("dense_1d_24", "dense", (24,), "kamani", 3, 21)
    golden_layer("dense_2d_8x10", "dense", (8, 10), "burgers", 2, 22)
    golden_layer("dense_2d_12x12_sine_uniform", "dense", (12, 12), "sine", 1, 23, uniform=True)
    golden_layer("dense_3d_8x8x8", "dense", (8, 8, 8), "gl", 1, 24)

    lay, inp = golden_layer("mg_2d_16x16_g2", "mg", (16, 16), "burgers", 2, 31, n_grid=2, dsf=True)
    golden_mg_stages("mg_2d_16x16_g2", lay, inp, 41)
    lay, inp = golden_layer("mg_2d_32x32_g3", "mg", (32, 32), "burgers", 2, 32, n_grid=3, dsf=True)
    golden_mg_stages("mg_2d_32x32_g3", lay, inp, 42)
    lay, inp = golden_layer("mg_2d_16x32_g2_nodsf", "mg", (16, 32), "transport", 3, 33, n_grid=2, dsf=False)
    lay, inp = golden_layer("mg_3d_8x16x16_g2_nodsf", "mg", (8, 16, 16), "gl", 2, 34, n_grid=2, dsf=False)
    golden_mg_stages("mg_3d_8x16x16_g2_nodsf", lay, inp, 44)
    lay, inp = golden_layer("mg_3d_16x16x16_g2", "mg", (16, 16, 16), "gl", 1, 35, n_grid=2, dsf=True, uniform=True)
