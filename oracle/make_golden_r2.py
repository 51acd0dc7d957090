"""Round-2 golden fixtures from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Adds the cases the round-1 review asked for: 3-D multigrid with three levels and downsample_first=False (the
reference's own Ginzburg-Landau default, discovery/ginzburg_landau.py:52-57,241-243, and one size up), the
dense 32x32 sine layer of fit/sine_pde_dense.py:106-119 (BASELINE config 1), a four-level 2-D Burgers grid, and
a Kamani-shaped batch (discovery/kamani.py:153-165).

    cd /tmp/scratch && PYTHONPATH=/root/reference:/root/repo/oracle/stubs \
        python /root/repo/oracle/make_golden_r2.py /root/repo/tests/golden [case ...]

Large outputs are stored as strided samples (every `stride`-th entry of the flattened array) plus the 2-norm of
the full array, so each file stays well below 5 MB; inputs are stored in full.
"""
import os
import sys
import time

import numpy as np
import torch

OUT = sys.argv[1] if len(sys.argv) > 1 else "/root/repo/tests/golden"
ONLY = set(sys.argv[2:])

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle.make_golden as MG  # noqa: E402  (defines the helpers; its __main__ block does not run)
from oracle.cases import IV_LISTS, make_inputs  # noqa: E402
from solver.multigrid import MultigridLayer  # noqa: E402
from solver.pde_layer_dense import PDEDenseLayer  # noqa: E402

MG.OUT = OUT


def sample(a, stride):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return a[::stride].copy()


def golden_stages_seeded(name, layer, inp, seed, stride):
    """V-cycle, Gauss-Seidel, normal matvec and transfers of a large case.  The input vectors are NOT stored: the
    test regenerates them with torch.randn(n, generator=manual_seed(seed)) in the same order (v, x0) and checks
    the stored norm and leading entries; outputs are strided samples + norms."""
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
        cA, cr = mg.fill_coarse_grids(coeffs, rhs, iv, steps)          # qp_dual_sparse_multigrid_normal_kkt.py:28-47
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
        gs5 = mg.smooth_gs(AL[0], AU[0], v.numpy(), x0.numpy(), nsteps=5)
        vcyc = mg.v_cycle_gs_start(AtA_list, v, AL, AU, L, n_step=1, back=False)
        save = dict(seed=seed, stride=stride, n=n, v_norm=float(v.norm()), x0_norm=float(x0.norm()),
                    v_head=MG.t2n(v[:8]), x0_head=MG.t2n(x0[:8]))
        for k, a in (("Atb", MG.t2n(Atb)), ("Kv", MG.t2n(Kv)), ("gs5", np.asarray(gs5)), ("vcycle", MG.t2n(vcyc))):
            save[k] = sample(a, stride)
            save[k + "_norm"] = float(np.linalg.norm(a))
    np.savez_compressed(os.path.join(OUT, f"stagesS_{name}.npz"), **save)
    print("stagesS", name, flush=True)


def golden_layer_sampled(name, kind, dims, iv_name, bs, seed, n_grid=2, dsf=True, uniform=False, stride=1,
                         stages_seed=None, stages_stride=7):
    """Like make_golden.golden_layer, run once under default dtype float64 (exact d_rhs, no fp32 add_pad buffer,
    lp_pde_central_diff.py:1634); outputs strided."""
    if ONLY and name not in ONLY:
        return
    iv_list = IV_LISTS[iv_name]
    t = time.time()
    if kind == "dense":
        layer = PDEDenseLayer(bs=bs, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, init_index_mi_list=iv_list,
                              n_iv_steps=1, double_ret=True, solver_dbl=True)
    else:
        layer = MultigridLayer(bs=bs, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=n_grid, evolution=False,
                               downsample_first=dsf, init_index_mi_list=iv_list, n_iv_steps=1, double_ret=True,
                               solver_dbl=True)
    t_build = time.time() - t
    n_init = layer.pde.num_added_initial_constraints
    inp = make_inputs(dims, bs, n_init, seed, uniform=uniform)
    torch.set_default_dtype(torch.float64)
    try:
        t1 = time.time()
        out = MG.run_layer(layer, inp, inp["loss_w"])
        t_run = time.time() - t1
    finally:
        torch.set_default_dtype(torch.float32)
    save = dict(kind=kind, dims=np.array(dims), iv_name=iv_name, bs=bs, seed=seed, n_grid=n_grid, dsf=dsf,
                uniform=uniform, stride=stride, coeffs=inp["coeffs"], rhs=inp["rhs"], iv_rhs=inp["iv_rhs"],
                loss_seed=seed, ref_build_s=t_build, ref_fwd_bwd_s=t_run)
    if stride == 1:
        save["loss_w"] = inp["loss_w"]
    # with stride > 1 the loss weights are regenerated from the seed by oracle.cases.make_inputs (norm stored)
    save["loss_w_norm"] = float(np.linalg.norm(inp["loss_w"]))
    for c, s in enumerate(inp["steps"]):
        save[f"steps{c}"] = s
    for k, v in out.items():
        if k == "info" or k.startswith("d_steps") or k == "d_iv_rhs":
            save[k] = v
        else:
            save[k] = sample(v, stride)
            save[k + "_norm"] = float(np.linalg.norm(v))
    np.savez_compressed(os.path.join(OUT, f"layer_{name}.npz"), **save)
    print("layer", name, "build %.1fs run %.1fs" % (t_build, t_run), out.get("info"), flush=True)
    if stages_seed is not None:
        if stride == 1 and np.prod(dims) * bs < 16384:
            MG.golden_mg_stages(name, layer, inp, stages_seed)
        else:
            golden_stages_seeded(name, layer, inp, stages_seed, stages_stride)


if __name__ == "__main__":
    torch.manual_seed(0)
    # dense: Kamani-shaped batch and the 32x32 sine layer (BASELINE configs 2 and 1)
    golden_layer_sampled("dense_1d_24_b256", "dense", (24,), "kamani", 256, 51)
    golden_layer_sampled("dense_2d_32x32_sine", "dense", (32, 32), "sine", 1, 52, uniform=True)
    golden_layer_sampled("dense_2d_32x32_sine_nonuniform", "dense", (32, 32), "sine", 1, 53)
    # 2-D four levels
    golden_layer_sampled("mg_2d_64x64_g4", "mg", (64, 64), "burgers", 2, 54, n_grid=4, dsf=True, stages_seed=64)
    # 3-D three levels, time axis kept: the reference's own GL default, then one size up
    golden_layer_sampled("mg_3d_8x32x32_g3_nodsf", "mg", (8, 32, 32), "gl", 2, 55, n_grid=3, dsf=False,
                         stages_seed=65)
    golden_layer_sampled("mg_3d_16x32x32_g3_nodsf", "mg", (16, 32, 32), "gl", 2, 56, n_grid=3, dsf=False, stride=3)
