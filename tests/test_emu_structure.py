"""CPU-only checks of the host-side logic (plan index tables, axis tables of K, wavefront Gauss-Seidel
ordering, V-cycle / FGMRES orchestration, gradient formulas, Python layers) by running the shared
per-element bodies through the test-only host emulator and comparing with the reference's golden vectors.
The CUDA kernels call the same bodies; their parity proper is tested on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from mech_nn_discovery_pde_b200 import _lib
from oracle.cases import IV_LISTS
from tests.emu.emu_lib import emu_library
from tests.helpers import GOLDEN, StageRunner, load_layer_case, rel, run_layer_case

import os

STAGE_CASES = ["mg_2d_16x16_g2", "mg_2d_32x32_g3", "mg_3d_8x16x16_g2_nodsf"]


@pytest.mark.parametrize("name", STAGE_CASES)
def test_stages_vs_reference(name):
    lib = emu_library()
    s = np.load(os.path.join(GOLDEN, f"stages_{name}.npz"))
    z, dims, steps = load_layer_case(name)
    B = int(z["bs"])
    sr = StageRunner(lib, "cpu", dims, IV_LISTS[str(z["iv_name"])], B, int(z["n_grid"]), bool(z["dsf"]), z["coeffs"],
                     steps)
    assert int(sr.info[3]) == 0
    assert rel(sr.stage(_lib.STAGE_ATB, 0, z["rhs"], z["iv_rhs"]), s["Atb"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_APPLY_K, 0, s["v"]), s["Kv"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=1), s["gs1"]) < 1e-12
    gs3 = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=3)
    assert rel(gs3, s["gs3"]) < 1e-12
    # pipelined sweeps are order independent inside a step: identical bits either way
    lib.dll.pdeop_emu_set_gs_reverse(1)
    try:
        gs3r = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=3)
    finally:
        lib.dll.pdeop_emu_set_gs_reverse(0)
    assert np.array_equal(gs3, gs3r)
    assert rel(sr.stage(_lib.STAGE_RESTRICT, 0, s["v"], out_level=1), s["restrict"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_PROLONG, 1, s["vc"], out_level=0), s["prolong"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_VCYCLE, 0, s["v"]), s["vcycle"]) < 1e-9
    if "v1" in s.files:
        assert rel(sr.stage(_lib.STAGE_APPLY_K, 1, s["v1"]), s["K1v1"]) < 1e-13


LAYER_CASES = ["dense_1d_24", "dense_2d_8x10", "dense_2d_12x12_sine_uniform", "mg_2d_16x16_g2",
               "mg_2d_16x32_g2_nodsf"]


@pytest.mark.parametrize("name", ["dense_2d_8x10", "mg_2d_16x16_g2"])
def test_sparse_tensor_adapter(name):
    """QPFunction fed with torch.sparse constraint tensors (the reference's argument types,
    qp_dual_sparse_multigrid_normal_kkt.py:25-33): same solution and the same gradients w.r.t. coeffs / rhs / iv_rhs /
    steps as the dense-carrier path and the reference."""
    lib = emu_library()
    z, out = run_layer_case(lib, "cpu", name, sparse=True)
    z2, ref = run_layer_case(lib, "cpu", name, sparse=False)
    assert rel(out["u"], ref["u"]) < 1e-14
    assert rel(out["d_coeffs"], ref["d_coeffs"]) < 1e-13
    assert rel(out["d_rhs"], ref["d_rhs"]) < 1e-13
    for c in range(len(out["d_steps"])):
        assert rel(out["d_steps"][c], ref["d_steps"][c]) < 1e-12
        assert rel(out["d_steps"][c], z[f"d_steps{c}"]) < 1e-6


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_vs_reference(name):
    lib = emu_library()
    z, out = run_layer_case(lib, "cpu", name)
    dense = str(z["kind"]) == "dense"
    tol_x, tol_g = (1e-8, 2e-7) if dense else (1e-8, 1e-8)
    assert rel(out["u"], z["u"]) < tol_x
    assert rel(out["d_coeffs"], z["d_coeffs"]) < tol_g
    assert rel(out["d_rhs"], z["d_rhs"]) < tol_g
    assert rel(out["d_iv_rhs"], z["d_iv_rhs"]) < tol_g
    for c in range(len(out["d_steps"])):
        assert rel(out["d_steps"][c], z[f"d_steps{c}"]) < max(tol_g, 1e-7)
    if not dense:
        info = z["info"]
        assert int(out["info_fwd"][0]) == int(info[0, 0]) and int(out["info_bwd"][0]) == int(info[1, 0])
        assert abs(out["info_fwd"][1] - info[0, 1]) <= 1e-6 * info[0, 1]
        assert abs(out["info_bwd"][1] - info[1, 1]) <= 1e-6 * info[1, 1]


def test_1d_multigrid_vs_oracle():
    """1-D grids go through the same kernels (D=1 template); no golden from the reference scripts, so the
    oracle (pinned on 2-D/3-D goldens and 1-D dense goldens) is the checker."""
    import torch
    from mech_nn_discovery_pde_b200 import MultigridLayer
    from oracle import pde_oracle as O
    from oracle.cases import make_inputs
    lib = emu_library()
    dims, B, n_grid = (32,), 2, 2
    iv = IV_LISTS["kamani"]
    st = O.build_structure(dims, iv)
    inp = make_inputs(dims, B, st.n_init, seed=5)
    g_out = inp["loss_w"].reshape(B, -1)
    ref = O.mg_layer(dims, iv, inp["coeffs"], inp["rhs"], inp["iv_rhs"], inp["steps"], n_grid, True, grad_out=g_out)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=n_grid, downsample_first=True,
                           init_index_mi_list=iv, n_iv_steps=1, _library=lib)
    coeffs = t(inp["coeffs"]).requires_grad_(True)
    rhs = t(inp["rhs"]).requires_grad_(True)
    ivr = t(inp["iv_rhs"]).requires_grad_(True)
    steps = [t(s).requires_grad_(True) for s in inp["steps"]]
    u0, u, _ = layer(coeffs, rhs, ivr, list(steps))
    (u * t(inp["loss_w"]).reshape(u.shape)).sum().backward()
    assert rel(u.detach().numpy().reshape(B, -1), ref.x) < 1e-8
    assert rel(coeffs.grad.numpy(), ref.d_coeffs) < 1e-8
    assert rel(rhs.grad.numpy(), ref.d_rhs) < 1e-8
    assert rel(steps[0].grad.numpy(), ref.d_steps[0]) < 1e-7


def test_rhs_grad_fp32_quirk_and_knobs():
    """PDEConfig knobs are read at call time; the fp32 add_pad quirk of the reference is reproducible on request."""
    from mech_nn_discovery_pde_b200.config import PDEConfig

    class Cfg(PDEConfig):
        rhs_grad_fp32_quirk = True
    lib = emu_library()
    z, out = run_layer_case(lib, "cpu", "mg_2d_16x16_g2", config=Cfg)
    assert np.array_equal(out["d_rhs"], out["d_rhs"].astype(np.float32).astype(np.float64))
    assert rel(out["d_rhs"], z["d_rhs_fp32quirk"]) < 1e-7

    class Short(PDEConfig):
        mg_fgmres_max_iter_forward = 20
        mg_fgmres_max_iter_backward = 10
    z, out = run_layer_case(lib, "cpu", "mg_2d_16x16_g2", config=Short)
    assert int(out["info_fwd"][0]) == 20 and int(out["info_bwd"][0]) == 10


def test_qpfunction_factory_surface():
    """The reference's QPFunction factories and pde.build_*_tensor helpers: gradients go to the constraint values,
    rhs and iv_rhs, none to coeffs / steps_list (qp_dual_sparse_multigrid_normal_kkt.py:162)."""
    import torch
    from mech_nn_discovery_pde_b200 import MultigridLayer
    from mech_nn_discovery_pde_b200.solver.qp_dual_sparse_multigrid_normal_kkt import QPFunction
    lib = emu_library()
    z, dims, steps = load_layer_case("mg_2d_16x16_g2")
    B = int(z["bs"])
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=2, downsample_first=True,
                           init_index_mi_list=IV_LISTS["burgers"], n_iv_steps=1, _library=lib)
    pde, mg = layer.pde, layer.mg_solver
    qpf = QPFunction(pde, mg, 1)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    coeffs = t(z["coeffs"]).requires_grad_(True)
    steps_t = [t(s).requires_grad_(True) for s in steps]
    eq = pde.build_equation_tensor(coeffs.detach()).requires_grad_(True)
    assert eq.shape == (B, pde.num_added_equation_constraints, layer.n_orders)
    dc = tuple(v.detach().requires_grad_(True) for v in pde.build_derivative_tensor(steps_t))
    rhs = t(z["rhs"]).requires_grad_(True)
    x = qpf(eq, rhs, t(z["iv_rhs"]), dc, coeffs, steps_t)
    assert x.shape == (B, pde.var_set.num_vars)
    assert rel(x.detach().numpy().reshape(B, 1, -1, layer.n_orders), z["u"]) < 1e-8
    (x * t(z["loss_w"]).reshape(B, -1)).sum().backward()
    assert coeffs.grad is None and all(s.grad is None for s in steps_t)
    # dA restricted to the equation rows == reference's d_coeffs on those rows
    g = pde.equation_grid_pointers().numpy()
    assert rel(eq.grad.numpy(), z["d_coeffs"][:, g, :]) < 1e-8
    assert rel(rhs.grad.numpy(), z["d_rhs"]) < 1e-8
    # pad helpers round trip (lp_pde_central_diff.py:1632-1705)
    r = pde.remove_pad(rhs.detach(), coeffs=False)
    back = pde.add_pad(r).reshape(B, -1)
    mask = np.zeros(pde.var_set.grid_size, bool)
    mask[g] = True
    assert np.array_equal(back.numpy()[:, mask], rhs.detach().numpy()[:, mask]) and (back.numpy()[:, ~mask] == 0).all()


def test_n_ind_dim_batches_like_reference():
    """n_ind_dim > 1 multiplies the batch: (bs, n_ind_dim) instances share one block system (multigrid.py:589-605)."""
    import torch
    from mech_nn_discovery_pde_b200 import MultigridLayer
    lib = emu_library()
    z, dims, steps = load_layer_case("mg_2d_16x16_g2")   # batch 2 -> bs=1, n_ind_dim=2
    layer = MultigridLayer(bs=1, coord_dims=dims, order=2, n_ind_dim=2, n_iv=1, n_grid=2, downsample_first=True,
                           init_index_mi_list=IV_LISTS["burgers"], n_iv_steps=1, _library=lib)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    G, M = layer.grid_size, layer.n_orders
    u0, u, eps = layer(t(z["coeffs"]).reshape(1, 2, G, M), t(z["rhs"]).reshape(1, 2, G), t(z["iv_rhs"]).reshape(1, 2, -1),
                       [t(s).reshape(1, 2, -1) for s in steps])
    assert u.shape == (1, 2, G, M) and u0.shape == (1, 2, G) and eps is None
    assert rel(u.detach().numpy().reshape(2, 1, G, M), z["u"]) < 1e-8


def _fgmres_call(lib, device, plan, persist, scratch, b, cfg):
    import ctypes
    import torch
    dev = torch.device(device)
    bt = torch.as_tensor(b, dtype=torch.float64, device=dev).contiguous()
    x = torch.zeros_like(bt)
    info = torch.zeros(4, dtype=torch.float64, device=dev)
    hess = torch.zeros((cfg.restart + 1) * cfg.restart, dtype=torch.float64, device=dev)
    lib.check(lib.dll.pdeop_fgmres(plan.handle, ctypes.byref(cfg), 0, _lib._ptr(bt), _lib._ptr(x), _lib._ptr(info),
                                   _lib._ptr(hess), _lib._ptr(persist), _lib._ptr(scratch),
                                   _lib.current_stream_ptr(dev)))
    return x.cpu().numpy(), info.cpu().numpy(), hess.cpu().numpy().reshape(cfg.restart + 1, cfg.restart)


def check_fgmres_control_flow(lib, device):
    """Device-side convergence flag (fgmres.py:76-78,134): zero rhs returns zeros with 0 iterations; a loose
    tolerance stops after the first restart cycle; the Hessenberg matrix matches the oracle's."""
    from oracle import pde_oracle as O
    z, dims, steps = load_layer_case("mg_2d_16x16_g2")
    iv = IV_LISTS[str(z["iv_name"])]
    B = int(z["bs"])
    sr = StageRunner(lib, device, dims, iv, B, 2, True, z["coeffs"], steps)
    mg = O.mg_setup(dims, iv, z["coeffs"], z["rhs"], z["iv_rhs"], steps, 2, True)
    n = B * sr.plan.n
    cfg = sr.plan.cfg(False)
    x, info, _ = _fgmres_call(lib, device, sr.plan, sr.persist, sr.scratch, np.zeros(n), cfg)
    assert int(info[0]) == 0 and not x.any()
    rng = np.random.default_rng(0)
    b = mg.K_list[0] @ rng.standard_normal(n)       # consistent right-hand side
    # reference run, full length, with trace
    tr = {}
    xr, (it_r, rn_r) = O.fgmres(mg.K_list[0], b, lambda v: O.v_cycle_start(mg, v), restart=10, maxiter=40, trace=tr)
    x, info, H = _fgmres_call(lib, device, sr.plan, sr.persist, sr.scratch, b, cfg)
    assert int(info[0]) == it_r and abs(info[1] - rn_r) <= 1e-6 * rn_r
    assert rel(x, xr) < 1e-8
    assert rel(H, tr["H"][-1]) < 1e-7                      # Hessenberg of the last restart cycle
    # loose absolute tolerance: stop at the first residual check that passes (after k+1 restart cycles)
    rn = tr["r_norms"]
    k = next(i for i in range(len(rn) - 1) if rn[i + 1] < 0.9 * rn[i])
    cfg2 = sr.plan.cfg(False)
    cfg2.atol = float(0.5 * (rn[k] + rn[k + 1]))
    x1, info1, _ = _fgmres_call(lib, device, sr.plan, sr.persist, sr.scratch, b, cfg2)
    assert int(info1[0]) == 10 * (k + 1) and abs(info1[1] - rn[k + 1]) <= 1e-6 * rn[k + 1]
    x1r, (it1, _) = O.fgmres(mg.K_list[0], b, lambda v: O.v_cycle_start(mg, v), restart=10, maxiter=40,
                             atol=cfg2.atol)
    assert it1 == 10 * (k + 1) and rel(x1, x1r) < 1e-8


def test_fgmres_control_flow():
    check_fgmres_control_flow(emu_library(), "cpu")


SURFACE_CASES = ["dense_2d_8x10_order1", "mg_2d_16x16_order1", "dense_1d_24_nind3"]


def run_surface_case(lib, device, name):
    """Total order 1 and n_ind_dim > 1 through the layer surface, against the unmodified reference
    (oracle/make_golden_surface.py; lp_pde_central_diff.py:304-315, pde_layer_dense.py:83-125)."""
    import os
    import torch
    from oracle.cases import IV_LISTS
    from tests.helpers import GOLDEN
    from mech_nn_discovery_pde_b200 import MultigridLayer, PDEDenseLayer
    z = np.load(os.path.join(GOLDEN, f"surf_{name}.npz"))
    dims = tuple(int(v) for v in z["dims"])
    bs, n_ind, order = int(z["bs"]), int(z["n_ind"]), int(z["order"])
    iv = IV_LISTS[str(z["iv_name"])]
    dev = torch.device(device)
    kw = dict(bs=bs, coord_dims=dims, order=order, n_ind_dim=n_ind, n_iv=1, init_index_mi_list=iv, n_iv_steps=1,
              double_ret=True, solver_dbl=True, _library=lib)
    if str(z["kind"]) == "dense":
        layer = PDEDenseLayer(**kw)
    else:
        layer = MultigridLayer(n_grid=int(z["n_grid"]), downsample_first=True, **kw)
    G = int(np.prod(dims))
    M = layer.n_orders
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    coeffs = t(z["coeffs"]).reshape(bs, n_ind, G, M).requires_grad_(True)
    rhs = t(z["rhs"]).reshape(bs, n_ind, G).requires_grad_(True)
    ivr = t(z["iv_rhs"]).reshape(bs, n_ind, -1).requires_grad_(True)
    steps = [t(z[f"steps{c}"]).reshape(bs, n_ind, -1).requires_grad_(True) for c in range(len(dims))]
    u0, u, eps = layer(coeffs, rhs, ivr, list(steps))
    assert tuple(u.shape) == tuple(z["u"].shape) and tuple(u0.shape) == tuple(z["u0"].shape)
    (u * t(z["loss_w"]).reshape(u.shape)).sum().backward()
    tol_g = 2e-7 if str(z["kind"]) == "dense" else 1e-8
    assert rel(u.detach().cpu().numpy(), z["u"]) < 1e-8
    assert rel(coeffs.grad.cpu().numpy(), z["d_coeffs"]) < tol_g
    assert rel(rhs.grad.cpu().numpy(), z["d_rhs"]) < tol_g
    assert rel(ivr.grad.cpu().numpy(), z["d_iv_rhs"]) < tol_g
    for c in range(len(dims)):
        assert rel(steps[c].grad.cpu().numpy(), z[f"d_steps{c}"]) < 1e-7
    if "info" in z.files:
        f, _ = layer.solver_info()
        assert f[0] == int(z["info"][0, 0])


@pytest.mark.parametrize("name", SURFACE_CASES)
def test_surface_order1_and_n_ind_dim(name):
    run_surface_case(emu_library(), "cpu", name)


def _converged_case(lib, device, dims, iv_name, B, n_grid, dsf, smoother, seed=123, max_iter=2500):
    """Converged mode (per-instance PCG, symmetric V-cycle) against the EXACT least-squares solution: the dense
    layer's Cholesky solve of the same system (semantics of solver/cg.py:51-147)."""
    import torch
    from oracle import pde_oracle as O
    from oracle.cases import IV_LISTS, make_inputs
    from mech_nn_discovery_pde_b200 import MultigridLayer, PDEConfig

    class Cfg(PDEConfig):
        solver_mode = "converged"
        mg_pcg_rtol = 1e-8
        mg_pcg_max_iter = max_iter
        mg_smoother = smoother
        mg_smoother_sweeps = 8

    iv = IV_LISTS[iv_name]
    st = O.build_structure(dims, iv)
    inp = make_inputs(dims, B, st.n_init, seed=seed)
    inp["rhs"][1] *= 1e-3          # instances converge at different iterations: exercises the per-instance masks
    dev = torch.device(device)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=n_grid, downsample_first=dsf,
                           init_index_mi_list=iv, n_iv_steps=1, _library=lib)
    layer.config = Cfg
    coeffs = t(inp["coeffs"]).requires_grad_(True)
    rhs = t(inp["rhs"]).requires_grad_(True)
    ivr = t(inp["iv_rhs"]).requires_grad_(True)
    steps = [t(s).requires_grad_(True) for s in inp["steps"]]
    u0, u, _ = layer(coeffs, rhs, ivr, list(steps))
    (u * t(inp["loss_w"]).reshape(u.shape)).sum().backward()
    f, b = layer.solver_info()
    exact = O.dense_layer(dims, iv, inp["coeffs"], inp["rhs"], inp["iv_rhs"], inp["steps"],
                          grad_out=inp["loss_w"].reshape(B, -1))
    # every instance reached the relative tolerance before the cap, forward and backward
    assert f[0] < Cfg.mg_pcg_max_iter and f[1] <= Cfg.mg_pcg_rtol, f
    assert b[0] < Cfg.mg_pcg_max_iter and b[1] <= Cfg.mg_pcg_rtol, b
    x = u.detach().cpu().numpy().reshape(B, -1)
    for i in range(B):      # per instance: small instances are not hidden behind large ones
        assert rel(x[i], exact.x[i]) < 1e-6
    assert rel(coeffs.grad.cpu().numpy(), exact.d_coeffs) < 1e-5
    assert rel(rhs.grad.cpu().numpy(), exact.d_rhs) < 1e-5
    assert rel(ivr.grad.cpu().numpy(), exact.d_iv_rhs) < 1e-5
    return f, b


@pytest.mark.parametrize("smoother", ["chebyshev", "jacobi"])
def test_converged_mode_vs_exact_solution(smoother):
    _converged_case(emu_library(), "cpu", (16, 16), "burgers", 3, 2, True, smoother)


# dims, iv list, batch, n_grid, downsample_first, sweeps, threads per CTA of the model (0: the kernel's 512)
LINE_CASES = [
    ((16, 16, 16), "gl", 1, 2, True, 5, 0),      # one CTA per instance
    ((16, 16, 16), "gl", 1, 2, True, 3, 64),     # four CTAs of four rows: progress counters, parked row
    ((16, 16, 16), "gl", 1, 2, True, 1, 96),     # a single sweep; CTAs of six rows, the last one not full
    ((16, 20, 16), "gl", 1, 2, True, 2, 80),     # rows that are not a multiple of the warp size, four CTAs
    ((16, 16), "burgers", 2, 2, True, 5, 0),     # 2-D: one row
    ((32, 24), "burgers", 1, 2, True, 3, 0),
]
_LINE_RUNNERS = {}


def _line_runner(dims, ivn, B, n_grid, dsf):
    """Operator set-up on the emulator (the dense coarsest factor dominates it): once per geometry."""
    from oracle import pde_oracle as O
    from oracle.cases import make_inputs
    key = (dims, ivn, B, n_grid, dsf)
    if key not in _LINE_RUNNERS:
        iv = IV_LISTS[ivn]
        st = O.build_structure(dims, iv)
        inp = make_inputs(dims, B, st.n_init, seed=7)
        _LINE_RUNNERS[key] = StageRunner(emu_library(), "cpu", dims, iv, B, n_grid, dsf, inp["coeffs"], inp["steps"])
    return _LINE_RUNNERS[key]


@pytest.mark.parametrize("case", LINE_CASES)
def test_line_marching_gs_matches_sequential_sweeps(case):
    """The line-marching Gauss-Seidel kernel bodies (csrc/pdeop_gs_line.h) under the emulator's discrete-event model of
    the CUDA kernel -- random thread interleavings within what the split CTA barrier, the cp.async waits and the
    inter-CTA progress counters allow -- reproduce the sequential lexicographic sweeps bit for bit, for two random
    schedules per case (solver/multigrid.py:399-405)."""
    import ctypes
    dims, ivn, B, n_grid, dsf, sweeps, max_threads = case
    lib = emu_library()
    sr = _line_runner(dims, ivn, B, n_grid, dsf)
    n = sr.level_n(0)
    rng = np.random.default_rng(1)
    b, x0 = rng.standard_normal(B * n), rng.standard_normal(B * n)
    sr.plan.set_tuning("gs_pipe", 0)
    ref = sr.stage(_lib.STAGE_GS, 0, b, x0, count=sweeps)
    assert lib.dll.pdeop_emu_line_last() == 0
    sr.plan.set_tuning("gs_pipe", 5)
    lib.dll.pdeop_emu_set_line_max_threads(ctypes.c_int(max_threads))
    try:
        for seed in (1, 2):
            lib.dll.pdeop_emu_set_line_seed(ctypes.c_uint(seed))
            got = sr.stage(_lib.STAGE_GS, 0, b, x0, count=sweeps)
            assert lib.dll.pdeop_emu_line_last() == 1, "line kernel model not used, or it deadlocked"
            assert np.array_equal(got, ref)
    finally:
        lib.dll.pdeop_emu_set_line_max_threads(ctypes.c_int(0))
        sr.plan.set_tuning("gs_pipe", 0)
