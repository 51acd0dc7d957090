"""Autograd operators over the C ABI (include/pdeop.h).

``PdePlan`` owns one native plan (index tables of every multigrid level) and the scratch buffer.
``MGSolveFn`` / ``DenseSolveFn`` are the torch.autograd.Function pair that replaces the reference's
``QPFunctionFn`` (solver/qp_dual_sparse_multigrid_normal_kkt.py:21-164,
solver/qp_dual_dense_normal_kkt.py:19-120).  Inputs and outputs are dense tensors in the reference's
layouts; the native side never sees a sparse tensor.
"""
import ctypes

import torch

from . import _lib
from .config import PDEConfig

c_void_p = ctypes.c_void_p


def level_dims(coord_dims, n_grid, downsample_first):
    """Level extents (multigrid.py:88-102): halve every axis, or every axis but the first."""
    dims = [int(v) for v in coord_dims]
    out = []
    for _ in range(n_grid):
        out.append(tuple(dims))
        dims = [v // 2 if (c > 0 or downsample_first) else v for c, v in enumerate(dims)]
    return out


def iv_descriptors(init_index_mi_list, dims_list):
    """Evaluate the (coord, mi_index, begin, end) lambdas at every level's dims (multigrid.py:296-306)."""
    desc = []
    for dims in dims_list:
        lvl = []
        for f in init_index_mi_list:
            pair = f(*dims)
            lvl.append([int(pair[1])] + [int(v) for v in pair[2]] + [int(v) for v in pair[3]])
        desc.append(lvl)
    return desc


class PdePlan:
    def __init__(self, coord_dims, order, batch, n_grid, downsample_first, init_index_mi_list, library=None):
        self.lib = library if library is not None else _lib.get_library()
        self.coord_dims = tuple(int(v) for v in coord_dims)
        self.d = len(self.coord_dims)
        self.batch = int(batch)
        self.n_grid = int(n_grid)
        self.downsample_first = bool(downsample_first)
        self.dims_list = level_dims(self.coord_dims, self.n_grid, self.downsample_first)
        self.iv_desc = iv_descriptors(init_index_mi_list, self.dims_list)
        self.handle = self.lib.plan_create(self.coord_dims, order, self.batch, self.n_grid, self.downsample_first,
                                           self.iv_desc)
        q = self.lib.query
        self.M = q(self.handle, _lib.Q_M, 0)
        self.G = q(self.handle, _lib.Q_G, 0)
        self.n = self.G * self.M
        self.n_eq = q(self.handle, _lib.Q_N_EQ, 0)
        self.n_init = q(self.handle, _lib.Q_N_INIT, 0)
        self.Ntot = [q(self.handle, _lib.Q_NTOT, l) for l in range(self.n_grid)]
        self.Ftot = [q(self.handle, _lib.Q_FTOT, l) for l in range(self.n_grid)]
        self.persist_bytes = q(self.handle, _lib.Q_PERSIST_BYTES, 0)
        self._scratch = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def scratch(self, device, restart):
        nbytes = self.lib.query(self.handle, _lib.Q_SCRATCH_BYTES, int(restart))
        key = (str(device), nbytes)
        buf = self._scratch.get(key)
        if buf is None:
            self._scratch.clear()
            buf = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
            self._scratch[key] = buf
        return buf

    def new_persist(self, device):
        return torch.empty(self.persist_bytes // 8, dtype=torch.float64, device=device)

    def cfg(self, back=False, config=PDEConfig):
        c = _lib.SolverCfg()
        c.gs_pre = int(config.mg_gauss_seidel_steps_pre)
        c.gs_post = int(config.mg_gauss_seidel_steps_post)
        c.mg_steps = int(config.mg_steps_backward if back else config.mg_steps_forward)
        c.max_iter = int(config.mg_fgmres_max_iter_backward if back else config.mg_fgmres_max_iter_forward)
        c.restart = int(config.mg_fgmres_restarts_backward if back else config.mg_fgmres_restarts_forward)
        c.atol = float(getattr(config, "mg_fgmres_atol", 1e-5))
        c.gs_variant = int(getattr(config, "gs_variant", 0))
        return c


def _ptr_array(tensors):
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _check_inputs(plan, coeffs, rhs, iv_rhs, cv, fv, bv):
    B = plan.batch
    assert coeffs.shape == (B, plan.G, plan.M), f"coeffs {tuple(coeffs.shape)}"
    assert rhs.shape == (B, plan.G), f"rhs {tuple(rhs.shape)}"
    assert iv_rhs.shape == (B, plan.n_init), f"iv_rhs {tuple(iv_rhs.shape)} expected {(B, plan.n_init)}"
    assert cv.shape == (B, plan.Ntot[0], 2, 6) and fv.shape == (B, plan.Ftot[0], 4) and bv.shape == fv.shape
    for t in (coeffs, rhs, iv_rhs, cv, fv, bv):
        assert t.dtype == torch.float64, "pdeop kernels compute in fp64"


def _raise_if_not_spd(info, config):
    if getattr(config, "check_factorization", True):
        bad = int(info[3].item())
        if bad != 0:
            raise torch.linalg.LinAlgError(
                f"pdeop: Cholesky factorisation failed, leading minor of order {bad} is not positive-definite")


class _Holder:
    """Per-call state kept for the backward pass (what the reference stashes on ctx,
    qp_dual_sparse_multigrid_normal_kkt.py:66-76): the operator tables, coarse coefficients and the
    coarsest Cholesky factor live in ``persist``."""
    __slots__ = ("plan", "persist", "coarse", "config", "info_fwd", "info_bwd")


def new_holder(plan, coarse, config):
    h = _Holder()
    h.plan = plan
    h.coarse = coarse
    h.config = config
    h.persist = None
    h.info_fwd = None
    h.info_bwd = None
    return h


def _alloc_grads(plan, cv, fv, bv, dev):
    B = plan.batch
    d_coeffs = torch.empty(B, plan.G, plan.M, dtype=torch.float64, device=dev)
    d_rhs = torch.empty(B, plan.G, dtype=torch.float64, device=dev)
    d_iv = torch.empty(B, plan.n_init, dtype=torch.float64, device=dev)
    return d_coeffs, d_rhs, d_iv, torch.empty_like(cv), torch.empty_like(fv), torch.empty_like(bv)


class MGSolveFn(torch.autograd.Function):
    """x = FGMRES(A^T A, A^T b) with a V-cycle preconditioner; backward by implicit differentiation."""

    @staticmethod
    def forward(ctx, coeffs, rhs, iv_rhs, cv, fv, bv, holder):
        plan = holder.plan
        lib = plan.lib
        coeffs, rhs, iv_rhs, cv, fv, bv = [t.contiguous() for t in (coeffs, rhs, iv_rhs, cv, fv, bv)]
        _check_inputs(plan, coeffs, rhs, iv_rhs, cv, fv, bv)
        dev = coeffs.device
        cfg = plan.cfg(False, holder.config)
        scratch = plan.scratch(dev, max(cfg.restart, int(holder.config.mg_fgmres_restarts_backward)))
        holder.persist = plan.new_persist(dev)
        cvs = [cv] + [t[0] for t in holder.coarse]
        fvs = [fv] + [t[1] for t in holder.coarse]
        bvs = [bv] + [t[2] for t in holder.coarse]
        x = torch.empty(plan.batch, plan.n, dtype=torch.float64, device=dev)
        info = torch.zeros(4, dtype=torch.float64, device=dev)
        lib.check(lib.dll.pdeop_mg_forward(plan.handle, ctypes.byref(cfg), _lib._ptr(coeffs), _lib._ptr(rhs),
                                           _lib._ptr(iv_rhs), _ptr_array(cvs), _ptr_array(fvs), _ptr_array(bvs),
                                           _lib._ptr(holder.persist), _lib._ptr(scratch), _lib._ptr(x),
                                           _lib._ptr(info), _lib.current_stream_ptr(dev)))
        holder.info_fwd = info
        _raise_if_not_spd(info, holder.config)
        ctx.holder = holder
        ctx.save_for_backward(rhs, cv, fv, bv, x)
        return x

    @staticmethod
    def backward(ctx, grad_x):
        holder = ctx.holder
        plan = holder.plan
        lib = plan.lib
        rhs, cv, fv, bv, x = ctx.saved_tensors
        dev = x.device
        grad_x = grad_x.to(torch.float64).contiguous()
        cfg = plan.cfg(True, holder.config)
        scratch = plan.scratch(dev, max(cfg.restart, int(holder.config.mg_fgmres_restarts_forward)))
        d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv = _alloc_grads(plan, cv, fv, bv, dev)
        info = torch.zeros(4, dtype=torch.float64, device=dev)
        lib.check(lib.dll.pdeop_mg_backward(plan.handle, ctypes.byref(cfg), _lib._ptr(rhs), _lib._ptr(cv),
                                            _lib._ptr(fv), _lib._ptr(bv), _lib._ptr(holder.persist),
                                            _lib._ptr(scratch), _lib._ptr(x), _lib._ptr(grad_x), _lib._ptr(d_coeffs),
                                            _lib._ptr(d_rhs), _lib._ptr(d_iv), _lib._ptr(d_cv), _lib._ptr(d_fv),
                                            _lib._ptr(d_bv), _lib._ptr(info), _lib.current_stream_ptr(dev)))
        holder.info_bwd = info
        if getattr(holder.config, "rhs_grad_fp32_quirk", False):
            d_rhs = d_rhs.float().double()   # lp_pde_central_diff.py:1634
        return d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, None


class DenseSolveFn(torch.autograd.Function):
    """x = (A^T A)^-1 A^T b by dense Cholesky; backward reuses the factor (qp_dual_dense_normal_kkt.py:23-118)."""

    @staticmethod
    def forward(ctx, coeffs, rhs, iv_rhs, cv, fv, bv, holder):
        plan = holder.plan
        lib = plan.lib
        coeffs, rhs, iv_rhs, cv, fv, bv = [t.contiguous() for t in (coeffs, rhs, iv_rhs, cv, fv, bv)]
        _check_inputs(plan, coeffs, rhs, iv_rhs, cv, fv, bv)
        dev = coeffs.device
        scratch = plan.scratch(dev, 1)
        holder.persist = plan.new_persist(dev)
        x = torch.empty(plan.batch, plan.n, dtype=torch.float64, device=dev)
        info = torch.zeros(4, dtype=torch.float64, device=dev)
        lib.check(lib.dll.pdeop_dense_forward(plan.handle, _lib._ptr(coeffs), _lib._ptr(rhs), _lib._ptr(iv_rhs),
                                              _lib._ptr(cv), _lib._ptr(fv), _lib._ptr(bv), _lib._ptr(holder.persist),
                                              _lib._ptr(scratch), _lib._ptr(x), _lib._ptr(info),
                                              _lib.current_stream_ptr(dev)))
        holder.info_fwd = info
        _raise_if_not_spd(info, holder.config)
        ctx.holder = holder
        ctx.save_for_backward(rhs, cv, fv, bv, x)
        return x

    @staticmethod
    def backward(ctx, grad_x):
        holder = ctx.holder
        plan = holder.plan
        lib = plan.lib
        rhs, cv, fv, bv, x = ctx.saved_tensors
        dev = x.device
        grad_x = grad_x.to(torch.float64).contiguous()
        scratch = plan.scratch(dev, 1)
        d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv = _alloc_grads(plan, cv, fv, bv, dev)
        info = torch.zeros(4, dtype=torch.float64, device=dev)
        lib.check(lib.dll.pdeop_dense_backward(plan.handle, _lib._ptr(rhs), _lib._ptr(cv), _lib._ptr(fv),
                                               _lib._ptr(bv), _lib._ptr(holder.persist), _lib._ptr(scratch),
                                               _lib._ptr(x), _lib._ptr(grad_x), _lib._ptr(d_coeffs), _lib._ptr(d_rhs),
                                               _lib._ptr(d_iv), _lib._ptr(d_cv), _lib._ptr(d_fv), _lib._ptr(d_bv),
                                               _lib._ptr(info), _lib.current_stream_ptr(dev)))
        holder.info_bwd = info
        if getattr(holder.config, "rhs_grad_fp32_quirk", False):
            d_rhs = d_rhs.float().double()
        return d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, None
