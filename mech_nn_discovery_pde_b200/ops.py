"""Torch custom operators over the C ABI (include/pdeop.h).

``PdePlan`` owns one native plan (index tables of every multigrid level) and its scratch buffer.  The solve is
registered with ``torch.library`` as

    torch.ops.pdeop.mg_solve / mg_solve_backward          multigrid-preconditioned FGMRES
    torch.ops.pdeop.dense_solve / dense_solve_backward    dense Cholesky

with fake-tensor implementations and ``register_autograd`` formulas; they replace the reference's
``QPFunctionFn`` pair (solver/qp_dual_sparse_multigrid_normal_kkt.py:21-164,
solver/qp_dual_dense_normal_kkt.py:19-120) and its CuPy dependency.  Inputs and outputs are dense tensors in the
reference's layouts; the per-call operator state the reference stashes on ``ctx`` (:66-76) is the ``persist``
output tensor, owned by the autograd graph alone.  ``MGSolveFn`` / ``DenseSolveFn`` keep the round-1 call
signature on top of the ops.
"""
import ctypes
import itertools
import weakref
from typing import List, Tuple

import torch
from torch import Tensor

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


_PLAN_IDS = itertools.count(1)
_PLANS = weakref.WeakValueDictionary()   # op argument `plan` (an int) -> PdePlan


def plan_from_id(plan_id):
    plan = _PLANS.get(int(plan_id))
    if plan is None:
        raise _lib.PdeopError(f"pdeop: plan {plan_id} does not exist (any more)")
    return plan


class PdePlan:
    def __init__(self, coord_dims, order, batch, n_grid, downsample_first, init_index_mi_list, library=None,
                 evolution=False, chain=None, gs_pipe=None, device=None):
        self.lib = library if library is not None else _lib.get_library()
        self.coord_dims = tuple(int(v) for v in coord_dims)
        self.d = len(self.coord_dims)
        self.batch = int(batch)
        self.order = int(order)
        self.n_grid = int(n_grid)
        self.downsample_first = bool(downsample_first)
        self.evolution = bool(evolution)
        self.dims_list = level_dims(self.coord_dims, self.n_grid, self.downsample_first)
        self.iv_desc = iv_descriptors(init_index_mi_list, self.dims_list)
        # the plan's tables live on ONE device: the requested one, else the current one
        self.device = None
        if self.lib.backend.startswith("cuda"):
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            if self.device.type != "cuda":
                raise _lib.PdeopError(f"pdeop: device {device} is not a CUDA device; this package has no CPU path")
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
        with self.device_guard():
            self.handle = self.lib.plan_create(self.coord_dims, order, self.batch, self.n_grid,
                                               self.downsample_first, self.iv_desc, evolution=evolution, chain=chain,
                                               gs_pipe=gs_pipe)
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
        self.id = next(_PLAN_IDS)
        _PLANS[self.id] = self

    def device_guard(self):
        """Context manager making the plan's device current (the native entry points check it)."""
        if self.device is None:
            import contextlib
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)

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
        return _cfg_struct(knobs_of(config, back), float(getattr(config, "mg_fgmres_atol", 1e-5)))

    # ---- per-plan instrumentation and kernel-variant switch (pdeop.h) ----
    def profile_enable(self, on=True):
        self.lib.profile_enable(self.handle, on)

    def profile_collect(self):
        return self.lib.profile_collect(self.handle)

    def set_tuning(self, key, value):
        self.lib.set_tuning(self.handle, key, value)

    def get_tuning(self, key):
        return self.lib.get_tuning(self.handle, key)


def knobs_of(config, back):
    """[gs_pre, gs_post, mg_steps, max_iter, restart, gs_variant | mode, pcg_max_iter, smoother, sweeps, power_iters]
    from a PDEConfig-like object (config.py:13-29); mode 1 = converged (per-instance PCG)."""
    mode = {"reference": 0, "converged": 1}[getattr(config, "solver_mode", "reference")]
    smoother = {"jacobi": 0, "chebyshev": 1}[getattr(config, "mg_smoother", "chebyshev")]
    return [int(config.mg_gauss_seidel_steps_pre), int(config.mg_gauss_seidel_steps_post),
            int(config.mg_steps_backward if back else config.mg_steps_forward),
            int(config.mg_fgmres_max_iter_backward if back else config.mg_fgmres_max_iter_forward),
            int(config.mg_fgmres_restarts_backward if back else config.mg_fgmres_restarts_forward),
            int(getattr(config, "gs_variant", 0)),
            mode, int(getattr(config, "mg_pcg_max_iter", 1000)), smoother, int(getattr(config, "mg_smoother_sweeps", 8)),
            int(getattr(config, "mg_power_iters", 60))]


def fparams_of(config):
    """[atol (reference mode, fgmres.py:22), rtol (converged mode), jacobi_w (config.py:29), Chebyshev interval ratio]"""
    return [float(getattr(config, "mg_fgmres_atol", 1e-5)), float(getattr(config, "mg_pcg_rtol", 1e-8)),
            float(getattr(config, "jacobi_w", 0.4)), float(getattr(config, "mg_cheb_ratio", 30.0))]


def _pcg_struct(knobs, fparams):
    c = _lib.PcgCfg()
    c.max_iter, c.smoother, c.sweeps, c.power_iters = int(knobs[7]), int(knobs[8]), int(knobs[9]), int(knobs[10])
    c.rtol, c.jacobi_w, c.cheb_ratio = float(fparams[1]), float(fparams[2]), float(fparams[3])
    return c


def _cfg_struct(knobs, atol):
    c = _lib.SolverCfg()
    c.gs_pre, c.gs_post, c.mg_steps, c.max_iter, c.restart, c.gs_variant = [int(v) for v in knobs[:6]]
    c.atol = float(atol)
    return c


def _ptr_array(tensors):
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _check_inputs(plan, coeffs, rhs, iv_rhs, cv, fv, bv):
    B = plan.batch
    want = {"coeffs": (coeffs, (B, plan.G, plan.M)), "rhs": (rhs, (B, plan.G)), "iv_rhs": (iv_rhs, (B, plan.n_init)),
            "cv": (cv, (B, plan.Ntot[0], 2, 6)), "fv": (fv, (B, plan.Ftot[0], 4)), "bv": (bv, (B, plan.Ftot[0], 4))}
    for name, (t, shape) in want.items():
        if tuple(t.shape) != shape:
            raise ValueError(f"pdeop: {name} has shape {tuple(t.shape)}, expected {shape}")
        if t.dtype != torch.float64:
            raise ValueError(f"pdeop: {name} is {t.dtype}; the kernels compute in fp64")
        if plan.device is not None and t.device != plan.device:
            raise ValueError(f"pdeop: {name} is on {t.device} but the plan lives on {plan.device}")


def _raise_if_not_spd(info):
    bad = int(info[3].item())
    if bad != 0:
        raise torch.linalg.LinAlgError(
            f"pdeop: Cholesky factorisation failed, leading minor of order {bad} is not positive-definite")


# ---------------------------------------------------------------------------------------------------------------
# torch.library operators.  `plan` is the integer id of a live PdePlan; `knobs` = knobs_of(config, back).
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("pdeop::mg_solve", mutates_args=())
def mg_solve(coeffs: Tensor, rhs: Tensor, iv_rhs: Tensor, cv: List[Tensor], fv: List[Tensor], bv: List[Tensor],
             plan: int, knobs: List[int], knobs_bwd: List[int], fparams: List[float],
             flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    """x (B,n), persist, info[4] = FGMRES(A^T A, A^T b) with a V-cycle preconditioner
    (qp_dual_sparse_multigrid_normal_kkt.py:25-79).  cv/fv/bv: line values of every level."""
    pl = plan_from_id(plan)
    lib = pl.lib
    coeffs, rhs, iv_rhs = coeffs.contiguous(), rhs.contiguous(), iv_rhs.contiguous()
    cv = [t.contiguous() for t in cv]
    fv = [t.contiguous() for t in fv]
    bv = [t.contiguous() for t in bv]
    _check_inputs(pl, coeffs, rhs, iv_rhs, cv[0], fv[0], bv[0])
    if not (len(cv) == len(fv) == len(bv) == pl.n_grid):
        raise ValueError("pdeop: one set of line values per multigrid level is required")
    dev = coeffs.device
    scratch = pl.scratch(dev, max(int(knobs[4]), int(knobs_bwd[4]), pl.n_grid + 1, 5))
    persist = pl.new_persist(dev)
    x = torch.empty(pl.batch, pl.n, dtype=torch.float64, device=dev)
    info = torch.zeros(4, dtype=torch.float64, device=dev)
    args = (_lib._ptr(coeffs), _lib._ptr(rhs), _lib._ptr(iv_rhs), _ptr_array(cv), _ptr_array(fv), _ptr_array(bv),
            _lib._ptr(persist), _lib._ptr(scratch), _lib._ptr(x), _lib._ptr(info), _lib.current_stream_ptr(dev))
    with pl.device_guard():
        if knobs[6] == 1:     # converged mode: per-instance PCG
            cfg = _pcg_struct(knobs, fparams)
            lib.check(lib.dll.pdeop_mg_forward_converged(pl.handle, ctypes.byref(cfg), *args))
        else:
            cfg = _cfg_struct(knobs, fparams[0])
            lib.check(lib.dll.pdeop_mg_forward(pl.handle, ctypes.byref(cfg), *args))
    return x, persist, info


@mg_solve.register_fake
def _(coeffs, rhs, iv_rhs, cv, fv, bv, plan, knobs, knobs_bwd, fparams, flags):
    pl = plan_from_id(plan)
    return (coeffs.new_empty(pl.batch, pl.n), coeffs.new_empty(pl.persist_bytes // 8), coeffs.new_empty(4))


@torch.library.custom_op("pdeop::mg_solve_backward", mutates_args=())
def mg_solve_backward(grad_x: Tensor, x: Tensor, rhs: Tensor, cv0: Tensor, fv0: Tensor, bv0: Tensor, persist: Tensor,
                      plan: int, knobs: List[int], fparams: List[float]
                      ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """(d_coeffs, d_rhs, d_iv_rhs, d_cv, d_fv, d_bv, info) by implicit differentiation: dz = (A^T A)^-1 g with the
    same operator and preconditioner left in `persist` (qp_dual_sparse_multigrid_normal_kkt.py:81-162)."""
    pl = plan_from_id(plan)
    lib = pl.lib
    dev = x.device
    grad_x = grad_x.to(torch.float64).contiguous()
    scratch = pl.scratch(dev, max(int(knobs[4]), pl.n_grid + 1, 5))
    B = pl.batch
    d_coeffs = torch.empty(B, pl.G, pl.M, dtype=torch.float64, device=dev)
    d_rhs = torch.empty(B, pl.G, dtype=torch.float64, device=dev)
    d_iv = torch.empty(B, pl.n_init, dtype=torch.float64, device=dev)
    d_cv, d_fv, d_bv = torch.empty_like(cv0), torch.empty_like(fv0), torch.empty_like(bv0)
    info = torch.zeros(4, dtype=torch.float64, device=dev)
    args = (_lib._ptr(rhs), _lib._ptr(cv0), _lib._ptr(fv0), _lib._ptr(bv0), _lib._ptr(persist), _lib._ptr(scratch),
            _lib._ptr(x), _lib._ptr(grad_x), _lib._ptr(d_coeffs), _lib._ptr(d_rhs), _lib._ptr(d_iv), _lib._ptr(d_cv),
            _lib._ptr(d_fv), _lib._ptr(d_bv), _lib._ptr(info), _lib.current_stream_ptr(dev))
    with pl.device_guard():
        if knobs[6] == 1:
            cfg = _pcg_struct(knobs, fparams)
            lib.check(lib.dll.pdeop_mg_backward_converged(pl.handle, ctypes.byref(cfg), *args))
        else:
            cfg = _cfg_struct(knobs, fparams[0])
            lib.check(lib.dll.pdeop_mg_backward(pl.handle, ctypes.byref(cfg), *args))
    return d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, info


@mg_solve_backward.register_fake
def _(grad_x, x, rhs, cv0, fv0, bv0, persist, plan, knobs, fparams):
    pl = plan_from_id(plan)
    B = pl.batch
    return (x.new_empty(B, pl.G, pl.M), x.new_empty(B, pl.G), x.new_empty(B, pl.n_init), torch.empty_like(cv0),
            torch.empty_like(fv0), torch.empty_like(bv0), x.new_empty(4))


@torch.library.custom_op("pdeop::dense_solve", mutates_args=())
def dense_solve(coeffs: Tensor, rhs: Tensor, iv_rhs: Tensor, cv: Tensor, fv: Tensor, bv: Tensor, plan: int,
                flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    """x (B,n), persist, info[4] = (A^T A)^-1 A^T b by dense Cholesky (qp_dual_dense_normal_kkt.py:23-56)."""
    pl = plan_from_id(plan)
    lib = pl.lib
    coeffs, rhs, iv_rhs, cv, fv, bv = [t.contiguous() for t in (coeffs, rhs, iv_rhs, cv, fv, bv)]
    _check_inputs(pl, coeffs, rhs, iv_rhs, cv, fv, bv)
    dev = coeffs.device
    scratch = pl.scratch(dev, 1)
    persist = pl.new_persist(dev)
    x = torch.empty(pl.batch, pl.n, dtype=torch.float64, device=dev)
    info = torch.zeros(4, dtype=torch.float64, device=dev)
    with pl.device_guard():
        lib.check(lib.dll.pdeop_dense_forward(pl.handle, _lib._ptr(coeffs), _lib._ptr(rhs), _lib._ptr(iv_rhs),
                                              _lib._ptr(cv), _lib._ptr(fv), _lib._ptr(bv), _lib._ptr(persist),
                                              _lib._ptr(scratch), _lib._ptr(x), _lib._ptr(info),
                                              _lib.current_stream_ptr(dev)))
    return x, persist, info


@dense_solve.register_fake
def _(coeffs, rhs, iv_rhs, cv, fv, bv, plan, flags):
    pl = plan_from_id(plan)
    return (coeffs.new_empty(pl.batch, pl.n), coeffs.new_empty(pl.persist_bytes // 8), coeffs.new_empty(4))


@torch.library.custom_op("pdeop::dense_solve_backward", mutates_args=())
def dense_solve_backward(grad_x: Tensor, x: Tensor, rhs: Tensor, cv: Tensor, fv: Tensor, bv: Tensor, persist: Tensor,
                         plan: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Gradients of dense_solve; reuses the factor in `persist` (qp_dual_dense_normal_kkt.py:58-118)."""
    pl = plan_from_id(plan)
    lib = pl.lib
    dev = x.device
    grad_x = grad_x.to(torch.float64).contiguous()
    scratch = pl.scratch(dev, 1)
    B = pl.batch
    d_coeffs = torch.empty(B, pl.G, pl.M, dtype=torch.float64, device=dev)
    d_rhs = torch.empty(B, pl.G, dtype=torch.float64, device=dev)
    d_iv = torch.empty(B, pl.n_init, dtype=torch.float64, device=dev)
    d_cv, d_fv, d_bv = torch.empty_like(cv), torch.empty_like(fv), torch.empty_like(bv)
    info = torch.zeros(4, dtype=torch.float64, device=dev)
    with pl.device_guard():
        lib.check(lib.dll.pdeop_dense_backward(pl.handle, _lib._ptr(rhs), _lib._ptr(cv), _lib._ptr(fv), _lib._ptr(bv),
                                               _lib._ptr(persist), _lib._ptr(scratch), _lib._ptr(x),
                                               _lib._ptr(grad_x), _lib._ptr(d_coeffs), _lib._ptr(d_rhs),
                                               _lib._ptr(d_iv), _lib._ptr(d_cv), _lib._ptr(d_fv), _lib._ptr(d_bv),
                                               _lib._ptr(info), _lib.current_stream_ptr(dev)))
    return d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, info


@dense_solve_backward.register_fake
def _(grad_x, x, rhs, cv, fv, bv, persist, plan):
    pl = plan_from_id(plan)
    B = pl.batch
    return (x.new_empty(B, pl.G, pl.M), x.new_empty(B, pl.G), x.new_empty(B, pl.n_init), torch.empty_like(cv),
            torch.empty_like(fv), torch.empty_like(bv), x.new_empty(4))


# flags: bit 0 = rhs_grad_fp32_quirk (lp_pde_central_diff.py:1634), bit 1 = check_factorization
FLAG_RHS_FP32, FLAG_CHECK_SPD = 1, 2

# info of the most recent backward of each plan (what the reference logs, :105-107); read by solver_info()
_LAST_BWD_INFO = {}


def _mg_setup_context(ctx, inputs, output):
    coeffs, rhs, iv_rhs, cv, fv, bv, plan, knobs, knobs_bwd, fparams, flags = inputs
    x, persist, info = output
    ctx.plan, ctx.knobs_bwd, ctx.fparams, ctx.flags, ctx.n_levels = plan, list(knobs_bwd), list(fparams), flags, len(cv)
    # `persist` (15-32 GB at the Ginzburg-Landau sizes) and `info` are outputs only so that the graph owns them: without
    # this autograd would materialise a zero gradient of the same size for each of them in every backward
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(rhs, cv[0], fv[0], bv[0], x, persist, info)


def _mg_backward(ctx, grad_x, grad_persist, grad_info):
    rhs, cv0, fv0, bv0, x, persist, info = ctx.saved_tensors
    if grad_x is None:
        grad_x = torch.zeros_like(x)
    if ctx.flags & FLAG_CHECK_SPD:
        _raise_if_not_spd(info)   # lazily: the forward never synchronises the host (cholesky_ex check, multigrid.py:439)
    d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, info_b = torch.ops.pdeop.mg_solve_backward(
        grad_x, x, rhs, cv0, fv0, bv0, persist, ctx.plan, ctx.knobs_bwd, ctx.fparams)
    _LAST_BWD_INFO[ctx.plan] = info_b
    if ctx.flags & FLAG_RHS_FP32:
        d_rhs = d_rhs.float().double()
    none = [None] * (ctx.n_levels - 1)
    return (d_coeffs, d_rhs, d_iv, [d_cv] + none, [d_fv] + none, [d_bv] + none, None, None, None, None, None)


torch.library.register_autograd("pdeop::mg_solve", _mg_backward, setup_context=_mg_setup_context)


def _dense_setup_context(ctx, inputs, output):
    coeffs, rhs, iv_rhs, cv, fv, bv, plan, flags = inputs
    x, persist, info = output
    ctx.plan, ctx.flags = plan, flags
    ctx.set_materialize_grads(False)    # see _mg_setup_context
    ctx.save_for_backward(rhs, cv, fv, bv, x, persist, info)


def _dense_backward(ctx, grad_x, grad_persist, grad_info):
    rhs, cv, fv, bv, x, persist, info = ctx.saved_tensors
    if grad_x is None:
        grad_x = torch.zeros_like(x)
    if ctx.flags & FLAG_CHECK_SPD:
        _raise_if_not_spd(info)
    d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, info_b = torch.ops.pdeop.dense_solve_backward(
        grad_x, x, rhs, cv, fv, bv, persist, ctx.plan)
    _LAST_BWD_INFO[ctx.plan] = info_b
    if ctx.flags & FLAG_RHS_FP32:
        d_rhs = d_rhs.float().double()
    return d_coeffs, d_rhs, d_iv, d_cv, d_fv, d_bv, None, None


torch.library.register_autograd("pdeop::dense_solve", _dense_backward, setup_context=_dense_setup_context)


def config_flags(config):
    return ((FLAG_RHS_FP32 if getattr(config, "rhs_grad_fp32_quirk", False) else 0)
            | (FLAG_CHECK_SPD if getattr(config, "check_factorization", True) else 0))


class _Holder:
    """What a layer keeps of its most recent call: the plan, the knobs, and the two 4-double info tensors
    {iters, r_norm, b_norm, chol_info}.  It does NOT own the persist buffer -- the autograd graph does, so the
    operator state of a call is released with its graph."""
    __slots__ = ("plan", "coarse", "config", "info_fwd", "__weakref__")

    @property
    def info_bwd(self):
        return _LAST_BWD_INFO.get(self.plan.id)

    def check_factorization(self):
        """Raises torch.linalg.LinAlgError if the forward's Cholesky failed (synchronises)."""
        if self.info_fwd is not None:
            _raise_if_not_spd(self.info_fwd)


def new_holder(plan, coarse, config):
    h = _Holder()
    h.plan = plan
    h.coarse = coarse
    h.config = config
    h.info_fwd = None
    _LAST_BWD_INFO.pop(plan.id, None)
    return h


def mg_solve_call(coeffs, rhs, iv_rhs, cv, fv, bv, holder):
    """x = torch.ops.pdeop.mg_solve(...) for one layer call; records info on the holder."""
    plan, config = holder.plan, holder.config
    cvs = [cv] + [t[0] for t in holder.coarse]
    fvs = [fv] + [t[1] for t in holder.coarse]
    bvs = [bv] + [t[2] for t in holder.coarse]
    x, _persist, info = torch.ops.pdeop.mg_solve(coeffs, rhs, iv_rhs, cvs, fvs, bvs, plan.id, knobs_of(config, False),
                                                 knobs_of(config, True), fparams_of(config), config_flags(config))
    holder.info_fwd = info.detach()
    return x


def dense_solve_call(coeffs, rhs, iv_rhs, cv, fv, bv, holder):
    x, _persist, info = torch.ops.pdeop.dense_solve(coeffs, rhs, iv_rhs, cv, fv, bv, holder.plan.id,
                                                    config_flags(holder.config))
    holder.info_fwd = info.detach()
    return x


class MGSolveFn:
    """Round-1 call signature: ``MGSolveFn.apply(coeffs, rhs, iv_rhs, cv, fv, bv, holder)``."""
    apply = staticmethod(mg_solve_call)


class DenseSolveFn:
    apply = staticmethod(dense_solve_call)
