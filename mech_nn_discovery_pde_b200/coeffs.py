"""Callers either side of the PDE layer, fused.

``CoeffBuilder``: basis functions of the data fields x learned parameters -> ``coeffs (B,G,M)``, ``rhs (B,G)`` in one
kernel with a one-pass backward, replacing the dozen elementwise kernels + ``torch.zeros`` + strided channel writes
of ``Model.solve`` in the discovery scripts (discovery/ginzburg_landau.py:354-374,
discovery/burgers_dparam_viscous.py:261-279, discovery/kamani.py:252-271).

``data_loss``: mean |u0 - target|^p and its gradient in one kernel (discovery/ginzburg_landau.py:486-510).

Both are ``torch.library`` operators over the C ABI (include/pdeop.h: pdeop_coeff_forward/backward,
pdeop_loss_forward); no CPU path.
"""
import ctypes
from typing import List, Tuple

import torch
from torch import Tensor

from . import _lib

c_void_p = ctypes.c_void_p
KIND_INT, KIND_ABS_POW, KIND_NONE = 0, 1, 2


def _bind(lib):
    d = lib.dll
    if getattr(d, "_coeff_bound", False):
        return d
    ci, cd, ll = ctypes.c_int, ctypes.c_double, ctypes.c_longlong
    PI, PD, PP = ctypes.POINTER(ci), ctypes.POINTER(cd), ctypes.POINTER(c_void_p)
    d.pdeop_coeff_forward.argtypes = [ll, ci, ci, ci, ci, PI, PI, PI, c_void_p, PD, PP, c_void_p, c_void_p, c_void_p,
                                      c_void_p]
    d.pdeop_coeff_backward.argtypes = [ll, ci, ci, ci, ci, PI, PI, PI, c_void_p, PD, PP, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, PP, c_void_p]
    d.pdeop_loss_forward.argtypes = [ll, c_void_p, c_void_p, ci, c_void_p, c_void_p, c_void_p]
    d.pdeop_coeff_last_error.restype = ctypes.c_char_p
    d._coeff_bound = True
    return d


def _check(d, rc):
    if rc != 0:
        raise _lib.PdeopError("pdeop: " + d.pdeop_coeff_last_error().decode())


def _arrays(spec: List[int], c0: List[float], F: int):
    """spec = [M, NT, NP, pair_out..., pair_term..., kind (NT*F)...]"""
    M, NT, NP = spec[0], spec[1], spec[2]
    po = (ctypes.c_int * NP)(*spec[3:3 + NP])
    pt = (ctypes.c_int * NP)(*spec[3 + NP:3 + 2 * NP])
    kd = (ctypes.c_int * (NT * F))(*spec[3 + 2 * NP:3 + 2 * NP + NT * F])
    c0a = (ctypes.c_double * (M + 1))(*c0)
    return M, NT, NP, po, pt, kd, c0a


@torch.library.custom_op("pdeop::coeff_build", mutates_args=())
def coeff_build(fields: List[Tensor], w: Tensor, expo: Tensor, spec: List[int], c0: List[float]) -> Tuple[Tensor, Tensor]:
    lib = _lib.get_library()
    d = _bind(lib)
    F = len(fields)
    M, NT, NP, po, pt, kd, c0a = _arrays(spec, c0, F)
    fields = [f.to(torch.float64).contiguous() for f in fields]
    shape = fields[0].shape
    npts = fields[0].numel()
    dev = fields[0].device
    w = w.to(torch.float64).contiguous()
    expo = expo.to(torch.float64).contiguous()
    coeffs = torch.empty(*shape, M, dtype=torch.float64, device=dev)
    rhs = torch.empty(*shape, dtype=torch.float64, device=dev)
    fp = (c_void_p * F)(*[f.data_ptr() for f in fields])
    with torch.cuda.device(dev):
        _check(d, d.pdeop_coeff_forward(npts, M, F, NT, NP, po, pt, kd, _lib._ptr(expo), c0a, fp, _lib._ptr(w),
                                        _lib._ptr(coeffs), _lib._ptr(rhs), _lib.current_stream_ptr(dev)))
    return coeffs, rhs


@coeff_build.register_fake
def _(fields, w, expo, spec, c0):
    f = fields[0]
    return f.new_empty(*f.shape, spec[0], dtype=torch.float64), f.new_empty(f.shape, dtype=torch.float64)


@torch.library.custom_op("pdeop::coeff_build_backward", mutates_args=())
def coeff_build_backward(d_coeffs: Tensor, d_rhs: Tensor, fields: List[Tensor], w: Tensor, expo: Tensor,
                         spec: List[int], c0: List[float]) -> Tuple[Tensor, Tensor, List[Tensor]]:
    lib = _lib.get_library()
    d = _bind(lib)
    F = len(fields)
    M, NT, NP, po, pt, kd, c0a = _arrays(spec, c0, F)
    fields = [f.to(torch.float64).contiguous() for f in fields]
    npts = fields[0].numel()
    dev = fields[0].device
    w = w.to(torch.float64).contiguous()
    expo = expo.to(torch.float64).contiguous()
    d_coeffs = d_coeffs.to(torch.float64).contiguous()
    d_rhs = d_rhs.to(torch.float64).contiguous()
    d_w = torch.empty(NP, dtype=torch.float64, device=dev)
    d_expo = torch.empty(NT, F, dtype=torch.float64, device=dev)
    d_fields = [torch.empty_like(f) for f in fields]
    fp = (c_void_p * F)(*[f.data_ptr() for f in fields])
    dfp = (c_void_p * F)(*[f.data_ptr() for f in d_fields])
    with torch.cuda.device(dev):
        _check(d, d.pdeop_coeff_backward(npts, M, F, NT, NP, po, pt, kd, _lib._ptr(expo), c0a, fp, _lib._ptr(w),
                                         _lib._ptr(d_coeffs), _lib._ptr(d_rhs), _lib._ptr(d_w), _lib._ptr(d_expo), dfp,
                                         _lib.current_stream_ptr(dev)))
    return d_w, d_expo, d_fields


@coeff_build_backward.register_fake
def _(d_coeffs, d_rhs, fields, w, expo, spec, c0):
    return w.new_empty(w.shape, dtype=torch.float64), expo.new_empty(expo.shape, dtype=torch.float64), \
        [torch.empty_like(f, dtype=torch.float64) for f in fields]


def _cb_setup(ctx, inputs, output):
    fields, w, expo, spec, c0 = inputs
    ctx.spec, ctx.c0, ctx.nf = list(spec), list(c0), len(fields)
    ctx.save_for_backward(w, expo, *fields)


def _cb_backward(ctx, d_coeffs, d_rhs):
    w, expo, *fields = ctx.saved_tensors
    d_w, d_expo, d_fields = torch.ops.pdeop.coeff_build_backward(d_coeffs, d_rhs, list(fields), w, expo, ctx.spec,
                                                                 ctx.c0)
    return list(d_fields), d_w, d_expo, None, None


torch.library.register_autograd("pdeop::coeff_build", _cb_backward, setup_context=_cb_setup)


class CoeffBuilder(torch.nn.Module):
    """coeffs[..., o] = c0[o] + sum_k w[k] * term_{t_k}(fields),  rhs = the same for o = M.

    ``terms``: list of per-field factor specs, one tuple per term, each entry ``None`` (factor absent), an ``int``
    power of the signed field value, or ``("abs", i)`` = |field|^expo with a real exponent taken from entry ``i`` of the
    ``exponents`` tensor passed to ``forward`` (learnable).  ``pairs``: list of ``(output, term)``; ``w[k]`` of
    ``forward`` weighs pair k.  ``const``: {output: constant}.

    Ginzburg-Landau (discovery/ginzburg_landau.py:354-374), fields (up0, vp0):
        terms = [(0,0),(1,0),(2,0),(0,1),(0,2),(1,1),(0,3)]           # 1,u,u^2,v,v^2,uv,v^3
        pairs = [(0,t) for t in range(6)] + [(5,0),(5,1),(5,2)] + [(6,0),(6,1),(6,2)] + [(7,3),(7,4),(7,6)]
        const = {1: 1.0};  w = cat(params[0][:6], params[1][:3], params[2][:3], params[3][:3])
    """

    def __init__(self, n_orders, n_fields, terms, pairs, const=None):
        super().__init__()
        self.M, self.F = int(n_orders), int(n_fields)
        kinds, self._expo_src, base = [], [], []
        for t in terms:
            assert len(t) == self.F
            for e in t:
                if e is None:
                    kinds.append(KIND_NONE); base.append(0.0); self._expo_src.append(-1)
                elif isinstance(e, tuple):
                    kinds.append(KIND_ABS_POW); base.append(0.0); self._expo_src.append(int(e[1]))
                else:
                    kinds.append(KIND_INT); base.append(float(int(e))); self._expo_src.append(-1)
        self.NT, self.NP = len(terms), len(pairs)
        self.spec = [self.M, self.NT, self.NP] + [int(p[0]) for p in pairs] + [int(p[1]) for p in pairs] + kinds
        self.c0 = [float((const or {}).get(o, 0.0)) for o in range(self.M + 1)]
        self.register_buffer("_expo_base", torch.tensor(base, dtype=torch.float64).reshape(self.NT, self.F),
                             persistent=False)
        src = torch.tensor(self._expo_src, dtype=torch.long).reshape(self.NT, self.F)
        self.register_buffer("_expo_idx", src.clamp(min=0), persistent=False)
        self.register_buffer("_expo_mask", (src >= 0), persistent=False)

    def forward(self, fields, w, exponents=None):
        """fields: list of F tensors of one shape (B,G); w (NP,); exponents: 1-D tensor indexed by the ("abs", i)
        entries.  Returns coeffs (B,G,M), rhs (B,G) in fp64."""
        dev = fields[0].device
        if self._expo_base.device != dev:      # once: the small index buffers follow the data (no per-call H2D copy)
            self._expo_base = self._expo_base.to(dev)
            self._expo_idx = self._expo_idx.to(dev)
            self._expo_mask = self._expo_mask.to(dev)
        expo = self._expo_base
        if exponents is not None:
            expo = torch.where(self._expo_mask, exponents.to(torch.float64)[self._expo_idx], expo)
        return torch.ops.pdeop.coeff_build(list(fields), w, expo, self.spec, self.c0)


@torch.library.custom_op("pdeop::data_loss", mutates_args=())
def data_loss_op(u0: Tensor, target: Tensor, p: int) -> Tuple[Tensor, Tensor]:
    """(loss[1], grad) with loss = mean |u0 - target|^p, grad = d loss / d u0."""
    lib = _lib.get_library()
    d = _bind(lib)
    u0c = u0.to(torch.float64).contiguous()
    tc = target.to(torch.float64).contiguous()
    grad = torch.empty_like(u0c)
    loss = torch.empty(1, dtype=torch.float64, device=u0c.device)
    with torch.cuda.device(u0c.device):
        _check(d, d.pdeop_loss_forward(u0c.numel(), _lib._ptr(u0c), _lib._ptr(tc), int(p), _lib._ptr(grad),
                                       _lib._ptr(loss), _lib.current_stream_ptr(u0c.device)))
    return loss, grad


@data_loss_op.register_fake
def _(u0, target, p):
    return u0.new_empty(1, dtype=torch.float64), torch.empty_like(u0, dtype=torch.float64)


def _dl_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _dl_backward(ctx, g_loss, g_grad):
    (grad,) = ctx.saved_tensors
    return grad * g_loss, None, None


torch.library.register_autograd("pdeop::data_loss", _dl_backward, setup_context=_dl_setup)


def data_loss(u0, target, p=1):
    """mean |u0 - target|^p as a 0-d tensor; the gradient w.r.t. u0 is produced by the same kernel."""
    loss, _ = torch.ops.pdeop.data_loss(u0, target, int(p))
    return loss[0]
