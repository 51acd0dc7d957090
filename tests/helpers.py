"""Shared helpers for the parity tests: run the product's layers / stages through the C ABI on a given
library (the CUDA library on the GPU box; the host emulator in CPU-only structure tests)."""
import ctypes
import os

import numpy as np
import torch

from mech_nn_discovery_pde_b200 import _lib
from mech_nn_discovery_pde_b200.config import PDEConfig
from mech_nn_discovery_pde_b200.ops import PdePlan, _ptr_array
from mech_nn_discovery_pde_b200.solver.line_values import coarsen_steps, line_values
from mech_nn_discovery_pde_b200.solver.multigrid import MultigridLayer
from mech_nn_discovery_pde_b200.solver.pde_layer_dense import PDEDenseLayer
from oracle.cases import IV_LISTS, make_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_layer_case(name):
    z = np.load(os.path.join(GOLDEN, f"layer_{name}.npz"))
    dims = tuple(int(v) for v in z["dims"])
    steps = [z[f"steps{c}"] for c in range(len(dims))]
    return z, dims, steps


def golden_stride(z):
    """Round-2 fixtures (oracle/make_golden_r2.py, make_golden_port.py) store large outputs as every stride-th entry
    of the flattened array plus the norm of the full array."""
    return int(z["stride"]) if "stride" in z.files else 1


def rel_sampled(got, z, key):
    """Relative error of `got` against golden field `key`, honouring the fixture's sampling stride; also checks the
    full-array norm when the fixture stores it."""
    g = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(z[key], dtype=np.float64).reshape(-1)
    e = rel(g[::golden_stride(z)], ref)
    if key + "_norm" in z.files:
        nrm = float(z[key + "_norm"])
        e = max(e, abs(float(np.linalg.norm(g)) - nrm) / max(nrm, 1e-300))
    return e


def golden_loss_w(z, dims):
    """Loss weights of a layer fixture: stored, or (sampled fixtures) regenerated from the seed and checked."""
    if "loss_w" in z.files:
        return z["loss_w"]
    inp = make_inputs(dims, int(z["bs"]), z["iv_rhs"].shape[1], int(z["seed"]), uniform=bool(z["uniform"]))
    assert np.array_equal(inp["coeffs"], z["coeffs"]), "seeded inputs do not reproduce the fixture's"
    assert abs(np.linalg.norm(inp["loss_w"]) - float(z["loss_w_norm"])) < 1e-12 * float(z["loss_w_norm"])
    return inp["loss_w"]


def seeded_stage_vectors(s):
    """(v, x0) of a stagesS_* fixture, regenerated as oracle/make_golden_r2.py drew them."""
    g = torch.Generator().manual_seed(int(s["seed"]))
    n = int(s["n"])
    v = torch.randn(n, generator=g, dtype=torch.float64).numpy()
    x0 = torch.randn(n, generator=g, dtype=torch.float64).numpy()
    assert np.array_equal(v[:8], s["v_head"]) and np.array_equal(x0[:8], s["x0_head"])
    assert abs(np.linalg.norm(v) - float(s["v_norm"])) < 1e-12 * float(s["v_norm"])
    return v, x0


class StageRunner:
    """Operator set-up + single multigrid building blocks through pdeop_mg_setup / pdeop_stage."""

    def __init__(self, lib, device, dims, iv_list, B, n_grid, dsf, coeffs, steps, config=PDEConfig, chain=None):
        self.lib = lib
        self.dev = torch.device(device)
        self.plan = PdePlan(dims, 2, B, n_grid, dsf, iv_list, library=lib, chain=chain,
                            device=self.dev if self.dev.type == "cuda" else None)
        self.B = B
        self.cfg = self.plan.cfg(False, config)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(self.dev)
        steps_t = [t(s).reshape(B, -1) for s in steps]
        vals = [line_values(steps_t)]
        cur = steps_t
        for l in range(1, n_grid):
            cur = coarsen_steps(cur, self.plan.dims_list[l - 1], dsf)
            vals.append(line_values(cur))
        self.vals = vals
        self.persist = self.plan.new_persist(self.dev)
        self.scratch = self.plan.scratch(self.dev, self.cfg.restart)
        self.info = torch.zeros(4, dtype=torch.float64, device=self.dev)
        coeffs_t = t(coeffs).reshape(B, self.plan.G, self.plan.M).contiguous()
        self.keep = (coeffs_t, vals)
        lib.check(lib.dll.pdeop_mg_setup(self.plan.handle, _lib._ptr(coeffs_t), _ptr_array([v[0] for v in vals]),
                                         _ptr_array([v[1] for v in vals]), _ptr_array([v[2] for v in vals]),
                                         _lib._ptr(self.persist), _lib._ptr(self.scratch), _lib._ptr(self.info),
                                         _lib.current_stream_ptr(self.dev)))

    def level_n(self, level):
        return self.lib.query(self.plan.handle, _lib.Q_G, level) * self.plan.M

    def stage(self, stage, level, in1, in2=None, count=0, out_level=None, gs_variant=None):
        t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(
            self.dev).contiguous()
        a, b = t(in1), t(in2)
        out_level = level if out_level is None else out_level
        out = torch.zeros(self.B * self.level_n(out_level), dtype=torch.float64, device=self.dev)
        cfg = self.cfg
        if gs_variant is not None:
            cfg = self.plan.cfg(False)
            cfg.gs_variant = gs_variant
        self.lib.check(self.lib.dll.pdeop_stage(self.plan.handle, ctypes.byref(cfg), stage, level, count,
                                                _lib._ptr(a), _lib._ptr(b), _lib._ptr(out), _lib._ptr(self.persist),
                                                _lib._ptr(self.scratch), _lib.current_stream_ptr(self.dev)))
        return out.cpu().numpy()


def run_layer_case(lib, device, name, config=PDEConfig, sparse=False):
    """Forward + backward of the product layer on a golden case; returns dict of numpy results.
    sparse=True: the layer hands torch.sparse constraint tensors to its QPFunction, as the reference does."""
    z, dims, steps = load_layer_case(name)
    iv = IV_LISTS[str(z["iv_name"])]
    B = int(z["bs"])
    dev = torch.device(device)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    if str(z["kind"]) == "dense":
        layer = PDEDenseLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, init_index_mi_list=iv,
                              n_iv_steps=1, double_ret=True, solver_dbl=True, _library=lib)
    else:
        layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=int(z["n_grid"]),
                               evolution=False, downsample_first=bool(z["dsf"]), init_index_mi_list=iv, n_iv_steps=1,
                               double_ret=True, solver_dbl=True, _library=lib)
    layer.config = config
    layer.sparse_constraints = sparse
    coeffs = t(z["coeffs"]).requires_grad_(True)
    rhs = t(z["rhs"]).requires_grad_(True)
    ivr = t(z["iv_rhs"]).requires_grad_(True)
    st = [t(s).requires_grad_(True) for s in steps]
    u0, u, eps = layer(coeffs, rhs, ivr, list(st))
    assert eps is None
    loss = (u * t(golden_loss_w(z, dims)).reshape(u.shape)).sum()
    loss.backward()
    out = dict(u=u.detach().cpu().numpy(), u0=u0.detach().cpu().numpy(), d_coeffs=coeffs.grad.cpu().numpy(),
               d_rhs=rhs.grad.cpu().numpy(), d_iv_rhs=ivr.grad.cpu().numpy(),
               d_steps=[s.grad.cpu().numpy() for s in st])
    h = layer.last_holder
    out["info_fwd"] = h.info_fwd.cpu().numpy()
    out["info_bwd"] = h.info_bwd.cpu().numpy()
    return z, out
