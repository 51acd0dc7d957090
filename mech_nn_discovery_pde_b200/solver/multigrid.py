"""``MultigridLayer``: the reference's sparse multigrid PDE layer (solver/multigrid.py:536-623) on the
B200-native solve.  Same constructor, same ``forward(coeffs, rhs, iv_rhs, steps_list) -> (u0, u, eps)``.
"""
import numpy as np
import torch
import torch.nn as nn

from ..config import PDEConfig
from ..ops import PdePlan
from . import qp_dual_sparse_multigrid_normal_kkt as MGS
from .line_values import coarsen_steps, embed_order, line_values
from .lp_pde_central_diff import PDESYSLP


class MultigridSolver:
    """Level hierarchy of the V-cycle (multigrid.py:45-112).  Holds the native plan; the coarse operators are
    rediscretised on the device each forward call (multigrid.py:115-163)."""

    def __init__(self, bs, order, n_ind_dim, n_iv, init_index_mi_list, coord_dims, n_iv_steps, solver_dbl=True,
                 evolution=False, downsample_first=True, gamma=0.5, alpha=0.1, double_ret=False, n_grid=2,
                 device=None, _library=None):
        self.coord_dims = tuple(int(v) for v in coord_dims)
        self.n_coord = len(self.coord_dims)
        self.order = order
        self.n_ind_dim = n_ind_dim
        self.n_iv = n_iv
        self.bs = bs
        self.device = device
        self.solver_dbl = solver_dbl
        self.init_index_mi_list = init_index_mi_list
        self.evolution = evolution
        self.downsample_first = downsample_first
        self.n_grid = n_grid
        if n_grid < 2:
            raise ValueError("MultigridLayer needs n_grid >= 2")
        self.plan = PdePlan(self.coord_dims, order, bs * n_ind_dim, n_grid, downsample_first, init_index_mi_list,
                            library=_library, evolution=evolution, device=device)
        self.dim_list = list(self.plan.dims_list)
        self.size_list = [int(np.prod(d)) for d in self.dim_list]
        for dims in self.dim_list:
            assert min(dims) >= 8, "every multigrid level needs extents >= 8 (multigrid.py:95)"
        self.pde_list = []
        for l, dims in enumerate(self.dim_list):
            n_init = self.plan.lib.query(self.plan.handle, 4, l)
            self.pde_list.append(PDESYSLP(bs * n_ind_dim, dims, order, n_iv, init_index_mi_list, n_init,
                                          evolution=evolution, dtype=torch.float64))

    @torch.no_grad()
    def coarse_line_values(self, steps_list):
        """Line values of levels 1.. from pairwise-summed spacings (multigrid.py:139, 271-285)."""
        out = []
        cur = [s.detach() for s in steps_list]
        for l in range(1, self.n_grid):
            cur = coarsen_steps(cur, self.dim_list[l - 1], self.downsample_first)
            out.append(embed_order(*line_values(cur, self.order)))
        return out


class MultigridLayer(nn.Module):
    """Multigrid layer (multigrid.py:536-623)."""

    def __init__(self, bs, order, n_ind_dim, n_iv, init_index_mi_list, coord_dims, n_iv_steps, solver_dbl=True,
                 evolution=False, downsample_first=True, gamma=0.5, alpha=0.1, double_ret=False, n_grid=2,
                 device=None, _library=None):
        super().__init__()
        self.step_size = 0.01
        self.coord_dims = tuple(int(v) for v in coord_dims)
        self.n_coord = len(self.coord_dims)
        self.order = order
        self.n_ind_dim = n_ind_dim
        self.n_dim = 1
        self.n_equations = 1
        self.n_iv = n_iv
        self.n_iv_steps = 1
        self.bs = bs
        self.device = device
        self.solver_dbl = solver_dbl
        self.evolution = evolution
        self.double_ret = double_ret
        # the reference hard-codes the fp64 multigrid solver whatever solver_dbl says (multigrid.py:569-570)
        self.mg_solver = MultigridSolver(bs, order, n_ind_dim, n_iv, init_index_mi_list, coord_dims, n_iv_steps,
                                         solver_dbl=True, n_grid=n_grid, evolution=evolution,
                                         downsample_first=downsample_first, device=device, _library=_library)
        self.pde = self.mg_solver.pde_list[0]
        self.pde.plan = self.mg_solver.plan
        self.n_orders = len(self.pde.var_set.mi_list)
        self.grid_size = self.pde.var_set.grid_size
        self.step_grid_shape = self.pde.step_grid_shape
        self.qpf = MGS.QPFunction(self.pde, self.mg_solver, self.n_iv, gamma=gamma, alpha=alpha, double_ret=double_ret)

    @property
    def config(self):
        return self.qpf.config

    @config.setter
    def config(self, value):
        self.qpf.config = value

    @property
    def last_holder(self):
        return self.qpf.last_holder

    def forward(self, coeffs, rhs, iv_rhs, steps_list):
        B = self.bs * self.n_ind_dim
        self.mg_solver.device = rhs.device
        coeffs = coeffs.reshape(B, self.grid_size, self.n_orders)
        rhs = rhs.reshape(B, self.grid_size)
        if iv_rhs is not None:
            iv_rhs = iv_rhs.reshape(B, -1)
        else:
            iv_rhs = rhs.new_zeros(B, 0)
        for i in range(self.n_coord):   # the reference re-assigns the caller's list entries (multigrid.py:596-598)
            steps_list[i] = steps_list[i].reshape(B, self.coord_dims[i] - 1)
        # kernels compute in fp64 (multigrid.py:601-605)
        coeffs = coeffs.double()
        rhs = rhs.double()
        iv_rhs = iv_rhs.double()
        steps = [s.double() for s in steps_list]

        # same call sequence as the reference (multigrid.py:607-611)
        sparse = getattr(self, "sparse_constraints", False)   # True: torch.sparse tensors, as the reference passes
        derivative_constraints = self.pde.build_derivative_tensor(steps, sparse=sparse)
        eq_constraints = self.pde.build_equation_tensor(coeffs, sparse=sparse)
        x = self.qpf(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps)
        eps = None
        u = self.pde.get_solution_reshaped(x)
        u = u.reshape(self.bs, self.n_ind_dim, *u.shape[1:])
        u0 = u[:, :, :, 0]
        return u0, u, eps

    def solver_info(self):
        """(iters, r_norm) of the last forward and backward FGMRES solves (fgmres.py:182); syncs."""
        h = self.last_holder
        f = None if h is None or h.info_fwd is None else (int(h.info_fwd[0].item()), float(h.info_fwd[1].item()))
        b = None if h is None or h.info_bwd is None else (int(h.info_bwd[0].item()), float(h.info_bwd[1].item()))
        return f, b
