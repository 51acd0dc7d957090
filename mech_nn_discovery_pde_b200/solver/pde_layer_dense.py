"""``PDEDenseLayer``: the reference's dense-KKT PDE layer (solver/pde_layer_dense.py:38-125) on the
B200-native dense path.  Same constructor, same ``forward(coeffs, rhs, iv_rhs, steps_list) -> (u0, u, eps)``.
"""
import torch
import torch.nn as nn

from ..ops import PdePlan
from . import qp_dual_dense_normal_kkt as MGS
from .lp_pde_central_diff import PDESYSLP


class PDEDenseLayer(nn.Module):
    """Dense PDE layer (pde_layer_dense.py:38-125)."""

    def __init__(self, bs, order, n_ind_dim, n_iv, init_index_mi_list, coord_dims, n_iv_steps, solver_dbl=True,
                 evolution=False, gamma=0.5, alpha=0.1, double_ret=False, device=None, _library=None):
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
        self.plan = PdePlan(self.coord_dims, order, bs * n_ind_dim, 1, True, init_index_mi_list, library=_library,
                            device=device)
        # the reference builds the dense layer's constraints with evolution=False whatever is passed
        # (pde_layer_dense.py:72-75)
        self.pde = PDESYSLP(bs * n_ind_dim, self.coord_dims, order, n_iv, init_index_mi_list, self.plan.n_init,
                            evolution=False, dtype=torch.float64)
        self.pde.plan = self.plan
        self.n_orders = len(self.pde.var_set.mi_list)
        self.grid_size = self.pde.var_set.grid_size
        self.step_grid_shape = self.pde.step_grid_shape
        self.qpf = MGS.QPFunction(self.pde, double_ret=double_ret)

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
        coeffs = coeffs.reshape(B, self.grid_size, self.n_orders)
        rhs = rhs.reshape(B, self.grid_size)
        if iv_rhs is not None:
            iv_rhs = iv_rhs.reshape(B, -1)
        else:
            iv_rhs = rhs.new_zeros(B, 0)
        for i in range(self.n_coord):   # pde_layer_dense.py:95-97
            steps_list[i] = steps_list[i].reshape(B, self.coord_dims[i] - 1)
        # solver_dbl=False (pde_layer_dense.py:64-69,101-105: the reference then factors A^T A in the inputs' precision):
        # here fp32 is a STORAGE format only -- operands are widened, the normal equations (cond ~1e10)
        # are formed, factored and solved in fp64, and the results are rounded back to the inputs' dtype
        out_dtype = torch.float64 if self.solver_dbl else coeffs.dtype
        coeffs = coeffs.double()
        rhs = rhs.double()
        iv_rhs = iv_rhs.double()
        steps = [s.double() for s in steps_list]

        # same call sequence as the reference (pde_layer_dense.py:107-110)
        sparse = getattr(self, "sparse_constraints", False)   # True: torch.sparse tensors, as the reference passes
        derivative_constraints = self.pde.build_derivative_tensor(steps, sparse=sparse)
        eq_constraints = self.pde.build_equation_tensor(coeffs, sparse=sparse)
        x = self.qpf(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps)
        eps = None
        u = self.pde.get_solution_reshaped(x)
        u = u.reshape(self.bs, self.n_ind_dim, *u.shape[1:])
        if u.dtype != out_dtype:
            u = u.to(out_dtype)
        u0 = u[:, :, :, 0]
        return u0, u, eps

    def solver_info(self):
        """(info_forward, info_backward) of the last call as (0, 0.0) pairs -- the dense path has no Krylov loop; kept
        so that callers can treat both layers alike.  Raises torch.linalg.LinAlgError if the factorisation failed
        (cholesky_ex(check_errors=True), qp_dual_dense_normal_kkt.py:39); syncs."""
        h = self.last_holder
        if h is None or h.info_fwd is None:
            return None, None
        h.check_factorization()
        return (0, 0.0), (None if h.info_bwd is None else (0, 0.0))
