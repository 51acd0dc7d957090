"""``QPFunction(pde)``: the reference's dense QP function factory (solver/qp_dual_dense_normal_kkt.py:19-120) over
the B200-native dense path.  Same callable signature and gradient routing as the sparse variant."""
import torch

from ..config import PDEConfig
from ..ops import DenseSolveFn, new_holder


def QPFunction(pde, double_ret=True, config=PDEConfig):
    def fn(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps_list):
        B, G, M = pde.bs, pde.var_set.grid_size, pde.var_set.n_vars_per_step
        g = pde.equation_grid_pointers(rhs.device)
        full = coeffs.detach().reshape(B, G, M).to(torch.float64).index_copy(1, g, eq_constraints.to(torch.float64))
        cv, fv, bv = derivative_constraints
        holder = new_holder(pde.plan, [], fn.config)
        x = DenseSolveFn.apply(full, rhs, iv_rhs, cv, fv, bv, holder)
        fn.last_holder = holder
        return x

    fn.config = config
    fn.last_holder = None
    return fn
