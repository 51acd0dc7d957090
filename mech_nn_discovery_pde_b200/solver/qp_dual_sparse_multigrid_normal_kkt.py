"""``QPFunction(pde, mg, n_iv, ...)``: the reference's sparse multigrid QP function factory
(solver/qp_dual_sparse_multigrid_normal_kkt.py:21-164) over the B200-native solve.

Returned callable: ``fn(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps_list) -> x (B, n)``.
As in the reference, gradients flow to ``eq_constraints``, ``rhs``, ``iv_rhs`` and ``derivative_constraints``
(dA, drhs, div_rhs, dD) and NOT to ``coeffs`` / ``steps_list`` (:162), which only feed the coarse-grid
rediscretisation (:28).  ``eq_constraints`` / ``derivative_constraints`` are the dense value carriers produced by
``pde.build_equation_tensor`` / ``pde.build_derivative_tensor`` (the reference uses sparse tensors holding the
same numbers)."""
import torch

from ..config import PDEConfig
from ..ops import MGSolveFn, new_holder


def QPFunction(pde, mg, n_iv, gamma=1, alpha=1, double_ret=True, config=PDEConfig):
    def fn(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps_list):
        B, G, M = pde.bs, pde.var_set.grid_size, pde.var_set.n_vars_per_step
        g = pde.equation_grid_pointers(rhs.device)
        # equation-row values come from eq_constraints (gradient path); the rows without an equation only feed
        # the coarse operators through linear interpolation (multigrid.py:243-256) and carry no gradient
        full = coeffs.detach().reshape(B, G, M).to(torch.float64).index_copy(1, g, eq_constraints.to(torch.float64))
        cv, fv, bv = derivative_constraints
        coarse = mg.coarse_line_values([s.detach() for s in steps_list])
        holder = new_holder(mg.plan, coarse, fn.config)
        x = MGSolveFn.apply(full, rhs, iv_rhs, cv, fv, bv, holder)
        fn.last_holder = holder
        return x

    fn.config = config
    fn.last_holder = None
    return fn
