"""``QPFunction(pde, mg, n_iv, ...)``: the reference's sparse multigrid QP function factory
(solver/qp_dual_sparse_multigrid_normal_kkt.py:21-164) over the B200-native solve.

Returned callable: ``fn(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps_list) -> x (B, n)``.
As in the reference, gradients flow to ``eq_constraints``, ``rhs``, ``iv_rhs`` and ``derivative_constraints``
(dA, drhs, div_rhs, dD) and NOT to ``coeffs`` / ``steps_list`` (:162), which only feed the coarse-grid
rediscretisation (:28).  ``eq_constraints`` / ``derivative_constraints`` are either the dense value carriers produced
by ``pde.build_equation_tensor`` / ``pde.build_derivative_tensor`` or, as in the reference, the ``torch.sparse``
tensors those methods return with ``sparse=True``: their values are pulled in construction order, and the gradients
come back as sparse tensors on the same pattern (dA exactly as the reference's; dD with each line value's gradient
-- the reference's per-nonzero dD summed over the grid directions the value is expanded along -- on the first
nonzero of that value, so that d(steps) through ``build_derivative_tensor`` is the same)."""
import torch

from ..config import PDEConfig
from .line_values import embed_order
from .lp_pde_central_diff import SparseValues
from ..ops import MGSolveFn, new_holder


def QPFunction(pde, mg, n_iv, gamma=1, alpha=1, double_ret=True, config=PDEConfig):
    def fn(eq_constraints, rhs, iv_rhs, derivative_constraints, coeffs, steps_list):
        B, G, M = pde.bs, pde.var_set.grid_size, pde.var_set.n_vars_per_step
        g = pde.equation_grid_pointers(rhs.device)
        if eq_constraints.is_sparse:            # the reference's own argument type (:25-33): values in construction order
            eq_constraints = SparseValues.apply(eq_constraints).reshape(B, -1, M)
        if torch.is_tensor(derivative_constraints):
            derivative_constraints = pde.line_values_from_sparse(derivative_constraints)
        # equation-row values come from eq_constraints (gradient path); the rows without an equation only feed
        # the coarse operators through linear interpolation (multigrid.py:243-256) and carry no gradient
        full = coeffs.detach().reshape(B, G, M).to(torch.float64).index_copy(1, g, eq_constraints.to(torch.float64))
        cv, fv, bv = embed_order(*derivative_constraints)
        plan = mg.plan
        if full.shape[-1] < plan.M:      # total order 1: the kernels' (u, u_c, u_cc) layout with zero u_cc columns
            full = torch.cat([full, full.new_zeros(B, G, plan.M - M)], dim=-1)
        coarse = mg.coarse_line_values([s.detach() for s in steps_list])
        holder = new_holder(mg.plan, coarse, fn.config)
        x = MGSolveFn.apply(full, rhs, iv_rhs, cv, fv, bv, holder)
        fn.last_holder = holder
        if M < plan.M:
            x = x.reshape(B, G, plan.M)[..., :M].reshape(B, G * M)
        return x

    fn.config = config
    fn.last_holder = None
    return fn
