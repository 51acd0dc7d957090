"""Host-side mirror of the reference's constraint bookkeeping (solver/lp_pde_central_diff.py).

The reference's ``PDESYSLP`` enumerates constraints with per-grid-point Python loops and stores COO
index lists; here every count is closed form and the index tables the kernels need are built inside
the native plan (csrc/pdeop_solver.cpp).  This class keeps the attributes callers read
(``var_set.*``, ``num_added_*_constraints``, ``step_grid_shape``, ``get_solution_reshaped``).
"""
from itertools import combinations

import numpy as np
import torch

from .line_values import line_values


class QPVariableSet:
    """Variable numbering: var = ravel(grid_index) * M + mi (lp_pde_central_diff.py:96-107);
    mi_list = [u | first derivatives by coord | second derivatives by coord] (:304-315)."""

    def __init__(self, coord_dims, order):
        self.coord_dims = tuple(int(v) for v in coord_dims)
        self.n_coord = len(self.coord_dims)
        self.grid_size = int(np.prod(self.coord_dims))
        self.order = order
        n = self.n_coord
        zero = (tuple(0 for _ in range(n)),)
        first = tuple(tuple(1 if i in comb else 0 for i in range(n)) for comb in combinations(range(n), 1))
        second = tuple(tuple(2 if i in comb else 0 for i in range(n)) for comb in combinations(range(n), 1))
        if order == 2:
            self.mi_list = zero + first + second
            self.mi_list_repr = ["u"] + [f"u_x{i}" for i in range(n)] + [f"u_x{i}x{i}" for i in range(n)]
            self.t_deriv_mi_index = [1, 1 + n]
        elif order == 1:
            self.mi_list = zero + first
            self.mi_list_repr = ["u"] + [f"u_x{i}" for i in range(n)]
            self.t_deriv_mi_index = [1]
        else:
            raise ValueError("unsupported total order")
        self.mi_indices = list(range(len(self.mi_list)))
        self.mi_to_index = {mi: i for i, mi in enumerate(self.mi_list)}
        self.taylor_mi_indices = [0]
        self.n_vars_per_step = len(self.mi_list)
        self.num_vars = self.n_vars_per_step * self.grid_size
        self.num_pde_vars = self.num_vars
        self.multi_index_shape = (self.grid_size, self.n_vars_per_step)
        self.num_added_eps_vars = 0
        self.sorted_central_mi_indices = {
            c: sorted([mi for mi in self.mi_list if mi[c] != 0], key=lambda mi: mi[c]) for c in range(n)}
        self.order_count = {}
        for c in range(n):
            cnt = {}
            for mi in self.mi_list:
                cnt[mi[c]] = cnt.get(mi[c], 0) + 1
            self.order_count[c] = cnt

    def get_variable_from_mi_index(self, index):
        grid_pointer = int(np.ravel_multi_index(tuple(index[0]), self.coord_dims, order="C"))
        return grid_pointer * self.n_vars_per_step + int(index[1])


class PDESYSLP:
    """Closed-form counts of the constraint system A = [equation; initial; derivative] rows
    (lp_pde_central_diff.py:1063-1139).  Row counts:
      n_eq     = (n_0-1) * prod_{c>=1}(n_c-2)                  (:228-235, :751)
      central  = 2 d G, forward = backward = sum_c (n_c-1) G/n_c (:869-884, :993-1006)
    """

    def __init__(self, bs, coord_dims, order, n_iv, init_index_mi_list, n_init, evolution=False, dtype=None):
        if evolution:
            raise NotImplementedError("evolution=True equation rows are not implemented (no shipped script uses them)")
        self.bs = bs
        self.coord_dims = tuple(int(v) for v in coord_dims)
        self.n_coord = len(self.coord_dims)
        self.order = order
        self.n_iv = n_iv
        self.init_index_mi_list = init_index_mi_list
        self.evolution = evolution
        self.dtype = dtype
        self.var_set = QPVariableSet(self.coord_dims, order)
        G = self.var_set.grid_size
        d = self.n_coord
        n_eq = self.coord_dims[0] - 1
        for c in range(1, d):
            n_eq *= self.coord_dims[c] - 2
        self.num_added_equation_constraints = int(n_eq)
        self.num_added_initial_constraints = int(n_init)
        n_central = order * d * G
        n_fb = sum((n - 1) * G // n for n in self.coord_dims)
        self.num_added_derivative_constraints = int(n_central + 2 * n_fb)
        self.num_added_constraints = (self.num_added_equation_constraints + self.num_added_initial_constraints
                                      + self.num_added_derivative_constraints)
        self.num_constraints = self.num_added_constraints
        self.tc_count = order + 2
        self._eq_g = None
        self.plan = None   # native plan, attached by the layer that owns this object
        step_coords = np.array(self.coord_dims)
        self.step_grid_size = {}
        self.step_grid_shape = {}
        for i in range(d):
            one_hot = np.array([1 if k == i else 0 for k in range(d)])
            self.step_grid_size[i] = int(np.prod(step_coords - one_hot))
            self.step_grid_shape[i] = tuple(int(v) for v in (step_coords - one_hot))

    # ---- per-call constraint values (the reference returns torch.sparse tensors here; the kernels consume
    # ---- the same numbers un-expanded, so these return dense carriers) -------------------------------------
    def equation_grid_pointers(self, device=None):
        """C-order grid pointers of the points that carry an equation row: first-axis index >= 1 and strictly
        inside every other axis (lp_pde_central_diff.py:228-235, 748-764)."""
        if self._eq_g is None:
            idx = np.indices(self.coord_dims).reshape(self.n_coord, -1)
            keep = idx[0] != 0
            for c in range(1, self.n_coord):
                keep &= (idx[c] != 0) & (idx[c] != self.coord_dims[c] - 1)
            self._eq_g = torch.as_tensor(np.nonzero(keep)[0], dtype=torch.long)
            assert self._eq_g.numel() == self.num_added_equation_constraints
        if device is not None and self._eq_g.device != device:
            self._eq_g = self._eq_g.to(device)
        return self._eq_g

    def remove_pad(self, values, coeffs=True):
        """(B, G, M) -> (B, n_eq, M) or (B, G) -> (B, n_eq): rows of the interior points (:1686-1705)."""
        g = self.equation_grid_pointers(values.device)
        if coeffs:
            return values.reshape(self.bs, self.var_set.grid_size, -1).index_select(1, g)
        return values.reshape(self.bs, self.var_set.grid_size).index_select(1, g)

    def add_pad(self, eq_values):
        """(B, n_eq) -> (B, *coord_dims) with zeros on the rows that carry no equation (:1632-1647); fp64."""
        g = self.equation_grid_pointers(eq_values.device)
        out = eq_values.new_zeros(eq_values.shape[0], self.var_set.grid_size)
        out.index_copy_(1, g, eq_values)
        return out.reshape(eq_values.shape[0], *self.coord_dims)

    def build_equation_tensor(self, coeffs):
        """Values of the equation rows, (B, n_eq, M), in the reference's value order (row-major over equation
        rows, then multi-index; :1707-1719).  Dense carrier instead of torch.sparse_coo_tensor."""
        return self.remove_pad(coeffs, coeffs=True)

    def build_derivative_tensor(self, steps_list):
        """Values of the derivative rows per line position: (central (B,Ntot,2,6), forward (B,Ftot,4), backward
        (B,Ftot,4)) -- build_derivative_values (:1618-1630) before its expansion over the grid."""
        return line_values(steps_list)

    def get_solution_reshaped(self, x):
        """(B, n) -> (B, G, M)  (lp_pde_central_diff.py:486-494)."""
        x = x[:, :self.var_set.num_vars]
        return x.reshape(-1, *self.var_set.multi_index_shape)
