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


class SparseValues(torch.autograd.Function):
    """values of a (possibly uncoalesced) sparse COO tensor in construction order; the gradient is a sparse tensor on
    the same pattern (what the reference's backward returns for dA / dD, qp_dual_sparse_multigrid_normal_kkt.py:132-162)."""

    @staticmethod
    def forward(ctx, sp):
        ctx.idx = sp._indices()
        ctx.shape = sp.shape
        return sp._values().clone()

    @staticmethod
    def backward(ctx, g):
        return torch.sparse_coo_tensor(ctx.idx, g, ctx.shape, check_invariants=False)


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

    def build_equation_tensor(self, coeffs, sparse=False):
        """Values of the equation rows, (B, n_eq, M), in the reference's value order (row-major over equation
        rows, then multi-index; :1707-1719).  Default: dense carrier.  sparse=True: the reference's own return type,
        a torch.sparse_coo_tensor (B, n_eq, n) whose values are these numbers in construction order."""
        vals = self.remove_pad(coeffs, coeffs=True)
        if not sparse:
            return vals
        idx = self.equation_indices(vals.device)
        return torch.sparse_coo_tensor(idx, vals.reshape(-1), (self.bs, self.num_added_equation_constraints,
                                                               self.var_set.num_vars), dtype=vals.dtype,
                                       check_invariants=False)

    def build_derivative_tensor(self, steps_list, sparse=False):
        """Values of the derivative rows per line position: (central (B,Ntot,2,6), forward (B,Ftot,4), backward
        (B,Ftot,4)) -- build_derivative_values (:1618-1630) before its expansion over the grid.  sparse=True: the
        reference's return type, torch.sparse_coo_tensor (B, n_deriv, n) holding the expanded per-nonzero values in
        construction order [central | forward | backward] (:1721-1731)."""
        cv, fv, bv = line_values(steps_list, self.order)
        if not sparse:
            return cv, fv, bv
        rows, cols, src = self.derivative_structure(cv.device)
        B = cv.shape[0]
        flat = torch.cat([cv.reshape(B, -1), fv.reshape(B, -1), bv.reshape(B, -1)], dim=1)
        vals = flat.index_select(1, src)                          # expansion over the other grid directions
        nnz = rows.numel()
        bidx = torch.arange(B, device=cv.device).repeat_interleave(nnz)
        idx = torch.stack([bidx, rows.repeat(B), cols.repeat(B)])
        return torch.sparse_coo_tensor(idx, vals.reshape(-1), (B, self.num_added_derivative_constraints,
                                                               self.var_set.num_vars), dtype=vals.dtype,
                                       check_invariants=False)

    # ---- sparse-tensor adapter (the reference's QPFunction takes torch.sparse tensors,
    # ---- qp_dual_sparse_multigrid_normal_kkt.py:25-33): closed-form COO structure in construction order -----------
    def equation_indices(self, device=None):
        """(3, B*n_eq*M) [batch, row, col] of the equation block: row r = r-th interior point in C order, columns
        g*M + m (lp_pde_central_diff.py:746-764, 1171-1190)."""
        if getattr(self, "_eq_idx", None) is None:
            g = self.equation_grid_pointers()
            M = self.var_set.n_vars_per_step
            n_eq = g.numel()
            rows = torch.arange(n_eq).repeat_interleave(M)
            cols = (g.cpu()[:, None] * M + torch.arange(M)[None, :]).reshape(-1)
            b = torch.arange(self.bs).repeat_interleave(n_eq * M)
            self._eq_idx = torch.stack([b, rows.repeat(self.bs), cols.repeat(self.bs)])
        if device is not None and self._eq_idx.device != device:
            self._eq_idx = self._eq_idx.to(device)
        return self._eq_idx

    def derivative_structure(self, device=None):
        """(rows, cols, src) of the derivative block's nonzeros in construction order: central rows coordinate-major,
        grid C order, derivative order ascending, 5 stencil columns + the own derivative channel (:886-1006); forward
        rows [u, u_c, u_cc, u(next)] over the points off the far edge, backward rows [u, u_c, u_cc, u(prev)] over
        the points off the near edge (:785-884).  src maps each nonzero to its entry of the flattened line values
        [central (Ntot,order,6) | forward (Ftot,order+2) | backward (Ftot,order+2)]."""
        if getattr(self, "_d_struct", None) is None:
            dims, d, order = self.coord_dims, self.n_coord, self.order
            M = self.var_set.n_vars_per_step
            G = self.var_set.grid_size
            idx = np.indices(dims).reshape(d, G)
            gptr = np.arange(G)
            strides = np.array([int(np.prod(dims[c + 1:])) for c in range(d)], dtype=np.int64)
            chan = lambda c, k: 1 + c if k == 1 else 1 + d + c
            ks = list(range(1, order + 1))
            tc = order + 2
            rows, cols, src = [], [], []
            r0 = 0
            cvoff = np.concatenate([[0], np.cumsum(dims)])[:-1]
            fvoff = np.concatenate([[0], np.cumsum([n - 1 for n in dims])])[:-1]
            n_cv = int(sum(dims)) * order * 6
            n_fv = int(sum(n - 1 for n in dims)) * tc
            for c in range(d):
                i = idx[c]
                nc = dims[c]
                off = np.tile(np.arange(5) - 2, (G, 1))
                off[i <= 1] = np.arange(5)
                off[(i > 1) & (i >= nc - 2)] = -np.arange(5)
                ucols = (gptr[:, None] + off * strides[c]) * M
                for kk, k in enumerate(ks):
                    own = gptr * M + chan(c, k)
                    cc = np.concatenate([ucols, own[:, None]], axis=1)                     # (G, 6)
                    rr = r0 + gptr * order + kk
                    ss = ((cvoff[c] + i)[:, None] * order + kk) * 6 + np.arange(6)[None, :]
                    rows.append((np.repeat(rr, 6), kk))
                    cols.append(cc.reshape(-1))
                    src.append(ss.reshape(-1))
                # interleave the per-order blocks so that entries follow (grid point, order) as constructed
                r_k = [rows[-order + j][0].reshape(G, 6) for j in range(order)]
                c_k = [cols[-order + j].reshape(G, 6) for j in range(order)]
                s_k = [src[-order + j].reshape(G, 6) for j in range(order)]
                del rows[-order:], cols[-order:], src[-order:]
                rows.append((np.stack(r_k, axis=1).reshape(-1), 0))
                cols.append(np.stack(c_k, axis=1).reshape(-1))
                src.append(np.stack(s_k, axis=1).reshape(-1))
                r0 += G * order
            rows = [r[0] for r in rows]
            for c in range(d):      # forward
                sel = gptr[idx[c] != dims[c] - 1]
                ent = [sel * M] + [sel * M + chan(c, k) for k in ks] + [(sel + strides[c]) * M]
                rows.append(np.repeat(r0 + np.arange(sel.shape[0]), tc))
                cols.append(np.stack(ent, axis=1).reshape(-1))
                src.append((n_cv + (fvoff[c] + idx[c][sel])[:, None] * tc + np.arange(tc)[None, :]).reshape(-1))
                r0 += sel.shape[0]
            for c in range(d):      # backward: line value i belongs to line position i + 1
                sel = gptr[idx[c] != 0]
                ent = [sel * M] + [sel * M + chan(c, k) for k in ks] + [(sel - strides[c]) * M]
                rows.append(np.repeat(r0 + np.arange(sel.shape[0]), tc))
                cols.append(np.stack(ent, axis=1).reshape(-1))
                src.append((n_cv + n_fv + (fvoff[c] + idx[c][sel] - 1)[:, None] * tc
                            + np.arange(tc)[None, :]).reshape(-1))
                r0 += sel.shape[0]
            assert r0 == self.num_added_derivative_constraints
            self._d_struct = tuple(torch.as_tensor(np.concatenate(a), dtype=torch.long) for a in (rows, cols, src))
        if device is not None and self._d_struct[0].device != device:
            self._d_struct = tuple(t.to(device) for t in self._d_struct)
        return self._d_struct

    def line_values_from_sparse(self, derivative_constraints):
        """Inverse of build_derivative_tensor(sparse=True): per-line values (cv, fv, bv) from the per-nonzero values
        of a derivative-constraint sparse tensor in construction order (one representative nonzero per line value);
        differentiable, so the gradient comes back as a sparse tensor on the same pattern."""
        rows, cols, src = self.derivative_structure(derivative_constraints.device)
        B = self.bs
        vals = SparseValues.apply(derivative_constraints).reshape(B, -1)
        n_line = int(src.max().item()) + 1
        rep = torch.full((n_line,), -1, dtype=torch.long, device=vals.device)
        pos = torch.arange(src.numel() - 1, -1, -1, device=vals.device)
        rep[src.flip(0)] = pos                       # first occurrence wins
        flat = vals.index_select(1, rep)
        d, order = self.n_coord, self.order
        ncv = sum(self.coord_dims) * order * 6
        nfv = sum(n - 1 for n in self.coord_dims) * (order + 2)
        cv = flat[:, :ncv].reshape(B, -1, order, 6)
        fv = flat[:, ncv:ncv + nfv].reshape(B, -1, order + 2)
        bv = flat[:, ncv + nfv:].reshape(B, -1, order + 2)
        return cv.contiguous(), fv.contiguous(), bv.contiguous()

    def get_solution_reshaped(self, x):
        """(B, n) -> (B, G, M)  (lp_pde_central_diff.py:486-494)."""
        x = x[:, :self.var_set.num_vars]
        return x.reshape(-1, *self.var_set.multi_index_shape)
