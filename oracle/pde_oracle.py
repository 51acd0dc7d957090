"""CPU oracle for the differentiable PDE-layer solve  --  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch CPU restatement (numpy / scipy / torch-CPU) of the algorithm the
reference (alpz/mech-nn-discovery-pde) runs on its PDE-layer hot path.  It is the checker the
parity tests, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg compare the CUDA
product against.  Nothing in the product package may import it.

Parity status: PINNED against the reference itself.  The reference ships no tests or golden
vectors for this path, so the pin is: ``oracle/make_golden.py`` imports the
unmodified reference from /root/reference in the build container (with the three stub modules in
``oracle/stubs``), runs it on seeded inputs and commits inputs+outputs under ``tests/golden``;
``tests/test_oracle_vs_golden.py`` checks this restatement against those vectors, and against the
one known-answer the reference holds (uniform-step finite-difference constants,
solver/lp_pde_central_diff.py:929-937,981-984).

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg as sla
import torch


# --------------------------------------------------------------------------------------------
# Knobs  (config.py:13-29)
# --------------------------------------------------------------------------------------------
class OracleConfig:
    mg_gauss_seidel_steps_pre = 5
    mg_gauss_seidel_steps_post = 5
    mg_steps_forward = 1
    mg_steps_backward = 1
    mg_fgmres_max_iter_forward = 40
    mg_fgmres_restarts_forward = 10
    mg_fgmres_max_iter_backward = 40
    mg_fgmres_restarts_backward = 10


# --------------------------------------------------------------------------------------------
# Structure  (solver/lp_pde_central_diff.py:33-348 QPVariableSet, :746-1139 PDESYSLP build side)
# --------------------------------------------------------------------------------------------
@dataclass
class Structure:
    dims: Tuple[int, ...]
    order: int
    d: int
    M: int
    G: int
    n: int
    # equation rows: grid pointers (C order) of points carrying an equation row
    eq_g: np.ndarray
    # initial rows: variable index of each initial/boundary row, in construction order
    init_var: np.ndarray
    # derivative rows, COO in construction order (row index local to the derivative block)
    d_row: np.ndarray
    d_col: np.ndarray
    n_central: int
    n_fwd: int
    n_bwd: int
    tc: int  # entries per forward/backward row

    @property
    def n_eq(self):
        return int(self.eq_g.shape[0])

    @property
    def n_init(self):
        return int(self.init_var.shape[0])

    @property
    def n_deriv(self):
        return self.n_central + self.n_fwd + self.n_bwd

    @property
    def rows(self):
        return self.n_eq + self.n_init + self.n_deriv


def channel_of(d: int, order: int, coord: int, k: int) -> int:
    """mi_list = [u | first derivs by coord | second derivs by coord]  (lp_pde_central_diff.py:304-315)."""
    if k == 1:
        return 1 + coord
    assert order == 2 and k == 2
    return 1 + d + coord


def build_structure(dims: Sequence[int], iv_list: Sequence[Callable], order: int = 2) -> Structure:
    dims = tuple(int(x) for x in dims)
    d = len(dims)
    assert order in (1, 2)
    M = 1 + d * order
    G = int(np.prod(dims))
    n = G * M
    idx = np.indices(dims).reshape(d, G)  # idx[c, g], C order  (:45)
    gptr = np.arange(G)
    strides = np.array([int(np.prod(dims[c + 1:])) for c in range(d)], dtype=np.int64)

    # equation rows: skip t==0 and spatial edges  (:228-235, :748-764)
    keep = idx[0] != 0
    for c in range(1, d):
        keep &= (idx[c] != 0) & (idx[c] != dims[c] - 1)
    eq_g = gptr[keep]

    # initial rows  (:1008-1033): for each iv lambda, grid points inside [begin, end] in C order
    init_var = []
    for f in iv_list:
        pair = f(*dims)
        mi_index = int(pair[1])
        rb = np.array(pair[2]).reshape(d, 1)
        re = np.array(pair[3]).reshape(d, 1)
        inside = np.all((idx >= rb) & (idx <= re), axis=0)
        init_var.append(gptr[inside] * M + mi_index)
    init_var = np.concatenate(init_var) if init_var else np.zeros(0, dtype=np.int64)

    # derivative rows  (:1057-1060): central, forward, backward
    rows = []
    cols = []
    r0 = 0
    ks = (1, 2) if order == 2 else (1,)
    # central (:993-1006, :886-991): coord-major, grid C order, orders ascending, 5 stencil cols + own channel
    for c in range(d):
        i = idx[c]
        nc = dims[c]
        left = i <= 1
        right = (~left) & (i >= nc - 2)
        for kk, k in enumerate(ks):
            off = np.empty((G, 5), dtype=np.int64)
            j = np.arange(5)
            off[:] = (j - 2)[None, :]
            off[left] = j[None, :]
            off[right] = -j[None, :]
            ucols = (gptr[:, None] + off * strides[c]) * M  # u channel = 0
            own = gptr * M + channel_of(d, order, c, k)
            ccols = np.concatenate([ucols, own[:, None]], axis=1)  # (G, 6)
            rr = r0 + gptr * len(ks) + kk
            rows.append(np.repeat(rr, 6))
            cols.append(ccols.reshape(-1))
        r0 += G * len(ks)
    n_central = r0
    # interleave so that construction order is (grid, k) -- rows carry it, sort by row keeps entry order
    # forward (:869-875, :785-863): [u, u_c, (u_cc), u(next)]
    tc = order + 2
    nf0 = r0
    for c in range(d):
        sel = gptr[idx[c] != dims[c] - 1]
        ent = [sel * M] + [sel * M + channel_of(d, order, c, k) for k in ks] + [(sel + strides[c]) * M]
        ccols = np.stack(ent, axis=1)
        rr = r0 + np.arange(sel.shape[0])
        rows.append(np.repeat(rr, tc))
        cols.append(ccols.reshape(-1))
        r0 += sel.shape[0]
    n_fwd = r0 - nf0
    nb0 = r0
    for c in range(d):
        sel = gptr[idx[c] != 0]
        ent = [sel * M] + [sel * M + channel_of(d, order, c, k) for k in ks] + [(sel - strides[c]) * M]
        ccols = np.stack(ent, axis=1)
        rr = r0 + np.arange(sel.shape[0])
        rows.append(np.repeat(rr, tc))
        cols.append(ccols.reshape(-1))
        r0 += sel.shape[0]
    n_bwd = r0 - nb0
    d_row = np.concatenate(rows)
    d_col = np.concatenate(cols)
    # central entries were emitted k-major per coord; the reference emits (grid, k) interleaved.
    # Re-order the central block so entry order equals construction order (stable sort on row).
    nce = n_central * 6
    o = np.argsort(d_row[:nce], kind="stable")
    d_row[:nce] = d_row[:nce][o]
    d_col[:nce] = d_col[:nce][o]
    return Structure(dims, order, d, M, G, n, eq_g, init_var, d_row, d_col, n_central, n_fwd, n_bwd, tc)


# --------------------------------------------------------------------------------------------
# Per-call derivative-row values  (lp_pde_central_diff.py:1300-1630)
# --------------------------------------------------------------------------------------------
def _vander_weights(nodes: torch.Tensor) -> torch.Tensor:
    """nodes (..., 5) -> weights (..., 5, 2): columns = first / second derivative at 0  (:1334-1341, :1415-1422)."""
    ones = torch.ones_like(nodes)
    p2 = nodes.pow(2)
    mat = torch.stack([ones, nodes, p2, nodes * p2, p2 * p2], dim=-2)
    rhs = torch.tensor([[0, 1, 0, 0, 0], [0, 0, 2, 0, 0]], dtype=nodes.dtype).T
    return torch.linalg.solve(mat, rhs.expand(*mat.shape[:-2], 5, 2))


def central_line_values(steps: torch.Tensor, order: int = 2) -> torch.Tensor:
    """steps (B, n-1) -> (B, n, order, 6): per line position, per derivative order, the 6 row values.

    Follows solve_5pt_stencil_edge (:1300-1396) for positions {0,1} and {n-2,n-1} (including the
    reference's choice of spacings steps[1:5]/steps[2:6] resp. steps[-3:-1]... for the edge rows) and
    solve_5pt_central_stencil (:1398-1492) for the interior.
    """
    B, nm1 = steps.shape
    # left edge, positions 0,1  (:1320-1332)
    begin = torch.zeros_like(steps[:, 0:2])
    s1, s2, s3, s4 = steps[:, 1:3], steps[:, 2:4], steps[:, 3:5], steps[:, 4:6]
    r1 = s1
    r2 = r1 + s2
    r3 = r2 + s3
    r4 = r3 + s4
    wl = _vander_weights(torch.stack([begin, r1, r2, r3, r4], dim=-1))
    hl = s1
    # right edge, positions n-2, n-1  (:1304-1318)
    end = torch.zeros_like(steps[:, -2:])
    t1, t2, t3, t4 = steps[:, -3:-1], steps[:, -4:-2], steps[:, -5:-3], steps[:, -6:-4]
    l1 = -t1
    l2 = l1 - t2
    l3 = l2 - t3
    l4 = l3 - t4
    wr = _vander_weights(torch.stack([end, l1, l2, l3, l4], dim=-1))
    hr = t1
    # interior, positions 2..n-3  (:1402-1414)
    n1 = steps[:, 2:-1]
    n2 = steps[:, 3:]
    p1 = steps[:, 1:-2]
    p2 = steps[:, :-3]
    center = torch.zeros_like(n1)
    wc = _vander_weights(torch.stack([-p1 - p2, -p1, center, n1, n1 + n2], dim=-1))
    hc = n1
    w = torch.cat([wl, wc, wr], dim=1)  # (B, n, 5, 2)
    h = torch.cat([hl, hc, hr], dim=1)  # (B, n)
    outs = []
    for k in range(1, order + 1):
        hk = h.unsqueeze(-1) ** k
        outs.append(torch.cat([w[..., k - 1] * hk, -hk], dim=-1))  # (:1349-1350, :1429-1430)
    return torch.stack(outs, dim=2)


def forward_line_values(steps: torch.Tensor, order: int = 2) -> torch.Tensor:
    """(B, n-1) -> (B, n-1, order+2): [1, h, h^2/2, -1]  (:785-848, :1550-1581)."""
    ent = [steps ** j / math.factorial(j) for j in range(order + 1)] + [-(steps ** 0)]
    return torch.stack(ent, dim=-1)


def backward_line_values(steps: torch.Tensor, order: int = 2) -> torch.Tensor:
    """(B, n-1) -> (B, n-1, order+2): [1, -h, h^2/2, -1]; entry i belongs to line position i+1  (:849-861, :1583-1615)."""
    ms = -steps
    ent = [ms ** j / math.factorial(j) for j in range(order + 1)] + [(ms ** 0) / (-1.0)]
    return torch.stack(ent, dim=-1)


def _expand_line(st: Structure, c: int, t: torch.Tensor, n_line: int) -> torch.Tensor:
    """(B, n_line, *tail) -> (B, *dims with dims[c]=n_line, *tail) flattened to (B, -1)  (:1271-1283, :1385-1393)."""
    B = t.shape[0]
    tail = t.shape[2:]
    shp = [B] + [1] * st.d + list(tail)
    shp[1 + c] = n_line
    exp = [B] + list(st.dims) + list(tail)
    exp[1 + c] = n_line
    return t.reshape(shp).expand(exp).reshape(B, -1)


def derivative_values(st: Structure, steps_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """All derivative-row nnz values in construction order [central | forward | backward]  (:1618-1630)."""
    cv, fv, bv = [], [], []
    for c in range(st.d):
        s = steps_list[c]
        cv.append(_expand_line(st, c, central_line_values(s, st.order), st.dims[c]))
        fv.append(_expand_line(st, c, forward_line_values(s, st.order), st.dims[c] - 1))
        bv.append(_expand_line(st, c, backward_line_values(s, st.order), st.dims[c] - 1))
    return torch.cat(cv + fv + bv, dim=1)


# --------------------------------------------------------------------------------------------
# Assembly  (lp_pde_central_diff.py:1686-1781)
# --------------------------------------------------------------------------------------------
def eq_values(st: Structure, coeffs: np.ndarray) -> np.ndarray:
    """coeffs (B, G, M) -> (B, n_eq, M): remove_pad (:1686-1705)."""
    return coeffs.reshape(coeffs.shape[0], st.G, st.M)[:, st.eq_g, :]


def assemble_A(st: Structure, coeffs: np.ndarray, dvals: np.ndarray) -> List[sp.csr_matrix]:
    """One CSR A=[eq; init; deriv] per instance  (fill_constraints_torch :1766-1781)."""
    B = coeffs.shape[0]
    ev = eq_values(st, coeffs)
    e_row = np.repeat(np.arange(st.n_eq), st.M)
    e_col = (st.eq_g[:, None] * st.M + np.arange(st.M)[None, :]).reshape(-1)
    i_row = st.n_eq + np.arange(st.n_init)
    i_col = st.init_var
    dr = st.n_eq + st.n_init + st.d_row
    row = np.concatenate([e_row, i_row, dr])
    col = np.concatenate([e_col, i_col, st.d_col])
    out = []
    for b in range(B):
        val = np.concatenate([ev[b].reshape(-1), np.ones(st.n_init), dvals[b]])
        out.append(sp.coo_matrix((val, (row, col)), shape=(st.rows, st.n)).tocsr())
    return out


def assemble_b(st: Structure, rhs: np.ndarray, iv_rhs: np.ndarray) -> np.ndarray:
    """b = [rhs(interior) | iv_rhs | 0]  (:1738, :1759)."""
    B = rhs.shape[0]
    r = rhs.reshape(B, st.G)[:, st.eq_g]
    return np.concatenate([r, iv_rhs.reshape(B, -1), np.zeros((B, st.n_deriv))], axis=1)


# --------------------------------------------------------------------------------------------
# Grid transfer  (solver/multigrid.py:243-391; F.interpolate linear, align_corners=True)
# --------------------------------------------------------------------------------------------
def interp_linear_axis(x: np.ndarray, axis: int, n_out: int) -> np.ndarray:
    n_in = x.shape[axis]
    if n_out == n_in:
        return x
    scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
    src = scale * np.arange(n_out)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.where(i0 < n_in - 1, i0 + 1, i0)
    l1 = src - i0
    l0 = 1.0 - l1
    shp = [1] * x.ndim
    shp[axis] = n_out
    return np.take(x, i0, axis=axis) * l0.reshape(shp) + np.take(x, i1, axis=axis) * l1.reshape(shp)


def interp_linear(x: np.ndarray, new_dims: Sequence[int]) -> np.ndarray:
    """x (..., *dims) -> (..., *new_dims): separable linear interpolation, align_corners=True."""
    d = len(new_dims)
    for c in range(d):
        x = interp_linear_axis(x, x.ndim - d + c, int(new_dims[c]))
    return x


def coarse_steps(steps_list, dims, downsample_first):
    """pairwise sums of fine steps dropping the last  (multigrid.py:271-285)."""
    out = []
    for c, s in enumerate(steps_list):
        s = s.reshape(s.shape[0], dims[c] - 1)
        if c == 0 and not downsample_first:
            out.append(s)
        else:
            out.append(s[:, :-1].reshape(s.shape[0], dims[c] // 2 - 1, 2).sum(-1))
    return out


def coarse_iv(iv_rhs: np.ndarray, iv_list, old_dims, new_dims) -> np.ndarray:
    """multigrid.py:287-337."""
    B = iv_rhs.shape[0]
    iv_rhs = iv_rhs.reshape(B, -1)
    parts = []
    off = 0
    for f in iv_list:
        po = f(*old_dims)
        pn = f(*new_dims)
        so = tuple(int(v) for v in (np.array(po[3]) + 1 - np.array(po[2])))
        sn = tuple(int(v) for v in (np.array(pn[3]) + 1 - np.array(pn[2])))
        size = int(np.prod(so))
        blk = iv_rhs[:, off:off + size].reshape(B, *so)
        off += size
        parts.append(interp_linear(blk, sn).reshape(B, -1))
    if not parts:
        return np.zeros((B, 0))
    return np.concatenate(parts, axis=1)


def level_dims(dims, n_grid, downsample_first):
    """multigrid.py:88-102."""
    out = []
    cur = np.array(dims)
    for _ in range(n_grid):
        assert cur.min() >= 8
        out.append(tuple(int(v) for v in cur))
        cur = cur.copy()
        if downsample_first:
            cur = cur // 2
        else:
            cur[1:] = cur[1:] // 2
    return out


# --------------------------------------------------------------------------------------------
# Multigrid hierarchy + V-cycle  (solver/multigrid.py:115-240, 399-498)
# --------------------------------------------------------------------------------------------
@dataclass
class MGState:
    B: int
    dims_list: List[Tuple[int, ...]]
    st_list: List[Structure]
    K_list: list            # per level: block-diag CSR K (levels < last) or dense (B, n, n) (last)
    L_list: list            # tril(K) CSR per non-coarsest level
    U_list: list            # triu(K,1) CSR per non-coarsest level
    chol: Optional[list]    # coarsest Cholesky factors (list of (c, lower))
    A0: List[sp.csr_matrix]
    b0: np.ndarray
    Atb0: np.ndarray
    cfg: type = OracleConfig


def _vec_to_grid(v, B, st):
    """flat (B*n) -> (B, M, *dims)  (multigrid.py:345-351)."""
    return v.reshape(B, st.G, st.M).transpose(0, 2, 1).reshape(B, st.M, *st.dims)


def _grid_to_vec(x, B, st):
    return x.reshape(B, st.M, st.G).transpose(0, 2, 1).reshape(-1)


def mg_setup(dims, iv_list, coeffs, rhs, iv_rhs, steps_list, n_grid, downsample_first, order=2,
             cfg=OracleConfig) -> MGState:
    """fill_coarse_grids + make_AtA + make_coarse_AtA_matrices + factor_coarsest
    (qp_dual_sparse_multigrid_normal_kkt.py:28-47, multigrid.py:115-240,438-440)."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    B = coeffs.shape[0]
    dl = level_dims(dims, n_grid, downsample_first)
    st_list = [build_structure(dm, iv_list, order) for dm in dl]
    K_list, L_list, U_list = [], [], []
    cur_c = coeffs.reshape(B, st_list[0].G, st_list[0].M)
    cur_steps = [torch.as_tensor(np.asarray(s), dtype=torch.float64).reshape(B, -1) for s in steps_list]
    A0 = b0 = Atb0 = None
    chol = None
    for l, st in enumerate(st_list):
        if l > 0:
            prev = st_list[l - 1]
            g = cur_c.transpose(0, 2, 1).reshape(B, prev.M, *prev.dims)
            g = interp_linear(g, st.dims)                                       # downsample_coeffs :243-256
            cur_c = g.reshape(B, st.M, st.G).transpose(0, 2, 1)
            cur_steps = coarse_steps(cur_steps, prev.dims, downsample_first)    # :271-285
        dv = derivative_values(st, cur_steps).numpy()
        A = assemble_A(st, cur_c, dv)
        if l == 0:
            A0 = A
            b0 = assemble_b(st, np.asarray(rhs, dtype=np.float64), np.asarray(iv_rhs, dtype=np.float64))
            Atb0 = np.concatenate([A[b].T @ b0[b] for b in range(B)])
        last = (l == n_grid - 1) and n_grid > 1
        if last:
            Kd = np.stack([(A[b].T @ A[b]).toarray() for b in range(B)])       # :216-218
            K_list.append(Kd)
            chol = [sla.cho_factor(Kd[b], lower=True, check_finite=True) for b in range(B)]  # :438-440
        else:
            K = sp.block_diag([(A[b].T @ A[b]) for b in range(B)], format="csr")  # :224-226
            K_list.append(K)
            L_list.append(sp.tril(K, k=0, format="csr"))                          # :238
            U_list.append(sp.triu(K, k=1, format="csr"))                          # :239
    return MGState(B, dl, st_list, K_list, L_list, U_list, chol, A0, b0, Atb0, cfg)


def smooth_gs(L, U, b, x, nsteps):
    """x <- tril(K)^-1 (b - triu(K,1) x), nsteps times  (multigrid.py:399-405)."""
    for _ in range(nsteps):
        x = spla.spsolve_triangular(L, b - U @ x, lower=True)
    return x


def restrict(mg: MGState, l: int, r: np.ndarray) -> np.ndarray:
    """multigrid.py:340-363."""
    g = _vec_to_grid(r, mg.B, mg.st_list[l])
    return _grid_to_vec(interp_linear(g, mg.dims_list[l + 1]), mg.B, mg.st_list[l + 1])


def prolong(mg: MGState, l: int, x: np.ndarray) -> np.ndarray:
    """level l -> l-1  (multigrid.py:366-391)."""
    g = _vec_to_grid(x, mg.B, mg.st_list[l])
    return _grid_to_vec(interp_linear(g, mg.dims_list[l - 1]), mg.B, mg.st_list[l - 1])


def solve_coarsest(mg: MGState, b: np.ndarray) -> np.ndarray:
    """multigrid.py:442-450."""
    bb = b.reshape(mg.B, -1)
    return np.concatenate([sla.cho_solve(mg.chol[i], bb[i]) for i in range(mg.B)])


def v_cycle(mg: MGState, l: int, b: np.ndarray, x: np.ndarray) -> np.ndarray:
    """multigrid.py:453-487."""
    n_grid = len(mg.st_list)
    x = smooth_gs(mg.L_list[l], mg.U_list[l], b, x, mg.cfg.mg_gauss_seidel_steps_pre)
    r = b - mg.K_list[l] @ x
    rH = restrict(mg, l, r)
    if l == n_grid - 2:
        dH = solve_coarsest(mg, rH)
    else:
        dH = v_cycle(mg, l + 1, rH, np.zeros_like(rH))
    x = x + prolong(mg, l + 1, dH)
    x = smooth_gs(mg.L_list[l], mg.U_list[l], b, x, mg.cfg.mg_gauss_seidel_steps_post)
    return x


def v_cycle_start(mg: MGState, b: np.ndarray, back: bool = False) -> np.ndarray:
    """multigrid.py:490-498."""
    x = np.zeros_like(b)
    n_step = mg.cfg.mg_steps_backward if back else mg.cfg.mg_steps_forward
    for _ in range(n_step):
        x = v_cycle(mg, 0, b, x)
    return x


# --------------------------------------------------------------------------------------------
# FGMRES  (solver/fgmres.py:21-182)
# --------------------------------------------------------------------------------------------
def fgmres(K, b, precond, restart=10, maxiter=40, atol=1e-5, trace: Optional[dict] = None):
    """Restarted flexible GMRES, classical Gram-Schmidt, global norms over the whole batch."""
    n = b.shape[0]
    x = np.zeros_like(b)
    if np.linalg.norm(b) == 0:                          # :76-78
        return b, (0, 0.0)
    restart = min(restart, n)
    V = np.empty((n, restart))
    Z = np.empty((n, restart))
    H = np.zeros((restart + 1, restart))               # allocated once (:96-101)
    e = np.zeros(restart + 1)
    iters = 0
    while True:
        r = b - K @ x                                   # :124
        r_norm = np.linalg.norm(r)
        if trace is not None:
            trace.setdefault("r_norms", []).append(float(r_norm))
        if r_norm <= atol or iters >= maxiter:          # :134
            break
        v = r / r_norm
        V[:, 0] = v
        e[0] = r_norm
        for j in range(restart):                        # :141
            z = precond(v)
            Z[:, j] = z
            u = K @ z
            h = V[:, :j + 1].T @ u                      # classical Gram-Schmidt (:103-109)
            u = u - V[:, :j + 1] @ h
            H[:j + 1, j] = h
            H[j + 1, j] = np.linalg.norm(u)
            if j + 1 < restart:
                v = u / H[j + 1, j]
                V[:, j + 1] = v
        y = np.linalg.lstsq(H, e, rcond=None)[0]        # :166
        if trace is not None:
            trace.setdefault("H", []).append(H.copy())
            trace.setdefault("y", []).append(y.copy())
        x = x + Z @ y
        iters += restart
    return x, (iters, float(r_norm))


# --------------------------------------------------------------------------------------------
# Layer forward / backward  (qp_dual_sparse_multigrid_normal_kkt.py:25-162, qp_dual_dense_normal_kkt.py:23-118)
# --------------------------------------------------------------------------------------------
@dataclass
class LayerResult:
    x: np.ndarray                 # (B, n)
    lam: np.ndarray               # (B, rows)
    info_fwd: Tuple[int, float] = (0, 0.0)
    dz: Optional[np.ndarray] = None
    info_bwd: Tuple[int, float] = (0, 0.0)
    d_coeffs: Optional[np.ndarray] = None      # (B, G, M)
    d_rhs: Optional[np.ndarray] = None         # (B, G)
    d_iv_rhs: Optional[np.ndarray] = None      # (B, n_init)
    d_steps: Optional[List[np.ndarray]] = None  # list of (B, n_c-1)
    d_dvals: Optional[np.ndarray] = None       # (B, nnz_deriv) gradient w.r.t. derivative-row values


def _grads(st: Structure, A, x, lam, dz, steps_t, dvals_t, B):
    """Gradient formulas shared by both back-ends  (qp_...multigrid...:112-162)."""
    dnu = -np.stack([A[b] @ dz[b] for b in range(B)])                 # :112-114
    db = -dnu
    d_rhs = np.zeros((B, st.G))
    d_rhs[:, st.eq_g] = db[:, :st.n_eq]                                # add_pad :1632-1647 (fp64 here)
    d_iv = db[:, st.n_eq:st.n_eq + st.n_init]
    # dA = mask * (lam dz^T + dnu x^T)  (:132-136, lp_...:2047-2078)
    d_coeffs = np.zeros((B, st.G, st.M))
    xg = x.reshape(B, st.G, st.M)
    dzg = dz.reshape(B, st.G, st.M)
    d_coeffs[:, st.eq_g, :] = (lam[:, :st.n_eq, None] * dzg[:, st.eq_g, :]
                               + dnu[:, :st.n_eq, None] * xg[:, st.eq_g, :])
    # dD  (:141-159, lp_...:1971-1998)
    o = st.n_eq + st.n_init
    r = o + st.d_row
    d_dvals = lam[:, r] * dz[:, st.d_col] + dnu[:, r] * x[:, st.d_col]
    d_steps = None
    if steps_t is not None:
        g = torch.autograd.grad(dvals_t, steps_t, grad_outputs=torch.as_tensor(d_dvals), allow_unused=True)
        d_steps = [None if t is None else t.numpy() for t in g]
    return d_coeffs, d_rhs, d_iv, d_steps, d_dvals


def dense_layer(dims, iv_list, coeffs, rhs, iv_rhs, steps_list, grad_out=None, order=2) -> LayerResult:
    """PDEDenseLayer forward (+ backward for upstream gradient grad_out (B, n))  (qp_dual_dense_normal_kkt.py:23-118)."""
    st = build_structure(dims, iv_list, order)
    coeffs = np.asarray(coeffs, dtype=np.float64)
    B = coeffs.shape[0]
    steps_t = [torch.as_tensor(np.asarray(s), dtype=torch.float64).reshape(B, -1).clone().requires_grad_(True)
               for s in steps_list]
    dvals_t = derivative_values(st, steps_t)
    A = assemble_A(st, coeffs, dvals_t.detach().numpy())
    b = assemble_b(st, np.asarray(rhs, dtype=np.float64), np.asarray(iv_rhs, dtype=np.float64))
    x = np.empty((B, st.n))
    facs = []
    for i in range(B):
        Ad = A[i].toarray()
        K = Ad.T @ Ad                                                  # :30-33
        fac = sla.cho_factor(K, lower=True)                            # :39
        facs.append(fac)
        x[i] = sla.cho_solve(fac, Ad.T @ b[i])                         # :40
    lam = b - np.stack([A[i] @ x[i] for i in range(B)])               # :42-43
    res = LayerResult(x=x, lam=lam)
    if grad_out is not None:
        g = np.asarray(grad_out, dtype=np.float64).reshape(B, st.n)
        dz = np.stack([sla.cho_solve(facs[i], g[i]) for i in range(B)])  # :65
        res.dz = dz
        (res.d_coeffs, res.d_rhs, res.d_iv_rhs, res.d_steps, res.d_dvals) = _grads(st, A, x, lam, dz, steps_t, dvals_t, B)
    return res


def mg_layer(dims, iv_list, coeffs, rhs, iv_rhs, steps_list, n_grid, downsample_first, grad_out=None,
             order=2, cfg=OracleConfig, trace: Optional[dict] = None) -> LayerResult:
    """MultigridLayer forward (+ backward)  (qp_dual_sparse_multigrid_normal_kkt.py:25-162)."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    B = coeffs.shape[0]
    mg = mg_setup(dims, iv_list, coeffs, rhs, iv_rhs, steps_list, n_grid, downsample_first, order, cfg)
    st = mg.st_list[0]
    K0 = mg.K_list[0]
    tf = {} if trace is not None else None
    x, info = fgmres(K0, mg.Atb0, lambda v: v_cycle_start(mg, v, back=False),
                     restart=cfg.mg_fgmres_restarts_forward, maxiter=cfg.mg_fgmres_max_iter_forward, trace=tf)
    if trace is not None:
        trace["fwd"] = tf
    x = x.reshape(B, st.n)
    lam = mg.b0 - np.stack([mg.A0[i] @ x[i] for i in range(B)])      # :62-63
    res = LayerResult(x=x, lam=lam, info_fwd=info)
    if grad_out is not None:
        if callable(grad_out):          # upstream gradient as a function of the solution (loss = f(x))
            grad_out = grad_out(x)
        g = np.asarray(grad_out, dtype=np.float64).reshape(-1)
        tb = {} if trace is not None else None
        dz, info_b = fgmres(K0, g, lambda v: v_cycle_start(mg, v, back=True),
                            restart=cfg.mg_fgmres_restarts_backward, maxiter=cfg.mg_fgmres_max_iter_backward, trace=tb)
        if trace is not None:
            trace["bwd"] = tb
        dz = dz.reshape(B, st.n)
        res.dz = dz
        res.info_bwd = info_b
        steps_t = [torch.as_tensor(np.asarray(s), dtype=torch.float64).reshape(B, -1).clone().requires_grad_(True)
                   for s in steps_list]
        dvals_t = derivative_values(st, steps_t)
        (res.d_coeffs, res.d_rhs, res.d_iv_rhs, res.d_steps, res.d_dvals) = _grads(st, mg.A0, x, lam, dz, steps_t, dvals_t, B)
    return res
