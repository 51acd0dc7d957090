"""Per-call values of the derivative ("smoothness") rows, per grid LINE position.

The reference expands these numbers to one value per nonzero of the derivative block
(solver/lp_pde_central_diff.py:1618-1630 build_derivative_values); the kernels here consume them
un-expanded: per coordinate c and line position i

    central  (B, n_c, 2, 6)   rows  sum_j w_j u(i+o_j) - h^k u_{c^k}(i) = 0,  k = 1, 2
    forward  (B, n_c-1, 4)    rows  u(i) + h u_c(i) + h^2/2 u_cc(i) - u(i+1) = 0
    backward (B, n_c-1, 4)    rows  u(i+1) - h u_c(i+1) + h^2/2 u_cc(i+1) - u(i) = 0

All of it stays in PyTorch so gradients reach ``steps_list`` by ordinary autograd (tiny tensors).
"""
import torch


def _fd_weights(nodes):
    """Finite-difference weights at 0 for 5 nodes: (..., 5) -> (..., 5, 2) [first, second derivative].

    Same construction as the reference (Vandermonde system solved with torch.linalg.solve,
    lp_pde_central_diff.py:1334-1341, 1415-1422) so both see identical rounding on identical devices.
    """
    sq = nodes * nodes
    vander = torch.stack([torch.ones_like(nodes), nodes, sq, sq * nodes, sq * sq], dim=-2)
    # right-hand sides e_1 and 2 e_2 (first and second derivative), built on the device without a host scalar copy
    k = torch.arange(5, device=nodes.device)
    target = torch.stack([(k == 1).to(nodes.dtype), 2.0 * (k == 2).to(nodes.dtype)], dim=-1)
    rhs = target.expand(*vander.shape[:-2], 5, 2)
    if vander.is_cuda and torch.cuda.is_current_stream_capturing():
        # same LU kernels without the singularity check, which reads a status back to the host (not allowed while a
        # CUDA graph is being captured; bench.py captures the dense workloads' whole step)
        return torch.linalg.solve_ex(vander, rhs, check_errors=False)[0]
    return torch.linalg.solve(vander, rhs)


def central_line_values(steps, order=2):
    """steps (B, n-1) -> (B, n, order, 6).

    Positions 0,1 and n-2,n-1 use one-sided 5-point stencils, the rest centred ones
    (lp_pde_central_diff.py:1000-1006).  The reference builds the one-sided node sets from
    steps[1:5]/steps[2:6] (left) and steps[-3:-1]... (right) rather than from the spacings adjacent to the
    point itself (lp_pde_central_diff.py:1304-1324); that choice is reproduced because it is part of A.
    """
    zero2 = torch.zeros_like(steps[:, :2])
    # left end: nodes 0, s1, s1+s2, ... with s_k = steps[:, k:k+2]
    acc = zero2
    nodes = [acc]
    for k in range(1, 5):
        acc = acc + steps[:, k:k + 2]
        nodes.append(acc)
    w_left = _fd_weights(torch.stack(nodes, dim=-1))
    h_left = steps[:, 1:3]
    # right end: nodes 0, -t1, -t1-t2, ... with t_k = steps[:, -(k+2):-k]
    acc = zero2
    nodes = [acc]
    n1 = steps.shape[1]
    for k in range(1, 5):
        acc = acc - steps[:, n1 - k - 2:n1 - k]
        nodes.append(acc)
    w_right = _fd_weights(torch.stack(nodes, dim=-1))
    h_right = steps[:, n1 - 3:n1 - 1]
    # interior positions 2..n-3
    nxt, nxt2 = steps[:, 2:-1], steps[:, 3:]
    prv, prv2 = steps[:, 1:-2], steps[:, :-3]
    w_mid = _fd_weights(torch.stack([-prv - prv2, -prv, torch.zeros_like(nxt), nxt, nxt + nxt2], dim=-1))
    h_mid = nxt
    w = torch.cat([w_left, w_mid, w_right], dim=1)           # (B, n, 5, 2)
    h = torch.cat([h_left, h_mid, h_right], dim=1).unsqueeze(-1)  # (B, n, 1)
    rows = []
    for k in range(1, order + 1):
        hk = h ** k
        rows.append(torch.cat([w[..., k - 1] * hk, -hk], dim=-1))   # (:1349-1350, :1429-1430)
    return torch.stack(rows, dim=2)


def forward_line_values(steps, order=2):
    """(B, n-1) -> (B, n-1, order+2) = [1, h, (h^2/2,) -1]  (lp_pde_central_diff.py:785-848, 1550-1581)."""
    one = torch.ones_like(steps)
    mid = [steps] if order == 1 else [steps, steps * steps / 2.0]
    return torch.stack([one] + mid + [-one], dim=-1)


def backward_line_values(steps, order=2):
    """(B, n-1) -> (B, n-1, order+2) = [1, -h, (h^2/2,) -1]; entry i belongs to line position i+1 (:849-861, 1583-1615)."""
    one = torch.ones_like(steps)
    mid = [-steps] if order == 1 else [-steps, steps * steps / 2.0]
    return torch.stack([one] + mid + [-one], dim=-1)


def line_values(steps_list, order=2):
    """[(B, n_c-1)]_c -> cv (B, Ntot, order, 6), fv (B, Ftot, order+2), bv (B, Ftot, order+2), coordinate-major."""
    cv = torch.cat([central_line_values(s, order) for s in steps_list], dim=1).contiguous()
    fv = torch.cat([forward_line_values(s, order) for s in steps_list], dim=1).contiguous()
    bv = torch.cat([backward_line_values(s, order) for s in steps_list], dim=1).contiguous()
    return cv, fv, bv


def embed_order(cv, fv, bv):
    """Line values of a total-order-1 system in the kernels' order-2 shapes: a zero second-derivative central row and
    a zero u_cc entry in the forward/backward rows (the native plan then keeps the u_cc unknowns decoupled)."""
    if cv.shape[2] == 2:
        return cv, fv, bv
    cv2 = torch.cat([cv, torch.zeros_like(cv)], dim=2)
    z = torch.zeros_like(fv[..., :1])
    fv2 = torch.cat([fv[..., :2], z, fv[..., 2:]], dim=-1)
    bv2 = torch.cat([bv[..., :2], z, bv[..., 2:]], dim=-1)
    return cv2.contiguous(), fv2.contiguous(), bv2.contiguous()


def coarsen_steps(steps_list, dims, downsample_first):
    """Coarse-level spacings: sums of consecutive fine pairs, last fine step dropped; the first axis only
    when it is coarsened too (multigrid.py:271-285)."""
    out = []
    for c, s in enumerate(steps_list):
        s = s.reshape(s.shape[0], dims[c] - 1)
        if c == 0 and not downsample_first:
            out.append(s)
        else:
            out.append(s[:, :-1].reshape(s.shape[0], dims[c] // 2 - 1, 2).sum(dim=-1))
    return out
