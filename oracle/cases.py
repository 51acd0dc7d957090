"""Shared seeded test cases (iv_lists and synthetic inputs)  --  TEST INFRASTRUCTURE ONLY.

iv_lists restate the callers' boundary/initial specifications:
  gl        discovery/ginzburg_landau.py:225-237
  burgers   discovery/burgers_dparam_viscous.py:170-173
  kamani    discovery/kamani.py:153-156
  sine      fit/sine_pde_dense.py:111-115
  transport examples/2. sparse_multigrid_solver_transport.ipynb (cell 3)
Each lambda returns (coord, mi_index, range_begin[], range_end[]) evaluated at a level's dims
(solver/multigrid.py:296-306).
"""
import numpy as np
import torch

IV_LISTS = {
    "gl": [
        lambda nt, nx, ny: (0, 0, [0, 0, 0], [0, nx - 1, ny - 1]),
        lambda nt, nx, ny: (1, 0, [1, 0, 0], [nt - 1, 0, ny - 1]),
        lambda nt, nx, ny: (2, 0, [1, 1, 0], [nt - 1, nx - 1, 0]),
        lambda nt, nx, ny: (1, 0, [1, nx - 1, 1], [nt - 1, nx - 1, ny - 1]),
        lambda nt, nx, ny: (2, 0, [1, 1, ny - 1], [nt - 1, nx - 2, ny - 1]),
    ],
    "burgers": [
        lambda nx, ny: (0, 0, [0, 0], [0, ny - 2]),
        lambda nx, ny: (1, 0, [1, 0], [nx - 1, 0]),
        lambda nx, ny: (1, 0, [0, ny - 1], [nx - 1, ny - 1]),
    ],
    "kamani": [
        lambda nt: (0, 0, [0], [0]),
    ],
    "sine": [
        lambda nx, ny: (0, 0, [0, 0], [0, ny - 2]),
        lambda nx, ny: (1, 0, [1, 0], [nx - 1, 0]),
        lambda nx, ny: (0, 0, [nx - 1, 1], [nx - 1, ny - 2]),
        lambda nx, ny: (1, 0, [0, ny - 1], [nx - 1, ny - 1]),
    ],
    "transport": [
        lambda nt, nx: (0, 0, [0, 0], [0, nx - 1]),
    ],
}


def make_inputs(dims, bs, n_init, seed, uniform=False):
    """Seeded fp64 inputs: coeffs (bs,G,M), rhs (bs,G), iv_rhs (bs,n_init), steps [(bs,n_c-1)], loss weights (bs,1,G,M)."""
    d = len(dims)
    M = 1 + 2 * d
    G = int(np.prod(dims))
    g = torch.Generator().manual_seed(seed)
    coeffs = 0.3 * torch.randn(bs, G, M, generator=g, dtype=torch.float64)
    coeffs[..., 1] += 1.0
    coeffs[..., 1 + d:] -= 0.5
    rhs = 0.1 * torch.randn(bs, G, generator=g, dtype=torch.float64)
    iv = 0.5 * torch.randn(bs, n_init, generator=g, dtype=torch.float64)
    base = [0.1, 0.3, 0.25][:d]
    if uniform:
        steps = [torch.full((bs, n - 1), h, dtype=torch.float64) for n, h in zip(dims, base)]
    else:
        steps = [h * (0.75 + 0.5 * torch.rand(bs, n - 1, generator=g, dtype=torch.float64)) for n, h in zip(dims, base)]
    loss_w = torch.randn(bs, 1, G, M, generator=g, dtype=torch.float64)
    return dict(coeffs=coeffs.numpy(), rhs=rhs.numpy(), iv_rhs=iv.numpy(), steps=[s.numpy() for s in steps],
                loss_w=loss_w.numpy())
