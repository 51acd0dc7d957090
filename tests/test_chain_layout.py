"""CPU-only: the chain solver's algebra and index formulas (csrc/pdeop_cuda.cu: k_make_w, chain_row, k_band_chain),
restated in numpy with the same compact layouts, against a dense solve.

L = D W~ with D = blockdiag(L_kk) and W~ = D^-1 L (identity diagonal blocks), hence L^-1 = W~^-1 D^-1 and
L^-T = D^-T W~^-T: both triangular solves are chains of band GEMVs framed by block-diagonal products.
  Wc [r][pw]   row r of block row k holds W[r, c_lo(k) : k0],  c_lo(k) = max(0, k0 - pw),  pw = ceil32(bw)
  WTc[c][pwt]  row c of block column k holds W[(k+1)*blk :, c] (zero where W stores nothing)
"""
import numpy as np
import pytest

BLK = 128


def _chain_solve(A, bw, rhs):
    n = A.shape[0]
    L = np.linalg.cholesky(A)
    nblk = (n + BLK - 1) // BLK
    pw = (bw + 31) & ~31
    nbmax = (pw + BLK - 1) // BLK
    pwt = nbmax * BLK
    nw = nbmax + 1
    Linv = [np.linalg.inv(L[k * BLK:min(n, (k + 1) * BLK), k * BLK:min(n, (k + 1) * BLK)]) for k in range(nblk)]
    Wc = np.zeros((n, pw))
    WTc = np.zeros((n, pwt))
    for k in range(1, nblk):                      # k_make_w
        k0 = k * BLK
        w = min(BLK, n - k0)
        c_lo = max(0, k0 - pw)
        Wk = Linv[k] @ L[k0:k0 + w, c_lo:k0]
        Wc[k0:k0 + w, :k0 - c_lo] = Wk
        for cc in range(c_lo, k0):
            off = k0 - (cc // BLK + 1) * BLK
            WTc[cc, off:off + w] = Wk[:, cc - c_lo]

    def chain(direction, W, vin):                 # k_band_chain with its vector window of nw blocks
        win = np.zeros((nw, BLK))
        vout = np.zeros(n)
        for s in range(nblk):
            k = s if direction == 0 else nblk - 1 - s
            kn = k - 1 if direction == 0 else k + 1
            k0 = k * BLK
            new_block = np.zeros(BLK)
            for i in range(BLK):
                g = k0 + i
                if g >= n:
                    continue
                acc = 0.0
                if s > 0:                          # chain_row
                    if direction == 0:
                        c_lo = max(0, k0 - pw)
                        ncols = k0 - c_lo
                        new, old, v0 = W[g, ncols - BLK:ncols], W[g, :ncols - BLK], c_lo
                    else:
                        nl = min(n - (k + 1) * BLK, BLK)
                        new = np.zeros(BLK)
                        new[:nl] = W[g, :nl]
                        rho = min((g + pw) // BLK, nblk - 1)
                        ol = (rho - (k + 1)) * BLK if rho > k + 1 else 0
                        old, v0 = W[g, BLK:BLK + ol], (k + 2) * BLK
                    for e in range(len(old)):
                        gv = v0 + e
                        acc += old[e] * win[(gv // BLK) % nw, gv % BLK]
                    acc += new @ win[kn % nw]
                new_block[i] = vin[g] - acc
            w = min(BLK, n - k0)
            win[k % nw, :] = 0.0
            win[k % nw, :w] = new_block[:w]
            vout[k0:k0 + w] = new_block[:w]
        return vout

    u = np.concatenate([Linv[k] @ rhs[k * BLK:(k + 1) * BLK] for k in range(nblk)])
    y = chain(0, Wc, u)
    v = chain(1, WTc, y)
    return np.concatenate([Linv[k].T @ v[k * BLK:(k + 1) * BLK] for k in range(nblk)])


@pytest.mark.parametrize("n,bw", [(1280, 300), (1100, 300), (1536, 700), (1024, 100)])
def test_chain_algebra_and_layouts(n, bw):
    rng = np.random.default_rng(n + bw)
    A = np.zeros((n, n))
    for i in range(n):
        lo = max(0, i - bw)
        A[i, lo:i + 1] = 0.1 * rng.standard_normal(i + 1 - lo)
    A = A + A.T + np.eye(n) * (0.5 * bw + 5.0)
    rhs = rng.standard_normal(n)
    x = _chain_solve(A, bw, rhs)
    ref = np.linalg.solve(A, rhs)
    assert np.linalg.norm(x - ref) < 1e-12 * np.linalg.norm(ref)
