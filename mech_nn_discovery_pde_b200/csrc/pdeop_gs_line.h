// pdeop -- LINE-MARCHING wavefront Gauss-Seidel: per-thread bodies (host+device).
//
// Same sweep, same canonical arithmetic (gather_axis_fma / gs_back1_apply / gs_fin_compute of pdeop_elem.h) and
// therefore the same bits as the hyperplane kernels -- what changes is WHO updates a point and WHERE its operands
// come from.  Reference semantics: solver/multigrid.py:399-405 (x <- tril(K)^-1 (b - triu(K,1) x), lexicographic).
//
// Schedule.  A thread owns one grid LINE (i0, i1, *) and marches along it: at step n it finishes the point at
// virtual position v = n - sigma - prime, sigma = i0 + i1 + delta * cta (v = sweep * N2 + i2: the nu sweeps of a
// smoothing call follow each other on the line without a gap).  Points on one hyperplane i0+i1+i2 are finished in
// the same step, so every lexicographic dependency (points at distance <= 4 along the axes) was finished at least
// one step earlier; a CTA owns `RPC` whole grid rows (i0), an instance spans C CTAs.
//
// Operands.  Nothing is re-gathered through per-point index tables:
//   * u of the CTA's lines lives in a shared-memory RING [row][v mod 12][i1]: positions v-4 .. v+5 of every line,
//     new values written by the owner when it finishes a point, old values fetched by the owner 7 positions ahead with
//     cp.async (LDGSTS, no registers) -- all 24 u neighbours of a point along the three axes are plain shared-memory
//     reads at [row+o][v][i1], [row][v][i1+o], [row][v+o][i1];
//   * the derivative channels of the in-row axis use the same kind of ring (8 slots), those of the row axis a 4-slot
//     ring of NEW values (their two forward neighbours are read from global memory), those of the marching axis stay
//     in registers;
//   * the couplings at distance 3 and 4 to derivative unknowns exist only towards the two positions next to either
//     end of an axis (one-sided stencils, lp_pde_central_diff.py:1000-1006): those values come from a small 16-slot
//     ring of the four end lines of a row, from four registers (marching axis), or from two rows of global memory;
//   * b, equation coefficients and reciprocal diagonals are streamed once per point (coalesced: consecutive i1 are
//     consecutive in the wave layout) behind an L2 prefetch issued one step earlier; the K tables and the rowbase
//     columns of the rows a CTA reads are staged in shared memory once per call.
// Measured slower than the hyperplane kernels on every level (profiles/r2_gs_experiments.md); opt-in (gs_pipe = 5).
// A step is split as in the pipelined hyperplane kernel: A(n) finishes a point (the D backward distance-1 couplings,
// which were written in step n-1, then the sequential channel solve), ARRIVE, B(n) gathers everything else for the
// point of step n+1, WAIT.  The barrier is a CTA-wide split mbarrier; CTAs of an instance exchange rows through global
// memory (ld.global.cg) guarded by per-CTA progress counters, with `delta` steps of skew between neighbouring CTAs so
// that what a CTA needs from its predecessor was finished delta steps earlier.
#pragma once
#include "pdeop_elem.h"

namespace pdeop {

constexpr int kLineRU = 12;      // slots of the u ring: positions v-4 .. v+5 are live, v+6 and v+7 in flight
constexpr int kLineRE = 16;      // slots of the end-line rings
constexpr int kLineRP = 8;       // slots of the in-row derivative rings
constexpr int kLineRN = 4;       // slots of the row-axis derivative rings (new values only)
constexpr int kLineLaU = 7;      // look-ahead of the u ring: position v+7 is requested in B(n)
constexpr int kLineLaP = 5;      // look-ahead of the in-row derivative rings
constexpr int kLinePrime = 8;    // steps before the first line starts (fills the rings)
constexpr int kLineDelta = 3;    // extra skew per CTA of an instance
constexpr int kLinePad = 8;      // zero doubles on either side of every ring
constexpr int kLineMaxThreads = 512;
constexpr int kLineMinExtent = 12;   // the low and the high end zone of an axis must not overlap

struct LineGeom {
    int N1, N2;      // lines per row, points per line
    int RPC;         // rows per CTA
    int C;           // CTAs per instance
    int LPC;         // lines per CTA = RPC * N1
    int NL;          // lines per instance = N0 * N1
    int VT;          // virtual positions per line = nsweeps * N2
    int NS;          // steps per call
    int threads;     // threads per CTA (LPC rounded up to whole warps)
    // shared-memory offsets in doubles: K tables, u ring, in-row derivative rings, row-axis new-value rings, end-line
    // rings, parking row for the previous CTA's last row
    int o_tab, o_ru, o_rp, o_rq, o_np, o_nq, o_ep, o_eq, o_park, o_rb, o_end;
    int NC;          // columns of the shared-memory copy of the rowbase table: rows r0-4 .. r0+RPC+3 of the grid
};

// Per-channel base pointers of the batch's vectors (channel-planar [instance][channel][wave index]): with the
// instance's offset folded into the 32-bit index, every global access is one [pointer constant + 32-bit index * 8].
struct LineStreams {
    const double* b[7];
    const double* coef[7];
    const double* dinv[7];
    double* x[7];
};
inline void line_streams(const LevelDev& L, const double* coef, const double* dinv, const double* b, double* x,
                         LineStreams& S) {
    for (int m = 0; m < 7; ++m) {
        const size_t o = (size_t)(m < L.M ? m : 0) * L.G;
        S.b[m] = b + o;
        S.coef[m] = coef + o;
        S.dinv[m] = dinv + o;
        S.x[m] = x + o;
    }
}

// Geometry of one smoothing call; returns false when the level does not fit this kernel (the caller falls back to the
// hyperplane kernels).  ps: table pitch in shared memory (max extent + 8 <= ps).
inline bool line_geom(const LevelDev& L, int nsweeps, int ps, int max_cluster, size_t max_smem, int max_threads,
                      LineGeom& g) {
    if (L.D < 2 || ps <= 0 || nsweeps < 1) return false;
    const int N0 = L.N[0], N1 = L.N[1], N2 = L.N[2];
    if ((size_t)L.M * L.G >= ((size_t)1 << 31)) return false;
    if (max_threads > kLineMaxThreads) max_threads = kLineMaxThreads;
    if (N1 > max_threads || N1 < kLineMinExtent || N2 < 16) return false;
    if (L.D == 3 && N0 < kLineMinExtent) return false;
    int rpc = max_threads / N1;
    if (rpc > N0) rpc = N0;
    g.N1 = N1;
    g.N2 = N2;
    g.C = (N0 + rpc - 1) / rpc;
    if (g.C > max_cluster) return false;
    // balance the rows over the CTAs (e.g. 32 rows, 3 CTAs of <= 12 instead of 16+16)
    g.RPC = (N0 + g.C - 1) / g.C;
    // a CTA reads rows up to 4 away: they must belong to its direct neighbours
    if (g.C > 1 && g.RPC < 4) return false;
    g.LPC = g.RPC * N1;
    g.NL = N0 * N1;
    g.VT = nsweeps * N2;
    g.NS = kLinePrime + g.VT + (N0 - 1) + (N1 - 1) + kLineDelta * (g.C - 1);
    g.threads = (g.LPC + 31) / 32 * 32;
    int o = 0;
    g.o_tab = o;
    o += L.D * kTabEntries * ps + kLinePad;
    g.o_ru = o;
    o += g.RPC * kLineRU * N1 + kLinePad;
    g.o_rp = o;
    o += g.RPC * kLineRP * N1 + kLinePad;
    g.o_rq = o;
    o += g.RPC * kLineRP * N1 + kLinePad;
    g.o_np = o;
    if (L.D == 3) o += g.RPC * kLineRN * N1 + kLinePad;
    g.o_nq = o;
    if (L.D == 3) o += g.RPC * kLineRN * N1 + kLinePad;
    g.o_ep = o;
    o += g.RPC * 4 * kLineRE;
    g.o_eq = o;
    o += g.RPC * 4 * kLineRE;
    g.o_park = o;
    if (L.D == 3) o += 3 * N1;
    g.o_rb = o;
    g.NC = g.RPC + 8;
    o += ((L.S + 8) * g.NC + 1) / 2;      // ints
    g.o_end = o;
    return (size_t)o * sizeof(double) + 256 <= max_smem;
}

// ring addresses (doubles from the start of shared memory); rr = local row in [0, RPC), vv = virtual position
// u ring: slot = virtual position mod kLineRU (LineCtx::su tracks it; line_uslot adds an offset)
PDEOP_HD int line_ru(const LineGeom& g, int rr, int i1, int slot) {
    return g.o_ru + (rr * kLineRU + slot) * g.N1 + i1;
}
PDEOP_HD int line_uslot(int su, int k) {   // slot of virtual position v + k, -kLineRU < k < kLineRU
    int t = su + k;
    if (t >= kLineRU) t -= kLineRU;
    if (t < 0) t += kLineRU;
    return t;
}
PDEOP_HD int line_rp(const LineGeom& g, int rr, int i1, int vv) {
    return g.o_rp + (rr * kLineRP + (vv & (kLineRP - 1))) * g.N1 + i1;
}
PDEOP_HD int line_rq(const LineGeom& g, int rr, int i1, int vv) {
    return g.o_rq + (rr * kLineRP + (vv & (kLineRP - 1))) * g.N1 + i1;
}
PDEOP_HD int line_np(const LineGeom& g, int rr, int i1, int vv) {
    return g.o_np + (rr * kLineRN + (vv & (kLineRN - 1))) * g.N1 + i1;
}
PDEOP_HD int line_nq(const LineGeom& g, int rr, int i1, int vv) {
    return g.o_nq + (rr * kLineRN + (vv & (kLineRN - 1))) * g.N1 + i1;
}
// end-line rings: k = 0, 1, 2, 3 <-> i1 = 0, 1, N1-2, N1-1
PDEOP_HD int line_ep(const LineGeom& g, int rr, int k, int vv) {
    return g.o_ep + (rr * 4 + k) * kLineRE + (vv & (kLineRE - 1));
}
PDEOP_HD int line_eq(const LineGeom& g, int rr, int k, int vv) {
    return g.o_eq + (rr * 4 + k) * kLineRE + (vv & (kLineRE - 1));
}
PDEOP_HD int line_end_index(int i, int n) { return i <= 1 ? i : (i >= n - 2 ? i - (n - 4) : -1); }
// the two end positions a point at position i couples to at distance 3 or 4: 0,1 or n-2,n-1 (returns the first)
PDEOP_HD int line_far_first(int i, int n) { return i <= 5 ? 0 : n - 2; }

template <int D>
struct LineCtx {
    int i0, i1, r;        // grid row, position in the row, row within the CTA
    int v;                // virtual position of the point A finishes in the current step (threads without a line: far
                          // below zero, so that nothing ever becomes active)
    int i2;               // v mod N2 (once v >= 0)
    int su;               // v mod kLineRU: slot of the u ring
    int w;                // wave index of the point A finishes in the current step (set by the previous B)
    int wh1, wh2;         // wave indices of the next two points of the line (at the start of B: of v+1 and v+2)
    double res[1 + 2 * D];            // partial residual: b - all couplings but the backward distance-1 ones
    double cc[1 + 2 * D], di[1 + 2 * D];
    double po0, qo0;      // old row-axis derivative values of that point (D == 3)
    double po2, qo2;      // old marching-axis derivative values of that point
    double h1p, h1q, h2p, h2q;   // old marching-axis derivative values of the next two points (as wh1, wh2)
    double pn1, qn1, pn2, qn2;   // new marching-axis derivative values at v and v-1 (after A)
    double fp0, fq0, fp1, fq1;   // marching-axis derivative values at the two end positions in reach (0,1 / N2-2,N2-1)
};

// position on the line of virtual position v + k (0 <= k < N2), or a negative number before the line starts
template <int D>
PDEOP_HD int line_pos(const LineCtx<D>& c, const LineGeom& g, int k) {
    if (c.v < 0) return c.v + k;
    const int p = c.i2 + k;
    return p >= g.N2 ? p - g.N2 : p;
}

template <int D>
PDEOP_HD void line_init(const LevelDev& L, const LineGeom& g, int cta, int tid, LineCtx<D>& c) {
    const int line = cta * g.LPC + tid;
    const bool valid = tid < g.LPC && line < g.NL;
    c.i0 = valid ? line / g.N1 : 0;
    c.i1 = valid ? line - c.i0 * g.N1 : 0;
    c.r = valid ? c.i0 - cta * g.RPC : 0;
    c.v = valid ? -(c.i0 + c.i1 + kLineDelta * cta) - kLinePrime : -(1 << 28);
    c.i2 = 0;
    c.su = ((c.v % kLineRU) + kLineRU) % kLineRU;
    c.w = c.wh1 = c.wh2 = 0;
#pragma unroll
    for (int m = 0; m < 1 + 2 * D; ++m) c.res[m] = c.cc[m] = c.di[m] = 0.0;
    c.po0 = c.qo0 = c.po2 = c.qo2 = c.h1p = c.h1q = c.h2p = c.h2q = 0.0;
    c.pn1 = c.qn1 = c.pn2 = c.qn2 = 0.0;
    c.fp0 = c.fq0 = c.fp1 = c.fq1 = 0.0;
    (void)L;
}

// end of a step: advance the virtual position
template <int D>
PDEOP_HD void line_advance(const LineGeom& g, LineCtx<D>& c) {
    c.v += 1;
    c.su = c.su + 1 == kLineRU ? 0 : c.su + 1;
    if (c.v >= 1) c.i2 = c.i2 + 1 == g.N2 ? 0 : c.i2 + 1;
}

// Shared memory of one CTA before the first step: the instance's K tables with pitch PS, zeroed rings (out-of-range
// neighbours must read finite values), and the rowbase columns of the rows this CTA reads (r0-4 .. r0+RPC+3; zero for
// rows outside the grid: such neighbours only ever meet exact-zero table entries).  Strided over the CTA's threads.
template <int D, int PS>
PDEOP_HD void line_smem_fill(const LevelDev& L, const LineGeom& g, const double* __restrict__ Ti, int cta, double* sm,
                             int tid, int nthr) {
    int maxn = L.N[0] > L.N[1] ? L.N[0] : L.N[1];
    maxn = (maxn > L.N[2] ? maxn : L.N[2]) + 2 * kTabPad;
    double* Ts = sm + g.o_tab;
    for (int i = tid; i < D * kTabEntries * PS; i += nthr) {
        const int row = i / PS, pos = i - row * PS;
        Ts[i] = pos < maxn ? Ti[(size_t)row * kTabPitch + pos] : 0.0;
    }
    for (int i = g.o_tab + D * kTabEntries * PS + tid; i < g.o_rb; i += nthr) sm[i] = 0.0;
    int* rbs = reinterpret_cast<int*>(sm + g.o_rb);
    const int nrb = (L.S + 8) * g.NC, N0 = L.N[0], r0 = cta * g.RPC;
    for (int i = tid; i < nrb; i += nthr) {
        const int s4 = i / g.NC, col = i - s4 * g.NC;
        const int row = r0 - 4 + col;
        rbs[i] = (row >= 0 && row < N0) ? L.rowbase[s4 * N0 + row] : 0;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A(n): finish the point at virtual position v.  x: the instance's iterate in global memory, sm: shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <int D, int PS>
PDEOP_HD void line_fin(const LevelDev& L, const LineGeom& g, double* sm, double* x, LineCtx<D>& c) {
    if (c.v < 0 || c.v >= g.VT) return;
    constexpr int M = 1 + 2 * D;
    constexpr int mp1 = 1 + (D - 2), mq1 = 1 + D + (D - 2);   // in-row axis channels
    constexpr int mp2 = 1 + (D - 1), mq2 = 1 + D + (D - 1);   // marching axis channels
    const int v = c.v, r = c.r, i1 = c.i1;
    const int au = line_ru(g, r, i1, c.su), ap = line_rp(g, r, i1, v), aq = line_rq(g, r, i1, v);
    FinLoads<D> f;
    f.xl[0] = sm[au];
    f.xl[mp1] = sm[ap];
    f.xl[mq1] = sm[aq];
    f.xl[mp2] = c.po2;
    f.xl[mq2] = c.qo2;
    if (D == 3) {
        f.xl[1] = c.po0;
        f.xl[1 + D] = c.qo0;
        // row above: this CTA's rings, or (first row of a CTA) the parking row filled by this thread in B(n-1)
        f.nb.un[0] = sm[r > 0 ? au - kLineRU * g.N1 : g.o_park + i1];
        f.nb.pn[0] = sm[r > 0 ? line_np(g, r - 1, i1, v) : g.o_park + g.N1 + i1];
        f.nb.qn[0] = sm[r > 0 ? line_nq(g, r - 1, i1, v) : g.o_park + 2 * g.N1 + i1];
    }
    f.nb.un[D - 2] = sm[au - 1];
    f.nb.pn[D - 2] = sm[ap - 1];
    f.nb.qn[D - 2] = sm[aq - 1];
    f.nb.un[D - 1] = sm[line_ru(g, r, i1, line_uslot(c.su, -1))];
    f.nb.pn[D - 1] = c.pn1;
    f.nb.qn[D - 1] = c.qn1;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        f.c[m] = c.cc[m];
        f.di[m] = c.di[m];
    }
    gs_fin_compute<D, PS>(L, sm + g.o_tab, x, c.w, c.i0, c.i1, c.i2, f, c.res);
    sm[au] = f.xl[0];
    sm[ap] = f.xl[mp1];
    sm[aq] = f.xl[mq1];
    if (D == 3) {
        sm[line_np(g, r, i1, v)] = f.xl[1];
        sm[line_nq(g, r, i1, v)] = f.xl[1 + D];
    }
    const int k = line_end_index(i1, g.N1);
    if (k >= 0) {
        sm[line_ep(g, r, k, v)] = f.xl[mp1];
        sm[line_eq(g, r, k, v)] = f.xl[mq1];
    }
    c.pn2 = c.pn1;
    c.qn2 = c.qn1;
    c.pn1 = f.xl[mp2];
    c.qn1 = f.xl[mq2];
    if (c.i2 == 0) {
        c.fp0 = c.pn1;
        c.fq0 = c.qn1;
    }
    if (c.i2 == 1) {
        c.fp1 = c.pn1;
        c.fq1 = c.qn1;
    }
}

// The off-point couplings of one axis without the backward distance-1 neighbour, in the canonical order of
// gather_axis_fma<D, PITCH, SPLIT = true> (pdeop_elem.h) -- same operations on the same operands, hence the same bits --
// with one shortcut: the far (|o| = 3, 4) couplings to derivative unknowns are exact zeros unless the neighbour is one
// of the two positions next to an end of the axis (stencil_offsets), so only those two are multiplied (skipping
// fma(0, x, s) leaves s unchanged).  un: u at o = -4..-1, 1..4 (j = 0..7, j = 3 unused); pn, qn: derivative values at
// o = -2, +1, +2 (e = 0, 1, 2); pa, qa / pb, qb: at positions first / first + 1 (read only when axis_end3(i)).
template <int D, int PITCH>
PDEOP_HD void line_axis_fma(const double* __restrict__ T, int n, int a, int i, const double un[8], const double pn[3],
                            const double qn[3], int first, double pa, double qa, double pb, double qb,
                            double acc[1 + 2 * D]) {
    const double* __restrict__ Ta = axis_table<PITCH>(T, a) + (i + kTabPad);
    double au = 0.0, ap = 0.0, aq = 0.0;
#pragma unroll
    for (int j = 2; j < 6; ++j) {
        const int o = j < 4 ? j - 4 : j - 3;
        if (o == -1) continue;
        const int e = o == -2 ? 0 : o;
        au = fma(Ta[(T_UU + o + 4) * PITCH], un[j], au);
        au = fma(Ta[(T_UP - o + 4) * PITCH + o], pn[e], au);
        au = fma(Ta[(T_UQ - o + 4) * PITCH + o], qn[e], au);
        ap = fma(Ta[(T_UP + o + 4) * PITCH], un[j], ap);
        aq = fma(Ta[(T_UQ + o + 4) * PITCH], un[j], aq);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = j < 4 ? j - 4 : j - 3;
        if (o >= -2 && o <= 2) continue;
        au = fma(Ta[(T_UU + o + 4) * PITCH], un[j], au);
    }
    if (axis_end1(i, n)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = j < 4 ? j - 4 : j - 3;
            if (o >= -2 && o <= 2) continue;
            ap = fma(Ta[(T_UP + o + 4) * PITCH], un[j], ap);
            aq = fma(Ta[(T_UQ + o + 4) * PITCH], un[j], aq);
        }
    }
    if (axis_end3(i, n)) {
        const int oa = first - i, ob = oa + 1;     // ascending offsets, as the canonical loop visits them
        if ((oa >= -4 && oa <= -3) || (oa >= 3 && oa <= 4)) {
            au = fma(Ta[(T_UP - oa + 4) * PITCH + oa], pa, au);
            au = fma(Ta[(T_UQ - oa + 4) * PITCH + oa], qa, au);
        }
        if ((ob >= -4 && ob <= -3) || (ob >= 3 && ob <= 4)) {
            au = fma(Ta[(T_UP - ob + 4) * PITCH + ob], pb, au);
            au = fma(Ta[(T_UQ - ob + 4) * PITCH + ob], qb, au);
        }
    }
    acc[0] += au;
    acc[1 + a] += ap;
    acc[1 + D + a] += aq;
}

// ------------------------------------------------------------------------------------------------------------------
// B(n): everything the point at virtual position v+1 needs except its backward distance-1 neighbours; ring look-ahead.
// IO: ldcg (read at the L2 coherence point: rows other CTAs write), cp8 (asynchronous 8-byte copy global -> shared),
// prefetch (into L2).
// ------------------------------------------------------------------------------------------------------------------
template <int D, int PS, class IO>
PDEOP_HD void line_pre(const LevelDev& L, const LineGeom& g, double* sm, const LineStreams& S, unsigned ioff,
                       LineCtx<D>& c, IO& io) {
    constexpr int M = 1 + 2 * D;
    constexpr int mp1 = 1 + (D - 2), mq1 = 1 + D + (D - 2);
    constexpr int mp2 = 1 + (D - 1), mq2 = 1 + D + (D - 1);
    const unsigned gm1 = (unsigned)L.G - 1u;
    const int N0 = L.N[0], i0 = c.i0, i1 = c.i1, r = c.r;
    // rowbase[(s + 4) * N0 + row] of the grid rows r0-4 .. r0+RPC+3 at rbs[(s + 4) * NC + row - r0 + 4]
    const int* __restrict__ rbs = reinterpret_cast<const int*>(sm + g.o_rb);
    const int NC = g.NC, col = r + 4, s0 = i0 + i1 + 4;
    // ---- ring look-ahead (independent of whether this thread has a point to prepare) ----
    {
        const int tu = c.v + kLineLaU;
        if (tu >= 0 && tu < g.VT) {
            const unsigned w = (unsigned)(rbs[(s0 + line_pos(c, g, kLineLaU)) * NC + col] + i1) + ioff;
            io.cp8(sm + line_ru(g, r, i1, line_uslot(c.su, kLineLaU)), S.x[0] + w);
            // the derivative values this line and its row-axis neighbours read with plain loads: into L2 now
            io.prefetch(S.x[mp2] + w);
            io.prefetch(S.x[mq2] + w);
            if (D == 3) {
                io.prefetch(S.x[1] + w);
                io.prefetch(S.x[1 + D] + w);
            }
            const int k = line_end_index(i1, g.N1);
            if (k >= 0) {
                io.cp8(sm + line_ep(g, r, k, tu), S.x[mp1] + w);
                io.cp8(sm + line_eq(g, r, k, tu), S.x[mq1] + w);
            }
        }
        const int tp = c.v + kLineLaP;
        if (tp >= 0 && tp < g.VT) {
            const unsigned w = (unsigned)(rbs[(s0 + line_pos(c, g, kLineLaP)) * NC + col] + i1) + ioff;
            io.cp8(sm + line_rp(g, r, i1, tp), S.x[mp1] + w);
            io.cp8(sm + line_rq(g, r, i1, tp), S.x[mq1] + w);
        }
    }
    const int vq = c.v + 1;
    // ---- marching-axis derivative window: the old values two positions ahead of the point being prepared ----
    // (at the start of B: h1 = old values at v+1, h2 = at v+2; f2 = at v+3 is loaded now)
    double f2p = 0.0, f2q = 0.0;
    int w2 = 0;
    {
        const int t2 = vq + 2;
        if (t2 >= 0 && t2 < g.VT) {
            w2 = rbs[(s0 + line_pos(c, g, 3)) * NC + col] + i1;
            f2p = io.ldown(S.x[mp2] + ((unsigned)w2 + ioff));    // own line: only this thread writes it
            f2q = io.ldown(S.x[mq2] + ((unsigned)w2 + ioff));
        }
    }
    if (vq >= 0 && vq < g.VT) {
        const int i2q = line_pos(c, g, 1);
        const unsigned wq = (unsigned)c.wh1;
        const int* __restrict__ rb = rbs + (s0 + i2q) * NC + col;    // neighbour row o at rb[o * (NC + 1)]
        const int suq = line_uslot(c.su, 1);                          // u-ring slot of the point being prepared
        // L2 prefetch of the streamed operands of the NEXT point (loaded at the end of the next B)
        if (vq + 1 < g.VT) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                io.prefetch(S.b[m] + ((unsigned)c.wh2 + ioff));
                io.prefetch(S.coef[m] + ((unsigned)c.wh2 + ioff));
                io.prefetch(S.dinv[m] + ((unsigned)c.wh2 + ioff));
            }
        }
        // ---- global operands of the row axis, issued before the rings are read ----
        double gu[8], gpn[3], gqn[3];     // u of rows outside the CTA; p, q at o = -2 (outside the CTA), +1, +2
        double fra = 0.0, frb = 0.0, fsa = 0.0, fsb = 0.0;   // p, q of the two end rows in reach
        bool e3r = false;
        int far0 = 0;
        if (D == 3) {
            const double* __restrict__ xu = S.x[0];
            const double* __restrict__ xp = S.x[1];
            const double* __restrict__ xq = S.x[1 + D];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = j < 4 ? j - 4 : j - 3;
                gu[j] = 0.0;
                if (o == -1) continue;
                const int rr = r + o;
                const bool in = rr >= 0 && rr < g.RPC;
                const unsigned wn = clamp_wave(rb[o * (NC + 1)] + i1, gm1) + ioff;
                if (!in) gu[j] = io.ldcg(xu + wn);
                if (o == -2) {
                    gpn[0] = gqn[0] = 0.0;
                    if (!in) {
                        gpn[0] = io.ldcg(xp + wn);
                        gqn[0] = io.ldcg(xq + wn);
                    }
                } else if (o == 1 || o == 2) {
                    gpn[o] = io.ldcg(xp + wn);
                    gqn[o] = io.ldcg(xq + wn);
                }
            }
            // previous CTA's last row, backward distance 1: finished delta steps ago; parked for A(n+1)
            if (r == 0 && i0 > 0) {
                const unsigned wn = clamp_wave(rb[-(NC + 1)] + i1, gm1) + ioff;
                sm[g.o_park + i1] = io.ldcg(xu + wn);
                sm[g.o_park + g.N1 + i1] = io.ldcg(xp + wn);
                sm[g.o_park + 2 * g.N1 + i1] = io.ldcg(xq + wn);
            }
            e3r = axis_end3(i0, N0);
            far0 = line_far_first(i0, N0);
            if (e3r) {
                const int* __restrict__ rf = rb + (far0 - i0) * (NC + 1);
                const unsigned wa = (unsigned)(rf[0] + i1) + ioff;
                const unsigned wb = (unsigned)(rf[NC + 1] + i1) + ioff;
                fra = io.ldcg(xp + wa);
                fsa = io.ldcg(xq + wa);
                frb = io.ldcg(xp + wb);
                fsb = io.ldcg(xq + wb);
            }
        }
        double acc0[M], acc1[M], acc2[M];
#pragma unroll
        for (int m = 0; m < M; ++m) acc0[m] = acc1[m] = acc2[m] = 0.0;
        const double* __restrict__ Ts = sm + g.o_tab;
        // ---- in-row axis (internal axis 1): rings ----
        {
            double un[8], pn[3], qn[3];
            const int first = line_far_first(i1, g.N1), k0 = i1 <= 5 ? 0 : 2;
            const int au = line_ru(g, r, i1, suq), ap = line_rp(g, r, i1, vq), aq = line_rq(g, r, i1, vq);
            const int ae = line_ep(g, r, k0, vq), af = line_eq(g, r, k0, vq);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = j < 4 ? j - 4 : j - 3;
                un[j] = o == -1 ? 0.0 : sm[au + o];
            }
            pn[0] = sm[ap - 2];
            pn[1] = sm[ap + 1];
            pn[2] = sm[ap + 2];
            qn[0] = sm[aq - 2];
            qn[1] = sm[aq + 1];
            qn[2] = sm[aq + 2];
            line_axis_fma<D, PS>(Ts, g.N1, D - 2, i1, un, pn, qn, first, sm[ae], sm[af], sm[ae + kLineRE],
                                 sm[af + kLineRE], acc1);
        }
        // ---- marching axis (internal axis 2): u from the ring, derivative values from registers ----
        {
            double un[8], pn[3], qn[3];
            const int ub = line_ru(g, r, i1, 0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = j < 4 ? j - 4 : j - 3;
                un[j] = o == -1 ? 0.0 : sm[ub + line_uslot(suq, o) * g.N1];
            }
            pn[0] = c.pn2;
            pn[1] = c.h2p;
            pn[2] = f2p;
            qn[0] = c.qn2;
            qn[1] = c.h2q;
            qn[2] = f2q;
            line_axis_fma<D, PS>(Ts, g.N2, D - 1, i2q, un, pn, qn, line_far_first(i2q, g.N2), c.fp0, c.fq0, c.fp1,
                                 c.fq1, acc2);
        }
        // ---- row axis (internal axis 0): this CTA's rows from the rings, the others from the loads above ----
        if (D == 3) {
            double un[8], pn[3], qn[3];
            const int au = line_ru(g, r, i1, suq);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = j < 4 ? j - 4 : j - 3;
                un[j] = gu[j];
                if (o == -1) continue;
                const int rr = r + o;
                if (rr >= 0 && rr < g.RPC) un[j] = sm[au + o * (kLineRU * g.N1)];
            }
            pn[0] = gpn[0];
            qn[0] = gqn[0];
            if (r >= 2) {
                pn[0] = sm[line_np(g, r - 2, i1, vq)];
                qn[0] = sm[line_nq(g, r - 2, i1, vq)];
            }
            pn[1] = gpn[1];
            pn[2] = gpn[2];
            qn[1] = gqn[1];
            qn[2] = gqn[2];
            line_axis_fma<D, PS>(Ts, N0, 0, i0, un, pn, qn, far0, fra, fsa, frb, fsb, acc0);
        }
        // ---- streamed operands of this point (prefetched into L2 one step ago) ----
        double bl[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            bl[m] = io.ldstream(S.b[m] + (wq + ioff));
            c.cc[m] = io.ldstream(S.coef[m] + (wq + ioff));
            c.di[m] = io.ldstream(S.dinv[m] + (wq + ioff));
        }
        const bool eq = coord_eq(L.coord[wq]);
        if (D == 3) {
            c.po0 = io.ldown(S.x[1] + (wq + ioff));
            c.qo0 = io.ldown(S.x[1 + D] + (wq + ioff));
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            // the axis sums in the canonical order: ((0 + row axis) + in-row axis) + marching axis
            const double acc = D == 3 ? (acc0[m] + acc1[m]) + acc2[m] : acc1[m] + acc2[m];
            c.res[m] = bl[m] - acc;
            if (!eq) c.cc[m] = 0.0;
        }
        // the far marching-axis values switch from the low end (positions 0, 1: written by A) to the high end
        if (i2q == 6) {
            const unsigned wa = (unsigned)(rbs[(s0 + g.N2 - 2) * NC + col] + i1) + ioff;
            const unsigned wb = (unsigned)(rbs[(s0 + g.N2 - 1) * NC + col] + i1) + ioff;
            c.fp0 = io.ldown(S.x[mp2] + wa);
            c.fq0 = io.ldown(S.x[mq2] + wa);
            c.fp1 = io.ldown(S.x[mp2] + wb);
            c.fq1 = io.ldown(S.x[mq2] + wb);
        }
    }
    // shift the marching-axis window: next point's old values become this point's, two-ahead becomes next
    c.w = c.wh1;
    c.wh1 = c.wh2;
    c.wh2 = w2;
    c.po2 = c.h1p;
    c.qo2 = c.h1q;
    c.h1p = c.h2p;
    c.h1q = c.h2q;
    c.h2p = f2p;
    c.h2q = f2q;
}

}  // namespace pdeop
