// pdeop -- per-element bodies of the stencil kernels (host+device).
//
// Every function here handles ONE element (a grid point of one instance, a line position, ...).
// The CUDA kernels in pdeop_cuda.cu map threads onto these bodies; the test-only CPU emulator in
// tests/emu loops over them, so index algebra can be checked against the oracle without a GPU.
//
// Reference semantics restated here (paths relative to the reference repo):
//   row structure / values     solver/lp_pde_central_diff.py:746-1033, 1300-1630
//   K = A^T A, A^T b           solver/multigrid.py:210-240
//   lexicographic Gauss-Seidel solver/multigrid.py:399-405
//   linear grid transfer       solver/multigrid.py:243-268, 340-391 (F.interpolate, align_corners=True)
//   gradients                  solver/qp_dual_sparse_multigrid_normal_kkt.py:112-162
#pragma once
#include <math.h>
#include "pdeop_common.h"

// Load policy for vectors that other thread blocks of the SAME kernel may be writing (the in-place
// Gauss-Seidel iterate): ld.global.cg reads at the L2 coherence point.  Everything else uses LdPlain.
struct LdPlain {
    static PDEOP_HD double ld(const double* p) { return *p; }
};
struct LdL2 {
    static PDEOP_HD double ld(const double* p) {
#if defined(__CUDA_ARCH__)
        return __ldcg(p);
#else
        return *p;
#endif
    }
};

// Streaming load (ld.global.cs: evict-first in L1 and L2) for data a sweep touches once per point (right-hand
// side, equation coefficients, reciprocal diagonals): keeps the L2 for the iterate, whose every value is gathered
// 24 times within +-4 wavefront steps.  PDEOP_NO_STREAM_HINT disables it (A/B testing).
PDEOP_HD double ld_stream(const double* p) {
#if defined(__CUDA_ARCH__) && defined(PDEOP_STREAM_NOALLOC)
    // A/B variant (tools/gs_bench.py, PDEOP_VARIANT_SO): no L1 allocation at all for the streams
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#elif defined(__CUDA_ARCH__) && !defined(PDEOP_NO_STREAM_HINT)
    return __ldcs(p);
#else
    return *p;
#endif
}

#ifndef PDEOP_ATOMIC_ADD
#if defined(__CUDA_ARCH__)
#define PDEOP_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define PDEOP_ATOMIC_ADD(p, v) (*(p) += (v))
#endif
#endif

// A/B switch (tools/apply_micro.py): one gather axis in flight in K apply (86 registers; 80 and 3 CTAs/SM with
// -DPDEOP_APPLY_MINB=3) instead of two (98 registers).  Measured: 345 vs 360 us per fine-level launch -- occupancy
// is not what limits the kernel; default unchanged.
#ifndef PDEOP_APPLY_OVERLAP
#define PDEOP_APPLY_OVERLAP true
#endif

namespace pdeop {

// 5-point stencil offsets of the central-difference row at line position i (lp_pde_central_diff.py:1000-1006):
// one-sided at the two positions next to either end, centred otherwise.
PDEOP_HD void stencil_offsets(int i, int n, int o[5]) {
    if (i <= 1) {
        for (int j = 0; j < 5; ++j) o[j] = j;
    } else if (i >= n - 2) {
        for (int j = 0; j < 5; ++j) o[j] = -j;
    } else {
        for (int j = 0; j < 5; ++j) o[j] = j - 2;
    }
}

// ------------------------------------------------------------------------------------------------
// Axis tables of K = A^T A.  One call per (instance, active axis a, padded position ip).
//   cv [Ntot][2][6]  central rows: 5 stencil values on u + value on the own derivative channel
//   fv [Ftot][4]     forward  rows at position i   : u[i], u_c[i], u_cc[i], u[i+1]
//   bv [Ftot][4]     backward rows at position i+1 : u[i+1], u_c[i+1], u_cc[i+1], u[i]
// Output Ta[e*P + ip], e in [0,30):
//   T_UU+o+4 : K[u(i), u(i+o)]     T_UP+o+4 : K[u_c(i), u(i+o)]     T_UQ+o+4 : K[u_cc(i), u(i+o)]
//   T_PP, T_QQ, T_PQ : K[u_c(i),u_c(i)], K[u_cc(i),u_cc(i)], K[u_c(i),u_cc(i)]      (axis part only)
// ------------------------------------------------------------------------------------------------
PDEOP_HD void build_table_elem(const LevelDev& L, int a, int ip, const double* cv, const double* fv,
                               const double* bv, double* Ta) {
    const int n = L.N[3 - L.D + a];
    const int i = ip - kTabPad;
    double t[kTabEntries];
    for (int e = 0; e < kTabEntries; ++e) t[e] = 0.0;
    if (i >= 0 && i < n) {
        const double* cva = cv + (size_t)L.cvoff[a] * 12;
        const double* fva = fv + (size_t)L.fvoff[a] * 4;
        const double* bva = bv + (size_t)L.fvoff[a] * 4;
        int lo = i - 4 < 0 ? 0 : i - 4;
        int hi = i + 4 > n - 1 ? n - 1 : i + 4;
        for (int ir = lo; ir <= hi; ++ir) {
            int o[5];
            stencil_offsets(ir, n, o);
            int ji = -1;
            for (int j = 0; j < 5; ++j)
                if (ir + o[j] == i) ji = j;
            if (ji < 0) continue;
            for (int k = 0; k < 2; ++k) {
                const double* row = cva + ((size_t)ir * 2 + k) * 6;
                const double wi = row[ji];
                for (int j2 = 0; j2 < 5; ++j2) t[T_UU + (ir + o[j2] - i) + 4] += wi * row[j2];
            }
        }
        {
            int o[5];
            stencil_offsets(i, n, o);
            const double* r1 = cva + ((size_t)i * 2 + 0) * 6;
            const double* r2 = cva + ((size_t)i * 2 + 1) * 6;
            for (int j = 0; j < 5; ++j) {
                t[T_UP + o[j] + 4] += r1[5] * r1[j];
                t[T_UQ + o[j] + 4] += r2[5] * r2[j];
            }
            t[T_PP] += r1[5] * r1[5];
            t[T_QQ] += r2[5] * r2[5];
        }
        if (i <= n - 2) {  // forward row at i
            const double* f = fva + (size_t)i * 4;
            t[T_UU + 4] += f[0] * f[0];
            t[T_UU + 5] += f[0] * f[3];
            t[T_UP + 4] += f[1] * f[0];
            t[T_UP + 5] += f[1] * f[3];
            t[T_UQ + 4] += f[2] * f[0];
            t[T_UQ + 5] += f[2] * f[3];
            t[T_PP] += f[1] * f[1];
            t[T_QQ] += f[2] * f[2];
            t[T_PQ] += f[1] * f[2];
        }
        if (i >= 1) {  // forward row at i-1 touches u(i) as "next"
            const double* f = fva + (size_t)(i - 1) * 4;
            t[T_UU + 4] += f[3] * f[3];
            t[T_UU + 3] += f[0] * f[3];
        }
        if (i >= 1) {  // backward row at i
            const double* g = bva + (size_t)(i - 1) * 4;
            t[T_UU + 4] += g[0] * g[0];
            t[T_UU + 3] += g[0] * g[3];
            t[T_UP + 4] += g[1] * g[0];
            t[T_UP + 3] += g[1] * g[3];
            t[T_UQ + 4] += g[2] * g[0];
            t[T_UQ + 3] += g[2] * g[3];
            t[T_PP] += g[1] * g[1];
            t[T_QQ] += g[2] * g[2];
            t[T_PQ] += g[1] * g[2];
        }
        if (i <= n - 2) {  // backward row at i+1 touches u(i) as "previous"
            const double* g = bva + (size_t)i * 4;
            t[T_UU + 4] += g[3] * g[3];
            t[T_UU + 5] += g[0] * g[3];
        }
    }
    if (L.order == 1 && i >= 0 && i < n) t[T_QQ] += 1.0;   // decoupled unit diagonal on the unused u_cc channel
    for (int e = 0; e < kTabEntries; ++e) Ta[(size_t)e * kTabPitch + ip] = t[e];
}

// wave index of the neighbour at offset o along internal axis AX
template <int AX>
PDEOP_HD int neighbor_pos(const LevelDev& L, int s, int i0, int i1, int o) {
    if (AX == 2) return L.rowbase[(s + o + 4) * L.N[0] + i0] + i1;
    if (AX == 1) return L.rowbase[(s + o + 4) * L.N[0] + i0] + i1 + o;
    return L.rowbase[(s + o + 4) * L.N[0] + i0 + o] + i1;
}

// table of (instance, active axis a): entry e of line position i is Tab(T,a)[e*kTabPitch + i + kTabPad]
template <int PITCH>
PDEOP_HD const double* axis_table(const double* T, int a) { return T + a * (kTabEntries * PITCH); }

// ------------------------------------------------------------------------------------------------
// acc[m] = sum over OFF-POINT couplings  K[(g,m),(g',m')] x[g',m'],  g' = g + o e_a, o in [-4,4]\{0}
// T: instance tables [D][30][kTabPitch];  x: instance vector [M][G].
// Branch-free: out-of-range neighbours are predicated loads (value 0); their coefficients are exact zeros
// because the tables are zero-padded by 4 positions on either side.  All table loads are
// [position pointer + compile-time offset]; vector loads use one pointer per channel plane and axis.
// ------------------------------------------------------------------------------------------------
// PITCH: table pitch (kTabPitch for tables in global memory, a smaller compile-time pitch for the copy the
// Gauss-Seidel kernel stages in shared memory); rowbase: L.rowbase or its shared-memory copy.
// keeps the compiler from sinking a batch of loads down to their first use (device only)
#if defined(__CUDA_ARCH__)
#define PDEOP_LOAD_FENCE() asm volatile("" ::: "memory")
#else
#define PDEOP_LOAD_FENCE()
#endif

// Clamp a (possibly out-of-range or negative) wave index into [0, G).  An out-of-range neighbour is only ever
// multiplied by an exact-zero table entry, so its value is irrelevant as long as the address is valid and the
// value finite: one unsigned minimum replaces the bounds predicates, zero fills and 64-bit sign extensions.
PDEOP_HD unsigned clamp_wave(int w, unsigned gm1) {
    const unsigned u = (unsigned)w;
    return u < gm1 ? u : gm1;
}

// wave index (unclamped) of the neighbour at offset o along internal axis ax; rb = rowbase + (s+4)*N0 + i0
PDEOP_HD int neighbor_wave(const int* __restrict__ rb, int ax, int N0, int i1, int o) {
    if (ax == 2) return rb[o * N0] + i1;          // row (s+o, i0)
    if (ax == 1) return rb[o * N0] + i1 + o;      // row (s+o, i0), column i1+o
    return rb[o * N0 + o] + i1;                   // row (s+o, i0+o)
}

// Neighbour values of one axis (offsets o = -4..-1, 1..4 at j = 0..7).
struct AxisNb {
    double un[8], pn[8], qn[8];
};

// Structural zeros of K along an axis of extent n (lp_pde_central_diff.py:1000-1006): a derivative row reaches 3
// or 4 positions away only through the one-sided stencils of the two positions next to either end.  Hence, for
// |o| >= 3:  K[u_c/u_cc(i), u(i+o)] = 0 unless i is such an end position      (axis_end1)
//            K[u(i), u_c/u_cc(i+o)] = 0 unless i+o is                         (axis_end3: i in 3..5 or n-6..n-4)
// Pure index tests: the skipped terms are exact zeros.
PDEOP_HD bool axis_end1(int i, int n) { return i <= 1 || i >= n - 2; }
PDEOP_HD bool axis_end3(int i, int n) { return (i >= 3 && i <= 5) || (i >= n - 6 && i <= n - 4); }

// Issue the loads of one axis.  SPLIT: without the backward distance-1 neighbour (o = -1), see k_gather.
template <int D, class LD, bool SPLIT>
PDEOP_HD void gather_axis_load(const LevelDev& L, const int* __restrict__ rb, const double* x, int a, int i, int i1,
                               AxisNb& nb) {
    const unsigned G = (unsigned)L.G, gm1 = G - 1u;
    const int ax = 3 - D + a;
    const int n = L.N[ax], N0 = L.N[0];
    const unsigned gp = (unsigned)(1 + a) * G, gq = (unsigned)(1 + D + a) * G;   // plane offsets (M*G < 2^31)
    const bool e3 = axis_end3(i, n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = j < 4 ? j - 4 : j - 3;
        if (SPLIT && o == -1) continue;
        const unsigned wn = clamp_wave(neighbor_wave(rb, ax, N0, i1, o), gm1);
        nb.un[j] = LD::ld(x + wn);
        if (o >= -2 && o <= 2) {
            nb.pn[j] = LD::ld(x + (wn + gp));
            nb.qn[j] = LD::ld(x + (wn + gq));
        } else {
            nb.pn[j] = e3 ? LD::ld(x + (wn + gp)) : 0.0;
            nb.qn[j] = e3 ? LD::ld(x + (wn + gq)) : 0.0;
        }
    }
}

// Canonical FMA order per axis (every kernel that inlines this body, and the host emulator, produce the same bits):
//   near offsets o = -2,-1,1,2 : au += UU u, au += UPn p, au += UQn q, ap += UP u, aq += UQ u
//   far u terms  o = -4,-3,3,4 : au += UU u
//   if axis_end1(i), far       : ap += UP u, aq += UQ u
//   if axis_end3(i), far       : au += UPn p, au += UQn q
template <int D, int PITCH, bool SPLIT>
PDEOP_HD void gather_axis_fma(const LevelDev& L, const double* __restrict__ T, int a, int i, const AxisNb& nb,
                              double acc[1 + 2 * D]) {
    const int n = L.N[3 - D + a];
    const double* __restrict__ Ta = axis_table<PITCH>(T, a) + (i + kTabPad);
    double au = 0.0, ap = 0.0, aq = 0.0;
#pragma unroll
    for (int j = 2; j < 6; ++j) {
        const int o = j < 4 ? j - 4 : j - 3;
        if (SPLIT && o == -1) continue;
        au = fma(Ta[(T_UU + o + 4) * PITCH], nb.un[j], au);
        au = fma(Ta[(T_UP - o + 4) * PITCH + o], nb.pn[j], au);
        au = fma(Ta[(T_UQ - o + 4) * PITCH + o], nb.qn[j], au);
        ap = fma(Ta[(T_UP + o + 4) * PITCH], nb.un[j], ap);
        aq = fma(Ta[(T_UQ + o + 4) * PITCH], nb.un[j], aq);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = j < 4 ? j - 4 : j - 3;
        if (o >= -2 && o <= 2) continue;
        au = fma(Ta[(T_UU + o + 4) * PITCH], nb.un[j], au);
    }
    if (axis_end1(i, n)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = j < 4 ? j - 4 : j - 3;
            if (o >= -2 && o <= 2) continue;
            ap = fma(Ta[(T_UP + o + 4) * PITCH], nb.un[j], ap);
            aq = fma(Ta[(T_UQ + o + 4) * PITCH], nb.un[j], aq);
        }
    }
    if (axis_end3(i, n)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = j < 4 ? j - 4 : j - 3;
            if (o >= -2 && o <= 2) continue;
            au = fma(Ta[(T_UP - o + 4) * PITCH + o], nb.pn[j], au);
            au = fma(Ta[(T_UQ - o + 4) * PITCH + o], nb.qn[j], au);
        }
    }
    acc[0] += au;
    acc[1 + a] += ap;
    acc[1 + D + a] += aq;
}

// acc[m] = sum over OFF-POINT couplings  K[(g,m),(g',m')] x[g',m'],  g' = g + o e_a.
// SPLIT = false: all offsets o in [-4,4] \ {0}.
// SPLIT = true : all but the backward distance-1 neighbours (o = -1), which gs_back1 adds.  Under wavefront
//   Gauss-Seidel those are the only values written in the immediately preceding step, so everything gathered
//   here can be done one step ahead, off the critical path of the sweep.
// The loads of axis a+1 are issued before the FMAs of axis a (at most two axes in flight): the kernels built on
// this body are bound by dependent memory round trips, not by bandwidth.
// BSUB: also loads b (issued with the last axis batch, when the registers of the first axes are free again) and
// returns acc[m] = b[m] - sum instead of the sum.
template <int D, class LD, int PITCH, bool SPLIT, bool BSUB, bool OVERLAP = true>
PDEOP_HD void k_gather(const LevelDev& L, const int* __restrict__ rowbase, const double* __restrict__ T,
                       const double* x, const double* __restrict__ b, unsigned w, int i0, int i1, int i2,
                       double acc[1 + 2 * D]) {
    constexpr int M = 1 + 2 * D;
    const int idx[3] = {i0, i1, i2};
    const int* __restrict__ rb = rowbase + (i0 + i1 + i2 + 4) * L.N[0] + i0;
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = 0.0;
    AxisNb nb[D];
    double bl[M];
    if (OVERLAP) gather_axis_load<D, LD, SPLIT>(L, rb, x, 0, idx[3 - D], i1, nb[0]);
#pragma unroll
    for (int a = 0; a < D; ++a) {
        if (OVERLAP) {
            if (a + 1 < D) gather_axis_load<D, LD, SPLIT>(L, rb, x, a + 1, idx[3 - D + a + 1], i1, nb[a + 1]);
        } else {   // one axis in flight: fewer registers, more resident warps (streaming kernels without barriers)
            gather_axis_load<D, LD, SPLIT>(L, rb, x, a, idx[3 - D + a], i1, nb[a]);
        }
        if (BSUB && a == D - 1) {
#pragma unroll
            for (int m = 0; m < M; ++m) bl[m] = ld_stream(b + ((unsigned)m * (unsigned)L.G + w));
        }
        PDEOP_LOAD_FENCE();
        gather_axis_fma<D, PITCH, SPLIT>(L, T, a, idx[3 - D + a], nb[a], acc);
    }
    if (BSUB) {
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = bl[m] - acc[m];
    }
}

// Backward distance-1 neighbours (o = -1 along every active axis): loads, then r[m] -= K[m, neighbour] x[neighbour].
// Split in two so that a caller can put these loads in flight together with its other loads.
template <int D>
struct Back1 {
    double un[D], pn[D], qn[D];
};

template <int D, class LD>
PDEOP_HD void gs_back1_load(const LevelDev& L, const int* __restrict__ rowbase, const double* x, int i0, int i1, int i2,
                            Back1<D>& nb) {
    const unsigned G = (unsigned)L.G, gm1 = G - 1u;
    const int N0 = L.N[0];
    const int* __restrict__ rb = rowbase + (i0 + i1 + i2 + 4) * N0 + i0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const unsigned wn = clamp_wave(neighbor_wave(rb, 3 - D + a, N0, i1, -1), gm1);
        nb.un[a] = LD::ld(x + wn);
        nb.pn[a] = LD::ld(x + (wn + (unsigned)(1 + a) * G));
        nb.qn[a] = LD::ld(x + (wn + (unsigned)(1 + D + a) * G));
    }
}

template <int D, int PITCH>
PDEOP_HD void gs_back1_apply(const double* __restrict__ T, int i0, int i1, int i2, const Back1<D>& nb,
                             double r[1 + 2 * D]) {
    const int idx[3] = {i0, i1, i2};
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const double* __restrict__ Ta = axis_table<PITCH>(T, a) + (idx[3 - D + a] + kTabPad);
        r[0] = fma(-Ta[(T_UU + 3) * PITCH], nb.un[a], r[0]);
        r[0] = fma(-Ta[(T_UP + 5) * PITCH - 1], nb.pn[a], r[0]);
        r[0] = fma(-Ta[(T_UQ + 5) * PITCH - 1], nb.qn[a], r[0]);
        r[1 + a] = fma(-Ta[(T_UP + 3) * PITCH], nb.un[a], r[1 + a]);
        r[1 + D + a] = fma(-Ta[(T_UQ + 3) * PITCH], nb.un[a], r[1 + D + a]);
    }
}

// per-point local quantities: equation-row coefficients c[m] (zero where the point has no equation
// row), initial-row multiplicities, and the axis tables' same-point entries.
template <int D>
struct PointLocal {
    double c[1 + 2 * D];
    double ini[1 + 2 * D];
    double uu;          // sum over axes of K[u,u] axis part
    double up[D], uq[D], pp[D], qq[D], pq[D];
};

template <int D, int PITCH>
PDEOP_HD void load_axis_local(const double* __restrict__ T, int i0, int i1, int i2, PointLocal<D>& pl) {
    const int idx[3] = {i0, i1, i2};
    pl.uu = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const double* __restrict__ Ta = axis_table<PITCH>(T, a) + (idx[3 - D + a] + kTabPad);
        pl.uu += Ta[(T_UU + 4) * PITCH];
        pl.up[a] = Ta[(T_UP + 4) * PITCH];
        pl.uq[a] = Ta[(T_UQ + 4) * PITCH];
        pl.pp[a] = Ta[T_PP * PITCH];
        pl.qq[a] = Ta[T_QQ * PITCH];
        pl.pq[a] = Ta[T_PQ * PITCH];
    }
}

template <int D>
PDEOP_HD void load_local(const LevelDev& L, const double* __restrict__ T, const double* __restrict__ coef, int w,
                         int i0, int i1, int i2, int flags, PointLocal<D>& pl) {
    const int G = L.G;
    const bool eq = flags & 1;
#pragma unroll
    for (int m = 0; m < 1 + 2 * D; ++m) {
        pl.c[m] = eq ? coef[m * G + w] : 0.0;   // (ld_stream here makes K apply 25 % slower: measured)
        pl.ini[m] = (double)((flags >> (4 + 2 * m)) & 3);
    }
    load_axis_local<D, kTabPitch>(T, i0, i1, i2, pl);
}

// diagonal of K at channel m of a point
template <int D>
PDEOP_HD double k_diag(const PointLocal<D>& pl, int m) {
    double d = pl.c[m] * pl.c[m] + pl.ini[m];
    if (m == 0) d += pl.uu;
    else if (m <= D) d += pl.pp[m - 1];
    else d += pl.qq[m - 1 - D];
    return d;
}

// dinv[m][w] = 1 / K[(w,m),(w,m)]
template <int D>
PDEOP_HD void dinv_elem(const LevelDev& L, const double* __restrict__ T, const double* __restrict__ coef,
                        double* __restrict__ dinv, int w) {
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    PointLocal<D> pl;
    load_local<D>(L, T, coef, w, i0, i1, i2, L.flags[w], pl);
#pragma unroll
    for (int m = 0; m < 1 + 2 * D; ++m) dinv[m * L.G + w] = 1.0 / k_diag<D>(pl, m);
}

// y = K x at one point (mode 0) or y = b - K x (mode 1)
template <int D>
PDEOP_HD void apply_k_elem(const LevelDev& L, const double* __restrict__ T, const double* __restrict__ coef,
                           const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ y,
                           int w, int mode) {
    constexpr int M = 1 + 2 * D;
    const int G = L.G;
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    const int flags = L.flags[w];
    // own-point operands first: the gather's load fences would otherwise pin them behind it as one more dependent
    // memory round trip at the end
    PointLocal<D> pl;
    double xl[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        xl[m] = x[m * G + w];
        pl.c[m] = (flags & 1) ? coef[m * G + w] : 0.0;
    }
    double acc[M];
    k_gather<D, LdPlain, kTabPitch, false, false, PDEOP_APPLY_OVERLAP>(L, L.rowbase, T, x, nullptr, 0u, i0, i1, i2, acc);
#pragma unroll
    for (int m = 0; m < M; ++m) pl.ini[m] = (double)((flags >> (4 + 2 * m)) & 3);
    load_axis_local<D, kTabPitch>(T, i0, i1, i2, pl);
    double cs = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) cs += pl.c[m] * xl[m];
    double yl[M];
#pragma unroll
    for (int m = 0; m < M; ++m) yl[m] = acc[m] + pl.c[m] * cs + pl.ini[m] * xl[m];
    yl[0] += pl.uu * xl[0];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        yl[0] += pl.up[a] * xl[1 + a] + pl.uq[a] * xl[1 + D + a];
        yl[1 + a] += pl.up[a] * xl[0] + pl.pp[a] * xl[1 + a] + pl.pq[a] * xl[1 + D + a];
        yl[1 + D + a] += pl.uq[a] * xl[0] + pl.pq[a] * xl[1 + a] + pl.qq[a] * xl[1 + D + a];
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int k = m * G + w;
        y[k] = mode ? (b[k] - yl[m]) : yl[m];
    }
}

// One lexicographic Gauss-Seidel update of the M unknowns of point w (channel order 0..M-1):
//   x_j <- (b_j - sum_{k != j} K_jk x_k) / K_jj   with already-updated values for k < j.
// Split in two halves so that the wavefront kernel can run the first one a step ahead:
//   gs_pre_elem : r[m] = b[m] - (all off-point couplings except the backward distance-1 neighbours)
//   gs_fin_elem : own-point loads and the backward distance-1 loads in one batch, r -= those couplings, then the
//                 sequential channel solve and the store.
// The equation-row part of the point block is rank one (c c^T): the running sum S = c.x is kept up to
// date as channels are updated, and the division is a multiplication by the precomputed reciprocal diagonal.
template <int D, class LD, int PITCH>
PDEOP_HD void gs_pre_elem(const LevelDev& L, const int* __restrict__ rowbase, const double* __restrict__ T,
                          const double* __restrict__ b, const double* x, int w, int i0, int i1, int i2,
                          double r[1 + 2 * D]) {
    k_gather<D, LD, PITCH, true, true>(L, rowbase, T, x, b, (unsigned)w, i0, i1, i2, r);
}

// operands of the finishing half: old own-point values, equation coefficients, reciprocal diagonals, backward
// distance-1 neighbours -- one batch of loads
template <int D>
struct FinLoads {
    double xl[1 + 2 * D], c[1 + 2 * D], di[1 + 2 * D];
    Back1<D> nb;
};

template <int D, class LD>
PDEOP_HD void gs_fin_load(const LevelDev& L, const int* __restrict__ rowbase, const double* __restrict__ coef,
                          const double* __restrict__ dinv, const double* x, int w, int i0, int i1, int i2, bool eq,
                          FinLoads<D>& f) {
    constexpr int M = 1 + 2 * D;
    const unsigned G = (unsigned)L.G, uw = (unsigned)w;
    gs_back1_load<D, LD>(L, rowbase, x, i0, i1, i2, f.nb);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        f.xl[m] = LD::ld(x + ((unsigned)m * G + uw));
        f.c[m] = eq ? ld_stream(coef + ((unsigned)m * G + uw)) : 0.0;
        f.di[m] = ld_stream(dinv + ((unsigned)m * G + uw));
    }
}

template <int D, int PITCH>
PDEOP_HD void gs_fin_compute(const LevelDev& L, const double* __restrict__ T, double* x, int w, int i0, int i1, int i2,
                             FinLoads<D>& f, double r[1 + 2 * D]) {
    constexpr int M = 1 + 2 * D;
    const unsigned G = (unsigned)L.G, uw = (unsigned)w;
    double* xl = f.xl;
    const double* c = f.c;
    const double* di = f.di;
    gs_back1_apply<D, PITCH>(T, i0, i1, i2, f.nb, r);
    PointLocal<D> pl;
    load_axis_local<D, PITCH>(T, i0, i1, i2, pl);
    // Canonical arithmetic (every kernel that inlines this body, and the host emulator, produce the same bits):
    //   S   = ((c_u x_u + s_0) + s_1) + ...,   s_a = fma(c_qa, x_qa, c_pa x_pa)
    //   u   : off = c_u t + l_0 + l_1 + ...,   l_a = fma(uq_a, x_qa, up_a x_pa)
    //   p_a : off = fma(pq_a, x_qa, fma(up_a, x_u, c_pa t));   q_a : off = fma(pq_a, x_pa, fma(uq_a, x_u, c_qa t))
    //   with t = fma(-c_m, x_m, S),  x_m <- (r_m - off) * dinv_m,  S <- fma(c_m, x_m, t)
    double S = c[0] * xl[0];
#pragma unroll
    for (int a = 0; a < D; ++a) S = S + fma(c[1 + D + a], xl[1 + D + a], c[1 + a] * xl[1 + a]);
    {   // u
        const double t = fma(-c[0], xl[0], S);
        double off = c[0] * t;
#pragma unroll
        for (int a = 0; a < D; ++a) off = off + fma(pl.uq[a], xl[1 + D + a], pl.up[a] * xl[1 + a]);
        const double xn = (r[0] - off) * di[0];
        S = fma(c[0], xn, t);
        xl[0] = xn;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {   // u_c channels
        const int m = 1 + a;
        const double t = fma(-c[m], xl[m], S);
        const double off = fma(pl.pq[a], xl[1 + D + a], fma(pl.up[a], xl[0], c[m] * t));
        const double xn = (r[m] - off) * di[m];
        S = fma(c[m], xn, t);
        xl[m] = xn;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {   // u_cc channels
        const int m = 1 + D + a;
        const double t = fma(-c[m], xl[m], S);
        const double off = fma(pl.pq[a], xl[1 + a], fma(pl.uq[a], xl[0], c[m] * t));
        const double xn = (r[m] - off) * di[m];
        S = fma(c[m], xn, t);
        xl[m] = xn;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) x[(unsigned)m * G + uw] = xl[m];
}

template <int D, class LD, int PITCH>
PDEOP_HD void gs_fin_elem(const LevelDev& L, const int* __restrict__ rowbase, const double* __restrict__ T,
                          const double* __restrict__ coef, const double* __restrict__ dinv, double* x, int w, int i0,
                          int i1, int i2, bool eq, double r[1 + 2 * D]) {
    FinLoads<D> f;
    gs_fin_load<D, LD>(L, rowbase, coef, dinv, x, w, i0, i1, i2, eq, f);
    PDEOP_LOAD_FENCE();
    gs_fin_compute<D, PITCH>(L, T, x, w, i0, i1, i2, f, r);
}

// Both halves back to back, coordinates (L.coord[w]) supplied by the caller.  (Measured and dropped for the unsplit
// kernel: issuing the finishing half's loads before the last gather axis is consumed spills -- 65 doubles live --,
// and pulling them into L2 with prefetch.global.L2 is slower on the fine level, 3.05 vs 2.83 ms per call.)
template <int D, class LD, int PITCH>
PDEOP_HD void gs_elem_cf(const LevelDev& L, const int* __restrict__ rowbase, const double* __restrict__ T,
                         const double* __restrict__ coef, const double* __restrict__ dinv,
                         const double* __restrict__ b, double* x, int w, int cf) {
    int i0, i1, i2;
    unpack_coord(cf, i0, i1, i2);
    double r[1 + 2 * D];
    gs_pre_elem<D, LD, PITCH>(L, rowbase, T, b, x, w, i0, i1, i2, r);
    gs_fin_elem<D, LD, PITCH>(L, rowbase, T, coef, dinv, x, w, i0, i1, i2, coord_eq(cf), r);
}

// (the one-launch-per-step cross-check kernel and the host emulator)
template <int D, class LD, int PITCH>
PDEOP_HD void gs_elem(const LevelDev& L, const int* __restrict__ rowbase, const double* __restrict__ T,
                      const double* __restrict__ coef, const double* __restrict__ dinv,
                      const double* __restrict__ b, double* x, int w) {
    gs_elem_cf<D, LD, PITCH>(L, rowbase, T, coef, dinv, b, x, w, L.coord[w]);
}

// ------------------------------------------------------------------------------------------------
// Linear grid transfer (align_corners=True), C channels, wave layout on both sides.
// out[m][wo] (=|+=) interp(in)[m] at the output point wo.
// ------------------------------------------------------------------------------------------------
PDEOP_HD void interp_axis(int i, int n_in, int n_out, int& lo, int& hi, double& l0, double& l1) {
    const double scale = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
    const double src = scale * (double)i;
    lo = (int)src;
    if (lo > n_in - 1) lo = n_in - 1;
    hi = lo + (lo < n_in - 1 ? 1 : 0);
    l1 = src - (double)lo;
    l0 = 1.0 - l1;
}

// CC > 0: channel count at compile time (the channel loop unrolls: all gathers of a point in flight together)
template <int CC = 0>
PDEOP_HD void interp_elem(const LevelDev& Li, const LevelDev& Lo, int C, const double* __restrict__ in,
                          double* __restrict__ out, int wo, int add) {
    if (CC > 0) C = CC;
    int i0, i1, i2;
    unpack_coord(Lo.coord[wo], i0, i1, i2);
    int l[3], h[3];
    double w0[3], w1[3];
    interp_axis(i0, Li.N[0], Lo.N[0], l[0], h[0], w0[0], w1[0]);
    interp_axis(i1, Li.N[1], Lo.N[1], l[1], h[1], w0[1], w1[1]);
    interp_axis(i2, Li.N[2], Lo.N[2], l[2], h[2], w0[2], w1[2]);
    // Corners with an exactly zero weight are not read (l1 == 0 means l0 == 1: w0*a + 0*b == w0*a bit for bit).  An
    // axis that is not coarsened (downsample_first=False keeps the time axis) always has l1 == 0: half the gathers.
    const bool z0 = w1[0] != 0.0, z1 = w1[1] != 0.0, z2 = w1[2] != 0.0;
    const int p000 = wave_pos(Li, l[0], l[1], l[2]);
    const int p001 = z2 ? wave_pos(Li, l[0], l[1], h[2]) : p000;
    const int p010 = z1 ? wave_pos(Li, l[0], h[1], l[2]) : p000;
    const int p011 = (z1 && z2) ? wave_pos(Li, l[0], h[1], h[2]) : p000;
    const int p100 = z0 ? wave_pos(Li, h[0], l[1], l[2]) : p000;
    const int p101 = (z0 && z2) ? wave_pos(Li, h[0], l[1], h[2]) : p000;
    const int p110 = (z0 && z1) ? wave_pos(Li, h[0], h[1], l[2]) : p000;
    const int p111 = (z0 && z1 && z2) ? wave_pos(Li, h[0], h[1], h[2]) : p000;
#pragma unroll
    for (int m = 0; m < (CC > 0 ? CC : C); ++m) {
        const double* __restrict__ s = in + (size_t)m * Li.G;
        double a0 = w0[2] * s[p000];
        if (z2) a0 += w1[2] * s[p001];
        double lo = w0[1] * a0;
        if (z1) {
            double a1 = w0[2] * s[p010];
            if (z2) a1 += w1[2] * s[p011];
            lo += w1[1] * a1;
        }
        double v = w0[0] * lo;
        if (z0) {
            double b0 = w0[2] * s[p100];
            if (z2) b0 += w1[2] * s[p101];
            double hi = w0[1] * b0;
            if (z1) {
                double b1 = w0[2] * s[p110];
                if (z2) b1 += w1[2] * s[p111];
                hi += w1[1] * b1;
            }
            v += w1[0] * hi;
        }
        const size_t k = (size_t)m * Lo.G + wo;
        out[k] = add ? out[k] + v : v;
    }
}

// Transpose of the linear prolongation coarse -> fine: fine point wf adds w * in[wf] to every coarse corner its
// prolongated value reads with weight w (R = P^T; the converged-mode V-cycle needs it to stay symmetric).
PDEOP_HD void restrict_t_elem(const LevelDev& Lf, const LevelDev& Lc, int C, const double* __restrict__ in,
                              double* out, int wf) {
    int i0, i1, i2;
    unpack_coord(Lf.coord[wf], i0, i1, i2);
    int l[3], h[3];
    double w0[3], w1[3];
    interp_axis(i0, Lc.N[0], Lf.N[0], l[0], h[0], w0[0], w1[0]);
    interp_axis(i1, Lc.N[1], Lf.N[1], l[1], h[1], w0[1], w1[1]);
    interp_axis(i2, Lc.N[2], Lf.N[2], l[2], h[2], w0[2], w1[2]);
    for (int c0 = 0; c0 < 2; ++c0) {
        const double a0 = c0 ? w1[0] : w0[0];
        if (a0 == 0.0) continue;
        for (int c1 = 0; c1 < 2; ++c1) {
            const double a1 = c1 ? w1[1] : w0[1];
            if (a1 == 0.0) continue;
            for (int c2 = 0; c2 < 2; ++c2) {
                const double a2 = c2 ? w1[2] : w0[2];
                if (a2 == 0.0) continue;
                const int pc = wave_pos(Lc, c0 ? h[0] : l[0], c1 ? h[1] : l[1], c2 ? h[2] : l[2]);
                const double wt = a0 * a1 * a2;
                for (int m = 0; m < C; ++m)
                    PDEOP_ATOMIC_ADD(&out[(size_t)m * Lc.G + pc], wt * in[(size_t)m * Lf.G + wf]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// layout conversion between the operator surface (B,G,M) natural order and wave/planar order
// ------------------------------------------------------------------------------------------------
PDEOP_HD void pack_elem(const LevelDev& L, const double* __restrict__ api, double* __restrict__ wave, int w) {
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    const size_t g = (size_t)nat_index(L, i0, i1, i2);
    for (int m = 0; m < L.M; ++m) wave[(size_t)m * L.G + w] = api[g * L.M + m];
}

PDEOP_HD void unpack_elem(const LevelDev& L, const double* __restrict__ wave, double* __restrict__ api, int w) {
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    const size_t g = (size_t)nat_index(L, i0, i1, i2);
    for (int m = 0; m < L.M; ++m) api[g * L.M + m] = wave[(size_t)m * L.G + w];
}

// A^T b restricted to equation rows: atb[m][w] = c[m] * rhs[g]   (initial rows added by atb_init_elem)
PDEOP_HD void atb_elem(const LevelDev& L, const double* __restrict__ coef, const double* __restrict__ rhs_nat,
                       double* __restrict__ atb, int w) {
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    const bool eq = L.flags[w] & 1;
    const double r = eq ? rhs_nat[nat_index(L, i0, i1, i2)] : 0.0;
    for (int m = 0; m < L.M; ++m) atb[(size_t)m * L.G + w] = eq ? coef[(size_t)m * L.G + w] * r : 0.0;
}

PDEOP_HD void atb_init_elem(const LevelDev& L, const double* __restrict__ iv_rhs, double* atb, int k) {
    PDEOP_ATOMIC_ADD(&atb[(size_t)L.init_m[k] * L.G + L.init_w[k]], iv_rhs[k]);
}

// ------------------------------------------------------------------------------------------------
// Dense K for one point: writes the M rows of the point into the pre-zeroed n x n matrix (n = M*G) in BAND
// ordering: unknown (w,m) -> band[w]*M + m.
// ------------------------------------------------------------------------------------------------
template <int D>
PDEOP_HD void dense_elem(const LevelDev& L, const double* __restrict__ T, const double* __restrict__ coef,
                         double* __restrict__ Kd, int w) {
    constexpr int M = 1 + 2 * D;
    const int G = L.G;
    constexpr int P = kTabPitch;
    const size_t n = (size_t)M * G;
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    const int flags = L.flags[w];
    const int s = i0 + i1 + i2;
    const int idx[3] = {i0, i1, i2};
    PointLocal<D> pl;
    load_local<D>(L, T, coef, w, i0, i1, i2, flags, pl);
    const size_t r0 = (size_t)L.band[w] * M;   // first row/col of this point
    // local block
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < M; ++k) {
            double v = pl.c[m] * pl.c[k];
            if (m == k) v += pl.ini[m];
            Kd[(r0 + m) * n + r0 + k] = v;
        }
    Kd[r0 * n + r0] += pl.uu;
    for (int a = 0; a < D; ++a) {
        const size_t ru = r0, rp = r0 + 1 + a, rq = r0 + 1 + D + a;
        Kd[ru * n + rp] += pl.up[a];
        Kd[rp * n + ru] += pl.up[a];
        Kd[ru * n + rq] += pl.uq[a];
        Kd[rq * n + ru] += pl.uq[a];
        Kd[rp * n + rp] += pl.pp[a];
        Kd[rq * n + rq] += pl.qq[a];
        Kd[rp * n + rq] += pl.pq[a];
        Kd[rq * n + rp] += pl.pq[a];
    }
    // neighbours
    for (int a = 0; a < D; ++a) {
        const int ax = 3 - D + a;
        const int nn = L.N[ax];
        const int i = idx[ax];
        const double* Ta = axis_table<kTabPitch>(T, a) + (i + kTabPad);
        const size_t ru = r0, rp = r0 + 1 + a, rq = r0 + 1 + D + a;
        for (int o = -4; o <= 4; ++o) {
            if (o == 0) continue;
            const int ii = i + o;
            if (ii < 0 || ii >= nn) continue;
            int wn;
            if (ax == 2) wn = neighbor_pos<2>(L, s, i0, i1, o);
            else if (ax == 1) wn = neighbor_pos<1>(L, s, i0, i1, o);
            else wn = neighbor_pos<0>(L, s, i0, i1, o);
            const size_t c0 = (size_t)L.band[wn] * M;
            Kd[ru * n + c0] = Ta[(T_UU + o + 4) * P];
            Kd[ru * n + c0 + 1 + a] = Ta[(T_UP - o + 4) * P + o];
            Kd[ru * n + c0 + 1 + D + a] = Ta[(T_UQ - o + 4) * P + o];
            Kd[rp * n + c0] = Ta[(T_UP + o + 4) * P];
            Kd[rq * n + c0] = Ta[(T_UQ + o + 4) * P];
        }
    }
}

// wave/planar vector [m][w] <-> band-ordered vector [band[w]*M + m]
PDEOP_HD void to_band_elem(const LevelDev& L, const double* __restrict__ wave, double* __restrict__ bandv, int w) {
    const size_t r0 = (size_t)L.band[w] * L.M;
    for (int m = 0; m < L.M; ++m) bandv[r0 + m] = wave[(size_t)m * L.G + w];
}
PDEOP_HD void from_band_elem(const LevelDev& L, const double* __restrict__ bandv, double* __restrict__ wave, int w) {
    const size_t r0 = (size_t)L.band[w] * L.M;
    for (int m = 0; m < L.M; ++m) wave[(size_t)m * L.G + w] = bandv[r0 + m];
}

// ------------------------------------------------------------------------------------------------
// Gradients at one point (qp_dual_sparse_multigrid_normal_kkt.py:112-162):
//   lam = b - A x, dnu = -A dz;  dA_r = lam_r dz^T + dnu_r x^T on the pattern of row r;  db = A dz.
// Outputs: d_coeffs (G,M) natural, d_rhs (G) natural, and atomically accumulated gradients of the
// per-line row values d_cv [Ntot][2][6], d_fv [Ftot][4], d_bv [Ftot][4] (the reference's per-nnz dD
// summed over the grid directions a line value is expanded along).
// ------------------------------------------------------------------------------------------------
template <int D>
PDEOP_HD void grad_elem(const LevelDev& L, const double* __restrict__ coef, const double* __restrict__ rhs_nat,
                        const double* __restrict__ cv, const double* __restrict__ fv, const double* __restrict__ bv,
                        const double* __restrict__ x, const double* __restrict__ dz, double* __restrict__ d_coeffs,
                        double* __restrict__ d_rhs, double* d_cv, double* d_fv, double* d_bv, int w) {
    constexpr int M = 1 + 2 * D;
    const int G = L.G;
    int i0, i1, i2;
    unpack_coord(L.coord[w], i0, i1, i2);
    const int flags = L.flags[w];
    const int s = i0 + i1 + i2;
    const int idx[3] = {i0, i1, i2};
    const size_t g = (size_t)nat_index(L, i0, i1, i2);
    double xl[M], zl[M];
    for (int m = 0; m < M; ++m) {
        xl[m] = x[(size_t)m * G + w];
        zl[m] = dz[(size_t)m * G + w];
    }
    if (flags & 1) {
        double cx = 0.0, cz = 0.0;
        for (int m = 0; m < M; ++m) {
            const double c = coef[(size_t)m * G + w];
            cx += c * xl[m];
            cz += c * zl[m];
        }
        const double lam = rhs_nat[g] - cx;
        for (int m = 0; m < M; ++m) d_coeffs[g * M + m] = lam * zl[m] - cz * xl[m];
        d_rhs[g] = cz;
    } else {
        for (int m = 0; m < M; ++m) d_coeffs[g * M + m] = 0.0;
        d_rhs[g] = 0.0;
    }
    for (int a = 0; a < D; ++a) {
        const int ax = 3 - D + a;
        const int n = L.N[ax];
        const int i = idx[ax];
        // central rows (i, k)
        int o[5];
        stencil_offsets(i, n, o);
        double ux[5], uz[5];
        for (int j = 0; j < 5; ++j) {
            int wn;
            if (o[j] == 0) wn = w;
            else if (ax == 2) wn = neighbor_pos<2>(L, s, i0, i1, o[j]);
            else if (ax == 1) wn = neighbor_pos<1>(L, s, i0, i1, o[j]);
            else wn = neighbor_pos<0>(L, s, i0, i1, o[j]);
            ux[j] = x[wn];
            uz[j] = dz[wn];
        }
        for (int k = 0; k < 2; ++k) {
            const size_t ro = ((size_t)(L.cvoff[a] + i) * 2 + k) * 6;
            const double* row = cv + ro;
            const int ch = k == 0 ? 1 + a : 1 + D + a;
            double rx = row[5] * xl[ch], rz = row[5] * zl[ch];
            for (int j = 0; j < 5; ++j) {
                rx += row[j] * ux[j];
                rz += row[j] * uz[j];
            }
            for (int j = 0; j < 5; ++j) PDEOP_ATOMIC_ADD(&d_cv[ro + j], -rx * uz[j] - rz * ux[j]);
            PDEOP_ATOMIC_ADD(&d_cv[ro + 5], -rx * zl[ch] - rz * xl[ch]);
        }
        // forward row at this point: u, u_c, u_cc, u(next)
        if (i <= n - 2) {
            int wn;
            if (ax == 2) wn = neighbor_pos<2>(L, s, i0, i1, 1);
            else if (ax == 1) wn = neighbor_pos<1>(L, s, i0, i1, 1);
            else wn = neighbor_pos<0>(L, s, i0, i1, 1);
            const size_t ro = (size_t)(L.fvoff[a] + i) * 4;
            const double* f = fv + ro;
            const double vx[4] = {xl[0], xl[1 + a], xl[1 + D + a], x[wn]};
            const double vz[4] = {zl[0], zl[1 + a], zl[1 + D + a], dz[wn]};
            double rx = 0.0, rz = 0.0;
            for (int j = 0; j < 4; ++j) {
                rx += f[j] * vx[j];
                rz += f[j] * vz[j];
            }
            for (int j = 0; j < 4; ++j) PDEOP_ATOMIC_ADD(&d_fv[ro + j], -rx * vz[j] - rz * vx[j]);
        }
        // backward row at this point: u, u_c, u_cc, u(prev); table entry i-1
        if (i >= 1) {
            int wn;
            if (ax == 2) wn = neighbor_pos<2>(L, s, i0, i1, -1);
            else if (ax == 1) wn = neighbor_pos<1>(L, s, i0, i1, -1);
            else wn = neighbor_pos<0>(L, s, i0, i1, -1);
            const size_t ro = (size_t)(L.fvoff[a] + i - 1) * 4;
            const double* f = bv + ro;
            const double vx[4] = {xl[0], xl[1 + a], xl[1 + D + a], x[wn]};
            const double vz[4] = {zl[0], zl[1 + a], zl[1 + D + a], dz[wn]};
            double rx = 0.0, rz = 0.0;
            for (int j = 0; j < 4; ++j) {
                rx += f[j] * vx[j];
                rz += f[j] * vz[j];
            }
            for (int j = 0; j < 4; ++j) PDEOP_ATOMIC_ADD(&d_bv[ro + j], -rx * vz[j] - rz * vx[j]);
        }
    }
}

// d(iv_rhs)[k] = (A dz)_init,k = dz at the row's variable
PDEOP_HD void grad_init_elem(const LevelDev& L, const double* __restrict__ dz, double* __restrict__ d_iv, int k) {
    d_iv[k] = dz[(size_t)L.init_m[k] * L.G + L.init_w[k]];
}

}  // namespace pdeop
