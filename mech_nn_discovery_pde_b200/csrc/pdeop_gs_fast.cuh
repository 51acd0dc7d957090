// pdeop -- wavefront Gauss-Seidel, kernel for the LATENCY-BOUND levels (device only; included by pdeop_cuda.cu).
//
// Same schedule and the same canonical arithmetic as k_gs_cluster / gs_elem_cf (pdeop_elem.h) -- one cluster per
// instance, sweep k on hyperplane t - 5k, cluster barrier per step, bit-identical results (tested against the
// one-launch-per-step kernel).  On a level whose hyperplane steps hold about one point per thread, the time of a
// step is one thread's dependency chain: coordinates -> neighbour indices -> two or three dependent L2 round trips
// of gathers -> DRAM/L2-cold own operands -> the sequential channel solve.  This kernel shortens that chain
// (measured on B200, batch 32, 5 sweeps: 32x32x32 0.781 vs 0.826 ms, 32x16x16 0.333 vs 0.380 ms; on the
// throughput-bound fine level it is SLOWER than k_gs_cluster, 2.6 vs 2.3 ms, see profiles/README.md):
//   * the instance's K tables are re-laid in shared memory as one 34-double RECORD per (axis, line position):
//     everything a point needs from an axis sits in 15 aligned 16-byte loads (LDS.128; stride 34 doubles = 4 banks
//     mod 32, conflict-free per quarter warp) instead of ~45 8-byte loads with per-entry address arithmetic; the
//     couplings that only exist next to the ends of a line (one-sided stencils) live in a second, rarely read table;
//   * the operands a point reads exactly once per sweep -- b, coefficients, reciprocal diagonal, its own old values:
//     3/4 of the algorithmic bytes, all DRAM-cold -- are staged by cp.async (LDGSTS) into a private shared-memory
//     slot ONE POINT AHEAD (the first point of the next step is requested before the cluster barrier), so their DRAM
//     latency overlaps the previous point's channel solve, the barrier and this point's gather instead of sitting
//     in its dependency chain;
//   * 384 threads per CTA at 168 registers: all 48 neighbour values of a point are gathered in ONE round trip
//     (k_gs_cluster keeps two of three axes in flight at 128 registers);
//   * vector planes are addressed as [opaque 64-bit plane base + 32-bit index]: one IMAD.WIDE per load; the
//     coordinates of a thread's next point are loaded one point ahead.
// Reference semantics: solver/multigrid.py:399-405 (x <- tril(K)^-1 (b - triu(K,1) x), lexicographic).
#pragma once

namespace pdeop {

constexpr int kRec = 34;        // doubles per (axis, position) record
constexpr int kRare = 18;       // doubles per (axis, position) rare record (16 used)
// record layout: e = 0..3 <-> o = -2,-1,1,2 and far offset f(e) = -4,-3,3,4
//   [6e+0] UU(o)  [6e+1] UPn(o)  [6e+2] UQn(o)  [6e+3] UP(o)  [6e+4] UQ(o)  [6e+5] UU(f(e))
//   [24] UU(0) [25] UP(0) [26] UQ(0) [27] PP [28] QQ [29] PQ
// rare layout: [2e+0] UP(f(e)) [2e+1] UQ(f(e))  [8+2e+0] UPn(f(e)) [8+2e+1] UQn(f(e))
__device__ __forceinline__ int rec_near_off(int e) { return e < 2 ? e - 2 : e - 1; }
__device__ __forceinline__ int rec_far_off(int e) { return e < 2 ? e - 4 : e + 1; }

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int D>
struct GsFastSmem {
    // [records D*P*kRec | rare D*P*kRare | slots 4*M*THREADS | rowbase ints | hstart ints]
    static __host__ __device__ size_t bytes(int P, int threads, int nrb, int S, bool stage) {
        constexpr int M = 1 + 2 * D;
        return ((size_t)D * P * (kRec + kRare) + (stage ? (size_t)4 * M * threads : 0)) * sizeof(double) +
               ((size_t)nrb + S + 1 + 3) / 4 * 4 * sizeof(int);
    }
};

// Re-lay one instance's K tables ([entry][position], global memory) as records in shared memory.
template <int D>
__device__ __forceinline__ void stage_records(const LevelDev& L, const double* __restrict__ Ti, int P, double* rec,
                                              double* rare, int tid, int nthr) {
    for (int i = tid; i < D * P; i += nthr) {
        const int a = i / P, pos = i - a * P;
        const double* Ta = Ti + (size_t)a * kTabEntries * kTabPitch + (pos + kTabPad);
        double* R = rec + (size_t)i * kRec;
        double* Q = rare + (size_t)i * kRare;
        const bool in = pos < L.N[3 - D + a];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int on = rec_near_off(e), of = rec_far_off(e);
            R[6 * e + 0] = in ? Ta[(T_UU + on + 4) * kTabPitch] : 0.0;
            R[6 * e + 1] = in ? Ta[(T_UP - on + 4) * kTabPitch + on] : 0.0;
            R[6 * e + 2] = in ? Ta[(T_UQ - on + 4) * kTabPitch + on] : 0.0;
            R[6 * e + 3] = in ? Ta[(T_UP + on + 4) * kTabPitch] : 0.0;
            R[6 * e + 4] = in ? Ta[(T_UQ + on + 4) * kTabPitch] : 0.0;
            R[6 * e + 5] = in ? Ta[(T_UU + of + 4) * kTabPitch] : 0.0;
            Q[2 * e + 0] = in ? Ta[(T_UP + of + 4) * kTabPitch] : 0.0;
            Q[2 * e + 1] = in ? Ta[(T_UQ + of + 4) * kTabPitch] : 0.0;
            Q[8 + 2 * e + 0] = in ? Ta[(T_UP - of + 4) * kTabPitch + of] : 0.0;
            Q[8 + 2 * e + 1] = in ? Ta[(T_UQ - of + 4) * kTabPitch + of] : 0.0;
        }
        R[24] = in ? Ta[(T_UU + 4) * kTabPitch] : 0.0;
        R[25] = in ? Ta[(T_UP + 4) * kTabPitch] : 0.0;
        R[26] = in ? Ta[(T_UQ + 4) * kTabPitch] : 0.0;
        R[27] = in ? Ta[T_PP * kTabPitch] : 0.0;
        R[28] = in ? Ta[T_QQ * kTabPitch] : 0.0;
        R[29] = in ? Ta[T_PQ * kTabPitch] : 0.0;
        R[30] = R[31] = R[32] = R[33] = 0.0;
        Q[16] = Q[17] = 0.0;
    }
}

// One point.  Arithmetic identical to gs_elem_cf<D, ., .>: see the canonical-order comments in pdeop_elem.h.
template <int D, int THREADS, bool STAGE, bool ALLAX>
__device__ __forceinline__ void gs_fast_point(const LevelDev& L, const double* __restrict__ rec,
                                              const double* __restrict__ rare, int P, const int* __restrict__ rbs,
                                              const double* const (&xp)[1 + 2 * D], double* const (&xw)[1 + 2 * D],
                                              const double* slot, const double* bo, const double* co,
                                              const double* dvo, unsigned w, int cf) {
    constexpr int M = 1 + 2 * D;
    int i0, i1, i2;
    unpack_coord(cf, i0, i1, i2);
    const int idx[3] = {i0, i1, i2};
    const unsigned gm1 = (unsigned)L.G - 1u;
    const int N0 = L.N[0];
    const int* __restrict__ rb = rbs + (i0 + i1 + i2 + 4) * N0 + i0;
    double au[D], ap[D], aq[D];
    double b1u[D], b1p[D], b1q[D];    // backward distance-1 neighbours (applied after r = b - acc)
    const double2* Rp[D];
    // ---- loads: ALLAX: every neighbour load of the point in flight together (one memory round trip, 48 values in
    // registers); otherwise the loads of axis a+1 are issued before the FMAs of axis a (two axes in flight) ----
    unsigned wn[D][8];
    double un[D][8], pn[D][4], qn[D][4];
    auto issue_axis = [&](int a) {
        const int ax = 3 - D + a;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = j < 4 ? j - 4 : j - 3;
            wn[a][j] = clamp_wave(neighbor_wave(rb, ax, N0, i1, o), gm1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) un[a][j] = xp[0][wn[a][j]];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            pn[a][e] = xp[1 + a][wn[a][e + 2]];
            qn[a][e] = xp[1 + D + a][wn[a][e + 2]];
        }
    };
    if (ALLAX) {
#pragma unroll
        for (int a = 0; a < D; ++a) issue_axis(a);
        PDEOP_LOAD_FENCE();
    } else {
        issue_axis(0);
    }
    // ---- per axis, the canonical FMA order ----
#pragma unroll
    for (int a = 0; a < D; ++a) {
        if (!ALLAX) {
            if (a + 1 < D) issue_axis(a + 1);
            PDEOP_LOAD_FENCE();
        }
        const int ax = 3 - D + a;
        const int i = idx[ax], n = L.N[ax];
        const double2* __restrict__ R = reinterpret_cast<const double2*>(rec + (size_t)(a * P + i) * kRec);
        Rp[a] = R;
        double u_ = 0.0, p_ = 0.0, q_ = 0.0;
        double farc[4];
        // near offsets o = -2, 1, 2 (o = -1 is the backward distance-1 neighbour, applied later)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double2 v0 = R[3 * e], v1 = R[3 * e + 1], v2 = R[3 * e + 2];
            farc[e] = v2.y;
            if (e == 1) continue;
            u_ = fma(v0.x, un[a][e + 2], u_);
            u_ = fma(v0.y, pn[a][e], u_);
            u_ = fma(v1.x, qn[a][e], u_);
            p_ = fma(v1.y, un[a][e + 2], p_);
            q_ = fma(v2.x, un[a][e + 2], q_);
        }
        // far u terms o = -4, -3, 3, 4
#pragma unroll
        for (int e = 0; e < 4; ++e) u_ = fma(farc[e], un[a][e < 2 ? e : e + 4], u_);
        const bool e1 = axis_end1(i, n), e3 = axis_end3(i, n);
        if (e1 || e3) {
            const double2* __restrict__ Q = reinterpret_cast<const double2*>(rare + (size_t)(a * P + i) * kRare);
            if (e1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double2 v = Q[e];
                    p_ = fma(v.x, un[a][e < 2 ? e : e + 4], p_);
                    q_ = fma(v.y, un[a][e < 2 ? e : e + 4], q_);
                }
            }
            if (e3) {
                double pf[4], qf[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned wf = wn[a][e < 2 ? e : e + 4];
                    pf[e] = xp[1 + a][wf];
                    qf[e] = xp[1 + D + a][wf];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double2 v = Q[4 + e];
                    u_ = fma(v.x, pf[e], u_);
                    u_ = fma(v.y, qf[e], u_);
                }
            }
        }
        au[a] = u_;
        ap[a] = p_;
        aq[a] = q_;
        b1u[a] = un[a][3];
        b1p[a] = pn[a][1];
        b1q[a] = qn[a][1];
    }
    // r = b - (sum over axes), accumulated in axis order from 0.0 like k_gather
    double r[M], xl[M], c[M], di[M];
    const bool eq = coord_eq(cf);
    const unsigned G = (unsigned)L.G;
    if (STAGE) {
        cp_async_wait_all();
#pragma unroll
        for (int m = 0; m < M; ++m) {
            r[m] = slot[m * THREADS];
            xl[m] = slot[(M + m) * THREADS];
            c[m] = eq ? slot[(2 * M + m) * THREADS] : 0.0;
            di[m] = slot[(3 * M + m) * THREADS];
        }
    } else {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            r[m] = ld_stream(bo + ((unsigned)m * G + w));
            xl[m] = xp[m][w];
            c[m] = eq ? ld_stream(co + ((unsigned)m * G + w)) : 0.0;
            di[m] = ld_stream(dvo + ((unsigned)m * G + w));
        }
        PDEOP_LOAD_FENCE();
    }
    {
        double acc0 = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) acc0 += au[a];
        r[0] = r[0] - acc0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
            r[1 + a] = r[1 + a] - ap[a];
            r[1 + D + a] = r[1 + D + a] - aq[a];
        }
    }
    // backward distance-1 couplings (gs_back1_apply): record entries [6..10] = UU, UPn, UQn, UP, UQ at o = -1;
    // same-point entries [24..29] = uu, up, uq, pp, qq, pq
    double up[D], uq[D], pq[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const double2 v0 = Rp[a][3], v1 = Rp[a][4], v2 = Rp[a][5];
        r[0] = fma(-v0.x, b1u[a], r[0]);
        r[0] = fma(-v0.y, b1p[a], r[0]);
        r[0] = fma(-v1.x, b1q[a], r[0]);
        r[1 + a] = fma(-v1.y, b1u[a], r[1 + a]);
        r[1 + D + a] = fma(-v2.x, b1u[a], r[1 + D + a]);
        const double2 l0 = Rp[a][12], l2 = Rp[a][14];
        up[a] = l0.y;
        uq[a] = Rp[a][13].x;
        pq[a] = l2.y;
    }
    // channel solve (gs_fin_compute)
    double S = c[0] * xl[0];
#pragma unroll
    for (int a = 0; a < D; ++a) S = S + fma(c[1 + D + a], xl[1 + D + a], c[1 + a] * xl[1 + a]);
    {
        const double t = fma(-c[0], xl[0], S);
        double off = c[0] * t;
#pragma unroll
        for (int a = 0; a < D; ++a) off = off + fma(uq[a], xl[1 + D + a], up[a] * xl[1 + a]);
        const double xn = (r[0] - off) * di[0];
        S = fma(c[0], xn, t);
        xl[0] = xn;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const int m = 1 + a;
        const double t = fma(-c[m], xl[m], S);
        const double off = fma(pq[a], xl[1 + D + a], fma(up[a], xl[0], c[m] * t));
        const double xn = (r[m] - off) * di[m];
        S = fma(c[m], xn, t);
        xl[m] = xn;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const int m = 1 + D + a;
        const double t = fma(-c[m], xl[m], S);
        const double off = fma(pq[a], xl[1 + a], fma(uq[a], xl[0], c[m] * t));
        const double xn = (r[m] - off) * di[m];
        S = fma(c[m], xn, t);
        xl[m] = xn;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) xw[m][w] = xl[m];
}

template <int D, int THREADS, bool SINGLE, bool STAGE = true, bool ALLAX = (THREADS <= 384)>
__global__ void __launch_bounds__(THREADS, 1) k_gs_fast(LevelDev L, const double* __restrict__ T,
                                                        const double* __restrict__ coef,
                                                        const double* __restrict__ dinv,
                                                        const double* __restrict__ b, double* x, int nsweeps, int P,
                                                        const int* done) {
    if (done && *done) return;
    constexpr int M = 1 + 2 * D;
    extern __shared__ __align__(16) unsigned char gs_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = SINGLE ? 1 : (int)cluster.num_blocks();
    const int rank = SINGLE ? 0 : (int)cluster.block_rank();
    const int ib = blockIdx.x / csize;
    const int tid = (((int)threadIdx.x >> 5) * csize + rank) * 32 + ((int)threadIdx.x & 31);
    const int nthreads = csize * THREADS;
    const size_t o = (size_t)ib * M * L.G;
    const double* Ti = T + (size_t)ib * D * kTabEntries * kTabPitch;
    double* rec = reinterpret_cast<double*>(gs_smem);
    double* rare = rec + (size_t)D * P * kRec;
    double* slots = rare + (size_t)D * P * kRare;
    int* rbs = reinterpret_cast<int*>(slots + (STAGE ? (size_t)4 * M * THREADS : 0));
    const int nrb = (L.S + 8) * L.N[0] + 8;
    int* hss = rbs + nrb;
    // ---- stage: records, rare records, index tables ----
    stage_records<D>(L, Ti, P, rec, rare, threadIdx.x, THREADS);
    for (int i = threadIdx.x; i < nrb; i += THREADS) rbs[i] = L.rowbase[i - 4];
    for (int i = threadIdx.x; i <= L.S; i += THREADS) hss[i] = L.hstart[i];
    __syncthreads();
    const int* rbuse = rbs + 4;
    const int* hs = hss;

    // opaque plane base pointers: every vector access is [64-bit plane base + 32-bit index]
    const double* xp[M];
    double* xw[M];
    const double *bo = b + o, *co = coef + o, *dvo = dinv + o;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double* p = x + o + (size_t)m * L.G;
        asm volatile("" : "+l"(p));
        __builtin_assume(__isGlobal(p));
        xw[m] = p;
        xp[m] = p;
    }
    asm volatile("" : "+l"(bo), "+l"(co), "+l"(dvo));
    __builtin_assume(__isGlobal(bo));
    __builtin_assume(__isGlobal(co));
    __builtin_assume(__isGlobal(dvo));
    const unsigned G = (unsigned)L.G;
    double* slot = slots + threadIdx.x;

    const int lag = kGsLagUnsplit;
    const int steps = L.S + lag * (nsweeps - 1);
    auto step_sweeps = [&](int t, int& k_lo, int& k_hi) -> int {
        k_lo = (t - (L.S - 1) + lag - 1) / lag;
        if (k_lo < 0) k_lo = 0;
        k_hi = t / lag;
        if (k_hi > nsweeps - 1) k_hi = nsweeps - 1;
        int total = 0;
        for (int k = k_lo; k <= k_hi; ++k) {
            const int s = t - lag * k;
            total += hs[s + 1] - hs[s];
        }
        return total;
    };
    auto point_of = [&](int t, int k_lo, int k_hi, int idx) -> int {
        int rem = idx;
        for (int k = k_lo; k <= k_hi; ++k) {
            const int s = t - lag * k;
            const int h0 = hs[s], cnt = hs[s + 1] - h0;
            if (rem < cnt) return h0 + rem;
            rem -= cnt;
        }
        return -1;
    };
    // request the single-use operands of point w into this thread's slot: [b | x_old | coef | dinv] x M
    auto prefetch = [&](unsigned w) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const unsigned k = (unsigned)m * G + w;
            cp_async8(slot + m * THREADS, bo + k);
            cp_async8(slot + (M + m) * THREADS, xp[m] + w);
            cp_async8(slot + (2 * M + m) * THREADS, co + k);
            cp_async8(slot + (3 * M + m) * THREADS, dvo + k);
        }
        cp_async_commit();
    };
    int k_lo = 0, k_hi = 0;
    int total = step_sweeps(0, k_lo, k_hi);
    int w = tid < total ? point_of(0, k_lo, k_hi, tid) : -1;
    int cf = w >= 0 ? L.coord[w] : 0;
    if (STAGE && w >= 0) prefetch((unsigned)w);
    for (int t = 0; t < steps; ++t) {
        // enumeration of the next step: its first point is requested before this step's barrier
        int nk_lo = 0, nk_hi = 0;
        const int ntotal = t + 1 < steps ? step_sweeps(t + 1, nk_lo, nk_hi) : 0;
        int wN = -1, cfN = 0;
        if (w < 0) {
            wN = tid < ntotal ? point_of(t + 1, nk_lo, nk_hi, tid) : -1;
            cfN = wN >= 0 ? L.coord[wN] : 0;
            if (STAGE && wN >= 0) prefetch((unsigned)wN);
        }
        int idx = tid;
        while (w >= 0) {
            // this thread's next point: in this step, else the first one of the next step
            const int idx_n = idx + nthreads;
            const bool more = idx_n < total;
            int w_n;
            if (more) w_n = point_of(t, k_lo, k_hi, idx_n);
            else w_n = tid < ntotal ? point_of(t + 1, nk_lo, nk_hi, tid) : -1;
            const int cf_n = w_n >= 0 ? L.coord[w_n] : 0;
            gs_fast_point<D, THREADS, STAGE, ALLAX>(L, rec, rare, P, rbuse, xp, xw, slot, bo, co, dvo, (unsigned)w, cf);
            // the slot has been consumed: request the next point's operands
            if (STAGE && w_n >= 0) prefetch((unsigned)w_n);
            if (!more) {
                wN = w_n;
                cfN = cf_n;
                break;
            }
            w = w_n;
            cf = cf_n;
            idx = idx_n;
        }
        // release/acquire at cluster scope; the acquire side invalidates L1 (CCTL.IVALL)
        if (SINGLE) __syncthreads();
        else cluster.sync();
        w = wN;
        cf = cfN;
        k_lo = nk_lo;
        k_hi = nk_hi;
        total = ntotal;
    }
    cp_async_wait_all();
}

}  // namespace pdeop
