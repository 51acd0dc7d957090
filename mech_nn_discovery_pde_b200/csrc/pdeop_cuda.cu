// pdeop -- CUDA backend for sm_100a (B200).  Implements pdeop_backend.h.
//
// Kernel families (all fp64, HBM-bound streams unless noted; see DESIGN.md for the rooflines):
//   * stencil kernels over grid points in wave order: K apply / residual, wavefront Gauss-Seidel,
//     linear grid transfer, A^T b, dense coarse K, gradients               (bodies in pdeop_elem.h)
//   * Krylov vector kernels with fused multi-dot / multi-axpy / norm steps: warp-shuffle + block
//     reductions, block partials re-summed in a fixed order by the consumer kernel (deterministic,
//     no host synchronisation; the convergence flag lives in device memory)
//   * batched dense band Cholesky for the coarsest multigrid level and the dense layer (outer panels of 256
//     columns, left-looking inside a panel, trailing updates on fp64 DMMA), and its triangular solves: the chain
//     solver (one persistent cluster kernel per direction: TMA bulk copies, mbarrier stages, st.async exchange
//     through distributed shared memory) or, for odd / small n, one launch per block row
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include "pdeop_backend.h"
#include "pdeop_elem.h"
#include "pdeop_lstsq.h"

namespace cg = cooperative_groups;

#ifndef PDEOP_APPLY_MINB
#define PDEOP_APPLY_MINB 2
#endif
// load policy of the Gauss-Seidel iterate.  An L2 evict-last hint on these loads (createpolicy + ld.L2::cache_hint)
// was measured and changes nothing once the streamed operands carry evict-first (ld_stream).
#ifndef PDEOP_GS_LD
#define PDEOP_GS_LD LdPlain
#endif

namespace pdeop {

static thread_local cudaError_t g_cuda_err = cudaSuccess;
static inline void note(cudaError_t e) {
    if (e != cudaSuccess && g_cuda_err == cudaSuccess) g_cuda_err = e;
}
// statistics only (bench.py reports it): kernels launched by this library, all plans and threads
static std::atomic<long long> g_launches{0};
#define PDEOP_LAUNCH_CHECK() note(cudaGetLastError())
#define PDEOP_COUNT(k) (g_launches.fetch_add((k), std::memory_order_relaxed))
long long be_launch_count() { return g_launches.load(std::memory_order_relaxed); }

int be_current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return dev;
}

// Per-device caches (a process may drive several GPUs, one plan each): SM count, the largest dynamic shared-memory
// size already granted to a kernel, and cluster-occupancy answers.  Keyed by (device, kernel); guarded by a mutex
// because plans on different devices may be driven from different host threads.
static std::mutex g_cache_mu;
static std::map<std::pair<int, const void*>, size_t> g_smem_granted;
static void ensure_dyn_smem(const void* kern, size_t bytes) {
    const int dev = be_current_device();
    std::lock_guard<std::mutex> lk(g_cache_mu);
    size_t& cur = g_smem_granted[{dev, kern}];
    if (bytes > cur) {
        note(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
}

void* be_event_create() {
    cudaEvent_t e = nullptr;
    note(cudaEventCreate(&e));
    return (void*)e;
}
void be_event_destroy(void* ev) {
    if (ev) cudaEventDestroy((cudaEvent_t)ev);
}
void be_event_record(void* ev, stream_t st) { note(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st)); }
float be_event_elapsed_ms(void* start, void* stop) {
    float ms = 0.f;
    note(cudaEventSynchronize((cudaEvent_t)stop));
    note(cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return ms;
}

const char* be_name() { return "cuda-sm100a"; }

void* be_alloc(size_t bytes) {
    void* p = nullptr;
    note(cudaMalloc(&p, bytes ? bytes : 1));
    return p;
}
void be_free(void* p) {
    if (p) cudaFree(p);
}
void be_upload(void* dst, const void* src, size_t bytes) {
    if (bytes) note(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
}
void be_zero(stream_t st, void* p, size_t bytes) {
    if (bytes) note(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)st));
}
int be_last_error(char* buf, int len) {
    cudaError_t e = g_cuda_err;
    if (e == cudaSuccess) e = cudaPeekAtLastError();
    if (e == cudaSuccess) return 0;
    snprintf(buf, len, "%s", cudaGetErrorString(e));
    g_cuda_err = cudaSuccess;
    cudaGetLastError();
    return 1;
}

constexpr int kThreads = 256;
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// =================================================================================================
// element-wise stencil kernels: one thread per grid point (wave order => coalesced own-point access)
// =================================================================================================
__global__ void __launch_bounds__(kThreads) k_build_tables(LevelDev L, const double* __restrict__ cv,
                                                           const double* __restrict__ fv,
                                                           const double* __restrict__ bv, double* __restrict__ T) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y, b = blockIdx.z;
    if (ip >= L.P) return;
    build_table_elem(L, a, ip, cv + (size_t)b * L.Ntot * 12, fv + (size_t)b * L.Ftot * 4, bv + (size_t)b * L.Ftot * 4,
                     T + ((size_t)b * L.D + a) * kTabEntries * kTabPitch);
}

void be_build_tables(stream_t st, const LevelDev& L, int B, const double* cv, const double* fv, const double* bv,
                     double* T) {
    dim3 grid(cdiv(L.P, kThreads), L.D, B);
    k_build_tables<<<grid, kThreads, 0, (cudaStream_t)st>>>(L, cv, fv, bv, T);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(kThreads) k_pack(LevelDev L, const double* __restrict__ api,
                                                   double* __restrict__ wave) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    pack_elem(L, api + o, wave + o, w);
}
__global__ void __launch_bounds__(kThreads) k_unpack(LevelDev L, const double* __restrict__ wave,
                                                     double* __restrict__ api) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    unpack_elem(L, wave + o, api + o, w);
}
void be_pack(stream_t st, const LevelDev& L, int B, const double* api, double* wave) {
    k_pack<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, (cudaStream_t)st>>>(L, api, wave);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}
void be_unpack(stream_t st, const LevelDev& L, int B, const double* wave, double* api) {
    k_unpack<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, (cudaStream_t)st>>>(L, wave, api);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

template <int CC>
__global__ void __launch_bounds__(kThreads) k_interp(LevelDev Li, LevelDev Lo, int C, const double* __restrict__ in,
                                                     double* __restrict__ out, int add, const int* done) {
    if (done && *done) return;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= Lo.G) return;
    interp_elem<CC>(Li, Lo, C, in + (size_t)blockIdx.y * C * Li.G, out + (size_t)blockIdx.y * C * Lo.G, w, add);
}
void be_interp(stream_t st, const LevelDev& Li, const LevelDev& Lo, int B, int C, const double* in, double* out,
               int add, const int* done) {
    const dim3 grid(cdiv(Lo.G, kThreads), B);
    cudaStream_t cs = (cudaStream_t)st;
    if (C == 7) k_interp<7><<<grid, kThreads, 0, cs>>>(Li, Lo, C, in, out, add, done);
    else if (C == 5) k_interp<5><<<grid, kThreads, 0, cs>>>(Li, Lo, C, in, out, add, done);
    else if (C == 3) k_interp<3><<<grid, kThreads, 0, cs>>>(Li, Lo, C, in, out, add, done);
    else k_interp<0><<<grid, kThreads, 0, cs>>>(Li, Lo, C, in, out, add, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(kThreads) k_atb(LevelDev L, const double* __restrict__ coef,
                                                  const double* __restrict__ rhs_nat, double* __restrict__ atb) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    atb_elem(L, coef + o, rhs_nat + (size_t)blockIdx.y * L.G, atb + o, w);
}
__global__ void __launch_bounds__(kThreads) k_atb_init(LevelDev L, const double* __restrict__ iv_rhs, double* atb) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.n_init) return;
    atb_init_elem(L, iv_rhs + (size_t)blockIdx.y * L.n_init, atb + (size_t)blockIdx.y * L.M * L.G, k);
}
void be_atb(stream_t st, const LevelDev& L, int B, const double* coef, const double* rhs_nat, const double* iv_rhs,
            double* atb) {
    k_atb<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, (cudaStream_t)st>>>(L, coef, rhs_nat, atb);
    PDEOP_COUNT(1 + (L.n_init > 0));
    PDEOP_LAUNCH_CHECK();
    if (L.n_init > 0) {
        k_atb_init<<<dim3(cdiv(L.n_init, kThreads), B), kThreads, 0, (cudaStream_t)st>>>(L, iv_rhs, atb);
        PDEOP_LAUNCH_CHECK();
    }
}

// (3 CTAs per SM -- __launch_bounds__(kThreads, 3), 80 registers -- spills 400 bytes per thread: not used)
template <int D>
__global__ void __launch_bounds__(kThreads, PDEOP_APPLY_MINB) k_apply(LevelDev L, const double* __restrict__ T,
                                                    const double* __restrict__ coef, const double* __restrict__ x,
                                                    const double* __restrict__ b, double* __restrict__ y, int mode,
                                                    const int* done) {
    if (done && *done) return;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    apply_k_elem<D>(L, T + (size_t)blockIdx.y * L.D * kTabEntries * kTabPitch, coef + o, x + o, b ? b + o : nullptr, y + o, w,
                    mode);
}
void be_apply_k(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, const double* x,
                const double* b, double* y, int mode, const int* done) {
    dim3 grid(cdiv(L.G, kThreads), B);
    cudaStream_t s = (cudaStream_t)st;
    if (L.D == 1) k_apply<1><<<grid, kThreads, 0, s>>>(L, T, coef, x, b, y, mode, done);
    else if (L.D == 2) k_apply<2><<<grid, kThreads, 0, s>>>(L, T, coef, x, b, y, mode, done);
    else k_apply<3><<<grid, kThreads, 0, s>>>(L, T, coef, x, b, y, mode, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

// =================================================================================================
// Wavefront Gauss-Seidel.  Lexicographic GS on K couples a point only with points at distance <= 4
// along the grid axes, so all points of a hyperplane s = i0+i1+i2 are independent, and sweep k+1 may
// work on hyperplane s-5 while sweep k works on s (kGsLag).  One thread-block cluster owns one
// instance: per step it updates the points of up to `nsweeps` hyperplanes (one per in-flight sweep),
// then the cluster barrier (release/acquire) orders the in-place iterate for the next step.
// The iterate is read at the L2 coherence point (LdL2) because other CTAs of the cluster write it.
// =================================================================================================
#ifdef PDEOP_GS_TIMING
// debug build only (tools/): [0] steps, [1] cycles in point updates, [2] cycles in the barrier, [3] point updates
__device__ unsigned long long g_gs_dbg[8];
extern "C" void pdeop_gs_dbg_read(unsigned long long* out) {
    cudaMemcpyFromSymbol(out, g_gs_dbg, sizeof(g_gs_dbg));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_gs_dbg, z, sizeof(z));
}
#endif

// PS > 0: the instance's K tables are staged in shared memory with compile-time pitch PS (max extent + 8 <= PS),
// together with the level's rowbase/hstart index tables: the cluster barrier invalidates L1 every step, so
// anything read from global memory is re-fetched from L2 each step, while shared memory stays put.
// PS == 0: tables too large for shared memory, read from global memory.
template <int D, int THREADS, int MINB, class LD, int PS, bool SINGLE>
__global__ void __launch_bounds__(THREADS, MINB) k_gs_cluster(LevelDev L, const double* __restrict__ T,
                                                              const double* __restrict__ coef,
                                                              const double* __restrict__ dinv,
                                                              const double* __restrict__ b, double* x, int nsweeps,
                                                              const int* done) {
    if (done && *done) return;
    extern __shared__ __align__(16) unsigned char gs_smem[];
    // SINGLE: one CTA per instance (small levels): block barrier, no L1 invalidation between steps
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = SINGLE ? 1 : (int)cluster.num_blocks();
    const int rank = SINGLE ? 0 : (int)cluster.block_rank();
    const int ib = blockIdx.x / csize;
    // Point idx of a step is handled by "thread" tid = idx mod nthreads.  Warps are dealt round-robin over the
    // cluster's CTAs (warp-sized chunks keep every access coalesced): a step with fewer points than the cluster has
    // threads -- most steps on the coarser levels -- spreads over all SMs of the cluster instead of filling CTA 0
    // first, so each SM's load/store pipe sees 1/csize of the step's gathers.
    const int tid = (((int)threadIdx.x >> 5) * csize + rank) * 32 + ((int)threadIdx.x & 31);
    const int nthreads = csize * blockDim.x;
    const size_t o = (size_t)ib * L.M * L.G;
    const double* Ti = T + (size_t)ib * L.D * kTabEntries * kTabPitch;
    constexpr int PITCH = PS > 0 ? PS : kTabPitch;
    const double* Tuse = Ti;
    const int* rbuse = L.rowbase;
    const int* hs = L.hstart;
    if (PS > 0) {
        double* Ts = reinterpret_cast<double*>(gs_smem);
        int* rbs = reinterpret_cast<int*>(Ts + D * kTabEntries * PS);
        const int nrb = (L.S + 8) * L.N[0] + 8;
        int* hss = rbs + nrb;
        int maxn = L.N[0] > L.N[1] ? L.N[0] : L.N[1];
        maxn = (maxn > L.N[2] ? maxn : L.N[2]) + 2 * kTabPad;
        constexpr int PSD = PS > 0 ? PS : 1;
        for (int i = threadIdx.x; i < D * kTabEntries * PS; i += blockDim.x) {
            const int row = i / PSD, pos = i - row * PSD;
            Ts[i] = pos < maxn ? Ti[(size_t)row * kTabPitch + pos] : 0.0;
        }
        for (int i = threadIdx.x; i < nrb; i += blockDim.x) rbs[i] = L.rowbase[i - 4];
        for (int i = threadIdx.x; i <= L.S; i += blockDim.x) hss[i] = L.hstart[i];
        __syncthreads();
        Tuse = Ts;
        rbuse = rbs + 4;
        hs = hss;
    }
    const int steps = L.S + kGsLagUnsplit * (nsweeps - 1);
    // points of step t: the hyperplanes s_k = t - lag*k of the sweeps k in flight, concatenated
    auto step_sweeps = [&](int t, int& k_lo, int& k_hi) -> int {
        k_lo = (t - (L.S - 1) + kGsLagUnsplit - 1) / kGsLagUnsplit;
        if (k_lo < 0) k_lo = 0;
        k_hi = t / kGsLagUnsplit;
        if (k_hi > nsweeps - 1) k_hi = nsweeps - 1;
        int total = 0;
        for (int k = k_lo; k <= k_hi; ++k) {
            const int s = t - kGsLagUnsplit * k;
            total += hs[s + 1] - hs[s];
        }
        return total;
    };
    auto point_of = [&](int t, int k_lo, int k_hi, int idx) -> int {
        int rem = idx;
        for (int k = k_lo; k <= k_hi; ++k) {
            const int s = t - kGsLagUnsplit * k;
            const int h0 = hs[s], cnt = hs[s + 1] - h0;
            if (rem < cnt) return h0 + rem;
            rem -= cnt;
        }
        return -1;
    };
    // (Prefetching the next point's coordinates, as the pipelined kernel does, was measured here and is slower:
    // 2.90 vs 2.83 ms per fine-level call.)
    for (int t = 0; t < steps; ++t) {
        int k_lo, k_hi;
        const int total = step_sweeps(t, k_lo, k_hi);
        for (int idx = tid; idx < total; idx += nthreads) {
            const int w = point_of(t, k_lo, k_hi, idx);
            gs_elem<D, LD, PITCH>(L, rbuse, Tuse, coef + o, dinv + o, b + o, x + o, w);
        }
        // release/acquire at cluster scope; the acquire side invalidates L1 (CCTL.IVALL), so the next
        // step's plain loads of x see what the other CTAs of the cluster wrote in this one
        if (SINGLE) __syncthreads();
        else cluster.sync();
    }
}

// -------------------------------------------------------------------------------------------------
// Software-pipelined wavefront Gauss-Seidel (production kernel).
//
// A point update needs exactly D*3 values written in the immediately preceding step: u, u_c, u_cc of the
// backward distance-1 neighbour along each axis.  Everything else it reads (the other 7 offsets per axis, b)
// was final one step earlier (kGsLag = radius + 2).  So each step is split in two halves around a SPLIT
// cluster barrier:
//     A(t): finish the points of step t   = stashed partial residual - backward-1 couplings, channel solve, store
//     barrier.cluster.arrive.release
//     B(t): pre-gather for the points of step t+1 (gs_pre_elem) into the stash
//     barrier.cluster.wait.acquire
// The critical path between two barriers is one batch of loads plus the channel solve; the long gather runs in
// the shadow of the barrier and of the other CTAs' A halves.  The stash of a thread's first point of a step stays
// in registers; further points (levels whose busiest step has more points than the cluster has threads) go through
// a small per-thread global stash that stays L2 resident (it is rewritten every step).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr int kStashSlots = 8;   // doubles per stashed point: M <= 7 residuals + (wave index, coord) packed

template <int D, int THREADS, int PS, bool SINGLE>
__global__ void __launch_bounds__(THREADS, 1) k_gs_pipe(LevelDev L, const double* __restrict__ T,
                                                        const double* __restrict__ coef,
                                                        const double* __restrict__ dinv,
                                                        const double* __restrict__ b, double* x, double* stash,
                                                        size_t stash_stride, int nsweeps, const int* done) {
    if (done && *done) return;
    constexpr int M = 1 + 2 * D;
    extern __shared__ __align__(16) unsigned char gs_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = SINGLE ? 1 : (int)cluster.num_blocks();
    const int rank = SINGLE ? 0 : (int)cluster.block_rank();
    const int ib = blockIdx.x / csize;
    // Point idx of a step is handled by "thread" tid = idx mod nthreads.  Warps are dealt round-robin over the
    // cluster's CTAs (warp-sized chunks keep every access coalesced): a step with fewer points than the cluster has
    // threads -- most steps on the coarser levels -- spreads over all SMs of the cluster instead of filling CTA 0
    // first, so each SM's load/store pipe sees 1/csize of the step's gathers.
    const int tid = (((int)threadIdx.x >> 5) * csize + rank) * 32 + ((int)threadIdx.x & 31);
    const int nthreads = csize * blockDim.x;
    const size_t o = (size_t)ib * L.M * L.G;
    const double* Ti = T + (size_t)ib * L.D * kTabEntries * kTabPitch;
    constexpr int PITCH = PS > 0 ? PS : kTabPitch;
    const double* Tuse = Ti;
    const int* rbuse = L.rowbase;
    const int* hs = L.hstart;
    // shared memory: [first-point stash: kStashSlots x THREADS doubles | K tables | rowbase | hstart]
    double* st0 = reinterpret_cast<double*>(gs_smem) + threadIdx.x;
    if (PS > 0) {
        double* Ts = reinterpret_cast<double*>(gs_smem) + kStashSlots * THREADS;
        int* rbs = reinterpret_cast<int*>(Ts + D * kTabEntries * PS);
        const int nrb = (L.S + 8) * L.N[0] + 8;
        int* hss = rbs + nrb;
        int maxn = L.N[0] > L.N[1] ? L.N[0] : L.N[1];
        maxn = (maxn > L.N[2] ? maxn : L.N[2]) + 2 * kTabPad;
        constexpr int PSD = PS > 0 ? PS : 1;
        for (int i = threadIdx.x; i < D * kTabEntries * PS; i += blockDim.x) {
            const int row = i / PSD, pos = i - row * PSD;
            Ts[i] = pos < maxn ? Ti[(size_t)row * kTabPitch + pos] : 0.0;
        }
        for (int i = threadIdx.x; i < nrb; i += blockDim.x) rbs[i] = L.rowbase[i - 4];
        for (int i = threadIdx.x; i <= L.S; i += blockDim.x) hss[i] = L.hstart[i];
        __syncthreads();
        Tuse = Ts;
        rbuse = rbs + 4;
        hs = hss;
    }
    const double* bo = b + o;
    const double* co = coef + o;
    const double* dvo = dinv + o;
    double* xo = x + o;
    double* st = stash + (size_t)ib * stash_stride + tid;
    // opaque per-instance base pointers: keeps every access at [64-bit base + 32-bit element index] (one
    // IMAD.WIDE.U32) instead of re-deriving base + instance offset + index in 64-bit arithmetic per load
    asm volatile("" : "+l"(bo), "+l"(co), "+l"(dvo), "+l"(xo), "+l"(st));
    __builtin_assume(__isGlobal(bo));
    __builtin_assume(__isGlobal(co));
    __builtin_assume(__isGlobal(dvo));
    __builtin_assume(__isGlobal(xo));
    __builtin_assume(__isGlobal(st));
    const int G = L.G;
    const int steps = L.S + kGsLag * (nsweeps - 1);

    // points of step t: the hyperplanes s_k = t - lag*k of the sweeps k in flight, concatenated
    auto step_sweeps = [&](int t, int& k_lo, int& k_hi) -> int {
        k_lo = (t - (L.S - 1) + kGsLag - 1) / kGsLag;
        if (k_lo < 0) k_lo = 0;
        k_hi = t / kGsLag;
        if (k_hi > nsweeps - 1) k_hi = nsweeps - 1;
        int total = 0;
        for (int k = k_lo; k <= k_hi; ++k) {
            const int s = t - kGsLag * k;
            total += hs[s + 1] - hs[s];
        }
        return total;
    };
    auto point_of = [&](int t, int k_lo, int k_hi, int idx) -> int {
        int rem = idx;
        for (int k = k_lo; k <= k_hi; ++k) {
            const int s = t - kGsLag * k;
            const int h0 = hs[s], cnt = hs[s + 1] - h0;
            if (rem < cnt) return h0 + rem;
            rem -= cnt;
        }
        return -1;
    };
    // this thread's first point of the step after the one being pre-gathered, and its coordinates: loaded a
    // whole half-step before they are needed
    int wN = -1, cfN = 0;
    auto first_point = [&](int t) {
        wN = -1;
        cfN = 0;
        if (t < steps) {
            int k_lo, k_hi;
            const int total = step_sweeps(t, k_lo, k_hi);
            if (tid < total) {
                wN = point_of(t, k_lo, k_hi, tid);
                cfN = L.coord[wN];
            }
        }
    };

    // B half: pre-gather every point this thread owns in step t; returns how many
    auto pregather = [&](int t) -> int {
        int k_lo, k_hi;
        const int total = step_sweeps(t, k_lo, k_hi);
        int w = wN, cf = cfN;
        first_point(t + 1);
        int n = 0;
        while (w >= 0) {
            // coordinates of the next point of this step: in flight during this point's gather
            const int idx_next = tid + (n + 1) * nthreads;
            const int w_next = idx_next < total ? point_of(t, k_lo, k_hi, idx_next) : -1;
            const int cf_next = w_next >= 0 ? L.coord[w_next] : 0;
            int i0, i1, i2;
            unpack_coord(cf, i0, i1, i2);
            // what the finishing half reads from DRAM-cold streams: pull it into L2 now.  (Helps on the levels
            // this kernel is used for by default; on the fine level, where the L2 thrashes, it is a loss: the
            // prefetch allocates with normal priority and undoes the evict-first hint of ld_stream.)
#pragma unroll
            for (int m = 0; m < M; ++m) {
                prefetch_l2(dvo + ((unsigned)m * (unsigned)G + (unsigned)w));
                if (coord_eq(cf)) prefetch_l2(co + ((unsigned)m * (unsigned)G + (unsigned)w));
            }
            double r[M];
            gs_pre_elem<D, PDEOP_GS_LD, PITCH>(L, rbuse, Tuse, bo, xo, w, i0, i1, i2, r);
            if (n == 0) {   // a thread's first point of a step: shared-memory stash (private slot, no sync needed)
#pragma unroll
                for (int m = 0; m < M; ++m) st0[m * THREADS] = r[m];
                st0[(kStashSlots - 1) * THREADS] = __hiloint2double(cf, w);
            } else {
                double* sp = st + (size_t)(n - 1) * kStashSlots * nthreads;
#pragma unroll
                for (int m = 0; m < M; ++m) sp[(unsigned)(m * nthreads)] = r[m];
                sp[(unsigned)((kStashSlots - 1) * nthreads)] = __hiloint2double(cf, w);
            }
            w = w_next;
            cf = cf_next;
            ++n;
        }
        return n;
    };

    first_point(0);
    int nA = pregather(0);
#ifdef PDEOP_GS_TIMING
    const bool dbg = blockIdx.x == 0 && threadIdx.x == 0;
    long long cA = 0, cB = 0, cW = 0, nP = 0;
#endif
    for (int t = 0; t < steps; ++t) {
#ifdef PDEOP_GS_TIMING
        const long long c0 = clock64();
        nP += nA;
#endif
        // A half: finish the points of step t
        if (nA > 0) {
            double r[M];
#pragma unroll
            for (int m = 0; m < M; ++m) r[m] = st0[m * THREADS];
            const double pk = st0[(kStashSlots - 1) * THREADS];
            const int w = __double2loint(pk), cf = __double2hiint(pk);
            int i0, i1, i2;
            unpack_coord(cf, i0, i1, i2);
            gs_fin_elem<D, PDEOP_GS_LD, PITCH>(L, rbuse, Tuse, co, dvo, xo, w, i0, i1, i2, coord_eq(cf), r);
        }
        for (int n = 1; n < nA; ++n) {
            const double* sp = st + (size_t)(n - 1) * kStashSlots * nthreads;
            double r[M];
#pragma unroll
            for (int m = 0; m < M; ++m) r[m] = sp[(unsigned)(m * nthreads)];
            const double pk = sp[(unsigned)((kStashSlots - 1) * nthreads)];
            const int w = __double2loint(pk), cf = __double2hiint(pk);
            int i0, i1, i2;
            unpack_coord(cf, i0, i1, i2);
            gs_fin_elem<D, PDEOP_GS_LD, PITCH>(L, rbuse, Tuse, co, dvo, xo, w, i0, i1, i2, coord_eq(cf), r);
        }
        // release the stores of A(t); the acquire side invalidates L1 (CCTL.IVALL), so the plain loads of the
        // next halves see what the other CTAs of the cluster wrote
#ifdef PDEOP_GS_TIMING
        const long long c1 = clock64();
#endif
        if (!SINGLE) cluster_arrive_release();
        nA = t + 1 < steps ? pregather(t + 1) : 0;
#ifdef PDEOP_GS_TIMING
        const long long c2 = clock64();
#endif
        if (SINGLE) __syncthreads();
        else cluster_wait_acquire();
#ifdef PDEOP_GS_TIMING
        const long long c3 = clock64();
        cA += c1 - c0;
        cB += c2 - c1;
        cW += c3 - c2;
#endif
    }
#ifdef PDEOP_GS_TIMING
    if (dbg) {
        g_gs_dbg[0] += steps;
        g_gs_dbg[1] += cA;
        g_gs_dbg[2] += cB;
        g_gs_dbg[3] += cW;
        g_gs_dbg[4] += nP;
    }
#endif
}

}  // namespace pdeop
#include "pdeop_gs_fast.cuh"
#include "pdeop_gs_line.cuh"
namespace pdeop {

// cross-check variant: one launch per step, no intra-kernel synchronisation
template <int D>
__global__ void __launch_bounds__(kThreads) k_gs_step(LevelDev L, const double* __restrict__ T,
                                                      const double* __restrict__ coef,
                                                      const double* __restrict__ dinv, const double* __restrict__ b,
                                                      double* x, int nsweeps, int t, const int* done) {
    if (done && *done) return;
    const int k = blockIdx.z;
    const int s = t - kGsLag * k;
    if (s < 0 || s >= L.S) return;
    const int h0 = L.hstart[s], cnt = L.hstart[s + 1] - h0;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    gs_elem<D, LdL2, kTabPitch>(L, L.rowbase, T + (size_t)blockIdx.y * L.D * kTabEntries * kTabPitch, coef + o, dinv + o,
                                b + o, x + o, h0 + j);
}

static int num_sms() {
    static std::map<int, int> per_dev;
    const int dev = be_current_device();
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int& n = per_dev[dev];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Process defaults, read once from the environment and never written afterwards:
//   PDEOP_GS_SMEM = 0 disables the shared-memory staging of the tables (A/B testing)
//   PDEOP_GS_SINGLE = n: levels whose busiest step has at most n*512 points run with one CTA per instance
//   PDEOP_GS_PIPE: default of the per-plan "gs_pipe" switch: 0 = unsplit cluster kernel everywhere, 1 = software-
//     pipelined kernel everywhere, 2 (default) = pipelined kernel on the latency-bound levels (busiest step <= 3
//     points per thread), where it was measured faster (B200: 32x32x32 1.00 vs 1.04 ms, 32x16x16 0.48 vs 0.50 ms per
//     5 sweeps, batch 32); on throughput-bound levels (32x64x64: 3.30 vs 3.08 ms) the stash round trip costs more
//     than the overlap gains
struct GsEnv {
    int threads, smem, single, pipe;
    GsEnv() {
        // 128 registers/thread; 256/384/1024-thread variants were measured and are not faster; 576 / 640 threads
        // (96 registers, ~150 bytes of spills) take 2.93 / 3.07 instead of 2.31 ms per fine-level call, 768 threads 4.2 ms
        threads = 512;
        const char* m = getenv("PDEOP_GS_SMEM");
        smem = (m && atoi(m) == 0) ? 0 : 1;
        const char* sg = getenv("PDEOP_GS_SINGLE");
        single = sg ? atoi(sg) : 0;
        const char* pp = getenv("PDEOP_GS_PIPE");
        pipe = pp ? atoi(pp) : 2;
    }
};
static const GsEnv& gs_env() {
    static const GsEnv e;
    return e;
}
int be_default_gs_pipe() { return gs_env().pipe; }

// Shrinks the cluster size until all instances' clusters are co-resident (one wave): a cluster needs its CTAs on
// SMs of one GPC, so e.g. 16 clusters of 8 do not fit a B200 although 128 SMs would suffice (measured: 2 waves,
// 3.35 ms vs 3.10 ms for 32 clusters of 4 with twice the work).  Results cached per (kernel, smem, cluster size).
template <class K>
static void fit_cluster_wave(cudaLaunchConfig_t& cfg, K kern, int B) {
    // answers of cudaOccupancyMaxActiveClusters per (device, kernel, shared memory, cluster size)
    static std::map<std::tuple<int, const void*, size_t, unsigned>, int> cache;
    const int dev = be_current_device();
    unsigned& cs = cfg.attrs[0].val.clusterDim.x;
    while (cs > 1) {
        const auto key = std::make_tuple(dev, (const void*)kern, (size_t)cfg.dynamicSmemBytes, cs);
        int nmax;
        {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            auto it = cache.find(key);
            if (it == cache.end()) {
                nmax = 0;
                cfg.gridDim = dim3((unsigned)(B * cs));
                if (cudaOccupancyMaxActiveClusters(&nmax, kern, &cfg) != cudaSuccess) {
                    cudaGetLastError();
                    nmax = 1 << 30;   // cannot tell: keep the size
                }
                cache[key] = nmax;
            } else {
                nmax = it->second;
            }
        }
        if (nmax >= B) break;
        cs /= 2;
    }
    cfg.gridDim = dim3((unsigned)(B * cs));
}

template <int D, int THREADS, int MINB, int PS, bool SINGLE>
static void launch_gs_inst(cudaLaunchConfig_t& cfg, const LevelDev& L, const double* T, const double* coef,
                           const double* dinv, const double* b, double* x, int nsweeps, const int* done) {
    auto kern = k_gs_cluster<D, THREADS, MINB, PDEOP_GS_LD, PS, SINGLE>;
    size_t smem = 0;
    if (PS > 0) {
        smem = (size_t)D * kTabEntries * PS * sizeof(double) + ((size_t)(L.S + 8) * L.N[0] + 8 + L.S + 1) * sizeof(int);
        ensure_dyn_smem((const void*)kern, smem);
    }
    cfg.dynamicSmemBytes = smem;
    if (!SINGLE) fit_cluster_wave(cfg, kern, (int)(cfg.gridDim.x / cfg.attrs[0].val.clusterDim.x));
    note(cudaLaunchKernelEx(&cfg, kern, L, T, coef, dinv, b, x, nsweeps, done));
}

template <int D, int THREADS, int PS, bool SINGLE>
static void launch_gs_pipe(cudaLaunchConfig_t& cfg, const LevelDev& L, const double* T, const double* coef,
                           const double* dinv, const double* b, double* x, double* stash, size_t stash_stride,
                           int nsweeps, const int* done) {
    auto kern = k_gs_pipe<D, THREADS, PS, SINGLE>;
    size_t smem = (size_t)kStashSlots * THREADS * sizeof(double);
    if (PS > 0)
        smem += (size_t)D * kTabEntries * PS * sizeof(double) + ((size_t)(L.S + 8) * L.N[0] + 8 + L.S + 1) * sizeof(int);
    ensure_dyn_smem((const void*)kern, smem);
    cfg.dynamicSmemBytes = smem;
    if (!SINGLE) fit_cluster_wave(cfg, kern, (int)(cfg.gridDim.x / cfg.attrs[0].val.clusterDim.x));
    note(cudaLaunchKernelEx(&cfg, kern, L, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done));
}

// k_gs_fast (pdeop_gs_fast.cuh): returns false when its shared-memory layout does not fit the level
template <int D, bool SINGLE, bool STAGE = true, int THREADS = 384>
static bool launch_gs_fast(cudaLaunchConfig_t& cfg, const LevelDev& L, const double* T, const double* coef,
                           const double* dinv, const double* b, double* x, int nsweeps, const int* done) {
    cfg.blockDim = dim3(THREADS);
    int P = L.N[0] > L.N[1] ? L.N[0] : L.N[1];
    P = P > L.N[2] ? P : L.N[2];
    const int nrb = (L.S + 8) * L.N[0] + 8;
    const size_t smem = GsFastSmem<D>::bytes(P, THREADS, nrb, L.S, STAGE);
    if (smem > (size_t)227 * 1024) return false;
    auto kern = k_gs_fast<D, THREADS, SINGLE, STAGE>;
    ensure_dyn_smem((const void*)kern, smem);
    cfg.dynamicSmemBytes = smem;
    if (!SINGLE) fit_cluster_wave(cfg, kern, (int)(cfg.gridDim.x / cfg.attrs[0].val.clusterDim.x));
    note(cudaLaunchKernelEx(&cfg, kern, L, T, coef, dinv, b, x, nsweeps, P, done));
    return true;
}

// k_gs_line (pdeop_gs_line.cuh): returns false when the level does not fit it (1-D, rows longer than a CTA, lines
// shorter than 16 points, more than 8 CTAs per instance, shared memory)
template <int D, int PS>
static bool launch_gs_line_ps(cudaStream_t s, const LevelDev& L, int B, const double* T, const double* coef,
                              const double* dinv, const double* b, double* x, double* stash, int nsweeps,
                              const int* done) {
    LineGeom g;
    if (!line_geom(L, nsweeps, PS, 8, (size_t)227 * 1024, kLineMaxThreads, g)) return false;
    if (g.C != 1 && g.C != 2 && g.C != 4 && g.C != 8) return false;
    if (g.C > 1 && !stash) return false;
    auto kern = k_gs_line<D, PS>;
    const size_t smem = (size_t)g.o_end * sizeof(double);
    ensure_dyn_smem((const void*)kern, smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(B * g.C));
    cfg.blockDim = dim3((unsigned)g.threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g.C > 1 ? 1 : 0;
    long long* flags = reinterpret_cast<long long*>(stash);
    if (g.C > 1) note(cudaMemsetAsync(flags, 0, (size_t)B * kLineFlagWords * sizeof(long long), s));
    LineStreams S;
    line_streams(L, coef, dinv, b, x, S);
    note(cudaLaunchKernelEx(&cfg, kern, L, g, T, S, x, flags, done));
    return true;
}

template <int D>
static bool launch_gs_line(cudaStream_t s, const LevelDev& L, int B, const double* T, const double* coef,
                           const double* dinv, const double* b, double* x, double* stash, int nsweeps,
                           const int* done) {
    int maxn = L.N[0] > L.N[1] ? L.N[0] : L.N[1];
    maxn = (maxn > L.N[2] ? maxn : L.N[2]) + 2 * kTabPad;
    if (maxn <= 40) return launch_gs_line_ps<D, 40>(s, L, B, T, coef, dinv, b, x, stash, nsweeps, done);
    if (maxn <= 72) return launch_gs_line_ps<D, 72>(s, L, B, T, coef, dinv, b, x, stash, nsweeps, done);
    if (maxn <= 136) return launch_gs_line_ps<D, 136>(s, L, B, T, coef, dinv, b, x, stash, nsweeps, done);
    if (maxn <= 264) return launch_gs_line_ps<D, 264>(s, L, B, T, coef, dinv, b, x, stash, nsweeps, done);
    return false;
}
template <>
bool launch_gs_line<1>(cudaStream_t, const LevelDev&, int, const double*, const double*, const double*, const double*,
                       double*, double*, int, const int*) {
    return false;
}

template <int D>
static void launch_gs_cluster(cudaStream_t s, const LevelDev& L, int B, const double* T, const double* coef,
                              const double* dinv, const double* b, double* x, double* stash, size_t stash_stride,
                              int nsweeps, const int* done, int gs_pipe) {
    const GsEnv& env = gs_env();
    // gs_pipe 5: the line-marching kernel wherever it fits, the mode-2 choice elsewhere (measured slower than the
    // hyperplane kernels on every level, profiles/r2_gs_experiments.md: kept as a tested variant, not the default)
    if (gs_pipe == 5) {
        if (launch_gs_line<D>(s, L, B, T, coef, dinv, b, x, stash, nsweeps, done)) return;
        gs_pipe = 2;
    }
    const int threads = env.threads;
    const int per_sm = 1;   // CTAs per SM (register-limited: 128 regs/thread x 512 threads)
    int maxn = L.N[0] > L.N[1] ? L.N[0] : L.N[1];
    maxn = (maxn > L.N[2] ? maxn : L.N[2]) + 2 * kTabPad;
    int ps = 0;
    if (env.smem) {
        if (maxn <= 40) ps = 40;
        else if (maxn <= 72) ps = 72;
        else if (maxn <= 136) ps = 136;
        else if (maxn <= 264 && D <= 2) ps = 264;
        // index tables must fit next to the K tables
        const size_t smem = (size_t)D * kTabEntries * ps * 8 + ((size_t)(L.S + 8) * L.N[0] + 8 + L.S + 1) * 4;
        if (smem + (size_t)kStashSlots * threads * 8 > (size_t)(per_sm == 2 ? 110 : 220) * 1024) ps = 0;
    }
    // cluster size: as many CTAs per instance as fit on the chip at once, capped by the portable
    // maximum (8) and by the work of one step (largest hyperplane x sweeps in flight)
    int maxh = 0;
    {
        int a = L.N[0] * L.N[1], bb = L.N[0] * L.N[2], c = L.N[1] * L.N[2];
        maxh = a < bb ? a : bb;
        maxh = maxh < c ? maxh : c;
    }
    int want = (maxh * (nsweeps < 5 ? nsweeps : 5) + threads - 1) / threads;
    int fit = per_sm * num_sms() / (B > 0 ? B : 1);
    int csize = 1;
    while (csize * 2 <= 8 && csize * 2 <= fit && csize < want) csize *= 2;
    const bool single = csize == 1 || want <= env.single;
    if (single) csize = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(B * csize));
    cfg.blockDim = dim3(threads);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define PDEOP_GS_DISPATCH(SG)                                                                            \
    switch (ps) {                                                                                        \
        case 40: launch_gs_inst<D, 512, 1, 40, SG>(cfg, L, T, coef, dinv, b, x, nsweeps, done); break;   \
        case 72: launch_gs_inst<D, 512, 1, 72, SG>(cfg, L, T, coef, dinv, b, x, nsweeps, done); break;   \
        case 136: launch_gs_inst<D, 512, 1, 136, SG>(cfg, L, T, coef, dinv, b, x, nsweeps, done); break; \
        case 264: launch_gs_inst<D, 512, 1, 264, SG>(cfg, L, T, coef, dinv, b, x, nsweeps, done); break; \
        default: launch_gs_inst<D, 512, 1, 0, SG>(cfg, L, T, coef, dinv, b, x, nsweeps, done); break;    \
    }
#define PDEOP_GS_PIPE_DISPATCH(SG)                                                                                   \
    switch (ps) {                                                                                                    \
        case 40: launch_gs_pipe<D, 512, 40, SG>(cfg, L, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done); break;   \
        case 72: launch_gs_pipe<D, 512, 72, SG>(cfg, L, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done); break;   \
        case 136: launch_gs_pipe<D, 512, 136, SG>(cfg, L, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done); break; \
        case 264: launch_gs_pipe<D, 512, 264, SG>(cfg, L, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done); break; \
        default: launch_gs_pipe<D, 512, 0, SG>(cfg, L, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done); break;    \
    }
    // (2-D grids: the unsplit kernel is faster -- Burgers 256x256, batch 64: 78.1 vs 74.2 solves/s)
    // Latency-bound level: the busiest step holds at most 3 points per thread.  gs_pipe: 0 unsplit cluster kernel on
    // every level; 1 software-pipelined kernel on every level; 2 (default) k_gs_fast (cp.async-staged operands, all
    // gathers of a point in flight) on the latency-bound 3-D levels, unsplit kernel elsewhere; 3 k_gs_fast on every
    // level it fits; 4 as 2 with the software-pipelined kernel instead of k_gs_fast (the round-1 default).
    const bool latency_level = D == 3 && want <= 3 * csize;
    const bool use_pipe = stash && (gs_pipe == 1 || (gs_pipe == 4 && latency_level));
    if (!single && (gs_pipe == 3 || (gs_pipe == 2 && latency_level))) {
        const bool ok = launch_gs_fast<D, false>(cfg, L, T, coef, dinv, b, x, nsweeps, done);
        if (ok) return;
        cfg.blockDim = dim3(threads);
    }
    if (use_pipe) {
        if (single) { PDEOP_GS_PIPE_DISPATCH(true) } else { PDEOP_GS_PIPE_DISPATCH(false) }
    } else if (single) { PDEOP_GS_DISPATCH(true) } else { PDEOP_GS_DISPATCH(false) }
#undef PDEOP_GS_PIPE_DISPATCH
#undef PDEOP_GS_DISPATCH
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_dinv(LevelDev L, const double* __restrict__ T,
                                                   const double* __restrict__ coef, double* __restrict__ dinv) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    dinv_elem<D>(L, T + (size_t)blockIdx.y * L.D * kTabEntries * kTabPitch, coef + o, dinv + o, w);
}
void be_dinv(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, double* dinv) {
    dim3 grid(cdiv(L.G, kThreads), B);
    cudaStream_t s = (cudaStream_t)st;
    if (L.D == 1) k_dinv<1><<<grid, kThreads, 0, s>>>(L, T, coef, dinv);
    else if (L.D == 2) k_dinv<2><<<grid, kThreads, 0, s>>>(L, T, coef, dinv);
    else k_dinv<3><<<grid, kThreads, 0, s>>>(L, T, coef, dinv);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

void be_gs(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, const double* dinv,
           const double* b, double* x, double* stash, size_t stash_stride, int nsweeps, const int* done, int variant,
           int gs_pipe) {
    if (nsweeps <= 0) return;
    cudaStream_t s = (cudaStream_t)st;
    if (variant == 0) {
        if (L.D == 1) launch_gs_cluster<1>(s, L, B, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done, gs_pipe);
        else if (L.D == 2) launch_gs_cluster<2>(s, L, B, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done, gs_pipe);
        else launch_gs_cluster<3>(s, L, B, T, coef, dinv, b, x, stash, stash_stride, nsweeps, done, gs_pipe);
        PDEOP_COUNT(1);
        PDEOP_LAUNCH_CHECK();
        return;
    }
    int maxh = 1;
    {
        int a = L.N[0] * L.N[1], bb = L.N[0] * L.N[2], c = L.N[1] * L.N[2];
        maxh = a < bb ? a : bb;
        maxh = maxh < c ? maxh : c;
    }
    const int steps = L.S + kGsLag * (nsweeps - 1);
    dim3 grid(cdiv(maxh, kThreads), B, nsweeps);
    for (int t = 0; t < steps; ++t) {
        if (L.D == 1) k_gs_step<1><<<grid, kThreads, 0, s>>>(L, T, coef, dinv, b, x, nsweeps, t, done);
        else if (L.D == 2) k_gs_step<2><<<grid, kThreads, 0, s>>>(L, T, coef, dinv, b, x, nsweeps, t, done);
        else k_gs_step<3><<<grid, kThreads, 0, s>>>(L, T, coef, dinv, b, x, nsweeps, t, done);
        PDEOP_COUNT(1);
    }
    PDEOP_LAUNCH_CHECK();
}

// =================================================================================================
// dense coarse operator and gradients
// =================================================================================================
template <int D>
__global__ void __launch_bounds__(kThreads) k_dense(LevelDev L, const double* __restrict__ T,
                                                    const double* __restrict__ coef, double* __restrict__ Kd) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t n = (size_t)L.M * L.G;
    dense_elem<D>(L, T + (size_t)blockIdx.y * L.D * kTabEntries * kTabPitch, coef + (size_t)blockIdx.y * n,
                  Kd + (size_t)blockIdx.y * n * n, w);
}
// Row r of every matrix: columns [r - bwz, r].  The blocked factorisation reads at most bw + kOuter - 1 columns left
// of the diagonal (a trailing-update tile row against the first column of its panel), the solves at most bw + 158;
// nothing reads right of the diagonal.  At n = 14336, bw = 1798, B = 32 this writes 7.9 GB instead of 52.6 GB.
__global__ void __launch_bounds__(256) k_zero_band(int n, int bwz, double* Kd) {
    const int r = blockIdx.x;
    double* row = Kd + ((size_t)blockIdx.y * n + r) * n;
    const int lo = r - bwz > 0 ? r - bwz : 0;
    for (int c = lo + threadIdx.x; c <= r; c += blockDim.x) row[c] = 0.0;
}
void be_zero_dense(stream_t st, int B, int n, int bw, double* Kd) {
    const int bwz = bw + 384;   // kOuter (256) + a tile of slack
    if (bwz >= n - 1 || n < 1024) {
        note(cudaMemsetAsync(Kd, 0, (size_t)B * n * n * sizeof(double), (cudaStream_t)st));
        return;
    }
    k_zero_band<<<dim3(n, B), 256, 0, (cudaStream_t)st>>>(n, bwz, Kd);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

void be_dense(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, double* Kd) {
    dim3 grid(cdiv(L.G, kThreads), B);
    cudaStream_t s = (cudaStream_t)st;
    if (L.D == 1) k_dense<1><<<grid, kThreads, 0, s>>>(L, T, coef, Kd);
    else if (L.D == 2) k_dense<2><<<grid, kThreads, 0, s>>>(L, T, coef, Kd);
    else k_dense<3><<<grid, kThreads, 0, s>>>(L, T, coef, Kd);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_grads(LevelDev L, const double* __restrict__ coef,
                                                    const double* __restrict__ rhs_nat, const double* __restrict__ cv,
                                                    const double* __restrict__ fv, const double* __restrict__ bv,
                                                    const double* __restrict__ x, const double* __restrict__ dz,
                                                    double* __restrict__ d_coeffs, double* __restrict__ d_rhs,
                                                    double* d_cv, double* d_fv, double* d_bv) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const int ib = blockIdx.y;
    const size_t n = (size_t)L.M * L.G;
    const size_t oc = (size_t)ib * L.Ntot * 12, of = (size_t)ib * L.Ftot * 4;
    grad_elem<D>(L, coef + ib * n, rhs_nat + (size_t)ib * L.G, cv + oc, fv + of, bv + of, x + ib * n, dz + ib * n,
                 d_coeffs + ib * n, d_rhs + (size_t)ib * L.G, d_cv + oc, d_fv + of, d_bv + of, w);
}
__global__ void __launch_bounds__(kThreads) k_grad_init(LevelDev L, const double* __restrict__ dz,
                                                        double* __restrict__ d_iv) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.n_init) return;
    grad_init_elem(L, dz + (size_t)blockIdx.y * L.M * L.G, d_iv + (size_t)blockIdx.y * L.n_init, k);
}
void be_grads(stream_t st, const LevelDev& L, int B, const double* coef, const double* rhs_nat, const double* cv,
              const double* fv, const double* bv, const double* x, const double* dz, double* d_coeffs, double* d_rhs,
              double* d_iv, double* d_cv, double* d_fv, double* d_bv) {
    dim3 grid(cdiv(L.G, kThreads), B);
    cudaStream_t s = (cudaStream_t)st;
    if (L.D == 1) k_grads<1><<<grid, kThreads, 0, s>>>(L, coef, rhs_nat, cv, fv, bv, x, dz, d_coeffs, d_rhs, d_cv, d_fv, d_bv);
    else if (L.D == 2) k_grads<2><<<grid, kThreads, 0, s>>>(L, coef, rhs_nat, cv, fv, bv, x, dz, d_coeffs, d_rhs, d_cv, d_fv, d_bv);
    else k_grads<3><<<grid, kThreads, 0, s>>>(L, coef, rhs_nat, cv, fv, bv, x, dz, d_coeffs, d_rhs, d_cv, d_fv, d_bv);
    PDEOP_COUNT(1 + (L.n_init > 0));
    PDEOP_LAUNCH_CHECK();
    if (L.n_init > 0) {
        k_grad_init<<<dim3(cdiv(L.n_init, kThreads), B), kThreads, 0, s>>>(L, dz, d_iv);
        PDEOP_LAUNCH_CHECK();
    }
}

// =================================================================================================
// Krylov vector kernels
// =================================================================================================
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum of a block's values; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm /*[8]*/) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// deterministic re-summation of nblk block partials by one warp (fixed order for a given nblk)
__device__ __forceinline__ double warp_sum_partials(const double* __restrict__ p, int nblk) {
    double a = 0.0;
    for (int i = threadIdx.x & 31; i < nblk; i += 32) a += p[i];
    return warp_sum(a);
}

static inline int reduce_blocks(size_t n) {
    size_t b = (n + (size_t)kThreads * 4 - 1) / ((size_t)kThreads * 4);
    if (b < 1) b = 1;
    if (b > (size_t)kReduceBlocks) b = kReduceBlocks;
    return (int)b;
}

// partial[(k0+k)*kReduceBlocks + blk] = sum_i V[(k0+k)*ldv + i] * w[i],  k < K
// vectors of even length and stride on 16-byte boundaries can be streamed as double2
__device__ __forceinline__ bool vec2_ok(size_t n, size_t ld, const void* a, const void* b) {
    return (n % 2 == 0) && (ld % 2 == 0) && (((uintptr_t)a | (uintptr_t)b) % 16 == 0);
}

template <int K>
__global__ void __launch_bounds__(kThreads) k_dots(size_t n, const double* __restrict__ V, size_t ldv, int k0,
                                                   const double* __restrict__ w, double* __restrict__ partial,
                                                   const int* done) {
    if (done && *done) return;
    __shared__ double sm[8];
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec2_ok(n, ldv, V, w)) {   // 16-byte loads: half the instructions, twice the contiguous bytes per warp and stream
        const size_t n2 = n / 2, ld2 = ldv / 2;
        const double2* __restrict__ w2 = reinterpret_cast<const double2*>(w);
        const double2* __restrict__ V2 = reinterpret_cast<const double2*>(V);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            const double2 wi = w2[i];
            double2 vk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) vk[k] = V2[(size_t)(k0 + k) * ld2 + i];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += vk[k].x * wi.x + vk[k].y * wi.y;
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const double wi = w[i];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += V[(size_t)(k0 + k) * ldv + i] * wi;
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double r = block_sum(acc[k], sm);
        if (threadIdx.x == 0) partial[(size_t)(k0 + k) * kReduceBlocks + blockIdx.x] = r;
    }
}

static void launch_dots(cudaStream_t s, size_t n, const double* V, size_t ldv, int nv, const double* w,
                        double* partial, const int* done) {
    const int nb = reduce_blocks(n);
    int k0 = 0;
    // up to 10 vectors (the default restart length) in one pass over w
#define PDEOP_DOTS_CASE(KK) \
    case KK: k_dots<KK><<<nb, kThreads, 0, s>>>(n, V, ldv, 0, w, partial, done); k0 = nv; PDEOP_COUNT(1); break;
    switch (nv) {
        PDEOP_DOTS_CASE(3) PDEOP_DOTS_CASE(5) PDEOP_DOTS_CASE(6) PDEOP_DOTS_CASE(7) PDEOP_DOTS_CASE(9) PDEOP_DOTS_CASE(10)
        default: break;
    }
#undef PDEOP_DOTS_CASE
    while (k0 < nv) {
        const int rem = nv - k0;
        if (rem >= 8) { k_dots<8><<<nb, kThreads, 0, s>>>(n, V, ldv, k0, w, partial, done); k0 += 8; }
        else if (rem >= 4) { k_dots<4><<<nb, kThreads, 0, s>>>(n, V, ldv, k0, w, partial, done); k0 += 4; }
        else if (rem >= 2) { k_dots<2><<<nb, kThreads, 0, s>>>(n, V, ldv, k0, w, partial, done); k0 += 2; }
        else { k_dots<1><<<nb, kThreads, 0, s>>>(n, V, ldv, k0, w, partial, done); k0 += 1; }
        PDEOP_COUNT(1);
    }
    PDEOP_LAUNCH_CHECK();
}

// =================================================================================================
// Converged mode: per-instance PCG vector kernels, polynomial smoother update, transpose restriction
// =================================================================================================
__global__ void __launch_bounds__(kThreads) k_bdot(size_t n, const double* __restrict__ a, const double* __restrict__ c,
                                                   double* out, const int* done) {
    if (done && *done) return;
    __shared__ double sm[8];
    const size_t o = (size_t)blockIdx.y * n;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc = fma(a[o + i], c[o + i], acc);
    const double r = block_sum(acc, sm);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, r);
}
static inline unsigned bvec_blocks(size_t n, int B) {
    size_t want = (n + (size_t)kThreads * 4 - 1) / ((size_t)kThreads * 4);
    const size_t cap = (size_t)(4 * 148 + B - 1) / B;   // ~4 CTAs per SM over the whole batch
    if (want > cap) want = cap;
    return (unsigned)(want < 1 ? 1 : want);
}
void be_bdot(stream_t st, size_t n, int B, const double* a, const double* c, double* out, const int* done) {
    k_bdot<<<dim3(bvec_blocks(n, B), B), kThreads, 0, (cudaStream_t)st>>>(n, a, c, out, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__device__ __forceinline__ double safe_ratio(double a, double b) {
    const double q = a / b;
    return isfinite(q) ? q : 0.0;   // cg.py:112,125 nan_to_num
}

__global__ void __launch_bounds__(kThreads) k_pcg_xr(size_t n, double* __restrict__ x, double* __restrict__ r,
                                                     const double* __restrict__ p, const double* __restrict__ Ap,
                                                     const double* rz, const double* pAp, const double* active,
                                                     double* rr, const int* done) {
    if (done && *done) return;
    __shared__ double sm[8];
    const int b = blockIdx.y;
    const double alpha = active[b] != 0.0 ? safe_ratio(rz[b], pAp[b]) : 0.0;
    const size_t o = (size_t)b * n;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        x[o + i] = fma(alpha, p[o + i], x[o + i]);
        const double rn = fma(-alpha, Ap[o + i], r[o + i]);
        r[o + i] = rn;
        acc = fma(rn, rn, acc);
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0) atomicAdd(rr + b, s);
}
void be_pcg_xr(stream_t st, size_t n, int B, double* x, double* r, const double* p, const double* Ap, const double* rz,
               const double* pAp, const double* active, double* rr, const int* done) {
    k_pcg_xr<<<dim3(bvec_blocks(n, B), B), kThreads, 0, (cudaStream_t)st>>>(n, x, r, p, Ap, rz, pAp, active, rr, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(kThreads) k_pcg_p(size_t n, double* __restrict__ p, const double* __restrict__ z,
                                                    const double* rz_new, const double* rz, const double* active,
                                                    const int* done) {
    if (done && *done) return;
    const int b = blockIdx.y;
    const double beta = active[b] != 0.0 ? safe_ratio(rz_new[b], rz[b]) : 0.0;
    const size_t o = (size_t)b * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[o + i] = fma(beta, p[o + i], z[o + i]);
}
void be_pcg_p(stream_t st, size_t n, int B, double* p, const double* z, const double* rz_new, const double* rz,
              const double* active, const int* done) {
    k_pcg_p<<<dim3(bvec_blocks(n, B), B), kThreads, 0, (cudaStream_t)st>>>(n, p, z, rz_new, rz, active, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void k_pcg_scalars(int B, double* active, double* rz, double* rz_new, double* pAp, double* rr, double* bnorm,
                              double rtol, FgmresState* s, int first) {
    __shared__ int any_active;
    __shared__ double worst;
    if (threadIdx.x == 0) {
        any_active = 0;
        worst = 0.0;
    }
    __syncthreads();
    if (!first && s->done) return;
    double my_worst = 0.0;
    int my_any = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        if (first) {
            bnorm[b] = sqrt(rr[b]);
            active[b] = bnorm[b] > 0.0 ? 1.0 : 0.0;   // cg.py:69 cont_mask
            rz[b] = 0.0;
        } else {
            const double rel = bnorm[b] > 0.0 ? sqrt(rr[b]) / bnorm[b] : 0.0;
            if (!(rel > rtol)) active[b] = 0.0;       // cg.py:131-133 res_mask
            my_worst = fmax(my_worst, rel);
            rz[b] = rz_new[b];
        }
        rz_new[b] = 0.0;
        pAp[b] = 0.0;
        rr[b] = 0.0;
        if (active[b] != 0.0) my_any = 1;
    }
    if (my_any) atomicOr(&any_active, 1);
    // max over the block through shared memory (values are non-negative: integer compare of the bit pattern)
    atomicMax(reinterpret_cast<unsigned long long*>(&worst), (unsigned long long)__double_as_longlong(my_worst));
    __syncthreads();
    if (threadIdx.x == 0) {
        if (first) {
            s->iters = 0;
            s->rnorm = 0.0;
        } else {
            s->iters += 1;
            s->rnorm = worst;
        }
        s->done = any_active ? 0 : 1;
    }
}
void be_pcg_scalars(stream_t st, int B, double* active, double* rz, double* rz_new, double* pAp, double* rr,
                    double* bnorm, double rtol, FgmresState* state, int first) {
    k_pcg_scalars<<<1, 256, 0, (cudaStream_t)st>>>(B, active, rz, rz_new, pAp, rr, bnorm, rtol, state, first);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(kThreads) k_poly_update(size_t n, double* __restrict__ x, double* __restrict__ d,
                                                          const double* __restrict__ r, const double* __restrict__ dinv,
                                                          double c1, double c2, const double* lam, int mode,
                                                          const int* done) {
    if (done && *done) return;
    const int b = blockIdx.y;
    double c2b = c2;
    if (mode == 1) c2b = c2 / lam[b];
    else if (mode == 2) c2b = fmin(c2, 1.8 / lam[b]);
    const size_t o = (size_t)b * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double dn = fma(c1, d[o + i], c2b * dinv[o + i] * r[o + i]);
        d[o + i] = dn;
        x[o + i] += dn;
    }
}
void be_copy(stream_t st, void* dst, const void* src, size_t bytes) {
    if (bytes) note(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)st));
}
void be_poly_update(stream_t st, size_t n, int B, double* x, double* d, const double* r, const double* dinv, double c1,
                    double c2, const double* lam, int mode, const int* done) {
    k_poly_update<<<dim3(bvec_blocks(n, B), B), kThreads, 0, (cudaStream_t)st>>>(n, x, d, r, dinv, c1, c2, lam, mode, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

// out (coarse) += P^T in (fine): every fine point scatters to the <= 8 coarse corners its prolongation reads
__global__ void __launch_bounds__(kThreads) k_restrict_t(LevelDev Lf, LevelDev Lc, int C, const double* __restrict__ in,
                                                         double* out, const int* done) {
    if (done && *done) return;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= Lf.G) return;
    restrict_t_elem(Lf, Lc, C, in + (size_t)blockIdx.y * C * Lf.G, out + (size_t)blockIdx.y * C * Lc.G, w);
}
void be_restrict_t(stream_t st, const LevelDev& Lf, const LevelDev& Lc, int B, int C, const double* in, double* out,
                   const int* done) {
    note(cudaMemsetAsync(out, 0, (size_t)B * C * Lc.G * sizeof(double), (cudaStream_t)st));
    k_restrict_t<<<dim3(cdiv(Lf.G, kThreads), B), kThreads, 0, (cudaStream_t)st>>>(Lf, Lc, C, in, out, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(kThreads) k_power_scale(size_t n, double* __restrict__ v, const double* __restrict__ Kv,
                                                          const double* __restrict__ dinv, const double* nrm2, double* lam,
                                                          int phase) {
    const int b = blockIdx.y;
    const size_t o = (size_t)b * n;
    if (phase == 0) {   // v <- dinv .* Kv
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            v[o + i] = dinv[o + i] * Kv[o + i];
        return;
    }
    const double nn = sqrt(nrm2[b]);
    if (blockIdx.x == 0 && threadIdx.x == 0) lam[b] = nn;
    const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[o + i] *= inv;
}
void be_power_step(stream_t st, size_t n, int B, double* v, const double* Kv, const double* dinv, double* lam,
                   double* tmpB) {
    cudaStream_t s = (cudaStream_t)st;
    const dim3 grid(bvec_blocks(n, B), B);
    k_power_scale<<<grid, kThreads, 0, s>>>(n, v, Kv, dinv, nullptr, lam, 0);
    note(cudaMemsetAsync(tmpB, 0, sizeof(double) * B, s));
    k_bdot<<<grid, kThreads, 0, s>>>(n, v, v, tmpB, nullptr);
    k_power_scale<<<grid, kThreads, 0, s>>>(n, v, nullptr, nullptr, tmpB, lam, 1);
    PDEOP_COUNT(3);
    PDEOP_LAUNCH_CHECK();
}

void be_state_reset(stream_t st, FgmresState* s) {
    note(cudaMemsetAsync(s, 0, sizeof(FgmresState), (cudaStream_t)st));
}

__global__ void k_begin_final(FgmresState* s, int nblk) {
    const double v = warp_sum_partials(state_partials(s, 0), nblk);
    if (threadIdx.x == 0) {
        s->bnorm = sqrt(v);
        s->iters = 0;
        s->rnorm = 0.0;
        s->done = (s->bnorm == 0.0) ? 1 : 0;   // fgmres.py:76-78
    }
}
void be_fg_begin(stream_t st, size_t n, const double* b, double* x, FgmresState* s) {
    cudaStream_t cs = (cudaStream_t)st;
    note(cudaMemsetAsync(x, 0, n * sizeof(double), cs));
    launch_dots(cs, n, b, 0, 1, b, state_partials(s, 0), nullptr);
    k_begin_final<<<1, 32, 0, cs>>>(s, reduce_blocks(n));
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void k_resnorm_final(FgmresState* s, int nblk, int maxiter, double atol) {
    if (s->done) return;
    const double v = warp_sum_partials(state_partials(s, 0), nblk);
    if (threadIdx.x == 0) {
        const double rn = sqrt(v);
        s->rnorm = rn;
        if (rn <= atol || s->iters >= maxiter) s->done = 1;   // fgmres.py:134
        else s->e[0] = rn;
    }
}
void be_fg_resnorm(stream_t st, size_t n, const double* r, FgmresState* s, int maxiter, double atol) {
    cudaStream_t cs = (cudaStream_t)st;
    launch_dots(cs, n, r, 0, 1, r, state_partials(s, 0), &s->done);
    k_resnorm_final<<<1, 32, 0, cs>>>(s, reduce_blocks(n), maxiter, atol);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(kThreads) k_first(size_t n, const double* __restrict__ r, double* __restrict__ V0,
                                                    const FgmresState* s) {
    if (s->done) return;
    const double rn = s->rnorm;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) V0[i] = r[i] / rn;
}
void be_fg_first(stream_t st, size_t n, const double* r, double* V0, FgmresState* s) {
    k_first<<<reduce_blocks(n), kThreads, 0, (cudaStream_t)st>>>(n, r, V0, s);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

// w -= sum_{k<=j} h_k V_k with h_k re-summed from the dot partials; block partials of ||w||^2 -> area 1
// KC = j + 1 at compile time (loads of all vectors in flight together), KC = 0: run-time loop
template <int KC>
__global__ void __launch_bounds__(kThreads) k_axpy_norm(size_t n, int j, int restart, const double* __restrict__ V,
                                                        double* __restrict__ w, FgmresState* s, int nblk) {
    if (s->done) return;
    __shared__ double h[kMaxRestart];
    __shared__ double sm[8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = wid; k <= j; k += (blockDim.x >> 5)) {
        const double v = warp_sum_partials(state_partials(s, 0) + (size_t)k * kReduceBlocks, nblk);
        if (lane == 0) {
            h[k] = v;
            if (blockIdx.x == 0) s->H[k * restart + j] = v;
        }
    }
    __syncthreads();
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (KC > 0 && vec2_ok(n, n, V, w)) {
        const size_t n2 = n / 2;
        double2* w2 = reinterpret_cast<double2*>(w);
        const double2* __restrict__ V2 = reinterpret_cast<const double2*>(V);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            double2 wi = w2[i];
            double2 vk[KC > 0 ? KC : 1];
#pragma unroll
            for (int k = 0; k < KC; ++k) vk[k] = V2[(size_t)k * n2 + i];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                wi.x -= h[k] * vk[k].x;
                wi.y -= h[k] * vk[k].y;
            }
            w2[i] = wi;
            acc += wi.x * wi.x + wi.y * wi.y;
        }
    } else
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double wi = w[i];
        if (KC > 0) {
            double vk[KC > 0 ? KC : 1];
#pragma unroll
            for (int k = 0; k < KC; ++k) vk[k] = V[(size_t)k * n + i];
#pragma unroll
            for (int k = 0; k < KC; ++k) wi -= h[k] * vk[k];
        } else {
            for (int k = 0; k <= j; ++k) wi -= h[k] * V[(size_t)k * n + i];
        }
        w[i] = wi;
        acc += wi * wi;
    }
    const double r = block_sum(acc, sm);
    if (threadIdx.x == 0) state_partials(s, 1)[blockIdx.x] = r;
}

// H[j+1][j] = ||w||, V_{j+1} = w / ||w||
__global__ void __launch_bounds__(kThreads) k_scale_next(size_t n, int j, int restart, double* __restrict__ V,
                                                         const double* __restrict__ w, FgmresState* s, int nblk) {
    if (s->done) return;
    __shared__ double nn_s;
    if (threadIdx.x < 32) {
        const double v = warp_sum_partials(state_partials(s, 1), nblk);
        if (threadIdx.x == 0) {
            nn_s = sqrt(v);
            if (blockIdx.x == 0) s->H[(j + 1) * restart + j] = nn_s;
        }
    }
    __syncthreads();
    if (j + 1 >= restart) return;
    const double nn = nn_s;
    double* vn = V + (size_t)(j + 1) * n;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec2_ok(n, n, V, w)) {
        double2* vn2 = reinterpret_cast<double2*>(vn);
        const double2* __restrict__ w2 = reinterpret_cast<const double2*>(w);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += stride) {
            const double2 a = w2[i];
            vn2[i] = make_double2(a.x / nn, a.y / nn);
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) vn[i] = w[i] / nn;
}

void be_fg_cgs(stream_t st, size_t n, int j, int restart, double* V, double* w, FgmresState* s) {
    cudaStream_t cs = (cudaStream_t)st;
    const int nb = reduce_blocks(n);
    launch_dots(cs, n, V, n, j + 1, w, state_partials(s, 0), &s->done);
#define PDEOP_AXPY_CASE(KK) case KK: k_axpy_norm<KK><<<nb, kThreads, 0, cs>>>(n, j, restart, V, w, s, nb); break;
    switch (j + 1) {
        PDEOP_AXPY_CASE(1) PDEOP_AXPY_CASE(2) PDEOP_AXPY_CASE(3) PDEOP_AXPY_CASE(4) PDEOP_AXPY_CASE(5)
        PDEOP_AXPY_CASE(6) PDEOP_AXPY_CASE(7) PDEOP_AXPY_CASE(8) PDEOP_AXPY_CASE(9) PDEOP_AXPY_CASE(10)
        default: k_axpy_norm<0><<<nb, kThreads, 0, cs>>>(n, j, restart, V, w, s, nb); break;
    }
#undef PDEOP_AXPY_CASE
    k_scale_next<<<(j + 1 < restart) ? nb : 1, kThreads, 0, cs>>>(n, j, restart, V, w, s, nb);
    PDEOP_COUNT(2);
    PDEOP_LAUNCH_CHECK();
}

__global__ void k_lstsq(FgmresState* s, int restart) {
    if (s->done) return;
    if (threadIdx.x == 0) hessenberg_lstsq(s->H, s->e, restart, s->y);
}
__global__ void __launch_bounds__(kThreads) k_update(size_t n, int restart, const double* __restrict__ Z,
                                                     double* __restrict__ x, FgmresState* s) {
    if (s->done) return;
    __shared__ double y[kMaxRestart];
    if (threadIdx.x < restart) y[threadIdx.x] = s->y[threadIdx.x];
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec2_ok(n, n, Z, x)) {
        double2* x2 = reinterpret_cast<double2*>(x);
        const double2* __restrict__ Z2 = reinterpret_cast<const double2*>(Z);
        const size_t n2 = n / 2;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            double2 xi = x2[i];
#pragma unroll 5
            for (int k = 0; k < restart; ++k) {
                const double2 z = Z2[(size_t)k * n2 + i];
                xi.x += y[k] * z.x;
                xi.y += y[k] * z.y;
            }
            x2[i] = xi;
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double xi = x[i];
        for (int k = 0; k < restart; ++k) xi += y[k] * Z[(size_t)k * n + i];
        x[i] = xi;
    }
}
__global__ void k_iters_add(FgmresState* s, int restart) {
    if (s->done) return;
    s->iters += restart;
}
void be_fg_update(stream_t st, size_t n, int restart, const double* Z, double* x, FgmresState* s) {
    cudaStream_t cs = (cudaStream_t)st;
    k_lstsq<<<1, 32, 0, cs>>>(s, restart);
    k_update<<<reduce_blocks(n), kThreads, 0, cs>>>(n, restart, Z, x, s);
    k_iters_add<<<1, 1, 0, cs>>>(s, restart);
    PDEOP_COUNT(3);
    PDEOP_LAUNCH_CHECK();
}

__global__ void k_info(const FgmresState* s, double* info4) {
    info4[0] = (double)s->iters;
    info4[1] = s->rnorm;
    info4[2] = s->bnorm;
    info4[3] = (double)s->chol_info;
}
void be_fg_info(stream_t st, const FgmresState* s, double* info4) {
    k_info<<<1, 1, 0, (cudaStream_t)st>>>(s, info4);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}
void be_fg_hess(stream_t st, const FgmresState* s, int restart, double* hess_out) {
    note(cudaMemcpyAsync(hess_out, s->H, sizeof(double) * (restart + 1) * restart, cudaMemcpyDeviceToDevice,
                         (cudaStream_t)st));
}

// =================================================================================================
// Batched dense Cholesky (lower, in place, row-major n x n per instance).
// Two-level blocked right-looking: outer panels of kOuter columns, inner steps of kInner columns
// (diag factor in shared memory, row-wise triangular solve of the panel, SYRK of the rest of the outer
// panel), then one deep SYRK of the trailing matrix per outer panel (depth kOuter => compute-bound).
// =================================================================================================
constexpr int kInner = 32;
constexpr int kOuter = 256;

__global__ void __launch_bounds__(256) k_chol_diag(int n, double* A, size_t strideA, int k0, int nb,
                                                   FgmresState* st) {
    __shared__ double a[kInner][kInner + 1];
    double* Ab = A + (size_t)blockIdx.x * strideA;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int i = idx / nb, j = idx % nb;
        a[i][j] = (j <= i) ? Ab[(size_t)(k0 + i) * n + k0 + j] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        if (tid == 0) {
            double d = a[j][j];
            if (!(d > 0.0)) {
                atomicCAS(&st->chol_info, 0, k0 + j + 1);
                d = 1.0;
            }
            a[j][j] = sqrt(d);
        }
        __syncthreads();
        const double djj = a[j][j];
        for (int i = j + 1 + tid; i < nb; i += blockDim.x) a[i][j] /= djj;
        __syncthreads();
        const int m = nb - j - 1;
        for (int idx = tid; idx < m * m; idx += blockDim.x) {
            const int i = j + 1 + idx / m, k = j + 1 + idx % m;
            if (k <= i) a[i][k] -= a[i][j] * a[k][j];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int i = idx / nb, j = idx % nb;
        if (j <= i) Ab[(size_t)(k0 + i) * n + k0 + j] = a[i][j];
    }
}

// rows r in [k0+nb, n):  A[r, k0:k0+nb] <- A[r, k0:k0+nb] * L11^-T
__global__ void __launch_bounds__(128) k_chol_trsm(int n, double* A, size_t strideA, int k0, int nb, int r1) {
    __shared__ double l[kInner][kInner + 1];
    double* Ab = A + (size_t)blockIdx.y * strideA;
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        const int i = idx / nb, j = idx % nb;
        l[i][j] = Ab[(size_t)(k0 + i) * n + k0 + j];
    }
    __syncthreads();
    // reciprocal diagonal once per CTA: 32 fp64 divisions per row (~40 instructions each) become multiplications
    __shared__ double rinv[kInner];
    if (threadIdx.x < nb) rinv[threadIdx.x] = 1.0 / l[threadIdx.x][threadIdx.x];
    __syncthreads();
    const int r = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r1) return;
    double* row = Ab + (size_t)r * n + k0;
    double v[kInner];
#pragma unroll
    for (int j = 0; j < kInner; ++j) v[j] = j < nb ? row[j] : 0.0;
#pragma unroll
    for (int j = 0; j < kInner; ++j) {
        if (j < nb) {
            double s = v[j];
#pragma unroll
            for (int t = 0; t < j; ++t) s -= v[t] * l[j][t];
            v[j] = s * rinv[j];
        }
    }
#pragma unroll
    for (int j = 0; j < kInner; ++j)
        if (j < nb) row[j] = v[j];
}

// C[i][j] -= sum_{t in [p0,p1)} A[i][t] A[j][t]   for i in [r0,r1), j in [c0,c1), i >= j
// Tensor-core path for fp64: mma.sync.m8n8k4.f64 (DMMA; fp64 has no tcgen05 kind).  128x128 tile per CTA.
// Both operands come from the same row-major panel: A fragment = P[i][k], B fragment (column-major k x n) = P[j][k].
// Shared tiles are [row][k] with a 20-double pitch: a half-warp's 64-bit fragment loads hit 32 distinct banks.
// The next K-chunk is prefetched into registers while the current one is multiplied.
constexpr int kSyrkT = 128;
constexpr int kSyrkK = 16;
constexpr int kSyrkLd = 20;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 16 warps as 4 (rows) x 4 (columns), each a 32x32 sub-tile = 4x4 m8n8 accumulator tiles (32 doubles per lane):
// with 8 warps and 64 accumulators per lane (206 registers) only 2 warps per scheduler were resident and the tensor
// pipe idled 50 % of the time behind their global-load and fixed-latency stalls (profiles/r1_syrk_dmma_tensor_pipe.csv).
constexpr int kSyrkThreads = 512;
// bw: half-bandwidth of the factor.  Row i of L is zero left of column i - bw, so a tile whose first row is i0 only
// needs the panel columns t >= i0 - bw: the lowest tile rows of a trailing update skip most of the panel depth
// (17 % of the tile work at bw = 1798, depth 256; the skipped products are exact zeros).
// TN: tile width.  128: 16 warps as 4 x 4, each a 32x32 sub-tile; 32 (block-column updates of the left-looking
// panel factorisation): 16 warps as 16 x 1, each 8 rows x 32 columns.
template <int TN>
__global__ void __launch_bounds__(kSyrkThreads, 1) k_syrk(int n, double* A, size_t strideA, int r0, int r1, int c0, int c1,
                                                          int p0, int p1, int bw) {
    constexpr int MT = TN == 128 ? 4 : 1, NT = 4;
    const int i0 = r0 + blockIdx.x * kSyrkT, j0 = c0 + blockIdx.y * TN;
    if (i0 + kSyrkT - 1 < j0) return;   // tile entirely above the diagonal
    if (i0 - bw > p0) p0 += (i0 - bw - p0) / kSyrkK * kSyrkK;
    if (p0 >= p1) return;
    __shared__ double As[kSyrkT][kSyrkLd];
    __shared__ double Bs[TN][kSyrkLd];
    double* Ab = A + (size_t)blockIdx.z * strideA;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int rbase = TN == 128 ? (warp & 3) * 32 : warp * 8, cbase = TN == 128 ? (warp >> 2) * 32 : 0;
    const int gid = lane >> 2, tig = lane & 3;
    const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row lr of the tile, 4 consecutive k
    const bool arow = i0 + lr < r1, brow = lr < TN && j0 + lr < c1;
    const double* ap = Ab + (size_t)(i0 + lr) * n + lk;
    const double* bp = Ab + (size_t)(j0 + (lr < TN ? lr : 0)) * n + lk;
    double acc[MT][NT][2];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    double ra[4], rb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = p0 + lk + q;
        ra[q] = (arow && k < p1) ? ap[p0 + q] : 0.0;
        rb[q] = (brow && k < p1) ? bp[p0 + q] : 0.0;
    }
    for (int kk = p0; kk < p1; kk += kSyrkK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            As[lr][lk + q] = ra[q];
            if (lr < TN) Bs[lr][lk + q] = rb[q];
        }
        __syncthreads();
        if (kk + kSyrkK < p1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = kk + kSyrkK + lk + q;
                ra[q] = (arow && k < p1) ? ap[kk + kSyrkK + q] : 0.0;
                rb[q] = (brow && k < p1) ? bp[kk + kSyrkK + q] : 0.0;
            }
        }
#pragma unroll
        for (int ks = 0; ks < kSyrkK / 4; ++ks) {
            double af[MT], bf[NT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) af[mt] = As[rbase + mt * 8 + gid][ks * 4 + tig];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bf[nt] = Bs[cbase + nt * 8 + gid][ks * 4 + tig];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int i = i0 + rbase + mt * 8 + gid;
        if (i >= r1) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j0 + cbase + nt * 8 + 2 * tig + e;
                if (j < c1 && i >= j) Ab[(size_t)i * n + j] -= acc[mt][nt][e];
            }
        }
    }
}

static void launch_syrk(cudaStream_t s, int B, int n, double* A, int r0, int r1, int c0, int c1, int p0, int p1,
                        int bw) {
    if (r0 >= r1 || c0 >= c1 || p0 >= p1) return;
    if (c1 - c0 <= 32) {
        dim3 grid(cdiv(r1 - r0, kSyrkT), 1, B);
        k_syrk<32><<<grid, kSyrkThreads, 0, s>>>(n, A, (size_t)n * n, r0, r1, c0, c1, p0, p1, bw);
    } else {
        dim3 grid(cdiv(r1 - r0, kSyrkT), cdiv(c1 - c0, kSyrkT), B);
        k_syrk<128><<<grid, kSyrkThreads, 0, s>>>(n, A, (size_t)n * n, r0, r1, c0, c1, p0, p1, bw);
    }
    PDEOP_COUNT(1);
}

// Band-limited: column k of L is nonzero only in rows [k, k+bw], so every update stops bw rows below its panel.
// Inverse of one kSolveBlk diagonal block of L (lower triangular): thread j builds column j by forward
// substitution; L entries are warp-uniform (broadcast) reads, the column lives in the output (coalesced).
__global__ void __launch_bounds__(kSolveBlk) k_trtri(int n, const double* __restrict__ Lf, size_t strideL,
                                                     double* Linv, int nblk) {
    const int kb = blockIdx.x, ib = blockIdx.y;
    const int k0 = kb * kSolveBlk;
    const int w = k0 + kSolveBlk < n ? kSolveBlk : n - k0;
    const double* Lb = Lf + (size_t)ib * strideL + (size_t)k0 * n + k0;
    double* Ib = Linv + ((size_t)ib * nblk + kb) * kSolveBlk * kSolveBlk;
    const int j = threadIdx.x;
    for (int i = 0; i < kSolveBlk; ++i) {
        double v = 0.0;
        if (j < w && i < w && i >= j) {
            v = (i == j) ? 1.0 : 0.0;
            for (int t = j; t < i; ++t) v -= Lb[(size_t)i * n + t] * Ib[(size_t)t * kSolveBlk + j];
            v /= Lb[(size_t)i * n + i];
        }
        Ib[(size_t)i * kSolveBlk + j] = v;
    }
    // transpose copy (upper triangular), used by the backward solve with the same warp-per-row kernel
    __syncthreads();
    double* It = Linv + ((size_t)gridDim.y * nblk + (size_t)ib * nblk + kb) * kSolveBlk * kSolveBlk;
    for (int i = 0; i < kSolveBlk; ++i) It[(size_t)j * kSolveBlk + i] = Ib[(size_t)i * kSolveBlk + j];
}

// =================================================================================================
// Chain solver, set-up half.  W = blockdiag(L_kk)^-1 L restricted to the band: with it
//     L^-1 = Wt^-1 D^-1   and   L^-T = D^-T Wt^-T     (D = blockdiag(L_kk), Wt = W with identity diagonal blocks),
// so both triangular solves become a chain of pure band GEMVs (no triangular block solve between two chain
// steps) framed by two batched block-diagonal products that are off the chain's critical path.
// W is stored twice, compactly (the n x n factor array keeps only the band, 1/8 of its bytes at n = 14336):
//   Wc [ib][r][pw]  : row r of block row k holds columns c_lo(k) .. k0-1,  c_lo(k) = max(0, k0 - pw)
//   WTc[ib][c][pwt] : row c of block column k holds rows (k+1)*256 .. of W[., c]  (zero where W stores nothing)
// so that the forward chain streams rows of Wc and the backward chain rows of WTc with the same kernel.
// One CTA: block row k, a tile of 32 band columns; thread i owns row i of the block, 32 accumulators.
// =================================================================================================
constexpr int kWTile = 32;
__global__ void __launch_bounds__(kSolveBlk) k_make_w(int n, int nblk, int pw, int pwt, const double* __restrict__ Lf,
                                                      size_t strideL, const double* __restrict__ LinvT,
                                                      double* __restrict__ Wc, double* __restrict__ WTc) {
    extern __shared__ __align__(16) double mw_p[];   // [kSolveBlk][kWTile]
    const int k = blockIdx.y + 1, ib = blockIdx.z;
    const int k0 = k * kSolveBlk;
    const int w = k0 + kSolveBlk < n ? kSolveBlk : n - k0;
    const int c_lo = k0 - pw > 0 ? k0 - pw : 0;
    const int ct0 = c_lo + blockIdx.x * kWTile;
    if (ct0 >= k0) return;
    const double* Lb = Lf + (size_t)ib * strideL;
    for (int idx = threadIdx.x; idx < kSolveBlk * kWTile; idx += blockDim.x) {
        const int j = idx / kWTile, c = idx % kWTile;
        mw_p[idx] = j < w ? Lb[(size_t)(k0 + j) * n + ct0 + c] : 0.0;
    }
    __syncthreads();
    const int i = threadIdx.x;
    const double* It = LinvT + ((size_t)ib * nblk + k) * kSolveBlk * kSolveBlk;   // It[j][i] = inv(L_kk)[i][j]
    double acc[kWTile];
#pragma unroll
    for (int c = 0; c < kWTile; ++c) acc[c] = 0.0;
    const int jmax = i | 31;   // warp-uniform: the last row of this warp
    for (int j = 0; j <= jmax; ++j) {
        const double a = It[(size_t)j * kSolveBlk + i];   // exact zero for j > i
        const double2* pj = reinterpret_cast<const double2*>(mw_p + j * kWTile);
#pragma unroll
        for (int c2 = 0; c2 < kWTile / 2; ++c2) {
            const double2 v = pj[c2];
            acc[2 * c2] = fma(a, v.x, acc[2 * c2]);
            acc[2 * c2 + 1] = fma(a, v.y, acc[2 * c2 + 1]);
        }
    }
    if (i >= w) return;
    double* wrow = Wc + ((size_t)ib * n + k0 + i) * pw + (ct0 - c_lo);
#pragma unroll
    for (int c = 0; c < kWTile; ++c) wrow[c] = acc[c];
#pragma unroll
    for (int c = 0; c < kWTile; ++c) {
        const int cc = ct0 + c;
        WTc[((size_t)ib * n + cc) * pwt + (k0 + i - (cc / kSolveBlk + 1) * kSolveBlk)] = acc[c];
    }
}

static inline double* align_doubles(double* p, size_t bytes) {
    return (double*)(((uintptr_t)p + bytes - 1) / bytes * bytes);
}
// chain-solver arrays inside the Linv area
static void chain_arrays(const ChainLayout& cl, int B, int n, double* Linv, double** Wc, double** WTc) {
    double* p = align_doubles(Linv + 2 * (size_t)B * cl.nblk * kSolveBlk * kSolveBlk, 256);
    *Wc = p;
    *WTc = p + (size_t)B * n * cl.pw;
}
bool be_default_chain() {
    static const bool on = [] {
        const char* e = getenv("PDEOP_CHAIN");   // 0: per-block-row launches (A/B testing)
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// =================================================================================================
// Small dense systems (the dense layer on short time grids: Kamani ODE, n = 72, thousands of instances;
// qp_dual_dense_normal_kkt.py:27-66).  The blocked band path above costs ~25 launches of mostly idle tiles per
// call at this size; here ONE CTA factors one instance in shared memory (right-looking, column by column), and
// ONE WARP solves one instance: both triangular sweeps walk ROWS of L (forward: dot products, backward: axpy
// updates), so L is streamed coalesced from global memory exactly once per sweep and needs no staging.
// =================================================================================================
constexpr int kSmallN = 96;

__global__ void __launch_bounds__(256) k_chol_small(int n, double* A, FgmresState* st) {
    extern __shared__ __align__(16) double cs_a[];   // [n][n+1]
    const int ld = n + 1;
    double* Ab = A + (size_t)blockIdx.x * n * n;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        cs_a[i * ld + j] = j <= i ? Ab[idx] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        if (tid == 0) {
            double d = cs_a[j * ld + j];
            if (!(d > 0.0)) {
                atomicCAS(&st->chol_info, 0, j + 1);
                d = 1.0;
            }
            cs_a[j * ld + j] = sqrt(d);
        }
        __syncthreads();
        const double rinv = 1.0 / cs_a[j * ld + j];
        for (int i = j + 1 + tid; i < n; i += blockDim.x) cs_a[i * ld + j] *= rinv;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = tid; idx < m * m; idx += blockDim.x) {
            const int r = idx / m, c = idx - r * m;
            if (c <= r) {
                const int i = j + 1 + r, k = j + 1 + c;
                cs_a[i * ld + k] = fma(-cs_a[i * ld + j], cs_a[k * ld + j], cs_a[i * ld + k]);
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        if (j <= i) Ab[idx] = cs_a[i * ld + j];
    }
}

// out = (L L^T)^-1 rhs, one warp per instance; vectors in band ordering.  Lane l holds entries l, l+32, l+64.
__global__ void __launch_bounds__(256) k_solve_small(int n, int B, const double* __restrict__ Lf,
                                                     const double* __restrict__ rhs, double* __restrict__ out,
                                                     const int* done) {
    if (done && *done) return;
    const int ib = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ib >= B) return;
    const int lane = threadIdx.x & 31;
    const double* Lb = Lf + (size_t)ib * n * n;
    constexpr int R = kSmallN / 32;
    double y[R];
#pragma unroll
    for (int r = 0; r < R; ++r) y[r] = lane + 32 * r < n ? rhs[(size_t)ib * n + lane + 32 * r] : 0.0;
    // forward: y_i = (b_i - sum_{k<i} L[i][k] y_k) / L[i][i], row i read contiguously
    for (int i = 0; i < n; ++i) {
        const double* row = Lb + (size_t)i * n;
        double a = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = lane + 32 * r;
            if (k < i) a = fma(row[k], y[r], a);
        }
        a = warp_sum(a);
        const double dii = row[i];
        if (lane == (i & 31)) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (r == (i >> 5)) y[r] = (y[r] - a) / dii;
        }
    }
    // backward: x_i = y_i / L[i][i]; then y_k -= L[i][k] x_i for k < i (row i again)
    for (int i = n - 1; i >= 0; --i) {
        const double* row = Lb + (size_t)i * n;
        const double dii = row[i];
        double xi = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r == (i >> 5)) xi = y[r];
        xi = __shfl_sync(0xffffffffu, xi, i & 31) / dii;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = lane + 32 * r;
            if (k < i) y[r] = fma(-row[k], xi, y[r]);
            else if (k == i) y[r] = xi;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (lane + 32 * r < n) out[(size_t)ib * n + lane + 32 * r] = y[r];
}

// Per-plan helper objects: a side stream (highest priority: its short, latency-bound panel kernels must win SM slots
// against the main stream's deep updates) and two events.  PDEOP_FACTOR_LOOKAHEAD=0 disables the look-ahead (A/B).
struct FactorAux {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_cols = nullptr, ev_panel = nullptr;
};
void* be_aux_create() {
    const char* e = getenv("PDEOP_FACTOR_LOOKAHEAD");
    if (e && atoi(e) == 0) return nullptr;
    FactorAux* a = new FactorAux();
    int lo = 0, hi = 0;
    note(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    note(cudaStreamCreateWithPriority(&a->side, cudaStreamNonBlocking, hi));
    note(cudaEventCreateWithFlags(&a->ev_cols, cudaEventDisableTiming));
    note(cudaEventCreateWithFlags(&a->ev_panel, cudaEventDisableTiming));
    if (!a->side || !a->ev_cols || !a->ev_panel) {
        be_aux_destroy(a);
        return nullptr;
    }
    return a;
}
void be_aux_destroy(void* aux) {
    FactorAux* a = (FactorAux*)aux;
    if (!a) return;
    if (a->side) cudaStreamDestroy(a->side);
    if (a->ev_cols) cudaEventDestroy(a->ev_cols);
    if (a->ev_panel) cudaEventDestroy(a->ev_panel);
    delete a;
}

// factor the block columns [K0, K1) of one outer panel (left-looking inside the panel); all of A's columns [K0, K1)
// must carry the trailing updates of the earlier panels
static void factor_panel(cudaStream_t s, int B, int n, int bw, double* Kd, FgmresState* state, int K0, int K1) {
    const size_t strideA = (size_t)n * n;
    for (int k0 = K0; k0 < K1; k0 += kInner) {
        const int nb = k0 + kInner < K1 ? kInner : K1 - k0;
        const long long lim = (long long)k0 + nb + bw;
        const int r1 = lim < n ? (int)lim : n;
        // bring block column k0 up to date with the panel columns factored so far (depth k0 - K0) just before it is
        // factored -- every block column is read-modify-written once instead of once per earlier inner block
        launch_syrk(s, B, n, Kd, k0, r1, k0, k0 + nb, K0, k0, bw);
        k_chol_diag<<<B, 256, 0, s>>>(n, Kd, strideA, k0, nb, state);
        PDEOP_COUNT(1);
        if (k0 + nb < r1) {
            k_chol_trsm<<<dim3(cdiv(r1 - k0 - nb, 128), B), 128, 0, s>>>(n, Kd, strideA, k0, nb, r1);
            PDEOP_COUNT(1);
        }
    }
}

void be_cholesky(stream_t st, int B, int n, int bw, double* Kd, double* Linv, FgmresState* state, bool use_chain,
                 void* aux) {
    cudaStream_t s = (cudaStream_t)st;
    if (n <= kSmallN) {
        const size_t smem = (size_t)n * (n + 1) * sizeof(double);
        ensure_dyn_smem((const void*)k_chol_small, smem);
        k_chol_small<<<B, 256, smem, s>>>(n, Kd, state);
        PDEOP_COUNT(1);
        PDEOP_LAUNCH_CHECK();
        return;
    }
    const size_t strideA = (size_t)n * n;
    FactorAux* fa = (FactorAux*)aux;
    if (!fa || n <= 2 * kOuter) {
        for (int K0 = 0; K0 < n; K0 += kOuter) {
            const int K1 = K0 + kOuter < n ? K0 + kOuter : n;
            factor_panel(s, B, n, bw, Kd, state, K0, K1);
            // trailing matrix: columns [K1, K1+bw), depth kOuter
            const long long lim = (long long)K1 + bw;
            const int r1 = lim < n ? (int)lim : n;
            launch_syrk(s, B, n, Kd, K1, r1, K1, r1, K0, K1, bw);
        }
    } else {
        // Look-ahead.  The trailing update of panel k is split by columns: (a) the columns of panel k+1, (b) the rest.
        // Panel k+1 reads and writes columns [K1, K2) only, (b) writes columns >= K2 and reads panel k's columns: the two
        // are independent, so panel k+1 (a chain of short latency-bound kernels on a few SMs) runs on the side stream
        // while the main stream streams (b) through the tensor pipe.  (kOuter is a multiple of the tile size: the
        // row range of (b) starts at K2, the tiles above the diagonal are never launched.)
        factor_panel(s, B, n, bw, Kd, state, 0, kOuter < n ? kOuter : n);
        for (int K0 = 0; K0 < n; K0 += kOuter) {
            const int K1 = K0 + kOuter < n ? K0 + kOuter : n;
            if (K1 >= n) break;
            const int K2 = K1 + kOuter < n ? K1 + kOuter : n;
            const long long lim = (long long)K1 + bw;
            const int r1 = lim < n ? (int)lim : n;
            const int ca = K2 < r1 ? K2 : r1;
            launch_syrk(s, B, n, Kd, K1, r1, K1, ca, K0, K1, bw);             // (a)
            note(cudaEventRecord(fa->ev_cols, s));
            note(cudaStreamWaitEvent(fa->side, fa->ev_cols, 0));
            factor_panel(fa->side, B, n, bw, Kd, state, K1, K2);              // panel k+1
            note(cudaEventRecord(fa->ev_panel, fa->side));
            if (ca < r1) launch_syrk(s, B, n, Kd, ca, r1, ca, r1, K0, K1, bw);   // (b)
            note(cudaStreamWaitEvent(s, fa->ev_panel, 0));
        }
    }
    const int nblk = (n + kSolveBlk - 1) / kSolveBlk;
    k_trtri<<<dim3(nblk, B), kSolveBlk, 0, s>>>(n, Kd, strideA, Linv, nblk);
    PDEOP_COUNT(1);
    const ChainLayout cl = be_chain_layout(n, bw);
    if (cl.use && use_chain) {
        double *Wc, *WTc;
        chain_arrays(cl, B, n, Linv, &Wc, &WTc);
        note(cudaMemsetAsync(WTc, 0, (size_t)B * n * cl.pwt * sizeof(double), s));   // entries W does not store
        const int smem = kSolveBlk * kWTile * (int)sizeof(double);
        ensure_dyn_smem((const void*)k_make_w, smem);
        const double* LinvT = Linv + (size_t)B * nblk * kSolveBlk * kSolveBlk;
        k_make_w<<<dim3(cdiv(cl.pw, kWTile), nblk - 1, B), kSolveBlk, smem, s>>>(n, nblk, cl.pw, cl.pwt, Kd, strideA, LinvT,
                                                                                Wc, WTc);
        PDEOP_COUNT(1);
    }
    PDEOP_LAUNCH_CHECK();
}

// =================================================================================================
// Blocked triangular solves  out = L^-T L^-1 rhs  (HBM-bound on L: each triangle is read once).
// Block rows of kSolveBlk; per block: a GEMV with everything already solved + a small triangular solve.
// =================================================================================================
// t[i] = r[i] - sum_{c in [c_lo,k0)} L[i][c] y[c]   for rows i in [k0, k0+w): one warp per row
__global__ void __launch_bounds__(256) k_fwd_gemv(int n, const double* __restrict__ Lf, size_t strideL, int k0, int w,
                                                  int c_lo, const double* r, const double* y, double* t,
                                                  const int* done) {
    if (done && *done) return;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= w) return;
    const int lane = threadIdx.x & 31;
    const int ib = blockIdx.y;
    const double* Lr = Lf + (size_t)ib * strideL + (size_t)(k0 + row) * n;
    const double* yb = y + (size_t)ib * n;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int c = c_lo + lane;
    for (; c + 96 < k0; c += 128) {
        a0 += Lr[c] * yb[c];
        a1 += Lr[c + 32] * yb[c + 32];
        a2 += Lr[c + 64] * yb[c + 64];
        a3 += Lr[c + 96] * yb[c + 96];
    }
    for (; c < k0; c += 32) a0 += Lr[c] * yb[c];
    const double a = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) t[(size_t)ib * n + k0 + row] = r[(size_t)ib * n + k0 + row] - a;
}

// y_k = M_kk t_k for a triangular kSolveBlk block (upper=0: lower triangular, upper=1: upper): one warp per row
__global__ void __launch_bounds__(256) k_blk_mv(int n, const double* __restrict__ Linv, int nblk, int kb, int w,
                                                const double* t, double* y, int upper, const int* done) {
    if (done && *done) return;
    __shared__ double ts[kSolveBlk];
    const int ib = blockIdx.y;
    const int k0 = kb * kSolveBlk;
    for (int i = threadIdx.x; i < w; i += blockDim.x) ts[i] = t[(size_t)ib * n + k0 + i];
    __syncthreads();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= w) return;
    const int lane = threadIdx.x & 31;
    const double* Ir = Linv + ((size_t)ib * nblk + kb) * kSolveBlk * kSolveBlk + (size_t)row * kSolveBlk;
    double a = 0.0;
    const int c0 = upper ? (row & ~31) : 0, c1 = upper ? w : row + 1;
    for (int c = c0 + lane; c < c1; c += 32)
        if (!upper || c >= row) a += Ir[c] * ts[c];
    a = warp_sum(a);
    if (lane == 0) y[(size_t)ib * n + k0 + row] = a;
}

// y[c] -= sum_{r in [k0,k0+w)} L[r][c] z[r]  for c in [c_lo,k0), and y[k0+i] = z[i] for the block itself;
// z = zsrc[k0..k0+w).  A CTA owns a tile of 128 columns; its 8 warps take the rows r = warp (mod 8) and read
// 4 x 256 B = 1 KB contiguous per row (DRAM-page friendly, like the forward GEMV); the 8 partial sums per column
// are combined through shared memory in a fixed order.
constexpr int kUpdCols = 128;
__global__ void __launch_bounds__(256) k_bwd_update(int n, const double* __restrict__ Lf, size_t strideL, int k0,
                                                    int w, int c_lo, int c_min, const double* __restrict__ zsrc,
                                                    double* y, const int* done) {
    if (done && *done) return;
    __shared__ double zs[kSolveBlk];
    __shared__ double part[8][kUpdCols];
    const int ib = blockIdx.y;
    double* yb = y + (size_t)ib * n;
    for (int i = threadIdx.x; i < w; i += blockDim.x) zs[i] = zsrc[(size_t)ib * n + k0 + i];
    __syncthreads();
    const int cbase = c_lo + blockIdx.x * kUpdCols;
    if (cbase >= k0) {   // tiles past the band columns copy the solved block into y
        const int i = (cbase - k0) + threadIdx.x;
        if (threadIdx.x < kUpdCols && i < w) yb[k0 + i] = zs[i];
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Lc = Lf + (size_t)ib * strideL + (size_t)k0 * n + cbase + lane;
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ok[q] = cbase + lane + 32 * q < k0 && cbase + lane + 32 * q >= c_min;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int r = warp;
    for (; r + 24 < w; r += 32) {   // 4 rows x 4 segments = 16 independent loads in flight per lane
        double l[4][4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int q = 0; q < 4; ++q) l[v][q] = ok[q] ? Lc[(size_t)(r + 8 * v) * n + 32 * q] : 0.0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const double zv = zs[r + 8 * v];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += l[v][q] * zv;
        }
    }
    for (; r < w; r += 8) {
        const double z0 = zs[r];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += (ok[q] ? Lc[(size_t)r * n + 32 * q] : 0.0) * z0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) part[warp][lane + 32 * q] = acc[q];
    __syncthreads();
    if (threadIdx.x < kUpdCols) {
        const int c = cbase + threadIdx.x;
        if (c < k0 && c >= c_min) {
            double sum = 0.0;
#pragma unroll
            for (int v = 0; v < 8; ++v) sum += part[v][threadIdx.x];
            yb[c] -= sum;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_to_band(LevelDev L, const double* __restrict__ wave,
                                                      double* __restrict__ bandv, const int* done) {
    if (done && *done) return;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    to_band_elem(L, wave + o, bandv + o, w);
}
__global__ void __launch_bounds__(kThreads) k_from_band(LevelDev L, const double* __restrict__ bandv,
                                                        double* __restrict__ wave, const int* done) {
    if (done && *done) return;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.G) return;
    const size_t o = (size_t)blockIdx.y * L.M * L.G;
    from_band_elem(L, bandv + o, wave + o, w);
}

// =================================================================================================
// Chain solver, solve half: y = Wt^-1 u (dir 0, rows of Wc) or v = Wt^-T y (dir 1, rows of WTc) as ONE persistent
// kernel per direction.  A cluster of 4 CTAs owns one instance; CTA q owns rows 32q..32q+31 of every 128-row block.
// Per chain step (one block, ascending for dir 0, descending for dir 1) a CTA streams its 32 band rows
// (~0.45 MB) from HBM with TMA bulk copies issued by two producer threads:
//   * the "old" part of each row (everything except the block solved in the previous step) goes, in chunks,
//     through shared-memory stages (mbarrier full/empty pairs, three private stages per consumer warp) and is
//     consumed by 8 warps, one row each, while the previous block's result is still in flight between the CTAs;
//   * the "new" part (32 rows x 128 columns of the previous step's block) is prefetched into its own buffer during
//     the old phase, so that when the previous block arrives only a 32 x 128 product remains on the critical path.
// Results are exchanged through distributed shared memory: every warp sends its 4 values into all four CTAs'
// vector windows (a ring of nw blocks) with st.async, which completes transaction bytes on the receiving CTA's
// "block ready" mbarrier; no fence, no cluster-wide barrier and no kernel boundary sits between two steps.
// HBM-bound: 8*(band entries) bytes per direction and instance, each read once.
// =================================================================================================
constexpr int kChainCta = 4;         // cluster size
constexpr int kChainRows = kSolveBlk / kChainCta;   // rows of a block per CTA (32)
static_assert(kChainRows == 32, "the new-block producer maps one lane to one row");
constexpr int kChainWarps = 8;       // consumer warps
constexpr int kChainDepth = 3;       // stages per consumer warp (each warp owns its stages: a parity wait is only
                                     // meaningful for a waiter that observes every phase of its barrier in order)
constexpr int kChainStages = kChainWarps * kChainDepth;
constexpr int kChainCHMax = 896;     // doubles per stage (one chunk of a row's old part): 24 stages = 168 KB in flight;
                                     // the launcher picks the largest multiple of 32 <= this that fits shared memory
constexpr int kChainThreads = (kChainWarps + 2) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
// bounded spin: a protocol error traps instead of hanging the GPU.  The bound is minutes of polling (each try_wait
// suspends for a while before it returns false), not a latency budget: a slow co-tenant must not become a trap.
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    for (long long spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1LL << 30)) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t a, uint32_t parity) {
    for (long long spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1LL << 30)) __trap();
    }
}
// asynchronous 8-byte store into CTA `rank`'s shared memory (same offset as the local address) that completes
// 8 transaction bytes on that CTA's mbarrier: data and "it has arrived" travel together, no release fence
__device__ __forceinline__ void st_async_remote(uint32_t local_addr, double v, uint32_t local_mbar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra, rm;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %3;\n\t"
        "mapa.shared::cluster.u32 rm, %2, %3;\n\t"
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [ra], %1, [rm];\n\t}" ::"r"(local_addr),
        "l"(__double_as_longlong(v)), "r"(local_mbar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}

struct ChainRow {
    const double* new_src;   // previous step's block columns of this row (null: none)
    int new_len;
    const double* old_src;   // everything older
    int old_len;             // multiple of 32
};

// geometry of local row j of step block k; v0 = vector index of old_src[0]
__device__ __forceinline__ ChainRow chain_row(int dir, int n, int nblk, int pw, int pwt, const double* Wi, int k, int s,
                                              int g) {
    ChainRow r;
    r.new_src = nullptr;
    r.new_len = 0;
    r.old_src = nullptr;
    r.old_len = 0;
    if (g >= n || s == 0) return r;
    const int k0 = k * kSolveBlk;
    if (dir == 0) {
        const int c_lo = k0 - pw > 0 ? k0 - pw : 0;
        const int ncols = k0 - c_lo;
        const double* base = Wi + (size_t)g * pw;
        r.new_src = base + (ncols - kSolveBlk);
        r.new_len = kSolveBlk;
        r.old_src = base;
        r.old_len = ncols - kSolveBlk;
    } else {
        const double* base = Wi + (size_t)g * pwt;
        const int left = n - (k + 1) * kSolveBlk;
        r.new_src = base;
        r.new_len = left < kSolveBlk ? left : kSolveBlk;
        int rho = (g + pw) / kSolveBlk;   // last block row whose band reaches column g
        if (rho > nblk - 1) rho = nblk - 1;
        r.old_src = base + kSolveBlk;
        r.old_len = rho > k + 1 ? (rho - (k + 1)) * kSolveBlk : 0;
    }
    return r;
}

__global__ void __launch_bounds__(kChainThreads, 1) k_band_chain(int dir, int n, int nblk, int pw, int pwt, int nw, int ch,
                                                                const double* __restrict__ W,
                                                                const double* __restrict__ vin,
                                                                double* __restrict__ vout, const int* done) {
    if (done && *done) return;
    extern __shared__ __align__(128) unsigned char ch_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    const int ib = blockIdx.x / kChainCta;
    double* newbuf = reinterpret_cast<double*>(ch_smem);                  // [kChainRows][kSolveBlk]
    constexpr int ns = kChainStages;
    double* ring = newbuf + kChainRows * kSolveBlk;                       // [kChainStages][ch]
    double* win = ring + (size_t)ns * ch;                           // [nw][kSolveBlk]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(win + (size_t)nw * kSolveBlk);
    // bars: full[ns] | empty[ns] | full_new | empty_new | ready[2]
    const uint32_t b_full = smem_u32(bars), b_empty = smem_u32(bars + ns);
    const uint32_t b_full_new = smem_u32(bars + 2 * ns), b_empty_new = smem_u32(bars + 2 * ns + 1);
    const uint32_t b_ready = smem_u32(bars + 2 * ns + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kChainRows * kSolveBlk; i += blockDim.x) newbuf[i] = 0.0;
    for (int i = tid; i < nw * kSolveBlk; i += blockDim.x) win[i] = 0.0;
    if (tid == 0) {
        for (int i = 0; i < ns; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        mbar_init(b_full_new, 1);
        mbar_init(b_empty_new, kChainWarps);
        mbar_init(b_ready, 1);       // one local arming arrival per phase + 8 bytes per delivered value
        mbar_init(b_ready + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();   // barriers and windows of all four CTAs are ready before anyone stores into them

    const size_t pitch = dir == 0 ? (size_t)pw : (size_t)pwt;
    const double* Wi = W + (size_t)ib * n * pitch;
    const double* vi = vin + (size_t)ib * n;
    double* vo = vout + (size_t)ib * n;

    if (warp == kChainWarps) {
        // ---- producers of the old chunks: lane wv feeds consumer warp wv (a single issuing thread is too slow:
        // ~0.25 us per chunk of address arithmetic, waiting and issuing, 64 chunks per step).  Consumer warp wv owns
        // rows wv + 8m and stages wv*depth ..; its chunks are numbered lq in its own order (row, then chunk),
        // stage = lq % depth, parity = (lq / depth) & 1.
        if (lane < kChainWarps) {
            const int wv = lane;
            unsigned lq = 0;
            for (int s = 1; s < nblk; ++s) {
                const int k = dir == 0 ? s : nblk - 1 - s;
                const int gl = k * kSolveBlk + q * kChainRows;
                const ChainRow last = chain_row(dir, n, nblk, pw, pwt, Wi, k, s, min(gl + kChainRows - 1, n - 1));
                const int nch = (last.old_len + ch - 1) / ch;
                for (int m = 0; m < kChainRows / kChainWarps; ++m) {
                    const ChainRow r = chain_row(dir, n, nblk, pw, pwt, Wi, k, s, gl + wv + kChainWarps * m);
                    for (int c = 0; c < nch; ++c, ++lq) {
                        const unsigned stage = (unsigned)wv * kChainDepth + lq % kChainDepth;
                        const unsigned par = (lq / kChainDepth) & 1u;
                        mbar_wait(b_empty + 8 * stage, par ^ 1u);
                        int len = r.old_len - c * ch;
                        if (len > ch) len = ch;
                        if (len > 0) {
                            mbar_expect_tx(b_full + 8 * stage, (uint32_t)len * 8u);
                            bulk_g2s(smem_u32(ring + (size_t)stage * ch), r.old_src + (size_t)c * ch,
                                     (uint32_t)len * 8u, b_full + 8 * stage);
                        } else {
                            mbar_arrive(b_full + 8 * stage);
                        }
                    }
                }
            }
        }
    } else if (warp == kChainWarps + 1) {
        // ---- producers of the new-block buffer: lane j copies local row j (kChainRows == 32) ----
        for (int s = 1; s < nblk; ++s) {
            const int k = dir == 0 ? s : nblk - 1 - s;
            const int gl = k * kSolveBlk + q * kChainRows;
            const ChainRow r = chain_row(dir, n, nblk, pw, pwt, Wi, k, s, gl + lane);
            uint32_t total = (uint32_t)r.new_len * 8u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) {
                mbar_wait(b_empty_new, ((unsigned)(s - 1) & 1u) ^ 1u);
                mbar_expect_tx(b_full_new, total);
            }
            __syncwarp();
            if (r.new_len > 0)
                bulk_g2s(smem_u32(newbuf + lane * kSolveBlk), r.new_src, (uint32_t)r.new_len * 8u, b_full_new);
        }
    } else {
        // ---- consumers: warp wv owns local rows wv, wv+8, ... ----
        const int wv = warp;
        constexpr int RW = kChainRows / kChainWarps;   // rows per warp (4)
        unsigned seq_base = 0;
#ifdef PDEOP_GS_TIMING
        long long tOld = 0, tReady = 0, tNew = 0, tFin = 0;
#endif
        for (int s = 0; s < nblk; ++s) {
#ifdef PDEOP_GS_TIMING
            const long long c0 = clock64();
            long long c1 = c0, c2 = c0, c3 = c0;
#endif
            const int k = dir == 0 ? s : nblk - 1 - s;
            const int kn = dir == 0 ? k - 1 : k + 1;
            const int gl = k * kSolveBlk + q * kChainRows;
            if (wv == 0 && lane == 0) {
                // arm this step's "block ready" phase: its previous phase (step s-2) completed before this warp's
                // previous new phase; deliveries that arrive before the arming just run the byte count negative
                const int wk = n - k * kSolveBlk < kSolveBlk ? n - k * kSolveBlk : kSolveBlk;
                mbar_expect_tx(b_ready + 8 * ((unsigned)s & 1u), (uint32_t)wk * 8u);
            }
            double uin[RW], acc[RW];
#pragma unroll
            for (int m = 0; m < RW; ++m) {
                const int g = gl + wv + kChainWarps * m;
                uin[m] = g < n ? vi[g] : 0.0;
                acc[m] = 0.0;
            }
            int nch = 0;
            if (s > 0) {
                const ChainRow last = chain_row(dir, n, nblk, pw, pwt, Wi, k, s, min(gl + kChainRows - 1, n - 1));
                nch = (last.old_len + ch - 1) / ch;
                const int k0 = k * kSolveBlk;
                const int v0 = dir == 0 ? (k0 - pw > 0 ? k0 - pw : 0) : (k + 2) * kSolveBlk;
                // old phase: needs blocks solved two or more steps ago, all delivered before the previous new phase
#pragma unroll
                for (int m = 0; m < RW; ++m) {
                    const int j = wv + kChainWarps * m;
                    const ChainRow r = chain_row(dir, n, nblk, pw, pwt, Wi, k, s, gl + j);
                    for (int c = 0; c < nch; ++c) {
                        const unsigned lq = seq_base + (unsigned)(m * nch + c);
                        const unsigned stage = (unsigned)wv * kChainDepth + lq % kChainDepth;
                        const unsigned par = (lq / kChainDepth) & 1u;
                        mbar_wait(b_full + 8 * stage, par);
                        int len = r.old_len - c * ch;
                        if (len > ch) len = ch;
                        if (len > 0) {
                            const double* wc = ring + (size_t)stage * ch;
                            const int gv = v0 + c * ch + 2 * lane;   // vector index of this lane's first pair
                            int slot = (gv / kSolveBlk) % nw, off = gv % kSolveBlk;
                            double a = acc[m];
                            for (int e = 2 * lane; e < len; e += 64) {
                                const double2 wv2 = *reinterpret_cast<const double2*>(wc + e);
                                const double2 yv2 = *reinterpret_cast<const double2*>(win + slot * kSolveBlk + off);
                                a = fma(wv2.x, yv2.x, a);
                                a = fma(wv2.y, yv2.y, a);
                                off += 64;
                                if (off >= kSolveBlk) {
                                    off -= kSolveBlk;
                                    slot = slot + 1 == nw ? 0 : slot + 1;
                                }
                            }
                            acc[m] = a;
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(b_empty + 8 * stage);
                    }
                }
                seq_base += (unsigned)(RW * nch);
#ifdef PDEOP_GS_TIMING
                c1 = clock64();
#endif
                // new phase: the previous step's block has arrived in this CTA's window
                mbar_wait_cluster(b_ready + 8 * ((unsigned)(s - 1) & 1u), ((unsigned)(s - 1) >> 1) & 1u);
#ifdef PDEOP_GS_TIMING
                c2 = clock64();
#endif
                mbar_wait(b_full_new, (unsigned)(s - 1) & 1u);
#ifdef PDEOP_GS_TIMING
                c3 = clock64();
#endif
                const double* yn = win + (kn % nw) * kSolveBlk;
#pragma unroll
                for (int m = 0; m < RW; ++m) {
                    const double* wr = newbuf + (wv + kChainWarps * m) * kSolveBlk;
                    double a = acc[m];
#pragma unroll
                    for (int it = 0; it < kSolveBlk / 64; ++it) {
                        const double2 wv2 = *reinterpret_cast<const double2*>(wr + 64 * it + 2 * lane);
                        const double2 yv2 = *reinterpret_cast<const double2*>(yn + 64 * it + 2 * lane);
                        a = fma(wv2.x, yv2.x, a);
                        a = fma(wv2.y, yv2.y, a);
                    }
                    acc[m] = a;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(b_empty_new);
            }
            // reduce, finish, and deliver this warp's 8 values to all four CTAs
            double mine = 0.0;
#pragma unroll
            for (int m = 0; m < RW; ++m) {
                const double tot = warp_sum(acc[m]);
                const double yv = uin[m] - tot;
                if ((lane & (RW - 1)) == m) mine = yv;
            }
            {
                const int m = lane & (RW - 1), dst = lane / RW;   // lanes 0..RW*4-1 = 4 CTAs x RW rows
                const int jl = wv + kChainWarps * m;
                const int g = gl + jl;
                if (dst < kChainCta && g < n) {
                    const double* slotp = win + (k % nw) * kSolveBlk + q * kChainRows + jl;
                    st_async_remote(smem_u32(slotp), mine, b_ready + 8 * ((unsigned)s & 1u), (uint32_t)dst);
                    if (dst == 0) vo[g] = mine;
                }
            }
#ifdef PDEOP_GS_TIMING
            const long long c4 = clock64();
            tOld += c1 - c0;
            tReady += c2 - c1;
            tNew += c3 - c2;
            tFin += c4 - c3;
#endif
        }
#ifdef PDEOP_GS_TIMING
        if (blockIdx.x == 0 && tid == 0) {
            g_gs_dbg[0] += nblk;
            g_gs_dbg[1] += tOld;
            g_gs_dbg[2] += tReady;
            g_gs_dbg[3] += tNew;
            g_gs_dbg[4] += tFin;
        }
#endif
    }
    __syncthreads();
    cluster.sync();   // no CTA leaves while others may still store into its window or arrive on its barriers
}

static size_t chain_smem_bytes(int nw, int ch) {
    return ((size_t)kChainRows * kSolveBlk + (size_t)kChainStages * ch + (size_t)nw * kSolveBlk) * sizeof(double) +
           (size_t)(2 * kChainStages + 4) * 8;
}

// y_k = M_kk t_k for every block at once: the block-diagonal products that frame the chain.  One CTA per (block,
// instance), thread r owns row r and walks the columns; it reads the TRANSPOSED copy MT of the triangular block
// (MT[c][r] = M[r][c]), so that every step is one coalesced row segment and only the triangle is read.
//   upper = 0: M lower triangular (pass MT = transposed inverse blocks);  upper = 1: M upper triangular (pass the
//   untransposed inverse blocks).
__global__ void __launch_bounds__(kSolveBlk) k_blk_mv_all(int n, const double* __restrict__ MT, int nblk, const double* t,
                                                          double* y, int upper, const int* done) {
    if (done && *done) return;
    __shared__ double ts[kSolveBlk];
    const int kb = blockIdx.x, ib = blockIdx.y;
    const int k0 = kb * kSolveBlk;
    const int w = k0 + kSolveBlk < n ? kSolveBlk : n - k0;
    const int r = threadIdx.x;
    ts[r] = r < w ? t[(size_t)ib * n + k0 + r] : 0.0;
    __syncthreads();
    const double* Mb = MT + ((size_t)ib * nblk + kb) * kSolveBlk * kSolveBlk + r;
    double a0 = 0.0, a1 = 0.0;
    if (!upper) {
        // warp-uniform bound: the last row of this warp; entries above the diagonal are stored zeros
        const int cmax = (r | 31) < w ? (r | 31) : w - 1;
        int c = 0;
        for (; c + 1 <= cmax; c += 2) {
            a0 = fma(Mb[(size_t)c * kSolveBlk], ts[c], a0);
            a1 = fma(Mb[(size_t)(c + 1) * kSolveBlk], ts[c + 1], a1);
        }
        if (c <= cmax) a0 = fma(Mb[(size_t)c * kSolveBlk], ts[c], a0);
    } else {
        int c = r & ~31;
        for (; c + 1 < w; c += 2) {
            a0 = fma(Mb[(size_t)c * kSolveBlk], ts[c], a0);
            a1 = fma(Mb[(size_t)(c + 1) * kSolveBlk], ts[c + 1], a1);
        }
        if (c < w) a0 = fma(Mb[(size_t)c * kSolveBlk], ts[c], a0);
    }
    if (r < w) y[(size_t)ib * n + k0 + r] = a0 + a1;
}

static void launch_chain(cudaStream_t s, int dir, int B, int n, const ChainLayout& cl, const double* W, const double* vin,
                         double* vout, const int* done) {
    const int nw = cl.nbmax + 1;
    int ch = kChainCHMax;
    while (ch > 64 && chain_smem_bytes(nw, ch) > (size_t)227 * 1024) ch -= 32;
    const size_t smem = chain_smem_bytes(nw, ch);
    ensure_dyn_smem((const void*)k_band_chain, smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(B * kChainCta));
    cfg.blockDim = dim3(kChainThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kChainCta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    note(cudaLaunchKernelEx(&cfg, k_band_chain, dir, n, cl.nblk, cl.pw, cl.pwt, nw, ch, W, vin, vout, done));
    PDEOP_COUNT(1);
}

void be_chol_solve(stream_t st, const LevelDev& L, int B, const double* Lf, const double* Linv, const double* rhs,
                   double* out, double* work, const int* done, bool use_chain) {
    cudaStream_t s = (cudaStream_t)st;
    const int n = L.M * L.G;
    const int bw = L.bw;
    const size_t strideL = (size_t)n * n;
    const int nblk = (n + kSolveBlk - 1) / kSolveBlk;
    const double* LinvT = Linv + (size_t)B * nblk * kSolveBlk * kSolveBlk;   // transposed inverse blocks
    double* rb = work;                  // right-hand side in band ordering; reused as block temporary
    double* y = work + (size_t)B * n;   // solution in band ordering
    k_to_band<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, s>>>(L, rhs, rb, done);
    PDEOP_COUNT(1);
    if (n <= kSmallN) {   // one warp per instance (k_chol_small left the factor in Lf)
        k_solve_small<<<cdiv(B, 8), 256, 0, s>>>(n, B, Lf, rb, y, done);
        k_from_band<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, s>>>(L, y, out, done);
        PDEOP_COUNT(2);
        PDEOP_LAUNCH_CHECK();
        return;
    }
    const ChainLayout cl = be_chain_layout(n, bw);
    if (cl.use && use_chain) {
        // L^-1 = Wt^-1 D^-1, L^-T = D^-T Wt^-T: block-diagonal product, two chains, block-diagonal product
        double *Wc, *WTc;
        chain_arrays(cl, B, n, const_cast<double*>(Linv), &Wc, &WTc);
        k_blk_mv_all<<<dim3(nblk, B), kSolveBlk, 0, s>>>(n, LinvT, nblk, rb, y, 0, done);
        launch_chain(s, 0, B, n, cl, Wc, y, rb, done);
        launch_chain(s, 1, B, n, cl, WTc, rb, y, done);
        k_blk_mv_all<<<dim3(nblk, B), kSolveBlk, 0, s>>>(n, Linv, nblk, y, rb, 1, done);
        k_from_band<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, s>>>(L, rb, out, done);
        PDEOP_COUNT(3);
        PDEOP_LAUNCH_CHECK();
        return;
    }
    // forward: y = L^-1 rb.  Per block row: t = rb_k - L[k, band] y (GEMV, in place in rb), y_k = Linv_kk t
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * kSolveBlk;
        const int w = k0 + kSolveBlk < n ? kSolveBlk : n - k0;
        if (k0 > 0) {
            int c_lo = k0 - bw;
            if (c_lo < 0) c_lo = 0;
            c_lo &= ~31;
            k_fwd_gemv<<<dim3(cdiv(w, 8), B), 256, 0, s>>>(n, Lf, strideL, k0, w, c_lo, rb, y, rb, done);
            PDEOP_COUNT(1);
        }
        k_blk_mv<<<dim3(cdiv(w, 8), B), 256, 0, s>>>(n, Linv, nblk, kb, w, rb, y, 0, done);
        PDEOP_COUNT(1);
    }
    // backward: y <- L^-T y.  Per block row (descending): z_k = Linv_kk^T y_k (into rb), then the update kernel
    // stores z_k into y and subtracts L[k rows, c]^T z_k from the band columns left of the block
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int k0 = kb * kSolveBlk;
        const int w = k0 + kSolveBlk < n ? kSolveBlk : n - k0;
        k_blk_mv<<<dim3(cdiv(w, 8), B), 256, 0, s>>>(n, LinvT, nblk, kb, w, y, rb, 1, done);
        PDEOP_COUNT(1);
        int c_lo = k0 - bw;
        if (c_lo < 0) c_lo = 0;
        c_lo &= ~31;
        // column tiles over the band columns [c_lo,k0) (rounded up to whole tiles) plus tiles copying the block
        const int band_tiles = (int)cdiv(k0 - c_lo, kUpdCols);
        k_bwd_update<<<dim3(band_tiles + cdiv(w, kUpdCols), B), 256, 0, s>>>(n, Lf, strideL, k0, w,
                                                                            k0 - band_tiles * kUpdCols, c_lo, rb, y, done);
        PDEOP_COUNT(1);
    }
    k_from_band<<<dim3(cdiv(L.G, kThreads), B), kThreads, 0, s>>>(L, y, out, done);
    PDEOP_COUNT(1);
    PDEOP_LAUNCH_CHECK();
}

}  // namespace pdeop
