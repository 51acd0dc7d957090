// pdeop -- line-marching wavefront Gauss-Seidel: the CUDA kernel around the bodies of pdeop_gs_line.h
// (device only; included by pdeop_cuda.cu after pdeop_gs_fast.cuh, whose cp.async helpers it uses).
#pragma once
#include "pdeop_gs_line.h"

namespace pdeop {

struct LineDevIO {
    unsigned long long pol;   // L2 evict-first access policy for the once-per-sweep streams
    __device__ __forceinline__ void init() {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    }
    __device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
    // streamed once per sweep (b, coefficients, reciprocal diagonals): no L1 allocation, first out of L2
    __device__ __forceinline__ double ldstream(const double* p) {
        double v;
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
        return v;
    }
    // values of the thread's own line: no L1 allocation (the small L1 left beside the rings holds the index tables)
    __device__ __forceinline__ double ldown(const double* p) {
        double v;
        asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void cp8(double* smem_dst, const double* gsrc) { cp_async8(smem_dst, gsrc); }
    __device__ __forceinline__ void prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
};

__device__ __forceinline__ void line_bar_init(uint32_t a, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void line_bar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
// bounded spin (try_wait suspends for a short, implementation-defined time; a long suspend-time hint was measured
// slower: the wake-up sits on the critical path of every step); a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void line_bar_wait(uint32_t a, uint32_t parity) {
    for (int spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1 << 24)) __trap();
        __nanosleep(20);      // fewer polls: a polling warp takes issue slots from the warps it waits for
    }
}
// progress counters of the CTAs of an instance (global memory): release / acquire at GPU scope
__device__ __forceinline__ void line_publish(long long* p, long long v) {
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long line_peek(const long long* p) {
    long long v;
    asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long line_await(const long long* p, long long seen, long long need) {
    if (seen >= need) return seen;
    for (long long spin = 0;; ++spin) {
        seen = line_peek(p);
        if (seen >= need) return seen;
        if (spin > (1LL << 26)) __trap();
        __nanosleep(32);
    }
}

constexpr int kLineFlagWords = 8;   // progress counters per instance (C <= 8)

// One CTA per group of `RPC` grid rows of one instance; grid = B * C CTAs, launched as clusters of C (co-scheduling).
// flags: B * kLineFlagWords 64-bit counters, zero on entry.
template <int D, int PS>
__global__ void __launch_bounds__(kLineMaxThreads, 1) k_gs_line(LevelDev L, LineGeom g, const double* __restrict__ T,
                                                              LineStreams S, double* x, long long* flags,
                                                              const int* done) {
    if (done && *done) return;
    extern __shared__ __align__(16) unsigned char line_smem[];
    __shared__ unsigned long long line_bar;
    double* sm = reinterpret_cast<double*>(line_smem);
    const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    const int ib = (int)blockIdx.x / g.C, cta = (int)blockIdx.x - ib * g.C;
    const size_t o = (size_t)ib * L.M * L.G;
    const double* Ti = T + (size_t)ib * L.D * kTabEntries * kTabPitch;
    line_smem_fill<D, PS>(L, g, Ti, cta, sm, tid, nthr);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&line_bar);
    if (tid == 0) line_bar_init(bar, nthr >> 5);
    __syncthreads();
    LineCtx<D> c;
    line_init<D>(L, g, cta, tid, c);
    LineDevIO io;
    io.init();
    long long* fl = flags + (size_t)ib * kLineFlagWords;
    long long seen_lo = 0, seen_hi = 0;
    for (int n = 0; n < g.NS; ++n) {
        line_fin<D, PS>(L, g, sm, x + o, c);
        cp_async_wait_all();          // the look-ahead copies of the previous step have landed
        __syncwarp();                 // one arrival per warp: the lanes' writes are ordered before lane 0's release
        if ((tid & 31) == 0) line_bar_arrive(bar);
        if (g.C > 1 && tid < 32) {
            if (tid == 0) {
                // every thread's A(n-1) is complete (this thread passed WAIT(n-1)): publish n finished steps, then make
                // sure B(n+2) may run: the previous CTA has published n, the next one n + 11 - N2
                line_publish(fl + cta, (long long)n);
                if (cta > 0) seen_lo = line_await(fl + cta - 1, seen_lo, (long long)n + 3 - kLineDelta);
                if (cta + 1 < g.C) seen_hi = line_await(fl + cta + 1, seen_hi, (long long)n + 8 + kLineDelta - g.N2);
            }
            __syncwarp();
        }
        line_pre<D, PS>(L, g, sm, S, (unsigned)o, c, io);
        cp_async_commit();
        line_advance<D>(g, c);
        if ((tid & 31) == 0) line_bar_wait(bar, (uint32_t)(n & 1));
        __syncwarp();
    }
    if (g.C > 1 && tid == 0) line_publish(fl + cta, (long long)g.NS);
}

}  // namespace pdeop
