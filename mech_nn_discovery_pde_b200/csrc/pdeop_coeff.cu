// pdeop -- fused coefficient builder and fused loss epilogue (row f3).
//
// The step right before the PDE layer in every discovery model is "basis functions of the data fields x learned
// parameters -> coeffs (B,G,M), rhs (B,G)": discovery/ginzburg_landau.py:354-374 (monomials of two fields),
// discovery/burgers_dparam_viscous.py:261-279 (powers of one field), discovery/kamani.py:252-271 (power laws
// |s|^e with LEARNED exponents).  The reference does it with a dozen elementwise kernels, a torch.zeros and
// strided channel writes per call.  Here one kernel writes coeffs and rhs in the operator-surface layout, and one
// kernel produces every gradient (weights, exponents, fields) in a single pass.
//
//   out[p, o] = c0[o] + sum_{k : out_k = o} w_k * term_{t_k}(p),     o < M -> coeffs[p*M + o],  o == M -> rhs[p]
//   term_t(p) = prod_f pw(field_f[p], expo[t,f], kind[t,f]),  kind 0: integer power of the signed value,
//                                                             kind 1: |value|^expo (real, differentiable exponent),
//                                                             kind 2: factor absent
// The step right after the layer (ginzburg_landau.py:486-510) is an L1 / L2 data loss on u0: pdeop_loss_* fuses
// the difference, the norm and the mean with the gradient that feeds the layer's backward.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "../../include/pdeop.h"

namespace {

constexpr int kMaxPairs = 32;
constexpr int kMaxTerms = 16;
constexpr int kMaxFields = 3;
constexpr int kMaxOut = 8;
constexpr int kThreads = 256;

struct CoeffDesc {
    int M, F, NT, NP;
    int pair_out[kMaxPairs], pair_term[kMaxPairs];
    int kind[kMaxTerms][kMaxFields];
    const double* expo;   // device, (NT, F): learned exponents stay on the device (no host sync per step)
    double c0[kMaxOut];
    const double* fields[kMaxFields];
    double* d_fields[kMaxFields];
};

__device__ __forceinline__ double ipow(double x, int e) {
    double r = 1.0;
    for (int i = 0; i < e; ++i) r *= x;
    return r;
}

__device__ __forceinline__ double factor(double v, int kind, double e) {
    if (kind == 0) return ipow(v, (int)e);
    if (kind == 1) return pow(fabs(v), e);
    return 1.0;
}
// d factor / d v
__device__ __forceinline__ double dfactor(double v, int kind, double e) {
    if (kind == 0) return (int)e == 0 ? 0.0 : e * ipow(v, (int)e - 1);
    if (kind == 1) return v == 0.0 ? 0.0 : e * pow(fabs(v), e - 1.0) * (v > 0.0 ? 1.0 : -1.0);
    return 0.0;
}

__global__ void __launch_bounds__(kThreads) k_coeff_fwd(CoeffDesc d, long long npts, const double* __restrict__ w,
                                                        double* __restrict__ coeffs, double* __restrict__ rhs) {
    __shared__ double ws[kMaxPairs];
    __shared__ double es[kMaxTerms][kMaxFields];
    if (threadIdx.x < d.NP) ws[threadIdx.x] = w[threadIdx.x];
    if (threadIdx.x < kMaxTerms * kMaxFields) {
        const int t = threadIdx.x / kMaxFields, f = threadIdx.x % kMaxFields;
        es[t][f] = (t < d.NT && f < d.F) ? d.expo[t * d.F + f] : 0.0;
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
        double fv[kMaxFields];
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f) fv[f] = f < d.F ? d.fields[f][p] : 0.0;
        double term[kMaxTerms];
#pragma unroll
        for (int t = 0; t < kMaxTerms; ++t) {
            double v = 1.0;
            if (t < d.NT) {
#pragma unroll
                for (int f = 0; f < kMaxFields; ++f)
                    if (f < d.F) v *= factor(fv[f], d.kind[t][f], es[t][f]);
            }
            term[t] = v;
        }
        double out[kMaxOut];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) out[o] = d.c0[o];
#pragma unroll
        for (int k = 0; k < kMaxPairs; ++k) {
            if (k < d.NP) {
                const double tv = term[d.pair_term[k]];
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o)
                    if (o == d.pair_out[k]) out[o] = fma(ws[k], tv, out[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
            if (o < d.M) coeffs[p * d.M + o] = out[o];
            else if (o == d.M) rhs[p] = out[o];
        }
    }
}

// one pass: dw[k] = sum_p dOut[p, out_k] term_{t_k}(p);  dexpo[t,f] (kind 1) = sum_p g_t(p) term_t(p) ln|field_f|;
// d_fields[f][p] = sum_t g_t(p) d term_t / d field_f,   g_t(p) = sum_{k : t_k = t} w_k dOut[p, out_k]
__global__ void __launch_bounds__(kThreads) k_coeff_bwd(CoeffDesc d, long long npts, const double* __restrict__ w,
                                                        const double* __restrict__ d_coeffs,
                                                        const double* __restrict__ d_rhs, double* dw, double* dexpo) {
    __shared__ double ws[kMaxPairs];
    __shared__ double es[kMaxTerms][kMaxFields];
    __shared__ double red[kThreads / 32];
    if (threadIdx.x < d.NP) ws[threadIdx.x] = w[threadIdx.x];
    if (threadIdx.x < kMaxTerms * kMaxFields) {
        const int t = threadIdx.x / kMaxFields, f = threadIdx.x % kMaxFields;
        es[t][f] = (t < d.NT && f < d.F) ? d.expo[t * d.F + f] : 0.0;
    }
    __syncthreads();
    double accw[kMaxPairs];
    double acce[kMaxTerms];   // one learned exponent per term at most is accumulated per field loop below
#pragma unroll
    for (int k = 0; k < kMaxPairs; ++k) accw[k] = 0.0;
    double accE[kMaxTerms][kMaxFields];
#pragma unroll
    for (int t = 0; t < kMaxTerms; ++t)
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f) accE[t][f] = 0.0;
    (void)acce;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
        double fv[kMaxFields];
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f) fv[f] = f < d.F ? d.fields[f][p] : 0.0;
        double dout[kMaxOut];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) dout[o] = o < d.M ? d_coeffs[p * d.M + o] : (o == d.M ? d_rhs[p] : 0.0);
        double fac[kMaxTerms][kMaxFields], term[kMaxTerms], g[kMaxTerms];
#pragma unroll
        for (int t = 0; t < kMaxTerms; ++t) {
            double v = 1.0;
#pragma unroll
            for (int f = 0; f < kMaxFields; ++f) {
                fac[t][f] = (t < d.NT && f < d.F) ? factor(fv[f], d.kind[t][f], es[t][f]) : 1.0;
                v *= fac[t][f];
            }
            term[t] = v;
            g[t] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < kMaxPairs; ++k) {
            if (k < d.NP) {
                double dv = 0.0;
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o)
                    if (o == d.pair_out[k]) dv = dout[o];
                const int t = d.pair_term[k];
#pragma unroll
                for (int tt = 0; tt < kMaxTerms; ++tt)
                    if (tt == t) {
                        accw[k] = fma(dv, term[tt], accw[k]);
                        g[tt] = fma(ws[k], dv, g[tt]);
                    }
            }
        }
        double df[kMaxFields];
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f) df[f] = 0.0;
#pragma unroll
        for (int t = 0; t < kMaxTerms; ++t) {
            if (t < d.NT) {
#pragma unroll
                for (int f = 0; f < kMaxFields; ++f) {
                    if (f < d.F && d.kind[t][f] != 2) {
                        double others = 1.0;
#pragma unroll
                        for (int f2 = 0; f2 < kMaxFields; ++f2)
                            if (f2 != f) others *= fac[t][f2];
                        df[f] = fma(g[t] * others, dfactor(fv[f], d.kind[t][f], es[t][f]), df[f]);
                        if (d.kind[t][f] == 1 && fv[f] != 0.0) accE[t][f] = fma(g[t] * term[t], log(fabs(fv[f])), accE[t][f]);
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f)
            if (f < d.F && d.d_fields[f]) d.d_fields[f][p] = df[f];
    }
    // block reduction of every accumulator, one atomic per block and slot
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    auto block_add = [&](double v, double* dst) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        if (wid == 0) {
            double r = lane < kThreads / 32 ? red[lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == 0) atomicAdd(dst, r);
        }
    };
#pragma unroll
    for (int k = 0; k < kMaxPairs; ++k)
        if (k < d.NP) block_add(accw[k], dw + k);
#pragma unroll
    for (int t = 0; t < kMaxTerms; ++t)
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f)
            if (t < d.NT && f < d.F && d.kind[t][f] == 1) block_add(accE[t][f], dexpo + t * d.F + f);
}

// ---- fused data loss (row f3): loss = mean(|u0 - target|^p), p = 1 or 2; grad = d loss / d u0 --------------------
__global__ void __launch_bounds__(kThreads) k_loss(long long n, const double* __restrict__ u0,
                                                   const double* __restrict__ target, int p, double scale,
                                                   double* __restrict__ grad, double* loss) {
    __shared__ double red[kThreads / 32];
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double dlt = u0[i] - target[i];
        if (p == 1) {
            acc += fabs(dlt);
            if (grad) grad[i] = dlt > 0.0 ? scale : (dlt < 0.0 ? -scale : 0.0);
        } else {
            acc = fma(dlt, dlt, acc);
            if (grad) grad[i] = 2.0 * scale * dlt;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        double r = lane < kThreads / 32 ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) atomicAdd(loss, r * scale);
    }
}

thread_local std::string g_cerr;
int cfail(const char* m) {
    g_cerr = m;
    return 1;
}

int fill_desc(CoeffDesc& d, int M, int F, int NT, int NP, const int* pair_out, const int* pair_term, const int* kind,
              const double* expo, const double* c0, const double* const* fields) {
    if (M < 1 || M + 1 > kMaxOut) return cfail("coeff builder: M out of range");
    if (F < 1 || F > kMaxFields) return cfail("coeff builder: 1..3 fields");
    if (NT < 1 || NT > kMaxTerms) return cfail("coeff builder: 1..16 terms");
    if (NP < 1 || NP > kMaxPairs) return cfail("coeff builder: 1..32 (output, term) pairs");
    d.M = M; d.F = F; d.NT = NT; d.NP = NP;
    for (int k = 0; k < kMaxPairs; ++k) {
        d.pair_out[k] = k < NP ? pair_out[k] : -1;
        d.pair_term[k] = k < NP ? pair_term[k] : -1;
        if (k < NP && (pair_out[k] < 0 || pair_out[k] > M || pair_term[k] < 0 || pair_term[k] >= NT))
            return cfail("coeff builder: pair index out of range");
    }
    for (int t = 0; t < kMaxTerms; ++t)
        for (int f = 0; f < kMaxFields; ++f) {
            d.kind[t][f] = (t < NT && f < F) ? kind[t * F + f] : 2;
        }
    d.expo = expo;
    for (int o = 0; o < kMaxOut; ++o) d.c0[o] = o <= M ? c0[o] : 0.0;
    for (int f = 0; f < kMaxFields; ++f) {
        d.fields[f] = f < F ? fields[f] : nullptr;
        d.d_fields[f] = nullptr;
    }
    return 0;
}

int grid_for(long long n) {
    long long b = (n + kThreads - 1) / kThreads;
    if (b > 148 * 8) b = 148 * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" const char* pdeop_coeff_last_error(void) { return g_cerr.c_str(); }

extern "C" int pdeop_coeff_forward(long long npts, int M, int F, int NT, int NP, const int* pair_out,
                                   const int* pair_term, const int* kind, const double* expo, const double* c0,
                                   const double* const* fields, const double* w, double* coeffs, double* rhs,
                                   void* stream) {
    CoeffDesc d;
    if (fill_desc(d, M, F, NT, NP, pair_out, pair_term, kind, expo, c0, fields)) return 1;
    k_coeff_fwd<<<grid_for(npts), kThreads, 0, (cudaStream_t)stream>>>(d, npts, w, coeffs, rhs);
    return cudaGetLastError() == cudaSuccess ? 0 : cfail("coeff builder: launch failed");
}

extern "C" int pdeop_coeff_backward(long long npts, int M, int F, int NT, int NP, const int* pair_out,
                                    const int* pair_term, const int* kind, const double* expo, const double* c0,
                                    const double* const* fields, const double* w, const double* d_coeffs,
                                    const double* d_rhs, double* d_w, double* d_expo, double* const* d_fields,
                                    void* stream) {
    CoeffDesc d;
    if (fill_desc(d, M, F, NT, NP, pair_out, pair_term, kind, expo, c0, fields)) return 1;
    for (int f = 0; f < F; ++f) d.d_fields[f] = d_fields ? d_fields[f] : nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(d_w, 0, sizeof(double) * NP, s);
    cudaMemsetAsync(d_expo, 0, sizeof(double) * NT * F, s);
    k_coeff_bwd<<<grid_for(npts), kThreads, 0, s>>>(d, npts, w, d_coeffs, d_rhs, d_w, d_expo);
    return cudaGetLastError() == cudaSuccess ? 0 : cfail("coeff builder: launch failed");
}

extern "C" int pdeop_loss_forward(long long n, const double* u0, const double* target, int p, double* grad,
                                  double* loss, void* stream) {
    if (p != 1 && p != 2) return cfail("loss: p must be 1 or 2");
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(loss, 0, sizeof(double), s);
    k_loss<<<grid_for(n), kThreads, 0, s>>>(n, u0, target, p, 1.0 / (double)n, grad, loss);
    return cudaGetLastError() == cudaSuccess ? 0 : cfail("loss: launch failed");
}
