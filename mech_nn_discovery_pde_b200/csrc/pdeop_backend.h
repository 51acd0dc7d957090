// pdeop -- backend interface between the host-side orchestration (pdeop_solver.cpp) and the kernels.
// The product implements it with CUDA kernels for sm_100a (pdeop_cuda.cu).  tests/emu implements it
// with plain loops over the same per-element bodies so the orchestration and index algebra can be
// checked on a machine without a GPU; that emulator is test infrastructure and is never loaded by
// the product package.
#pragma once
#include <stddef.h>
#include "pdeop_common.h"

namespace pdeop {

typedef void* stream_t;  // cudaStream_t in the CUDA build

const char* be_name();
void* be_alloc(size_t bytes);
void be_free(void* p);
void be_upload(void* dst, const void* src, size_t bytes);
void be_zero(stream_t st, void* p, size_t bytes);
int be_last_error(char* buf, int len);  // 0 if no pending error
// timing events on the launching stream (cudaEvent_t in the CUDA build) and the launch counter
void* be_event_create();
void be_event_destroy(void* ev);
void be_event_record(void* ev, stream_t st);
float be_event_elapsed_ms(void* start, void* stop);  // waits for `stop`
long long be_launch_count();
// device the calling thread is bound to (cudaGetDevice), -1 without one; be_set_device binds it
int be_current_device();
// process defaults of the per-plan kernel-variant switches (environment variables PDEOP_GS_PIPE / PDEOP_CHAIN,
// read when a plan is created; afterwards the plan's own copy is what every call uses)
int be_default_gs_pipe();
bool be_default_chain();

void be_build_tables(stream_t st, const LevelDev& L, int B, const double* cv, const double* fv, const double* bv,
                     double* T);
void be_pack(stream_t st, const LevelDev& L, int B, const double* api, double* wave);
void be_unpack(stream_t st, const LevelDev& L, int B, const double* wave, double* api);
void be_interp(stream_t st, const LevelDev& Li, const LevelDev& Lo, int B, int C, const double* in, double* out,
               int add, const int* done);
void be_atb(stream_t st, const LevelDev& L, int B, const double* coef, const double* rhs_nat, const double* iv_rhs,
            double* atb);
// y = K x (mode 0) or y = b - K x (mode 1)
void be_apply_k(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, const double* x,
                const double* b, double* y, int mode, const int* done);
// nsweeps lexicographic Gauss-Seidel sweeps, in place.  variant 0: production kernel; 1: one launch per
// hyperplane step (test cross-check).  stash: B * stash_stride doubles of scratch for the production kernel
// (stash_stride >= be_gs_stash_doubles(L)).
// gs_pipe: 0 unsplit cluster kernel, 1 software-pipelined kernel, 2 pipelined on the latency-bound levels only
void be_gs(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, const double* dinv,
           const double* b, double* x, double* stash, size_t stash_stride, int nsweeps, const int* done, int variant,
           int gs_pipe);
// per-instance stash size: 8 doubles per point of the busiest step, rounded up to whole rounds of the largest
// cluster (8 CTAs x 512 threads)
inline size_t be_gs_stash_doubles(const LevelDev& L) { return 8 * ((size_t)L.G + 4096); }
// dinv[m][w] = 1 / K[(w,m),(w,m)]: reciprocal diagonal of K, computed once per operator set-up
void be_dinv(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, double* dinv);
// zero what the factorisation and the solves will read of the B dense n x n matrices: the lower band (half-bandwidth bw
// plus the panel/tile slack of the blocked algorithms); nothing else of the array is ever read
void be_zero_dense(stream_t st, int B, int n, int bw, double* Kd);
void be_dense(stream_t st, const LevelDev& L, int B, const double* T, const double* coef, double* Kd);
// ---- dense coarsest level: storage of the factor's inverse diagonal blocks and (chain solver) scaled band ----
// Linv area (doubles): [inverse diagonal blocks | their transposes | pad | Wc | WTc], see be_cholesky.
// The chain solver (persistent cluster kernel, pdeop_cuda.cu) needs 16-byte aligned rows: n even, and at least four
// block rows to be worth it.
struct ChainLayout {
    int use;       // 1: chain solver applies to (n, bw)
    int nblk;      // block rows of kSolveBlk
    int pw;        // Wc row pitch  = band columns per row, rounded up to 32
    int nbmax;     // block columns a block row's band touches
    int pwt;       // WTc row pitch = nbmax * kSolveBlk
};
inline ChainLayout be_chain_layout(int n, int bw) {
    ChainLayout c;
    c.nblk = (n + kSolveBlk - 1) / kSolveBlk;
    c.pw = (bw + 31) & ~31;
    c.nbmax = (c.pw + kSolveBlk - 1) / kSolveBlk;
    c.pwt = c.nbmax * kSolveBlk;
    c.use = (n % 2 == 0) && c.nblk >= 8 && c.nbmax + 1 <= 32;   // vector window of nbmax+1 blocks must fit shared memory
    return c;
}
// use_chain (decided ONCE per plan: layout and solver choice come from the same snapshot): be_cholesky builds and
// be_chol_solve reads the compact scaled band in the Linv area; the dense n x n array is then only needed during
// operator set-up
inline size_t be_chol_linv_doubles(int B, int n, int bw, bool use_chain) {
    const ChainLayout c = be_chain_layout(n, bw);
    size_t tot = 2 * (size_t)B * c.nblk * kSolveBlk * kSolveBlk;
    if (c.use && use_chain) tot += 64 + (size_t)B * n * ((size_t)c.pw + c.pwt);
    return tot;
}

// in-place lower Cholesky of B dense n x n matrices; state->chol_info set on a non-positive pivot
// (bw = half-bandwidth of the matrix in the dense ordering; nothing outside the band is touched)
// Linv: be_chol_linv_doubles(B, n, bw) doubles receiving the inverses of the diagonal blocks of L, their transposes
// and, when the chain solver applies, the block-scaled band W = blockdiag(L_kk)^-1 L in two compact layouts
// aux: per-plan helper objects of the backend (be_aux_create), or nullptr: with them the panel factorisation of outer
// panel k+1 runs on a second, higher-priority stream while the main stream applies panel k's trailing update to the
// columns beyond panel k+1 (look-ahead); joined before the call returns control of the data to `st`
void be_cholesky(stream_t st, int B, int n, int bw, double* Kd, double* Linv, FgmresState* state, bool use_chain,
                 void* aux);
// per-plan helper objects (CUDA build: one high-priority side stream and two events; emulator: nothing)
void* be_aux_create();
void be_aux_destroy(void* aux);
// out = (L L^T)^-1 rhs for the dense system of level L (n = M*G); rhs/out in wave/planar layout, the
// factor in band ordering; work: 2*B*n doubles
void be_chol_solve(stream_t st, const LevelDev& L, int B, const double* Lf, const double* Linv, const double* rhs,
                   double* out, double* work, const int* done, bool use_chain);
void be_grads(stream_t st, const LevelDev& L, int B, const double* coef, const double* rhs_nat, const double* cv,
              const double* fv, const double* bv, const double* x, const double* dz, double* d_coeffs,
              double* d_rhs, double* d_iv, double* d_cv, double* d_fv, double* d_bv);

// ---- converged mode: per-instance PCG with a symmetric V-cycle (cg.py:51-147 semantics) ----
// All vectors are (B, n) contiguous; s_* are device arrays of B doubles.
// out[b] (+)= sum_i a[b,i] * c[b,i]   (out must be zeroed by the caller: be_zero)
void be_bdot(stream_t st, size_t n, int B, const double* a, const double* c, double* out, const int* done);
// alpha_b = active_b ? rz_b / pAp_b (0 when not finite) : 0;  x += alpha p;  r -= alpha Ap;  rr[b] += sum r^2
void be_pcg_xr(stream_t st, size_t n, int B, double* x, double* r, const double* p, const double* Ap, const double* rz,
               const double* pAp, const double* active, double* rr, const int* done);
// beta_b = active_b ? rz_new_b / rz_b (0 when not finite) : 0;  p = z + beta p
void be_pcg_p(stream_t st, size_t n, int B, double* p, const double* z, const double* rz_new, const double* rz,
              const double* active, const int* done);
// per-instance bookkeeping after an iteration: active_b &= (sqrt(rr_b) > rtol * bnorm_b); rz = rz_new; zero rz_new,
// pAp, rr; iters += 1; relres = max_b sqrt(rr_b)/bnorm_b; done = no instance active.  first != 0: initialise (active_b =
// bnorm_b > 0, iters = 0) from bb[b] = ||b_b||^2 stored in rr
void be_pcg_scalars(stream_t st, int B, double* active, double* rz, double* rz_new, double* pAp, double* rr,
                    double* bnorm, double rtol, FgmresState* state, int first);
// polynomial smoother step: d = c1 d + w_b * dinv .* r;  x += d, with w_b = c2 (mode 0), c2 / lam[b] (mode 1:
// Chebyshev coefficients are in units of lambda_max) or min(c2, 1.8 / lam[b]) (mode 2: weighted Jacobi kept convergent)
void be_poly_update(stream_t st, size_t n, int B, double* x, double* d, const double* r, const double* dinv, double c1,
                    double c2, const double* lam, int mode, const int* done);
void be_copy(stream_t st, void* dst, const void* src, size_t bytes);
// restriction by the TRANSPOSE of the linear prolongation (full-weighting-like, R = P^T): out (coarse) = P^T in (fine)
void be_restrict_t(stream_t st, const LevelDev& Lf, const LevelDev& Lc, int B, int C, const double* in, double* out,
                   const int* done);
// power iteration step for lambda_max(D^-1 K): v <- dinv .* Kv / ||dinv .* Kv||_b,  lam[b] = ||dinv .* Kv||_b (||v||_b = 1)
void be_power_step(stream_t st, size_t n, int B, double* v, const double* Kv, const double* dinv, double* lam, double* tmpB);

// FGMRES vector steps on flat vectors of length n (whole local batch: global norms, fgmres.py:76,128,158)
void be_fg_begin(stream_t st, size_t n, const double* b, double* x, FgmresState* s);
void be_fg_resnorm(stream_t st, size_t n, const double* r, FgmresState* s, int maxiter, double atol);
void be_fg_first(stream_t st, size_t n, const double* r, double* V0, FgmresState* s);
void be_fg_cgs(stream_t st, size_t n, int j, int restart, double* V, double* w, FgmresState* s);
void be_fg_update(stream_t st, size_t n, int restart, const double* Z, double* x, FgmresState* s);
void be_fg_info(stream_t st, const FgmresState* s, double* info4);
void be_fg_hess(stream_t st, const FgmresState* s, int restart, double* hess_out);
void be_state_reset(stream_t st, FgmresState* s);

}  // namespace pdeop
