/* pdeop.h -- C ABI of libpdeop.so, the B200-native (sm_100a) differentiable PDE-layer solve.
 *
 * This is the drop-in boundary for ONE path of alpz/mech-nn-discovery-pde: constraint assembly ->
 * normal-equation solve (multigrid-preconditioned FGMRES, or dense Cholesky) -> implicit-
 * differentiation backward.  The reference implements that path in Python on top of torch.sparse,
 * CuPy/cuSPARSE and cuSOLVER; each entry point below names the reference interface it replaces
 * (paths relative to the reference repo).  Plain pointers and sizes only: no torch types.
 *
 * Conventions
 *  - all data pointers are DEVICE pointers to contiguous fp64 unless stated otherwise;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous and stream-ordered, no
 *    call synchronises the host, allocates device memory or creates streams (plan_create excepted);
 *  - no mutable process-global state: kernel-variant switches and instrumentation belong to a plan; plans on
 *    different devices (or driven by different host threads) do not interact;
 *  - the caller owns every buffer (sizes from pdeop_plan_query); the library owns only the immutable
 *    plan (index tables, level dims);
 *  - return value 0 = ok, nonzero = error, text from pdeop_last_error();
 *  - operator-surface tensors use the reference's layouts: coeffs (B,G,M), rhs (B,G), iv_rhs
 *    (B,n_init), x (B,G*M) with variable index g*M+m (lp_pde_central_diff.py:96-107).
 *  - the per-call derivative-row values are passed per LINE position instead of per nonzero:
 *      cv (B,Ntot,2,6)  central rows     (lp_pde_central_diff.py:1300-1547)
 *      fv (B,Ftot,4)    forward rows     (:1550-1581)
 *      bv (B,Ftot,4)    backward rows    (:1583-1615)
 *    with Ntot = sum_c n_c, Ftot = sum_c (n_c-1), coordinate-major.  The reference's per-nonzero value
 *    vector is exactly these tables expanded over the remaining grid directions.
 */
#ifndef PDEOP_H
#define PDEOP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdeop_plan pdeop_plan;

/* Solver knobs; names follow config.py:13-29 (PDEConfig). */
typedef struct pdeop_solver_cfg {
    int gs_pre;        /* mg_gauss_seidel_steps_pre  */
    int gs_post;       /* mg_gauss_seidel_steps_post */
    int mg_steps;      /* mg_steps_forward / mg_steps_backward */
    int max_iter;      /* mg_fgmres_max_iter_*  */
    int restart;       /* mg_fgmres_restarts_*  */
    double atol;       /* fgmres.py:22 atol=1e-5 (absolute, on the batch-global residual) */
    int gs_variant;    /* 0 = production wavefront kernel, 1 = one launch per hyperplane (cross-check) */
} pdeop_solver_cfg;

enum pdeop_query {
    PDEOP_Q_NLEVELS = 0,
    PDEOP_Q_G = 1,            /* grid points of `level` */
    PDEOP_Q_M = 2,
    PDEOP_Q_N_EQ = 3,
    PDEOP_Q_N_INIT = 4,
    PDEOP_Q_NTOT = 5,
    PDEOP_Q_FTOT = 6,
    PDEOP_Q_PERSIST_BYTES = 7,  /* tables + coefficients of all levels + coarsest dense factor */
    PDEOP_Q_SCRATCH_BYTES = 8,  /* vectors; `level` argument carries the FGMRES restart length */
    PDEOP_Q_DIM0 = 16           /* +c: extent c of `level` */
};

enum pdeop_stage {
    PDEOP_STAGE_APPLY_K = 0,    /* out = K_level in1                 (multigrid.py:393-397 mult_AtA)        */
    PDEOP_STAGE_GS = 1,         /* out = GS^count(b=in1, x0=in2)      (multigrid.py:399-405 smooth_gs)       */
    PDEOP_STAGE_RESTRICT = 2,   /* level -> level+1                  (multigrid.py:340-363)                 */
    PDEOP_STAGE_PROLONG = 3,    /* level -> level-1                  (multigrid.py:366-391)                 */
    PDEOP_STAGE_VCYCLE = 4,     /* out = V-cycle(b=in1) from zero    (multigrid.py:490-498)                 */
    PDEOP_STAGE_COARSE_SOLVE = 5, /* coarsest Cholesky solve         (multigrid.py:442-450)                 */
    PDEOP_STAGE_ATB = 6         /* out = A^T b at level 0; in1 = rhs (B,G), in2 = iv_rhs  (multigrid.py:230)  */
};

/* Replaces PDESYSLP.build_constraints + MultigridSolver.__init__ (lp_pde_central_diff.py:1063-1139,
 * multigrid.py:46-112): closed-form index tables instead of per-grid-point Python loops.
 *   d, dims[d]     problem dimension (1..3) and grid extents
 *   order          total derivative order (2)
 *   n_grid         number of multigrid levels (1 = dense layer)
 *   downsample_first  halve the first axis too (multigrid.py:99-102)
 *   iv_desc        n_grid * n_iv * (1+2d) ints: per level and initial/boundary spec [mi, begin[d], end[d]]
 *                  (the iv lambdas of init_index_mi_list evaluated at that level's dims, multigrid.py:296-306) */
int pdeop_plan_create(int d, const int* dims, int order, int batch, int n_grid, int downsample_first, int n_iv,
                      const int* iv_desc, pdeop_plan** out);
/* Same with creation-time options (NULL = defaults).  The plan records the CUDA device that is current when it is
 * created; its tables live there and every later call must be made with that device current (checked). */
typedef struct pdeop_plan_opts {
    int evolution;      /* PDESYSLP(evolution=...) equation rows (lp_pde_central_diff.py:756-759); 0 default */
    int chain;          /* coarsest triangular solves: -1 default (environment PDEOP_CHAIN, else on), 0 one launch per
                           block row, 1 persistent chain kernels where they apply.  Fixed for the plan's lifetime: the
                           persist/scratch layout depends on it. */
    int gs_pipe;        /* -1 default (environment PDEOP_GS_PIPE, else 2); see pdeop_plan_set_tuning key 0 */
    int reserved[5];    /* must be 0 */
} pdeop_plan_opts;
int pdeop_plan_create_ex(int d, const int* dims, int order, int batch, int n_grid, int downsample_first, int n_iv,
                         const int* iv_desc, const pdeop_plan_opts* opts, pdeop_plan** out);
void pdeop_plan_destroy(pdeop_plan* plan);
int pdeop_plan_query(const pdeop_plan* plan, int what, int level, long long* out);
const char* pdeop_last_error(void);
const char* pdeop_backend_name(void);

/* Replaces QPFunctionFn.forward of solver/qp_dual_sparse_multigrid_normal_kkt.py:25-79
 * (fill_coarse_grids, make_AtA, make_coarse_AtA_matrices, factor_coarsest, fgmres_matvec, lam).
 * cv/fv/bv: arrays of n_grid device pointers (one per level).  x_out (B,G*M).  info_out: 4 doubles
 * {iters, r_norm, b_norm, chol_info} written on the stream (fgmres.py:182 returns (iters, r_norm)). */
int pdeop_mg_forward(pdeop_plan* plan, const pdeop_solver_cfg* cfg, const double* coeffs, const double* rhs,
                     const double* iv_rhs, const double* const* cv, const double* const* fv,
                     const double* const* bv, void* persist, void* scratch, double* x_out, double* info_out,
                     void* stream);

/* Replaces QPFunctionFn.backward of solver/qp_dual_sparse_multigrid_normal_kkt.py:81-162.
 * Reuses the operator and factor left in `persist` by the forward call.  Outputs: d_coeffs (B,G,M)
 * [= dA on the equation rows], d_rhs (B,G) [add_pad'ed, fp64], d_iv_rhs (B,n_init), and the
 * gradient w.r.t. the level-0 line values d_cv/d_fv/d_bv [= dD summed over expanded directions]. */
int pdeop_mg_backward(pdeop_plan* plan, const pdeop_solver_cfg* cfg, const double* rhs, const double* cv0,
                      const double* fv0, const double* bv0, void* persist, void* scratch, const double* x,
                      const double* grad_x, double* d_coeffs, double* d_rhs, double* d_iv_rhs, double* d_cv,
                      double* d_fv, double* d_bv, double* info_out, void* stream);

/* Replaces solver/qp_dual_dense_normal_kkt.py:23-56 / :58-118 (dense A^T A, cholesky_ex, cholesky_solve). */
int pdeop_dense_forward(pdeop_plan* plan, const double* coeffs, const double* rhs, const double* iv_rhs,
                        const double* cv0, const double* fv0, const double* bv0, void* persist, void* scratch,
                        double* x_out, double* info_out, void* stream);
int pdeop_dense_backward(pdeop_plan* plan, const double* rhs, const double* cv0, const double* fv0,
                         const double* bv0, void* persist, void* scratch, const double* x, const double* grad_x,
                         double* d_coeffs, double* d_rhs, double* d_iv_rhs, double* d_cv, double* d_fv,
                         double* d_bv, double* info_out, void* stream);

/* ---- converged mode -----------------------------------------------------------
 * The reference's live solver stops at its iteration cap far from convergence.  This mode
 * solves the same normal equations A^T A x = A^T b to a per-instance RELATIVE tolerance: preconditioned conjugate
 * gradients with per-instance dot products, step lengths and convergence masks (the semantics of
 * solver/cg.py:51-147 cg_matvec: converged instances stop moving, non-finite alpha/beta are zeroed), preconditioned
 * by one SYMMETRIC V-cycle: polynomial smoother on D^-1 K -- weighted Jacobi (solver/multigrid.py:407-416
 * smooth_jacobi, weight clamped to 1.8 / lambda_max) or Chebyshev over [1.1 lambda_max/cheb_ratio, 1.1 lambda_max] with
 * lambda_max(D^-1 K) estimated per instance and level by power iteration --, restriction by the transpose of the
 * linear prolongation, rediscretised coarse operators, dense Cholesky on the coarsest level.
 * info_out: {iterations, max_b ||r_b||/||b_b||, 0, chol_info}. */
typedef struct pdeop_pcg_cfg {
    int max_iter;      /* PCG iteration cap */
    double rtol;       /* per-instance relative residual tolerance */
    int smoother;      /* 0 weighted Jacobi, 1 Chebyshev */
    int sweeps;        /* smoother steps (Jacobi) / polynomial degree (Chebyshev), each side of the coarse correction */
    double jacobi_w;   /* config.jacobi_w (config.py:29) */
    int power_iters;   /* power-iteration steps for lambda_max(D^-1 K) */
    double cheb_ratio; /* Chebyshev smoothing interval [1.1 lambda_max / cheb_ratio, 1.1 lambda_max]; 30 if <= 1 */
} pdeop_pcg_cfg;
int pdeop_mg_forward_converged(pdeop_plan* plan, const pdeop_pcg_cfg* cfg, const double* coeffs, const double* rhs,
                               const double* iv_rhs, const double* const* cv, const double* const* fv,
                               const double* const* bv, void* persist, void* scratch, double* x_out, double* info_out,
                               void* stream);
int pdeop_mg_backward_converged(pdeop_plan* plan, const pdeop_pcg_cfg* cfg, const double* rhs, const double* cv0,
                                const double* fv0, const double* bv0, void* persist, void* scratch, const double* x,
                                const double* grad_x, double* d_coeffs, double* d_rhs, double* d_iv_rhs, double* d_cv,
                                double* d_fv, double* d_bv, double* info_out, void* stream);

/* Operator set-up only (levels' tables, coarse coefficients, coarsest factor): the part of
 * QPFunctionFn.forward before the Krylov solve.  Needed before pdeop_stage. */
int pdeop_mg_setup(pdeop_plan* plan, const double* coeffs, const double* const* cv, const double* const* fv,
                   const double* const* bv, void* persist, void* scratch, double* info_out, void* stream);

/* Single multigrid building blocks on vectors in the reference's variable order (B, G_level*M); used by
 * the parity tests, one per reference method (see enum pdeop_stage). */
int pdeop_stage(pdeop_plan* plan, const pdeop_solver_cfg* cfg, int stage, int level, int count, const double* in1,
                const double* in2, double* out, void* persist, void* scratch, void* stream);

/* FGMRES on a caller-supplied right-hand side b (B,G*M) with the operator in `persist`
 * (fgmres.py:21 fgmres_matvec); x_out (B,G*M), info_out as above; hess_out: (restart+1)*restart doubles
 * holding the Hessenberg matrix of the last restart cycle (may be NULL). */
int pdeop_fgmres(pdeop_plan* plan, const pdeop_solver_cfg* cfg, int back, const double* b, double* x_out,
                 double* info_out, double* hess_out, void* persist, void* scratch, void* stream);

/* Instrumentation (bench.py), PER PLAN: device time per kernel category measured with events recorded on the
 * launching stream around each kernel group while enabled.
 * Categories: 0 GS fine level, 1 GS coarser levels, 2 K-apply/residual fine, 3 K-apply coarser, 4 grid
 * transfer, 5 coarsest triangular solves, 6 coarsest factorisation, 7 Krylov vector kernels, 8 set-up,
 * 9 gradients, 10 layout conversion.  pdeop_plan_profile_collect waits for the recorded events and returns the
 * number of categories.  pdeop_launch_count: kernels launched by the library so far (process-wide statistic). */
int pdeop_plan_profile_enable(pdeop_plan* plan, int on);
int pdeop_plan_profile_collect(pdeop_plan* plan, double* ms_per_category, long long* count_per_category, int ncat);
long long pdeop_launch_count(void);
/* Kernel-variant switch of ONE plan, for A/B tests and profiling (no reference counterpart):
 *   key 0 "gs_pipe": 0 unsplit cluster Gauss-Seidel kernel on every level, 1 software-pipelined kernel on every
 *                    level, 2 (default) staged kernel (cp.async operand staging, all gathers of a point in flight) on
 *                    the latency-bound 3-D levels, 3 staged kernel wherever it fits, 4 software-pipelined kernel on
 *                    the latency-bound levels (round-1 default), 5 line-marching kernel (a thread owns a grid line,
 *                    neighbours from shared-memory rings filled by cp.async) wherever it fits, mode 2 elsewhere
 *   key 1 "chain"  : read-only (pdeop_plan_get_tuning); chosen at creation, see pdeop_plan_opts.chain
 * Returns 0, or nonzero with pdeop_last_error(). */
int pdeop_plan_set_tuning(pdeop_plan* plan, int key, int value);
int pdeop_plan_get_tuning(const pdeop_plan* plan, int key, int* value);

/* ---- callers either side of the solve -------------------------------------------------
 * f1, fused coefficient builder.  Replaces the basis-functions x learned-parameters assembly in front of the layer
 * (discovery/ginzburg_landau.py:354-374, discovery/burgers_dparam_viscous.py:261-279, discovery/kamani.py:252-271):
 *     out[p,o] = c0[o] + sum_{k<NP, pair_out[k]=o} w[k] * prod_{f<F} pw(fields[f][p], expo[t,f], kind[t,f]),  t = pair_term[k]
 * with p over the npts = B*G grid points, o < M written to coeffs (npts,M) and o = M to rhs (npts).
 * kind 0: integer power of the signed value, 1: |value|^expo (real exponent, differentiable), 2: factor absent.
 * pair_out/pair_term/kind/c0 are HOST arrays (NP<=32, NT<=16, F<=3, M<=7); expo (NT*F), w (NP), fields and outputs
 * are device pointers (`fields`, `d_fields`: host arrays of F device pointers): learned exponents never visit the host.  The backward produces, in one pass, d_w (NP),
 * d_expo (NT*F, nonzero for kind 1) and d_fields[f] (npts; NULL entries skipped). */
int pdeop_coeff_forward(long long npts, int M, int F, int NT, int NP, const int* pair_out, const int* pair_term,
                        const int* kind, const double* expo, const double* c0, const double* const* fields,
                        const double* w, double* coeffs, double* rhs, void* stream);
int pdeop_coeff_backward(long long npts, int M, int F, int NT, int NP, const int* pair_out, const int* pair_term,
                         const int* kind, const double* expo, const double* c0, const double* const* fields,
                         const double* w, const double* d_coeffs, const double* d_rhs, double* d_w, double* d_expo,
                         double* const* d_fields, void* stream);
/* f3, fused data-loss epilogue (discovery/ginzburg_landau.py:486-510): loss[0] = mean |u0 - target|^p (p = 1 or 2)
 * and grad = d loss / d u0 (may be NULL), one kernel. */
int pdeop_loss_forward(long long n, const double* u0, const double* target, int p, double* grad, double* loss,
                       void* stream);
const char* pdeop_coeff_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PDEOP_H */
