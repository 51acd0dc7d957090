"""Solver knobs of the PDE layer.  Attribute names mirror the reference's ``config.PDEConfig``
(config.py:13-29) and are read at call time, so assigning to the class attributes (as the reference's
notebooks do) changes the next forward/backward call."""


class PDEConfig:
    # multigrid options (config.py:13-26)
    mg_gauss_seidel_steps_pre = 5
    mg_gauss_seidel_steps_post = 5

    mg_steps_forward = 1
    mg_steps_backward = 1

    mg_fgmres_max_iter_forward = 40
    mg_fgmres_restarts_forward = 10

    mg_fgmres_max_iter_backward = 40
    mg_fgmres_restarts_backward = 10

    # accepted for compatibility; the reference never reads it on the live path (config.py:29)
    jacobi_w = 0.4

    # --- additions of this implementation -------------------------------------------------------
    # absolute tolerance on the batch-global residual (fgmres.py:22, atol=1e-5)
    mg_fgmres_atol = 1e-5
    # The reference's add_pad allocates its buffer with the default dtype, so d(loss)/d(rhs) is rounded
    # to fp32 even on the fp64 path (lp_pde_central_diff.py:1634).  False: keep fp64 (default).
    rhs_grad_fp32_quirk = False
    # Read the Cholesky status back after each forward and raise torch.linalg.LinAlgError like
    # cholesky_ex(check_errors=True) (multigrid.py:439).  Costs one host sync per forward.
    check_factorization = True
    # 0: production wavefront Gauss-Seidel kernel; 1: one launch per hyperplane (debug cross-check)
    gs_variant = 0

    # --- converged mode (north-star subsystems (b)/(c)) --------------------
    # "reference": the reference's live algorithm -- FGMRES(10) capped at 40 iterations with a Gauss-Seidel V-cycle, the
    #   batch sharing one Krylov space (iterate parity with the reference, usually far from converged);
    # "converged": per-instance PCG to a relative tolerance with a symmetric V-cycle (polynomial smoother, R = P^T),
    #   validated against the exact least-squares solution (solver/cg.py:51-147 semantics).
    solver_mode = "reference"
    mg_pcg_rtol = 1e-8
    mg_pcg_max_iter = 1000
    mg_smoother = "chebyshev"      # or "jacobi" (weighted Jacobi with jacobi_w, solver/multigrid.py:407-416)
    mg_smoother_sweeps = 8
    # power iterations for lambda_max(D^-1 K): the estimate converges from below, and an under-estimate makes the
    # Chebyshev smoother amplify the top of the spectrum (12 iterations: 3-D Ginzburg-Landau cases diverge; measured)
    mg_power_iters = 60
    mg_cheb_ratio = 30.0           # Chebyshev smoothing interval [1.1 lambda_max / ratio, 1.1 lambda_max]
