"""Times the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py) on the host CPU  --  TEST INFRASTRUCTURE.

    python oracle/run_ref.py <workload> <batch>        # prints one JSON line

Runs in its own process from a scratch working directory: importing the reference's QP modules creates
./logs/misc/<n>/ (extras/source.py:7-21).  Inputs are bench.py's synthetic inputs for the workload; the layer is
the reference's own MultigridLayer / PDEDenseLayer (solver/multigrid.py:536, solver/pde_layer_dense.py:38) with
`config.py` default knobs; loss = sum(u0^2) like the GPU arm.
"""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(1, ROOT)


def main():
    wname, B = sys.argv[1], int(sys.argv[2])
    os.chdir(tempfile.mkdtemp(prefix="pdeop_ref_"))
    import torch
    import bench
    torch.set_num_threads(os.cpu_count() or 1)
    wl = bench.WORKLOADS[wname]
    iv = bench.IV_LISTS[wl["iv"]]
    t0 = time.time()
    if wl.get("dense"):
        from solver.pde_layer_dense import PDEDenseLayer
        layer = PDEDenseLayer(bs=B, coord_dims=wl["dims"], order=2, n_ind_dim=1, n_iv=1, init_index_mi_list=iv,
                              n_iv_steps=1, double_ret=True, solver_dbl=True)
    else:
        from solver.multigrid import MultigridLayer
        layer = MultigridLayer(bs=B, coord_dims=wl["dims"], order=2, n_ind_dim=1, n_iv=1, n_grid=wl["n_grid"],
                               evolution=False, downsample_first=wl["dsf"], init_index_mi_list=iv, n_iv_steps=1,
                               double_ret=True, solver_dbl=True)
    t_build = time.time() - t0
    inp = bench.synth_inputs(wl, B, 1234)
    n_init = layer.pde.num_added_initial_constraints
    iv_rhs = 0.5 * torch.randn(B, n_init, generator=inp["gen"], dtype=torch.float64)
    theta = bench.theta_init(wl, "cpu")
    t1 = time.time()
    coeffs = bench.assemble_coeffs(wl, inp["coeffs_base"], inp["field"], theta)
    u0, u, _ = layer(coeffs, inp["rhs"], iv_rhs, list(inp["steps"]))
    t_fwd = time.time() - t1
    t2 = time.time()
    (u0 * u0).sum().backward()
    t_bwd = time.time() - t2
    print(json.dumps({"value": B / (t_fwd + t_bwd), "unit": "solves/s", "kind": "reference", "cores": os.cpu_count(),
                      "workload": wl["desc"], "batch": B, "structure_build_s": t_build, "forward_s": t_fwd,
                      "backward_s": t_bwd, "theta_grad": [float(v) for v in theta.grad],
                      "sample": f"unmodified reference (CuPy -> numpy/scipy stubs), {B} instances, full forward+backward; "
                                f"one-time structure build {t_build:.1f}s not counted"}))


if __name__ == "__main__":
    main()
