"""Full-size fixtures of the BENCHMARKED configuration from the pinned oracle port  --  TEST INFRASTRUCTURE ONLY.

The unmodified reference needs ~25 min of pure-Python structure building plus hours of solve at 32x64x64
(SURVEY.md section 3.4), so the fixtures for bench.py's default workload (Ginzburg-Landau 32x64x64, n_grid=4,
downsample_first=False) come from oracle/pde_oracle.py, which tests/test_oracle_vs_golden.py pins against the
reference on every smaller case, including the 3-D three-level downsample_first=False goldens of
oracle/make_golden_r2.py.  Inputs are bench.py's own synthetic inputs (same generator, same seed), so the test
compares exactly what the benchmark runs.

    python oracle/make_golden_port.py tests/golden gl32 1      # workload, batch  (~10 CPU-minutes per instance)
    python oracle/make_golden_port.py tests/golden gl64 1 487  # the grid of BASELINE configuration 5, sampling stride
                                                               # (57 CPU-minutes, ~12 GB)

Stored: FGMRES (iters, r_norm) forward and backward, norms and strided samples of u, d_coeffs, d_rhs, full
d_iv_rhs and d_steps.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic inputs of the benchmark)
from oracle import pde_oracle as O  # noqa: E402


def main():
    out_dir, wname, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
    stride = int(sys.argv[4]) if len(sys.argv) > 4 else 61
    wl = bench.WORKLOADS[wname]
    iv = bench.IV_LISTS[wl["iv"]]
    dims = wl["dims"]
    t0 = time.time()
    st = O.build_structure(dims, iv)
    inp = bench.synth_inputs(wl, B, 1234)
    iv_rhs = (0.5 * torch.randn(B, st.n_init, generator=inp["gen"], dtype=torch.float64)).numpy()
    theta = bench.theta_init(wl, "cpu").detach()
    coeffs = bench.assemble_coeffs(wl, inp["coeffs_base"], inp["field"], theta).numpy()
    rhs = inp["rhs"].numpy()
    steps = [s.numpy() for s in inp["steps"]]
    # forward, then the benchmark's loss sum(u0^2): upstream gradient 2 u0 on channel 0
    def upstream(x):
        g = np.zeros((B, st.G, st.M))
        g[:, :, 0] = 2.0 * x.reshape(B, st.G, st.M)[:, :, 0]
        return g.reshape(B, -1)
    res = O.mg_layer(dims, iv, coeffs, rhs, iv_rhs, steps, wl["n_grid"], wl["dsf"], grad_out=upstream)
    t_all = time.time() - t0
    save = dict(workload=wname, bs=B, seed=1234, stride=stride, info=np.array([res.info_fwd, res.info_bwd], dtype=np.float64),
                coeffs_norm=float(np.linalg.norm(coeffs)), rhs_norm=float(np.linalg.norm(rhs)),
                iv_rhs_norm=float(np.linalg.norm(iv_rhs)), d_iv_rhs=res.d_iv_rhs, oracle_seconds=t_all)
    for k, a in (("u", res.x), ("d_coeffs", res.d_coeffs), ("d_rhs", res.d_rhs)):
        a = np.asarray(a).reshape(-1)
        save[k] = a[::stride].copy()
        save[k + "_norm"] = float(np.linalg.norm(a))
    for c, s in enumerate(res.d_steps):
        save[f"d_steps{c}"] = s
    np.savez_compressed(os.path.join(out_dir, f"port_{wname}_b{B}.npz"), **save)
    print("port", wname, B, "info", res.info_fwd, res.info_bwd, "%.0fs" % t_all, flush=True)


if __name__ == "__main__":
    main()
