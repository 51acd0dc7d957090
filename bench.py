#!/usr/bin/env python
"""Benchmark of the differentiable PDE-layer solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload gl32|gl_ref|burgers|gl64]

A "step" is one forward+backward of the multigrid PDE layer over one batch of synthetic instances
(SURVEY.md section 8(d) inputs).  Default workload: Ginzburg-Landau 32x64x64 space-time grid, 32 instances
per GPU, n_grid=4, downsample_first=False (as in the reference's GL script), fp64, config.py default knobs.  One process per GPU; the batch
is sharded (weak scaling: 32 instances per GPU) with no communication inside the solve; the only
collective is the all-reduce of the learned-parameter gradient.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm on the host CPU
(the oracle port: the Python reference itself cannot travel to the GPU box) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dims, iv, per-GPU batch, n_grid, downsample_first, steps h
    # downsample_first=False as in the reference's GL script (discovery/ginzburg_landau.py:241-243): the time
    # axis is never coarsened, 32x64x64 -> 32x32x32 -> 32x16x16 -> 32x8x8
    "gl32": dict(dims=(32, 64, 64), iv="gl", batch=32, n_grid=4, dsf=False, h=(0.1, 0.3906, 0.3906),
                 desc="Ginzburg-Landau 32x64x64, batch 32/GPU, n_grid=4, downsample_first=False"),
    "gl32_dsf": dict(dims=(32, 64, 64), iv="gl", batch=32, n_grid=3, dsf=True, h=(0.1, 0.3906, 0.3906),
                     desc="Ginzburg-Landau 32x64x64, batch 32/GPU, n_grid=3, downsample_first=True"),
    "gl64": dict(dims=(64, 128, 128), iv="gl", batch=32, n_grid=4, dsf=True, h=(0.1, 0.3906, 0.3906),
                 desc="Ginzburg-Landau 64x128x128, batch 32/GPU, n_grid=4, downsample_first=True"),
    "gl_ref": dict(dims=(8, 32, 32), iv="gl", batch=32, n_grid=3, dsf=False, h=(0.1, 0.3906, 0.3906),
                   desc="Ginzburg-Landau reference default 8x32x32, batch 32/GPU, n_grid=3, downsample_first=False"),
    "burgers": dict(dims=(256, 256), iv="burgers", batch=64, n_grid=6, dsf=True, h=(0.025, 20.0 / 256),
                    desc="Burgers 256x256, batch 64/GPU, n_grid=6, downsample_first=True"),
}

IV_LISTS = {
    # discovery/ginzburg_landau.py:225-237
    "gl": [
        lambda nt, nx, ny: (0, 0, [0, 0, 0], [0, nx - 1, ny - 1]),
        lambda nt, nx, ny: (1, 0, [1, 0, 0], [nt - 1, 0, ny - 1]),
        lambda nt, nx, ny: (2, 0, [1, 1, 0], [nt - 1, nx - 1, 0]),
        lambda nt, nx, ny: (1, 0, [1, nx - 1, 1], [nt - 1, nx - 1, ny - 1]),
        lambda nt, nx, ny: (2, 0, [1, 1, ny - 1], [nt - 1, nx - 2, ny - 1]),
    ],
    # discovery/burgers_dparam_viscous.py:170-173
    "burgers": [
        lambda nx, ny: (0, 0, [0, 0], [0, ny - 2]),
        lambda nx, ny: (1, 0, [1, 0], [nx - 1, 0]),
        lambda nx, ny: (1, 0, [0, ny - 1], [nx - 1, ny - 1]),
    ],
}


def synth_inputs(wl, B, seed):
    """fp64 synthetic inputs on the host."""
    dims = wl["dims"]
    d = len(dims)
    G = int(np.prod(dims))
    M = 1 + 2 * d
    g = torch.Generator().manual_seed(seed)
    coeffs = torch.zeros(B, G, M, dtype=torch.float64)
    field = torch.randn(B, G, generator=g, dtype=torch.float64)
    if wl["iv"] == "gl":
        coeffs[..., 1] = 1.0
    else:  # burgers: u_t + p u_x - 0.1 u_xx
        coeffs[..., 1] = 1.0
        field = torch.rand(B, G, generator=g, dtype=torch.float64)
    rhs = 0.1 * torch.randn(B, G, generator=g, dtype=torch.float64)
    steps = [torch.full((B, n - 1), h, dtype=torch.float64) for n, h in zip(dims, wl["h"])]
    return dict(coeffs_base=coeffs, field=field, rhs=rhs, steps=steps, G=G, M=M, d=d, gen=g)


def assemble_coeffs(wl, base, field, theta):
    """The step before the op in the discovery scripts: learned parameters -> coeffs (ginzburg_landau.py:354-374)."""
    coeffs = base.clone()
    d = len(wl["dims"])
    if wl["iv"] == "gl":
        coeffs[..., 0] = theta[0] * field
        coeffs[..., 1 + d + 1] = theta[1]
        coeffs[..., 1 + d + 2] = theta[2]
    else:
        coeffs[..., 2] = theta[0] * field
        coeffs[..., 4] = theta[1]
    return coeffs


def theta_init(wl, device):
    if wl["iv"] == "gl":
        return torch.tensor([0.1, -1.0, -1.0], dtype=torch.float64, device=device, requires_grad=True)
    return torch.tensor([1.0, -0.1], dtype=torch.float64, device=device, requires_grad=True)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(plan, cfgs, B):
    """Per-launch algorithmic bytes of the HBM-bound kernel groups."""
    from mech_nn_discovery_pde_b200 import _lib
    w = 8
    M = plan.M
    G0 = plan.G
    nc = plan.lib.query(plan.handle, _lib.Q_G, plan.n_grid - 1) * M
    # coarsest level in band ordering (longest axis outermost): half-bandwidth 4*inner*M + M - 1
    cd = plan.dims_list[-1]
    inner = int(np.prod(cd)) // max(cd)
    bw = min(4 * inner * M + M - 1, nc - 1)
    nsw = cfgs["gs_pre"]
    return {
        "gs_fine": 4 * M * w * G0 * B * nsw,               # per sweep: read x, b, coeffs; write x
        "apply_fine": 3 * M * w * G0 * B,                   # read z, coeffs; write y  (residual: +b)
        "coarse_solve": 2 * nc * (bw + 1) * w * B,          # band of L streamed once per triangular solve
    }


def run_ours(args, wl):
    from mech_nn_discovery_pde_b200 import MultigridLayer, PDEConfig, _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or wl["batch"]
    dims = wl["dims"]
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=wl["n_grid"],
                           downsample_first=wl["dsf"], init_index_mi_list=IV_LISTS[wl["iv"]], n_iv_steps=1)
    lib = _lib.get_library()
    plan = layer.mg_solver.plan
    inp = synth_inputs(wl, B, 1234 + rank)
    n_init = plan.n_init
    iv_host = 0.5 * torch.randn(B, n_init, generator=inp["gen"], dtype=torch.float64)
    host = dict(base=inp["coeffs_base"].pin_memory(), field=inp["field"].pin_memory(), rhs=inp["rhs"].pin_memory(),
                iv=iv_host.pin_memory(), steps=[s.pin_memory() for s in inp["steps"]])
    theta = theta_init(wl, dev)

    def to_dev():
        return dict(base=host["base"].to(dev, non_blocking=True), field=host["field"].to(dev, non_blocking=True),
                    rhs=host["rhs"].to(dev, non_blocking=True), iv=host["iv"].to(dev, non_blocking=True),
                    steps=[s.to(dev, non_blocking=True) for s in host["steps"]])

    def step(dv):
        if theta.grad is not None:
            theta.grad = None
        coeffs = assemble_coeffs(wl, dv["base"], dv["field"], theta)
        u0, u, _ = layer(coeffs, dv["rhs"], dv["iv"], list(dv["steps"]))
        loss = (u0 * u0).sum()          # upstream gradient g = 2 u0
        loss.backward()
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(theta.grad)   # the only collective: learned-parameter gradient
        return loss

    resident = to_dev()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step(resident)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.profile_enable(True)
    launches0 = lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(resident)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.launch_count() - launches0
    prof = lib.profile_collect()
    lib.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    fwd_info, bwd_info = layer.solver_info()

    # end to end through the public layer with HOST buffers: H2D of the step's inputs and D2H of the loss inside
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        dv = to_dev()
        loss = step(dv)
        _ = float(loss.item())
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    h2d = sum(t.numel() * 8 for t in [host["base"], host["field"], host["rhs"], host["iv"]] + host["steps"])

    tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    solves = B * world * args.steps
    value = solves / (ms / 1e3)
    e2e_value = solves / (ms_e2e / 1e3)

    peaks = {}
    peak_src = "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    cfgs = dict(gs_pre=int(PDEConfig.mg_gauss_seidel_steps_pre))
    ab = algorithmic_bytes(plan, cfgs, B)
    total_prof_ms = sum(v[0] for v in prof.values()) or 1.0
    breakdown = {k: {"ms_per_step": v[0] / args.steps, "groups_per_step": v[1] / args.steps,
                     "share": v[0] / total_prof_ms} for k, v in prof.items() if v[1] > 0}
    for k in ab:
        if k in breakdown and prof[k][1] > 0:
            avg_s = prof[k][0] / 1e3 / prof[k][1]
            breakdown[k]["algorithmic_gbs"] = ab[k] / avg_s / 1e9
            breakdown[k]["frac_of_hbm_peak"] = ab[k] / avg_s / 1e9 / hbm_peak
    hbm_kernels = [k for k in ("gs_fine", "coarse_solve", "apply_fine") if k in breakdown]
    dom = max(hbm_kernels, key=lambda k: breakdown[k]["ms_per_step"])
    traffic = None
    try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if dom in tj and args.workload == "gl32" and B == wl["batch"]:
            traffic = tj[dom]["bytes"]
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": breakdown[dom]["algorithmic_gbs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": breakdown[dom]["frac_of_hbm_peak"], "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": ab[dom],
                "avg_launch_ms": prof[dom][0] / prof[dom][1], "share_of_step": breakdown[dom]["share"]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(wl, seconds_budget=args.cpu_budget)
    out = {
        "metric": "PDE-layer fwd+bwd solves/sec (GL grid, fp64)", "value": value, "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch_per_gpu": B, "global_batch": B * world,
                   "fgmres": "restart 10, maxiter 40, atol 1e-5, GS 5+5, 1 V-cycle", "parallelism": f"batch-shard x{world}",
                   "l2": "inputs larger than L2 (every vector >= 235 MB)"},
        "fgmres_info": {"forward": fwd_info, "backward": bwd_info},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": roofline, "breakdown": breakdown, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores (bounded sample)
# -------------------------------------------------------------------------------------------------------------
def cpu_sample(wl, seconds_budget=60.0, iters=1):
    """One instance of the workload on the CPU: operator set-up measured once, `iters` Arnoldi steps
    (V-cycle + normal matvec + Gram-Schmidt) measured and extrapolated to the 40+40 iterations the layer runs."""
    from oracle import pde_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dims = wl["dims"]
    iv = IV_LISTS[wl["iv"]]
    t0 = time.time()
    st = O.build_structure(dims, iv)
    inp = synth_inputs(wl, 1, 1234)
    theta = theta_init(wl, "cpu").detach()
    coeffs = assemble_coeffs(wl, inp["coeffs_base"], inp["field"], theta).numpy()
    iv_rhs = 0.5 * np.random.default_rng(0).standard_normal((1, st.n_init))
    mg = O.mg_setup(dims, iv, coeffs, inp["rhs"].numpy(), iv_rhs, [s.numpy() for s in inp["steps"]], wl["n_grid"],
                    wl["dsf"])
    t_setup = time.time() - t0
    K0 = mg.K_list[0]
    v = mg.Atb0 / np.linalg.norm(mg.Atb0)
    V = [v]
    t_it = []
    for j in range(iters):
        t1 = time.time()
        z = O.v_cycle_start(mg, V[-1])
        u = K0 @ z
        Vm = np.stack(V, axis=1)
        h = Vm.T @ u
        u = u - Vm @ h
        V.append(u / np.linalg.norm(u))
        t_it.append(time.time() - t1)
        if time.time() - t0 > seconds_budget:
            break
    t_iter = float(np.mean(t_it))
    n_iters = 80
    total = t_setup + n_iters * t_iter
    return {"value": 1.0 / total, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"1 instance of the workload: operator set-up measured once ({t_setup:.1f}s) + {len(t_it)} "
                      f"Arnoldi step(s) (V-cycle+matvec+CGS, {t_iter:.2f}s each) extrapolated to 40 fwd + 40 bwd "
                      f"iterations; scipy triangular solves are single-threaded, BLAS uses all cores",
            "setup_s": t_setup, "arnoldi_step_s": t_iter}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = None
    vals = []
    for i in range(max(1, min(args.steps, 3))):
        res = cpu_sample(wl, seconds_budget=args.cpu_budget, iters=1)
        vals.append(res["value"])
    value = float(np.mean(vals))
    res["value"] = value
    out = {"impl": "reference", "metric": "PDE-layer fwd+bwd solves/sec (GL grid, fp64)", "value": value,
           "unit": "solves/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl["desc"], "note": "reference algorithm on host CPU (oracle port), per-instance rate"},
           "cpu_baseline": res,
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gl32", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=120.0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, wl)


if __name__ == "__main__":
    main()
