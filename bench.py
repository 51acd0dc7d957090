#!/usr/bin/env python
"""Benchmark of the differentiable PDE-layer solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload gl32|gl32_dsf|gl_ref|burgers|gl64|kamani|sine]

A "step" is one forward+backward of the multigrid PDE layer over one batch of synthetic instances
(SURVEY.md section 8(d) inputs).  Default workload: Ginzburg-Landau 32x64x64 space-time grid, 32 instances
per GPU, n_grid=4, downsample_first=False (as in the reference's GL script), fp64, config.py default knobs.  One process per GPU; the batch
is sharded (weak scaling: 32 instances per GPU) with no communication inside the solve; the only
collective is the all-reduce of the learned-parameter gradient.

The other BASELINE.json configurations are selectable with --workload: `kamani` (dense layer, (24,) grid, batch
4096), `sine` (dense layer, 32x32, batch 1), `burgers` (256x256, batch 64), `gl64` (64x128x128, 32 per GPU: the
grid of the scaling configuration).

Prints ONE JSON line (rank 0).  `value` is timed with the library's instrumentation OFF; the per-kernel-group
breakdown and the roofline come from a second, separately timed pass with per-plan instrumentation on.  The step
(everything but the collective) is captured in a CUDA graph and replayed (`--graph auto`; `cuda_graph` in the JSON
line says whether it was): the dense workloads are host-launch bound, the multigrid ones gain ~2 %; the replay is
checked against an eager step bit for bit before it is used.
`--impl reference` times the reference's algorithm on the host CPU on a bounded sample: the oracle port for the
workload itself and, when oracle/_ref (the unmodified reference + stub modules, built by oracle/make_ref.py) is
present, the unmodified reference on its own default Ginzburg-Landau configuration as a cross-check.
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
    # dense layer (PDEDenseLayer): discovery/kamani.py:45-46,153-165 and fit/sine_pde_dense.py:98,106-119
    "kamani": dict(dims=(24,), iv="kamani", batch=4096, n_grid=1, dsf=True, h=(0.0314,), dense=True,
                   desc="Kamani ODE (24,) dense layer, batch 4096/GPU"),
    "sine": dict(dims=(32, 32), iv="sine", batch=1, n_grid=1, dsf=True, h=(0.05, 0.05), dense=True,
                 desc="sine fit 32x32 dense layer, batch 1"),
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
    # discovery/kamani.py:153-156
    "kamani": [
        lambda nt: (0, 0, [0], [0]),
    ],
    # fit/sine_pde_dense.py:111-115
    "sine": [
        lambda nx, ny: (0, 0, [0, 0], [0, ny - 2]),
        lambda nx, ny: (1, 0, [1, 0], [nx - 1, 0]),
        lambda nx, ny: (0, 0, [nx - 1, 1], [nx - 1, ny - 2]),
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
    elif wl["iv"] == "burgers":  # u_t + p u_x - 0.1 u_xx
        coeffs[..., 1] = 1.0
        field = torch.rand(B, G, generator=g, dtype=torch.float64)
    elif wl["iv"] == "kamani":   # c0 = 1 + 0.1 N, c1 = 0.1 + |N|
        coeffs[..., 1] = 0.1 + torch.randn(B, G, generator=g, dtype=torch.float64).abs()
    else:                        # sine: coefficients ~ N(0,1), constant over the grid
        coeffs[:] = torch.randn(B, 1, M, generator=g, dtype=torch.float64)
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
    elif wl["iv"] == "burgers":
        coeffs[..., 2] = theta[0] * field
        coeffs[..., 4] = theta[1]
    elif wl["iv"] == "kamani":
        coeffs[..., 0] = theta[0] + theta[1] * field
    else:
        coeffs[..., 0] = base[..., 0] * theta[0]
    return coeffs


def make_builder(wl):
    """The same assembly as assemble_coeffs through the fused coefficient-builder kernel:
    returns build(field, theta) -> coeffs (B,G,M); None where the workload's base coefficients vary per point."""
    from mech_nn_discovery_pde_b200.coeffs import CoeffBuilder
    d = len(wl["dims"])
    M = 1 + 2 * d
    if wl["iv"] == "gl":       # c0 = theta0 * field, c1 = 1, c5 = theta1, c6 = theta2
        cb = CoeffBuilder(M, 1, [(0,), (1,)], [(0, 1), (1 + d + 1, 0), (1 + d + 2, 0)], const={1: 1.0})
        return lambda field, theta: cb([field], theta)[0]
    if wl["iv"] == "burgers":  # c1 = 1, c2 = theta0 * field, c4 = theta1
        cb = CoeffBuilder(M, 1, [(0,), (1,)], [(2, 1), (4, 0)], const={1: 1.0})
        return lambda field, theta: cb([field], theta)[0]
    return None


def theta_init(wl, device):
    if wl["iv"] == "gl":
        return torch.tensor([0.1, -1.0, -1.0], dtype=torch.float64, device=device, requires_grad=True)
    if wl["iv"] == "kamani":
        return torch.tensor([1.0, 0.1], dtype=torch.float64, device=device, requires_grad=True)
    if wl["iv"] == "sine":
        return torch.tensor([1.0], dtype=torch.float64, device=device, requires_grad=True)
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


def algorithmic_bytes(plan, wl, cfgs, B):
    """Per-launch algorithmic bytes / flops of the kernel groups."""
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
        "factor_flops": B * nc * (bw + 1.0) * (bw + 2.0),   # band Cholesky: n*bw^2 multiply-adds = 2x flops / 2
    }


PARAM_GRAD_DOUBLES = 4 * (512 * 1024 + 1024 * 1024 + 1024 * 10)   # the GL model's four ParamNets


def run_ours(args, wl):
    from mech_nn_discovery_pde_b200 import MultigridLayer, PDEConfig, PDEDenseLayer
    from mech_nn_discovery_pde_b200 import parallel
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG at VERSION or above):
        # stdout carries ONE JSON line, so file descriptor 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    B = args.batch or wl["batch"]
    dims = wl["dims"]
    dense = bool(wl.get("dense"))
    if dense:
        layer = PDEDenseLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1,
                              init_index_mi_list=IV_LISTS[wl["iv"]], n_iv_steps=1, device=dev)
        plan = layer.plan
    else:
        layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=wl["n_grid"],
                               downsample_first=wl["dsf"], init_index_mi_list=IV_LISTS[wl["iv"]], n_iv_steps=1,
                               device=dev)
        plan = layer.mg_solver.plan
    lib = plan.lib
    inp = synth_inputs(wl, B, 1234 + rank)
    n_init = plan.n_init
    iv_host = 0.5 * torch.randn(B, n_init, generator=inp["gen"], dtype=torch.float64)
    host = dict(base=inp["coeffs_base"].pin_memory(), field=inp["field"].pin_memory(), rhs=inp["rhs"].pin_memory(),
                iv=iv_host.pin_memory(), steps=[s.pin_memory() for s in inp["steps"]])
    u0_host = torch.empty(B, 1, inp["G"], dtype=torch.float64).pin_memory()
    theta = theta_init(wl, dev)
    # stand-in for the learned ParamNets of the discovery model: their gradient is what the training step
    # all-reduces (~50 MB fp64).  theta's gradient is scattered into it so the collective carries live data.
    pnet = torch.zeros(PARAM_GRAD_DOUBLES, dtype=torch.float64, device=dev, requires_grad=True)
    pnet.grad = torch.zeros_like(pnet)

    def to_dev():
        return dict(base=host["base"].to(dev, non_blocking=True), field=host["field"].to(dev, non_blocking=True),
                    rhs=host["rhs"].to(dev, non_blocking=True), iv=host["iv"].to(dev, non_blocking=True),
                    steps=[s.to(dev, non_blocking=True) for s in host["steps"]])

    builder = None if args.no_fused_builder else make_builder(wl)

    def local_step(dv):
        if theta.grad is not None:
            theta.grad = None
        if builder is not None:
            coeffs = builder(dv["field"], theta)
        else:
            coeffs = assemble_coeffs(wl, dv["base"], dv["field"], theta)
        u0, u, _ = layer(coeffs, dv["rhs"], dv["iv"], list(dv["steps"]))
        loss = (u0 * u0).sum()          # upstream gradient g = 2 u0
        loss.backward()
        pnet.grad[:theta.numel()] = theta.grad
        return loss, u0

    def step(dv):
        out = local_step(dv)
        if world > 1:
            parallel.allreduce_param_grads([theta, pnet])   # the only collective: learned-parameter gradients
        return out

    resident = to_dev()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step(resident)
    torch.cuda.synchronize()

    # ---- CUDA graph of the step (everything but the collective).  The dense workloads (~100 short launches per step)
    # are host-launch bound: 1.6-1.8x; the multigrid workloads gain ~2 %.  The captured step reads the static device
    # tensors of `resident`; the e2e pass copies each step's host inputs into them.  The Cholesky status check moves out
    # of the captured backward (it reads a value back) to after the replay.  `auto` leaves the 64x128x128 workload eager:
    # a graph's private pool would hold a second 32 GB operator state next to the 100 GB of scratch.
    graph = None
    graph_launches = 0
    want_graph = args.graph == "on" or (args.graph == "auto" and args.workload != "gl64")
    if want_graph:
        import warnings
        warnings.filterwarnings("ignore", message="The AccumulateGrad node's stream does not match")
        try:
            class GraphCfg(PDEConfig):
                check_factorization = False
            layer.config = GraphCfg
            eager_step = step
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    local_step(resident)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = lib.launch_count()
            with torch.cuda.graph(g):
                loss_g, u0_g = local_step(resident)
                if os.environ.get("PDEOP_BENCH_FORCE_GRAPH_FAIL"):      # test hook for the fallback path
                    raise RuntimeError("forced capture failure")
            graph_launches = lib.launch_count() - l0
            loss_eager = float(local_step(resident)[0].item())
            g.replay()
            torch.cuda.synchronize()
            layer.last_holder.check_factorization()
            if float(loss_g.item()) != loss_eager:      # same kernels on the same inputs: same bits
                raise RuntimeError(f"graph replay loss {float(loss_g.item())!r} != eager loss {loss_eager!r}")
            graph = g

            def step(dv):       # noqa: F811  (dv must be `resident`: the graph reads those tensors)
                graph.replay()
                if world > 1:
                    parallel.allreduce_param_grads([theta, pnet])
                return loss_g, u0_g
        except Exception as e:   # noqa: BLE001
            import traceback
            print(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            if os.environ.get("PDEOP_BENCH_TRACE"):
                traceback.print_exc()
            graph = None
            layer.config = PDEConfig
            step = eager_step
            # leave nothing of the aborted capture behind: its private pool, half-built autograd state, stream state
            g = loss_g = u0_g = None
            theta.grad = None
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            for _ in range(max(args.warmup, 3)):
                step(resident)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pass 1: `value`, instrumentation off -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
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
    if graph is not None:      # replays launch no kernel from the host: count what one captured step contains
        launches = graph_launches * args.steps
    clocks = sampler.stop() if rank == 0 else None
    fwd_info, bwd_info = layer.solver_info()

    # ---- pass 2: end to end through the public layer with HOST buffers: H2D of the step's inputs, D2H of u0 and
    # ---- of the loss inside the timed region ---------------------------------------------------------------
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        if graph is not None:       # into the tensors the graph reads
            for k in ("base", "field", "rhs", "iv"):
                resident[k].copy_(host[k], non_blocking=True)
            for a, b_ in zip(resident["steps"], host["steps"]):
                a.copy_(b_, non_blocking=True)
            dv = resident
        else:
            dv = to_dev()
        loss, u0 = step(dv)
        u0_host.copy_(u0.detach().reshape(u0_host.shape), non_blocking=True)
        _ = float(loss.item())
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    h2d = sum(t.numel() * 8 for t in [host["base"], host["field"], host["rhs"], host["iv"]] + host["steps"])
    d2h = u0_host.numel() * 8 + 8

    # ---- pass 3: per-kernel-group device times (events on the launching stream), separately timed ------------
    prof_steps = max(1, min(args.steps, 3))
    if graph is not None:           # the instrumentation records events from the host: eager steps
        step = eager_step
    plan.profile_enable(True)
    barrier()
    for _ in range(prof_steps):
        step(resident)
    barrier()
    prof = plan.profile_collect()
    plan.profile_enable(False)

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
    ab = algorithmic_bytes(plan, wl, cfgs, B)
    total_prof_ms = sum(v[0] for v in prof.values()) or 1.0
    breakdown = {k: {"ms_per_step": v[0] / prof_steps, "groups_per_step": v[1] / prof_steps,
                     "share": v[0] / total_prof_ms} for k, v in prof.items() if v[1] > 0}
    for k in ("gs_fine", "apply_fine", "coarse_solve"):
        if k in breakdown and prof[k][1] > 0:
            avg_s = prof[k][0] / 1e3 / prof[k][1]
            breakdown[k]["algorithmic_gbs"] = ab[k] / avg_s / 1e9
            breakdown[k]["frac_of_hbm_peak"] = ab[k] / avg_s / 1e9 / hbm_peak
    if "factor" in breakdown:
        avg_s = prof["factor"][0] / 1e3 / prof["factor"][1]
        breakdown["factor"]["fp64_tflops"] = ab["factor_flops"] / avg_s / 1e12
    traffic = None
    if dense:
        # dense layer: the factorisation is the GEMM-shaped part (fp64 tensor pipe / FMA bound); B200 fp64 dense
        # peak is not in MEASURED_PEAKS.json, the nominal 40 TFLOP/s (B200_PROFILING.md) is used and said so
        dom = "factor"
        fl = breakdown[dom].get("fp64_tflops", 0.0)
        roofline = {"kernel": dom, "bound": "tensor", "achieved": fl, "peak": 40.0, "unit": "TFLOP/s",
                    "frac": fl / 40.0, "traffic": None, "peak_source": "nominal fp64 (no measured fp64 peak on this pool)",
                    "flops_per_launch": ab["factor_flops"], "avg_launch_ms": prof[dom][0] / prof[dom][1],
                    "share_of_step": breakdown[dom]["share"]}
    hbm_kernels = [k for k in ("gs_fine", "coarse_solve", "apply_fine") if k in breakdown and not dense]
    dom = max(hbm_kernels, key=lambda k: breakdown[k]["ms_per_step"]) if hbm_kernels else None
    try:   # DRAM bytes per launch of this kernel from the committed ncu --set full capture of this round (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if dom in tj and args.workload == tj.get("workload") and B == wl["batch"]:
            traffic = tj[dom]["bytes"]
    except Exception:
        pass
    if dom is not None:
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
                   "solver": ("dense Cholesky" if dense else "fgmres restart 10, maxiter 40, atol 1e-5, GS 5+5, 1 V-cycle"),
                   "parallelism": f"batch-shard x{world}",
                   "collective": f"all-reduce of {(PARAM_GRAD_DOUBLES + theta.numel()) * 8 / 1e6:.1f} MB fp64 parameter gradients per step",
                   "l2": "inputs larger than L2 (every vector >= 235 MB)" if not dense else
                         "working set per step larger than L2" if B * plan.n * plan.n * 8 > 126e6 else "L2 not flushed (working set < L2)"},
        "fgmres_info": {"forward": fwd_info, "backward": bwd_info},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "cuda_graph": graph is not None,
        "roofline": roofline, "breakdown": breakdown, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (bounded sample)
# -------------------------------------------------------------------------------------------------------------
def cpu_sample(wl, seconds_budget=60.0, iters=5):
    """One instance of the workload on the CPU with the oracle port of the reference algorithm.

    Multigrid workloads: operator set-up measured once; then `iters` Arnoldi steps (V-cycle + normal matvec +
    Gram-Schmidt) of the FORWARD solve and `iters` of the BACKWARD solve (right-hand side 2 u0 of the partial
    forward iterate: a real backward, same operator) are measured and extrapolated to the 40 + 40 iterations the
    layer runs; the gradient formulas are measured too.  Dense workloads: the full layer forward+backward."""
    from oracle import pde_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dims = wl["dims"]
    iv = IV_LISTS[wl["iv"]]
    t0 = time.time()
    st = O.build_structure(dims, iv)
    if wl.get("dense"):
        B = min(wl["batch"], 256)
        inp = synth_inputs(wl, B, 1234)
        theta = theta_init(wl, "cpu").detach()
        coeffs = assemble_coeffs(wl, inp["coeffs_base"], inp["field"], theta).numpy()
        iv_rhs = 0.5 * np.random.default_rng(0).standard_normal((B, st.n_init))
        steps = [s.numpy() for s in inp["steps"]]
        t1 = time.time()
        res = O.dense_layer(dims, iv, coeffs, inp["rhs"].numpy(), iv_rhs, steps)
        g = np.zeros((B, st.G, st.M))
        g[:, :, 0] = 2.0 * res.x.reshape(B, st.G, st.M)[:, :, 0]
        O.dense_layer(dims, iv, coeffs, inp["rhs"].numpy(), iv_rhs, steps, grad_out=g.reshape(B, -1))
        dt = time.time() - t1
        return {"value": B / dt, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"{B} instance(s) of the workload, full dense forward+backward with the oracle port "
                          f"({dt:.2f}s; LAPACK Cholesky, all cores)"}
    inp = synth_inputs(wl, 1, 1234)
    theta = theta_init(wl, "cpu").detach()
    coeffs = assemble_coeffs(wl, inp["coeffs_base"], inp["field"], theta).numpy()
    iv_rhs = 0.5 * np.random.default_rng(0).standard_normal((1, st.n_init))
    mg = O.mg_setup(dims, iv, coeffs, inp["rhs"].numpy(), iv_rhs, [s.numpy() for s in inp["steps"]], wl["n_grid"],
                    wl["dsf"])
    t_setup = time.time() - t0
    K0 = mg.K_list[0]

    def arnoldi(b, back):
        v = b / np.linalg.norm(b)
        V, Z, t_it = [v], [], []
        for j in range(iters):
            t1 = time.time()
            z = O.v_cycle_start(mg, V[-1], back=back)
            u = K0 @ z
            Vm = np.stack(V, axis=1)
            h = Vm.T @ u
            u = u - Vm @ h
            V.append(u / np.linalg.norm(u))
            Z.append(z)
            t_it.append(time.time() - t1)
            if time.time() - t0 > seconds_budget:
                break
        return Z, t_it
    Zf, tf = arnoldi(mg.Atb0, False)
    x_part = np.sum(Zf, axis=0)                       # some iterate in the forward Krylov space
    g = np.zeros((st.G, st.M))
    g[:, 0] = 2.0 * x_part.reshape(st.G, st.M)[:, 0]
    Zb, tb = arnoldi(g.reshape(-1), True)
    t1 = time.time()
    x2, dz2 = x_part.reshape(1, -1), np.sum(Zb, axis=0).reshape(1, -1)
    lam = mg.b0 - np.stack([mg.A0[0] @ x2[0]])
    O._grads(st, mg.A0, x2, lam, dz2, None, None, 1)
    t_grads = time.time() - t1
    tfm, tbm = float(np.mean(tf)), float(np.mean(tb))
    total = t_setup + 40 * tfm + 40 * tbm + t_grads
    return {"value": 1.0 / total, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"1 instance of the workload: operator set-up measured once ({t_setup:.1f}s) + {len(tf)} forward "
                      f"and {len(tb)} backward Arnoldi steps (V-cycle+matvec+CGS; {tfm:.2f}s / {tbm:.2f}s each) "
                      f"extrapolated to 40 + 40 iterations + gradients ({t_grads:.2f}s); scipy triangular solves are "
                      f"single-threaded, BLAS uses all cores",
            "setup_s": t_setup, "arnoldi_step_fwd_s": tfm, "arnoldi_step_bwd_s": tbm, "measured_s": time.time() - t0}


def reference_unmodified_sample():
    """The UNMODIFIED reference (oracle/_ref: the reference's Python files + the stub modules, see
    oracle/make_ref.py) on its own default Ginzburg-Landau configuration (8x32x32, n_grid 3, downsample_first
    False, discovery/ginzburg_landau.py:52-57,241-243), batch 2, forward+backward, in a subprocess."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    runner = os.path.join(ROOT, "oracle", "run_ref.py")
    if not os.path.isdir(os.path.join(ref_dir, "solver")):
        return {"unavailable": "oracle/_ref not built (python oracle/make_ref.py needs /root/reference)"}
    try:
        out = subprocess.run([sys.executable, runner, "gl_ref", "2"], capture_output=True, text=True, timeout=900)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if out.returncode != 0 or not line:
            return {"unavailable": "oracle/run_ref.py failed: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200]}
        return json.loads(line[-1])
    except Exception as e:   # noqa: BLE001
        return {"unavailable": f"oracle/run_ref.py: {e}"}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_sample(wl, seconds_budget=args.cpu_budget, iters=5)
    value = res["value"]
    unmod = None if args.no_unmodified_reference else reference_unmodified_sample()
    out = {"impl": "reference", "metric": "PDE-layer fwd+bwd solves/sec (GL grid, fp64)", "value": value,
           "unit": "solves/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl["desc"], "note": "reference algorithm on host CPU (oracle port), per-instance rate, "
                                                        "bounded sample extrapolated to the layer's 40+40 iterations"},
           "cpu_baseline": res, "reference_unmodified": unmod,
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
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="capture the step in a CUDA graph (auto: every workload but gl64, whose state would not fit twice)")
    ap.add_argument("--cpu-budget", type=float, default=120.0)
    ap.add_argument("--no-fused-builder", action="store_true",
                    help="assemble coeffs with PyTorch ops instead of the fused coefficient-builder kernel")
    ap.add_argument("--no-unmodified-reference", action="store_true",
                    help="reference arm: skip the unmodified-reference cross-check (oracle/_ref)")
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
