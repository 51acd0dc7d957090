"""CPU, world_size 2 over gloo: the N>1 path.  Each rank solves its contiguous shard of the batch with no
communication inside the solve; the result must equal a single-process solve OF THAT SHARD (the Krylov
space couples the instances of one shard: SURVEY.md section 0, fact 3), and the learned-parameter
gradient is the all-reduced sum.  The solve runs on the test-only host emulator here (no GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _layer_run(lib, dims, iv, coeffs_base, field, rhs, ivr, steps, theta):
    from mech_nn_discovery_pde_b200 import MultigridLayer
    B = coeffs_base.shape[0]
    layer = MultigridLayer(bs=B, coord_dims=dims, order=2, n_ind_dim=1, n_iv=1, n_grid=2, downsample_first=True,
                           init_index_mi_list=iv, n_iv_steps=1, _library=lib)
    coeffs = coeffs_base.clone()
    coeffs[..., 0] = theta[0] * field
    coeffs[..., 4] = coeffs[..., 4] + theta[1]
    u0, u, _ = layer(coeffs, rhs, ivr, [s.clone() for s in steps])
    (u0 * u0).sum().backward()
    return u.detach()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mech_nn_discovery_pde_b200.parallel import allreduce_param_grads, shard_batch
    from oracle.cases import IV_LISTS, make_inputs
    from oracle.pde_oracle import build_structure
    from tests.emu.emu_lib import emu_library
    lib = emu_library()
    dims, iv, Bg = (16, 16), IV_LISTS["burgers"], 4
    st = build_structure(dims, iv)
    inp = make_inputs(dims, Bg, st.n_init, seed=2024)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    g = torch.Generator().manual_seed(7)
    field = torch.randn(Bg, st.G, generator=g, dtype=torch.float64)
    full = dict(base=t(inp["coeffs"]), field=field, rhs=t(inp["rhs"]), ivr=t(inp["iv_rhs"]),
                steps=[t(s) for s in inp["steps"]])
    sh = {k: (shard_batch(v, rank, world) if k != "steps" else [shard_batch(s, rank, world) for s in v])
          for k, v in full.items()}
    theta = torch.tensor([0.3, -0.2], dtype=torch.float64, requires_grad=True)
    u = _layer_run(lib, dims, iv, sh["base"], sh["field"], sh["rhs"], sh["ivr"], sh["steps"], theta)
    local_grad = theta.grad.clone()
    allreduce_param_grads([theta])
    q.put((rank, u.numpy(), local_grad.numpy(), theta.grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, u, lg, ag = q.get(timeout=600)
        res[r] = (u, lg, ag)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # all-reduced gradient is the sum of the local ones, identical on both ranks
    total = res[0][1] + res[1][1]
    assert np.allclose(res[0][2], total, rtol=1e-13) and np.allclose(res[1][2], total, rtol=1e-13)
    # each shard equals a single-process run on that shard
    sys.path.insert(0, ROOT)
    from mech_nn_discovery_pde_b200.parallel import shard_batch, shard_bounds
    from oracle.cases import IV_LISTS, make_inputs
    from oracle.pde_oracle import build_structure
    from tests.emu.emu_lib import emu_library
    assert shard_bounds(5, 0, 2) == (0, 3) and shard_bounds(5, 1, 2) == (3, 5)
    lib = emu_library()
    dims, iv, Bg = (16, 16), IV_LISTS["burgers"], 4
    st = build_structure(dims, iv)
    inp = make_inputs(dims, Bg, st.n_init, seed=2024)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    g = torch.Generator().manual_seed(7)
    field = torch.randn(Bg, st.G, generator=g, dtype=torch.float64)
    for r in range(world):
        theta = torch.tensor([0.3, -0.2], dtype=torch.float64, requires_grad=True)
        u = _layer_run(lib, dims, iv, shard_batch(t(inp["coeffs"]), r, world), shard_batch(field, r, world),
                       shard_batch(t(inp["rhs"]), r, world), shard_batch(t(inp["iv_rhs"]), r, world),
                       [shard_batch(t(s), r, world) for s in inp["steps"]], theta)
        assert np.array_equal(u.numpy(), res[r][0])
        assert np.array_equal(theta.grad.numpy(), res[r][1])


def _overlap_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mech_nn_discovery_pde_b200.parallel import OverlappedGradReducer, allreduce_param_grads
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                              torch.nn.Linear(16, 3)).double()
    unused = torch.nn.Parameter(torch.ones(4, dtype=torch.float64))     # never receives a gradient
    params = list(net.parameters()) + [unused]
    red = OverlappedGradReducer(params, bucket_bytes=1024)              # several buckets
    out = []
    for step in range(2):                                              # re-arming between steps
        x = torch.randn(5, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(100 * step + rank))
        for p in params:
            p.grad = None
        (net(x) ** 2).sum().backward()
        local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
        red.finish()
        overlapped = [p.grad.clone() for p in params]
        for p, g in zip(params, local):                                # the one-bucket reference path on the same gradients
            p.grad = g.clone()
        allreduce_param_grads(params)
        out.append(([g.numpy() for g in overlapped], [p.grad.numpy() for p in params], len(red.buckets)))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_grad_reducer_matches_flat_allreduce():
    """Bucketed, hook-driven asynchronous all-reduce of the learned-parameter gradients equals the
    single flat all-reduce, over two steps, with several buckets and a parameter that receives no gradient."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        for overlapped, flat, nb in res[rank]:
            assert nb > 1
            for a, b in zip(overlapped, flat):
                assert np.array_equal(a, b)
    for a, b in zip(res[0][1][0], res[1][1][0]):                       # both ranks hold the same sums
        assert np.array_equal(a, b)
