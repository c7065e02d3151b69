"""world_size-2 gloo test (CPU) of the multi-GPU host logic: instance sharding and cost gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from elg_b200.dist import gather_costs, shard_batch, shard_slice


def test_shard_slice_partitions_exactly():
    for n in (0, 1, 7, 64, 1000, 10001):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s = shard_slice(n, r, world)
                seen += list(range(n))[s]
            assert seen == list(range(n))
            sizes = [shard_slice(n, r, world).stop - shard_slice(n, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_slice(10, 2, 2)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = {"loc": torch.arange(n * 6, dtype=torch.float32).reshape(n, 3, 2), "demand": torch.arange(n * 3).reshape(n, 3).float()}
    mine = shard_batch(batch, rank, world)
    local_cost = mine["loc"].sum(dim=(1, 2)) + mine["demand"].sum(dim=1)        # stands in for the per-instance best cost
    full = gather_costs(local_cost, n)
    q.put((rank, full.tolist(), mine["loc"].shape[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 10])
def test_two_rank_gloo_gather(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(n * 6, dtype=torch.float32).reshape(n, 3, 2).sum(dim=(1, 2)) + torch.arange(n * 3).reshape(n, 3).float().sum(dim=1)).tolist()
    for rank, full, cnt in got:
        assert full == want
        assert cnt == (n + 1 - rank) // 2 if n % 2 else cnt == n // 2


# ---- training path: data-parallel gradient = all-reduce(sum) of per-shard gradients / world ---------------------------
def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import train_helpers as TH
    from elg_b200.dist import allreduce_mean_gradient
    g = TH.TrainGolden("train_cvrp_n20")                       # 6 instances -> 3 per rank
    s = shard_slice(6, rank, world)
    prob = g.problem()
    from helpers import sub_problem
    _, _, grads, _ = TH.oracle_grads(g.kind, g.meta["model_params"], g.state_dict(), sub_problem(prob, list(range(6))[s]), g.M,
                                     g.tours()[s], g.reward()[s], True)
    keys = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1) for k in keys])
    scale = allreduce_mean_gradient(flat)
    q.put((rank, (flat * scale).tolist() if rank == 0 else None))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_equals_full_batch_gradient():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import train_helpers as TH
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = TH.TrainGolden("train_cvrp_n20")
    _, _, full, _ = TH.oracle_grads(g.kind, g.meta["model_params"], g.state_dict(), g.problem(), g.M, g.tours(), g.reward(), True)
    want = torch.cat([full[k].reshape(-1) for k in sorted(full)])
    have = torch.tensor(got[0])
    assert float((have - want).abs().max()) < 1e-4 * float(want.abs().max())
