"""Data-parallel training on the GPU with two ranks (both on cuda:0, gloo rendezvous on 127.0.0.1 -- the test box has one
GPU; the collectives and the host logic are the ones NCCL runs at N > 1): replicas that were BUILT from different random
parameters agree after construction (rank 0's weights are broadcast), stay bit-identical through the steps, through the
switch to joint training (fresh local-policy parameters drawn per rank) and in what rank 0 checkpoints."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import random
        from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict
        from elg_b200.trainer import Trainer
        torch.manual_seed(100 + rank); random.seed(100 + rank)
        mp_ = dict(DEFAULT_MODEL_PARAMS["cvrp"], ensemble=False)
        sd = {k: v for k, v in synthetic_state_dict("cvrp", seed=7 + rank).items() if "local_polic" not in k}   # differs per rank
        tr = Trainer("cvrp", mp_, sd, "cuda:0")

        def all_equal(t):
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t.contiguous())
            return all(torch.equal(out[0], o) for o in out[1:])

        ok = {"after_init": all_equal(tr.handle.weights)}
        for s in range(2):
            tr.step(synthetic_cvrp_batch(4, 20, seed=50 + 10 * s + rank), 20)      # different instances per rank
        ok["after_steps"] = all_equal(tr.handle.weights) and all_equal(tr.exp_avg)
        tr.add_local_policy()                                                       # per-rank random local parameters
        ok["after_switch"] = all_equal(tr.handle.weights)
        for s in range(2):
            tr.step(synthetic_cvrp_batch(4, 20, seed=90 + 10 * s + rank), 20)
        ok["after_joint_steps"] = all_equal(tr.handle.weights) and all_equal(tr.handle.derived)
        ok["moved"] = bool((tr.exp_avg != 0).any())
        q.put((rank, ok))
        dist.destroy_process_group()
    except Exception as e:      # surface the failure instead of a hang
        q.put((rank, {"error": repr(e)}))


def test_two_rank_replicas_stay_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, ok in res.items():
        assert "error" not in ok, ok
        assert all(ok.values()), (rank, ok)
