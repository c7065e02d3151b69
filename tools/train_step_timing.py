#!/usr/bin/env python
"""GPU: per-phase timing (CUDA events) of one CVRP100 / TSP100 REINFORCE step: B instances x M POMO rows, sample mode.

    python tools/train_step_timing.py [B] [steps] [chunk_steps] [cvrp|tsp]
"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elg_b200 import engine, _lib                            # noqa: E402
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch   # noqa: E402
from elg_b200.trainer import Trainer                         # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    kind = sys.argv[4] if len(sys.argv) > 4 else "cvrp"
    N, M = 100, 100
    dev = "cuda:0"
    tr = Trainer(kind, dict(DEFAULT_MODEL_PARAMS[kind]), synthetic_state_dict(kind, seed=1234, gain=1.0), dev, chunk_steps=chunk)
    names = ["load", "encode_train", "rollout_sample", "backward", "adam+prepare"]
    tot = {k: 0.0 for k in names}
    Ts = []
    for it in range(steps + 2):
        if kind == "cvrp":
            data = {k: v.to(dev) for k, v in synthetic_cvrp_batch(B, N, seed=100 + it).items()}
        else:
            data = synthetic_tsp_batch(B, N, seed=100 + it).to(dev)
        random.seed(it)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        ev[0].record()
        if kind == "cvrp":
            xy, dem = engine.load_problems("cvrp", data["loc"], data["depot"], data["demand"], 1)
        else:
            xy, dem = engine.load_problems("tsp", data, None, None, 1)
        ev[1].record()
        batch, saved = engine.encode_train(tr.handle, xy, dem)
        ev[2].record()
        tours, reward, logp, n_steps = engine.rollout(batch, M, random.sample(range(0, N if kind == "cvrp" else M), M), "sample", seed=it)
        ev[3].record()
        T = int(n_steps.max().item())
        grads, loss, ws = engine.reinforce_backward(batch, saved, M, tours, T, reward, logp, True, chunk, tr.grads)
        ev[4].record()
        tr.optimizer_step()
        ev[5].record()
        torch.cuda.synchronize()
        if it >= 2:
            for i, k in enumerate(names):
                tot[k] += ev[i].elapsed_time(ev[i + 1])
            Ts.append(T)
        print("step %d T=%d loss=%.5f mean cost=%.4f launches=%d total=%.1f ms" % (
            it, T, float(loss), float(-reward.mean()), _lib.launch_count() - l0, ev[0].elapsed_time(ev[5])), flush=True)
    s = sum(tot.values()) / steps
    print(kind + " B=%d M=%d N=%d chunk=%d: %.1f ms/step = %.1f instances/s; %s" % (
        B, M, N, chunk, s, B / s * 1e3, ", ".join("%s %.1f" % (k, tot[k] / steps) for k in names)))


if __name__ == "__main__":
    main()
