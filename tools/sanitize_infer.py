#!/usr/bin/env python
"""Small greedy rollouts through both tensor-core rollout kernels for compute-sanitizer:
    compute-sanitizer --tool memcheck  --kernel-name kns=elg python tools/sanitize_infer.py
    compute-sanitizer --tool racecheck --kernel-name kns=elg python tools/sanitize_infer.py
rollout_tc_kernel: CVRP30 / TSP30 x 8 aug (attention = tensor); rollout_stc_kernel: CVRP130 (M = 20) and TSP140 (M = 12) x 8 aug."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from elg_b200 import engine
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch

DEV = "cuda:0"
for kind, N, M, att in (("cvrp", 30, 30, "tensor"), ("tsp", 30, 30, "tensor"), ("cvrp", 130, 20, "auto"), ("tsp", 140, 12, "auto")):
    mp = dict(DEFAULT_MODEL_PARAMS[kind])
    h = engine.ModelHandle(kind, mp, synthetic_state_dict(kind, seed=3, gain=3.0), DEV, attention=att)
    if kind == "cvrp":
        d = synthetic_cvrp_batch(1, N, seed=5)
        xy, dem = engine.load_problems("cvrp", d["loc"].to(DEV), d["depot"].to(DEV), d["demand"].to(DEV), 8)
    else:
        xy, dem = engine.load_problems("tsp", synthetic_tsp_batch(1, N, seed=5).to(DEV), None, None, 8)
    batch = engine.encode(h, xy, dem)
    perm = list(range(M)) if kind == "tsp" else list(range(1, M + 1))
    tours, reward, _, n_steps = engine.rollout(batch, M, perm)
    torch.cuda.synchronize()
    print(kind, N, att, "T", int(n_steps.max()), "mean cost %.4f" % float(-reward.mean()), flush=True)
