#!/usr/bin/env python
"""Tiny driver for ncu captures: a few evaluation batches of the CVRP100 workload (no timing output)."""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from elg_b200.cvrp import CVRPEnv, CVRPModel
from elg_b200.cvrp.test import solve_batch
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=100)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--n", type=int, default=100)
args = ap.parse_args()
dev = "cuda:0"
mp = dict(DEFAULT_MODEL_PARAMS["cvrp"])
model = CVRPModel(**mp)
model.decoder.add_local_policy(dev)
model.load_state_dict(synthetic_state_dict("cvrp", seed=1234))
model = model.to(dev).eval().requires_grad_(False)
env = CVRPEnv(args.n, dev)
for i in range(args.iters):
    random.seed(i)
    data = {k: v.to(dev) for k, v in synthetic_cvrp_batch(args.batch, args.n, seed=100 + i).items()}
    no_aug, aug, sol, rew = solve_batch(model, env, data, 8)
    torch.cuda.synchronize()
    print("iter", i, "aug cost", float(aug.mean()), "T", sol.shape[2])
