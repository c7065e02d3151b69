#!/usr/bin/env python
"""A few CVRP100 REINFORCE steps (64 instances x 100 rollouts) for ncu launch lists / captures of the training path."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict   # noqa: E402
from elg_b200.trainer import Trainer                                                           # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda:0"
tr = Trainer("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]), synthetic_state_dict("cvrp", seed=1234, gain=1.0), dev, chunk_steps=128)
for it in range(steps):
    data = {k: v.to(dev) for k, v in synthetic_cvrp_batch(64, 100, seed=100 + it).items()}
    random.seed(it)
    out = tr.step(data, 100, seed=it)
    torch.cuda.synchronize()
    print("step", it, "loss", float(out["loss"]), "T", out["T"])
