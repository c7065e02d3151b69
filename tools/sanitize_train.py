#!/usr/bin/env python
"""One small training step (CVRP20 and TSP20, 3 instances) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python tools/sanitize_train.py
    compute-sanitizer --tool racecheck python tools/sanitize_train.py
"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch   # noqa: E402
from elg_b200.trainer import Trainer                                                                               # noqa: E402
from elg_b200 import generate_data as G                                                                            # noqa: E402

for kind in ("cvrp", "tsp"):
    tr = Trainer(kind, dict(DEFAULT_MODEL_PARAMS[kind]), synthetic_state_dict(kind, seed=3, gain=2.0), "cuda:0", chunk_steps=8)
    data = synthetic_cvrp_batch(3, 20, seed=1) if kind == "cvrp" else synthetic_tsp_batch(3, 20, seed=1)
    random.seed(1)
    out = tr.step(data, 20, seed=5)
    torch.cuda.synchronize()
    print(kind, "T", out["T"], "loss", float(out["loss"]), "|g|", float(out["grads"].norm()))
dist = {"data_type": "mixed", "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07}
d = G.generate_vrp_data(8, 50, dist, seed=1)
d2 = G.generate_tsp_data(8, 50, dict(dist, data_type="cluster"), seed=1)
torch.cuda.synchronize()
print("generators ok", float(d["loc"].mean()), float(d2.mean()))
