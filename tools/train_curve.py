#!/usr/bin/env python
"""GPU: a short REINFORCE run from random initialisation through the reference-semantics driver (train_loop.train):
CVRP100, 64 instances x 100 rollouts per step, lr 1e-4, global-only warm-up for the first T steps then joint training.
Prints the sampled best-of-POMO tour length every 50 steps and the greedy validation costs (uniform N = 100 / 200 / 500).
    python tools/train_curve.py [steps] [T] [log_step]
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elg_b200.synth import DEFAULT_MODEL_PARAMS            # noqa: E402
from elg_b200.train_loop import train                      # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
log_step = int(sys.argv[3]) if len(sys.argv) > 3 else 250
torch.manual_seed(924); np.random.seed(924); random.seed(924)
config = {"name": "ELG_curve", "training": "joint", "seed": 924,
          "params": {"problem_size": 100, "multiple_width": 100, "scale_norm": True, "T": T, "start_steps": 0,
                     "train_steps": steps - 1, "mixed": False, "train_batch_size": 64, "learning_rate": 1e-4, "log_step": log_step},
          "distribution": {"data_type": "uniform", "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07},
          "model_params": dict(DEFAULT_MODEL_PARAMS["cvrp"])}
out_dir = os.path.join(ROOT, "gpurun_out", "train_curve")
os.makedirs(out_dir, exist_ok=True)
t0 = time.time()
tr, hist = train("cvrp", config, "cuda:0", dir_path=out_dir, log_path=os.path.join(out_dir, "log.json"),
                 val_samples=(256, 32, 8), verbose=True)
torch.cuda.synchronize()
wall = time.time() - t0
log = json.load(open(os.path.join(out_dir, "log.json")))
curve = [float(np.mean([h[1] for h in hist[i:i + 50]])) for i in range(0, len(hist), 50)]
print(json.dumps({"steps": len(hist), "T_switch": T, "wall_s": wall, "sampled_best_of_pomo_len_per_50_steps": curve,
                  "val": log["result"], "log_step": log_step}))
