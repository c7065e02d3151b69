#!/usr/bin/env python
"""Time elg_encode alone (CUDA events on the launching stream): the encoder + decoder tables + neighbour lists of
`--batch` instances x 8 augmentations of the CVRP100 / TSP100 workload, inputs resident on the device."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from elg_b200 import engine
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch

ap = argparse.ArgumentParser()
ap.add_argument("--problem", default="cvrp")
ap.add_argument("--batch", type=int, default=1000)
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda:0")
mp = dict(DEFAULT_MODEL_PARAMS[args.problem])
handle = engine.ModelHandle(args.problem, mp, synthetic_state_dict(args.problem, seed=1234), dev)
if args.problem == "cvrp":
    d = {k: v.to(dev) for k, v in synthetic_cvrp_batch(args.batch, args.n, seed=100).items()}
    xy, dem = engine.load_problems("cvrp", d["loc"], d["depot"], d["demand"], aug=8)
else:
    xy, dem = engine.load_problems("tsp", synthetic_tsp_batch(args.batch, args.n, seed=100).to(dev), aug=8)
times = []
for i in range(args.iters + 2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    batch = engine.encode(handle, xy, dem)
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        times.append(e0.elapsed_time(e1))
    del batch
times.sort()
print(json.dumps({"problem": args.problem, "aug_instances": int(xy.shape[0]), "nodes": int(xy.shape[1]),
                  "encode_ms_median": times[len(times) // 2], "encode_ms_min": times[0]}))
