#!/usr/bin/env python
"""Run the CVRPLIB / TSPLIB drivers over a directory of instances with seeded random-init weights
(the released checkpoints are not available offline) and record per-instance wall time."""
import argparse
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--kind", choices=["vrplib", "tsplib"], required=True)
ap.add_argument("--path", required=True)
ap.add_argument("--limit", type=int, default=None)
ap.add_argument("--out", default="gpurun_out")
args = ap.parse_args()
problem = "cvrp" if args.kind == "vrplib" else "tsp"
cfg = {"name": "ELG_random_init", "use_cuda": True, "cuda_device_num": 0, "vrplib_set": "X", "load_checkpoint": None,
       "params": {"aug_factor": 8}, "model_params": dict(DEFAULT_MODEL_PARAMS[problem])}
random.seed(1234)
if problem == "cvrp":
    from elg_b200.cvrp import CVRPModel
    from elg_b200.cvrp.test_vrplib import VRPLib_Tester
    model = CVRPModel(**cfg["model_params"])
    model.decoder.add_local_policy("cuda:0")
    model.load_state_dict(synthetic_state_dict("cvrp", seed=1234, gain=3.0))
    cfg["vrplib_path"] = args.path
    res = VRPLib_Tester(cfg, model=model).test_on_vrplib(limit=args.limit, out_dir=args.out)
else:
    from elg_b200.tsp import TSPModel
    from elg_b200.tsp.test_tsplib import TSPLib_Tester
    model = TSPModel(**cfg["model_params"])
    model.decoder.add_local_policy("cuda:0")
    model.load_state_dict(synthetic_state_dict("tsp", seed=1234, gain=3.0))
    cfg["tsplib_path"] = args.path
    res = TSPLib_Tester(cfg, model=model).test_on_tsplib(limit=args.limit, out_dir=args.out)
rows = [(r["instance"], r["record"][0]["scale"], r["record"][0]["seconds"], r["record"][0]["best_cost"]) for r in res if "instance" in r]
print(json.dumps({"n": len(rows), "total_s": sum(r[2] for r in rows), "rows": rows}))
