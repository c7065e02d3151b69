#!/usr/bin/env python
"""BASELINE.json configs that are not bench lines, measured once each (GPU, CUDA events; CPU oracle beside where it
finishes in about a minute):
  configs[0]  TSP100 ELG-POMO greedy, x8 augmentation, 64 synthetic uniform instances   (GPU + CPU oracle, parity)
  configs[3]  CVRP1000 generalisation, x8 augmentation, POMO width 1000                  (GPU; feasibility check)
    python tools/config_runs.py > gpurun_out/config_runs.json
"""
import json
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch   # noqa: E402
from oracle import elg_oracle as O            # noqa: E402  (checker / CPU baseline only)

DEV = "cuda:0"


def gpu_time(fn, warm=2, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep, out


def tsp100():
    from elg_b200.tsp import TSPEnv, TSPModel
    from elg_b200.tsp.test import solve_batch
    mp = dict(DEFAULT_MODEL_PARAMS["tsp"])
    sd = synthetic_state_dict("tsp", seed=1234, gain=3.0)
    torch.manual_seed(0)
    data = torch.rand(64, 100, 2)
    model = TSPModel(**mp)
    model.decoder.add_local_policy(DEV)
    model.load_state_dict(sd)
    model = model.to(DEV).eval().requires_grad_(False)
    env = TSPEnv(100, DEV)
    dev_data = data.to(DEV)

    def run():
        random.seed(0)
        return solve_batch(model, env, dev_data, 8)
    ms, (no_aug, aug, sol, rew) = gpu_time(run)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        prob = O.load_tsp(data, 8)
        ref_t, _, ref_r = O.rollout(O.Weights(sd, "tsp", mp), prob, 100, O.start_permutation("tsp", 100, 100, seed=0), "greedy")
        ref_no_aug, ref_aug = O.best_of(ref_r, 8, 64)
    cpu_s = time.perf_counter() - t0
    same = (sol.cpu() == ref_t).all(dim=2).float().mean()
    return {"config": "TSP100 ELG-POMO greedy, x8 aug, 64 synthetic uniform instances (torch.manual_seed(0)), POMO=100, seeded random-init weights",
            "gpu_ms": ms, "gpu_instances_per_s": 64 / ms * 1e3, "cpu_s": cpu_s, "cpu_instances_per_s": 64 / cpu_s,
            "cpu_threads": torch.get_num_threads(), "speedup": cpu_s * 1e3 / ms,
            "tours_identical_frac": float(same), "aug_cost_gpu": float(aug.mean()), "aug_cost_cpu": float(ref_aug.mean()),
            "max_abs_cost_diff_per_instance": float((aug.cpu() - ref_aug).abs().max())}


def cvrp1000(n=4):
    from elg_b200.cvrp import CVRPEnv, CVRPModel
    from elg_b200.cvrp.test import solve_batch
    mp = dict(DEFAULT_MODEL_PARAMS["cvrp"])
    sd = synthetic_state_dict("cvrp", seed=1234, gain=3.0)
    data = synthetic_cvrp_batch(n, 1000, seed=4)
    model = CVRPModel(**mp)
    model.decoder.add_local_policy(DEV)
    model.load_state_dict(sd)
    model = model.to(DEV).eval().requires_grad_(False)
    env = CVRPEnv(1000, DEV)
    dev = {k: v.to(DEV) for k, v in data.items()}

    def run():
        random.seed(0)
        return solve_batch(model, env, dev, 8)
    ms, (no_aug, aug, sol, rew) = gpu_time(run, warm=1, rep=2)
    prob = O.load_cvrp(data["depot"], data["loc"], data["demand"], 8)
    O.check_feasible_cvrp(sol[:2, :50].cpu(), prob.demand[:2])
    return {"config": "CVRP1000 generalisation, x8 aug, POMO width 1000, %d synthetic uniform instances, seeded random-init weights" % n,
            "gpu_ms": ms, "gpu_instances_per_s": n / ms * 1e3, "rollout_steps_T": int(sol.shape[2]), "ms_per_decode_step": ms / int(sol.shape[2]),
            "aug_cost": float(aug.mean()), "feasible": True}


def cvrp_xxl(N=7000):
    """A synthetic instance of CVRPLIB Set-XXL size (Antwerp2: 7000 customers) with the reference driver's POMO width
    min(N, 1000): the largest shapes of the streamed kernel (220 mask words per row, neighbour lists of 55 chunks)."""
    from elg_b200.cvrp import CVRPEnv, CVRPModel
    from elg_b200.cvrp.test import solve_batch
    mp = dict(DEFAULT_MODEL_PARAMS["cvrp"])
    sd = synthetic_state_dict("cvrp", seed=1234, gain=3.0)
    data = synthetic_cvrp_batch(1, N, seed=7)
    data["demand"] = data["demand"] * (50.0 / 700.0)          # ~ Set-XXL capacities: about 140 customers per route
    model = CVRPModel(**mp)
    model.decoder.add_local_policy(DEV)
    model.load_state_dict(sd)
    model = model.to(DEV).eval().requires_grad_(False)
    env = CVRPEnv(1000, DEV)
    dev = {k: v.to(DEV) for k, v in data.items()}
    random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    no_aug, aug, sol, rew = solve_batch(model, env, dev, 8)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    prob = O.load_cvrp(data["depot"], data["loc"], data["demand"], 8)
    O.check_feasible_cvrp(sol[:1, :8].cpu(), prob.demand[:1])
    return {"config": "synthetic CVRP%d (Set-XXL size), x8 aug, POMO width 1000, seeded random-init weights" % N, "seconds": secs,
            "rollout_steps_T": int(sol.shape[2]), "ms_per_decode_step": 1e3 * secs / int(sol.shape[2]), "aug_cost": float(aug.mean()),
            "feasible": True}


if __name__ == "__main__":
    out = {"tsp100": tsp100(), "cvrp1000": cvrp1000()}
    if "--xxl" in sys.argv:
        out["cvrp_xxl"] = cvrp_xxl()
    print(json.dumps(out, indent=1))
