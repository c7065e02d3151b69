#!/usr/bin/env python
"""Greedy rollout of FEW aug-instances (fewer than SMs): fp32-pipe kernel vs the tensor-core kernel with its rows split over
1 / 2 / 3 CTAs per aug-instance (ELG_TC_TILES, read by rollout_tc_tiles).  CUDA events around elg_rollout.

    python tools/small_batch_timing.py [aug_instances] [M]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import torch
    from elg_b200 import engine
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict
    B, M, att = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    dev = torch.device("cuda:0")
    h = engine.ModelHandle("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]), synthetic_state_dict("cvrp", seed=1234), dev, attention=att)
    d = {k: v.to(dev) for k, v in synthetic_cvrp_batch(B, 100, seed=100).items()}
    xy, dem = engine.load_problems("cvrp", d["loc"], d["depot"], d["demand"], aug=1)
    batch = engine.encode(h, xy, dem)
    perm = list(range(1, M + 1))
    ts = []
    for i in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tours, reward, _, n_steps = engine.rollout(batch, M, perm)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-6s tiles=%-4s %7.3f ms  T=%d  mean cost %.5f" % (att, os.environ.get("ELG_TC_TILES", "auto"), sorted(ts)[2],
                                                            int(n_steps.max()), float(-reward.mean())))
else:
    B = sys.argv[1] if len(sys.argv) > 1 else "64"
    M = sys.argv[2] if len(sys.argv) > 2 else "100"
    for att, tiles in (("fp32", None), ("tensor", "1"), ("tensor", "2"), ("tensor", "3"), ("auto", None)):
        env = dict(os.environ)
        if tiles:
            env["ELG_TC_TILES"] = tiles
        subprocess.run([sys.executable, __file__, "--child", B, M, att], env=env)
