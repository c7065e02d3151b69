#!/usr/bin/env python
"""Per-phase cycle breakdown of the streamed tensor-core rollout kernel (debug build: `python -m elg_b200.build --variant
timing`; run with ELG_B200_LIB=elg_b200/csrc/libelg_b200_timing.so).  Usage: phase_timing_stc.py [cvrp|tsp] [N] [instances]"""
import ctypes as C
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from elg_b200 import _lib
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch

kind = sys.argv[1] if len(sys.argv) > 1 else "cvrp"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
NAMES = ["Q operand + tile-0 TMA issue", "neighbour search (128-entry chunks)", "features", "local scores / weights -> TMEM, MMA1 issue",
         "MMA1 wait + ol + MMA2 issue", "local finalize", "key tiles (online softmax, P V)", "O operand", "score tiles (arg-max)",
         "neighbour logits + select", "phase C (env step, masks)", "end barrier"]
dev = "cuda:0"
if kind == "cvrp":
    from elg_b200.cvrp import CVRPEnv as Env, CVRPModel as Model
    from elg_b200.cvrp.test import solve_batch
    data = {k: v.to(dev) for k, v in synthetic_cvrp_batch(n, N, seed=4).items()}
else:
    from elg_b200.tsp import TSPEnv as Env, TSPModel as Model
    from elg_b200.tsp.test import solve_batch
    data = synthetic_tsp_batch(n, N, seed=4).to(dev)
model = Model(**dict(DEFAULT_MODEL_PARAMS[kind]))
model.decoder.add_local_policy(dev)
model.load_state_dict(synthetic_state_dict(kind, seed=1234, gain=3.0))
model = model.to(dev).eval().requires_grad_(False)
env = Env(min(N, 1000), dev)
out = (C.c_ulonglong * 16)()
fn = _lib.lib.elg_debug_phase_clocks_stc
random.seed(0)
solve_batch(model, env, data, 8)
torch.cuda.synchronize()
fn(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
random.seed(0)
e0.record()
res = solve_batch(model, env, data, 8)
e1.record()
torch.cuda.synchronize()
fn(out, 1)
tot = float(sum(out))
print("%s N=%d, %d instance(s) x 8: %.1f ms, T=%d" % (kind, N, n, e0.elapsed_time(e1), res[2].shape[2]))
for nm, v in zip(NAMES, out):
    print("%-46s %6.2f %%" % (nm, 100.0 * v / tot))
