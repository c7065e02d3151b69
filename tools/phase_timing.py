#!/usr/bin/env python
"""Per-phase cycle breakdown of the rollout kernel (debug build only).

    ELG_NVCC_EXTRA=-DELG_PHASE_TIMING python -m elg_b200.build --force     # here
    gpurun -- python tools/phase_timing.py                                  # on the GPU box
    python -m elg_b200.build --force                                        # restore the shipped library

Thread 0 of every CTA accumulates clock64() deltas at the phase boundaries; the sum over CTAs is printed as
a share of the total and as microseconds per CTA-step."""
import ctypes as C
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from elg_b200 import _lib
from elg_b200.cvrp import CVRPEnv, CVRPModel
from elg_b200.cvrp.test import solve_batch
from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict

TC = os.environ.get("ELG_B200_ATTENTION") == "tensor"
NAMES = (["Q build", "softmax+P (4 rounds)", "B1 (2 passes)", "O operand", "B3", "select+C", "end barrier", "score MMA wait"] if TC else
         ["A", "A-barrier", "B1", "copy", "copy-barrier", "B3+C", "end-barrier", "-"])
dev = "cuda:0"
fn = _lib.lib.elg_debug_phase_clocks_tc if TC else _lib.lib.elg_debug_phase_clocks
model = CVRPModel(**dict(DEFAULT_MODEL_PARAMS["cvrp"]))
model.decoder.add_local_policy(dev)
model.load_state_dict(synthetic_state_dict("cvrp", seed=1234))
model = model.to(dev).eval().requires_grad_(False)
env = CVRPEnv(100, dev)
out = (C.c_ulonglong * 8)()
for i in range(2):
    random.seed(i)
    data = {k: v.to(dev) for k, v in synthetic_cvrp_batch(400, 100, seed=100 + i).items()}
    solve_batch(model, env, data, 8)
    torch.cuda.synchronize()
    fn(out, 1)
tot = float(sum(out))
for n, v in zip(NAMES, out):
    print("%-14s %6.2f %%" % (n, 100.0 * v / tot))
