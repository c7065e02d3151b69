#!/usr/bin/env python
"""Diagnostic for tests/test_gpu_sampler.py: per-node expected / observed counts and the kernel's own probabilities."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import Golden
from oracle import elg_oracle as O
from elg_b200 import engine
case, t, row = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
DEV = "cuda:0"; REPL, CALLS = 512, 200
g = Golden(case)
handle = engine.ModelHandle(g.kind, g.model_params(), g.state_dict(), DEV, attention="fp32")
prob = g.oracle_problem(); b0 = g.rows_b[0]; s = g.step(t)
cur, masked = s["cur"][:1], s["masked"][:1]
load = s["load"][:1] if g.kind == "cvrp" else None
first = g.tours()[b0:b0 + 1, :, 0] if g.kind == "tsp" else None
W = O.Weights(g.state_dict(), g.kind, g.model_params())
sub = O.Problem(prob.kind, prob.xy[b0:b0 + 1], None if prob.demand is None else prob.demand[b0:b0 + 1], prob.dist[b0:b0 + 1], None, 1)
cache = O.decoder_cache(W, O.encode(W, sub))
if g.kind == "tsp": O.set_first(W, cache, first)
logits = O.decode_logits(W, sub, cache, cur, masked, load)
p_ref = torch.softmax(logits.double(), dim=2)[0]
M, N1 = p_ref.shape
xy = sub.xy.expand(REPL, -1, -1).contiguous().to(DEV)
dem = None if sub.demand is None else sub.demand.expand(REPL, -1).contiguous().to(DEV)
batch = engine.encode(handle, xy, dem)
bits = engine.pack_mask_bits(masked.expand(REPL, -1, -1).contiguous().to(DEV))
curd = cur.expand(REPL, -1).contiguous().to(DEV)
loadd = None if load is None else load.expand(REPL, -1).contiguous().to(DEV)
firstd = None if first is None else first.expand(REPL, -1).contiguous().to(DEV)
counts = torch.zeros(M, N1, dtype=torch.float64, device=DEV)
psum = torch.zeros(M, N1, dtype=torch.float64, device=DEV)
_, _, klog = engine.decode_step(batch, M, curd, bits, load=loadd, first=firstd, mode="greedy", want_logits=True)
p_k = torch.softmax(klog[0].double().cpu(), dim=1)
for c in range(CALLS):
    sel, pr, _ = engine.decode_step(batch, M, curd, bits, load=loadd, first=firstd, mode="sample", seed=1000 + c, step=t)
    oh = torch.zeros(REPL, M, N1, device=DEV, dtype=torch.float64).scatter_(2, sel[:, :, None], 1.0)
    counts += oh.sum(0); psum += (oh * pr.double()[:, :, None]).sum(0)
counts, psum = counts.cpu(), psum.cpu()
n = REPL * CALLS
print("row", row, "replica-0 logits equal across replicas:", bool((klog == klog[:1]).all()))
for j in range(N1):
    if p_ref[row, j] > 0 or counts[row, j] > 0:
        print("node %3d  p_oracle %.6e  p_kernel(logits) %.6e  p_returned %.6e  expected %9.1f  observed %7d  z %.2f" % (
            j, p_ref[row, j], p_k[row, j], (psum[row, j] / counts[row, j]) if counts[row, j] else float("nan"), p_ref[row, j] * n, counts[row, j],
            (counts[row, j] - p_ref[row, j] * n) / max(1e-9, (p_ref[row, j] * n) ** 0.5)))
