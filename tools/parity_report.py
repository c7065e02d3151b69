#!/usr/bin/env python
"""Parity numbers of the CUDA path against the reference goldens and the fp64 oracle (run on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import ALL_CASES, Golden, compare_tours, sub_problem, tie_rows, top2_margin
from oracle import elg_oracle as O
from elg_b200 import engine

DEV = "cuda:0"
print("| case | N | rows | logits max / mean abs err vs reference | ... vs fp64 oracle (ours) | ... (reference vs fp64) | greedy flips / decisions (teacher-forced) | free-running identical tours | reward rel err (matched rows) |")
print("|---|---|---|---|---|---|---|---|---|")
for name in ALL_CASES:
    g = Golden(name)
    sd, mp = g.state_dict(), g.model_params()
    handle = engine.ModelHandle(g.kind, mp, sd, DEV)
    prob = g.oracle_problem()
    prob64 = g.oracle_problem(torch.float64)
    W64 = O.Weights(sd, g.kind, mp, torch.float64)
    rows = g.rows_b
    sub = engine.encode(handle, prob.xy[rows].to(DEV), None if prob.demand is None else prob.demand[rows].to(DEV))
    sub64 = sub_problem(prob64, rows)
    cache64 = O.decoder_cache(W64, O.encode(W64, sub64))
    first = g.tours()[rows][:, :, 0] if g.kind == "tsp" else None
    if g.kind == "tsp":
        O.set_first(W64, cache64, first)
    e_ref, e_64, r_64, flips, dec = [], [], [], 0, 0
    for t in g.steps:
        s = g.step(t)
        live = ~tie_rows(sub_problem(prob, rows), g.kind, s["cur"], s["masked"], mp["local_size"][0])
        bits = engine.pack_mask_bits(s["masked"].to(DEV))
        sel, _, lg = engine.decode_step(sub, g.M, s["cur"].to(DEV), bits, load=None if g.kind == "tsp" else s["load"].to(DEV),
                                        first=None if first is None else first.to(DEV), want_logits=True)
        lg, sel = lg.cpu(), sel.cpu()
        l64 = O.decode_logits(W64, sub64, cache64, s["cur"], s["masked"], None if g.kind == "tsp" else s["load"].double())
        fin = ~torch.isinf(s["logits"]) & live[:, :, None]
        e_ref.append((lg - s["logits"])[fin].abs())
        e_64.append((lg.double() - l64)[fin].abs())
        r_64.append((s["logits"].double() - l64)[fin].abs())
        flips += int((sel[live] != s["selected"][live]).sum())
        dec += int(live.sum())
    e_ref, e_64, r_64 = torch.cat(e_ref), torch.cat(e_64), torch.cat(r_64)
    batch = engine.encode(handle, prob.xy.to(DEV), None if prob.demand is None else prob.demand.to(DEV))
    if g.meta.get("lib"):
        un = prob.unscaled_xy.expand(batch.B, -1, -1).contiguous().to(DEV)
        batch.tables.unscaled = un.data_ptr()
    tours16, reward, _, n_steps = engine.rollout(batch, g.M, g.perm().tolist())
    T = int(n_steps.max())
    frac, same = compare_tours(tours16[:, :, :T].long().cpu(), g.tours())
    rel = float(((reward.cpu() - g.reward()).abs() / g.reward().abs())[same].max())
    print("| %s | %d | %d | %.1e / %.1e | %.1e / %.1e | %.1e / %.1e | %d / %d | %.2f %% | %.1e |" % (
        name, g.meta["N"], same.numel(), e_ref.max(), e_ref.mean(), e_64.max(), e_64.mean(), r_64.max(), r_64.mean(),
        flips, dec, 100 * frac, rel))
