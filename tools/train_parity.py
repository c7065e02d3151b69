#!/usr/bin/env python
"""GPU: gradient of the CUDA training path against the oracle's autograd and the reference fixtures, stage by stage.

    python tools/train_parity.py [case ...]          (cases: tests/golden/train_*.npz)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from elg_b200 import engine                                  # noqa: E402
from elg_b200.trainer import Trainer                         # noqa: E402
import train_helpers as TH                                   # noqa: E402


def run(name, verbose=True):
    g = TH.TrainGolden(name)
    dev = "cuda:0"
    mp = g.meta["model_params"]
    sd = g.state_dict()
    tr = Trainer(g.kind, mp, sd, dev, scale_norm=g.meta["scale_norm"])
    data = g.data()
    cv = g.kind == "cvrp"
    if cv:
        xy, dem = engine.load_problems("cvrp", data["loc"].to(dev), data["depot"].to(dev), data["demand"].to(dev), 1)
    else:
        xy, dem = engine.load_problems("tsp", data.to(dev), None, None, 1)
    B, N1 = int(xy.shape[0]), int(xy.shape[1])
    batch, saved = engine.encode_train(tr.handle, xy, dem)
    tours = g.tours()
    T = tours.shape[2]
    t_max = 2 * N1 + 2 if cv else N1
    tours16 = TH.pad_tours(tours, t_max).to(dev)
    reward = g.reward().to(dev)
    grads, loss, ws = engine.reinforce_backward(batch, saved, g.M, tours16, T, reward, None, g.meta["scale_norm"], 16)
    torch.cuda.synchronize()
    mine = tr.unpack(grads)
    J, logp, ref, extra = TH.oracle_grads(g.kind, mp, sd, g.problem(), g.M, tours, g.reward(), g.meta["scale_norm"], keep=True)
    lay = engine.train_workspace_layout(tr.handle, B, g.M, N1, t_max)
    wsf = ws.view(torch.float32)
    rows = B * N1
    ck = 0.36067376022224085
    dK = wsf[lay["dK"]:lay["dK"] + rows * 128].reshape(B, N1, 8, 16).permute(0, 2, 1, 3).cpu() * ck
    dV = wsf[lay["dV"]:lay["dV"] + rows * 128].reshape(B, N1, 8, 16).permute(0, 2, 1, 3).cpu()
    def rel(a, b):
        return float((a - b).abs().max() / (b.abs().max() + 1e-30))
    print("== %s  B=%d N1=%d M=%d T=%d  J(oracle)=%.6f J(ref)=%.6f" % (name, B, N1, g.M, T, float(J), float(g.z["J"])))
    print("   table grads vs oracle autograd: dK rel %.2e   dV rel %.2e" % (rel(dK, extra["k"]), rel(dV, extra["v"])))
    eo = TH.grad_errors(mine, ref)
    ef = TH.fixture_errors(g, mine)
    worst = 0.0
    for k in sorted(eo):
        worst = max(worst, eo[k], ef[k])
        if verbose or eo[k] > 1e-2:
            print("   %-58s vs oracle %.2e   vs reference %.2e" % (k, eo[k], ef[k]))
    print("   worst error (x tensor rms): %.3e" % worst)
    return worst


if __name__ == "__main__":
    names = sys.argv[1:] or ["train_cvrp_n20"]
    w = [run(n) for n in names]
    print("WORST", max(w))
