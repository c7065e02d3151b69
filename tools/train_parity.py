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
    mp = g.model_params()
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


def run_own(kind, N, n, M, gain, wseed=5):
    """Own sampled rollout; ours and the fp32 oracle both measured against the fp64 oracle on the same tours."""
    import random
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch
    from oracle import elg_oracle as O
    mp = dict(DEFAULT_MODEL_PARAMS[kind])
    sd = synthetic_state_dict(kind, seed=wseed, gain=gain)
    tr = Trainer(kind, mp, sd, "cuda:0")
    if kind == "cvrp":
        data = synthetic_cvrp_batch(n, N, seed=9)
        prob = lambda dt: O.load_cvrp(data["depot"], data["loc"], data["demand"], 1, dt)
    else:
        data = synthetic_tsp_batch(n, N, seed=9)
        prob = lambda dt: O.load_tsp(data, 1, dt)
    random.seed(4)
    out = tr.forward_backward(data, M, seed=123)
    torch.cuda.synchronize()
    T = out["T"]
    tours = out["tours"][:, :, :T].long().cpu()
    reward = out["reward"].cpu()
    mine = tr.unpack(out["grads"])
    J32, lp32, g32, _ = TH.oracle_grads(kind, mp, sd, prob(torch.float32), M, tours, reward, True)
    J64, lp64, g64, _ = TH.oracle_grads(kind, mp, sd, prob(torch.float64), M, tours, reward.double(), True, dtype=torch.float64)
    e_m = TH.grad_errors(mine, g64)
    e_o = TH.grad_errors(g32, g64)
    print("== own %s N=%d n=%d M=%d gain=%.1f T=%d: J ours %.6f fp32 %.6f fp64 %.6f; logp err ours %.2e fp32 %.2e" % (
        kind, N, n, M, gain, T, float(out["loss"]), float(J32), float(J64),
        float((out["logp"].cpu().double() - lp64).abs().max()), float((lp32.double() - lp64).abs().max())))
    for k in sorted(e_m, key=lambda k: -e_m[k])[:8]:
        print("   %-58s ours vs fp64 %.2e   fp32 oracle vs fp64 %.2e" % (k, e_m[k], e_o[k]))
    l2 = TH.grad_errors_l2(mine, g64)
    print("   worst Frobenius error: %s" % str(sorted(l2.items(), key=lambda kv: -kv[1])[:3]))
    k = max(e_m, key=lambda k: e_m[k])
    d = (mine[k].double() - g64[k]).abs()
    if d.dim() == 2:
        rowerr = d.pow(2).sum(1)
        top = rowerr.topk(3)
        print("   %s: error energy by output row: top3 rows %s hold %.1f%% of the squared error" % (
            k, top.indices.tolist(), 100 * float(top.values.sum() / rowerr.sum())))
    return max(e_m.values())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--own":
        kind, N, n, M, gain = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
        print("WORST", run_own(kind, N, n, M, gain))
    else:
        names = sys.argv[1:] or ["train_cvrp_n20"]
        w = [run(n, verbose=False) for n in names]
        print("WORST", max(w))
