#!/usr/bin/env python
"""Golden vectors for the TRAINING path: one REINFORCE step of the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/gen_golden.py for the rules; same layout: one subprocess per
problem family, fixtures under tests/golden/train_*.npz).

    python oracle/gen_golden_train.py

Per case the reference's own training-step body (CVRP/train.py:104-125, TSP/train.py:100-122) is executed
verbatim on a seeded batch: `model.pre_forward`, `rollout(..., eval_type='sample')`, the POMO shared
baseline, the max-advantage scaling, `J.backward()` and one `torch.optim.Adam(lr, weight_decay=1e-6)` step.
Recorded: the batch, the sampled tours, rewards, `log_prob`, `J`, and for every parameter tensor its gradient
sum, L2 norm and a strided sample of <= 512 entries (plus the same sample of the weights after the Adam step).
"""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ELG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
SAMPLE = 512

CASES = {
    "train_cvrp_n20": dict(problem="cvrp", n=6, N=20, M=20, seed=101, wseed=1234, gain=1.0, scale_norm=True),
    "train_cvrp_n50": dict(problem="cvrp", n=3, N=50, M=50, seed=102, wseed=7, gain=2.0, scale_norm=True),
    "train_cvrp_n100": dict(problem="cvrp", n=2, N=100, M=100, seed=103, wseed=1234, gain=1.0, scale_norm=True),
    # decoder without add_local_policy: the warm-up phase of `training: joint` (global policy + distance penalty)
    "train_cvrp_n20_global": dict(problem="cvrp", n=6, N=20, M=20, seed=104, wseed=1234, gain=1.0, scale_norm=True, local=False),
    "train_tsp_n20_global": dict(problem="tsp", n=6, N=20, M=20, seed=114, wseed=1234, gain=1.0, scale_norm=True, local=False),
    "train_tsp_n20": dict(problem="tsp", n=6, N=20, M=20, seed=111, wseed=1234, gain=1.0, scale_norm=True),
    "train_tsp_n50": dict(problem="tsp", n=3, N=50, M=50, seed=112, wseed=7, gain=2.0, scale_norm=True),
}


def sample_idx(numel):
    import numpy as np
    if numel <= SAMPLE:
        return np.arange(numel)
    return (np.arange(SAMPLE) * (numel // SAMPLE)).astype(np.int64)


def worker(problem, names):
    sys.path.insert(0, os.path.join(REF, problem.upper()))
    sys.path.insert(1, ROOT)
    import numpy as np
    import torch
    from torch.optim import Adam as Optimizer
    from elg_b200.synth import (DEFAULT_MODEL_PARAMS, state_dict_checksum, synthetic_cvrp_batch,
                                synthetic_state_dict, synthetic_tsp_batch)
    from utils import rollout
    if problem == "cvrp":
        from CVRPEnv import CVRPEnv as Env
        from CVRPModel import CVRPModel as Model
    else:
        from TSPEnv import TSPEnv as Env
        from TSPModel import TSPModel as Model
    torch.set_num_threads(8)
    for name in names:
        c = CASES[name]
        mp = dict(DEFAULT_MODEL_PARAMS[problem])
        sd = synthetic_state_dict(problem, seed=c["wseed"], gain=c["gain"])
        model = Model(**mp)
        if c.get("local", True):
            model.decoder.add_local_policy("cpu")
        else:
            sd = {k: v for k, v in sd.items() if ".local_polic" not in k}
        model.load_state_dict(sd)
        env = Env(c["M"], "cpu")
        optimizer = Optimizer(model.parameters(), lr=1e-4, weight_decay=1e-6)
        rec = {}
        if problem == "cvrp":
            batch = synthetic_cvrp_batch(c["n"], c["N"], seed=c["seed"])
            rec["depot"], rec["loc"], rec["demand"] = [batch[k].numpy() for k in ("depot", "loc", "demand")]
        else:
            batch = synthetic_tsp_batch(c["n"], c["N"], seed=c["seed"])
            rec["problems"] = batch.numpy()
        random.seed(c["seed"])
        torch.manual_seed(c["seed"])
        # ---- body of the reference's training loop -------------------------------------------
        model.train()
        env.load_random_problems(batch)
        reset_state, _, _ = env.reset()
        model.pre_forward(reset_state)
        solutions, probs, rewards = rollout(model=model, env=env, eval_type='sample')
        optimizer.zero_grad()
        bl_val = rewards.mean(dim=1)[:, None]
        log_prob = probs.log().sum(dim=1)
        advantage = rewards - bl_val
        J = - advantage * log_prob
        if c["scale_norm"]:
            norm_fac = advantage.max(dim=1)[0][:, None]
            if problem == "cvrp" or (norm_fac != 0.).all():
                J = J / norm_fac
        J = J.mean()
        J.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        optimizer.step()
        # ---------------------------------------------------------------------------------------
        rec["tours"] = solutions.numpy().astype(np.int16)
        rec["reward"] = rewards.detach().numpy()
        rec["log_prob"] = log_prob.detach().numpy()
        rec["J"] = np.array(float(J))
        new_sd = model.state_dict()
        keys = sorted(grads)
        for k in keys:
            g = grads[k].reshape(-1).double()
            idx = sample_idx(g.numel())
            rec["g_sum/" + k] = np.array(float(g.sum()))
            rec["g_norm/" + k] = np.array(float(g.norm()))
            rec["g_sample/" + k] = grads[k].reshape(-1).numpy()[idx]
            rec["w_after/" + k] = new_sd[k].reshape(-1).numpy()[idx]
        meta = dict(c)
        meta.update(wsum=state_dict_checksum(sd), T=int(solutions.shape[2]), torch=torch.__version__, model_params=mp,
                    keys=keys, lr=1e-4, weight_decay=1e-6)
        rec["meta"] = np.array(json.dumps(meta))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print("%-18s T=%d J=%.6f |g|=%.4e bytes=%d" % (name, solutions.shape[2], float(J),
              float(torch.cat([g.reshape(-1) for g in grads.values()]).norm()), os.path.getsize(path)), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        worker(sys.argv[2], sys.argv[3:])
        return
    want = sys.argv[1:] or list(CASES)
    for problem in ("cvrp", "tsp"):
        names = [k for k in want if CASES[k]["problem"] == problem]
        if names:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", problem] + names, check=True)


if __name__ == "__main__":
    main()
