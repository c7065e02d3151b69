#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (gaocrr/ELG) on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the reference is
mounted read-only at /root/reference; the GPU box has no reference, so the
outputs are committed as small fixtures under tests/golden/ and this script is
committed next to them.  Nothing under elg_b200/ imports this file.

    python oracle/gen_golden.py            # regenerate every fixture
    python oracle/gen_golden.py cvrp_n20   # one case

One subprocess per problem family, because CVRP/ and TSP/ both define top-level
modules called `models`, `utils` and `generate_data`.

What is recorded (per case):
  * the inputs (instances, weight seed/gain + checksum, POMO start permutation),
  * `encoded_nodes` of a few aug-instances (reference `CVRP/CVRPModel.py:32`,
    `TSP/TSPModel.py:22`),
  * for chosen steps: the Step_State *before* the decode (current node, load,
    ninf mask bits, finished) and the decoder's masked logits = the tensor fed
    to the final softmax (`CVRP/models.py:418-420`, `TSP/models.py:298-300`),
    captured by wrapping `F.softmax` inside the reference's `models` module,
  * the greedy tours (`rollout`, `CVRP/utils.py:7-29`, `TSP/utils.py:7-26`) and
    rewards, and for library-style cases the rounded unscaled cost
    (`CVRP/CVRPEnv.py:268-288`, `TSP/TSPEnv.py:174-184`).
"""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ELG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: problem, n instances, N nodes, M pomo width, aug, instance seed, weight seed, gain, recorded steps, recorded aug-instances
    "cvrp_n20": dict(problem="cvrp", n=3, N=20, M=20, aug=8, seed=11, wseed=1234, gain=1.0, steps="all", rows_b=[0, 1, 9, 23]),
    "cvrp_n20_sharp": dict(problem="cvrp", n=2, N=20, M=20, aug=8, seed=12, wseed=77, gain=6.0, steps="all", rows_b=[0, 5, 15]),
    "cvrp_n50": dict(problem="cvrp", n=2, N=50, M=50, aug=8, seed=13, wseed=1234, gain=3.0, steps=[2, 3, 10, 30, 55], rows_b=[0, 11]),
    "cvrp_n100": dict(problem="cvrp", n=1, N=100, M=100, aug=8, seed=1234, wseed=1234, gain=1.0, steps=[2, 7, 40, 100], rows_b=[0, 5]),
    "cvrp_n100_sharp": dict(problem="cvrp", n=1, N=100, M=100, aug=8, seed=99, wseed=5, gain=6.0, steps=[2, 50], rows_b=[3]),
    "cvrp_n20_noaug": dict(problem="cvrp", n=6, N=20, M=12, aug=1, seed=21, wseed=3, gain=2.0, steps=[2, 9], rows_b=[0, 5]),
    "cvrp_lib": dict(problem="cvrp", n=1, N=37, M=37, aug=8, seed=31, wseed=1234, gain=3.0, steps=[2, 20], rows_b=[0, 7], lib=True),
    "cvrp_x101": dict(problem="cvrp", n=1, N=100, M=100, aug=8, seed=71, wseed=1234, gain=3.0, steps=[2, 60], rows_b=[0, 6], lib=True, vrp_file="X-n101-k25"),
    "cvrp_x200": dict(problem="cvrp", n=1, N=199, M=199, aug=8, seed=72, wseed=1234, gain=3.0, steps=[2, 150], rows_b=[3], lib=True, vrp_file="X-n200-k36"),
    "cvrp_n200": dict(problem="cvrp", n=1, N=200, M=60, aug=8, seed=61, wseed=1234, gain=3.0, steps=[2, 90, 260], rows_b=[0, 5]),
    "tsp_n150": dict(problem="tsp", n=1, N=150, M=64, aug=8, seed=62, wseed=1234, gain=3.0, steps=[1, 70, 149], rows_b=[0, 5]),
    "tsp_n20": dict(problem="tsp", n=3, N=20, M=20, aug=8, seed=41, wseed=1234, gain=1.0, steps="all", rows_b=[0, 1, 9, 23]),
    "tsp_n20_sharp": dict(problem="tsp", n=2, N=20, M=20, aug=8, seed=42, wseed=77, gain=6.0, steps="all", rows_b=[0, 5, 15]),
    "tsp_n50": dict(problem="tsp", n=2, N=50, M=50, aug=8, seed=43, wseed=1234, gain=3.0, steps=[1, 2, 10, 30, 49], rows_b=[0, 11]),
    "tsp_n100": dict(problem="tsp", n=1, N=100, M=100, aug=8, seed=0, wseed=1234, gain=1.0, steps=[1, 7, 40, 80, 99], rows_b=[0, 5]),
    "tsp_n30_m10": dict(problem="tsp", n=4, N=30, M=10, aug=1, seed=44, wseed=3, gain=2.0, steps=[1, 15, 29], rows_b=[0, 3]),
    "tsp_lib": dict(problem="tsp", n=1, N=52, M=52, aug=8, seed=51, wseed=1234, gain=3.0, steps=[1, 30], rows_b=[0, 7], lib=True),
}


def worker(problem, names):
    """Runs inside a subprocess with the reference's sub-project on sys.path."""
    sys.path.insert(0, os.path.join(REF, problem.upper()))
    sys.path.insert(1, ROOT)
    import numpy as np
    import torch
    import models as ref_models
    from elg_b200.synth import (DEFAULT_MODEL_PARAMS, state_dict_checksum, synthetic_cvrp_batch,
                                synthetic_state_dict, synthetic_tsp_batch)
    from utils import rollout as ref_rollout
    if problem == "cvrp":
        from CVRPEnv import CVRPEnv as Env
        from CVRPModel import CVRPModel as Model
    else:
        from TSPEnv import TSPEnv as Env
        from TSPModel import TSPModel as Model

    torch.set_num_threads(8)
    for name in names:
        c = CASES[name]
        n, N, M, aug = c["n"], c["N"], c["M"], c["aug"]
        mp = dict(DEFAULT_MODEL_PARAMS[problem])
        sd = synthetic_state_dict(problem, seed=c["wseed"], gain=c["gain"])
        model = Model(**mp)
        model.decoder.add_local_policy("cpu")
        model.load_state_dict(sd)
        model.eval()
        model.requires_grad_(False)
        env = Env(M, "cpu")
        rec = {}
        if problem == "cvrp":
            batch = synthetic_cvrp_batch(n, N, seed=c["seed"])
            if c.get("vrp_file"):
                from elg_b200 import vrplib_io
                inst = vrplib_io.read_instance(os.path.join(OUT, "vrplib", c["vrp_file"] + ".vrp"))
                env.load_vrplib_problem(inst, aug_factor=aug)
                rec["lib_node_coord"], rec["lib_demand"] = inst["node_coord"], inst["demand"]
                rec["lib_capacity"] = np.array(inst["capacity"])
            elif c.get("lib"):
                # library-style instance: integer coordinates / demands, depot = node 0
                g = torch.Generator().manual_seed(c["seed"])
                coord = torch.randint(0, 1000, (N + 1, 2), generator=g).numpy().astype(np.float64)
                dem = torch.randint(1, 25, (N + 1,), generator=g).numpy().astype(np.float64)
                dem[0] = 0
                inst = {"node_coord": coord, "demand": dem, "capacity": 100, "depot": np.array([0])}
                env.load_vrplib_problem(inst, aug_factor=aug)
                rec["lib_node_coord"], rec["lib_demand"] = coord, dem
                rec["lib_capacity"] = np.array(100)
            else:
                env.load_random_problems(batch, aug_factor=aug)
                rec["depot"], rec["loc"], rec["demand"] = [batch[k].numpy() for k in ("depot", "loc", "demand")]
        else:
            problems = synthetic_tsp_batch(n, N, seed=c["seed"])
            if c.get("lib"):
                g = torch.Generator().manual_seed(c["seed"])
                coord = torch.randint(0, 2000, (N, 2), generator=g).numpy().astype(np.float64)
                unscaled = torch.tensor(coord, dtype=torch.float)[None]
                pts = (coord - np.min(coord)) / (np.max(coord) - np.min(coord))  # TSP/test_tsplib.py:128
                env.load_tsplib_problem(torch.tensor(pts, dtype=torch.float)[None], unscaled, aug)
                rec["lib_node_coord"] = coord
            else:
                env.load_random_problems(problems, aug_factor=aug)
                rec["problems"] = problems.numpy()

        steps = c["steps"]
        rows_b = c["rows_b"]
        st_rec = {}
        cur = {"t": 0, "logits": None}
        orig_softmax = ref_models.F.softmax

        def cap_softmax(x, dim=None, **kw):
            cur["logits"] = x.detach().clone()
            return orig_softmax(x, dim=dim, **kw)

        orig_step = model.one_step_rollout

        def cap_step(state, *a, **kw):
            t = cur["t"]
            cur["logits"] = None
            want = (steps == "all") or (t in steps)
            pre = None
            if want and state.current_node is not None:
                pre = dict(cur=state.current_node[rows_b].clone(), mask=state.ninf_mask[rows_b].clone())
                if problem == "cvrp":
                    pre["load"] = state.load[rows_b].clone()
                    pre["finished"] = state.finished[rows_b].clone()
            out = orig_step(state, *a, **kw)
            if pre is not None and cur["logits"] is not None:
                pre["logits"] = cur["logits"][rows_b]
                pre["selected"] = out[0][rows_b].clone()
                st_rec[t] = pre
            cur["t"] = t + 1
            return out

        ref_models.F.softmax = cap_softmax
        model.one_step_rollout = cap_step
        random.seed(c["seed"])
        reset_state, _, _ = env.reset()
        with torch.no_grad():
            model.pre_forward(reset_state)
            tours, _, reward = ref_rollout(model, env, "greedy")
        ref_models.F.softmax = orig_softmax

        rec["enc"] = model.encoded_nodes[rows_b].numpy()
        rec["tours"] = tours.numpy().astype(np.int16)
        rec["reward"] = reward.numpy()
        rec["perm"] = tours[0, :, 1 if problem == "cvrp" else 0].numpy().astype(np.int16)
        rec["rows_b"] = np.array(rows_b)
        ts = sorted(st_rec)
        rec["step_ids"] = np.array(ts)
        for t in ts:
            s = st_rec[t]
            rec["s%d_cur" % t] = s["cur"].numpy().astype(np.int16)
            rec["s%d_maskbits" % t] = np.packbits(torch.isinf(s["mask"]).numpy(), axis=-1, bitorder="little")
            rec["s%d_logits" % t] = s["logits"].numpy()
            rec["s%d_selected" % t] = s["selected"].numpy().astype(np.int16)
            if problem == "cvrp":
                rec["s%d_load" % t] = s["load"].numpy()
                rec["s%d_finished" % t] = s["finished"].numpy()
        meta = dict(c)
        meta.update(wsum=state_dict_checksum(sd), T=int(tours.shape[2]), torch=torch.__version__,
                    model_params=mp)
        rec["meta"] = np.array(json.dumps(meta))
        os.makedirs(OUT, exist_ok=True)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print("%-18s T=%d  mean best cost=%.4f  steps recorded=%d  bytes=%d" % (
            name, tours.shape[2], float((-reward).min(1)[0].mean()), len(ts),
            os.path.getsize(os.path.join(OUT, name + ".npz"))), flush=True)


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
