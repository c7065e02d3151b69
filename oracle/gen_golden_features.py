#!/usr/bin/env python
"""Golden per-step feature tensors from the UNMODIFIED reference environments (TEST INFRASTRUCTURE ONLY; same rules
as gen_golden.py): CVRPEnv.get_cur_feature (CVRP/CVRPEnv.py:291-318) and TSPEnv.get_local_feature
(TSP/TSPEnv.py:135-156), recorded after hand-driven steps.  The cvrp instance has demands that are multiples of
1/8, so that some rows reach load == 0 exactly (norm_demand = 0/0 = nan at the depot, +inf at customers).

    python oracle/gen_golden_features.py        ->  tests/golden/features_cvrp.npz, features_tsp.npz
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ELG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def worker(problem):
    sys.path.insert(0, os.path.join(REF, problem.upper()))
    import numpy as np
    import torch
    g = torch.Generator().manual_seed(7)
    rec = {}
    if problem == "cvrp":
        from CVRPEnv import CVRPEnv
        n, N, M = 3, 12, 6
        depot, loc = torch.rand(n, 1, 2, generator=g), torch.rand(n, N, 2, generator=g)
        demand = torch.randint(1, 4, (n, N), generator=g).float() / 8          # 1/8, 2/8, 3/8: loads stay exact
        env = CVRPEnv(M, "cpu")
        env.load_random_problems({"depot": depot, "loc": loc, "demand": demand}, aug_factor=8)
        env.reset()
        B = 8 * n
        # rows visit customers m+1, m+2, ... until the next would not fit, then the depot, and so on (always feasible)
        sel = torch.zeros(B, M, dtype=torch.long)
        env.step(sel)
        nxt = torch.arange(M)[None, :].expand(B, M).clone()
        visited = torch.zeros(B, M, N + 1, dtype=torch.bool)
        dem_all = torch.cat((torch.zeros(B, 1), demand.repeat(8, 1)), dim=1)
        steps = []
        for t in range(1, 22):
            cand = (nxt % N) + 1
            seen = visited.gather(2, cand[:, :, None]).squeeze(2)
            tries = 0
            while seen.any() and tries < N:
                nxt = torch.where(seen, nxt + 1, nxt)
                cand = (nxt % N) + 1
                seen = visited.gather(2, cand[:, :, None]).squeeze(2)
                tries += 1
            fits = env.load + 1e-6 >= dem_all.gather(1, cand)
            allv = visited[:, :, 1:].all(dim=2)
            sel = torch.where(fits & ~allv & ~seen, cand, torch.zeros_like(cand))
            sel = torch.where((sel == 0) & (env.current_node == 0) & ~allv, cand, sel)      # never the depot twice unless finished
            env.step(sel)
            visited.scatter_(2, sel[:, :, None], True)
            visited[:, :, 0] = False
            cd, th, rel, nd = env.get_cur_feature()
            steps.append(t)
            rec["s%d_cur" % t] = env.current_node.numpy().astype(np.int16)
            rec["s%d_load" % t] = env.load.numpy().copy()
            rec["s%d_dist" % t], rec["s%d_theta" % t] = cd.numpy().copy(), th.numpy().copy()
            rec["s%d_rel" % t], rec["s%d_nd" % t] = rel.numpy().copy(), nd.numpy().copy()
        rec["depot"], rec["loc"], rec["demand"] = depot.numpy(), loc.numpy(), demand.numpy()
        rec["steps"] = np.array(steps)
        zero_loads = sum(int((rec["s%d_load" % t] == 0).sum()) for t in steps)
        print("cvrp: %d steps, rows with load == 0: %d, nan entries: %d" % (len(steps), zero_loads,
              sum(int(np.isnan(rec["s%d_nd" % t]).sum()) for t in steps)))
        assert zero_loads > 0
    else:
        from TSPEnv import TSPEnv
        n, N, M = 2, 15, 15
        problems = torch.rand(n, N, 2, generator=g)
        env = TSPEnv(M, "cpu")
        env.load_random_problems(problems, aug_factor=8)
        env.reset()
        B = 8 * n
        steps = []
        for t in range(6):
            sel = ((torch.arange(M)[None, :] + 3 * t) % N).expand(B, M).clone()
            env.step(sel)
            cd, th, rel = env.get_local_feature()
            steps.append(t)
            rec["s%d_cur" % t] = env.current_node.numpy().astype(np.int16)
            rec["s%d_dist" % t], rec["s%d_theta" % t], rec["s%d_rel" % t] = cd.numpy().copy(), th.numpy().copy(), rel.numpy().copy()
        rec["problems"] = problems.numpy()
        rec["steps"] = np.array(steps)
        print("tsp: %d steps" % len(steps))
    np.savez_compressed(os.path.join(OUT, "features_%s.npz" % problem), **rec)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        for p in ("cvrp", "tsp"):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", p], check=True)
