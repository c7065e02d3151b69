#!/usr/bin/env python
"""Library-set parity: our drivers (elg_b200.tsp.test_tsplib / elg_b200.cvrp.test_vrplib, CUDA path) against the
per-instance results of the UNMODIFIED reference drivers recorded in tests/golden/lib (oracle/gen_golden_lib.py).

    gpurun -- python tools/library_parity.py [--out profiles/r02_library_parity]

Same instances (tests/golden/lib/*_inputs.npz), same seeded synthetic checkpoint, same `random.seed` before every instance
(the POMO start permutation).  For every instance: best cost / gap of both sides, the eight per-augmentation best costs,
the sum of all row rewards; for the "detail" instances every row reward and the tours of a row subset; for the tie-order
instances our tours against the reference run twice (unmodified, and with index-ordered distance ties).
Writes <out>.md (table, all rows) and <out>.json.  tests/test_gpu_library.py runs the same comparison as assertions."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tests", "golden", "lib")


def _config(problem):
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    return {"name": "ELG", "use_cuda": True, "cuda_device_num": 0, "vrplib_set": "X", "training": "joint", "load_checkpoint": None,
            "params": {"aug_factor": 8}, "model_params": dict(DEFAULT_MODEL_PARAMS[problem])}


def _model(problem, ref):
    from elg_b200.synth import state_dict_checksum, synthetic_state_dict
    sd = synthetic_state_dict(problem, seed=ref["wseed"], gain=ref["gain"])
    assert state_dict_checksum(sd) == ref["wsum"], "synthetic weights drifted from the fixture"
    if problem == "cvrp":
        from elg_b200.cvrp import CVRPModel as Model
    else:
        from elg_b200.tsp import TSPModel as Model
    m = Model(**_config(problem)["model_params"])
    m.decoder.add_local_policy("cuda:0")
    m.load_state_dict(sd)
    return m


def _compare(name, ref_row, res, rewards, tours, secs, stable_row=None):
    import torch
    rew = rewards.cpu()
    M = rew.shape[1]
    per_aug = (-rew).reshape(8, M).min(dim=1)[0]
    row = dict(instance=name, scale=ref_row["scale"], optimal=ref_row["optimal"], ref_best=ref_row["best_cost"],
               our_best=float(res["best_cost"]), ref_gap=ref_row["gap"], our_gap=float(res["gap"]),
               per_aug_equal=int(sum(float(a) == float(b) for a, b in zip(per_aug, ref_row["per_aug_best"]))),
               reward_sum_rel=abs(float(rew.double().sum()) - ref_row["reward_sum"]) / abs(ref_row["reward_sum"]),
               T_ref=ref_row["T"], T_ours=int(tours.shape[2]), seconds=round(secs, 3), ref_cpu_seconds=ref_row["seconds"])
    detail = os.path.join(LIB, "detail_%s.npz" % name)
    if os.path.exists(detail):
        z = np.load(detail)
        ref_rew = torch.tensor(z["reward"])
        row["rows_reward_equal"] = float((ref_rew == rew).float().mean())
        rt = torch.tensor(z["tours_rows"].astype(np.int64))
        ot = tours[:, :rt.shape[1]].cpu()
        T = max(rt.shape[2], ot.shape[2])
        pa = torch.zeros(rt.shape[0], rt.shape[1], T, dtype=torch.long); pa[:, :, :rt.shape[2]] = rt
        pb = torch.zeros_like(pa); pb[:, :, :ot.shape[2]] = ot
        row["rows_tour_equal"] = float((pa == pb).all(dim=2).float().mean())
    if stable_row is not None:
        row["stable_best"] = stable_row["best_cost"]
        row["stable_gap"] = stable_row["gap"]
        row["per_aug_equal_stable"] = int(sum(float(a) == float(b) for a, b in zip(per_aug, stable_row["per_aug_best"])))
        row["reward_sum_rel_stable"] = abs(float(rew.double().sum()) - stable_row["reward_sum"]) / abs(stable_row["reward_sum"])
    ties = os.path.join(LIB, "ties_%s.npz" % name)
    if os.path.exists(ties):
        z = np.load(ties)
        for key in ("unmodified", "stable"):
            rt = torch.tensor(z["tours_" + key].astype(np.int64))
            ot = tours.cpu()
            T = max(rt.shape[2], ot.shape[2])
            pa = torch.zeros(rt.shape[0], rt.shape[1], T, dtype=torch.long); pa[:, :, :rt.shape[2]] = rt
            pb = torch.zeros_like(pa); pb[:, :, :ot.shape[2]] = ot
            row["rows_tour_equal_vs_" + key] = float((pa == pb).all(dim=2).float().mean())
        row["ref_unmodified_vs_stable"] = json.loads(str(z["meta"]))["rows_identical"]
    return row


def run_set(kind, names=None):
    """kind = 'tsplib' | 'setx' -> list of per-instance comparison rows (reference order: by size)."""
    import torch
    problem = "tsp" if kind == "tsplib" else "cvrp"
    with open(os.path.join(LIB, kind + "_ref.json")) as f:
        ref = json.load(f)
    inputs = np.load(os.path.join(LIB, kind + "_inputs.npz"))
    stable = {}
    if os.path.exists(os.path.join(LIB, kind + "_ref_stable.json")):
        with open(os.path.join(LIB, kind + "_ref_stable.json")) as f:
            stable = {r["instance"]: r for r in json.load(f)["instances"]}
    model = _model(problem, ref)
    cfg = _config(problem)
    if problem == "cvrp":
        from elg_b200.cvrp.test_vrplib import VRPLib_Tester
        tester = VRPLib_Tester(cfg, model=model)
    else:
        from elg_b200.tsp.test_tsplib import TSPLib_Tester
        tester = TSPLib_Tester(cfg, model=model)
    rows = []
    for r in ref["instances"]:
        name = r["instance"]
        if names and name not in names:
            continue
        res = {}
        random.seed(ref["seed"])
        torch.cuda.synchronize()
        t0 = time.time()
        if problem == "cvrp":
            cap, opt = inputs[name + "/capopt"]
            inst = {"node_coord": inputs[name + "/coord"].astype(np.float64), "demand": inputs[name + "/demand"].astype(np.float64),
                    "capacity": float(cap), "depot": np.array([0])}
            tours, rewards = tester.test_on_one_ins(name=name, result_dict=res, instance=inst, solution=float(opt))
        else:
            tours, rewards = tester.test_on_one_ins(name=name, result_dict=res, instance=[inputs[name + "/coord"], float(inputs[name + "/opt"])])
        torch.cuda.synchronize()
        rows.append(_compare(name, r, res, rewards, tours, time.time() - t0, stable.get(name)))
    return rows, ref


def bins(kind, rows, key):
    """Mean gap per size bin exactly as the reference drivers print them (CVRP/test_vrplib.py:86-104, TSP/test_tsplib.py:95-124)."""
    scale = np.array([r["scale"] for r in rows])
    gap = np.array([r[key] for r in rows])
    edges = (("<=200", scale <= 200), ("200-500", (scale > 200) & (scale <= 500)), (">500", scale > 500), ("all", scale > 0))
    return {k: float(100 * gap[m].mean()) for k, m in edges if m.any()}


def summarize(kind, rows):
    n = len(rows)
    eq = sum(r["ref_best"] == r["our_best"] for r in rows)
    st = [r for r in rows if "stable_best" in r]
    return dict(set=kind, instances=n, best_cost_equal=eq, per_aug_equal=sum(r["per_aug_equal"] for r in rows), per_aug_total=8 * n,
                stable_instances=len(st), stable_best_cost_equal=sum(r["our_best"] == r["stable_best"] for r in st),
                stable_per_aug_equal=sum(r["per_aug_equal_stable"] for r in st), stable_per_aug_total=8 * len(st),
                gap_bins_stable=bins(kind, st, "stable_gap") if st else None,
                gap_bins_ours_on_stable_subset=bins(kind, st, "our_gap") if st else None,
                max_rel_best_diff=max(abs(r["our_best"] - r["ref_best"]) / r["ref_best"] for r in rows),
                gap_bins_ref=bins(kind, rows, "ref_gap"), gap_bins_ours=bins(kind, rows, "our_gap"),
                gpu_seconds=sum(r["seconds"] for r in rows), ref_cpu_seconds=sum(r["ref_cpu_seconds"] for r in rows))


def markdown(results):
    out = ["# Library-set parity (round 2): reference drivers on CPU vs elg_b200 on one B200", "",
           "Reference = the unmodified `TSP/test_tsplib.py:126-162` / `CVRP/test_vrplib.py:111-145` run in the build container "
           "(`oracle/gen_golden_lib.py`, seeded synthetic checkpoint, gain 3 — the released checkpoints are not available offline, "
           "so the gaps are those of an untrained policy; what is compared is reference vs ours).  Costs are the rounded "
           "unscaled tour lengths of the best of 8 augmentations x M POMO rows.  `aug=` counts how many of the eight "
           "per-augmentation best costs are identical.", "",
           "**Tie order.**  Library coordinates are integers, so many node pairs are exactly equidistant; the reference breaks such "
           "ties by whatever order `torch.topk`'s partial sort leaves, and its local policy is rank-aware.  The column "
           "`ref (index ties)` is the reference run once more with the ONLY change that `torch.topk` is a stable sort (equal "
           "distances in node-index order, `oracle/gen_golden_lib.py stable_*`): it differs from the unmodified reference on the "
           "lattice-like instances (pr*, u*, ts225, a280, pcb442, ...) exactly where we do, and our costs equal it.", ""]
    for kind, (rows, summ) in results.items():
        out += ["## %s: %d instances" % (kind, summ["instances"]), "",
                "* vs the unmodified reference: best cost identical on %d / %d instances, per-augmentation best on %d / %d." % (
                    summ["best_cost_equal"], summ["instances"], summ["per_aug_equal"], summ["per_aug_total"]),
                "* vs the reference with index-ordered ties: best cost identical on **%d / %d** instances, per-augmentation best on %d / %d; "
                "mean gaps (%%) reference %s | ours %s." % (summ["stable_best_cost_equal"], summ["stable_instances"], summ["stable_per_aug_equal"],
                                                          summ["stable_per_aug_total"], json.dumps(summ["gap_bins_stable"]),
                                                          json.dumps(summ["gap_bins_ours_on_stable_subset"])), "",
            "Mean gap per size bin (%%): reference %s | ours %s.  Time: reference %.0f s on 5-6 CPU threads, ours %.1f s." % (
                json.dumps(summ["gap_bins_ref"]), json.dumps(summ["gap_bins_ours"]), summ["ref_cpu_seconds"], summ["gpu_seconds"]), "",
            "| instance | N | optimum | reference best | ref (index ties) | our best | ref gap | our gap | aug= (ref / index ties) | rows identical | T ref/ours | s (ours) |",
            "|---|---|---|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            extra = ""
            if "rows_tour_equal" in r:
                extra = "tours %.1f %%, rewards %.1f %%" % (100 * r["rows_tour_equal"], 100 * r["rows_reward_equal"])
            if "rows_tour_equal_vs_stable" in r:
                extra = "vs unmodified %.2f %%, vs index-ordered ties %.2f %% (reference vs itself %.2f %%)" % (
                    100 * r["rows_tour_equal_vs_unmodified"], 100 * r["rows_tour_equal_vs_stable"], 100 * r["ref_unmodified_vs_stable"])
            out.append("| %s | %d | %.0f | %.0f | %s | %.0f | %.4f | %.4f | %d / %s | %s | %d/%d | %.3f |" % (
                r["instance"], r["scale"], r["optimal"], r["ref_best"], ("%.0f" % r["stable_best"]) if "stable_best" in r else "-",
                r["our_best"], r["ref_gap"], r["our_gap"], r["per_aug_equal"], r.get("per_aug_equal_stable", "-"),
                extra, r["T_ref"], r["T_ours"], r["seconds"]))
        out.append("")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_library_parity"))
    ap.add_argument("--sets", default="tsplib,setx")
    args = ap.parse_args()
    results = {}
    for kind in args.sets.split(","):
        rows, _ = run_set(kind)
        results[kind] = (rows, summarize(kind, rows))
        print(json.dumps(results[kind][1]))
    with open(args.out + ".md", "w") as f:
        f.write(markdown(results))
    with open(args.out + ".json", "w") as f:
        json.dump({k: {"summary": v[1], "rows": v[0]} for k, v in results.items()}, f, indent=1)


if __name__ == "__main__":
    main()
