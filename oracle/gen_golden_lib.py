#!/usr/bin/env python
"""Library-set goldens: run the UNMODIFIED reference drivers (gaocrr/ELG) on CPU over TSPLIB and CVRPLIB Set-X.

TEST INFRASTRUCTURE ONLY (same rules as gen_golden.py): runs in the build container where the reference is
mounted read-only at /root/reference; outputs are committed under tests/golden/lib/; nothing under elg_b200/
imports this file.

    python oracle/gen_golden_lib.py tsp            # all 49 TSPLIB files      (TSP/test_tsplib.py:126-162)
    python oracle/gen_golden_lib.py cvrp           # all 100 Set-X files      (CVRP/test_vrplib.py:111-145)
    python oracle/gen_golden_lib.py cvrp X-n101-k25 X-n1001-k43
    python oracle/gen_golden_lib.py ties           # tie-order floor on X-n101 / X-n200 (see below)
    python oracle/gen_golden_lib.py ties_tsp       # the same on lattice-like TSPLIB instances (pr107, u159, ts225, ...)
    python oracle/gen_golden_lib.py stable_tsp     # every instance of tsplib_ref.json again with index-ordered ties
    python oracle/gen_golden_lib.py stable_cvrp    #   -> {tsplib,setx}_ref_stable.json (summaries only)

What runs is the reference's own `TSPLib_Tester.test_on_one_ins` / `VRPLib_Tester.test_on_one_ins` (imported from
the reference tree, constructed through their own `__init__` with a checkpoint file in the reference format).  The
only things injected are (i) a module called `vrplib` — the reference imports that un-vendored PyPI parser; ours
(`elg_b200/vrplib_io.py`) provides the two calls it makes — and (ii) a wrapper around the module-level name
`rollout` that forwards to the reference's `rollout` and keeps what it returned (tours, rewards), plus, for the
"detail" instances, an `F.softmax` wrapper that keeps the masked logits of chosen steps (as gen_golden.py does).
`random.seed(SEED)` is set before each instance because the reference never seeds the POMO start permutation.

Weights: the released checkpoints are absent, so seeded synthetic ones (`elg_b200/synth.py`, gain 3).

Outputs
  tests/golden/lib/{tsplib,setx}_ref.json   per instance: scale, optimal, best_cost, gap, T, per-augmentation best
                                            cost, sum of all row rewards, seconds on this container's CPU
  tests/golden/lib/{tsplib,setx}_inputs.npz the instance data (coordinates, demands, capacity, optimum) so the GPU
                                            box, which has no reference tree, can run the same sets
  tests/golden/lib/detail_<name>.npz        for a few large instances: every row reward, the POMO permutation,
                                            the tours of a row subset, and pre-decode state + masked logits of
                                            chosen steps / aug-instances / rows
  tests/golden/lib/ties_<name>.npz          `ties` mode: the same instance run twice — unmodified, and with
                                            `torch.topk` replaced (inside the reference's `models` module only) by a
                                            STABLE sort, i.e. equal distances ordered by node index.  Integer
                                            coordinates give exactly equal distances, `torch.topk`'s order among them
                                            is an artefact of its partial-sort, and the local policy is rank-aware
                                            (positional encoding, CVRP/models.py:27-49,74), so the two runs differ;
                                            that difference is the floor for any implementation with another tie rule.
"""
import json
import os
import pickle
import random
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ELG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "lib")
SEED, WSEED, GAIN = 1234, 1234, 3.0
THREADS = int(os.environ.get("ELG_GOLDEN_THREADS", "6"))

# instances that also get a detail file: (steps recorded, aug-instances recorded, rows recorded)
DETAIL = {
    "X-n502-k39": dict(steps=[2, 300, 560], rows_b=[0, 5], rows_m=64, tour_rows=32),
    "X-n1001-k43": dict(steps=[2, 500, 1000], rows_b=[0, 6], rows_m=64, tour_rows=32),
    "pr1002": dict(steps=[1, 500, 1001], rows_b=[0, 3], rows_m=64, tour_rows=32),
    "rat575": dict(steps=[1, 300, 574], rows_b=[0, 7], rows_m=64, tour_rows=32),
}


def _setup(problem):
    sys.path.insert(0, os.path.join(REF, problem.upper()))
    sys.path.insert(1, ROOT)
    import torch
    from elg_b200 import vrplib_io
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_state_dict
    torch.set_num_threads(THREADS)
    shim = types.ModuleType("vrplib")
    shim.read_instance, shim.read_solution = vrplib_io.read_instance, vrplib_io.read_solution
    sys.modules["vrplib"] = shim
    sd = synthetic_state_dict(problem, seed=WSEED, gain=GAIN)
    ckpt = "/tmp/elg_golden_lib_%s.pt" % problem
    torch.save({"model_state_dict": sd, "step": 0}, ckpt)
    config = {"name": "ELG", "use_cuda": False, "cuda_device_num": 0, "vrplib_set": "X", "training": "joint",
              "load_checkpoint": ckpt, "params": {"aug_factor": 8}, "model_params": dict(DEFAULT_MODEL_PARAMS[problem])}
    return config, sd


def _capture(driver_mod, ref_models, tester, detail):
    """Wrap the driver's `rollout` name (and, for detail instances, the model's step + F.softmax)."""
    import torch
    got = {}
    ref_rollout = driver_mod.rollout
    model = tester.model
    orig_step = model.one_step_rollout
    orig_softmax = ref_models.F.softmax
    cur = {"t": 0, "logits": None}
    st_rec = {}

    def cap_softmax(x, dim=None, **kw):
        cur["logits"] = x
        return orig_softmax(x, dim=dim, **kw)

    def cap_step(state, *a, **kw):
        t = cur["t"]
        cur["logits"] = None
        pre = None
        if detail and t in detail["steps"] and state.current_node is not None:
            rb, rm = detail["rows_b"], detail["rows_m"]
            pre = dict(cur=state.current_node[rb, :rm].clone(), mask=state.ninf_mask[rb, :rm].clone())
            if hasattr(state, "load") and state.load is not None:
                pre["load"] = state.load[rb, :rm].clone()
                pre["finished"] = state.finished[rb, :rm].clone()
        out = orig_step(state, *a, **kw)
        if pre is not None and cur["logits"] is not None:
            pre["logits"] = cur["logits"][detail["rows_b"], :detail["rows_m"]].clone()
            pre["selected"] = out[0][detail["rows_b"], :detail["rows_m"]].clone()
            st_rec[t] = pre
        cur["logits"] = None
        cur["t"] = t + 1
        return out

    def cap_rollout(m, env, eval_type="greedy"):
        out = ref_rollout(m, env, eval_type)
        got["tours"], got["reward"] = out[0], out[2]
        return out

    driver_mod.rollout = cap_rollout
    if detail:
        ref_models.F.softmax = cap_softmax
        model.one_step_rollout = cap_step

    def restore():
        driver_mod.rollout = ref_rollout
        ref_models.F.softmax = orig_softmax
        if detail:
            del model.one_step_rollout
    return got, st_rec, restore


def _summ(name, res, got, secs, aug=8):
    rew = got["reward"]
    M = rew.shape[1]
    per_aug = (-rew).reshape(aug, M).min(dim=1)[0]
    return dict(instance=name, scale=int(res["scale"]), best_cost=float(res["best_cost"]), gap=float(res["gap"]),
                T=int(got["tours"].shape[2]), M=int(M), per_aug_best=[float(x) for x in per_aug],
                reward_sum=float(rew.double().sum()), seconds=round(secs, 2))


def _save_detail(name, problem, got, st_rec, detail, sd_sum):
    import numpy as np
    import torch
    rec = {}
    tours = got["tours"]
    rec["reward"] = got["reward"].numpy()
    rec["perm"] = tours[0, :, 1 if problem == "cvrp" else 0].numpy().astype(np.int16)
    rec["tours_rows"] = tours[:, :detail["tour_rows"]].numpy().astype(np.int16)
    rec["rows_b"] = np.array(detail["rows_b"])
    rec["step_ids"] = np.array(sorted(st_rec))
    for t in sorted(st_rec):
        s = st_rec[t]
        rec["s%d_cur" % t] = s["cur"].numpy().astype(np.int16)
        rec["s%d_maskbits" % t] = np.packbits(torch.isinf(s["mask"]).numpy(), axis=-1, bitorder="little")
        rec["s%d_logits" % t] = s["logits"].numpy()
        rec["s%d_selected" % t] = s["selected"].numpy().astype(np.int16)
        if "load" in s:
            rec["s%d_load" % t] = s["load"].numpy()
            rec["s%d_finished" % t] = s["finished"].numpy()
    rec["meta"] = np.array(json.dumps(dict(problem=problem, name=name, seed=SEED, wseed=WSEED, gain=GAIN, wsum=sd_sum,
                                           T=int(tours.shape[2]), M=int(tours.shape[1]), rows_m=detail["rows_m"],
                                           tour_rows=detail["tour_rows"])))
    np.savez_compressed(os.path.join(OUT, "detail_%s.npz" % name), **rec)


def _merge_json(path, rows, header):
    old = {}
    if os.path.exists(path):
        with open(path) as f:
            old = {r["instance"]: r for r in json.load(f)["instances"]}
    for r in rows:
        old[r["instance"]] = r
    with open(path, "w") as f:
        json.dump(dict(header, instances=sorted(old.values(), key=lambda r: (r["scale"], r["instance"]))), f, indent=0)


class _StableTies:
    """Context: inside the reference's `models` module (and for Tensor.topk with largest=False) torch.topk is replaced by
    a STABLE sort, i.e. equal distances keep node-index order."""
    def __init__(self, ref_models):
        self.m = ref_models

    def __enter__(self):
        import torch
        self.orig = torch.Tensor.topk
        orig = self.orig

        def topk(x, k, dim=-1, largest=True, sorted=True):
            if largest:
                return orig(x, k, dim, largest, sorted)
            v, i = torch.sort(x, dim=dim, descending=False, stable=True)
            return v.narrow(dim, 0, k), i.narrow(dim, 0, k)
        torch.Tensor.topk = topk
        return self

    def __exit__(self, *a):
        import torch
        torch.Tensor.topk = self.orig


def worker_cvrp(names, stable=False):
    import numpy as np
    import torch
    config, sd = _setup("cvrp")
    from elg_b200.synth import state_dict_checksum
    from elg_b200 import vrplib_io
    import models as ref_models
    import test_vrplib as drv
    tester = drv.VRPLib_Tester(config)
    base = os.path.join(REF, "CVRP", "VRPLib", "Vrp-Set-X")
    allnames = sorted((f[:-4] for f in os.listdir(base) if f.endswith(".vrp")), key=lambda s: int(s.split("-")[1][1:]))
    # inputs for the GPU box
    inputs = {}
    for nm in allnames:
        inst = vrplib_io.read_instance(os.path.join(base, nm + ".vrp"))
        inputs[nm + "/coord"] = inst["node_coord"].astype(np.int32)
        inputs[nm + "/demand"] = inst["demand"].astype(np.int32)
        inputs[nm + "/capopt"] = np.array([inst["capacity"], vrplib_io.read_solution(os.path.join(base, nm + ".sol"))["cost"]])
        assert (inputs[nm + "/coord"] == inst["node_coord"]).all() and (inputs[nm + "/demand"] == inst["demand"]).all()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "setx_inputs.npz"), **inputs)
    header = dict(source="CVRP/test_vrplib.py:111-145 (unmodified, CPU)", seed=SEED, wseed=WSEED, gain=GAIN,
                  wsum=state_dict_checksum(sd), torch=torch.__version__)
    if stable:
        return _run_stable("cvrp", names, drv, ref_models, tester, header, "setx",
                           lambda nm: dict(instance=os.path.join(base, nm + ".vrp"), solution=os.path.join(base, nm + ".sol")),
                           lambda nm: float(inputs[nm + "/capopt"][1]))
    for nm in (names or allnames):
        detail = DETAIL.get(nm)
        got, st_rec, restore = _capture(drv, ref_models, tester, detail)
        res = {}
        random.seed(SEED)
        t0 = time.time()
        tester.test_on_one_ins(name=nm, result_dict=res, instance=os.path.join(base, nm + ".vrp"),
                               solution=os.path.join(base, nm + ".sol"))
        secs = time.time() - t0
        restore()
        row = _summ(nm, res, got, secs)
        row["optimal"] = float(inputs[nm + "/capopt"][1])
        if detail:
            _save_detail(nm, "cvrp", got, st_rec, detail, header["wsum"])
        _merge_json(os.path.join(OUT, "setx_ref.json"), [row], header)
        print("%-14s N=%4d T=%4d best=%9.0f gap=%.4f  %.1fs" % (nm, row["scale"], row["T"], row["best_cost"], row["gap"], secs), flush=True)


def _run_stable(problem, names, drv, ref_models, tester, header, kind, make_args, optimum):
    """The instances already in <kind>_ref.json once more with index-ordered ties -> <kind>_ref_stable.json."""
    import torch
    with open(os.path.join(OUT, kind + "_ref.json")) as f:
        todo = [r["instance"] for r in json.load(f)["instances"]]
    header = dict(header, source=header["source"].replace("unmodified", "torch.topk -> stable sort: index-ordered distance ties"))
    for nm in (names or todo):
        got, _, restore = _capture(drv, ref_models, tester, None)
        res = {}
        random.seed(SEED)
        t0 = time.time()
        with torch.no_grad(), _StableTies(ref_models):
            tester.test_on_one_ins(name=nm, result_dict=res, **make_args(nm))
        secs = time.time() - t0
        restore()
        row = _summ(nm, res, got, secs)
        row["optimal"] = optimum(nm)
        _merge_json(os.path.join(OUT, kind + "_ref_stable.json"), [row], header)
        print("%-14s N=%4d best=%10.0f gap=%.4f  %.1fs (index-ordered ties)" % (nm, row["scale"], row["best_cost"], row["gap"], secs), flush=True)


def worker_tsp(names, stable=False):
    import numpy as np
    import torch
    config, sd = _setup("tsp")
    from elg_b200.synth import state_dict_checksum
    import models as ref_models
    import test_tsplib as drv
    tester = drv.TSPLib_Tester(config)
    base = os.path.join(REF, "TSP", "TSPLib")
    data = {}
    for f in os.listdir(base):
        if f.endswith(".pkl"):
            with open(os.path.join(base, f), "rb") as fh:
                data[f[:-4]] = pickle.load(fh)
    allnames = sorted(data, key=lambda k: (data[k][0].shape[0], k))
    inputs = {}
    for nm in allnames:
        inputs[nm + "/coord"] = np.asarray(data[nm][0], dtype=np.float64)
        inputs[nm + "/opt"] = np.array(float(data[nm][1]))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "tsplib_inputs.npz"), **inputs)
    header = dict(source="TSP/test_tsplib.py:126-162 (unmodified, CPU)", seed=SEED, wseed=WSEED, gain=GAIN,
                  wsum=state_dict_checksum(sd), torch=torch.__version__)
    if stable:
        return _run_stable("tsp", names, drv, ref_models, tester, header, "tsplib", lambda nm: dict(instance=data[nm]),
                           lambda nm: float(data[nm][1]))
    for nm in (names or allnames):
        detail = DETAIL.get(nm)
        got, st_rec, restore = _capture(drv, ref_models, tester, detail)
        res = {}
        random.seed(SEED)
        t0 = time.time()
        with torch.no_grad():
            tester.test_on_one_ins(name=nm, result_dict=res, instance=data[nm])
        secs = time.time() - t0
        restore()
        row = _summ(nm, res, got, secs)
        row["optimal"] = float(data[nm][1])
        if detail:
            _save_detail(nm, "tsp", got, st_rec, detail, header["wsum"])
        _merge_json(os.path.join(OUT, "tsplib_ref.json"), [row], header)
        print("%-10s N=%4d best=%10.0f gap=%.4f  %.1fs" % (nm, row["scale"], row["best_cost"], row["gap"], secs), flush=True)


def worker_ties(names, problem="cvrp"):
    """Unmodified reference vs the reference with index-ordered ties (stable sort instead of torch.topk)."""
    import numpy as np
    import torch
    config, sd = _setup(problem)
    from elg_b200.synth import state_dict_checksum
    import models as ref_models
    if problem == "cvrp":
        import test_vrplib as drv
        tester = drv.VRPLib_Tester(config)
        base = os.path.join(REF, "CVRP", "VRPLib", "Vrp-Set-X")
        default = ["X-n101-k25", "X-n200-k36"]
    else:
        import test_tsplib as drv
        tester = drv.TSPLib_Tester(config)
        base = os.path.join(REF, "TSP", "TSPLib")
        default = ["eil76", "pr107", "u159", "ts225", "a280", "pcb442"]

    class _StableTopk:
        """`torch` stand-in for the reference's models module: topk(largest=False) by stable sort."""
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def topk(x, k, dim=-1, largest=True, sorted=True):
            assert not largest
            v, i = torch.sort(x, dim=dim, descending=False, stable=True)
            return v.narrow(dim, 0, k), i.narrow(dim, 0, k)

    for nm in (names or default):
        runs = {}
        for mode in ("unmodified", "stable"):
            if mode == "stable":
                ref_models.torch = _StableTopk()
                orig_tensor_topk = torch.Tensor.topk
                torch.Tensor.topk = lambda self, k, dim=-1, largest=True, sorted=True: (
                    _StableTopk.topk(self, k, dim, largest, sorted) if not largest else orig_tensor_topk(self, k, dim, largest, sorted))
            got, _, restore = _capture(drv, ref_models, tester, None)
            res = {}
            random.seed(SEED)
            with torch.no_grad():
                if problem == "cvrp":
                    tester.test_on_one_ins(name=nm, result_dict=res, instance=os.path.join(base, nm + ".vrp"),
                                           solution=os.path.join(base, nm + ".sol"))
                else:
                    with open(os.path.join(base, nm + ".pkl"), "rb") as fh:
                        inst = pickle.load(fh)
                    tester.test_on_one_ins(name=nm, result_dict=res, instance=inst)
            restore()
            if mode == "stable":
                ref_models.torch = torch
                torch.Tensor.topk = orig_tensor_topk
            runs[mode] = (got["tours"], got["reward"], res["best_cost"])
        ta, tb = runs["unmodified"][0], runs["stable"][0]
        T = max(ta.shape[2], tb.shape[2])
        pa = torch.zeros(ta.shape[0], ta.shape[1], T, dtype=torch.long); pa[:, :, :ta.shape[2]] = ta
        pb = torch.zeros_like(pa); pb[:, :, :tb.shape[2]] = tb
        same = (pa == pb).all(dim=2)
        print("%s: unmodified vs index-ordered ties: %.2f %% of rows identical; best %.0f vs %.0f" % (
            nm, 100 * same.float().mean(), runs["unmodified"][2], runs["stable"][2]), flush=True)
        np.savez_compressed(os.path.join(OUT, "ties_%s.npz" % nm),
                            tours_unmodified=ta.numpy().astype(np.int16), reward_unmodified=runs["unmodified"][1].numpy(),
                            tours_stable=tb.numpy().astype(np.int16), reward_stable=runs["stable"][1].numpy(),
                            perm=ta[0, :, 1 if problem == "cvrp" else 0].numpy().astype(np.int16),
                            meta=np.array(json.dumps(dict(name=nm, seed=SEED, wseed=WSEED, gain=GAIN, problem=problem,
                                                          best_unmodified=float(runs["unmodified"][2]), best_stable=float(runs["stable"][2]),
                                                          wsum=state_dict_checksum(sd), rows_identical=float(same.float().mean())))))


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        if sys.argv[2] == "ties_tsp":
            worker_ties(sys.argv[3:], "tsp")
        elif sys.argv[2] == "stable_tsp":
            worker_tsp(sys.argv[3:], stable=True)
        elif sys.argv[2] == "stable_cvrp":
            worker_cvrp(sys.argv[3:], stable=True)
        else:
            {"cvrp": worker_cvrp, "tsp": worker_tsp, "ties": worker_ties}[sys.argv[2]](sys.argv[3:])
        return
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    for m in (["tsp", "cvrp", "ties", "ties_tsp"] if mode == "all" else [mode]):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", m] + sys.argv[2:], check=True)


if __name__ == "__main__":
    main()
