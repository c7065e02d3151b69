"""Library-instance drivers (CVRPLIB / TSPLIB): reader on CPU, end-to-end drivers on the GPU."""
import os
import pickle
import random

import numpy as np
import pytest
import torch

from helpers import GOLDEN, Golden
from elg_b200 import vrplib_io

VRP = os.path.join(GOLDEN, "vrplib")


def test_vrplib_reader_fields():
    inst = vrplib_io.read_instance(os.path.join(VRP, "X-n101-k25.vrp"))
    assert inst["dimension"] == 101 and inst["capacity"] == 206
    assert inst["node_coord"].shape == (101, 2) and inst["demand"].shape == (101,)
    assert tuple(inst["node_coord"][0]) == (365.0, 689.0) and inst["demand"][0] == 0 and inst["demand"][1] == 38
    assert list(inst["depot"]) == [0]
    sol = vrplib_io.read_solution(os.path.join(VRP, "X-n101-k25.sol"))
    assert sol["cost"] == 27591 and len(sol["routes"]) == 26
    # the optimal routes are feasible and cost what the file says (rounded EUC_2D)
    c, d = inst["node_coord"], inst["demand"]
    total = 0.0
    seen = set()
    for r in sol["routes"]:
        assert sum(d[j] for j in r) <= inst["capacity"]
        path = [0] + r + [0]
        total += sum(round(float(np.hypot(*(c[a] - c[b])))) for a, b in zip(path, path[1:]))
        seen.update(r)
    assert seen == set(range(1, 101)) and total == sol["cost"]


def _config(problem, name="ELG"):
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    return {"name": name, "use_cuda": True, "cuda_device_num": 0, "vrplib_set": "X", "load_checkpoint": None,
            "params": {"aug_factor": 8}, "model_params": dict(DEFAULT_MODEL_PARAMS[problem])}


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["cvrp_x101", "cvrp_x200"])
def test_vrplib_driver_matches_reference(case, tmp_path):
    from elg_b200.cvrp import CVRPModel
    from elg_b200.cvrp.test_vrplib import VRPLib_Tester
    g = Golden(case)
    cfg = _config("cvrp")
    model = CVRPModel(**cfg["model_params"])
    model.decoder.add_local_policy("cuda:0")
    model.load_state_dict(g.state_dict())
    tester = VRPLib_Tester(cfg, model=model)
    name = g.meta["vrp_file"]
    res = {}
    random.seed(g.meta["seed"])
    sols, rewards = tester.test_on_one_ins(name, res, os.path.join(VRP, name + ".vrp"), os.path.join(VRP, name + ".sol"))
    ref_best = float((-g.reward()).min())
    assert res["scale"] == g.meta["N"]
    assert abs(res["best_cost"] - ref_best) / ref_best < 2e-3
    same = (sols.cpu()[:, :, :min(sols.shape[2], g.T)] == g.tours()[:, :, :min(sols.shape[2], g.T)]).all(dim=2)
    assert same.float().mean() > 0.95
    assert torch.equal(rewards.cpu()[same], g.reward()[same])          # rounded integer costs
    assert res["gap"] == (res["best_cost"] - 27591.0) / 27591.0 if name == "X-n101-k25" else True


@pytest.mark.gpu
def test_tsplib_driver_matches_reference(tmp_path):
    from elg_b200.tsp import TSPModel
    from elg_b200.tsp.test_tsplib import TSPLib_Tester
    g = Golden("tsp_lib")
    coords = g.z["lib_node_coord"]
    d = tmp_path / "TSPLib"
    d.mkdir()
    with open(d / "synthetic52.pkl", "wb") as f:
        pickle.dump([coords, 12345.0], f)
    cfg = _config("tsp")
    cfg["tsplib_path"] = str(d)
    model = TSPModel(**cfg["model_params"])
    model.decoder.add_local_policy("cuda:0")
    model.load_state_dict(g.state_dict())
    tester = TSPLib_Tester(cfg, model=model)
    random.seed(g.meta["seed"])
    results = tester.test_on_tsplib(out_dir=str(tmp_path / "out"))
    rec = results[0]["record"][0]
    ref_best = float((-g.reward()).min())
    assert rec["scale"] == 52 and abs(rec["best_cost"] - ref_best) / ref_best < 2e-3
    assert os.path.exists(tmp_path / "out" / "ELG_tsplib.json")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cvrp", "tsp"])
def test_test_py_loop_matches_oracle_costs(kind, capsys):
    """CVRP/test.py:14-56 / TSP/test.py:14-56 drop-ins: best-of-POMO and best-of-augmentation costs per instance agree
    with the oracle's rollout on the same instances and start permutation."""
    import random
    import torch
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch
    from oracle import elg_oracle as O
    mp = dict(DEFAULT_MODEL_PARAMS[kind])
    sd = synthetic_state_dict(kind, seed=21, gain=3.0)
    dev = "cuda:0"
    if kind == "cvrp":
        from elg_b200.cvrp import CVRPEnv as Env, CVRPModel as Model
        from elg_b200.cvrp.test import solve_batch, test
        data = synthetic_cvrp_batch(4, 20, seed=8)
        batch = {k: v.to(dev) for k, v in data.items()}
        prob = O.load_cvrp(data["depot"], data["loc"], data["demand"], 8)
    else:
        from elg_b200.tsp import TSPEnv as Env, TSPModel as Model
        from elg_b200.tsp.test import solve_batch, test
        data = synthetic_tsp_batch(4, 20, seed=8)
        batch = data.to(dev)
        prob = O.load_tsp(data, 8)
    model = Model(**mp)
    model.decoder.add_local_policy(dev)
    model.load_state_dict(sd)
    model = model.to(dev)
    env = Env(20, dev)
    random.seed(5)
    no_aug, aug, sol, rew = solve_batch(model, env, batch, 8)
    perm = O.start_permutation(kind, 20, 20, seed=5)
    _, _, ref_r = O.rollout(O.Weights(sd, kind, mp), prob, 20, perm, "greedy")
    ref_no_aug, ref_aug = O.best_of(ref_r, 8, 4)
    assert float((aug.cpu() - ref_aug).abs().max()) < 1e-4 * float(ref_aug.max())
    assert float((no_aug.cpu() - ref_no_aug).abs().max()) < 1e-4 * float(ref_no_aug.max())
    random.seed(5)
    avg = test([batch], model, env, 8)
    assert abs(float(avg) - float(ref_aug.mean())) < 1e-4 * float(ref_aug.mean())
    assert "Aug cost" in capsys.readouterr().out


def test_lib_driver_gap_bins_and_records(tmp_path):
    """Host helpers shared by the two library drivers (elg_b200/lib_driver.py): size bins as the reference prints them
    (CVRP/test_vrplib.py:86-109: (0, 200], (200, 500], (500, ...)), record fields of the result file, JSON dump."""
    import json
    from elg_b200 import lib_driver as drv
    res = []
    for name, scale, opt, cost in (("a", 100, 100.0, 105.0), ("b", 200, 50.0, 51.0), ("c", 201, 10.0, 11.0), ("d", 900, 1000.0, 1100.0)):
        rec = {"run_idx": 0}
        drv.fill_record(rec, cost, scale, opt)
        assert rec["best_cost"] == cost and rec["scale"] == scale and abs(rec["gap"] - (cost - opt) / opt) < 1e-12
        res.append({"instance": name, "optimal": opt, "record": [rec]})
    drv.fill_record(None, 1.0, 1, 1.0)                       # the reference passes result_dict=None in ad-hoc calls
    bins = drv.gap_bins(res, [("<200", 0, 200), ("200-500", 200, 500), ("500-1000", 500, 10 ** 9)])
    assert abs(bins["<200"] - 100 * (0.05 + 0.02) / 2) < 1e-9          # scale 200 belongs to the first bin
    assert abs(bins["200-500"] - 10.0) < 1e-9 and abs(bins["500-1000"] - 10.0) < 1e-9
    assert abs(bins["total"] - 100 * (0.05 + 0.02 + 0.1 + 0.1) / 4) < 1e-9
    assert "200-500" not in drv.gap_bins(res[:2], [("<200", 0, 200), ("200-500", 200, 500)])    # empty bins are not reported
    drv.dump_results(res, str(tmp_path / "out"), "x.json")
    assert json.load(open(tmp_path / "out" / "x.json"))[3]["record"][0]["scale"] == 900
    with pytest.raises(RuntimeError):
        drv.require_cuda({"use_cuda": False, "cuda_device_num": 0})
