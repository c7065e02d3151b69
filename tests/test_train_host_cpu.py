"""CPU: host-side pieces of the training drivers that need no GPU (reference loop semantics)."""
import json

import numpy as np
import pytest
import torch


def test_logger_writes_reference_json_layout(tmp_path):
    from elg_b200.train_loop import Logger
    lg = Logger(str(tmp_path / "log" / "run"), {"name": "ELG", "params": {"problem_size": 100}})
    lg.log([15.9, 8.1, 14.5])
    lg.log([15.8, 8.0, 14.4])
    d = json.load(open(str(tmp_path / "log" / "run")))
    assert d["name"] == "ELG" and d["result"] == {"val_100": [15.9, 15.8], "val_200": [8.1, 8.0], "val_500": [14.5, 14.4]}


def test_mixed_distribution_sampling_follows_validation_gaps():
    """CVRP/train.py:96-99,143-148: data_type ~ softmax(gaps), gaps = (val - opt) / opt."""
    from elg_b200.train_loop import softmax
    opts = np.array([15.740834, 7.909336, 14.294179])
    gaps = (np.array([16.5, 8.0, 15.5]) - opts) / opts
    p = softmax(gaps)
    assert abs(p.sum() - 1) < 1e-12 and p[2] > p[0] > p[1]
    assert np.allclose(softmax(np.array([1, 1, 1])), 1 / 3)


def test_generator_front_end_checks():
    from elg_b200 import generate_data as G
    from elg_b200._lib import ElgError
    dist = {"data_type": np.array(["cluster"]), "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07}
    assert G._kind(dist) == (1, 3) and G._kind(dict(dist, data_type="mixed")) == (2, 1) and G._kind(dict(dist, data_type="uniform")) == (0, 1)
    with pytest.raises(KeyError):
        G._kind(dict(dist, data_type="gaussian"))
    with pytest.raises(KeyError):
        G.generate_vrp_data(4, 37, dict(dist, data_type="uniform"), device="cuda:0")       # CAPACITIES has no 37 (before any CUDA call)
    with pytest.raises(ElgError):
        G.generate_tsp_data(4, 20, dict(dist, data_type="uniform"), device="cpu")
    assert G.CAPACITIES == {10: 20., 20: 30., 50: 40., 100: 50., 200: 80., 500: 100., 1000: 250.}


def test_trainer_needs_cuda():
    from elg_b200._lib import ElgError
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_state_dict
    from elg_b200.trainer import Trainer
    with pytest.raises(ElgError):
        Trainer("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]), synthetic_state_dict("cvrp", seed=1), "cpu")


def test_training_modes_other_than_joint_are_rejected():
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    from elg_b200.train_loop import train
    cfg = {"training": "only_local", "params": {"problem_size": 20, "multiple_width": 20, "scale_norm": True, "T": 0, "start_steps": 0,
                                                "train_steps": 1, "mixed": False, "train_batch_size": 2, "learning_rate": 1e-4, "log_step": 10},
           "distribution": {"data_type": "uniform"}, "model_params": dict(DEFAULT_MODEL_PARAMS["cvrp"])}
    with pytest.raises(NotImplementedError):
        train("cvrp", cfg, "cuda:0", verbose=False)


def test_dataset_classes_read_reference_pickles(tmp_path):
    """VRPDataset / TSPDataset (CVRP/generate_data.py:108-171, TSP/generate_data.py:74-99) on files in the reference's
    pickle layout: [(depot, loc, demand, capacity), ...] and [[[x, y], ...], ...]; DataLoader collation as in test.py."""
    from torch.utils.data import DataLoader
    from elg_b200.generate_data import TSPDataset, VRPDataset, save_dataset
    rng = np.random.RandomState(0)
    vrp = [(rng.rand(2).tolist(), rng.rand(10, 2).tolist(), rng.randint(1, 10, 10).tolist(), 20.0) for _ in range(5)]
    save_dataset(vrp, str(tmp_path / "d" / "vrp10"))
    ds = VRPDataset(str(tmp_path / "d" / "vrp10.pkl"), num_samples=4, offset=1)
    assert len(ds) == 4 and ds[0]['loc'].shape == (10, 2) and ds[0]['depot'].shape == (2,)
    assert torch.allclose(ds[0]['demand'], torch.tensor(vrp[1][2], dtype=torch.float) / 20.0)
    batch = next(iter(DataLoader(ds, batch_size=4)))
    assert batch['loc'].shape == (4, 10, 2) and batch['demand'].shape == (4, 10) and batch['depot'].shape == (4, 2)
    tsp = rng.rand(6, 12, 2).tolist()
    save_dataset(tsp, str(tmp_path / "d" / "tsp12.pkl"))
    dt = TSPDataset(str(tmp_path / "d" / "tsp12.pkl"), num_samples=6)
    assert len(dt) == 6 and next(iter(DataLoader(dt, batch_size=3))).shape == (3, 12, 2)
