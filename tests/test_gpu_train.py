"""GPU: the CUDA training path (elg_encode_train + elg_rollout(sample) + elg_reinforce_backward + elg_adam_step)
against the reference's own training step (tests/golden/train_*.npz) and the oracle's autograd."""
import random

import numpy as np
import pytest
import torch

import train_helpers as TH

pytestmark = pytest.mark.gpu


def _setup(g, dev="cuda:0"):
    from elg_b200 import engine
    from elg_b200.trainer import Trainer
    tr = Trainer(g.kind, g.meta["model_params"], g.state_dict(), dev, scale_norm=g.meta["scale_norm"])
    data = g.data()
    if g.kind == "cvrp":
        xy, dem = engine.load_problems("cvrp", data["loc"].to(dev), data["depot"].to(dev), data["demand"].to(dev), 1)
    else:
        xy, dem = engine.load_problems("tsp", data.to(dev), None, None, 1)
    return engine, tr, xy, dem


@pytest.mark.parametrize("name", TH.TRAIN_CASES)
def test_gradient_matches_reference_training_step(name):
    """Teacher-forced on the reference's sampled tours: every parameter gradient within 2e-2 of the tensor's rms of the
    reference's J.backward() (sampled entries) and of the oracle's autograd (all entries)."""
    g = TH.TrainGolden(name)
    engine, tr, xy, dem = _setup(g)
    N1 = int(xy.shape[1])
    t_max = 2 * N1 + 2 if g.kind == "cvrp" else N1
    batch, saved = engine.encode_train(tr.handle, xy, dem)
    tours = g.tours()
    grads, loss, _ = engine.reinforce_backward(batch, saved, g.M, TH.pad_tours(tours, t_max).cuda(), tours.shape[2],
                                               g.reward().cuda(), torch.tensor(g.z["log_prob"]).cuda(), g.meta["scale_norm"], 8)
    torch.cuda.synchronize()
    mine = tr.unpack(grads)
    ef = TH.fixture_errors(g, mine)
    assert max(ef.values()) < 2e-2, sorted(ef.items(), key=lambda kv: -kv[1])[:3]
    _, _, ref, _ = TH.oracle_grads(g.kind, g.model_params(), g.state_dict(), g.problem(), g.M, tours, g.reward(),
                                   g.meta["scale_norm"])
    eo = TH.grad_errors(mine, ref)
    assert max(eo.values()) < 2e-2, sorted(eo.items(), key=lambda kv: -kv[1])[:3]
    assert abs(float(loss) - float(g.z["J"])) < 2e-3 * max(1.0, abs(float(g.z["J"])))


@pytest.mark.parametrize("chunk", [1, 5, 32])
def test_gradient_independent_of_chunking(chunk):
    g = TH.TrainGolden("train_cvrp_n20")
    engine, tr, xy, dem = _setup(g)
    batch, saved = engine.encode_train(tr.handle, xy, dem)
    tours = g.tours()
    t16 = TH.pad_tours(tours, 2 * 21 + 2).cuda()
    a, _, _ = engine.reinforce_backward(batch, saved, g.M, t16, tours.shape[2], g.reward().cuda(), None, True, 16)
    a = a.clone()
    b, _, _ = engine.reinforce_backward(batch, saved, g.M, t16, tours.shape[2], g.reward().cuda(), None, True, chunk)
    torch.cuda.synchronize()
    assert float((a - b).abs().max()) < 1e-4 * float(a.abs().max())


@pytest.mark.parametrize("kind,N", [("cvrp", 20), ("tsp", 20), ("cvrp", 100)])
def test_own_sampled_rollout_gradient_and_adam(kind, N):
    """Full step on our own sampled rollout: log-probs, loss and gradient agree with the oracle replaying the same
    tours; the Adam update agrees with the oracle's (torch.optim.Adam formula)."""
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch
    from elg_b200.trainer import Trainer
    from oracle import elg_oracle as O
    n, M = (3, N) if N <= 20 else (2, 100)
    mp = dict(DEFAULT_MODEL_PARAMS[kind])
    sd = synthetic_state_dict(kind, seed=5, gain=2.0)
    tr = Trainer(kind, mp, sd, "cuda:0")
    if kind == "cvrp":
        data = synthetic_cvrp_batch(n, N, seed=9)
        prob = O.load_cvrp(data["depot"], data["loc"], data["demand"], 1)
    else:
        data = synthetic_tsp_batch(n, N, seed=9)
        prob = O.load_tsp(data, 1)
    random.seed(4)
    w_before = tr.handle.weights.clone()
    out = tr.step(data, M, seed=123)
    torch.cuda.synchronize()
    T = out["T"]
    tours = out["tours"][:, :, :T].long().cpu()
    if kind == "cvrp":
        O.check_feasible_cvrp(tours, prob.demand)
    reward = out["reward"].cpu()
    J, logp, ref, _ = TH.oracle_grads(kind, mp, sd, prob, M, tours, reward, True)
    assert float((out["logp"].cpu() - logp).abs().max()) < 5e-3 * max(1.0, float(logp.abs().max()))
    assert abs(float(out["loss"]) - float(J)) < 5e-3 * max(1.0, abs(float(J)))
    # Frobenius error per tensor: on a freshly sampled rollout a ReLU pre-activation within rounding of zero may flip
    # between our forward and the oracle's, which moves one row of that layer's gradient (seen on layers.3 W1)
    eo = TH.grad_errors_l2(tr.unpack(out["grads"]), ref)
    assert max(eo.values()) < 2e-2, sorted(eo.items(), key=lambda kv: -kv[1])[:3]
    # Adam: compare on the packed buffer with the oracle formula applied to OUR gradient (isolates the optimizer)
    g = out["grads"].cpu()
    p, _, _ = O.adam_step(w_before.cpu(), g, torch.zeros_like(g), torch.zeros_like(g), 1)
    assert float((tr.handle.weights.cpu() - p).abs().max()) < 1e-6
    new_sd = tr.state_dict()
    assert set(new_sd) == set(k for k in sd if k in new_sd) and all(new_sd[k].shape == sd[k].shape for k in new_sd)


def test_training_loop_driver_learns_and_checkpoints(tmp_path):
    """The train() driver (reference loop semantics) on CVRP20: the sampled tour length drops over 60 Adam steps,
    a checkpoint in the reference's format is written and loads into the drop-in model, the JSON log has the keys."""
    import json
    import numpy as np
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    from elg_b200.train_loop import train
    torch.manual_seed(3); np.random.seed(3); random.seed(3)
    config = {"name": "t", "training": "joint", "seed": 3,
              "params": {"problem_size": 20, "multiple_width": 20, "scale_norm": True, "T": 0, "start_steps": 0,
                         "train_steps": 59, "mixed": True, "train_batch_size": 64, "learning_rate": 1e-4, "log_step": 60},
              "distribution": {"data_type": "uniform", "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07},
              "model_params": dict(DEFAULT_MODEL_PARAMS["cvrp"])}
    tr, hist = train("cvrp", config, "cuda:0", dir_path=str(tmp_path / "w"), log_path=str(tmp_path / "log.json"),
                     val_samples=(64, 64, 32), verbose=False)
    first = np.mean([h[1] for h in hist[:5]])
    last = np.mean([h[1] for h in hist[-5:]])
    assert last < first - 0.1, (first, last)
    ck = torch.load(str(tmp_path / "w" / "model_epoch_1.pt"))
    assert set(ck) == {"step", "model_state_dict", "optimizer_state_dict"}
    from elg_b200.cvrp import CVRPModel
    m = CVRPModel(**config["model_params"])
    m.decoder.add_local_policy("cpu")
    m.load_state_dict(ck["model_state_dict"])
    log = json.load(open(str(tmp_path / "log.json")))
    assert len(log["result"]["val_100"]) == 1 and len(log["result"]["val_500"]) == 1


def _own_rollout_check(kind, N, n, M, scale_norm=True, gain=1.5, tol=2e-2):
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch
    from elg_b200.trainer import Trainer
    from oracle import elg_oracle as O
    mp = dict(DEFAULT_MODEL_PARAMS[kind])
    sd = synthetic_state_dict(kind, seed=11, gain=gain)
    tr = Trainer(kind, mp, sd, "cuda:0", scale_norm=scale_norm)
    if kind == "cvrp":
        data = synthetic_cvrp_batch(n, N, seed=17)
        prob = O.load_cvrp(data["depot"], data["loc"], data["demand"], 1)
    else:
        data = synthetic_tsp_batch(n, N, seed=17)
        prob = O.load_tsp(data, 1)
    random.seed(6)
    out = tr.forward_backward(data, M, seed=77)
    torch.cuda.synchronize()
    tours = out["tours"][:, :, :out["T"]].long().cpu()
    J, logp, ref, _ = TH.oracle_grads(kind, mp, sd, prob, M, tours, out["reward"].cpu(), scale_norm)
    assert float((out["logp"].cpu() - logp).abs().max()) < 5e-3 * max(1.0, float(logp.abs().max()))
    assert abs(float(out["loss"]) - float(J)) < 5e-3 * max(1.0, abs(float(J)))
    eo = TH.grad_errors_l2(tr.unpack(out["grads"]), ref)
    assert max(eo.values()) < tol, sorted(eo.items(), key=lambda kv: -kv[1])[:3]


@pytest.mark.parametrize("kind,N", [("cvrp", 107), ("tsp", 112)])
def test_largest_supported_instance_uses_the_8_warp_variant(kind, N):
    """The resident limit (elg_rollout_resident): 108 nodes for cvrp (k = 40), 112 for tsp; the tables leave room for 8 warps."""
    _own_rollout_check(kind, N, 2, 40)


def test_tsp100_full_width():
    _own_rollout_check("tsp", 100, 2, 100)


def test_fewer_rows_than_warps_and_single_instance():
    _own_rollout_check("cvrp", 20, 1, 7)


def test_without_advantage_scaling():
    _own_rollout_check("cvrp", 20, 3, 20, scale_norm=False)


def test_tsp_scaling_skipped_when_an_instance_has_zero_max_advantage():
    """TSP/train.py:114-117: the max-advantage scaling applies only if every instance's maximum advantage is non-zero.
    M = 1 makes every advantage zero -> unscaled J = 0 and an exactly zero gradient (no NaN)."""
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_state_dict, synthetic_tsp_batch
    from elg_b200.trainer import Trainer
    tr = Trainer("tsp", dict(DEFAULT_MODEL_PARAMS["tsp"]), synthetic_state_dict("tsp", seed=11, gain=1.0), "cuda:0")
    out = tr.forward_backward(synthetic_tsp_batch(3, 20, seed=1), 1, start_nodes=[0], seed=5)
    torch.cuda.synchronize()
    assert float(out["loss"]) == 0.0 and float(out["grads"].abs().max()) == 0.0


def test_training_rejects_instances_beyond_the_resident_limit():
    from elg_b200._lib import ElgError
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict
    from elg_b200.trainer import Trainer
    tr = Trainer("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]), synthetic_state_dict("cvrp", seed=11, gain=1.0), "cuda:0")
    with pytest.raises(ElgError):
        tr.forward_backward(synthetic_cvrp_batch(1, 150, seed=1), 20, seed=5)


def test_checkpoint_resume_reproduces_the_next_step():
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict
    from elg_b200.trainer import Trainer
    mp = dict(DEFAULT_MODEL_PARAMS["cvrp"])
    sd = synthetic_state_dict("cvrp", seed=3, gain=1.0)
    perm = list(range(20))
    a = Trainer("cvrp", mp, sd, "cuda:0")
    for s in range(2):
        a.step(synthetic_cvrp_batch(4, 20, seed=30 + s), 20, start_nodes=perm, seed=100 + s)
    ck = a.checkpoint()
    b = Trainer("cvrp", mp, sd, "cuda:0")
    b.load_checkpoint(ck)
    assert b.step_count == 2 and torch.equal(a.handle.weights, b.handle.weights) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    oa = a.step(synthetic_cvrp_batch(4, 20, seed=32), 20, start_nodes=perm, seed=102)
    ob = b.step(synthetic_cvrp_batch(4, 20, seed=32), 20, start_nodes=perm, seed=102)
    torch.cuda.synchronize()
    assert torch.equal(oa["tours"], ob["tours"])
    # gradients are accumulated with floating-point atomics: equal up to summation order; Adam turns the rounding noise of
    # exactly-zero gradients (biases in front of an instance norm) into +-lr steps, so weights agree to a few lr there
    assert float((oa["grads"] - ob["grads"]).abs().max()) < 1e-4 * float(oa["grads"].abs().max())
    assert float((a.handle.weights - b.handle.weights).abs().max()) < 2.5e-4
    sd2 = b.state_dict()
    assert all(sd2[k].shape == sd[k].shape for k in sd2)


def test_joint_training_switches_on_the_local_policy_at_step_T(tmp_path):
    """CVRP/train.py:91-95: global policy + distance penalty alone until step T, then add_local_policy + new optimizer."""
    import numpy as np
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    from elg_b200.train_loop import train
    from elg_b200.trainer import Trainer
    torch.manual_seed(5); np.random.seed(5); random.seed(5)
    config = {"name": "t", "training": "joint", "seed": 5,
              "params": {"problem_size": 20, "multiple_width": 20, "scale_norm": True, "T": 3, "start_steps": 0,
                         "train_steps": 5, "mixed": False, "train_batch_size": 8, "learning_rate": 1e-4, "log_step": 100},
              "distribution": {"data_type": "uniform", "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07},
              "model_params": dict(DEFAULT_MODEL_PARAMS["cvrp"])}
    tr, hist = train("cvrp", dict(config, params=dict(config["params"], train_steps=2)), "cuda:0", verbose=False)
    assert not tr.has_local and tr.step_count == 3 and not any("local" in k for k in tr.state_dict())
    lo = min(tr._all_slots[k] for k in tr._local_keys)
    assert float(tr.handle.weights[lo:].abs().max()) == 0.0            # local slots (the tail of the buffer) untouched
    tr, hist = train("cvrp", config, "cuda:0", verbose=False)
    assert tr.has_local and tr.step_count == 3                           # 6 steps: 3 warm-up, optimizer restarted, 3 joint
    sd = tr.state_dict()
    assert any("local_policies.0.Wq.weight" in k for k in sd)
    assert float(tr.exp_avg[lo:].abs().max()) > 0.0                      # the local policy receives gradient now
    assert len(hist) == 6 and all(np.isfinite(h[0]) for h in hist)


def test_train_py_entry_point_reads_config_yml(tmp_path):
    """`python -m elg_b200.cvrp.train` in a directory with the reference's config.yml layout (CVRP/config.yml)."""
    import os
    import subprocess
    import sys
    import yaml
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    cfg = {"name": "ELG", "use_cuda": True, "cuda_device_num": 0, "logger": "no_logger", "load_checkpoint": None,
           "training": "joint", "seed": 924,
           "params": {"problem_size": 20, "multiple_width": 20, "scale_norm": True, "T": 2, "start_steps": 0, "train_steps": 3,
                      "mixed": False, "train_batch_size": 8, "test_size": 10, "test_batch_size": 10, "learning_rate": 1e-4,
                      "log_step": 4, "aug_factor": 8},
           "distribution": {"data_type": "uniform", "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07},
           "model_params": dict(DEFAULT_MODEL_PARAMS["cvrp"])}
    with open(tmp_path / "config.yml", "w") as f:
        yaml.safe_dump(cfg, f)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "elg_b200.cvrp.train"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Enable joint training." in r.stdout
    runs = os.listdir(tmp_path / "weights")
    assert len(runs) == 1 and os.path.exists(tmp_path / "weights" / runs[0] / "model_epoch_1.pt")
    assert len(os.listdir(tmp_path / "log")) == 1


def test_full_size_gradient_is_the_mean_of_shard_gradients():
    """Size-independent property at BASELINE's full training size (64 instances x 100 rollouts, CVRP100): J is a mean
    over instances of independent per-instance terms, so the gradient of the full batch equals the mean of the gradients
    of its two halves replayed on the same tours (this is also what data-parallel training relies on)."""
    from elg_b200 import engine
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict
    from elg_b200.trainer import Trainer
    tr = Trainer("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]), synthetic_state_dict("cvrp", seed=1234, gain=1.0), "cuda:0")
    data = synthetic_cvrp_batch(64, 100, seed=77)
    random.seed(1)
    out = tr.forward_backward(data, 100, seed=9)
    full = out["grads"].clone()
    T, tours, reward, logp = out["T"], out["tours"], out["reward"], out["logp"]
    assert 110 <= T <= 204 and float(full.abs().max()) > 0
    xy, dem = out["batch"].xy, out["batch"].demand
    parts = []
    for s in (slice(0, 32), slice(32, 64)):
        batch, saved = engine.encode_train(tr.handle, xy[s].contiguous(), dem[s].contiguous())
        g, loss, _ = engine.reinforce_backward(batch, saved, 100, tours[s].contiguous(), T, reward[s].contiguous(),
                                               logp[s].contiguous(), True, 128)
        parts.append(g.clone())
    torch.cuda.synchronize()
    mean = 0.5 * (parts[0] + parts[1])
    err = float((mean - full).abs().max()) / float(full.abs().max())
    assert err < 2e-4, err
    # and the loss reported for the full batch is the mean of the per-row terms: J = sum coef * logp
    adv = reward - reward.mean(dim=1, keepdim=True)
    J = float((-(adv / adv.max(dim=1, keepdim=True)[0]) * logp).mean())
    assert abs(float(out["loss"]) - J) < 1e-4 * max(1.0, abs(J))


@pytest.mark.parametrize("kind", ["cvrp", "tsp"])
def test_encode_train_equals_encode_bit_for_bit(kind):
    """elg_encode_train runs the inference encoder's kernels with per-layer buffers: every decoder table is identical."""
    from elg_b200 import engine
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch
    h = engine.ModelHandle(kind, dict(DEFAULT_MODEL_PARAMS[kind]), synthetic_state_dict(kind, seed=2, gain=2.0), "cuda:0", attention="fp32")
    if kind == "cvrp":
        d = synthetic_cvrp_batch(5, 100, seed=3)
        xy, dem = engine.load_problems("cvrp", d["loc"].cuda(), d["depot"].cuda(), d["demand"].cuda(), 1)
    else:
        xy, dem = engine.load_problems("tsp", synthetic_tsp_batch(5, 100, seed=3).cuda(), None, None, 1)
    a = engine.encode(h, xy, dem)
    b, saved = engine.encode_train(h, xy, dem)
    torch.cuda.synchronize()
    for name in ("enc", "k", "v", "qtab", "eb", "e"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    N1 = int(xy.shape[1])
    pair = 8 * ((N1 + 1) & ~1)
    per = 128 + pair + 16 * N1                          # ELG_NBR_NODE_BYTES: list + (distance, angle) pairs + rank-ordered records
    va, vb = a.nbr.view(-1, per), b.nbr.view(-1, per)
    assert torch.equal(va[:, :128 + 8 * N1], vb[:, :128 + 8 * N1])      # the pad entry of the pair table is never written
    assert torch.equal(va[:, 128 + pair:], vb[:, 128 + pair:])
    rows = xy.shape[0] * xy.shape[1]
    sv = saved.view(torch.float32)
    last_t2_off = (5 * (8 * 128 + 512) + 7 * 128 + 512) * rows          # layer 5, pre-norm sum of the second sub-layer
    t2 = sv[last_t2_off:last_t2_off + rows * 128].reshape(xy.shape[0], xy.shape[1], 128)
    # the encoded nodes are the instance norm of that sum: zero mean / unit variance per channel before the affine map
    assert torch.isfinite(t2).all() and float(t2.abs().max()) > 0
