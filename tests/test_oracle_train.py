"""CPU: the oracle's teacher-forced REINFORCE objective, its autograd gradient and its Adam step against the
reference's own training step (tests/golden/train_*.npz, made by oracle/gen_golden_train.py)."""
import numpy as np
import pytest
import torch

from oracle import elg_oracle as O
from train_helpers import TRAIN_CASES, TrainGolden, sample_idx


def oracle_grads(g, dtype=torch.float32):
    W = O.Weights(g.state_dict(), g.kind, g.model_params(), dtype).requires_grad_()
    J, logp = O.reinforce_loss(W, g.problem(dtype), g.M, g.tours(), g.reward().to(dtype), g.meta["scale_norm"])
    J.backward()
    return W, J.detach(), logp.detach(), {k: v.grad for k, v in W.sd.items()}


def check_grads_against_fixture(g, grads, tol=1e-3):
    """Per parameter tensor: sampled entries and the L2 norm agree with the reference's gradient to `tol` of the
    tensor's own rms.  Tensors whose exact gradient is zero (biases in front of an instance norm) hold rounding
    noise only, so the scale is floored at 1e-3 of the largest tensor rms."""
    rms = {k: float(g.z["g_norm/" + k]) / np.sqrt(grads[k].numel()) for k in g.meta["keys"]}
    floor = 1e-3 * max(rms.values())
    worst = 0.0
    for k in g.meta["keys"]:
        mine = grads[k].reshape(-1).double().cpu().numpy()
        ref = g.z["g_sample/" + k].astype(np.float64)
        idx = sample_idx(mine.size)
        scale = max(rms[k], floor)
        err = np.abs(mine[idx] - ref).max() / scale
        nerr = abs(np.linalg.norm(mine) - float(g.z["g_norm/" + k])) / np.sqrt(mine.size) / scale
        worst = max(worst, err, nerr)
        assert err < 20 * tol and nerr < tol, "%s: sample err %.3e, norm err %.3e (x rms %.3e)" % (k, err, nerr, scale)
    return worst


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_oracle_reinforce_gradient_matches_reference(name):
    g = TrainGolden(name)
    W, J, logp, grads = oracle_grads(g)
    ref_lp = torch.tensor(g.z["log_prob"])
    assert float((logp - ref_lp).abs().max()) < 2e-3 * max(1.0, float(ref_lp.abs().max()))
    assert abs(float(J) - float(g.z["J"])) < 2e-3 * max(1.0, abs(float(g.z["J"])))
    check_grads_against_fixture(g, grads)


@pytest.mark.parametrize("name", ["train_cvrp_n20", "train_tsp_n20"])
def test_oracle_adam_step_matches_reference(name):
    g = TrainGolden(name)
    W, _, _, grads = oracle_grads(g)
    rms = {k: float(g.z["g_norm/" + k]) / np.sqrt(grads[k].numel()) for k in g.meta["keys"]}
    for k in g.meta["keys"]:
        if rms[k] < 1e-3 * max(rms.values()):
            continue            # exact gradient is zero: Adam amplifies pure rounding noise to +-lr
        p = W.sd[k].detach()
        new_p, _, _ = O.adam_step(p, grads[k], torch.zeros_like(p), torch.zeros_like(p), 1, g.meta["lr"], weight_decay=g.meta["weight_decay"])
        idx = sample_idx(p.numel())
        ref = g.z["w_after/" + k]
        # the first Adam step moves every weight by ~lr * sign(g); entries whose gradient is ~0 are ill-conditioned
        gs = g.z["g_sample/" + k]
        ok = np.abs(gs) > 1e-3 * max(np.abs(gs).max(), 1e-30)
        diff = np.abs(new_p.reshape(-1).numpy()[idx] - ref)
        assert diff[ok].max(initial=0.0) < 2e-6, k
