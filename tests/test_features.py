"""Per-step feature tensors (SURVEY 8a rows `CVRPEnv.get_cur_feature` / `TSPEnv.get_local_feature`).

CPU: the oracle's `cur_feature` against tensors recorded from the unmodified reference environments
(tests/golden/features_*.npz, oracle/gen_golden_features.py), including rows whose load is exactly 0
(norm_demand = nan at the depot, +inf at the customers).
GPU: `elg_cur_feature` through the drop-in environments against the same fixtures, and two mixed-mode runs:
  * a reference-style Step_State (fp32 0/-inf ninf_mask, no packed bits) built by plain torch code drives OUR model;
  * the features OUR environment hands out drive a torch restatement of the reference model's feature consumer
    (the oracle's local policy fed from the environment tensors, CVRP/models.py:51-175).
"""
import os
import random
from dataclasses import dataclass

import numpy as np
import pytest
import torch

from helpers import GOLDEN, compare_tours
from oracle import elg_oracle as O


def _load(kind):
    z = np.load(os.path.join(GOLDEN, "features_%s.npz" % kind))
    if kind == "cvrp":
        prob = O.load_cvrp(torch.tensor(z["depot"]), torch.tensor(z["loc"]), torch.tensor(z["demand"]), 8)
    else:
        prob = O.load_tsp(torch.tensor(z["problems"]), 8)
    return z, prob


def _same_special(a, b):
    """nan at the same places, infinities equal with sign, finite entries compared by the caller"""
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    inf = torch.isinf(a)
    assert torch.equal(inf, torch.isinf(b))
    assert torch.equal(a[inf], b[inf])
    return ~(inf | torch.isnan(a))


@pytest.mark.parametrize("kind", ["cvrp", "tsp"])
def test_oracle_cur_feature_matches_reference(kind):
    z, prob = _load(kind)
    for t in z["steps"]:
        cur = torch.tensor(z["s%d_cur" % t].astype(np.int64))
        load = torch.tensor(z["s%d_load" % t]) if kind == "cvrp" else None
        out = O.cur_feature(prob, cur, load)
        assert torch.equal(out[0], torch.tensor(z["s%d_dist" % t]))
        assert torch.equal(out[1], torch.tensor(z["s%d_theta" % t]))
        assert torch.equal(out[2], torch.tensor(z["s%d_rel" % t]))
        if kind == "cvrp":
            ref = torch.tensor(z["s%d_nd" % t])
            fin = _same_special(out[3], ref)
            assert torch.equal(out[3][fin], ref[fin])
    if kind == "cvrp":      # the fixture does contain the load == 0 case
        assert sum(int((z["s%d_load" % t] == 0).sum()) for t in z["steps"]) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cvrp", "tsp"])
def test_env_features_match_reference(kind):
    """get_cur_feature / get_local_feature of the drop-in environments (elg_cur_feature) after the recorded steps."""
    z, prob = _load(kind)
    dev = "cuda:0"
    if kind == "cvrp":
        from elg_b200.cvrp import CVRPEnv
        M = z["s%d_cur" % z["steps"][0]].shape[1]
        env = CVRPEnv(M, dev)
        env.load_random_problems({k: torch.tensor(z[k]) for k in ("depot", "loc", "demand")}, aug_factor=8)
        env.reset()
        env.step(torch.zeros(prob.xy.shape[0], M, dtype=torch.long, device=dev))
    else:
        from elg_b200.tsp import TSPEnv
        M = z["s0_cur"].shape[1]
        env = TSPEnv(M, dev)
        env.load_random_problems(torch.tensor(z["problems"]), aug_factor=8)
        env.reset()
    n_nan = n_inf = 0
    for t in z["steps"]:
        cur = torch.tensor(z["s%d_cur" % t].astype(np.int64), device=dev)
        env.step(cur)
        if kind == "cvrp":
            assert torch.equal(env.load.cpu(), torch.tensor(z["s%d_load" % t]))          # sequential fp32 recurrence, bit-exact
            cd, th, rel, nd = [x.cpu() for x in env.get_cur_feature()]
        else:
            cd, th, rel = [x.cpu() for x in env.get_local_feature()]
        ref_d, ref_t, ref_r = (torch.tensor(z["s%d_%s" % (t, k)]) for k in ("dist", "theta", "rel"))
        assert torch.equal(rel, ref_r)                                                   # one fp32 subtraction
        # distances: torch's CPU norm(p=2) rounds like sqrt(fma(dy, dy, dx*dx)) on ~99 % of pairs, otherwise 1 ulp off
        assert (cd - ref_d).abs().max() <= 1.2e-7 * max(1.0, float(ref_d.max()))
        assert float((cd == ref_d).float().mean()) > 0.98
        assert (th - ref_t).abs().max() <= 5e-7                                          # atan2f: <= 2 ulp at |theta| <= pi
        if kind == "cvrp":
            ref_n = torch.tensor(z["s%d_nd" % t])
            fin = _same_special(nd, ref_n)
            assert torch.equal(nd[fin], ref_n[fin])                                      # IEEE division
            n_nan += int(torch.isnan(nd).sum()); n_inf += int(torch.isinf(nd).sum())
    if kind == "cvrp":
        assert n_nan > 0 and n_inf > 0


@dataclass
class _RefStyleState:
    """What the reference's CVRPEnv / TSPEnv hand to the model (CVRP/CVRPEnv.py:22-31, TSP/TSPEnv.py:14-21): plain tensors,
    fp32 ninf_mask of 0 / -inf, nothing of ours attached."""
    selected_count: int = 0
    load: torch.Tensor = None
    current_node: torch.Tensor = None
    ninf_mask: torch.Tensor = None
    finished: torch.Tensor = None
    BATCH_IDX: torch.Tensor = None
    POMO_IDX: torch.Tensor = None


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cvrp", "tsp"])
def test_reference_style_state_drives_our_model(kind):
    """A foreign environment (torch code on the GPU following the oracle's env step, fp32 0/-inf masks) + our model's
    one_step_rollout give the same tours as our fused rollout."""
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_cvrp_batch, synthetic_state_dict, synthetic_tsp_batch
    dev = "cuda:0"
    N = M = 20
    mp = dict(DEFAULT_MODEL_PARAMS[kind])
    sd = synthetic_state_dict(kind, seed=5, gain=3.0)
    if kind == "cvrp":
        from elg_b200.cvrp import CVRPEnv as Env, CVRPModel as Model, rollout
        data = synthetic_cvrp_batch(2, N, seed=9)
        prob = O.load_cvrp(data["depot"], data["loc"], data["demand"], 8)
    else:
        from elg_b200.tsp import TSPEnv as Env, TSPModel as Model, rollout
        data = synthetic_tsp_batch(2, N, seed=9)
        prob = O.load_tsp(data, 8)
    model = Model(**mp)
    model.decoder.add_local_policy(dev)
    model.load_state_dict(sd)
    model = model.to(dev).eval().requires_grad_(False)
    env = Env(M, dev)
    env.load_random_problems(data, 8)
    reset_state, _, _ = env.reset()
    model.pre_forward(reset_state)
    random.seed(11)
    fused, _, fused_r = rollout(model, env, "greedy")
    # foreign environment: the oracle's state machine on GPU tensors
    for name in ("xy", "demand", "dist"):
        v = getattr(prob, name)
        if v is not None:
            setattr(prob, name, v.to(dev))
    B = prob.xy.shape[0]
    model.pre_forward(reset_state)
    random.seed(11)
    acts = []
    if kind == "cvrp":
        st = O.cvrp_reset(prob, M)
        st.load, st.visited, st.masked, st.finished = st.load.to(dev), st.visited.to(dev), st.masked.to(dev), st.finished.to(dev)
    else:
        st = O.tsp_reset(prob, M)
        st.masked = st.masked.to(dev)
    done = False
    while not done:
        ninf = torch.zeros(B, M, prob.xy.shape[1], device=dev).masked_fill(st.masked, float("-inf"))
        rs = _RefStyleState(selected_count=st.count, load=getattr(st, "load", None), current_node=st.cur, ninf_mask=ninf,
                            finished=getattr(st, "finished", None),
                            BATCH_IDX=torch.arange(B, device=dev)[:, None].expand(B, M), POMO_IDX=torch.arange(M, device=dev)[None, :].expand(B, M))
        if kind == "cvrp":
            sel, _ = model.one_step_rollout(rs, None, None, None, norm_demand=None, eval_type="greedy")
            done = O.cvrp_env_step(prob, st, sel)
        else:
            sel, _ = model.one_step_rollout(rs, None, None, None, eval_type="greedy")
            done = O.tsp_env_step(prob, st, sel)
        acts.append(sel)
    tours = torch.stack(acts, dim=2).cpu()
    frac, _ = compare_tours(tours, fused.cpu())
    assert frac == 1.0, frac


@pytest.mark.gpu
def test_our_env_features_drive_reference_style_consumer():
    """The tensors our CVRPEnv hands out, consumed as the reference's local policy consumes them (oracle restatement of
    CVRP/models.py:51-175 taking cur_dist / theta / norm_demand from the caller), give the logits of the oracle that
    builds its own features -- including steps where norm_demand holds nan / inf in masked slots."""
    from elg_b200.cvrp import CVRPEnv
    from elg_b200.synth import DEFAULT_MODEL_PARAMS, synthetic_state_dict
    z, prob = _load("cvrp")
    dev = "cuda:0"
    mp = dict(DEFAULT_MODEL_PARAMS["cvrp"])
    mp["local_size"] = [6]
    W = O.Weights(synthetic_state_dict("cvrp", seed=3, gain=2.0), "cvrp", mp)
    cache = O.decoder_cache(W, O.encode(W, prob))
    M = z["s1_cur"].shape[1]
    env = CVRPEnv(M, dev)
    env.load_random_problems({k: torch.tensor(z[k]) for k in ("depot", "loc", "demand")}, aug_factor=8)
    env.reset()
    st = O.cvrp_reset(prob, M)
    sel0 = torch.zeros(prob.xy.shape[0], M, dtype=torch.long)
    env.step(sel0.to(dev)); O.cvrp_env_step(prob, st, sel0)
    specials = 0
    for t in z["steps"]:
        cur = torch.tensor(z["s%d_cur" % t].astype(np.int64))
        state, _, _ = env.step(cur.to(dev))
        O.cvrp_env_step(prob, st, cur)
        cd, th, rel, nd = [x.cpu() for x in env.get_cur_feature()]
        masked = torch.isinf(state.ninf_mask.cpu())
        assert torch.equal(masked, st.masked)
        specials += int((~torch.isfinite(nd)).sum())
        own = O.decode_logits(W, prob, cache, st.cur, st.masked, st.load)
        fed = O.decode_logits(W, prob, cache, st.cur, st.masked, st.load, feats=(cd, th, nd))
        assert torch.equal(torch.isinf(own), torch.isinf(fed)) and not torch.isnan(fed).any()
        fin = ~torch.isinf(own)
        assert (own[fin] - fed[fin]).abs().max() < 2e-4
    assert specials > 0
