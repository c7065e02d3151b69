"""Drop-in for the reference's TSP/utils.py: rollout (TSP/utils.py:7-26), x8 augmentation, seeding."""
import random

import numpy as np
import torch

from .. import engine
from ..cvrp.utils import augment_xy_data_by_8_fold, seed_everything  # noqa: F401  (same functions in both sub-projects)


def rollout(model, env, eval_type='greedy'):
    if getattr(model, "_elg_fused", False) and getattr(env, "_elg_fused", False):
        return _fused_rollout(model, env, eval_type)
    env.reset()
    actions, probs = [], []
    state, reward, done = env.pre_step()
    while not done:
        cur_dist, cur_theta, xy = env.get_local_feature()
        selected, one_step_prob = model.one_step_rollout(state, cur_dist=cur_dist, cur_theta=cur_theta, xy=xy,
                                                         eval_type=eval_type)
        state, reward, done = env.step(selected)
        actions.append(selected)
        probs.append(one_step_prob)
    actions = torch.stack(actions, 1)
    probs = None if eval_type == 'greedy' else torch.stack(probs, 1)
    return torch.transpose(actions, 1, 2), probs, reward


def _fused_rollout(model, env, eval_type):
    env.reset()
    batch = model._batch
    if batch is None or batch.xy.data_ptr() != env.problems.data_ptr():
        raise RuntimeError("model.pre_forward(reset_state) must be called on this env's problems before rollout")
    M = env.pomo_size
    start = random.sample(range(0, M), M)            # TSP/TSPModel.py:31
    batch.tables.unscaled = env._unscaled_dev.data_ptr() if env.tsplib else None
    tours16, reward, logp, n_steps = engine.rollout(batch, M, start, mode=eval_type, seed=model._next_seed())
    solutions = tours16[:, :, :env.problem_size].long()
    env._finish_fused(solutions)
    probs = None if eval_type == 'greedy' else _probs_from_logp(logp, env.problem_size)
    model._last_logp = logp
    return solutions, probs, reward


def check_feasible(pi):
    """Every node exactly once (TSP/utils.py:72-78); input (1, multi, problem)."""
    pi = pi.squeeze(0)
    want = torch.arange(pi.size(1), device=pi.device).view(1, -1).expand_as(pi)
    assert (want == pi.data.sort(1)[0]).all(), "Invalid tour"
