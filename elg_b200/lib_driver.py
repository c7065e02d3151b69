"""Shared machinery of the two library drivers (elg_b200.cvrp.test_vrplib / elg_b200.tsp.test_tsplib), which keep the
reference's class names, config keys, result-file layout and printouts (CVRP/test_vrplib.py:15-151,
TSP/test_tsplib.py:18-168) but are thin shells around the functions here:

    solve_instance(model, env, aug)      one greedy rollout of a loaded instance -> (best cost over aug x POMO, tours, rewards)
    run_set(entries, solve, ...)         the per-instance loop with timing and the record dicts of the result files
    gap_bins(results, edges, ...)        mean gap per size bin, as the reference prints them
"""
import json
import os
import time

import numpy as np
import torch


def require_cuda(config):
    if not config.get('use_cuda', True):
        raise RuntimeError("elg_b200 has no CPU path: set use_cuda: True")
    device = torch.device('cuda', config['cuda_device_num'])
    torch.cuda.set_device(device)
    return device


def load_model(model_cls, config, device, model=None, needs_local=True):
    """The reference's checkpoint protocol: add_local_policy BEFORE load_state_dict (CVRP/test.py:75-78)."""
    if model is None:
        model = model_cls(**config['model_params'])
        if needs_local:
            model.decoder.add_local_policy(device)
        model.load_state_dict(torch.load(config['load_checkpoint'], map_location=device)['model_state_dict'])
    model = model.to(device)
    model.eval()
    model.requires_grad_(False)
    return model


def solve_instance(model, env, rollout, aug_factor, width):
    """env already holds the instance.  Best of POMO, then best of augmentation (CVRP/test_vrplib.py:127-137)."""
    reset_state, _, _ = env.reset()
    model.pre_forward(reset_state)
    with torch.no_grad():
        solutions, _, rewards = rollout(model, env, 'greedy')
    best = -rewards.reshape(aug_factor, 1, width).max(dim=2)[0].max(dim=0)[0].float()
    return float(best.cpu()[0]), solutions, rewards


def run_set(entries, solve_one, repeat_times=1, echo_cost=False):
    """entries: iterable of (name, optimal, payload).  solve_one(name, payload, record) fills record['best_cost' / 'scale' / 'gap']."""
    results, total = [], 0.0
    for run_idx in range(repeat_times):
        for name, optimal, payload in entries:
            record = {'run_idx': run_idx}
            t0 = time.time()
            solve_one(name, payload, record)
            torch.cuda.synchronize()
            record['seconds'] = time.time() - t0
            total += record['seconds']
            results.append({'instance': name, 'optimal': optimal, 'record': [record]})
            print("Instance Name {}: gap {:.4f}".format(name, record['gap']))
            if echo_cost:
                print("cost: {}".format(record['best_cost']))
    return results, total


def fill_record(record, best_cost, scale, optimal):
    if record is not None:
        record['best_cost'] = best_cost
        record['scale'] = scale
        record['gap'] = (best_cost - optimal) / optimal


def gap_bins(results, edges):
    """edges: [(label, lo_exclusive, hi_inclusive)] on the instance scale -> {label: mean gap in %}, plus 'total'."""
    gap = np.array([r['record'][-1]['gap'] for r in results])
    scale = np.array([int(r['record'][-1]['scale']) for r in results])
    out = {}
    for label, lo, hi in edges:
        sel = (scale > lo) & (scale <= hi)
        if sel.any():
            out[label] = 100 * float(gap[sel].mean())
    out['total'] = 100 * float(gap.mean()) if len(gap) else float('nan')
    return out


def dump_results(results, out_dir, filename):
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, filename), 'w') as f:
            json.dump(results, f)
