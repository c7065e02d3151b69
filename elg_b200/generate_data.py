"""Drop-in for the reference's training-data generators (CVRP/generate_data.py:9-92, TSP/generate_data.py:9-57),
running on the device (elg_generate_problems).  Same signatures plus `device` / `seed`; same distribution dict
(config.yml `distribution`), same CAPACITIES table (KeyError for a problem_size that is not in it)."""
import ctypes as C

import torch

from . import _lib
from ._lib import ELG_CVRP, ELG_TSP, check, lib

# From VRP with RL paper https://arxiv.org/abs/1802.04240 (CVRP/generate_data.py:75-83)
CAPACITIES = {10: 20., 20: 30., 50: 40., 100: 50., 200: 80., 500: 100., 1000: 250.}
_KINDS = {"uniform": 0, "cluster": 1, "mixed": 2}
_counter = [0]


def _kind(distribution):
    dt = distribution["data_type"]
    dt = dt[0] if not isinstance(dt, str) else dt        # train.py stores np.random.choice(...) output (a 1-element array)
    if dt not in _KINDS:
        raise KeyError(dt)
    k = _KINDS[dt]
    nc = {0: 1, 1: distribution.get("n_cluster", 3), 2: distribution.get("n_cluster_mix", 1)}[k]
    return k, int(nc)


def _seed(seed):
    if seed is None:
        _counter[0] += 1
        seed = torch.initial_seed() * 1000003 + _counter[0]
    return int(seed) & (2 ** 64 - 1)


def _gen(problem, batch_size, problem_size, distribution, device, seed, capacity):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.ElgError("the generators run on a CUDA device; there is no CPU path")
    kind, nc = _kind(distribution)
    node = torch.empty((batch_size, problem_size, 2), dtype=torch.float32, device=dev)
    depot = torch.empty((batch_size, 1, 2), dtype=torch.float32, device=dev) if problem == ELG_CVRP else None
    demand = torch.empty((batch_size, problem_size), dtype=torch.float32, device=dev) if problem == ELG_CVRP else None
    p = lambda t: C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        check(lib.elg_generate_problems(problem, kind, batch_size, problem_size, nc, float(distribution.get("lower", 0.2)),
                                        float(distribution.get("upper", 0.8)), float(distribution.get("std", 0.07)),
                                        float(capacity), _seed(seed), p(depot), p(node), p(demand),
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return depot, node, demand


def generate_vrp_data(batch_size, problem_size, distribution, device="cuda:0", seed=None):
    capacity = CAPACITIES[problem_size]
    depot, node, demand = _gen(ELG_CVRP, batch_size, problem_size, distribution, device, seed, capacity)
    return {"loc": node, "demand": demand, "depot": depot}


def generate_tsp_data(batch_size, problem_size, distribution, device="cuda:0", seed=None):
    return _gen(ELG_TSP, batch_size, problem_size, distribution, device, seed, 1.0)[1]
