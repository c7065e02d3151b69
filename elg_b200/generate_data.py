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


# ---------------------------------------------------------------------------------------------- datasets
# Drop-ins for the reference's dataset helpers (CVRP/generate_data.py:91-171, TSP/generate_data.py:59-99): the same
# names, constructor arguments, pickle format ([depot, loc, demand, capacity(, types, types, grid)] rows for cvrp,
# coordinate lists for tsp) and items; implemented over one tensor-backed base class.
import os
import pickle

from torch.utils.data import Dataset


def check_extension(filename):
    """`name` -> `name.pkl` unless it already ends in .pkl"""
    root, ext = os.path.splitext(filename)
    return filename if ext == ".pkl" else filename + ".pkl"


def save_dataset(dataset, filename):
    """Pickle `dataset` (highest protocol) to check_extension(filename), creating the directory."""
    target = check_extension(filename)
    os.makedirs(os.path.dirname(target) or ".", exist_ok=True)
    with open(target, "wb") as fh:
        pickle.dump(dataset, fh, pickle.HIGHEST_PROTOCOL)


def make_instance(args):
    """One pickled cvrp row -> {'loc', 'demand', 'depot'} float tensors: demands over capacity, coordinates over the grid size."""
    depot, loc, demand, capacity = args[:4]
    grid = args[6] if len(args) > 4 else 1
    as_f = lambda v: torch.as_tensor(v, dtype=torch.float)
    return {"loc": as_f(loc) / grid, "demand": as_f(demand) / capacity, "depot": as_f(depot) / grid}


class _ListDataset(Dataset):
    """Instances from a reference .pkl file (rows [offset, offset + num_samples)) or from the device generators."""

    def __init__(self, filename, size, num_samples, offset, distribution, device, seed):
        super().__init__()
        if filename is not None:
            if os.path.splitext(filename)[1] != ".pkl":
                raise AssertionError("dataset files are .pkl")
            with open(filename, "rb") as fh:
                rows = pickle.load(fh)[offset:offset + num_samples]
            self.data = [self._from_row(r) for r in rows]
        else:
            self.data = self._generate(num_samples, size, distribution or {"data_type": "uniform"}, device, seed)
        self.size = len(self.data)

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return self.data[idx]


class VRPDataset(_ListDataset):
    def __init__(self, filename=None, size=100, num_samples=10000, offset=0, distribution=None, device="cuda:0", seed=None):
        super().__init__(filename, size, num_samples, offset, distribution, device, seed)

    _from_row = staticmethod(make_instance)

    @staticmethod
    def _generate(n, size, dist, device, seed):
        d = {k: v.cpu() for k, v in generate_vrp_data(n, size, dist, device, seed).items()}
        return [{"loc": d["loc"][i], "demand": d["demand"][i], "depot": d["depot"][i, 0]} for i in range(n)]


class TSPDataset(_ListDataset):
    def __init__(self, filename=None, size=100, num_samples=10000, offset=0, distribution=None, device="cuda:0", seed=None):
        super().__init__(filename, size, num_samples, offset, distribution, device, seed)

    @staticmethod
    def _from_row(row):
        return torch.as_tensor(row, dtype=torch.float)

    @staticmethod
    def _generate(n, size, dist, device, seed):
        return list(generate_tsp_data(n, size, dist, device, seed).cpu())
