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
# The reference's Dataset classes (CVRP/generate_data.py:108-171, TSP/generate_data.py:74-99): same constructor,
# same items; files are read on the host, generated data comes from the device generators above.
import os
import pickle

from torch.utils.data import Dataset


def check_extension(filename):
    return filename if os.path.splitext(filename)[1] == ".pkl" else filename + ".pkl"


def save_dataset(dataset, filename):
    filedir = os.path.split(filename)[0]
    if filedir and not os.path.isdir(filedir):
        os.makedirs(filedir)
    with open(check_extension(filename), 'wb') as f:
        pickle.dump(dataset, f, pickle.HIGHEST_PROTOCOL)


def make_instance(args):
    depot, loc, demand, capacity, *args = args
    grid_size = 1
    if len(args) > 0:
        depot_types, customer_types, grid_size = args
    return {'loc': torch.tensor(loc, dtype=torch.float) / grid_size,
            'demand': torch.tensor(demand, dtype=torch.float) / capacity,
            'depot': torch.tensor(depot, dtype=torch.float) / grid_size}


class VRPDataset(Dataset):
    def __init__(self, filename=None, size=100, num_samples=10000, offset=0, distribution=None, device="cuda:0", seed=None):
        super(VRPDataset, self).__init__()
        if filename is not None:
            assert os.path.splitext(filename)[1] == '.pkl'
            with open(filename, 'rb') as f:
                data = pickle.load(f)
            self.data = [make_instance(args) for args in data[offset:offset + num_samples]]
        else:
            dist = distribution if distribution is not None else {"data_type": "uniform"}
            data = generate_vrp_data(num_samples, size, dist, device, seed)
            self.data = [make_instance([data['depot'][i, 0].cpu().numpy(), data['loc'][i].cpu().numpy(),
                                        data['demand'][i].cpu().numpy(), 1.0]) for i in range(num_samples)]
        self.size = len(self.data)

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return self.data[idx]


class TSPDataset(Dataset):
    def __init__(self, filename=None, size=100, num_samples=10000, offset=0, distribution=None, device="cuda:0", seed=None):
        super(TSPDataset, self).__init__()
        if filename is not None:
            assert os.path.splitext(filename)[1] == '.pkl'
            with open(filename, 'rb') as f:
                data = pickle.load(f)
            self.data = [torch.FloatTensor(row) for row in data[offset:offset + num_samples]]
        else:
            dist = distribution if distribution is not None else {"data_type": "uniform"}
            data = generate_tsp_data(num_samples, size, dist, device, seed).cpu()
            self.data = [data[i] for i in range(num_samples)]
        self.size = len(self.data)

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return self.data[idx]
