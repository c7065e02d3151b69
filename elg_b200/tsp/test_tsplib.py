"""Drop-in for the reference's TSP/test_tsplib.py (TSP/test_tsplib.py:18-168): one TSPLIB instance at a time (pickles of
[coords (N,2) float64, optimal]), global min-max scaling, POMO width N, x8 augmentation, rounded unscaled tour length, gap
bins <=200 / <=500 / <=1002.  Same class / method names, config keys, printouts and result file; the work is in
elg_b200.lib_driver."""
import os
import pickle

import numpy as np
import torch

from .. import lib_driver as drv
from .TSPEnv import TSPEnv
from .TSPModel import TSPModel
from .utils import rollout

TSP_BINS = [("<=200", 0, 200), ("200-500", 200, 500), ("500-1002", 500, 1002)]


class TSPLib_Tester:

    def __init__(self, config, model=None):
        self.config = config
        self.device = drv.require_cuda(config)
        self.model = drv.load_model(TSPModel, config, self.device, model)
        self.tsplib_path = config.get('tsplib_path', 'TSPLib')
        self.repeat_times = 1
        self.aug_factor = config['params']['aug_factor']
        self.tsplib_results = None

    def test_on_tsplib(self, limit=None, out_dir='test_results', max_scale=None):
        entries = []
        for fname in sorted(f for f in os.listdir(self.tsplib_path) if f.endswith('.pkl'))[:limit]:
            with open(os.path.join(self.tsplib_path, fname), 'rb') as f:
                instance = pickle.load(f)
            if not (max_scale and len(instance[0]) > max_scale):
                entries.append((fname[:-4], instance[1], instance))
        results, total = drv.run_set(entries, lambda n, inst, rec: self.test_on_one_ins(n, rec, inst), self.repeat_times)
        drv.dump_results(results, out_dir, self.config['name'] + '_tsplib.json')
        kept = [r for r in results if r['record'][-1]['scale'] <= 1002]
        bins = drv.gap_bins(kept, TSP_BINS)
        print("Total average cost {:.2f}".format(np.mean([r['record'][-1]['best_cost'] for r in kept])))
        print("Total average gap {:.2f}%".format(bins['total']))
        for label, _, _ in TSP_BINS:
            if label in bins:
                print("{} average gap {:.2f}%".format(label, bins[label]))
        print("Average time: {:.2f}s".format(total / max(len(results), 1)))
        self.tsplib_results = results
        return results

    def test_on_one_ins(self, name, result_dict, instance):
        coords, optimal = instance[0], instance[1]
        unscaled = torch.tensor(coords, dtype=torch.float)[None]
        scaled = (coords - np.min(coords)) / (np.max(coords) - np.min(coords))        # one scale for both axes, TSP/test_tsplib.py:128
        problem_size = scaled.shape[0]
        env = TSPEnv(problem_size, self.device)
        env.load_tsplib_problem(torch.tensor(scaled, dtype=torch.float)[None], unscaled, self.aug_factor)
        best, solutions, rewards = drv.solve_instance(self.model, env, rollout, self.aug_factor, problem_size)
        drv.fill_record(result_dict, best, problem_size, optimal)
        return solutions, rewards


if __name__ == "__main__":
    import yaml
    with open('config.yml', 'r', encoding='utf-8') as config_file:
        config = yaml.load(config_file.read(), Loader=yaml.FullLoader)
    TSPLib_Tester(config=config).test_on_tsplib()
