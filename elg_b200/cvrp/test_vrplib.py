"""Drop-in for the reference's CVRP/test_vrplib.py (CVRP/test_vrplib.py:15-151): one CVRPLIB instance at a time, POMO width
min(N, 1000), x8 augmentation, rounded unscaled cost, gap to the .sol optimum, size bins.  Same class / method names, config
keys, printouts and result file; the work is in elg_b200.lib_driver."""
import os

import torch

from .. import lib_driver as drv
from .. import vrplib_io as vrplib
from .CVRPEnv import CVRPEnv
from .CVRPModel import CVRPModel
from .utils import rollout

X_BINS = [("<200", 0, 200), ("200-500", 200, 500), ("500-1000", 500, 10 ** 9)]


class VRPLib_Tester:

    def __init__(self, config, model=None):
        self.config = config
        self.device = drv.require_cuda(config)
        self.model = drv.load_model(CVRPModel, config, self.device, model, needs_local=config['model_params']['ensemble'])
        self.vrplib_path = config.get('vrplib_path') or ('VRPLib/Vrp-Set-X/' if config['vrplib_set'] == 'X' else "VRPLib/Vrp-Set-XXL/")
        self.repeat_times = 1
        self.aug_factor = config['params']['aug_factor']
        self.vrplib_results = None

    def test_on_vrplib(self, limit=None, out_dir='test_results'):
        names = sorted(f[:-4] for f in os.listdir(self.vrplib_path) if f.endswith('.vrp'))[:limit]
        path = lambda n, ext: os.path.join(self.vrplib_path, n + ext)
        entries = [(n, vrplib.read_solution(path(n, '.sol'))['cost'], n) for n in names]
        xxl = 'XXL' in self.vrplib_path
        results, total = drv.run_set(entries, lambda n, _, rec: self.test_on_one_ins(n, rec, path(n, '.vrp'), path(n, '.sol')),
                                     self.repeat_times, echo_cost=xxl)
        bins = drv.gap_bins(results, [] if xxl else X_BINS)
        summary = dict(bins, avg_time_s=total / max(len(results), 1))
        if xxl:
            print("{:.2f}%".format(bins['total']))
        else:
            for label, _, _ in X_BINS:
                if label in bins:
                    print("Average gap on subset of {}: {:.2f}%".format(label, bins[label]))
            print("Average gap total: {:.2f}%".format(bins['total']))
        print("Average time: {:.2f}s".format(summary['avg_time_s']))
        results.append(summary)
        drv.dump_results(results, out_dir, self.config['name'] + '_vrplib.json')
        self.vrplib_results = results
        return results

    def test_on_one_ins(self, name, result_dict, instance, solution):
        """instance: path of a .vrp file or a parsed dict; solution: path of the .sol file or the optimal cost."""
        if isinstance(instance, str):
            instance = vrplib.read_instance(instance)
        optimal = vrplib.read_solution(solution)['cost'] if isinstance(solution, str) else solution
        problem_size = instance['node_coord'].shape[0] - 1
        env = CVRPEnv(min(problem_size, 1000), self.device)            # CVRP/test_vrplib.py:116
        env.load_vrplib_problem(instance, aug_factor=self.aug_factor)
        best, solutions, rewards = drv.solve_instance(self.model, env, rollout, self.aug_factor, env.multi_width)
        drv.fill_record(result_dict, best, problem_size, optimal)
        return solutions, rewards


if __name__ == "__main__":
    import yaml
    with open('config.yml', 'r', encoding='utf-8') as config_file:
        config = yaml.load(config_file.read(), Loader=yaml.FullLoader)
    VRPLib_Tester(config=config).test_on_vrplib()
