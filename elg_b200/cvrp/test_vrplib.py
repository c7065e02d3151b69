"""Drop-in for the reference's CVRP/test_vrplib.py (CVRP/test_vrplib.py:15-151): one CVRPLIB instance at a
time, POMO width min(N, 1000), x8 augmentation, rounded unscaled cost, gap to the .sol optimum, size bins."""
import json
import os
import time

import numpy as np
import torch

from .. import vrplib_io as vrplib
from .CVRPEnv import CVRPEnv
from .CVRPModel import CVRPModel
from .utils import rollout


class VRPLib_Tester:

    def __init__(self, config, model=None):
        self.config = config
        model_params = config['model_params']
        if not config.get('use_cuda', True):
            raise RuntimeError("elg_b200 has no CPU path: set use_cuda: True")
        self.device = torch.device('cuda', config['cuda_device_num'])
        torch.cuda.set_device(self.device)
        if model is None:
            model = CVRPModel(**model_params)
            if model_params['ensemble']:
                model.decoder.add_local_policy(self.device)
            checkpoint = torch.load(config['load_checkpoint'], map_location=self.device)
            model.load_state_dict(checkpoint['model_state_dict'])
        self.model = model.to(self.device)
        self.vrplib_path = config.get('vrplib_path') or ('VRPLib/Vrp-Set-X/' if config['vrplib_set'] == 'X' else "VRPLib/Vrp-Set-XXL/")
        self.repeat_times = 1
        self.aug_factor = config['params']['aug_factor']
        self.vrplib_results = None

    def test_on_vrplib(self, limit=None, out_dir='test_results'):
        files = sorted(f for f in os.listdir(self.vrplib_path) if f.endswith('.vrp'))
        if limit:
            files = files[:limit]
        vrplib_results, total_time = [], 0.
        for t in range(self.repeat_times):
            for fname in files:
                name = fname[:-4]
                instance_file = os.path.join(self.vrplib_path, name + '.vrp')
                solution_file = os.path.join(self.vrplib_path, name + '.sol')
                optimal = vrplib.read_solution(solution_file)['cost']
                result_dict = {'run_idx': t}
                start_time = time.time()
                self.test_on_one_ins(name=name, result_dict=result_dict, instance=instance_file, solution=solution_file)
                torch.cuda.synchronize()
                result_dict['seconds'] = time.time() - start_time
                total_time += result_dict['seconds']
                vrplib_results.append({'instance': name, 'optimal': optimal, 'record': [result_dict]})
                print("Instance Name {}: gap {:.4f}".format(name, result_dict['gap']))
                if 'XXL' in self.vrplib_path:
                    print("cost: {}".format(result_dict['best_cost']))
        gaps = np.array([r['record'][-1]['gap'] for r in vrplib_results])
        scale = np.array([int(r['record'][-1]['scale']) for r in vrplib_results])
        summary = {"total": 100 * float(gaps.mean()), "avg_time_s": total_time / max(len(vrplib_results), 1)}
        if 'XXL' in self.vrplib_path:
            print("{:.2f}%".format(100 * gaps.mean()))
        else:
            for label, sel in (("<200", scale <= 200), ("200-500", (scale > 200) & (scale <= 500)), ("500-1000", scale > 500)):
                if sel.any():
                    summary[label] = 100 * float(gaps[sel].mean())
                    print("Average gap on subset of {}: {:.2f}%".format(label, summary[label]))
            print("Average gap total: {:.2f}%".format(summary["total"]))
        print("Average time: {:.2f}s".format(summary["avg_time_s"]))
        vrplib_results.append(summary)
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, self.config['name'] + '_vrplib.json'), 'w') as f:
                json.dump(vrplib_results, f)
        self.vrplib_results = vrplib_results
        return vrplib_results

    def test_on_one_ins(self, name, result_dict, instance, solution):
        instance = vrplib.read_instance(instance) if isinstance(instance, str) else instance
        optimal = vrplib.read_solution(solution)['cost'] if isinstance(solution, str) else solution
        problem_size = instance['node_coord'].shape[0] - 1
        multiple_width = min(problem_size, 1000)
        env = CVRPEnv(multiple_width, self.device)
        env.load_vrplib_problem(instance, aug_factor=self.aug_factor)
        reset_state, reward, done = env.reset()
        self.model.eval()
        self.model.requires_grad_(False)
        self.model.pre_forward(reset_state)
        with torch.no_grad():
            policy_solutions, policy_prob, rewards = rollout(self.model, env, 'greedy')
        aug_reward = rewards.reshape(self.aug_factor, 1, env.multi_width)
        max_pomo_reward, _ = aug_reward.max(dim=2)
        max_aug_pomo_reward, _ = max_pomo_reward.max(dim=0)
        best_cost = -max_aug_pomo_reward.float()
        if result_dict is not None:
            result_dict['best_cost'] = best_cost.cpu().numpy().tolist()[0]
            result_dict['scale'] = problem_size
            result_dict['gap'] = (result_dict['best_cost'] - optimal) / optimal
        return policy_solutions, rewards


if __name__ == "__main__":
    import yaml
    with open('config.yml', 'r', encoding='utf-8') as config_file:
        config = yaml.load(config_file.read(), Loader=yaml.FullLoader)
    VRPLib_Tester(config=config).test_on_vrplib()
