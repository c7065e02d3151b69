"""Drop-in for the reference's TSP/test_tsplib.py (TSP/test_tsplib.py:18-168): one TSPLIB instance at a time
(pickles of [coords (N,2) float64, optimal]), global min-max scaling, POMO width N, x8 augmentation,
rounded unscaled tour length, gap bins <=200 / <=500 / <=1002."""
import json
import os
import pickle
import time

import numpy as np
import torch

from .TSPEnv import TSPEnv
from .TSPModel import TSPModel
from .utils import rollout


class TSPLib_Tester:

    def __init__(self, config, model=None):
        self.config = config
        model_params = config['model_params']
        if not config.get('use_cuda', True):
            raise RuntimeError("elg_b200 has no CPU path: set use_cuda: True")
        self.device = torch.device('cuda', config['cuda_device_num'])
        torch.cuda.set_device(self.device)
        if model is None:
            model = TSPModel(**model_params)
            model.decoder.add_local_policy(self.device)
            checkpoint = torch.load(config['load_checkpoint'], map_location=self.device)
            model.load_state_dict(checkpoint['model_state_dict'])
        self.model = model.to(self.device)
        self.tsplib_path = config.get('tsplib_path', 'TSPLib')
        self.repeat_times = 1
        self.aug_factor = config['params']['aug_factor']
        self.tsplib_results = None

    def test_on_tsplib(self, limit=None, out_dir='test_results', max_scale=None):
        files = sorted(f for f in os.listdir(self.tsplib_path) if f.endswith('.pkl'))
        tsplib_results, total_time = [], 0.
        for t in range(self.repeat_times):
            for fname in files[:limit] if limit else files:
                name = fname[:-4]
                with open(os.path.join(self.tsplib_path, fname), 'rb') as f:
                    instance = pickle.load(f)
                if max_scale and len(instance[0]) > max_scale:
                    continue
                result_dict = {'run_idx': t}
                start_time = time.time()
                self.test_on_one_ins(name=name, result_dict=result_dict, instance=instance)
                torch.cuda.synchronize()
                result_dict['seconds'] = time.time() - start_time
                total_time += result_dict['seconds']
                tsplib_results.append({'instance': name, 'optimal': instance[1], 'record': [result_dict]})
                print("Instance Name {}: gap {:.4f}".format(name, result_dict['gap']))
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, self.config['name'] + '_tsplib.json'), 'w') as f:
                json.dump(tsplib_results, f)
        cost = np.array([r['record'][-1]['best_cost'] for r in tsplib_results])
        opt = np.array([r['optimal'] for r in tsplib_results])
        scale = np.array([r['record'][-1]['scale'] for r in tsplib_results])
        gap = (cost - opt) / opt
        keep = scale <= 1002
        print("Total average cost {:.2f}".format(cost[keep].mean()))
        print("Total average gap {:.2f}%".format(100 * gap[keep].mean()))
        for label, sel in (("<=200", scale <= 200), ("200-500", (scale > 200) & (scale <= 500)),
                           ("500-1002", (scale > 500) & (scale <= 1002))):
            if sel.any():
                print("{} average gap {:.2f}%".format(label, 100 * gap[sel].mean()))
        print("Average time: {:.2f}s".format(total_time / max(len(tsplib_results), 1)))
        self.tsplib_results = tsplib_results
        return tsplib_results

    def test_on_one_ins(self, name, result_dict, instance):
        unscaled_points = torch.tensor(instance[0], dtype=torch.float)[None, :, :]
        points = (instance[0] - np.min(instance[0])) / (np.max(instance[0]) - np.min(instance[0]))
        test_batch = torch.tensor(points, dtype=torch.float)[None, :, :]
        optimal = instance[1]
        problem_size = test_batch.shape[1]
        pomo_size = problem_size
        env = TSPEnv(pomo_size, self.device)
        env.load_tsplib_problem(test_batch, unscaled_points, self.aug_factor)
        reset_state, reward, done = env.reset()
        self.model.eval()
        self.model.requires_grad_(False)
        self.model.pre_forward(reset_state)
        policy_solutions, policy_prob, rewards = rollout(self.model, env, 'greedy')
        aug_reward = rewards.reshape(self.aug_factor, 1, pomo_size)
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
    TSPLib_Tester(config=config).test_on_tsplib()
