"""Drop-in for the reference's CVRP/test.py evaluation loop (CVRP/test.py:14-56).

`solve_batch` is one iteration of that loop (load + x8 augmentation, encode, greedy rollout,
best-of-POMO and best-of-augmentation); `test` is the loop itself with the reference's printout.
Run as a script it reads `config.yml` from the current directory like the reference.
"""
import time

import torch

from .CVRPEnv import CVRPEnv
from .CVRPModel import CVRPModel
from .utils import rollout


def solve_batch(model, env, batch, aug_factor):
    """-> (no_aug_cost (n,), aug_cost (n,), solutions (aug*n, M, T), rewards (aug*n, M))"""
    n = batch['loc'].shape[0]
    env.load_random_problems(batch, aug_factor)
    reset_state, _, _ = env.reset()
    with torch.no_grad():
        model.pre_forward(reset_state)
        solutions, probs, rewards = rollout(model=model, env=env, eval_type='greedy')
    aug_reward = rewards.reshape(aug_factor, n, env.multi_width)
    max_pomo_reward, _ = aug_reward.max(dim=2)          # best of POMO
    no_aug_cost = -max_pomo_reward[0, :].float()
    max_aug_pomo_reward, _ = max_pomo_reward.max(dim=0)  # best of augmentation
    aug_cost = -max_aug_pomo_reward.float()
    return no_aug_cost, aug_cost, solutions, rewards


def test(dataloader, model, env, aug_factor):
    model.eval()
    model.requires_grad_(False)
    avg_cost_total, no_avg_cost_total, t = 0., 0., 0
    start = time.time()
    for batch in dataloader:
        no_aug_cost, aug_cost, _, _ = solve_batch(model, env, batch, aug_factor)
        avg_cost_total += aug_cost.mean()
        no_avg_cost_total += no_aug_cost.mean()
        t += 1
    torch.cuda.synchronize()
    end = time.time()
    avg_cost_total /= t
    no_avg_cost_total /= t
    print("Aug cost: {:.4f}".format(avg_cost_total))
    print("no aug Avg cost: {:.4f}, Wall-clock time: {:.2f}s".format(no_avg_cost_total, float(end - start)))
    return avg_cost_total


def load_model(config, device):
    """Checkpoint loading exactly as the reference (CVRP/test.py:73-79)."""
    model_params = config['model_params']
    model = CVRPModel(**model_params)
    checkpoint = torch.load(config['load_checkpoint'], map_location=device)
    if model_params['ensemble']:
        model.decoder.add_local_policy(device)
    model.load_state_dict(checkpoint['model_state_dict'])
    return model.to(device)


if __name__ == "__main__":
    import pickle
    import yaml
    with open('config.yml', 'r', encoding='utf-8') as f:
        config = yaml.load(f.read(), Loader=yaml.FullLoader)
    device = "cuda:{}".format(config['cuda_device_num'])
    model = load_model(config, device)
    env = CVRPEnv(multi_width=config['params']['multiple_width'], device=device)
    with open(config['test_filename'], 'rb') as f:
        data = pickle.load(f)[:config['params']['test_size']]
    bs = config['params']['test_batch_size']
    batches = [{'depot': torch.FloatTensor([d[0] for d in data[i:i + bs]]),
                'loc': torch.FloatTensor([d[1] for d in data[i:i + bs]]),
                'demand': torch.FloatTensor([d[2] for d in data[i:i + bs]]) / float(data[0][3])}
               for i in range(0, len(data), bs)]
    test(batches, model, env, aug_factor=config['params']['aug_factor'])
