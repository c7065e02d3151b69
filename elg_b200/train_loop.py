"""Training driver shared by elg_b200/cvrp/train.py and elg_b200/tsp/train.py: the reference's `train()` /
`validate()` / `Logger` (CVRP/train.py:22-148, TSP/train.py:20-146, CVRP/utils.py:130-151) on top of `Trainer`.

Same config.yml keys, same checkpoint dict ('step', 'model_state_dict', 'optimizer_state_dict'), same JSON log
('result': val_100 / val_200 / val_500 lists), same mixed-distribution sampling (softmax of the validation gaps).
`training: joint` follows the reference: the global policy (+ distance penalty) trains alone until step T, then the
local policy is added with fresh parameters and a new optimizer (CVRP/train.py:91-95); `only_global` never adds it.
Differences, stated once: (1) `only_local` (CVRPModel_local) is not implemented; (2) validation files under data/ are used when present, otherwise seeded
batches from the device generators stand in for them; (3) with torch.distributed initialised every rank trains on its
own instances and the gradient is all-reduced (the reference has no multi-GPU training).
"""
import json
import os
import pickle
import random

import numpy as np
import torch

from . import generate_data as gen
from .trainer import Trainer


class Logger(object):
    def __init__(self, filename, config):
        self.filename = filename
        self.logger = dict(config)
        self.logger['result'] = {'val_100': [], 'val_200': [], 'val_500': []}

    def log(self, info):
        for k, v in zip(('val_100', 'val_200', 'val_500'), info):
            self.logger['result'][k].append(v)
        os.makedirs(os.path.dirname(self.filename) or '.', exist_ok=True)
        with open(self.filename, 'w') as f:
            json.dump(self.logger, f)


def softmax(x):
    return np.exp(x) / np.sum(np.exp(x), axis=0)


def _batches(problem, path, size, n, batch, distribution, device, seed):
    """Validation batches: the reference's pickled set if it exists, else a seeded generated one."""
    if path and os.path.exists(path):
        with open(path, 'rb') as f:
            data = pickle.load(f)[:n]
        for i in range(0, len(data), batch):
            chunk = data[i:i + batch]
            if problem == "cvrp":
                yield {'depot': torch.FloatTensor([d[0] for d in chunk]).to(device)[:, None, :],
                       'loc': torch.FloatTensor([d[1] for d in chunk]).to(device),
                       'demand': (torch.FloatTensor([d[2] for d in chunk]) / float(chunk[0][3])).to(device)}
            else:
                yield torch.FloatTensor(chunk).to(device)
        return
    for i in range(0, n, batch):
        nb = min(batch, n - i)
        if problem == "cvrp":
            yield gen.generate_vrp_data(nb, size, distribution, device, seed=seed + i)
        else:
            yield gen.generate_tsp_data(nb, size, distribution, device, seed=seed + i)


def test_rollout(problem, batches, env, model):
    from .cvrp.utils import rollout as cvrp_rollout
    from .tsp.utils import rollout as tsp_rollout
    rollout = cvrp_rollout if problem == "cvrp" else tsp_rollout
    avg_cost, num_batch = 0., 0.
    for batch in batches:
        env.load_random_problems(batch)
        reset_state, _, _ = env.reset()
        model.eval()
        with torch.no_grad():
            model.pre_forward(reset_state)
            solutions, probs, rewards = rollout(model=model, env=env, eval_type='greedy')
        avg_cost += float(-rewards.max(1)[0].mean())
        num_batch += 1.
    return avg_cost / max(num_batch, 1.)


def validate(problem, trainer, model_params, multiple_width, device, mixed, distribution, val_samples=(1000, 1000, 100)):
    if problem == "cvrp":
        from .cvrp import CVRPEnv as Env, CVRPModel as Model
        pre = 'vrp'
    else:
        from .tsp import TSPEnv as Env, TSPModel as Model
        pre = 'tsp'
    model = Model(**model_params)
    if trainer.has_local:
        model.decoder.add_local_policy(device)
    model.load_state_dict(trainer.state_dict())
    model = model.to(device).requires_grad_(False)
    out = []
    if mixed:
        sets = [('data/%s_uniform100_1000_seed1234.pkl' % pre, 100, val_samples[0], 'uniform'),
                ('data/%s_cluster100_1000_seed1234.pkl' % pre, 100, val_samples[1], 'cluster'),
                ('data/%s_mixed100_1000_seed1234.pkl' % pre, 100, val_samples[2], 'mixed')]
    else:      # file names as the reference spells them: CVRP/train.py:65-69 (vrp100_val), TSP/train.py:62-66 (tsp_100_val)
        sep = '' if problem == "cvrp" else '_'
        sets = [('data/%s%s%d_val.pkl' % (pre, sep, size), size, n, 'uniform') for size, n in zip((100, 200, 500), val_samples)]
    env = Env(multiple_width, device)      # one environment of width multiple_width for every set (CVRP/train.py:43, TSP/train.py:42)
    for path, size, n, dt in sets:
        if not os.path.exists(path):
            print("validation set %s not found: using a seeded generated set of the same size / distribution" % path, flush=True)
        d = dict(distribution, data_type=dt)
        batch = 10 if size == 500 else (1000 if (problem == "cvrp" or mixed) else 500)      # the reference's DataLoader batch sizes
        out.append(test_rollout(problem, _batches(problem, path, size, n, batch, d, device, 1234), env, model))
    return out


def train(problem, config, device, dir_path=None, log_path=None, max_steps=None, process_group=None, state_dict=None,
          val_samples=(1000, 1000, 100), verbose=True):
    """The reference's training loop.  `state_dict` (reference checkpoint format) or a freshly initialised model."""
    p = config['params']
    model_params = config['model_params']
    distribution = dict(config['distribution'])
    if state_dict is None:
        if problem == "cvrp":
            from .cvrp import CVRPModel as Model
        else:
            from .tsp import TSPModel as Model
        m = Model(**model_params)          # no local policy yet, as in the reference (CVRP/train.py:199)
        state_dict = m.state_dict()
    training = config.get('training', 'joint')
    if training not in ('joint', 'only_global'):
        raise NotImplementedError("`training: %s` is not implemented (joint and only_global are)" % training)
    tr = Trainer(problem, model_params, state_dict, device, lr=p['learning_rate'], weight_decay=1e-6,
                 scale_norm=p['scale_norm'], process_group=process_group)
    rank = torch.distributed.get_rank(process_group) if tr.world > 1 else 0
    file_logger = Logger(log_path, config) if (log_path and rank == 0) else None
    opts = np.array([15.740834, 7.909336, 14.294179]) if problem == "cvrp" else np.array([7.753418, 3.667576, 6.729566])   # CVRP/train.py:146, TSP/train.py:143
    gaps = np.array([1, 1, 1])
    n_steps = p['train_steps'] - p['start_steps'] + 1
    if max_steps is not None:
        n_steps = min(n_steps, max_steps)
    history = []
    for i in range(n_steps):
        # Enable joint training (CVRP/train.py:91-95)
        if i == p['T'] - p['start_steps'] and training == 'joint':
            if verbose and rank == 0:
                print("Enable joint training.")
            tr.add_local_policy()
        if p['mixed']:
            dis = np.random.choice(['uniform', 'cluster', 'mixed'], size=1, p=softmax(gaps))
            distribution['data_type'] = dis
        else:
            distribution['data_type'] = 'uniform'
        if problem == "cvrp":
            batch = gen.generate_vrp_data(p['train_batch_size'], p['problem_size'], distribution, device)
        else:
            batch = gen.generate_tsp_data(p['train_batch_size'], p['problem_size'], distribution, device)
        out = tr.step(batch, p['multiple_width'])
        history.append((float(out["loss"]), float(-out["reward"].max(1)[0].mean())))
        if verbose and rank == 0 and (i % 50 == 0):
            print("step %d  J %.5f  training length %.4f" % (i, history[-1][0], history[-1][1]), flush=True)
        if (i + 1) % p['log_step'] == 0:
            # rank 0 validates and checkpoints; the others wait at the broadcast (not inside the next step's gradient
            # all-reduce), and every rank gets the validation costs: the curriculum weights `gaps` stay identical
            val = torch.zeros(3, dtype=torch.float64, device=device)
            if rank == 0:
                val_info = validate(problem, tr, model_params, p['multiple_width'], device, p['mixed'], distribution, val_samples)
                val = torch.tensor(val_info, dtype=torch.float64, device=device)
                if file_logger is not None:
                    file_logger.log(val_info)
                if dir_path:
                    os.makedirs(dir_path, exist_ok=True)
                    ck = tr.checkpoint()
                    ck['step'] = i
                    torch.save(ck, dir_path + '/model_epoch_{}.pt'.format(int((i + 1) / p['log_step'])))
            if tr.world > 1:
                torch.distributed.broadcast(val, src=torch.distributed.get_global_rank(process_group, 0) if process_group is not None else 0,
                                            group=process_group)
            if p['mixed']:
                gaps = (val.cpu().numpy() - opts) / opts
    return tr, history


def main(problem):
    import datetime
    import yaml
    with open('config.yml', 'r', encoding='utf-8') as f:
        config = yaml.load(f.read(), Loader=yaml.FullLoader)
    if not config['use_cuda']:
        raise RuntimeError("elg_b200 has no CPU path: set use_cuda: True")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(config['cuda_device_num'])))
    device = "cuda:{}".format(local)
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    # seed_everything (CVRP/utils.py:121-128), offset per rank so that the ranks generate different training instances and
    # sampling streams; the PARAMETERS are rank 0's on every rank (Trainer.sync_weights broadcasts them at construction
    # and again when the local policy is added)
    seed = config['seed'] + int(os.environ.get("RANK", "0"))
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    ts = datetime.datetime.utcnow() + datetime.timedelta(hours=+8)
    ts_name = f'-ts{ts.month}-{ts.day}-{ts.hour}-{ts.minute}-{ts.second}'
    state_dict = None
    if config.get('load_checkpoint'):
        state_dict = torch.load(config['load_checkpoint'], map_location="cpu")['model_state_dict']
    train(problem, config, device, dir_path='weights/{}_{}_{}'.format(config['name'], ts_name, config['seed']),
          log_path='log/{}_{}'.format(config['name'], ts_name), state_dict=state_dict)
