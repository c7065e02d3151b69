"""REINFORCE training of ELG-POMO with the shared baseline: host side of the reference's training loop body
(CVRP/train.py:83-148, TSP/train.py:80-146) on top of the C ABI.

One `Trainer.step(batch)` = load_random_problems + pre_forward + rollout(sample) + J.backward() + Adam step.
Data-parallel training (one process per GPU): every rank runs the step on its own instances and the packed
gradient is summed with one NCCL all-reduce before the Adam step (grad_scale = 1 / world_size), so that equal
shards give the gradient of the global mean of J.  There is no CPU path.
"""
import random

import torch

from . import engine
from .dist import allreduce_mean_gradient


class Trainer:
    def __init__(self, problem, model_params, state_dict, device, lr=1e-4, weight_decay=1e-6, scale_norm=True,
                 betas=(0.9, 0.999), eps=1e-8, chunk_steps=128, process_group=None):
        self.problem = problem
        self.handle = engine.ModelHandle(problem, model_params, state_dict, device, attention="fp32")
        self.device = self.handle.device
        self.lr, self.weight_decay, self.scale_norm, self.betas, self.eps = lr, weight_decay, scale_norm, betas, eps
        self.chunk_steps = chunk_steps
        self.exp_avg = torch.zeros_like(self.handle.weights)
        self.exp_avg_sq = torch.zeros_like(self.handle.weights)
        self.grads = torch.zeros_like(self.handle.weights)
        self.step_count = 0
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        slots, _ = engine.weight_slots(problem, self.handle.desc)
        self._all_slots = slots
        self._local_keys = [k for k in slots if ".local_polic" in k]
        self._model_params = dict(model_params)
        # without a local policy (decoder.local False, CVRP/models.py:294-297) its keys are not part of the model
        self._slots = slots if self.handle.has_local else {k: o for k, o in slots.items() if k not in self._local_keys}
        self._shapes = {k: tuple(v.shape) for k, v in state_dict.items() if k in self._slots}
        self.global_step = 0          # never reset: seeds the sampling streams (the optimizer's step_count restarts with it)
        self.sync_weights()

    def sync_weights(self):
        """Data-parallel replicas must start from the same parameters: rank 0's packed weights are broadcast (every rank may
        have built its model from its own RNG, CVRP/train.py:199 under a per-rank seed) and the decoder tables re-folded.
        Collective: every rank of the group calls it (constructor, add_local_policy, load_checkpoint)."""
        if self.world > 1:
            src = torch.distributed.get_global_rank(self.pg, 0) if self.pg is not None else 0
            torch.distributed.broadcast(self.handle.weights, src=src, group=self.pg)
            engine.prepare_model(self.handle)

    @property
    def has_local(self):
        return self.handle.has_local

    def add_local_policy(self, local_state_dict=None):
        """The switch to joint training (CVRP/train.py:91-95): `model.decoder.add_local_policy(device)` with freshly
        initialised local-policy parameters (or the given ones) and a NEW optimizer (all Adam moments and the step
        count start again)."""
        from . import _lib
        from .params import LocalPolicy
        if self.handle.has_local:
            return
        if local_state_dict is None:
            lp = LocalPolicy(self.problem, dict(self._model_params, demand=self._model_params.get("demand", True)))
            pre = "decoder.local_policies.0." if self.problem == "cvrp" else "decoder.local_policy_0."
            local_state_dict = {pre + k: v for k, v in lp.state_dict().items()}
        self._slots = self._all_slots
        self._shapes.update({k: tuple(local_state_dict[k].shape) for k in self._local_keys})
        host = self.handle.weights.detach().cpu()
        for k in self._local_keys:
            v = local_state_dict[k].detach().to("cpu", torch.float32).reshape(-1)
            host[self._all_slots[k]:self._all_slots[k] + v.numel()] = v
        self.handle.weights.copy_(host)
        self.handle.desc.flags |= _lib.FLAG_ENSEMBLE
        self.handle.has_local = True
        engine.prepare_model(self.handle)
        self.sync_weights()           # the fresh local parameters were drawn per rank: rank 0's copy wins
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        self.step_count = 0

    # ---- one forward (sample rollout) + backward; returns the pieces so tests can look at them
    def forward_backward(self, data, M, start_nodes=None, seed=0):
        cv = self.problem == "cvrp"
        if cv:
            xy, dem = engine.load_problems("cvrp", data["loc"].to(self.device), data["depot"].to(self.device),
                                           data["demand"].to(self.device), 1)
        else:
            xy, dem = engine.load_problems("tsp", data.to(self.device), None, None, 1)
        N1 = int(xy.shape[1])
        if start_nodes is None:      # CVRP/CVRPModel.py:47, TSP/TSPModel.py:31
            start_nodes = random.sample(range(0, N1 - 1 if cv else M), M)
        batch, saved = engine.encode_train(self.handle, xy, dem)
        tours, reward, logp, n_steps = engine.rollout(batch, M, start_nodes, "sample", seed=seed)
        T = int(n_steps.max().item())
        grads, loss, ws = engine.reinforce_backward(batch, saved, M, tours, T, reward, logp, self.scale_norm,
                                                    self.chunk_steps, self.grads)
        return dict(batch=batch, tours=tours, reward=reward, logp=logp, T=T, grads=grads, loss=loss, ws=ws)

    def optimizer_step(self):
        scale = allreduce_mean_gradient(self.grads, self.pg)              # NCCL sum over NVLink (no-op on one GPU)
        self.step_count += 1
        engine.adam_step(self.handle.weights, self.grads, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                         self.betas[0], self.betas[1], self.eps, self.weight_decay, scale)
        engine.prepare_model(self.handle)

    def step(self, data, M, start_nodes=None, seed=None):
        if seed is None:       # a counter that survives the optimizer restart of add_local_policy: no sampling stream is replayed
            seed = (torch.initial_seed() * 1000003 + self.global_step) & (2 ** 63 - 1)
        out = self.forward_backward(data, M, start_nodes, seed)
        self.optimizer_step()
        self.global_step += 1
        return out

    # ---- reference-format state (CVRP/train.py:137-141)
    def state_dict(self):
        w = self.handle.weights.detach().cpu()
        return {k: w[off:off + int(torch.tensor(self._shapes[k]).prod())].reshape(self._shapes[k]).clone()
                for k, off in self._slots.items()}

    def unpack(self, flat):
        f = flat.detach().cpu()
        return {k: f[off:off + int(torch.tensor(self._shapes[k]).prod())].reshape(self._shapes[k]).clone()
                for k, off in self._slots.items()}

    def _pack(self, d, out):
        host = out.detach().cpu()
        for k, off in self._slots.items():
            v = d[k].detach().to("cpu", torch.float32).reshape(-1)
            host[off:off + v.numel()] = v
        out.copy_(host)

    def load_checkpoint(self, ck):
        """Resume from a dict written by `checkpoint()` (or by the reference: its torch Adam state is per-parameter and is
        accepted too when it carries 'state' / 'param_groups')."""
        self._pack(ck["model_state_dict"], self.handle.weights)
        engine.prepare_model(self.handle)
        opt = ck.get("optimizer_state_dict") or {}
        if "exp_avg" in opt:
            self._pack(opt["exp_avg"], self.exp_avg)
            self._pack(opt["exp_avg_sq"], self.exp_avg_sq)
            self.step_count = int(opt.get("step", ck.get("step", 0)))
        elif "state" in opt and opt["state"]:
            # torch.optim.Adam.state_dict(): parameters in model.parameters() order = state_dict order of the reference
            keys = [k for k in ck["model_state_dict"] if k in self._slots]
            st = opt["state"]
            self._pack({k: st[i]["exp_avg"] for i, k in enumerate(keys)}, self.exp_avg)
            self._pack({k: st[i]["exp_avg_sq"] for i, k in enumerate(keys)}, self.exp_avg_sq)
            self.step_count = int(st[0]["step"])
        else:
            self.exp_avg.zero_(); self.exp_avg_sq.zero_()
            self.step_count = 0
        self.global_step = max(int(ck.get("step", 0)), self.step_count)

    def checkpoint(self):
        """The reference's checkpoint dict (CVRP/train.py:137-141).  `optimizer_state_dict` has the layout of
        torch.optim.Adam.state_dict() -- per-parameter state in model.parameters() order (= state_dict order) plus
        param_groups -- so that the reference's `optimizer.load_state_dict` accepts it."""
        m, v = self.unpack(self.exp_avg), self.unpack(self.exp_avg_sq)
        keys = list(self._slots)
        state = {i: {"step": torch.tensor(float(self.step_count)), "exp_avg": m[k], "exp_avg_sq": v[k]} for i, k in enumerate(keys)}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(keys)))}
        return {"step": self.global_step, "model_state_dict": self.state_dict(),
                "optimizer_state_dict": {"state": state if self.step_count > 0 else {}, "param_groups": [group]}}
