"""Drop-in for the reference's TSPModel (TSP/TSPModel.py:10-64)."""
import random

import torch
import torch.nn as nn

from .. import engine
from ..params import Decoder, Encoder


class TSPModel(nn.Module):
    _elg_fused = True

    def __init__(self, **model_params):
        super().__init__()
        self.model_params = model_params
        self.encoder = Encoder("tsp", **model_params)
        self.decoder = Decoder("tsp", **model_params)
        self.encoded_nodes = None
        # shape: (batch, problem, EMBEDDING_DIM)
        self._handle, self._handle_key, self._batch = None, None, None
        self._first = None
        self._seed = 0

    def _get_handle(self, device):
        params = list(self.state_dict(keep_vars=True).items())
        key = (str(device),) + tuple((k, v.data_ptr(), v._version) for k, v in params)
        if self._handle is None or key != self._handle_key:
            self._handle = engine.ModelHandle("tsp", self.model_params, dict(params), device)
            self._handle_key = key
        return self._handle

    def _next_seed(self):
        self._seed += 1
        return (torch.initial_seed() * 1000003 + self._seed) & (2 ** 63 - 1)

    def pre_forward(self, reset_state):
        """Encoder + decoder caches (TSP/TSPModel.py:21-24); reads reset_state.problems (B, N, 2)."""
        xy = reset_state.problems
        self._batch = engine.encode(self._get_handle(xy.device), xy)
        self.encoded_nodes = self._batch.enc
        self._first = None

    def one_step_rollout(self, state, cur_dist, cur_theta, xy, eval_type):
        """One decode step (TSP/TSPModel.py:26-64); the first call picks the POMO start permutation."""
        batch_size, pomo_size = state.BATCH_IDX.size(0), state.BATCH_IDX.size(1)
        device = state.BATCH_IDX.device
        if state.current_node is None:
            selected = torch.tensor(random.sample(range(0, pomo_size), pomo_size), device=device)[None, :] \
                .expand(batch_size, pomo_size)
            prob = torch.ones(size=(batch_size, pomo_size), device=device)
            self._first = selected.contiguous()       # replaces decoder.set_q1: the kernel gathers Wq_first*enc[first]
            self._step = 0
        else:
            self._step += 1
            bits = getattr(state, "_mask_bits", None)
            if bits is None:
                bits = engine.pack_mask_bits(state.ninf_mask)
            selected, prob, logits = engine.decode_step(
                self._batch, pomo_size, state.current_node, bits, first=self._first, mode=eval_type,
                seed=self._next_seed(), step=self._step, want_logits=getattr(self, "_keep_logits", False))
            self._last_logits = logits
            if eval_type != 'sample':
                prob = None
        return selected, prob
