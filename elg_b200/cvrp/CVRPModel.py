"""Drop-in for the reference's CVRPModel (CVRP/CVRPModel.py:10-75).

Same constructor (`CVRPModel(**model_params)`), sub-module names (state_dict keys of the released
checkpoint), `pre_forward(reset_state)` and `one_step_rollout(state, cur_dist, cur_theta, xy,
norm_demand, eval_type)`; the computation runs in libelg_b200.so.
"""
import random

import torch
import torch.nn as nn

from .. import engine
from ..params import Decoder, Encoder


class CVRPModel(nn.Module):
    _elg_fused = True

    def __init__(self, **model_params):
        nn.Module.__init__(self)
        self.model_params = model_params
        self.encoder = Encoder("cvrp", **model_params)
        self.decoder = Decoder("cvrp", **model_params)
        self.encoded_nodes = None
        # shape: (batch, problem+1, embedding)
        self._handle, self._handle_key, self._batch = None, None, None
        self._seed = 0

    # ---- weights -> packed device buffer (re-packed only when a parameter changed) --------------
    def _get_handle(self, device):
        params = list(self.state_dict(keep_vars=True).items())
        key = (str(device),) + tuple((k, v.data_ptr(), v._version) for k, v in params)
        if self._handle is None or key != self._handle_key:
            self._handle = engine.ModelHandle("cvrp", self.model_params, dict(params), device)
            self._handle_key = key
        return self._handle

    def _next_seed(self):
        self._seed += 1
        return (torch.initial_seed() * 1000003 + self._seed) & (2 ** 63 - 1)

    def pre_forward(self, reset_state):
        """Encoder + decoder caches (CVRP/CVRPModel.py:21-34).  Reads depot_xy (B,1,2), node_xy (B,N,2),
        node_demand (B,N) from the reset state; `dist` is not needed."""
        src = getattr(reset_state, "_depot_node_xy", None)
        if src is not None:
            xy, dem = src, reset_state._depot_node_demand
        else:   # a reference Reset_State: rebuild the concatenated layout
            xy = torch.cat((reset_state.depot_xy, reset_state.node_xy), dim=1).contiguous()
            dem = torch.cat((torch.zeros_like(reset_state.node_demand[:, :1]), reset_state.node_demand), dim=1).contiguous()
        handle = self._get_handle(xy.device)
        self._batch = engine.encode(handle, xy, dem)
        self.encoded_nodes = self._batch.enc
        # shape: (batch, problem+1, embedding)

    def one_step_rollout(self, state, cur_dist, cur_theta, xy, norm_demand, eval_type):
        """One decode step (CVRP/CVRPModel.py:36-75).  The feature tensors are accepted for signature
        compatibility; the kernel recomputes what it needs from the node coordinates."""
        device = state.ninf_mask.device
        batch_size, multi_width = state.ninf_mask.shape[0], state.ninf_mask.shape[1]
        problem_size = state.ninf_mask.shape[2] - 1
        if state.selected_count == 0:      # first move: depot
            selected = torch.zeros(size=(batch_size, multi_width), dtype=torch.long, device=device)
            prob = torch.ones(size=(batch_size, multi_width), device=device)
        elif state.selected_count == 1:    # second move: POMO start nodes
            selected = torch.tensor(random.sample(range(0, problem_size), multi_width), device=device)[None, :] \
                .expand(batch_size, multi_width)
            prob = torch.ones(size=(batch_size, multi_width), device=device)
        else:
            bits = getattr(state, "_mask_bits", None)
            if bits is None:
                bits = engine.pack_mask_bits(state.ninf_mask)
            selected, prob, logits = engine.decode_step(
                self._batch, multi_width, state.current_node, bits, load=state.load, mode=eval_type,
                seed=self._next_seed(), step=state.selected_count, want_logits=getattr(self, "_keep_logits", False))
            self._last_logits = logits
            if eval_type != 'sample':
                prob = None
        return selected, prob
