"""Helpers for the training-path tests: reference training fixtures, oracle autograd gradients, comparisons."""
import json
import os

import numpy as np
import torch

from elg_b200.synth import state_dict_checksum, synthetic_state_dict
from oracle import elg_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN_CASES = ["train_cvrp_n20", "train_cvrp_n50", "train_cvrp_n100", "train_tsp_n20", "train_tsp_n50",
               "train_cvrp_n20_global", "train_tsp_n20_global"]
SAMPLE = 512


def sample_idx(numel):
    if numel <= SAMPLE:
        return np.arange(numel)
    return (np.arange(SAMPLE) * (numel // SAMPLE)).astype(np.int64)


class TrainGolden:
    """One recorded training step of the reference (oracle/gen_golden_train.py)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.kind, self.M = self.meta["problem"], self.meta["M"]

    def state_dict(self):
        sd = synthetic_state_dict(self.kind, seed=self.meta["wseed"], gain=self.meta["gain"])
        if not self.meta.get("local", True):      # reference decoder without add_local_policy
            sd = {k: v for k, v in sd.items() if ".local_polic" not in k}
        assert state_dict_checksum(sd) == self.meta["wsum"]
        return sd

    def model_params(self):
        """model_params as the oracle needs them: `ensemble` means "a local policy is present" there."""
        mp = dict(self.meta["model_params"])
        if not self.meta.get("local", True):
            mp["ensemble"] = False
        return mp

    def data(self):
        z = self.z
        if self.kind == "cvrp":
            return {"depot": torch.tensor(z["depot"]), "loc": torch.tensor(z["loc"]), "demand": torch.tensor(z["demand"])}
        return torch.tensor(z["problems"])

    def problem(self, dtype=torch.float32):
        d = self.data()
        if self.kind == "cvrp":
            return O.load_cvrp(d["depot"], d["loc"], d["demand"], 1, dtype)
        return O.load_tsp(d, 1, dtype)

    def tours(self):
        return torch.tensor(self.z["tours"].astype(np.int64))

    def reward(self):
        return torch.tensor(self.z["reward"])


def oracle_grads(kind, model_params, state_dict, prob, M, tours, reward, scale_norm=True, dtype=torch.float32, keep=False):
    """Autograd gradient of the oracle's teacher-forced REINFORCE objective.  keep=True also returns the gradients
    w.r.t. the encoded nodes and the decoder keys / values."""
    W = O.Weights(state_dict, kind, model_params, dtype).requires_grad_()
    enc = O.encode(W, prob)
    cache = O.decoder_cache(W, enc)
    if keep:
        enc.retain_grad(); cache.k.retain_grad(); cache.v.retain_grad()
    logp = O.teacher_forced_logp(W, prob, M, tours, cache=cache)
    J = (O.reinforce_coef(kind, reward.to(dtype), scale_norm) * logp).sum()
    J.backward()
    grads = {k: v.grad for k, v in W.sd.items()}
    extra = dict(enc=enc.grad, k=cache.k.grad, v=cache.v.grad) if keep else None
    return J.detach(), logp.detach(), grads, extra


def grad_errors(mine, ref, floor_frac=1e-3):
    """Per tensor: max |mine - ref| relative to max(rms(ref), floor_frac * largest rms).  Returns {key: err}."""
    rms = {k: float(ref[k].double().norm()) / np.sqrt(ref[k].numel()) for k in ref}
    floor = floor_frac * max(rms.values())
    return {k: float((mine[k].double().cpu() - ref[k].double()).abs().max()) / max(rms[k], floor) for k in ref}


def grad_errors_l2(mine, ref, floor_frac=1e-3):
    """Per tensor: ||mine - ref||_F / max(||ref||_F, floor_frac * largest tensor rms * sqrt(numel)).  Robust against a
    single ReLU unit whose pre-activation sits within rounding of zero (its whole gradient row flips on or off)."""
    rms = {k: float(ref[k].double().norm()) / np.sqrt(ref[k].numel()) for k in ref}
    floor = floor_frac * max(rms.values())
    return {k: float((mine[k].double().cpu() - ref[k].double()).norm()) / (max(rms[k], floor) * np.sqrt(ref[k].numel()))
            for k in ref}


def fixture_errors(g, mine, floor_frac=1e-3):
    rms = {k: float(g.z["g_norm/" + k]) / np.sqrt(mine[k].numel()) for k in g.meta["keys"]}
    floor = floor_frac * max(rms.values())
    out = {}
    for k in g.meta["keys"]:
        m = mine[k].reshape(-1).double().cpu().numpy()
        out[k] = float(np.abs(m[sample_idx(m.size)] - g.z["g_sample/" + k]).max()) / max(rms[k], floor)
    return out


def pad_tours(tours, t_max):
    B, M, T = tours.shape
    out = torch.zeros(B, M, t_max, dtype=torch.int16)
    out[:, :, :T] = tours.to(torch.int16)
    return out
