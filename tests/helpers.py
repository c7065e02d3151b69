"""Shared helpers for the parity tests: golden loading and oracle-side reconstruction."""
import json
import os

import numpy as np
import torch

from elg_b200.synth import DEFAULT_MODEL_PARAMS, state_dict_checksum, synthetic_state_dict
from oracle import elg_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CVRP_CASES = ["cvrp_n20", "cvrp_n20_sharp", "cvrp_n50", "cvrp_n100", "cvrp_n100_sharp", "cvrp_n20_noaug", "cvrp_lib", "cvrp_n200", "cvrp_x101", "cvrp_x200"]
TSP_CASES = ["tsp_n20", "tsp_n20_sharp", "tsp_n50", "tsp_n100", "tsp_n30_m10", "tsp_lib", "tsp_n150"]
ALL_CASES = CVRP_CASES + TSP_CASES


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.kind = self.meta["problem"]
        self.M, self.aug, self.T = self.meta["M"], self.meta["aug"], self.meta["T"]
        self.rows_b = [int(b) for b in self.z["rows_b"]]
        self.steps = [int(t) for t in self.z["step_ids"]]

    def state_dict(self):
        sd = synthetic_state_dict(self.kind, seed=self.meta["wseed"], gain=self.meta["gain"])
        assert state_dict_checksum(sd) == self.meta["wsum"], "synthetic weights drifted from the fixture"
        return sd

    def model_params(self):
        return dict(self.meta["model_params"])

    def perm(self):
        return torch.tensor(self.z["perm"].astype(np.int64))

    def tours(self):
        return torch.tensor(self.z["tours"].astype(np.int64))

    def reward(self):
        return torch.tensor(self.z["reward"])

    def oracle_problem(self, dtype=torch.float32):
        z = self.z
        if self.kind == "cvrp":
            if self.meta.get("lib"):
                return O.load_vrplib(z["lib_node_coord"], z["lib_demand"], float(z["lib_capacity"]), self.aug, dtype)
            return O.load_cvrp(torch.tensor(z["depot"]), torch.tensor(z["loc"]), torch.tensor(z["demand"]), self.aug, dtype)
        if self.meta.get("lib"):
            return O.load_tsplib(z["lib_node_coord"], self.aug, dtype)
        return O.load_tsp(torch.tensor(z["problems"]), self.aug, dtype)

    def step(self, t):
        """Recorded pre-decode state + reference logits for the recorded aug-instances at step t."""
        z = self.z
        logits = torch.tensor(z["s%d_logits" % t])
        N1 = logits.shape[-1]
        masked = torch.tensor(np.unpackbits(z["s%d_maskbits" % t], axis=-1, bitorder="little")[..., :N1].astype(bool))
        out = dict(cur=torch.tensor(z["s%d_cur" % t].astype(np.int64)), masked=masked, logits=logits,
                   selected=torch.tensor(z["s%d_selected" % t].astype(np.int64)))
        if self.kind == "cvrp":
            out["load"] = torch.tensor(z["s%d_load" % t])
            out["finished"] = torch.tensor(z["s%d_finished" % t])
        return out


def sub_problem(prob, rows):
    """Restrict an oracle Problem to some aug-instances."""
    return O.Problem(prob.kind, prob.xy[rows], None if prob.demand is None else prob.demand[rows], prob.dist[rows],
                     None if prob.unscaled_xy is None else (prob.unscaled_xy if prob.unscaled_xy.shape[0] == 1 else prob.unscaled_xy[rows]),
                     prob.aug)


def top2_margin(logits):
    """Gap between the best and second-best finite logit per row."""
    l = logits.clone().float()
    l[torch.isinf(l)] = -1e30
    v = l.topk(2, dim=-1)[0]
    return v[..., 0] - v[..., 1]


def compare_tours(t_a, t_b):
    """Fraction of (b, m) rows whose tours are identical (zero-padded to a common length)."""
    T = max(t_a.shape[2], t_b.shape[2])
    pa = torch.zeros(*t_a.shape[:2], T, dtype=torch.int64); pa[:, :, :t_a.shape[2]] = t_a
    pb = torch.zeros(*t_b.shape[:2], T, dtype=torch.int64); pb[:, :, :t_b.shape[2]] = t_b
    same = (pa == pb).all(dim=2)
    return float(same.float().mean()), same


def tie_rows(prob, kind, cur, masked, k):
    """Rows whose k+1 nearest valid nodes contain two exactly equal distances: torch.topk orders such ties
    in an implementation-defined way (SURVEY 'hard parts'), so rank-dependent logits are not comparable."""
    N1 = masked.shape[2]
    row = prob.dist.gather(1, cur[:, :, None].expand(-1, -1, N1))
    excl = masked.clone()
    if kind == "cvrp":
        excl[:, :, 0] = True
    d = row.masked_fill(excl, float("inf")).sort(dim=-1)[0][:, :, :k + 1]
    return ((d[:, :, 1:] == d[:, :, :-1]) & ~torch.isinf(d[:, :, 1:])).any(-1)
