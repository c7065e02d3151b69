"""Deterministic synthetic checkpoints and instances.

The released checkpoints (`CVRP/weights/ELG.pt`, `TSP/weights/ELG.pt`) are not
shipped with the reference tree (see `.MISSING_LARGE_BLOBS`), so parity and
benchmarks use seeded random-init weights in the reference's checkpoint format
(`{'model_state_dict': ...}`, reference `CVRP/test.py:75-78`).  Every tensor is
drawn from its own generator keyed by (seed, parameter name) so the result does
not depend on module construction order and both the reference model and the
B200 model can `load_state_dict` the same dict.

Key names / shapes follow the reference modules:
  CVRP: `CVRP/models.py:199-209,232-247,276-297,7-25`
  TSP : `TSP/models.py:134-142,156-172,206-225,7-22`
"""
import hashlib
import math
from collections import OrderedDict

import torch

DEFAULT_MODEL_PARAMS = {
    "cvrp": dict(ensemble=True, distance_penalty=True, positional=True, xi=-1, local_size=[40],
                 ensemble_size=1, demand=True, euclidean=False, embedding_dim=128,
                 encoder_layer_num=6, head_num=8, qkv_dim=16, logit_clipping=50, ff_hidden_dim=512,
                 local_att_hidden_dim=32, local_att_head_num=4, local_att_qkv_dim=8),
    "tsp": dict(ensemble=True, distance_penalty=True, positional=True, ensemble_size=1, xi=-1,
                local_size=[30], euclidean=False, embedding_dim=128, encoder_layer_num=6,
                head_num=8, qkv_dim=16, logit_clipping=50, ff_hidden_dim=512,
                local_att_hidden_dim=32, local_att_head_num=4, local_att_qkv_dim=8),
}


def state_dict_spec(problem, model_params=None):
    """Ordered list of (key, shape, kind) for the reference state_dict of `problem`."""
    p = dict(DEFAULT_MODEL_PARAMS[problem])
    if model_params:
        p.update(model_params)
    E, F, L = p["embedding_dim"], p["ff_hidden_dim"], p["encoder_layer_num"]
    HD = p["head_num"] * p["qkv_dim"]
    e, hd = p["local_att_hidden_dim"], p["local_att_head_num"] * p["local_att_qkv_dim"]
    spec = []

    def lin(name, out_f, in_f, bias=True):
        spec.append((name + ".weight", (out_f, in_f), "linear"))
        if bias:
            spec.append((name + ".bias", (out_f,), "linear_bias:%d" % in_f))

    if problem == "cvrp":
        lin("encoder.embedding_depot", E, 2)
        lin("encoder.embedding_node", E, 3)
        n1, ff, n2 = "add_n_normalization_1", "feed_forward", "add_n_normalization_2"
    else:
        lin("encoder.embedding", E, 2)
        n1, ff, n2 = "addAndNormalization1", "feedForward", "addAndNormalization2"
    for i in range(L):
        pre = "encoder.layers.%d." % i
        lin(pre + "Wq", HD, E, False)
        lin(pre + "Wk", HD, E, False)
        lin(pre + "Wv", HD, E, False)
        lin(pre + "multi_head_combine", E, HD)
        spec.append((pre + n1 + ".norm.weight", (E,), "norm_w"))
        spec.append((pre + n1 + ".norm.bias", (E,), "norm_b"))
        lin(pre + ff + ".W1", F, E)
        lin(pre + ff + ".W2", E, F)
        spec.append((pre + n2 + ".norm.weight", (E,), "norm_w"))
        spec.append((pre + n2 + ".norm.bias", (E,), "norm_b"))
    if problem == "cvrp":
        lin("decoder.Wq_last", HD, E + 1, False)
    else:
        lin("decoder.Wq_first", HD, E, False)
        lin("decoder.Wq_last", HD, E, False)
    lin("decoder.Wk", HD, E, False)
    lin("decoder.Wv", HD, E, False)
    lin("decoder.multi_head_combine", E, HD)
    if p.get("ensemble", True):
        feat = 3 if (problem == "cvrp" and p.get("demand", True)) else 2
        pols = (["decoder.local_policies.%d." % i for i in range(p["ensemble_size"])]
                if problem == "cvrp" else ["decoder.local_policy_0."])
        for pre in pols:
            spec.append((pre + "cur_token_emb", (e,), "token"))
            lin(pre + "init_emb", e, feat)
            lin(pre + "Wq", hd, e, False)
            lin(pre + "Wk", hd, e, False)
            lin(pre + "Wv", hd, e, False)
            lin(pre + "multi_head_combine", e, hd)
    return spec


def _gen_for(seed, key):
    h = hashlib.sha256(("%d/%s" % (seed, key)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synthetic_state_dict(problem, seed=1234, gain=1.0, model_params=None):
    """Reference-format `model_state_dict` with PyTorch-default-like init.

    `gain` multiplies the decoder / local-policy output projections so that the
    pre-tanh scores leave the near-zero regime of a fresh init (trained
    checkpoints have peaked logits); `gain=1` is a plain fresh init.
    """
    sd = OrderedDict()
    for key, shape, kind in state_dict_spec(problem, model_params):
        g = _gen_for(seed, key)
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0
        if kind == "linear":
            t = u / math.sqrt(shape[1])
        elif kind.startswith("linear_bias"):
            t = u / math.sqrt(int(kind.split(":")[1]))
        elif kind == "norm_w":
            t = 1.0 + 0.1 * u
        elif kind == "norm_b":
            t = 0.1 * u
        elif kind == "token":
            t = u
        else:
            raise ValueError(kind)
        if gain != 1.0 and key.startswith("decoder.") and "multi_head_combine.weight" in key:
            t = t * gain
        sd[key] = t.contiguous()
    return sd


def state_dict_checksum(sd):
    """Order-independent checksum of a state_dict (detects RNG drift across torch builds)."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def synthetic_cvrp_batch(n, problem_size, seed=1234, capacity=None):
    """Uniform CVRP instances exactly as the reference's uniform generator
    (`CVRP/generate_data.py:10-13,75-89`): depot,loc ~ U[0,1), demand = randint(1,10)/capacity."""
    caps = {10: 20., 20: 30., 50: 40., 100: 50., 200: 80., 500: 100., 1000: 250., 5000: 500.}
    if capacity is None:
        capacity = caps.get(problem_size, 50.0)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    depot = torch.rand((n, 1, 2), generator=g)
    loc = torch.rand((n, problem_size, 2), generator=g)
    demand = torch.randint(1, 10, (n, problem_size), generator=g).float() / capacity
    return {"depot": depot, "loc": loc, "demand": demand}


def synthetic_tsp_batch(n, problem_size, seed=1234):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand((n, problem_size, 2), generator=g)
