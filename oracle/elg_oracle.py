"""CPU oracle: a restatement of the ELG rollout hot path (gaocrr/ELG) in batched torch-CPU ops.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module, and only as the
checker / CPU baseline.  The product (elg_b200/) never imports it and has no CPU
fallback.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against tests/golden/*.npz, which oracle/gen_golden.py produced by running the
unmodified reference in the build container (the reference's own tree has no
golden vectors for this path, and its released checkpoints are absent, so the
pins are reference outputs on seeded synthetic weights).

The arithmetic follows the reference op for op where rounding matters (fp32
`F.linear` on the concatenated [embedding, load] query input, the precomputed
pairwise `norm(p=2)` distance matrix, `atan2` of relative coordinates, division
of demand by load, the two softmaxes, `50*tanh`), but it is written against
boolean masks and per-row semantics rather than the
reference's in-place +-inf tensors.  `dtype=torch.float64` gives a high-precision
"truth" used to measure how far either fp32 implementation is from exact.

Reference map (file:line under /root/reference):
  augment8            CVRP/utils.py:69-87, TSP/utils.py:89-107
  load_cvrp/load_tsp  CVRP/CVRPEnv.py:125-150, TSP/TSPEnv.py:53-67
  load_vrplib/tsplib  CVRP/CVRPEnv.py:84-123, TSP/test_tsplib.py:126-137 + TSP/TSPEnv.py:69-85
  encode              CVRP/models.py:199-269,506-562, TSP/models.py:134-194,387-423
  decoder_cache       CVRP/models.py:300-308, TSP/models.py:227-242
  decode_logits       CVRP/models.py:322-423 + 51-175, TSP/models.py:244-303 + 48-110
  cur_feature         CVRP/CVRPEnv.py:291-318, TSP/TSPEnv.py:135-156
  select              CVRP/CVRPModel.py:36-75, TSP/TSPModel.py:26-64
  cvrp_env_step       CVRP/CVRPEnv.py:190-249
  tsp_env_step        TSP/TSPEnv.py:108-133
  tour_length         CVRP/CVRPEnv.py:251-288, TSP/TSPEnv.py:158-184
  rollout             CVRP/utils.py:7-29, TSP/utils.py:7-26
  reinforce_loss      CVRP/train.py:112-121, TSP/train.py:107-119 (teacher-forced on recorded tours; torch
                      autograd through this module is the gradient oracle of the training path)
  adam_step           torch.optim.Adam(lr, weight_decay=1e-6) as constructed at CVRP/train.py:87
"""
import math
import random
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


# --------------------------------------------------------------------------- problem loading

def augment8(xy):
    """(b, n, 2) -> (8b, n, 2); block a holds transform a of every instance (row = a*b + i)."""
    x, y = xy[..., 0:1], xy[..., 1:2]
    variants = [(x, y), (1 - x, y), (x, 1 - y), (1 - x, 1 - y), (y, x), (1 - y, x), (y, 1 - x), (1 - y, 1 - x)]
    return torch.cat([torch.cat(v, dim=-1) for v in variants], dim=0)


def pairwise_dist(xy):
    return (xy[:, :, None, :] - xy[:, None, :, :]).norm(p=2, dim=-1)


@dataclass
class Problem:
    kind: str                      # 'cvrp' | 'tsp'
    xy: torch.Tensor               # (B, N1, 2)  node 0 is the depot for cvrp
    demand: torch.Tensor = None    # (B, N1)     demand[:, 0] == 0 (cvrp)
    dist: torch.Tensor = None      # (B, N1, N1)
    unscaled_xy: torch.Tensor = None
    aug: int = 1


def load_cvrp(depot, loc, demand, aug=1, dtype=torch.float32):
    depot = depot.to(dtype)
    if depot.dim() == 2:
        depot = depot[:, None, :]
    loc, demand = loc.to(dtype), demand.to(dtype)
    if aug == 8:
        depot, loc, demand = augment8(depot), augment8(loc), demand.repeat(8, 1)
    elif aug != 1:
        raise NotImplementedError
    xy = torch.cat((depot, loc), dim=1)
    dem = torch.cat((torch.zeros(xy.shape[0], 1, dtype=dtype), demand), dim=1)
    return Problem("cvrp", xy, dem, pairwise_dist(xy), None, aug)


def load_vrplib(node_coord, demand, capacity, aug=1, dtype=torch.float32):
    """Per-axis min-max scaling of library coordinates; node 0 is the depot."""
    coord = torch.as_tensor(node_coord, dtype=torch.float32)[None].to(dtype)
    dem = (torch.as_tensor(demand, dtype=torch.float32)[None] / capacity).to(dtype)
    lo, hi = coord.min(dim=1, keepdim=True)[0], coord.max(dim=1, keepdim=True)[0]
    xy = (coord - lo) / (hi - lo)
    unscaled = coord
    if aug == 8:
        xy, unscaled, dem = augment8(xy), augment8(unscaled), dem.repeat(8, 1)
    elif aug != 1:
        raise NotImplementedError
    return Problem("cvrp", xy, dem, pairwise_dist(xy), unscaled, aug)


def load_tsp(problems, aug=1, dtype=torch.float32, unscaled=None):
    xy = problems.to(dtype)
    if aug == 8:
        xy = augment8(xy)
    elif aug != 1:
        raise NotImplementedError
    return Problem("tsp", xy, None, pairwise_dist(xy), unscaled, aug)


def load_tsplib(node_coord, aug=1, dtype=torch.float32):
    """Global (both axes together) min-max scaling; the unscaled copy is NOT augmented."""
    import numpy as np
    c = np.asarray(node_coord)
    pts = (c - np.min(c)) / (np.max(c) - np.min(c))
    unscaled = torch.tensor(c, dtype=torch.float)[None].to(dtype)
    return load_tsp(torch.tensor(pts, dtype=torch.float)[None], aug, dtype, unscaled)


# --------------------------------------------------------------------------- weights

class Weights:
    """Reference-format state_dict cast to one dtype, with problem-specific key names resolved."""

    def __init__(self, state_dict, kind, model_params, dtype=torch.float32):
        self.kind, self.p, self.dtype = kind, model_params, dtype
        self.sd = {k: v.detach().to("cpu", dtype) for k, v in state_dict.items()}
        if kind == "cvrp":
            self.norm1, self.ff, self.norm2 = "add_n_normalization_1", "feed_forward", "add_n_normalization_2"
            self.local = "decoder.local_policies.0."
        else:
            self.norm1, self.ff, self.norm2 = "addAndNormalization1", "feedForward", "addAndNormalization2"
            self.local = "decoder.local_policy_0."

    def __getitem__(self, k):
        return self.sd[k]

    def requires_grad_(self, flag=True):
        for v in self.sd.values():
            v.requires_grad_(flag)
        return self


def _heads(t, h):
    b, n, _ = t.shape
    return t.reshape(b, n, h, -1).transpose(1, 2)


def _instance_norm(x, w, b, eps=1e-5):
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * w + b


# --------------------------------------------------------------------------- encoder

def encode(W, prob):
    p, H = W.p, W.p["head_num"]
    if W.kind == "cvrp":
        dep = F.linear(prob.xy[:, :1], W["encoder.embedding_depot.weight"], W["encoder.embedding_depot.bias"])
        feat = torch.cat((prob.xy[:, 1:], prob.demand[:, 1:, None]), dim=2)
        nod = F.linear(feat, W["encoder.embedding_node.weight"], W["encoder.embedding_node.bias"])
        x = torch.cat((dep, nod), dim=1)
    else:
        x = F.linear(prob.xy, W["encoder.embedding.weight"], W["encoder.embedding.bias"])
    for i in range(p["encoder_layer_num"]):
        pre = "encoder.layers.%d." % i
        q = _heads(F.linear(x, W[pre + "Wq.weight"]), H)
        k = _heads(F.linear(x, W[pre + "Wk.weight"]), H)
        v = _heads(F.linear(x, W[pre + "Wv.weight"]), H)
        att = torch.softmax(q @ k.transpose(2, 3) / math.sqrt(p["qkv_dim"]), dim=3) @ v
        att = att.transpose(1, 2).reshape(x.shape[0], x.shape[1], -1)
        mh = F.linear(att, W[pre + "multi_head_combine.weight"], W[pre + "multi_head_combine.bias"])
        x1 = _instance_norm(x + mh, W[pre + W.norm1 + ".norm.weight"], W[pre + W.norm1 + ".norm.bias"])
        hid = F.relu(F.linear(x1, W[pre + W.ff + ".W1.weight"], W[pre + W.ff + ".W1.bias"]))
        ff = F.linear(hid, W[pre + W.ff + ".W2.weight"], W[pre + W.ff + ".W2.bias"])
        x = _instance_norm(x1 + ff, W[pre + W.norm2 + ".norm.weight"], W[pre + W.norm2 + ".norm.bias"])
    return x


@dataclass
class DecoderCache:
    enc: torch.Tensor     # (B, N1, E)
    k: torch.Tensor       # (B, H, N1, D)
    v: torch.Tensor
    q_first: torch.Tensor = None   # tsp: (B, H, M, D)


def decoder_cache(W, enc):
    H = W.p["head_num"]
    return DecoderCache(enc, _heads(F.linear(enc, W["decoder.Wk.weight"]), H),
                        _heads(F.linear(enc, W["decoder.Wv.weight"]), H))


def _gather_nodes(enc, idx):
    return enc.gather(1, idx[:, :, None].expand(-1, -1, enc.shape[2]))


def set_first(W, cache, first):
    """TSP: cache the first-node query (TSP/models.py:237-242)."""
    cache.q_first = _heads(F.linear(_gather_nodes(cache.enc, first), W["decoder.Wq_first.weight"]), W.p["head_num"])


# --------------------------------------------------------------------------- decode step

def _position_table(n_pos, emb, dtype):
    half = emb // 2
    inc = math.log(10000.0) / max(half - 1, 1)
    inv = torch.exp(torch.arange(half, dtype=torch.float32) * -inc).to(dtype)
    ang = torch.arange(n_pos, dtype=torch.float32).to(dtype)[:, None] * inv[None, :]
    return torch.cat((torch.sin(ang), torch.cos(ang)), dim=1)


def cur_feature(prob, cur, load=None):
    """The per-step feature tensors the reference's environments hand to the model: CVRPEnv.get_cur_feature
    (CVRP/CVRPEnv.py:291-318) -> (cur_dist, cur_theta, relative_xy, norm_demand); TSPEnv.get_local_feature
    (TSP/TSPEnv.py:135-156) -> the first three.  cur_dist is a row of the precomputed distance matrix, theta =
    atan2(rel_y, rel_x), norm_demand = demand / load for EVERY node (0/0 = nan at the depot and +-inf at customers
    when the load is exactly 0 or slightly negative -- SURVEY A.6; the model only ever reads valid customers)."""
    B, N1, _ = prob.xy.shape
    cur_dist = prob.dist.gather(1, cur[:, :, None].expand(-1, -1, N1))
    rel = prob.xy[:, None, :, :] - prob.xy.gather(1, cur[:, :, None].expand(-1, -1, 2))[:, :, None, :]
    theta = torch.atan2(rel[..., 1], rel[..., 0])
    if prob.kind == "cvrp":
        return cur_dist, theta, rel, prob.demand[:, None, :] / load[:, :, None]
    return cur_dist, theta, rel


def _neighbourhood(W, prob, cur, masked, load, feats=None):
    """k nearest valid nodes per row, ascending distance; returns dict of (B, M, L) tensors.

    feats = (cur_dist, cur_theta, norm_demand | None): take the per-node features from the caller (what the reference's
    local_policy_att.forward receives from the environment, CVRP/models.py:51-60) instead of recomputing them.

    valid = not masked and (cvrp) not the depot.  Slots past a row's valid count are
    flagged `pad` and carry zero features, as in the reference (inf -> 0 padding).
    cvrp sequences get the depot prepended at position 0 with features (0, 0, 0).
    """
    B, M, N1 = masked.shape
    k = W.p["local_size"][0]
    row_dist = prob.dist.gather(1, cur[:, :, None].expand(-1, -1, N1)) if feats is None else feats[0]
    excl = masked.clone()
    if W.kind == "cvrp":
        excl[:, :, 0] = True
    n_valid = (~excl).sum(-1)
    kb = int(min(k, int(n_valid.max())))
    keyed = row_dist.masked_fill(excl, float("inf"))
    out = {"row_dist": row_dist}
    if kb > 0:
        if W.kind == "cvrp":
            d, idx = keyed[:, :, 1:].topk(kb, dim=-1, largest=False)
            idx = idx + 1
        else:
            d, idx = keyed.topk(kb, dim=-1, largest=False)
        pad = torch.isinf(d)
        d = d.masked_fill(pad, 0.0)
        dmax = d.max(-1, keepdim=True)[0]
    else:
        d = row_dist.new_zeros(B, M, 0)
        idx = cur.new_zeros(B, M, 0)
        pad = torch.zeros(B, M, 0, dtype=torch.bool)
        dmax = row_dist.new_zeros(B, M, 1)
    if feats is None:
        rel = prob.xy[:, None, :, :] - prob.xy.gather(1, cur[:, :, None].expand(-1, -1, 2))[:, :, None, :]
        theta_all = torch.atan2(rel[..., 1], rel[..., 0])
    else:
        theta_all = feats[1]
    theta = theta_all.gather(2, idx).masked_fill(pad, 0.0)
    out.update(d=d, idx=idx, pad=pad, dmax=dmax, theta=theta)
    if W.kind == "cvrp":
        nd_all = prob.demand[:, None, :] / load[:, :, None] if feats is None else feats[2]
        nd = nd_all.gather(2, idx).masked_fill(pad, 0.0)
        out["norm_demand"] = nd
    return out


def _local_scores(W, nb, masked):
    """Local attention policy scores for the neighbourhood sequence -> (B, M, L[+1])."""
    p = W.p
    e, h = p["local_att_hidden_dim"], p["local_att_head_num"]
    pre = W.local
    d, theta, pad = nb["d"], nb["theta"], nb["pad"]
    dn = d / (nb["dmax"] + 1e-6)
    if W.kind == "cvrp":
        dn = torch.where(nb["dmax"] != 0, dn, d)        # rows with dmax == 0 stay un-normalised
        feat = torch.stack((dn, theta, nb["norm_demand"]), dim=-1)
        feat = torch.cat((feat.new_zeros(feat.shape[0], feat.shape[1], 1, 3), feat), dim=2)
        seq_masked = torch.cat((masked[:, :, :1], pad | masked.gather(2, nb["idx"])), dim=2)
    else:
        feat = torch.stack((dn, theta), dim=-1)
        seq_masked = pad | masked.gather(2, nb["idx"])
    L = feat.shape[2]
    ik = F.linear(feat, W[pre + "init_emb.weight"], W[pre + "init_emb.bias"])
    if p["positional"]:
        ik = ik + _position_table(L, e, ik.dtype)[None, None]
    q = F.linear(W[pre + "cur_token_emb"], W[pre + "Wq.weight"]).reshape(h, -1)            # (h, dk)
    kk = F.linear(ik, W[pre + "Wk.weight"]).reshape(*ik.shape[:3], h, -1)                   # (B, M, L, h, dk)
    vv = F.linear(ik, W[pre + "Wv.weight"]).reshape(*ik.shape[:3], h, -1)
    s = torch.einsum("hd,bmlhd->bmhl", q, kk) / math.sqrt(p["local_att_qkv_dim"])
    s = s.masked_fill(seq_masked[:, :, None, :], NEG_INF)
    w = torch.softmax(s, dim=-1)
    o = torch.einsum("bmhl,bmlhd->bmhd", w, vv).reshape(ik.shape[0], ik.shape[1], -1)
    mh = F.linear(o, W[pre + "multi_head_combine.weight"], W[pre + "multi_head_combine.bias"])
    return torch.einsum("bme,bmle->bml", mh, ik) / math.sqrt(e)


def decode_logits(W, prob, cache, cur, masked, load=None, feats=None):
    """Masked logits (the tensor the reference feeds to its final softmax), shape (B, M, N1).
    feats: optional environment features for the local policy / distance penalty, see _neighbourhood.

    cur (B, M) int64; masked (B, M, N1) bool (True = -inf in the reference's ninf_mask);
    load (B, M) for cvrp.
    """
    p = W.p
    H, E = p["head_num"], p["embedding_dim"]
    B, M, N1 = masked.shape
    last = _gather_nodes(cache.enc, cur)
    if W.kind == "cvrp":
        q = _heads(F.linear(torch.cat((last, load[:, :, None]), dim=2), W["decoder.Wq_last.weight"]), H)
    else:
        q = cache.q_first + _heads(F.linear(last, W["decoder.Wq_last.weight"]), H)
    s = q @ cache.k.transpose(2, 3) / math.sqrt(p["qkv_dim"])
    s = s.masked_fill(masked[:, None, :, :], NEG_INF)
    o = (torch.softmax(s, dim=3) @ cache.v).transpose(1, 2).reshape(B, M, -1)
    mh = F.linear(o, W["decoder.multi_head_combine.weight"], W["decoder.multi_head_combine.bias"])
    score = mh @ cache.enc.transpose(1, 2) / math.sqrt(E)

    nb = _neighbourhood(W, prob, cur, masked, load, feats)
    if p["distance_penalty"]:
        pen = torch.full_like(score, float(p["xi"]))
        if W.kind == "cvrp":
            pen[:, :, 0] = -0.0
            val = torch.where(nb["dmax"] != 0, nb["d"] / nb["dmax"], nb["d"])
        else:
            val = nb["d"] / (nb["dmax"] + 1e-6)
        pen.scatter_(2, nb["idx"], -val)
        score = score + pen
    if p["ensemble"]:
        loc = _local_scores(W, nb, masked)
        idx = nb["idx"]
        if W.kind == "cvrp":
            idx = torch.cat((idx.new_zeros(B, M, 1), idx), dim=2)
        score = score + torch.zeros_like(score).scatter_(2, idx, loc)
    return (p["logit_clipping"] * torch.tanh(score)).masked_fill(masked, NEG_INF)


# --------------------------------------------------------------------------- environment

@dataclass
class CvrpState:
    cur: torch.Tensor = None         # (B, M) int64
    load: torch.Tensor = None        # (B, M)
    visited: torch.Tensor = None     # (B, M, N1) bool; depot bit = "at the depot"
    masked: torch.Tensor = None      # (B, M, N1) bool
    finished: torch.Tensor = None    # (B, M) bool
    count: int = 0


def cvrp_reset(prob, M):
    B, N1 = prob.demand.shape
    return CvrpState(None, torch.ones(B, M, dtype=prob.xy.dtype), torch.zeros(B, M, N1, dtype=torch.bool),
                     torch.zeros(B, M, N1, dtype=torch.bool), torch.zeros(B, M, dtype=torch.bool), 0)


def cvrp_env_step(prob, st, sel):
    at_depot = sel == 0
    st.count += 1
    st.cur = sel
    st.load = st.load - prob.demand.gather(1, sel)          # sequential fp32 recurrence
    st.load = torch.where(at_depot, torch.ones_like(st.load), st.load)
    st.visited.scatter_(2, sel[:, :, None], True)
    st.visited[:, :, 0] = at_depot
    too_big = st.load[:, :, None] + 1e-6 < prob.demand[:, None, :]
    st.masked = st.visited | too_big
    st.finished = st.finished | st.visited.all(dim=2)
    st.masked[:, :, 0] &= ~st.finished
    return bool(st.finished.all())


@dataclass
class TspState:
    cur: torch.Tensor = None
    masked: torch.Tensor = None
    count: int = 0


def tsp_reset(prob, M):
    B, N, _ = prob.xy.shape
    return TspState(None, torch.zeros(B, M, N, dtype=torch.bool), 0)


def tsp_env_step(prob, st, sel):
    st.count += 1
    st.cur = sel
    st.masked.scatter_(2, sel[:, :, None], True)
    return st.count == prob.xy.shape[1]


def tour_length(xy, tours, rounding=False):
    """Closed-tour length (B, M): sum_t |xy[tour_t] - xy[tour_{t+1 mod T}]|, optional per-edge rint."""
    B, M, T = tours.shape
    pts = xy[:, None, :, :].expand(B, M, -1, 2).gather(2, tours[:, :, :, None].expand(-1, -1, -1, 2))
    seg = ((pts - pts.roll(dims=2, shifts=-1)) ** 2).sum(3).sqrt()
    if rounding:
        seg = torch.round(seg)
    return seg.sum(2)


# --------------------------------------------------------------------------- rollout

def start_permutation(kind, N, M, seed=None):
    """POMO start nodes from Python's `random.sample` (CVRP/CVRPModel.py:47, TSP/TSPModel.py:31).
    cvrp samples range(0, N) (includes the depot, never node N); tsp samples range(0, M)."""
    if seed is not None:
        random.seed(seed)
    return torch.tensor(random.sample(range(0, N if kind == "cvrp" else M), M), dtype=torch.int64)


def select(logits, mode, generator=None):
    probs = torch.softmax(logits, dim=2)
    if mode == "greedy":
        return probs.argmax(dim=2), None
    B, M, N1 = probs.shape
    sel = probs.reshape(B * M, N1).multinomial(1, generator=generator).reshape(B, M)
    return sel, probs.gather(2, sel[:, :, None]).squeeze(2)


def rollout(W, prob, M, perm, mode="greedy", generator=None, cache=None, hook=None, max_steps=None):
    """Full construction rollout. Returns (tours (B, M, T) int64, probs (B, T, M) | None, reward (B, M))."""
    if cache is None:
        cache = decoder_cache(W, encode(W, prob))
    B = prob.xy.shape[0]
    acts, plist = [], []
    ones = torch.ones(B, M, dtype=prob.xy.dtype)
    if W.kind == "cvrp":
        st = cvrp_reset(prob, M)
        done = False
        while not done:
            if st.count == 0:
                sel, pr = torch.zeros(B, M, dtype=torch.int64), ones
            elif st.count == 1:
                sel, pr = perm[None, :].expand(B, M).clone(), ones
            else:
                logits = decode_logits(W, prob, cache, st.cur, st.masked, st.load)
                if hook is not None:
                    hook(st, logits)
                sel, pr = select(logits, mode, generator)
                if pr is not None and not (pr != 0).all():
                    pr = pr + 1e-6
            done = cvrp_env_step(prob, st, sel)
            acts.append(sel)
            plist.append(pr)
            if max_steps is not None and st.count >= max_steps:
                break
    else:
        st = tsp_reset(prob, M)
        done = False
        while not done:
            if st.count == 0:
                sel, pr = perm[None, :].expand(B, M).clone(), ones
                set_first(W, cache, sel)
            else:
                logits = decode_logits(W, prob, cache, st.cur, st.masked)
                if hook is not None:
                    hook(st, logits)
                while True:
                    sel, pr = select(logits, mode, generator)
                    if pr is None or (pr != 0).all():
                        break
            done = tsp_env_step(prob, st, sel)
            acts.append(sel)
            plist.append(pr)
            if max_steps is not None and st.count >= max_steps:
                break
    tours = torch.stack(acts, dim=2)
    if prob.unscaled_xy is not None:
        reward = -tour_length(prob.unscaled_xy.expand(B, -1, -1) if prob.kind == "tsp" else prob.unscaled_xy,
                              tours, rounding=True)
    else:
        reward = -tour_length(prob.xy, tours)
    probs = None if mode == "greedy" else torch.stack(plist, dim=1)
    return tours, probs, reward


# --------------------------------------------------------------------------- training objective

def teacher_forced_logp(W, prob, M, tours, cache=None, hook=None):
    """Sum over steps of log p(recorded action) per row, (B, M), differentiable w.r.t. the tensors in W.

    Replays `tours` (B, M, T) through the environment; forced steps (cvrp: depot + POMO start, tsp: POMO start)
    carry probability 1 (CVRP/CVRPModel.py:43-50, TSP/TSPModel.py:30-37), finished rows select the depot with
    probability 1.  This is `probs.log().sum(dim=1)` of CVRP/train.py:115 for the same action sequence."""
    if cache is None:
        cache = decoder_cache(W, encode(W, prob))
    B, _, T = tours.shape
    logp = torch.zeros(B, M, dtype=prob.xy.dtype)
    if W.kind == "cvrp":
        st = cvrp_reset(prob, M)
        for t in range(T):
            sel = tours[:, :, t]
            if t >= 2:
                logits = decode_logits(W, prob, cache, st.cur, st.masked, st.load)
                if hook is not None:
                    hook(t, st, logits)
                lp = torch.log_softmax(logits, dim=2).gather(2, sel[:, :, None]).squeeze(2)
                logp = logp + torch.where(st.finished, torch.zeros_like(lp), lp)
            cvrp_env_step(prob, st, sel)
    else:
        st = tsp_reset(prob, M)
        for t in range(T):
            sel = tours[:, :, t]
            if t == 0:
                set_first(W, cache, sel)
            else:
                logits = decode_logits(W, prob, cache, st.cur, st.masked.clone())   # the env updates its mask in place
                if hook is not None:
                    hook(t, st, logits)
                logp = logp + torch.log_softmax(logits, dim=2).gather(2, sel[:, :, None]).squeeze(2)
            tsp_env_step(prob, st, sel)
    return logp


def reinforce_coef(kind, reward, scale_norm=True):
    """dJ/dlogp per row for J = mean(-(r - mean_m r) * logp [/ max_m(r - mean_m r)])  (CVRP/train.py:113-121;
    TSP/train.py:114-117 skips the scaling unless every instance has a non-zero maximum advantage)."""
    adv = reward - reward.mean(dim=1, keepdim=True)
    coef = -adv
    if scale_norm:
        fac = adv.max(dim=1, keepdim=True)[0]
        if kind == "cvrp" or bool((fac != 0).all()):
            coef = coef / fac
    return coef / reward.numel()


def reinforce_loss(W, prob, M, tours, reward, scale_norm=True, hook=None):
    """J of CVRP/train.py:112-121 for recorded tours and rewards (rewards carry no gradient)."""
    logp = teacher_forced_logp(W, prob, M, tours, hook=hook)
    return (reinforce_coef(W.kind, reward, scale_norm) * logp).sum(), logp


def adam_step(p, g, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-6):
    """One torch.optim.Adam update (L2 weight decay added to the gradient), step counted from 1."""
    g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    denom = v.sqrt() / math.sqrt(1 - beta2 ** step) + eps
    return p - (lr / (1 - beta1 ** step)) * m / denom, m, v


def best_of(reward, aug, n):
    """(aug*n, M) rewards -> (no-aug cost (n,), aug cost (n,))  (CVRP/test.py:31-41)."""
    r = reward.reshape(aug, n, -1).max(dim=2)[0]
    return -r[0], -r.max(dim=0)[0]


def check_feasible_cvrp(tours, demand):
    """Known-answer invariant (CVRP/utils.py:90-119): every customer exactly once, capacity <= 1+1e-4."""
    B, M, T = tours.shape
    N = demand.shape[1] - 1
    srt = tours.sort(dim=2)[0]
    assert (srt[:, :, -N:] == torch.arange(1, N + 1)[None, None, :]).all(), "invalid tour"
    assert (srt[:, :, :-N] == 0).all(), "invalid tour"
    d = demand[:, None, :].expand(B, M, -1).gather(2, tours)
    used = torch.zeros(B, M, dtype=demand.dtype)
    for t in range(T):
        used = torch.where(tours[:, :, t] == 0, torch.zeros_like(used), used + d[:, :, t])
        assert (used <= 1 + 1e-4).all(), "capacity exceeded"
