"""Distribution of the sampling mode (CVRP/CVRPModel.py:59-68, TSP/TSPModel.py:46-57): Philox4x32 inverse-CDF draws of
`elg_decode_step(mode = sample)` follow softmax(logits) of the ORACLE, checked by a chi-square test with >= 1e5 draws per
row; the returned probability is the probability of the drawn node; masked (zero-probability) nodes are never drawn.

The reference's zero-probability quirks -- cvrp adds 1e-6 to every probability of the step if a sampled probability is 0
(CVRP/CVRPModel.py:67-68), tsp re-draws the step until none is (TSP/TSPModel.py:47-57) -- guard against
`torch.multinomial` returning a zero-probability entry, which inverse-CDF sampling cannot do (a zero-width interval is
never hit); the last assertion below is that statement: no drawn node ever has probability 0."""
import math

import numpy as np
import pytest
import torch

from helpers import Golden
from oracle import elg_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPL, CALLS = 512, 200          # 512 replicas of the instance per call x 200 calls = 102,400 draws per row


@pytest.mark.parametrize("case,t", [("cvrp_n20", 5), ("cvrp_n20_sharp", 9), ("tsp_n20", 3), ("tsp_n50", 10)])
def test_sample_frequencies_follow_oracle_softmax(case, t):
    from elg_b200 import engine
    g = Golden(case)
    handle = engine.ModelHandle(g.kind, g.model_params(), g.state_dict(), DEV, attention="fp32")
    prob = g.oracle_problem()
    b0 = g.rows_b[0]
    s = g.step(t)
    cur, masked = s["cur"][:1], s["masked"][:1]                       # the first recorded aug-instance: (1, M) / (1, M, N1)
    load = s["load"][:1] if g.kind == "cvrp" else None
    first = g.tours()[b0:b0 + 1, :, 0] if g.kind == "tsp" else None
    # oracle probabilities of that aug-instance at that state
    W = O.Weights(g.state_dict(), g.kind, g.model_params())
    sub = O.Problem(prob.kind, prob.xy[b0:b0 + 1], None if prob.demand is None else prob.demand[b0:b0 + 1], prob.dist[b0:b0 + 1], None, 1)
    cache = O.decoder_cache(W, O.encode(W, sub))
    if g.kind == "tsp":
        O.set_first(W, cache, first)
    logits = O.decode_logits(W, sub, cache, cur, masked, load)
    p_ref = torch.softmax(logits.double(), dim=2)[0]                  # (M, N1)
    M, N1 = p_ref.shape
    # the CUDA path: the same aug-instance replicated REPL times, every replica / call draws independently
    xy = sub.xy.expand(REPL, -1, -1).contiguous().to(DEV)
    dem = None if sub.demand is None else sub.demand.expand(REPL, -1).contiguous().to(DEV)
    batch = engine.encode(handle, xy, dem)
    bits = engine.pack_mask_bits(masked.expand(REPL, -1, -1).contiguous().to(DEV))
    curd = cur.expand(REPL, -1).contiguous().to(DEV)
    loadd = None if load is None else load.expand(REPL, -1).contiguous().to(DEV)
    firstd = None if first is None else first.expand(REPL, -1).contiguous().to(DEV)
    counts = torch.zeros(M, N1, dtype=torch.float64, device=DEV)
    worst_p = 0.0
    min_prob = 1.0
    for c in range(CALLS):
        sel, pr, _ = engine.decode_step(batch, M, curd, bits, load=loadd, first=firstd, mode="sample", seed=1000 + c, step=t)
        counts += torch.zeros(REPL, M, N1, device=DEV, dtype=torch.float64).scatter_(2, sel[:, :, None], 1.0).sum(0)
        if c < 3:      # the returned probability is the probability of the drawn node
            want = p_ref.to(DEV)[torch.arange(M, device=DEV)[None, :].expand(REPL, M), sel]
            worst_p = max(worst_p, float((pr.double() - want).abs().max()))
        min_prob = min(min_prob, float(pr.min()))
    counts = counts.cpu()
    n = REPL * CALLS
    assert worst_p < 5e-4, worst_p                                    # |dp| <= p |dlogit|, logits agree to ~1e-3 at worst
    assert min_prob > 0.0                                             # a zero-probability node is never drawn
    assert float(counts[masked[0]].sum()) == 0.0                      # masked nodes in particular
    # chi-square per row over the nodes with expected count >= 5 (the rest pooled into one bin)
    rows_tested = 0
    for m in range(M):
        e = p_ref[m] * n
        big = e >= 5
        if int(big.sum()) < 2:
            continue                                                  # (almost) deterministic row: nothing to test
        obs = torch.cat((counts[m][big], counts[m][~big].sum()[None]))
        exp = torch.cat((e[big], e[~big].sum()[None]))
        keep = exp >= 5                                               # the pooled rare nodes only if the pool itself is large enough
        if float(exp[-1]) < 5:                                        # otherwise a Poisson tail bound on the pool: P(X >= obs) > 1e-7
            lam, k = float(exp[-1]), int(obs[-1])
            tail = 1.0 - sum(math.exp(-lam) * lam ** i / math.factorial(i) for i in range(k))
            assert k == 0 or tail > 1e-7, (m, lam, k, tail)
        chi2 = float((((obs - exp) ** 2)[keep] / exp[keep]).sum())
        df = int(keep.sum()) - 1
        # Wilson-Hilferty: chi2 is below this bound with probability 1 - 3e-6 per row
        bound = df * (1 - 2 / (9 * df) + 4.5 * math.sqrt(2 / (9 * df))) ** 3
        assert chi2 < bound, (m, chi2, df, bound)
        rows_tested += 1
    assert rows_tested >= M // 2
