"""Library sets (north_star "optimality gap matched on the repo's TSPLIB/CVRPLIB sets"): our drivers on
the GPU against per-instance results of the UNMODIFIED reference drivers (tests/golden/lib, oracle/gen_golden_lib.py):
all 49 TSPLIB instances and 51 of CVRPLIB Set-X including X-n1001-k43 at M = 1000 rows (the large-N regime: > 10 mask
words, uint16 neighbour lists, streamed K/V/E tiles).

What can and cannot be identical.  Library coordinates are integers, so many node pairs are exactly equidistant; the
reference breaks such ties by whatever order `torch.topk`'s partial sort leaves them in, and its local policy is rank-aware
(positional encoding), so the tie order changes tours.  The reference run with index-ordered ties (stable sort, the only
change) is recorded next to the unmodified one (`*_ref_stable.json`, `ties_*.npz`): it differs from the unmodified
reference exactly where we do -- and OUR results equal it.  Hence the gates:
  * against the index-ordered-ties reference: best cost identical on (almost) every instance, size-bin mean gaps within
    1e-3 absolute;
  * against the unmodified reference: bin-mean gaps within 2e-2 absolute (the reference's own tie-order noise), best cost
    within 7 %; tours of the tie fixtures at least as close to the unmodified reference as the reference is to itself;
  * rounded costs are the exact integer length of our own tours (checked in test_gpu_parity / test_drivers)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
LIB = os.path.join(ROOT, "tests", "golden", "lib")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tsplib", "setx"])
def result(request):
    import library_parity as lp
    rows, ref = lp.run_set(request.param)
    stable = None
    path = os.path.join(LIB, request.param + "_ref_stable.json")
    if os.path.exists(path):
        with open(path) as f:
            stable = {r["instance"]: r for r in json.load(f)["instances"]}
    return request.param, rows, stable, lp


def test_costs_against_unmodified_reference(result):
    kind, rows, stable, lp = result
    assert len(rows) == (49 if kind == "tsplib" else 51)
    rel = np.array([abs(r["our_best"] - r["ref_best"]) / r["ref_best"] for r in rows])
    assert rel.max() < 0.07, rel.max()
    # instances without lattice structure are identical (random-uniform style coordinates: no exact ties to speak of)
    assert (rel == 0).sum() >= (25 if kind == "tsplib" else 45)
    a, b = lp.bins(kind, rows, "ref_gap"), lp.bins(kind, rows, "our_gap")
    for k in a:
        assert abs(a[k] - b[k]) / 100 < 2e-2, (k, a[k], b[k])
    for r in rows:
        assert r["T_ref"] == r["T_ours"] or kind == "setx"       # tsp rollouts always take N steps


def test_costs_against_reference_with_index_ordered_ties(result):
    kind, rows, stable, lp = result
    if not stable:
        pytest.skip("no *_ref_stable.json fixture")
    both = [(r, stable[r["instance"]]) for r in rows if r["instance"] in stable]
    assert len(both) >= 0.9 * len(rows)
    equal = sum(r["our_best"] == s["best_cost"] for r, s in both)
    assert equal >= len(both) - max(2, len(both) // 20), (equal, len(both), [(r["instance"], r["our_best"], s["best_cost"]) for r, s in both if r["our_best"] != s["best_cost"]])
    for r, s in both:
        assert abs(r["our_best"] - s["best_cost"]) / s["best_cost"] < 4e-3 or r["our_best"] == s["best_cost"], r["instance"]
    scale = np.array([r["scale"] for r, _ in both])
    ours = np.array([r["our_gap"] for r, _ in both])
    ref = np.array([s["gap"] for _, s in both])
    for m in (scale <= 200, (scale > 200) & (scale <= 500), scale > 500, scale > 0):
        if m.any():
            assert abs(ours[m].mean() - ref[m].mean()) < 1e-3


def test_tie_fixtures_and_large_instances(result):
    kind, rows, stable, lp = result
    by = {r["instance"]: r for r in rows}
    seen = 0
    for r in rows:
        if "rows_tour_equal_vs_stable" in r:
            seen += 1
            assert r["rows_tour_equal_vs_stable"] >= 0.995, r            # ours = the reference with index-ordered ties
            assert r["rows_tour_equal_vs_unmodified"] >= r["ref_unmodified_vs_stable"] - 0.01, r
    assert seen >= 2
    if kind == "setx":
        big = by["X-n1001-k43"]                                           # M = 1000 rows, N + 1 = 1001 nodes, T ~ 1800 steps
        assert big["our_best"] == big["ref_best"] and big["T_ours"] == big["T_ref"]
        assert big["rows_tour_equal"] > 0.8 and big["rows_reward_equal"] > 0.85
        assert by["X-n502-k39"]["our_best"] == by["X-n502-k39"]["ref_best"]
