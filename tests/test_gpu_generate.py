"""GPU: device data generators against the distributions of the reference's generate_vrp_data / generate_tsp_data
(CVRP/generate_data.py:9-92, TSP/generate_data.py:9-57): ranges, clamps, demand alphabet, cluster geometry, the
mutated half of `mixed`, determinism per seed."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DIST = {"data_type": "uniform", "n_cluster": 3, "n_cluster_mix": 1, "lower": 0.2, "upper": 0.8, "std": 0.07}


def test_uniform_cvrp():
    from elg_b200.generate_data import CAPACITIES, generate_vrp_data
    d = generate_vrp_data(512, 100, DIST, seed=1)
    assert d["loc"].shape == (512, 100, 2) and d["depot"].shape == (512, 1, 2) and d["demand"].shape == (512, 100)
    for t in (d["loc"], d["depot"]):
        assert float(t.min()) >= 0.0 and float(t.max()) < 1.0
    assert abs(float(d["loc"].mean()) - 0.5) < 5e-3 and abs(float(d["loc"].var()) - 1 / 12) < 3e-3
    q = d["demand"] * CAPACITIES[100]
    assert torch.equal(q, q.round()) and float(q.min()) == 1 and float(q.max()) == 9
    counts = torch.bincount(q.long().flatten(), minlength=10)[1:].float()
    assert float((counts / counts.sum() - 1 / 9).abs().max()) < 0.01
    d2 = generate_vrp_data(512, 100, DIST, seed=1)
    assert torch.equal(d["loc"], d2["loc"]) and torch.equal(d["demand"], d2["demand"])
    assert not torch.equal(d["loc"], generate_vrp_data(512, 100, DIST, seed=2)["loc"])
    with pytest.raises(KeyError):
        generate_vrp_data(4, 37, DIST, seed=1)


@pytest.mark.parametrize("problem", ["cvrp", "tsp"])
def test_cluster(problem):
    from elg_b200.generate_data import generate_tsp_data, generate_vrp_data
    dist = dict(DIST, data_type="cluster")
    N = 100
    if problem == "cvrp":
        d = generate_vrp_data(256, N, dist, seed=3)
        pts = torch.cat((d["loc"], d["depot"]), dim=1)          # the depot was one of the N + 1 clustered points
    else:
        pts = generate_tsp_data(256, N, dist, seed=3)
    assert float(pts.min()) >= 0.0 and float(pts.max()) <= 1.0
    # every point lies within 6 std of one of at most 3 centres inside [lower, upper]^2: per-instance k-means-free
    # check through the group structure (tsp keeps the generation order: 3 consecutive groups of 33, 33, 34)
    if problem == "tsp":
        g = [pts[:, :33], pts[:, 33:66], pts[:, 66:]]
        for grp in g:
            c = grp.mean(dim=1, keepdim=True)
            assert float(c.min()) > 0.2 - 0.05 and float(c.max()) < 0.8 + 0.05
            sd = (grp - c).pow(2).mean(dim=(1, 2)).sqrt()
            assert abs(float(sd.mean()) - 0.07) < 0.01
    # clustered data is far more concentrated than uniform data
    assert float(pts.var(dim=1).mean()) < 0.06


def test_mixed_tsp_half_mutated():
    from elg_b200.generate_data import generate_tsp_data
    dist = dict(DIST, data_type="mixed")
    pts = generate_tsp_data(256, 100, dist, seed=5)
    assert pts.shape == (256, 100, 2) and float(pts.min()) >= 0.0 and float(pts.max()) <= 1.0
    # one cluster of 50 points with std 0.07 around a centre in [0.2, 0.8]^2: the 50 points nearest to the densest spot
    med = pts.median(dim=1, keepdim=True)[0]
    d = (pts - med).norm(dim=2)
    near = (d < 0.25).float().sum(dim=1)
    assert float(near.mean()) > 50          # the cluster (50) plus the uniform points that happen to fall inside
    assert float(near.mean()) < 75
