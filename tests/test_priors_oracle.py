"""CPU checks of the prior post-processing restatement (oracle/priors_oracle.py) against tests/golden/priors.npz — a
fixture produced by EXECUTING the reference's own lines scripts/extract_priors.py:175-196 (tracing loop, quantile,
selection; tests/golden/make_golden_priors.py) — plus properties of the voxel grouping that hold for open3d's
algorithm by construction."""
import numpy as np
import pytest

from helpers import Fixture
from oracle import priors_oracle as PO


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_postprocess_matches_reference_lines(case):
    fx = Fixture("priors.npz")
    g = lambda k: fx.np(f"{case}/{k}")
    out = PO.postprocess_priors(g("in_points"), g("in_features"), g("in_colors"), g("in_densities"), np.zeros(3),
                                float(g("voxel_size")), float(g("hit_thr_ratio")))
    assert out["points"].dtype == np.float32 and out["features"].dtype == np.float16 and out["colors"].dtype == np.float32
    np.testing.assert_array_equal(out["hits"], g("hits"))
    np.testing.assert_array_equal(out["points"], g("points"))
    np.testing.assert_array_equal(out["features"], g("features"))
    np.testing.assert_array_equal(out["colors"], g("colors"))


def test_voxel_grouping_properties():
    g = np.random.default_rng(0)
    pts = g.uniform([-5, -3, -1], [5, 3, 2], (5000, 3)).astype(np.float32)
    centres, traces, vidx = PO.voxel_down_sample_and_trace(pts, 0.4)
    assert sum(len(t) for t in traces) == len(pts) and len(set(np.concatenate(traces))) == len(pts)      # a partition
    _, vmb = PO.voxel_keys(pts, 0.4)
    lo = vmb[None, :] + vidx * 0.4
    assert np.all(centres >= lo - 1e-9) and np.all(centres <= lo + 0.4 + 1e-9)       # centre of mass inside its voxel
    assert np.all(vidx >= 2)                   # the bound is min - 1 - voxel/2: at least 1.2 m = 3 voxels of margin ...
    assert np.all(np.lexsort((vidx[:, 2], vidx[:, 1], vidx[:, 0])) == np.arange(len(vidx)))         # ascending order
    # the lattice origin is (min - 1) - voxel / 2 in doubles: the minimum point itself sits 1.2 m = "3 voxels" above it,
    # which in double arithmetic is 2.9999999999999996 -> voxel 2 (open3d floors the same double)
    p2 = np.array([[0, 0, 0], [0.2 + 1.0, 0, 0]], np.float32)
    idx, vmb2 = PO.voxel_keys(p2, 0.4)
    assert vmb2[0] == -1.0 - 0.2 and idx[0, 0] == int(np.floor((0.0 - vmb2[0]) / 0.4)) == 2 and idx[1, 0] == 6


def test_pickle_reader_round_trip(tmp_path):
    import pickle
    fx = Fixture("priors.npz")
    d = {k: fx.np(f"a/{k}") for k in ("points", "features", "colors", "hits")}
    d["origin"] = np.array([10.0, -20.0, 1.0], np.float32)
    path = tmp_path / "extracted_priors.pkl"
    with open(path, "wb") as f:
        pickle.dump(d, f)
    with open(path, "rb") as f:
        xyz, feats, hits = PO.read_priors_like_city_prior(pickle.load(f))
    assert xyz.shape == d["points"].shape and feats.dtype == np.float16 and hits.shape == (len(xyz), 1)
    np.testing.assert_allclose(xyz[:, 2], d["points"][:, 2] + 1.0, rtol=1e-6)
    np.testing.assert_allclose(xyz[:, 0], -(d["points"][:, 0] + 10.0), rtol=1e-6)
    assert abs(float(hits.mean()) - 1.0) < 1e-5
