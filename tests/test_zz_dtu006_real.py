"""The reference's packaged example on REAL data, through the C-ABI on the GPU: real edge maps ->
polyline graphs (row f1) -> candidate sets (row f2: compatibility-graph communities, SfM-point components) -> pipelines
1 + 2 + 3 -> density limiter -> outlier filter, each stage identical to the CPU oracle (chains, observation lists, 2D coordinates bit for bit; 3D coordinates
within north_star's 1e-4; identical inlier index sets).  profiles/c1_real_dtu006.py is the same run as a script."""
import os
import sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "profiles"))


@pytest.mark.gpu
def test_real_dtu006_end_to_end_matches_oracle():
    import c1_real_dtu006 as run
    res = run.run(os.path.join(HERE, "golden"), 16)
    assert res["candidate_sets_pipeline1"] > 100 and res["candidate_sets_pipeline2"] > 100
    for k in ("pipeline1", "pipeline2", "pipeline3"):
        assert res[k]["identical"] and res[k]["points"] > 50000, res[k]
        assert res[k]["max_abs_xyz_diff"] < 1e-4, res[k]                 # north_star tolerance
    assert res["density_limiter"]["identical"] and 0 < res["density_limiter"]["kept"] < res["density_limiter"]["in"]
    assert res["filter"]["identical_inlier_sets"] and res["filter"]["identical_refined_xyz"]
    assert res["filter"]["edge_point_inliers"] > 1000
