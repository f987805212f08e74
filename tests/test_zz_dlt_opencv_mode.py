"""eg3d_params.dlt_wellposed = 2 — the DEFAULT since round 2: the get_min_max quirk's camera pair handed to OpenCV's own SVD
restated, i.e. what a reference linked against OpenCV 4.x computes also after a degenerate 2-view DLT — on the
GPU against the oracle, on a synthetic scene and on the packaged dtu006 example (all three pipelines)."""
import os
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _same(g, r):
    return (g.n_points == r.n_points and np.array_equal(g.obs_off, r.obs_off) and np.array_equal(g.obs_view, r.obs_view)
            and np.array_equal(g.obs_poly, r.obs_poly) and np.array_equal(g.obs_seg, r.obs_seg) and g.obs_xy.tobytes() == r.obs_xy.tobytes())


@pytest.mark.gpu
def test_opencv_faithful_mode_matches_oracle_synthetic_and_real():
    from edgegraph3d_b200 import lib as E, synthetic as syn, pipeline as P, real_scene
    from tests import oracle_lib as O
    assert E.default_params().dlt_wellposed == 2 and O.default_params().dlt_wellposed == 2
    sc = syn.make_scene(n_views=9, n_curves=20, seed=7, closed_frac=0.2, drop_view_frac=0.1)
    cands = syn.curve_candidate_sets(sc, seed=7)
    prm = E.default_params(dlt_wellposed=2)
    with E.DeviceScene(sc, prm) as dev:
        g, _ = dev.match_polyline_sets(cands)
    r = O.OracleScene(sc, prm).match_polyline_sets(cands, n_threads=16)
    assert _same(g, r) and g.n_points > 100 and np.abs(g.xyz - r.xyz).max() < 1e-4
    real, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    c1, c2, _ = P.candidate_sets(real)
    prm = E.default_params(dlt_wellposed=2, **P.REAL_DATA_CAPACITIES)
    odev = O.OracleDevice(real, prm, n_threads=16)
    with E.DeviceScene(real, prm) as dev:
        gp, _ = P.run_pipelines(dev, real, c1, c2)
    rp, _ = P.run_pipelines(odev, real, c1, c2)
    for g, r in zip(gp, rp):
        assert _same(g, r) and np.abs(g.xyz - r.xyz).max() < 1e-4
    assert [p.n_points for p in rp] == [163214, 115855, 184852]      # the OpenCV-faithful run of DESIGN.md §2a
