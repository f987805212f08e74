"""dtu006 real-camera case: the 25 cameras, 6268 SfM tracks and LMedS fundamental matrices of the
reference's packaged example (tests/golden/dtu006_sfm.npz, made by tests/golden/make_dtu006_fixture.py) with synthetic
curves projected into them.  CPU part: fixture integrity + OpenMVG JSON round trip; GPU part: pipelines 1-3, density
limiter and outlier filter through the C-ABI, identical to the CPU oracle."""
import json
import os
import numpy as np
import pytest
from edgegraph3d_b200 import synthetic as syn, openmvg_io as io
from edgegraph3d_b200.scene import PointSet

HERE = os.path.dirname(os.path.abspath(__file__))


def dtu_scene(n_curves=160, seed=6):
    z = np.load(os.path.join(HERE, "golden", "dtu006_sfm.npz"))
    sc = syn.make_scene(width=int(z["width"]), height=int(z["height"]), n_curves=n_curves, segs_per_curve=20, curve_len=0.04, seed=seed,
                        closed_frac=0.05, cameras=z["cameras"], fundamental=z["fundamental"], fundamental_valid=z["fundamental_valid"],
                        real_tracks=(z["track_xyz"], z["track_off"], z["track_view"], z["track_xy"]), centers=z["track_xyz"])
    return sc, z


def test_fixture_is_the_packaged_example():
    sc, z = dtu_scene(n_curves=8)
    assert sc.n_views == 25 and sc.n_tracks == 6268 and len(z["track_view"]) == 32890
    assert int(z["fundamental_valid"].sum()) == 590                                      # 10 ordered pairs share < 10 tracks
    # the camera matrices reproject the SfM tracks (radial distortion ignored as in the reference's parser)
    P = z["cameras"].astype(np.float64).reshape(-1, 3, 4)
    off, tv, txy, X = z["track_off"], z["track_view"], z["track_xy"], z["track_xyz"].astype(np.float64)
    mse = []
    for p in range(0, 6268, 13):
        o = slice(off[p], off[p + 1])
        h = P[tv[o]] @ np.append(X[p], 1)
        mse.append((((h[:, :2] / h[:, 2:3]) - txy[o]) ** 2).sum(1).mean())
    assert np.median(mse) < 2.0
    # LMedS F: observations of a track lie within a few px of each other's epipolar lines
    F = z["fundamental"].reshape(25, 25, 3, 3)
    d = []
    for p in range(0, 6268, 17):
        a, b = off[p], off[p + 1] - 1
        if z["fundamental_valid"][tv[a], tv[b]]:
            l = F[tv[a], tv[b]] @ np.append(txy[a], 1)
            d.append(abs(np.append(txy[b], 1) @ l) / np.hypot(l[0], l[1]))
    assert np.median(d) < 2.0


def test_openmvg_json_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    V = 4
    doc = {"sfm_data_version": "0.3", "root_path": "", "control_points": [],
           "views": [{"key": k, "value": {"polymorphic_id": 1, "ptr_wrapper": {"id": k, "data": {"local_path": "", "filename": "%04d.png" % k, "width": 640,
                                                                                          "height": 480, "id_view": k, "id_intrinsic": 0, "id_pose": k}}}} for k in range(V)],
           "intrinsics": [{"key": 0, "value": {"polymorphic_id": 2, "polymorphic_name": "pinhole_radial_k3", "ptr_wrapper": {"id": 9, "data": {
               "width": 640, "height": 480, "focal_length": 520.25, "principal_point": [321.5, 239.25], "disto_k3": [0.01, 0.0, 0.0]}}}}],
           "extrinsics": [], "structure": []}
    for k in range(V):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        doc["extrinsics"].append({"key": k, "value": {"rotation": q.tolist(), "center": rng.normal(size=3).tolist()}})
    d = io.load_sfm_data(doc)
    assert d["cameras"].shape == (V, 12) and d["width"] == 640
    # P = K [R | -R C] in float32
    R, C, K = d["R"][2].astype(np.float64), d["center"][2].astype(np.float64), d["K"][2].astype(np.float64)
    assert np.allclose(d["cameras"][2].reshape(3, 4), K @ np.concatenate([R, (-R @ C)[:, None]], 1), rtol=1e-5, atol=1e-4)
    xyz = rng.normal(size=(5, 3)).astype(np.float32)
    off = np.array([0, 3, 5, 8, 11, 14], np.int64)
    view = rng.integers(0, V, 14).astype(np.int32)
    xy = rng.uniform(0, 480, (14, 2)).astype(np.float32)
    inl = np.array([1, 1, 0, 1, 1], bool)
    out = tmp_path / "out.json"
    assert io.save_sfm_data(str(out), doc, xyz, off, view, xy, inliers=inl) == 4
    back = io.load_sfm_data(str(out))
    keep = np.where(inl)[0]
    assert np.array_equal(back["track_xyz"], xyz[keep])
    assert np.array_equal(back["track_xy"], np.concatenate([xy[off[i]:off[i + 1]] for i in keep]))
    assert np.array_equal(back["track_view"], np.concatenate([view[off[i]:off[i + 1]] for i in keep]))
    assert json.load(open(out))["structure"][0]["value"]["observations"][0]["value"]["id_feat"] == 0
    io.write_ply(str(tmp_path / "p.ply"), xyz)
    assert open(tmp_path / "p.ply").read().startswith("ply\nformat ascii 1.0\nelement vertex 5\n")


def test_lmeds_fundamentals_reproduce_the_fixture():
    cv2 = pytest.importorskip("cv2")
    z = np.load(os.path.join(HERE, "golden", "dtu006_sfm.npz"))
    if str(z["cv2_version"]) != cv2.__version__:
        pytest.skip("fixture made with another OpenCV version")
    # a view subset keeps this quick: tracks restricted to views 0..5
    keep_obs = z["track_view"] < 6
    off = np.concatenate([[0], np.cumsum([keep_obs[z["track_off"][p]:z["track_off"][p + 1]].sum() for p in range(6268)])]).astype(np.int64)
    F, valid = io.fundamental_from_tracks(6, off, z["track_view"][keep_obs], z["track_xy"][keep_obs])
    assert np.array_equal(valid, z["fundamental_valid"][:6, :6])
    assert np.allclose(F[valid == 1], z["fundamental"].reshape(25, 25, 9)[:6, :6][valid == 1], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_dtu006_cameras_pipelines_match_oracle():
    from edgegraph3d_b200 import lib as E
    from tests import oracle_lib as O
    sc, _ = dtu_scene()
    cands = syn.curve_candidate_sets(sc, seed=6)
    osc = O.OracleScene(sc)
    with E.DeviceScene(sc) as dev:
        g12, _ = dev.match_polyline_sets(cands, 0, 6)
        g3, _ = dev.match_refpoints(0, 1200)
        r12 = osc.match_polyline_sets(cands, 0, 6, n_threads=16)
        r3 = osc.match_refpoints(0, 1200, n_threads=16)
        for g, r in ((g12, r12), (g3, r3)):
            assert g.n_points == r.n_points and g.n_points > 100
            assert np.array_equal(g.obs_off, r.obs_off) and np.array_equal(g.obs_view, r.obs_view)
            assert np.array_equal(g.obs_poly, r.obs_poly) and np.array_equal(g.obs_seg, r.obs_seg)
            assert g.obs_xy.tobytes() == r.obs_xy.tobytes()
            assert np.abs(g.xyz - r.xyz).max() < 1e-4          # north_star tolerance
        allp = PointSet.concat([g12, g3])
        keep_g = dev.dedup_close_points(allp)
        keep_o = osc.dedup_close_points(allp)
        assert np.array_equal(keep_g, keep_o) and 0 < keep_g.sum() < allp.n_points
        kept = np.where(keep_g)[0]
        xyz = np.concatenate([sc.track_xyz, allp.xyz[kept]])
        lens = allp.obs_off[kept + 1] - allp.obs_off[kept]
        obs_off = np.concatenate([sc.track_off, int(sc.track_off[-1]) + np.cumsum(lens)])
        idx = np.concatenate([np.arange(allp.obs_off[i], allp.obs_off[i + 1]) for i in kept])
        obs_view = np.concatenate([sc.track_view, allp.obs_view[idx]])
        obs_xy = np.concatenate([sc.track_xy, allp.obs_xy[idx]])
        fx, inl, _ = dev.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks)
        ox, oinl = osc.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks, n_threads=16)[:2]
        assert np.array_equal(inl, oinl)                       # identical inlier index sets
        assert np.array_equal(fx[inl == 1], ox[oinl == 1])
        assert inl[:sc.n_tracks].mean() > 0.5                   # most real SfM tracks survive the 2.25 px^2 filter
