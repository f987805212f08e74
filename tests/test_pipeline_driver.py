"""Host glue of edgegraph3d_b200/pipeline.py (edge_matching / edge_reconstruction_pipeline shape) on the CPU: a small
synthetic OpenMVG scene written to disk as sfm_data JSON + edge PNGs goes through file reading, polyline-graph
construction (f1), LMedS fundamentals (f4), both candidate-set producers (f2) and the output writers (f3); the device
calls are served by the CPU oracle behind the same method names (the `-m gpu` run of the real example is
tests/test_zz_dtu006_real.py)."""
import json
import os
import numpy as np
import pytest
from edgegraph3d_b200 import pipeline as P, openmvg_io as io
from tests import oracle_lib as O


def write_scene(tmp, V=6, W=640, H=480):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    f, pp = 700.0, (W / 2, H / 2)
    doc = {"sfm_data_version": "0.3", "root_path": "", "control_points": [], "views": [], "extrinsics": [], "structure": [],
           "intrinsics": [{"key": 0, "value": {"polymorphic_id": 2, "polymorphic_name": "pinhole_radial_k3", "ptr_wrapper": {"id": 9, "data": {
               "width": W, "height": H, "focal_length": f, "principal_point": list(pp), "disto_k3": [0.0, 0.0, 0.0]}}}}]}
    cams = []
    for k in range(V):
        ang = 0.12 * (k - V / 2)
        R = np.array([[np.cos(ang), 0, -np.sin(ang)], [0, 1, 0], [np.sin(ang), 0, np.cos(ang)]])
        C = np.array([5 * np.sin(ang), 0.1 * k, -5 * np.cos(ang)])
        doc["views"].append({"key": k, "value": {"polymorphic_id": 1, "ptr_wrapper": {"id": k, "data": {"local_path": "", "filename": "%04d.png" % (2 * k),
                             "width": W, "height": H, "id_view": k, "id_intrinsic": 0, "id_pose": k}}}})
        doc["extrinsics"].append({"key": k, "value": {"rotation": R.tolist(), "center": C.tolist()}})
        K = np.array([[f, 0, pp[0]], [0, f, pp[1]], [0, 0, 1]])
        cams.append(K @ np.concatenate([R, (-R @ C)[:, None]], 1))
    # 3D curves: a few wavy space curves; SfM points = samples of them (+ noise) seen by every view
    curves = []
    for c in range(7):
        t = np.linspace(0, 1, 60)
        a, b = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        curves.append(a[None] * (1 - t[:, None]) + b[None] * t[:, None] + 0.15 * np.sin(6 * t)[:, None] * rng.normal(size=3)[None])
    os.makedirs(tmp / "edges")
    for k in range(V):
        im = np.zeros((H, W, 3), np.uint8)
        for X in curves:
            h = (cams[k] @ np.concatenate([X, np.ones((len(X), 1))], 1).T).T
            px = np.round(h[:, :2] / h[:, 2:3]).astype(np.int32)
            cv2.polylines(im, [px.reshape(-1, 1, 2)], False, (255, 255, 255), 1)
        cv2.imwrite(str(tmp / "edges" / ("%04d.png" % (2 * k))), im)
    pid = 0
    for X in curves:
        for j in range(3, 60, 4):
            obs = []
            for k in range(V):
                h = cams[k] @ np.append(X[j], 1)
                xy = h[:2] / h[2] + rng.normal(0, 0.4, 2)
                obs.append({"key": k, "value": {"id_feat": pid, "x": xy.tolist()}})
            doc["structure"].append({"key": pid, "value": {"X": X[j].tolist(), "observations": obs}})
            pid += 1
    json.dump(doc, open(tmp / "sfm_data.json", "w"))
    return doc


def test_edge_matching_host_glue(tmp_path):
    doc = write_scene(tmp_path)
    info = P.edge_matching(str(tmp_path / "sfm_data.json"), str(tmp_path / "edges"), str(tmp_path / "out"), _scene_factory=O.OracleDevice)
    assert info["candidate_sets_pipeline1"] > 0 and info["candidate_sets_pipeline2"] >= 0
    assert sum(info["points_per_pipeline"]) > 100 and 0 < info["kept_after_density_limiter"] < sum(info["points_per_pipeline"])
    n_sfm = len(doc["structure"])
    before = io.load_sfm_data(str(tmp_path / "out" / "before_filtering.json"))
    after = io.load_sfm_data(str(tmp_path / "out" / "output.json"))
    assert len(before["track_xyz"]) == n_sfm + info["kept_after_density_limiter"] == info["points_before_filtering"]
    assert 0 < len(after["track_xyz"]) == info["points_after_filtering"] < len(before["track_xyz"])
    # the first n_sfm points of before_filtering.json are the input's own, unchanged (add_3dpoints_to_sfmd appends)
    src = io.load_sfm_data(doc)
    assert np.array_equal(before["track_xyz"][:n_sfm], src["track_xyz"]) and np.array_equal(before["track_xy"][:len(src["track_xy"])], src["track_xy"])
    # every new point has >= 3 observations in distinct views of the rig and reprojects onto them
    P12 = src["cameras"].astype(np.float64).reshape(-1, 3, 4)
    for i in range(n_sfm, len(before["track_xyz"]), 7):
        o = slice(before["track_off"][i], before["track_off"][i + 1])
        v = before["track_view"][o]
        assert len(v) >= 3 and len(set(v.tolist())) == len(v)
        h = P12[v] @ np.append(before["track_xyz"][i].astype(np.float64), 1)
        assert (((h[:, :2] / h[:, 2:3]) - before["track_xy"][o]) ** 2).sum(1).mean() < 18.0    # em_GaussNewton accepts mse/(2n) < 9 per coordinate (triangulation.cpp:150-168)


def test_edge_matching_grows_capacities_on_demand(tmp_path):
    """EG3D_ERR_CAPACITY from the library (a chain longer than max_chain_points) makes the driver double the capacities and
    run again; any other error is passed on."""
    from edgegraph3d_b200 import lib as E, _abi as A
    write_scene(tmp_path)
    seen = []

    class Stingy(O.OracleDevice):
        def match_polyline_sets(self, cands, view_begin=0, view_end=None):
            seen.append(int(self.params.max_chain_points))
            if self.params.max_chain_points < 1000:
                raise E.Eg3dError(A.EG3D_ERR_CAPACITY, "a per-seed capacity was exceeded")
            return super().match_polyline_sets(cands, view_begin, view_end)

    info = P.edge_matching(str(tmp_path / "sfm_data.json"), str(tmp_path / "edges"), str(tmp_path / "out"), _scene_factory=Stingy)
    assert seen[:3] == [256, 512, 1024] and info["max_chain_points"] == 1024 and info["max_follow_points"] == 1280

    class Broken(O.OracleDevice):
        def match_polyline_sets(self, cands, view_begin=0, view_end=None):
            raise E.Eg3dError(A.EG3D_ERR_CUDA, "boom")

    with pytest.raises(E.Eg3dError) as ei:
        P.edge_matching(str(tmp_path / "sfm_data.json"), str(tmp_path / "edges"), str(tmp_path / "out2"), _scene_factory=Broken)
    assert ei.value.status == A.EG3D_ERR_CUDA
