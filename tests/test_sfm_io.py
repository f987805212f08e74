"""Row f3 in C++ (eg3d_sfm_load / eg3d_sfm_save / eg3d_write_ply, host code of libeg3d.so) against the Python reading of the same
reference functions (edgegraph3d_b200/openmvg_io.py: OpenMvgParser.cpp:75-153, 241-301; output_sfm_data.cpp:186-229;
output_point_cloud.cpp): cameras and tracks bit for bit — the camera matrices of openmvg_io are themselves pinned against the
reference's vendored glm (tests/test_glm_golden.py) — and a save -> load round trip.  When the reference tree is present (build
container only) the packaged example/dtu006/input.json is read by both."""
import json
import os
import numpy as np
import pytest
from edgegraph3d_b200 import lib as E, openmvg_io as io


def _doc(V, rng, n_pts=40, shuffled_keys=False):
    keys = list(range(V))
    if shuffled_keys:
        keys = [int(k) for k in rng.permutation(np.arange(3, 3 + 2 * V, 2))]       # pose keys that are not 0..V-1 and not sorted
    doc = {"sfm_data_version": "0.3", "root_path": "/some/where", "control_points": [],
           "views": [{"key": i, "value": {"polymorphic_id": 1, "ptr_wrapper": {"id": i, "data": {"local_path": "", "filename": "%04d.png" % i, "width": 640, "height": 480,
                                                                                          "id_view": i, "id_intrinsic": i % 2, "id_pose": k}}}} for i, k in enumerate(keys)],
           "intrinsics": [{"key": j, "value": {"polymorphic_id": 2, "polymorphic_name": "pinhole_radial_k3", "ptr_wrapper": {"id": 9 + j, "data": {
               "width": 640, "height": 480, "focal_length": 520.25 + 3.1 * j, "principal_point": [321.5 - j, 239.25 + j], "disto_k3": [0.01, -0.2, 0.0]}}}} for j in range(2)],
           "extrinsics": [], "structure": []}
    for k in keys:
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        doc["extrinsics"].append({"key": k, "value": {"rotation": q.tolist(), "center": rng.normal(size=3).tolist()}})
    for p in range(n_pts):
        obs = [{"key": int(k), "value": {"id_feat": int(rng.integers(0, 999)), "x": rng.uniform(0, 640, 2).tolist()}} for k in rng.choice(keys, size=int(rng.integers(2, V + 1)), replace=False)]
        doc["structure"].append({"key": p * 3 + 1, "value": {"X": rng.normal(size=3).tolist(), "observations": obs}})
    return doc


def _same(a, b):
    assert a["width"] == b["width"] and a["height"] == b["height"] and list(a["view_keys"]) == list(b["view_keys"])
    for k in ("cameras", "K", "R", "center", "t", "track_xyz", "track_off", "track_view", "track_xy"):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and x.dtype == y.dtype and x.tobytes() == y.tobytes(), k


@pytest.mark.parametrize("seed,V,shuffled", [(1, 4, False), (2, 9, True), (3, 25, True)])
def test_cpp_loader_equals_the_python_loader_bit_for_bit(tmp_path, seed, V, shuffled):
    doc = _doc(V, np.random.default_rng(seed), shuffled_keys=shuffled)
    p = tmp_path / "in.json"
    p.write_text(json.dumps(doc, indent=1 if seed % 2 else None))
    _same(E.sfm_load(p), io.load_sfm_data(str(p)))


def test_cpp_save_round_trips_and_keeps_the_original_sections(tmp_path):
    rng = np.random.default_rng(5)
    doc = _doc(6, rng, shuffled_keys=True)
    src = tmp_path / "in.json"
    src.write_text(json.dumps(doc))
    xyz = rng.normal(size=(7, 3)).astype(np.float32)
    off = np.array([0, 3, 5, 8, 11, 14, 14, 17], np.int64)            # one point without observations
    view = rng.integers(0, 6, 17).astype(np.int32)
    xy = rng.uniform(0, 480, (17, 2)).astype(np.float32)
    inl = np.array([1, 1, 0, 1, 1, 1, 0], bool)
    out_c, out_py = tmp_path / "c.json", tmp_path / "py.json"
    assert E.sfm_save(out_c, src, xyz, off, view, xy, inliers=inl) == 5
    assert io.save_sfm_data(str(out_py), doc, xyz, off, view, xy, inliers=inl) == 5
    a, b = json.load(open(out_c)), json.load(open(out_py))
    assert a == b                                                     # same document, number for number
    back = E.sfm_load(out_c)
    keep = np.where(inl)[0]
    assert np.array_equal(back["track_xyz"], xyz[keep])
    assert np.array_equal(back["track_xy"], np.concatenate([xy[off[i]:off[i + 1]] for i in keep]))
    assert np.array_equal(back["track_view"], np.concatenate([view[off[i]:off[i + 1]] for i in keep]))
    assert E.sfm_save(tmp_path / "all.json", src, xyz, off, view, xy) == 7


def test_cpp_ply_equals_the_python_ply(tmp_path):
    rng = np.random.default_rng(9)
    xyz = (rng.normal(size=(50, 3)) * np.array([1, 1e-5, 1e4])).astype(np.float32)
    rgb = rng.integers(0, 256, (50, 3)).astype(np.uint8)
    E.write_ply(tmp_path / "a.ply", xyz); io.write_ply(str(tmp_path / "b.ply"), xyz)
    assert (tmp_path / "a.ply").read_text() == (tmp_path / "b.ply").read_text()
    E.write_ply(tmp_path / "c.ply", xyz, rgb); io.write_ply(str(tmp_path / "d.ply"), xyz, rgb)
    assert (tmp_path / "c.ply").read_text() == (tmp_path / "d.ply").read_text()


def test_cpp_reader_on_unusual_but_valid_json_text(tmp_path):
    """Scientific notation, negative zero, integers where floats are expected, escaped strings, tabs / CRLF between tokens, unknown keys."""
    doc = _doc(3, np.random.default_rng(11))
    doc["root_path"] = 'C:\\data\\"dtu" A\n\ttab'                 # backslashes, quotes, control characters -> escapes in the file
    doc["extra"] = {"a": [1, 2, {"b": None}], "t": True, "f": False}
    doc["extrinsics"][1]["value"]["center"] = [-0.0, 1, 2.0]
    doc["structure"][0]["value"]["X"] = [1, -2, 3]
    text = json.dumps(doc)
    assert '"focal_length": 520.25' in text and '"principal_point": [321.5, 239.25]' in text
    text = text.replace('"focal_length": 520.25', '"focal_length": 5.2025E+2').replace('"principal_point": [321.5, 239.25]', '"principal_point": [3215e-1,\t239.25]')
    text = text.replace('"center": [', '"center":\r\n [ ', 1).replace(", ", " ,\n\t", 40)
    assert json.loads(text) == doc
    p = tmp_path / "odd.json"
    p.write_text(text)
    _same(E.sfm_load(p), io.load_sfm_data(str(p)))
    # the verbatim sections survive a save with their escapes intact
    out = tmp_path / "out.json"
    assert E.sfm_save(out, p, np.zeros((1, 3), np.float32), np.array([0, 1], np.int64), np.array([0], np.int32), np.array([[1.5, 2.5]], np.float32)) == 1
    back = json.load(open(out))
    assert back["root_path"] == doc["root_path"] and back["views"] == doc["views"] and back["extrinsics"] == doc["extrinsics"]


def test_bad_files_are_refused(tmp_path):
    with pytest.raises(E.Eg3dError):
        E.sfm_load(tmp_path / "missing.json")
    (tmp_path / "broken.json").write_text('{"views": [1, 2,')
    with pytest.raises(E.Eg3dError):
        E.sfm_load(tmp_path / "broken.json")
    (tmp_path / "empty.json").write_text('{"views": [], "intrinsics": [], "extrinsics": []}')
    with pytest.raises(E.Eg3dError):
        E.sfm_load(tmp_path / "empty.json")


REF_JSON = "/root/reference/example/dtu006/input.json"


@pytest.mark.skipif(not os.path.exists(REF_JSON), reason="the reference tree is only present in the build container")
def test_packaged_example_file_read_by_both():
    a, b = E.sfm_load(REF_JSON), io.load_sfm_data(REF_JSON)
    _same(a, b)
    assert a["cameras"].shape == (25, 12) and len(a["track_off"]) - 1 == 6268 and len(a["track_view"]) == 32890
