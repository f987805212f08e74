"""Oracle vs the cv2-generated known-answer vectors (tests/golden/make_golden_cv2.py): pins the oracle at the
OpenCV boundary — the only third-party arithmetic on the hot path."""
import ctypes as C
import os
import numpy as np
import pytest
from edgegraph3d_b200 import _abi as A
from edgegraph3d_b200.scene import FlatScene
from tests import oracle_lib as O


HERE = os.path.dirname(os.path.abspath(__file__))


def _scene_with_cameras(cams):
    V = cams.shape[0]
    # one dummy 2-vertex polyline per view: the GN / DLT entry points only read the cameras
    return FlatScene(640, 480, cams, np.zeros((V, V, 9)), np.zeros((V, V), np.uint8), np.arange(V + 1), np.arange(V + 1) * 2,
                     np.tile(np.array([[10, 10], [20, 20]], np.float32), (V, 1)), np.zeros(V, np.uint32), np.ones(V, np.uint32))


def test_epiline_bit_exact(golden):
    L = O.lib()
    out = np.zeros(3, np.float32)
    for F, p, ref in zip(golden["epi_F"], golden["epi_pts"], golden["epi_lines"]):
        F = np.ascontiguousarray(F)
        assert L.eg3d_oracle_epiline(A.ptr(F, A.c_f64p), float(p[0]), float(p[1]), A.ptr(out, A.c_f32p))
        assert out.tobytes() == ref.tobytes()


def test_dlt_matches_triangulatePoints(golden):
    L = O.lib()
    cams = golden["dlt_cams"]
    out = np.zeros(4, np.float32)
    worst = 0.0
    for va, vb, x1, x2, ref in zip(golden["dlt_va"], golden["dlt_vb"], golden["dlt_x1"], golden["dlt_x2"], golden["dlt_X4"]):
        P1, P2 = np.ascontiguousarray(cams[va]), np.ascontiguousarray(cams[vb])
        x1, x2 = np.ascontiguousarray(x1), np.ascontiguousarray(x2)
        L.eg3d_oracle_triangulate_dlt(A.ptr(P1, A.c_f32p), A.ptr(P2, A.c_f32p), A.ptr(x1, A.c_f32p), A.ptr(x2, A.c_f32p), A.ptr(out, A.c_f32p))
        a = out[:3].astype(np.float64) / float(out[3])
        b = ref[:3].astype(np.float64) / float(ref[3])
        worst = max(worst, np.abs(a - b).max() / max(1e-3, np.abs(b).max()))
    # only a GN initialiser: float32-cast agreement
    assert worst < 5e-6, worst


def test_opencv_faithful_dlt_is_bit_identical_to_triangulatePoints(golden):
    """The restatement of OpenCV's own Jacobi SVD (triangulate_dlt_opencv, eg3d_params.dlt_wellposed == 2 in the oracle)
    reproduces cv2.triangulatePoints BIT FOR BIT, sign included, on the 400 golden cases and — live, when cv2 is importable —
    on > 99 % of 900 further inputs incl. degenerate ones (the same camera twice, with the same or with another observation),
    where the null space is not a single direction and only the same algorithm in the same operation order lands on the same
    vector; the rest differ by a few float ulps (hypot, see below)."""
    L = O.lib()
    cams = golden["dlt_cams"]
    out = np.zeros(4, np.float32)
    for va, vb, x1, x2, ref in zip(golden["dlt_va"], golden["dlt_vb"], golden["dlt_x1"], golden["dlt_x2"], golden["dlt_X4"]):
        P1, P2 = np.ascontiguousarray(cams[va]), np.ascontiguousarray(cams[vb])
        x1, x2 = np.ascontiguousarray(x1), np.ascontiguousarray(x2)
        L.eg3d_oracle_triangulate_dlt_opencv(A.ptr(P1, A.c_f32p), A.ptr(P2, A.c_f32p), A.ptr(x1, A.c_f32p), A.ptr(x2, A.c_f32p), A.ptr(out, A.c_f32p))
        assert out.tobytes() == np.ascontiguousarray(ref, np.float32).tobytes()
    cv2 = pytest.importorskip("cv2")
    z = np.load(os.path.join(HERE, "golden", "dtu006_sfm.npz"))
    P = z["cameras"].reshape(-1, 3, 4).astype(np.float32)
    rng = np.random.default_rng(5)
    exact, worst = 0, 0.0
    for _ in range(300):
        a, b = rng.choice(len(P), 2, replace=False)
        x1 = rng.uniform([0, 0], [1600, 1200]).astype(np.float32)
        x2 = rng.uniform([0, 0], [1600, 1200]).astype(np.float32)
        for pa, pb, xa, xb in ((P[a], P[b], x1, x2), (P[a], P[a], x1, x1), (P[a], P[a], x1, x2)):
            pa, pb = np.ascontiguousarray(pa), np.ascontiguousarray(pb)
            ref = cv2.triangulatePoints(pa, pb, xa.reshape(2, 1), xb.reshape(2, 1)).reshape(4).astype(np.float32)
            L.eg3d_oracle_triangulate_dlt_opencv(A.ptr(pa.reshape(-1), A.c_f32p), A.ptr(pb.reshape(-1), A.c_f32p), A.ptr(xa, A.c_f32p), A.ptr(xb, A.c_f32p),
                                                 A.ptr(out, A.c_f32p))
            exact += out.tobytes() == ref.tobytes()
            worst = max(worst, float(np.abs(out - ref).max() / np.abs(ref).max()))
    # cv2 calls the C library's hypot (glibc: ~0.8 ulp, not correctly rounded, FMA-build dependent); the restatement uses a
    # correctly rounded one so that CPU and GPU agree: a difference of a few ulps of the float result in a handful of degenerate cases
    assert exact >= 0.99 * 900 and worst < 5e-7, (exact, worst)


def _run_gn(golden, name, fp64):
    sc = _scene_with_cameras(golden["cams"])
    osc = O.OracleScene(sc)
    init = golden[f"{name}_init"].astype(np.float32)
    return osc.gn_triangulate(golden[f"{name}_off"], golden[f"{name}_view"], golden[f"{name}_xy"], init, fp64)


def test_gn64_matches_cv2_rebuild(golden):
    xyz, mse, ok = _run_gn(golden, "gn64", True)
    assert np.array_equal(ok, golden["gn64_ok"])
    sel = golden["gn64_ok"] == 1
    # same operations in the same order on CV_64F: expect bit-level agreement after the float32 narrowing
    assert np.array_equal(xyz[sel], golden["gn64_X"][sel].astype(np.float32))
    assert np.array_equal(mse, golden["gn64_mse"].astype(np.float32))


def test_gn32_matches_cv2_rebuild(golden):
    xyz, mse, ok = _run_gn(golden, "gn32", False)
    assert np.array_equal(ok, golden["gn32_ok"])
    sel = golden["gn32_ok"] == 1
    assert np.array_equal(xyz[sel], golden["gn32_X"][sel].astype(np.float32))
    assert np.array_equal(mse, golden["gn32_mse"].astype(np.float32))


def test_degenerate_dlt_is_a_point_on_the_ray_not_a_defined_value():
    """Why eg3d_params.dlt_wellposed exists.  get_min_max's "last index" quirk (edge_graph_3d_utilities.hpp:86-88) can hand
    cv::triangulatePoints the SAME camera and the SAME observation twice (triangulation.cpp:290).  The DLT system then has
    rank 2: its null space is the whole back-projected ray, and the vector an SVD returns from it depends on the SVD
    implementation (and on how the library builds the system).  Real cv2 and the oracle's Jacobi SVD both return a point that
    reprojects onto the observation — but at unrelated depths, in front of or behind the camera."""
    cv2 = pytest.importorskip("cv2")
    L = O.lib()
    z = np.load(os.path.join(HERE, "golden", "dtu006_sfm.npz"))
    P = z["cameras"].reshape(-1, 3, 4).astype(np.float32)
    rng = np.random.default_rng(0)
    out = np.zeros(4, np.float32)
    depth_ratio = []
    for _ in range(40):
        v = int(rng.integers(0, len(P)))
        x = rng.uniform([200, 200], [1400, 1000]).astype(np.float32)
        Pv = np.ascontiguousarray(P[v])
        X = cv2.triangulatePoints(Pv, Pv, x.reshape(2, 1), x.reshape(2, 1)).reshape(4).astype(np.float64)
        L.eg3d_oracle_triangulate_dlt(A.ptr(Pv.reshape(-1), A.c_f32p), A.ptr(Pv.reshape(-1), A.c_f32p), A.ptr(x, A.c_f32p), A.ptr(x, A.c_f32p),
                                      A.ptr(out, A.c_f32p))
        d = []
        for Y in (X, out.astype(np.float64)):
            h = Pv.astype(np.float64) @ Y
            assert np.abs(h[:2] / h[2] - x).max() < 0.5          # on the ray of the observation ...
            d.append(h[2] / Y[3])
        depth_ratio.append(d[0] / d[1])
    r = np.abs(np.array(depth_ratio))
    assert (r > 3).sum() + (r < 1 / 3).sum() > 20               # ... at depths that have nothing to do with each other
    assert (np.array(depth_ratio) < 0).any()                     # sometimes on opposite sides of the camera


def test_what_the_reference_does_after_a_degenerate_dlt():
    """Consequence of the degenerate DLT for the accept / reject decision, measured on 1 500 real dtu006 tracks (>= 3 views):
    the FP64 Gauss-Newton of the matching path (pinned bit for bit against cv2 above) started from the point REAL cv2 returns
    for the degenerate system accepts about one track in five; started from a well-posed two-view initialiser it accepts all
    of them.  The oracle's own degenerate initialiser (eg3d_params.dlt_wellposed = 0) lands on the same decision as the cv2
    one in ~9 cases of 10; the well-posed policy (the default, = 1) only where cv2's start happens to converge."""
    cv2 = pytest.importorskip("cv2")
    from edgegraph3d_b200 import real_scene
    sc, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    osc, L = O.OracleScene(sc), O.lib()
    P = sc.cameras.reshape(-1, 3, 4)
    rng = np.random.default_rng(1)
    tracks = [t for t in range(sc.n_tracks) if sc.track_off[t + 1] - sc.track_off[t] >= 3]
    obs_off, ov, oxy, init_cv, init_or, init_ok = [0], [], [], [], [], []
    out = np.zeros(4, np.float32)
    for t in rng.choice(tracks, 1500, replace=False):
        o0, o1 = int(sc.track_off[t]), int(sc.track_off[t + 1])
        views, xy = sc.track_view[o0:o1], sc.track_xy[o0:o1]
        k = int(np.argmin(views))
        Pv, x = np.ascontiguousarray(P[views[k]]), np.ascontiguousarray(xy[k])
        X = cv2.triangulatePoints(Pv, Pv, x.reshape(2, 1), x.reshape(2, 1)).reshape(4)
        init_cv.append(X[:3] / X[3])
        L.eg3d_oracle_triangulate_dlt(A.ptr(Pv.reshape(-1), A.c_f32p), A.ptr(Pv.reshape(-1), A.c_f32p), A.ptr(x, A.c_f32p), A.ptr(x, A.c_f32p), A.ptr(out, A.c_f32p))
        init_or.append(out[:3] / out[3])
        j = (k + 1) % len(views)
        Y = cv2.triangulatePoints(Pv, np.ascontiguousarray(P[views[j]]), x.reshape(2, 1), xy[j].reshape(2, 1)).reshape(4)
        init_ok.append(Y[:3] / Y[3])
        ov += views.tolist(); oxy += xy.tolist(); obs_off.append(len(ov))
    obs_off, ov, oxy = np.array(obs_off, np.int64), np.array(ov, np.int32), np.array(oxy, np.float32)
    with np.errstate(all="ignore"):
        ok_cv = osc.gn_triangulate(obs_off, ov, oxy, np.array(init_cv, np.float32), fp64=1)[2]
        ok_or = osc.gn_triangulate(obs_off, ov, oxy, np.array(init_or, np.float32), fp64=1)[2]
        ok_wp = osc.gn_triangulate(obs_off, ov, oxy, np.array(init_ok, np.float32), fp64=1)[2]
    assert ok_wp.mean() > 0.95
    assert ok_cv.mean() < 0.4
    assert (ok_cv == ok_or).mean() > 0.8
