"""Row f4 without OpenCV: eg3d_fundamental_from_tracks (host C++ of libeg3d.so; generate_all_fundamental_matrices_from_Points,
geometric_utilities.cpp:754-820).  cv::findFundamentalMat(FM_LMEDS) is a randomised estimator, so the bar is not bits: the pair set
must be the reference's (>= 10 common tracks), exact data must give the cameras' own epipolar geometry, gross outliers must be
rejected, and on the real tracks of the packaged example the matrices must explain the tracks as well as the cv2 matrices of the
committed fixture do."""
import os
import numpy as np
import pytest
from edgegraph3d_b200 import lib as E, synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))


def _tracks(cams, X, rng, noise=0.0, outlier_frac=0.0, miss=0.2):
    V = cams.shape[0]
    P = cams.reshape(V, 3, 4).astype(np.float64)
    off, view, xy = [0], [], []
    for x in X:
        for v in range(V):
            if rng.random() < miss:
                continue
            h = P[v] @ np.r_[x, 1.0]
            p = h[:2] / h[2] + rng.normal(0, noise, 2)
            if rng.random() < outlier_frac:
                p = rng.uniform(0, 1000, 2)
            view.append(v); xy.append(p)
        off.append(len(view))
    return np.array(off, np.int64), np.array(view, np.int32), np.array(xy, np.float32).reshape(-1, 2)


def _epi_dist(F, a, b):
    l = (F.reshape(3, 3) @ np.c_[a, np.ones(len(a))].T).T
    return np.abs((l[:, :2] * b).sum(1) + l[:, 2]) / np.hypot(l[:, 0], l[:, 1])


def _pairs(V, off, view, xy):
    seen = [dict() for _ in range(V)]
    for p in range(len(off) - 1):
        for o in range(int(off[p]), int(off[p + 1])):
            seen[int(view[o])][p] = o
    def common(i, j):
        c = sorted(set(seen[i]) & set(seen[j]))
        return (np.array([xy[seen[i][p]] for p in c], np.float64).reshape(-1, 2), np.array([xy[seen[j][p]] for p in c], np.float64).reshape(-1, 2))
    return common


def _rig(seed, V=6):
    sc = syn.make_scene(n_views=V, n_curves=4, seed=seed)
    rng = np.random.default_rng(seed)
    X = rng.uniform(-0.4, 0.4, (120, 3)) + np.array([0, 0, 0.0])
    return sc.cameras.reshape(V, 12), X, rng


def test_exact_tracks_give_the_cameras_own_geometry():
    cams, X, rng = _rig(3)
    V = cams.shape[0]
    off, view, xy = _tracks(cams, X, rng)
    F, valid = E.fundamental_from_tracks(V, off, view, xy)
    Fa = E.camera_fundamentals(cams)
    common = _pairs(V, off, view, xy)
    assert valid.sum() == V * (V - 1)
    for i in range(V):
        for j in range(V):
            if i == j:
                assert not valid[i, j] and not F[i, j].any()
                continue
            a, b = common(i, j)
            assert _epi_dist(F[i, j], a, b).max() < 2e-2            # float32 observations of a ~1000 px image
            f, g = F[i, j] / np.linalg.norm(F[i, j]), Fa[i, j] / np.linalg.norm(Fa[i, j])
            assert min(np.abs(f - g).max(), np.abs(f + g).max()) < 1e-3
            assert abs(F[i, j, 8] - 1.0) < 1e-12 or abs(np.linalg.norm(F[i, j]) - 1.0) < 1e-12      # OpenCV's scaling
            assert abs(np.linalg.det(F[i, j].reshape(3, 3))) < 1e-9 * np.linalg.norm(F[i, j]) ** 3   # rank 2


def test_gross_outliers_are_rejected_and_the_call_is_deterministic():
    cams, X, rng = _rig(5)
    V = cams.shape[0]
    off, view, xy = _tracks(cams, X, rng, noise=0.3, outlier_frac=0.15)      # ~28 % of the CORRESPONDENCES have a wild end (LMedS breaks down at 50 %)
    clean_off, clean_view, clean_xy = _tracks(cams, X, np.random.default_rng(5), noise=0.0, outlier_frac=0.0, miss=0.0)
    F, valid = E.fundamental_from_tracks(V, off, view, xy)
    F2, valid2 = E.fundamental_from_tracks(V, off, view, xy)
    assert F.tobytes() == F2.tobytes() and np.array_equal(valid, valid2)
    common = _pairs(V, clean_off, clean_view, clean_xy)
    med = [np.median(_epi_dist(F[i, j], *common(i, j))) for i in range(V) for j in range(V) if i != j]
    assert np.median(med) < 0.5 and max(med) < 3.0, (np.median(med), max(med))


def test_pairs_with_fewer_than_ten_common_tracks_stay_invalid():
    cams, X, rng = _rig(7, V=4)
    off, view, xy = _tracks(cams, X[:9], rng, miss=0.0)       # 9 tracks, all views
    F, valid = E.fundamental_from_tracks(4, off, view, xy)
    assert not valid.any() and not F.any()
    F, valid = E.fundamental_from_tracks(4, off, view, xy, min_common=8)
    assert valid.sum() == 12


def test_degenerate_inputs_do_not_break_the_call():
    F, v = E.fundamental_from_tracks(3, np.array([0], np.int64), np.zeros(0, np.int32), np.zeros((0, 2), np.float32))
    assert not v.any()
    off = np.arange(0, 62, 2).astype(np.int64); view = np.tile([0, 1], 30).astype(np.int32)
    same = np.tile([[100., 100.], [200., 200.]], (30, 1)).astype(np.float32)               # 30 copies of one correspondence
    F, v = E.fundamental_from_tracks(2, off, view, same)
    assert not v.any() and np.isfinite(F).all()
    xs = np.linspace(0, 500, 30)
    line = np.stack([np.stack([xs, xs], 1), np.stack([xs + 5, xs * 0.5], 1)], 1).reshape(-1, 2).astype(np.float32)   # collinear points: rank-deficient
    F, v = E.fundamental_from_tracks(2, off, view, line)
    assert np.isfinite(F).all()
    with pytest.raises(E.Eg3dError):
        E.fundamental_from_tracks(2, off, np.full_like(view, 5), line)                   # a view id outside the scene


def test_real_dtu006_tracks_as_good_as_the_cv2_fixture():
    d = np.load(os.path.join(HERE, "golden", "dtu006_sfm.npz"))
    V = d["cameras"].shape[0]
    F, valid = E.fundamental_from_tracks(V, d["track_off"], d["track_view"], d["track_xy"])
    assert np.array_equal(valid, d["fundamental_valid"])                    # the reference's pair set (590 ordered pairs)
    common = _pairs(V, d["track_off"], d["track_view"], d["track_xy"])
    own, cv = [], []
    for i in range(V):
        for j in range(V):
            if valid[i, j]:
                a, b = common(i, j)
                own.append(np.median(_epi_dist(F[i, j], a, b))); cv.append(np.median(_epi_dist(d["fundamental"][i, j], a, b)))
    own, cv = np.array(own), np.array(cv)
    assert np.median(own) <= 1.05 * np.median(cv)                           # measured: 0.74 px against 0.82 px
    assert np.percentile(own, 90) <= 1.1 * np.percentile(cv, 90)
    assert own.max() <= 2.5                                                 # cv2's worst pair: 1.65 px


@pytest.mark.gpu
def test_native_fundamentals_drive_the_path_gpu_equals_oracle():
    """The packaged example with libeg3d.so's own LMedS matrices in place of the cv2 ones: a different (equally valid) input, on
    which the GPU path must still equal the oracle — pipeline 2, chains / observation lists / 2D bits / 3D."""
    import dataclasses
    from edgegraph3d_b200 import real_scene, pipeline as P, openmvg_io as io
    from tests import oracle_lib as O
    from tests.test_gpu_parity import assert_points_parity
    sc, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    F, valid = io.fundamental_from_tracks(sc.n_views, sc.track_off, sc.track_view, sc.track_xy, engine="native")
    sc = dataclasses.replace(sc, fundamental=F.reshape(sc.n_views, sc.n_views, 9), fundamental_valid=valid)
    _, cands2, _ = P.candidate_sets(sc)
    prm = E.default_params(**P.REAL_DATA_CAPACITIES)
    with E.DeviceScene(sc, prm) as dev:
        gpu, _ = dev.match_polyline_sets(cands2)
    ref = O.OracleDevice(sc, prm, n_threads=16).match_polyline_sets(cands2)
    ref = ref[0] if isinstance(ref, tuple) else ref
    assert gpu.n_points > 50000
    assert_points_parity(sc, gpu, ref)
