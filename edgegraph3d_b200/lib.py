"""ctypes binding of edgegraph3d_b200/libeg3d.so (the C-ABI of include/eg3d.h).

The library is the product: if it is missing or no CUDA device is usable every call raises — there is no CPU
fallback and nothing here imports the oracle.
"""
import ctypes as C
import os
import numpy as np
from . import _abi as A
from .scene import PointSet

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EG3D_LIB", os.path.join(_HERE, "libeg3d.so"))  # EG3D_LIB: A/B builds of the same library
_lib = None

EXPORTS = [
    "eg3d_last_error", "eg3d_camera_fundamentals", "eg3d_device_count", "eg3d_params_default", "eg3d_scene_create", "eg3d_scene_destroy",
    "eg3d_sample_seeds", "eg3d_epipolar_intersect", "eg3d_epipolar_intersect_device", "eg3d_hits_get", "eg3d_hits_free", "eg3d_match_seeds",
    "eg3d_match_polyline_sets", "eg3d_match_refpoints", "eg3d_points_get", "eg3d_points_device_get", "eg3d_points_free", "eg3d_gn_triangulate",
    "eg3d_gn_triangulate_device", "eg3d_dedup_close_points", "eg3d_filter",
    "eg3d_plg_from_edge_image", "eg3d_plg_get", "eg3d_plg_free",
    "eg3d_polyline_sets_from_refpoints", "eg3d_polyline_sets_get", "eg3d_polyline_sets_free",
    "eg3d_polyline_similarity_graph", "eg3d_similarity_graph_get", "eg3d_similarity_graph_communities", "eg3d_polyline_sets_from_communities",
    "eg3d_similarity_graph_free", "eg3d_triangulate_dlt_host", "eg3d_build_info", "eg3d_project_host",
    "eg3d_comm_unique_id", "eg3d_comm_create", "eg3d_comm_destroy", "eg3d_points_allgather",
    "eg3d_match_correspondences", "eg3d_refpoint_correspondences", "eg3d_corr_get", "eg3d_corr_free", "eg3d_fundamental_from_tracks",
    "eg3d_sfm_load", "eg3d_sfm_get", "eg3d_sfm_free", "eg3d_sfm_save", "eg3d_write_ply",
]


class Eg3dError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"eg3d status {status}: {msg}")
        self.status = status


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Eg3dError(-1, f"{LIB_PATH} is missing: build it with edgegraph3d_b200/csrc/build.sh "
                            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.eg3d_last_error.restype = C.c_char_p
    L.eg3d_device_count.restype = C.c_int
    L.eg3d_camera_fundamentals.argtypes = [A.c_f32p, C.c_int32, A.c_f64p]
    L.eg3d_fundamental_from_tracks.argtypes = [C.POINTER(A.SceneDesc), C.c_int32, A.c_f64p, A.c_u8p]
    L.eg3d_sfm_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.eg3d_sfm_get.argtypes = [C.c_void_p, C.POINTER(A.SfmView)]
    L.eg3d_sfm_free.argtypes = [C.c_void_p]
    L.eg3d_sfm_free.restype = None
    L.eg3d_sfm_save.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, A.c_f32p, A.c_i64p, A.c_i32p, A.c_f32p, A.c_u8p, A.c_i64p]
    L.eg3d_write_ply.argtypes = [C.c_char_p, C.c_int64, A.c_f32p, A.c_u8p]
    L.eg3d_params_default.argtypes = [C.POINTER(A.Params)]
    L.eg3d_scene_create.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Params), C.POINTER(C.c_void_p)]
    L.eg3d_scene_destroy.argtypes = [C.c_void_p]
    L.eg3d_sample_seeds.argtypes = [C.POINTER(A.SceneDesc), A.c_i32p, A.c_u32p, C.c_int64, C.c_float, C.c_int64,
                                    A.c_i32p, A.c_u32p, A.c_u32p, A.c_f32p, A.c_i32p, A.c_i64p]
    L.eg3d_epipolar_intersect.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates), C.POINTER(C.c_void_p), C.POINTER(A.Timing)]
    L.eg3d_epipolar_intersect_device.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates), C.POINTER(A.Timing)]
    L.eg3d_hits_get.argtypes = [C.c_void_p, A.c_i64p, A.c_i32p, C.POINTER(A.c_i64p), C.POINTER(C.POINTER(A.Hit))]
    L.eg3d_hits_free.argtypes = [C.c_void_p]
    L.eg3d_match_seeds.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates), C.POINTER(C.c_void_p), C.POINTER(A.Timing)]
    L.eg3d_match_polyline_sets.argtypes = [C.c_void_p, C.POINTER(A.Candidates), C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(A.Timing)]
    L.eg3d_match_refpoints.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(A.Timing)]
    L.eg3d_points_get.argtypes = [C.c_void_p, C.POINTER(A.PointsView)]
    L.eg3d_points_device_get.argtypes = [C.c_void_p, C.POINTER(A.PointsView)]
    L.eg3d_points_free.argtypes = [C.c_void_p]
    L.eg3d_gn_triangulate.argtypes = [C.c_void_p, C.c_int64, A.c_i64p, A.c_i32p, A.c_f32p, A.c_f32p, C.c_int, A.c_f32p, A.c_f32p,
                                      A.c_u8p, C.POINTER(A.Timing)]
    L.eg3d_gn_triangulate_device.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(A.Timing)]
    L.eg3d_dedup_close_points.argtypes = [C.c_void_p, C.POINTER(A.PointsView), A.c_u8p]
    L.eg3d_filter.argtypes = [C.c_void_p, C.c_int64, A.c_f32p, A.c_i64p, A.c_i32p, A.c_f32p, C.c_int64, C.c_float, C.c_int32,
                              A.c_u8p, C.POINTER(A.Timing)]
    L.eg3d_plg_from_edge_image.argtypes = [A.c_u8p, C.c_int32, C.c_int32, C.c_int32, A.c_u8p, C.c_int32, C.POINTER(C.c_void_p)]
    L.eg3d_plg_get.argtypes = [C.c_void_p, C.POINTER(A.PlgView)]
    L.eg3d_plg_free.argtypes = [C.c_void_p]
    L.eg3d_polyline_sets_from_refpoints.argtypes = [C.POINTER(A.SceneDesc), C.c_float, C.c_float, C.POINTER(C.c_void_p)]
    L.eg3d_polyline_sets_get.argtypes = [C.c_void_p, C.POINTER(A.Candidates), A.c_i64p, C.POINTER(A.c_i64p)]
    L.eg3d_polyline_sets_free.argtypes = [C.c_void_p]
    L.eg3d_polyline_similarity_graph.argtypes = [C.POINTER(A.SceneDesc), C.c_float, C.POINTER(C.c_void_p)]
    L.eg3d_similarity_graph_get.argtypes = [C.c_void_p, C.POINTER(A.SimilarityGraphView)]
    L.eg3d_similarity_graph_communities.argtypes = [C.c_void_p, A.c_i64p, A.c_f64p]
    L.eg3d_polyline_sets_from_communities.argtypes = [C.c_void_p, A.c_i64p, C.POINTER(C.c_void_p)]
    L.eg3d_similarity_graph_free.argtypes = [C.c_void_p]
    L.eg3d_build_info.restype = C.c_char_p
    L.eg3d_project_host.argtypes = [A.c_f32p, A.c_f32p, A.c_f32p]
    L.eg3d_triangulate_dlt_host.argtypes = [A.c_f32p, A.c_f32p, A.c_f32p, A.c_f32p, C.c_int32, A.c_f32p]
    L.eg3d_match_correspondences.argtypes = [C.c_void_p, C.c_int64, A.c_i32p, A.c_i64p, C.POINTER(A.Hit), C.POINTER(C.c_void_p), C.POINTER(A.Timing)]
    L.eg3d_refpoint_correspondences.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
    L.eg3d_corr_get.argtypes = [C.c_void_p, A.c_i64p, A.c_i32p, C.POINTER(A.c_i32p), C.POINTER(A.c_u32p), C.POINTER(A.c_u32p), C.POINTER(A.c_f32p), C.POINTER(A.c_i64p),
                                C.POINTER(A.c_i64p), C.POINTER(C.POINTER(A.Hit))]
    L.eg3d_corr_free.argtypes = [C.c_void_p]
    L.eg3d_comm_unique_id.argtypes = [A.c_u8p]
    L.eg3d_comm_create.argtypes = [C.c_void_p, A.c_u8p, C.c_int32, C.c_int32]
    L.eg3d_comm_destroy.argtypes = [C.c_void_p]
    L.eg3d_points_allgather.argtypes = [C.c_void_p, C.c_void_p, A.c_i64p, C.POINTER(C.c_void_p), C.POINTER(A.Timing)]
    _lib = L
    return L


def _check(st):
    if st != A.EG3D_OK:
        raise Eg3dError(st, load().eg3d_last_error().decode(errors="replace"))


def build_info():
    """{'EG3D_DLT_OPENCV': '0', ...}: the compile-time switches of the loaded library."""
    return dict(kv.split("=", 1) for kv in load().eg3d_build_info().decode().split())


def default_params(**overrides):
    p = A.Params()
    load().eg3d_params_default(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def camera_fundamentals(cameras):
    """F[a][b] with x_b^T F x_a = 0, from the camera matrices alone (host code; used for exact-safe pruning)."""
    cams = np.ascontiguousarray(cameras, np.float32).reshape(-1, 12)
    V = cams.shape[0]
    out = np.zeros((V, V, 9), np.float64)
    load().eg3d_camera_fundamentals(A.ptr(cams, A.c_f32p), V, A.ptr(out, A.c_f64p))
    return out


def fundamental_from_tracks(n_views, track_off, track_view, track_xy, min_common=10):
    """f4 (host C++, no OpenCV): least-median-of-squares F[i][j] from the SfM tracks seen by both views
    (generate_all_fundamental_matrices_from_Points, geometric_utilities.cpp:754-820).  -> (F [V][V][9] f64, valid [V][V] u8).
    Same estimator family as cv2.findFundamentalMat(FM_LMEDS), not bit-identical to it (openmvg_io.fundamental_from_tracks is the
    cv2 call that reproduces the committed dtu006 fixture)."""
    d = A.SceneDesc()
    off = np.ascontiguousarray(track_off, np.int64); tv = np.ascontiguousarray(track_view, np.int32)
    xy = np.ascontiguousarray(track_xy, np.float32).reshape(-1, 2)
    d.n_views = int(n_views); d.n_tracks = len(off) - 1
    d.track_off = A.ptr(off, A.c_i64p); d.track_view = A.ptr(tv, A.c_i32p); d.track_xy = A.ptr(xy, A.c_f32p)
    F = np.zeros((n_views, n_views, 9), np.float64); valid = np.zeros((n_views, n_views), np.uint8)
    _check(load().eg3d_fundamental_from_tracks(C.byref(d), int(min_common), A.ptr(F, A.c_f64p), A.ptr(valid, A.c_u8p)))
    return F, valid


def sfm_load(path):
    """f3 (host C++): OpenMVG sfm_data JSON -> the dict openmvg_io.load_sfm_data returns (same keys, same bits)."""
    L = load()
    h = C.c_void_p()
    _check(L.eg3d_sfm_load(str(path).encode(), C.byref(h)))
    try:
        v = A.SfmView()
        _check(L.eg3d_sfm_get(h, C.byref(v)))
        V, NT = int(v.n_views), int(v.n_tracks)
        arr = lambda p, shape, dt: (np.ctypeslib.as_array(p, shape=shape).astype(dt, copy=True) if int(np.prod(shape)) > 0 else np.zeros(shape, dt))
        off = arr(v.track_off, (NT + 1,), np.int64)
        nobs = int(off[-1])
        return dict(width=int(v.width), height=int(v.height), cameras=arr(v.cameras, (V, 12), np.float32), K=arr(v.K, (V, 3, 3), np.float32),
                    R=arr(v.R, (V, 3, 3), np.float32), center=arr(v.center, (V, 3), np.float32), t=arr(v.t, (V, 3), np.float32),
                    track_xyz=arr(v.track_xyz, (NT, 3), np.float32), track_off=off, track_view=arr(v.track_view, (nobs,), np.int32),
                    track_xy=arr(v.track_xy, (nobs, 2), np.float32), view_keys=[int(k) for k in arr(v.view_keys, (V,), np.int64)])
    finally:
        L.eg3d_sfm_free(h)


def sfm_save(path, original_path, xyz, obs_off, obs_view, obs_xy, inliers=None):
    """f3 (host C++): output_sfm_data — the original file's views / intrinsics / extrinsics + a rewritten `structure`.  -> points written."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3); off = np.ascontiguousarray(obs_off, np.int64)
    ov = np.ascontiguousarray(obs_view, np.int32); xy = np.ascontiguousarray(obs_xy, np.float32).reshape(-1, 2)
    keep = None if inliers is None else np.ascontiguousarray(np.asarray(inliers).astype(np.uint8))
    n = C.c_int64()
    _check(load().eg3d_sfm_save(str(path).encode(), str(original_path).encode(), len(off) - 1, A.ptr(xyz, A.c_f32p), A.ptr(off, A.c_i64p),
                                A.ptr(ov, A.c_i32p), A.ptr(xy, A.c_f32p), A.ptr(keep, A.c_u8p) if keep is not None else None, C.byref(n)))
    return int(n.value)


def write_ply(path, xyz, rgb=None):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    _check(load().eg3d_write_ply(str(path).encode(), len(xyz), A.ptr(xyz, A.c_f32p), A.ptr(rgb, A.c_u8p) if rgb is not None else None))


def polyline_sets_from_refpoints(scene, find_within_dist=10.0, mult=3.0):
    """f2 (host C++): polyline_matching_closeness_to_refpoints (polyline_matcher.cpp:75-168), the producer of pipeline 2's
    candidate sets.  -> (CandidateSets, ids of the contributing SfM points)."""
    from .scene import CandidateSets
    L = load()
    d = scene.desc()
    h = C.c_void_p()
    _check(L.eg3d_polyline_sets_from_refpoints(C.byref(d), find_within_dist, mult, C.byref(h)))
    try:
        c = A.Candidates(); n = C.c_int64(); rp = A.c_i64p()
        _check(L.eg3d_polyline_sets_get(h, C.byref(c), C.byref(n), C.byref(rp)))
        n_off = int(c.n_sets) * scene.n_views + 1
        off = np.ctypeslib.as_array(c.off, shape=(n_off,)).copy()
        ids = np.ctypeslib.as_array(c.polyline, shape=(int(off[-1]),)).copy() if off[-1] > 0 else np.zeros(0, np.uint32)
        ref = np.ctypeslib.as_array(rp, shape=(int(n.value),)).copy() if n.value > 0 else np.zeros(0, np.int64)
        return CandidateSets(int(c.n_sets), off, ids), ref
    finally:
        L.eg3d_polyline_sets_free(h)


def _candidate_sets_from_handle(h, n_views):
    from .scene import CandidateSets
    L = load()
    c = A.Candidates(); n = C.c_int64(); rp = A.c_i64p()
    _check(L.eg3d_polyline_sets_get(h, C.byref(c), C.byref(n), C.byref(rp)))
    off = np.ctypeslib.as_array(c.off, shape=(int(c.n_sets) * n_views + 1,)).copy()
    ids = np.ctypeslib.as_array(c.polyline, shape=(int(off[-1]),)).copy() if off[-1] > 0 else np.zeros(0, np.uint32)
    return CandidateSets(int(c.n_sets), off, ids)


class SimilarityGraph:
    """f2, pipeline 1's producer (host C++): the weighted (view, polyline) compatibility graph of
    polyline_matching_similarity_graph (polyline_matcher.cpp:222-336) + its communities + the candidate sets they give."""

    def __init__(self, scene, find_within_dist=10.0):
        L = load()
        self.n_views = scene.n_views
        d = scene.desc()
        self.h = C.c_void_p()
        _check(L.eg3d_polyline_similarity_graph(C.byref(d), find_within_dist, C.byref(self.h)))
        v = A.SimilarityGraphView()
        _check(L.eg3d_similarity_graph_get(self.h, C.byref(v)))
        n, m = int(v.n_nodes), int(v.n_edges)
        self.node_view = np.ctypeslib.as_array(v.node_view, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        self.node_polyline = np.ctypeslib.as_array(v.node_polyline, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        self.edge_a = np.ctypeslib.as_array(v.edge_a, shape=(m,)).copy() if m else np.zeros(0, np.int64)
        self.edge_b = np.ctypeslib.as_array(v.edge_b, shape=(m,)).copy() if m else np.zeros(0, np.int64)
        self.edge_weight = np.ctypeslib.as_array(v.edge_weight, shape=(m,)).copy() if m else np.zeros(0, np.float32)
        self.dimacs = C.string_at(v.dimacs, int(v.dimacs_len)).decode()

    def communities(self):
        """The library's deterministic sequential Louvain -> (community id per node, modularity)."""
        com = np.zeros(len(self.node_view), np.int64)
        q = C.c_double()
        _check(load().eg3d_similarity_graph_communities(self.h, A.ptr(com, A.c_i64p), C.cast(C.byref(q), A.c_f64p)))
        return com, q.value

    def candidate_sets(self, community):
        L = load()
        com = np.ascontiguousarray(community, np.int64)
        assert len(com) == len(self.node_view)
        h = C.c_void_p()
        _check(L.eg3d_polyline_sets_from_communities(self.h, A.ptr(com, A.c_i64p), C.byref(h)))
        try:
            return _candidate_sets_from_handle(h, self.n_views)
        finally:
            L.eg3d_polyline_sets_free(h)

    def close(self):
        if self.h:
            load().eg3d_similarity_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 calls this and ships it to the other ranks)."""
    uid = np.zeros(128, np.uint8)
    _check(load().eg3d_comm_unique_id(A.ptr(uid, A.c_u8p)))
    return uid.tobytes()


def sample_seeds(scene, views, polylines, spacing):
    """a3 seed sampler (host C++): polyline_matching.cpp:168-190."""
    L = load()
    d = scene.desc()
    views = np.ascontiguousarray(views, np.int32)
    polylines = np.ascontiguousarray(polylines, np.uint32)
    cap = 1 << 16
    while True:
        ov = np.zeros(cap, np.int32); op = np.zeros(cap, np.uint32); os_ = np.zeros(cap, np.uint32)
        oxy = np.zeros((cap, 2), np.float32); osrc = np.zeros(cap, np.int32)
        n = C.c_int64()
        st = L.eg3d_sample_seeds(C.byref(d), A.ptr(views, A.c_i32p), A.ptr(polylines, A.c_u32p), len(views), spacing, cap,
                                 A.ptr(ov, A.c_i32p), A.ptr(op, A.c_u32p), A.ptr(os_, A.c_u32p), A.ptr(oxy, A.c_f32p),
                                 A.ptr(osrc, A.c_i32p), C.byref(n))
        if st not in (A.EG3D_OK, A.EG3D_ERR_CAPACITY):      # capacity = "call again with a larger buffer" (n holds the count)
            _check(st)
        if n.value <= cap:
            k = int(n.value)
            return ov[:k].copy(), op[:k].copy(), os_[:k].copy(), oxy[:k].copy(), osrc[:k].copy()
        cap = int(n.value)


class DevicePoints:
    """Handle to a device-resident result (eg3d_points)."""

    def __init__(self, handle):
        self.h = handle

    def info(self):
        v = A.PointsView()
        _check(load().eg3d_points_device_get(self.h, C.byref(v)))
        return int(v.n_points), int(v.n_obs)

    def device_view(self):
        """eg3d_points_view whose pointers are DEVICE pointers (for device-side consumers, e.g. the all-gather)."""
        v = A.PointsView()
        _check(load().eg3d_points_device_get(self.h, C.byref(v)))
        return v

    def fetch(self):
        """Device -> page-locked host -> numpy copies."""
        v = A.PointsView()
        _check(load().eg3d_points_get(self.h, C.byref(v)))
        return PointSet.from_view(v)

    def fetch_raw(self):
        """D2H into the handle's pinned buffers only (what a C caller of eg3d_points_get pays); returns bytes moved."""
        v = A.PointsView()
        _check(load().eg3d_points_get(self.h, C.byref(v)))
        n, m = int(v.n_points), int(v.n_obs)
        return n * (12 + 4 + 4) + (n + 1) * 8 + m * (4 + 4 + 4 + 8)

    def free(self):
        if self.h:
            load().eg3d_points_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceScene:
    """Device-resident scene handle (eg3d_scene)."""

    def __init__(self, scene, params=None):
        L = load()
        self.scene = scene
        self.params = params if params is not None else default_params()
        self._desc = scene.desc()
        h = C.c_void_p()
        _check(L.eg3d_scene_create(C.byref(self._desc), C.byref(self.params), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            load().eg3d_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _points(self, handle, fetch=True):
        dp = DevicePoints(handle)
        if not fetch:
            return dp
        ps = dp.fetch()
        dp.free()
        return ps

    def epipolar_intersect(self, seeds, cands=None):
        L = load()
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        h = C.c_void_p()
        tm = A.Timing()
        _check(L.eg3d_epipolar_intersect(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, C.byref(h), C.byref(tm)))
        n = C.c_int64(); V = C.c_int32(); off = A.c_i64p(); hits = C.POINTER(A.Hit)()
        _check(L.eg3d_hits_get(h, C.byref(n), C.byref(V), C.byref(off), C.byref(hits)))
        cnt = int(n.value) * int(V.value) + 1
        off_np = np.ctypeslib.as_array(off, shape=(cnt,)).copy()
        nh = int(off_np[-1])
        dt = np.dtype([("polyline", np.uint32), ("segment", np.uint32), ("x", np.float32), ("y", np.float32)])
        if nh:
            buf = (C.c_char * (nh * 16)).from_address(C.addressof(hits.contents))
            hn = np.frombuffer(buf, dtype=dt).copy()
        else:
            hn = np.zeros(0, dt)
        L.eg3d_hits_free(h)
        return off_np, hn, int(V.value), tm.as_dict()

    def epipolar_intersect_device(self, seeds, cands=None):
        """K1 alone, result left on the device and released: returns the timing dict (the sweep microbenchmark)."""
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        tm = A.Timing()
        _check(load().eg3d_epipolar_intersect_device(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, C.byref(tm)))
        return tm.as_dict()

    def match_seeds(self, seeds, cands=None, fetch=True):
        L = load()
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        h = C.c_void_p()
        tm = A.Timing()
        st = L.eg3d_match_seeds(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, C.byref(h), C.byref(tm))
        self.last_timing = tm.as_dict()
        _check(st)
        return self._points(h, fetch), tm.as_dict()

    def match_polyline_sets(self, cands, view_begin=0, view_end=None, fetch=True):
        L = load()
        cd = cands.desc()
        ve = self.scene.n_views if view_end is None else view_end
        h = C.c_void_p()
        tm = A.Timing()
        st = L.eg3d_match_polyline_sets(self.h, C.byref(cd), view_begin, ve, C.byref(h), C.byref(tm))
        self.last_timing = tm.as_dict()
        _check(st)
        return self._points(h, fetch), tm.as_dict()

    def match_refpoints(self, tb=0, te=None, fetch=True):
        L = load()
        te = self.scene.n_tracks if te is None else te
        h = C.c_void_p()
        tm = A.Timing()
        _check(L.eg3d_match_refpoints(self.h, tb, te, C.byref(h), C.byref(tm)))
        return self._points(h, fetch), tm.as_dict()

    def match_correspondences(self, start_view, hit_off, hits, fetch=True):
        """B4: K3 alone on caller-supplied hit lists (CSR over (seed, view); `hits` is the structured array epipolar_intersect returns)."""
        sv = np.ascontiguousarray(start_view, np.int32)
        off = np.ascontiguousarray(hit_off, np.int64)
        hh = np.ascontiguousarray(hits)
        h = C.c_void_p()
        tm = A.Timing()
        st = load().eg3d_match_correspondences(self.h, len(sv), A.ptr(sv, A.c_i32p), A.ptr(off, A.c_i64p),
                                               C.cast(hh.ctypes.data, C.POINTER(A.Hit)), C.byref(h), C.byref(tm))
        self.last_timing = tm.as_dict()
        _check(st)
        return self._points(h, fetch), tm.as_dict()

    def refpoint_correspondences(self, tb=0, te=None):
        """B3 (EdgeManager side): the seeds of pipeline 3 and their per-view hit lists -> (SeedBatch, track per seed, hit_off, hits)."""
        from .scene import SeedBatch
        L = load()
        te = self.scene.n_tracks if te is None else te
        h = C.c_void_p()
        _check(L.eg3d_refpoint_correspondences(self.h, tb, te, C.byref(h)))
        try:
            n = C.c_int64(); V = C.c_int32(); view = A.c_i32p(); pl = A.c_u32p(); seg = A.c_u32p(); xy = A.c_f32p(); tr = A.c_i64p(); off = A.c_i64p(); hits = C.POINTER(A.Hit)()
            _check(L.eg3d_corr_get(h, C.byref(n), C.byref(V), C.byref(view), C.byref(pl), C.byref(seg), C.byref(xy), C.byref(tr), C.byref(off), C.byref(hits)))
            k = int(n.value)
            arr = lambda p, shape, dt: (np.ctypeslib.as_array(p, shape=shape).copy() if k else np.zeros(shape, dt))
            off_np = np.ctypeslib.as_array(off, shape=(k * int(V.value) + 1,)).copy()
            nh = int(off_np[-1])
            dt = np.dtype([("polyline", np.uint32), ("segment", np.uint32), ("x", np.float32), ("y", np.float32)])
            hn = np.frombuffer((C.c_char * (nh * 16)).from_address(C.addressof(hits.contents)), dtype=dt).copy() if nh else np.zeros(0, dt)
            seeds = SeedBatch(arr(view, (k,), np.int32), arr(pl, (k,), np.uint32), arr(seg, (k,), np.uint32), arr(xy, (k, 2), np.float32), None)
            return seeds, arr(tr, (k,), np.int64), off_np, hn
        finally:
            L.eg3d_corr_free(h)

    # --- multi-GPU exchange: the scene handle owns an NCCL communicator
    def comm_create(self, unique_id, rank, world):
        """Collective.  `unique_id`: the 128 bytes rank 0 got from lib.comm_unique_id() (shipped by the host's own channel)."""
        uid = np.frombuffer(bytes(unique_id), np.uint8).copy()
        assert uid.size == 128
        _check(load().eg3d_comm_create(self.h, A.ptr(uid, A.c_u8p), rank, world))
        self.has_comm = True

    def comm_destroy(self):
        _check(load().eg3d_comm_destroy(self.h))
        self.has_comm = False

    def points_allgather(self, mine, seed_global=None, fetch=False):
        """Collective: `mine` (DevicePoints) of every rank -> the merged records of all ranks in (global seed ordinal, chain
        position) order, on every rank.  seed_global: int64 per seed of the call that produced `mine` (None = rank-major)."""
        sg = None if seed_global is None else np.ascontiguousarray(seed_global, np.int64)
        h = C.c_void_p()
        tm = A.Timing()
        _check(load().eg3d_points_allgather(self.h, mine.h, A.ptr(sg, A.c_i64p) if sg is not None else None, C.byref(h), C.byref(tm)))
        return self._points(h, fetch), tm.as_dict()

    def gn_triangulate(self, obs_off, obs_view, obs_xy, init_xyz, fp64):
        L = load()
        n = len(obs_off) - 1
        obs_off = np.ascontiguousarray(obs_off, np.int64)
        obs_view = np.ascontiguousarray(obs_view, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32)
        init_xyz = np.ascontiguousarray(init_xyz, np.float32)
        xyz = np.zeros((n, 3), np.float32); mse = np.zeros(n, np.float32); ok = np.zeros(n, np.uint8)
        tm = A.Timing()
        _check(L.eg3d_gn_triangulate(self.h, n, A.ptr(obs_off, A.c_i64p), A.ptr(obs_view, A.c_i32p), A.ptr(obs_xy, A.c_f32p),
                                     A.ptr(init_xyz, A.c_f32p), int(fp64), A.ptr(xyz, A.c_f32p), A.ptr(mse, A.c_f32p),
                                     A.ptr(ok, A.c_u8p), C.byref(tm)))
        return xyz, mse, ok, tm.as_dict()

    def gn_triangulate_device(self, n, k, d_view, d_xy, d_init, fp64, d_xyz, d_mse, d_ok):
        """Device pointers (ints) on this scene's device; returns the timing dict."""
        tm = A.Timing()
        _check(load().eg3d_gn_triangulate_device(self.h, n, k, d_view, d_xy, d_init, int(fp64), d_xyz, d_mse, d_ok, C.byref(tm)))
        return tm.as_dict()

    def dedup_close_points(self, pts):
        keep = np.zeros(pts.n_points, np.uint8)
        v = pts.view_struct()
        _check(load().eg3d_dedup_close_points(self.h, C.byref(v), A.ptr(keep, A.c_u8p)))
        return keep

    def filter(self, xyz, obs_off, obs_view, obs_xy, first_edgepoint, gn_max_mse=2.25, forced_min_filter=-1):
        xyz = np.ascontiguousarray(xyz, np.float32).copy()
        obs_off = np.ascontiguousarray(obs_off, np.int64)
        obs_view = np.ascontiguousarray(obs_view, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32)
        n = len(obs_off) - 1
        inl = np.zeros(n, np.uint8)
        tm = A.Timing()
        _check(load().eg3d_filter(self.h, n, A.ptr(xyz, A.c_f32p), A.ptr(obs_off, A.c_i64p), A.ptr(obs_view, A.c_i32p),
                                  A.ptr(obs_xy, A.c_f32p), first_edgepoint, gn_max_mse, forced_min_filter, A.ptr(inl, A.c_u8p), C.byref(tm)))
        return xyz, inl, tm.as_dict()
