"""ctypes wrapper of the CPU oracle (oracle/libeg3d_oracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np
from edgegraph3d_b200 import _abi as A
from edgegraph3d_b200.scene import PointSet

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "libeg3d_oracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_ROOT, "oracle", f) for f in os.listdir(os.path.join(_ROOT, "oracle")) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(_ROOT, "include", "eg3d.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle")], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.eg3d_oracle_scene_create.restype = C.c_void_p
        L.eg3d_oracle_scene_create.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Params)]
        L.eg3d_oracle_scene_destroy.argtypes = [C.c_void_p]
        for name in ("eg3d_oracle_epipolar_intersect", "eg3d_oracle_match_seeds"):
            f = getattr(L, name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates), C.c_int]
        L.eg3d_oracle_match_polyline_sets.restype = C.c_void_p
        L.eg3d_oracle_match_polyline_sets.argtypes = [C.c_void_p, C.POINTER(A.Candidates), C.c_int32, C.c_int32, C.c_int]
        L.eg3d_oracle_match_refpoints.restype = C.c_void_p
        L.eg3d_oracle_match_refpoints.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        L.eg3d_oracle_dlt_stats.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
        L.eg3d_oracle_refpoint_hits.restype = C.c_void_p
        L.eg3d_oracle_refpoint_hits.argtypes = [C.c_void_p, C.c_int64, C.c_int64, A.c_i64p]
        L.eg3d_oracle_hits_get.argtypes = [C.c_void_p, A.c_i64p, A.c_i32p, C.POINTER(A.c_i64p), C.POINTER(C.POINTER(A.Hit))]
        L.eg3d_oracle_hits_free.argtypes = [C.c_void_p]
        L.eg3d_oracle_points_get.argtypes = [C.c_void_p, C.POINTER(A.PointsView)]
        L.eg3d_oracle_points_free.argtypes = [C.c_void_p]
        L.eg3d_oracle_points_seed_info.restype = C.c_int64
        L.eg3d_oracle_points_seed_info.argtypes = [C.c_void_p, C.POINTER(A.c_u8p), C.POINTER(A.c_i64p), C.POINTER(A.c_i32p)]
        L.eg3d_oracle_gn_triangulate.argtypes = [C.c_void_p, C.c_int64, A.c_i64p, A.c_i32p, A.c_f32p, A.c_f32p, C.c_int,
                                                 A.c_f32p, A.c_f32p, A.c_u8p, C.c_int]
        L.eg3d_oracle_dedup_close_points.argtypes = [C.c_void_p, C.POINTER(A.PointsView), A.c_u8p]
        L.eg3d_oracle_filter.argtypes = [C.c_void_p, C.c_int64, A.c_f32p, A.c_i64p, A.c_i32p, A.c_f32p, C.c_int64, C.c_float,
                                         C.c_int32, A.c_u8p, C.c_int]
        L.eg3d_oracle_sample_seeds.argtypes = [C.POINTER(A.SceneDesc), A.c_i32p, A.c_u32p, C.c_int64, C.c_float, C.c_int64,
                                               A.c_i32p, A.c_u32p, A.c_u32p, A.c_f32p, A.c_i32p, A.c_i64p]
        L.eg3d_oracle_epiline.argtypes = [A.c_f64p, C.c_float, C.c_float, A.c_f32p]
        L.eg3d_oracle_triangulate_dlt.argtypes = [A.c_f32p] * 5
        L.eg3d_oracle_triangulate_dlt_opencv.argtypes = [A.c_f32p] * 5
        L.eg3d_oracle_compute_projection.argtypes = [A.c_f32p] * 3
        L.eg3d_oracle_intersect_segment_line.argtypes = [A.c_f32p] * 3
        L.eg3d_oracle_intersect_segment_line_nqp.argtypes = [A.c_f32p, A.c_f32p, C.c_float, C.c_float, A.c_f32p]
        L.eg3d_oracle_squared_2d_distance.restype = C.c_float
        L.eg3d_oracle_squared_2d_distance.argtypes = [C.c_float] * 4
        L.eg3d_oracle_grid_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, A.c_u32p, C.c_int]
        L.eg3d_oracle_walk.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_int,
                                       A.c_f32p, C.c_float, A.c_u32p, A.c_f32p]
        L.eg3d_oracle_polyline_distancesq.restype = C.c_float
        L.eg3d_oracle_polyline_distancesq.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_float, A.c_u32p, A.c_f32p]
        L.eg3d_oracle_params_default.argtypes = [C.POINTER(A.Params)]
        _lib = L
    return _lib


def default_params(**overrides):
    p = A.Params()
    lib().eg3d_oracle_params_default(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def hits_to_numpy(get, handle):
    n = C.c_int64()
    V = C.c_int32()
    off = A.c_i64p()
    hits = C.POINTER(A.Hit)()
    get(handle, C.byref(n), C.byref(V), C.byref(off), C.byref(hits))
    cnt = int(n.value) * int(V.value) + 1
    off_np = np.ctypeslib.as_array(off, shape=(cnt,)).copy()
    nh = int(off_np[-1])
    dt = np.dtype([("polyline", np.uint32), ("segment", np.uint32), ("x", np.float32), ("y", np.float32)])
    if nh:
        buf = (C.c_char * (nh * 16)).from_address(C.addressof(hits.contents))
        h = np.frombuffer(buf, dtype=dt).copy()
    else:
        h = np.zeros(0, dt)
    return off_np.reshape(-1), h, int(V.value)


class OracleScene:
    def __init__(self, scene, params=None):
        self.scene = scene
        self.params = params if params is not None else default_params()
        self._desc = scene.desc()
        self.h = lib().eg3d_oracle_scene_create(C.byref(self._desc), C.byref(self.params))

    def close(self):
        if self.h:
            lib().eg3d_oracle_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _points(self, handle):
        v = A.PointsView()
        lib().eg3d_oracle_points_get(handle, C.byref(v))
        ps = PointSet.from_view(v)
        ub, tr, cs = A.c_u8p(), A.c_i64p(), A.c_i32p()
        n = lib().eg3d_oracle_points_seed_info(handle, C.byref(ub), C.byref(tr), C.byref(cs))
        # per seed of the call: undefined behaviour in the reference, SfM point id, candidate set
        ps.seed_ub = np.ctypeslib.as_array(ub, shape=(n,)).copy() if n else np.zeros(0, np.uint8)
        ps.seed_track = np.ctypeslib.as_array(tr, shape=(n,)).copy() if n else np.zeros(0, np.int64)
        ps.seed_set = np.ctypeslib.as_array(cs, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        lib().eg3d_oracle_points_free(handle)
        return ps

    def epipolar_intersect(self, seeds, cands=None, n_threads=8):
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        h = lib().eg3d_oracle_epipolar_intersect(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, n_threads)
        out = hits_to_numpy(lib().eg3d_oracle_hits_get, h)
        lib().eg3d_oracle_hits_free(h)
        return out

    def match_seeds(self, seeds, cands=None, n_threads=8):
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        return self._points(lib().eg3d_oracle_match_seeds(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, n_threads))

    def match_polyline_sets(self, cands, view_begin=0, view_end=None, n_threads=8):
        cd = cands.desc()
        ve = self.scene.n_views if view_end is None else view_end
        return self._points(lib().eg3d_oracle_match_polyline_sets(self.h, C.byref(cd), view_begin, ve, n_threads))

    def match_refpoints(self, tb=0, te=None, n_threads=8):
        te = self.scene.n_tracks if te is None else te
        return self._points(lib().eg3d_oracle_match_refpoints(self.h, tb, te, n_threads))

    def refpoint_hits(self, tb=0, te=None):
        te = self.scene.n_tracks if te is None else te
        n = C.c_int64()
        h = lib().eg3d_oracle_refpoint_hits(self.h, tb, te, C.byref(n))
        out = hits_to_numpy(lib().eg3d_oracle_hits_get, h)
        lib().eg3d_oracle_hits_free(h)
        return out

    def gn_triangulate(self, obs_off, obs_view, obs_xy, init_xyz, fp64, n_threads=8):
        n = len(obs_off) - 1
        obs_off = np.ascontiguousarray(obs_off, np.int64)
        obs_view = np.ascontiguousarray(obs_view, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32)
        init_xyz = np.ascontiguousarray(init_xyz, np.float32)
        xyz = np.zeros((n, 3), np.float32)
        mse = np.zeros(n, np.float32)
        ok = np.zeros(n, np.uint8)
        lib().eg3d_oracle_gn_triangulate(self.h, n, A.ptr(obs_off, A.c_i64p), A.ptr(obs_view, A.c_i32p), A.ptr(obs_xy, A.c_f32p),
                                         A.ptr(init_xyz, A.c_f32p), int(fp64), A.ptr(xyz, A.c_f32p), A.ptr(mse, A.c_f32p),
                                         A.ptr(ok, A.c_u8p), n_threads)
        return xyz, mse, ok

    def dedup_close_points(self, pts):
        keep = np.zeros(pts.n_points, np.uint8)
        v = pts.view_struct()
        lib().eg3d_oracle_dedup_close_points(self.h, C.byref(v), A.ptr(keep, A.c_u8p))
        return keep

    def filter(self, xyz, obs_off, obs_view, obs_xy, first_edgepoint, gn_max_mse=2.25, forced_min_filter=-1, n_threads=8):
        xyz = np.ascontiguousarray(xyz, np.float32).copy()
        obs_off = np.ascontiguousarray(obs_off, np.int64)
        obs_view = np.ascontiguousarray(obs_view, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32)
        n = len(obs_off) - 1
        inl = np.zeros(n, np.uint8)
        lib().eg3d_oracle_filter(self.h, n, A.ptr(xyz, A.c_f32p), A.ptr(obs_off, A.c_i64p), A.ptr(obs_view, A.c_i32p),
                                 A.ptr(obs_xy, A.c_f32p), first_edgepoint, gn_max_mse, forced_min_filter, A.ptr(inl, A.c_u8p), n_threads)
        return xyz, inl

    def grid_query(self, view, which, x, y, cap=64):
        out = np.zeros(cap, np.uint32)
        n = lib().eg3d_oracle_grid_query(self.h, view, which, x, y, A.ptr(out, A.c_u32p), cap)
        return out[:min(n, cap)].tolist()


def dlt_stats(reset=True):
    """(fresh 2-view DLT calls, calls that got the same camera twice) since the last reset — statistics of the oracle."""
    a, b = C.c_longlong(), C.c_longlong()
    lib().eg3d_oracle_dlt_stats(C.byref(a), C.byref(b), int(reset))
    return a.value, b.value


class OracleDevice:
    """Tests-only stand-in for edgegraph3d_b200.lib.DeviceScene: the same method names and return shapes, answered by the
    CPU oracle, so that host-side drivers (edgegraph3d_b200/pipeline.py, the multi-rank merge) can be exercised without a GPU
    and so that a GPU run can be compared with the oracle call for call."""

    def __init__(self, scene, params=None, n_threads=4):
        self.scene, self.params, self.n_threads = scene, params, n_threads
        self.osc = OracleScene(scene, params)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.osc.close()

    def match_polyline_sets(self, cands, view_begin=0, view_end=None):
        return self.osc.match_polyline_sets(cands, view_begin, view_end, n_threads=self.n_threads), {"total_ms": 0.0, "n_seeds": 0}

    def match_refpoints(self, tb=0, te=None):
        return self.osc.match_refpoints(tb, te, n_threads=self.n_threads), {"total_ms": 0.0, "n_seeds": 0}

    def dedup_close_points(self, pts):
        return self.osc.dedup_close_points(pts)

    def filter(self, xyz, obs_off, obs_view, obs_xy, first_edgepoint):
        fx, inl = self.osc.filter(xyz, obs_off, obs_view, obs_xy, first_edgepoint, n_threads=self.n_threads)[:2]
        return fx, inl, {"gn_ms": 0.0}


def sample_seeds(scene, views, polylines, spacing):
    """a3 seed sampler through the oracle (same contract as eg3d_sample_seeds)."""
    d = scene.desc()
    views = np.ascontiguousarray(views, np.int32)
    polylines = np.ascontiguousarray(polylines, np.uint32)
    cap = 1 << 16
    while True:
        ov = np.zeros(cap, np.int32); op = np.zeros(cap, np.uint32); os_ = np.zeros(cap, np.uint32)
        oxy = np.zeros((cap, 2), np.float32); osrc = np.zeros(cap, np.int32)
        n = C.c_int64()
        lib().eg3d_oracle_sample_seeds(C.byref(d), A.ptr(views, A.c_i32p), A.ptr(polylines, A.c_u32p), len(views), spacing, cap,
                                       A.ptr(ov, A.c_i32p), A.ptr(op, A.c_u32p), A.ptr(os_, A.c_u32p), A.ptr(oxy, A.c_f32p),
                                       A.ptr(osrc, A.c_i32p), C.byref(n))
        if n.value <= cap:
            k = int(n.value)
            return ov[:k].copy(), op[:k].copy(), os_[:k].copy(), oxy[:k].copy(), osrc[:k].copy()
        cap = int(n.value)
