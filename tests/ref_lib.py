"""ctypes binding of oracle/_ref/libref_path.so — the REFERENCE'S OWN hot-path code compiled unmodified from /root/reference
against stand-in headers (oracle/ref_path_wrapper.cpp, oracle/Makefile).  Test infrastructure: it checks the oracle, and only
tests import it.  The library is built in the build container (the reference tree does not exist on the GPU box; the prebuilt
.so travels there like every other built artefact)."""
import ctypes as C
import os
import numpy as np
from edgegraph3d_b200 import _abi as A
from edgegraph3d_b200.scene import PointSet

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_path.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.eg3d_ref_scene_create.restype = C.c_void_p
        L.eg3d_ref_scene_create.argtypes = [C.POINTER(A.SceneDesc)]
        L.eg3d_ref_last_error.restype = C.c_char_p
        L.eg3d_ref_last_error.argtypes = [C.c_void_p]
        for name in ("eg3d_ref_match_polyline_sets", "eg3d_ref_match_seeds", "eg3d_ref_match_refpoints", "eg3d_ref_epipolar_intersect"):
            getattr(L, name).restype = C.c_void_p
        L.eg3d_ref_match_polyline_sets.argtypes = [C.c_void_p, C.POINTER(A.Candidates)]
        L.eg3d_ref_match_seeds.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates)]
        L.eg3d_ref_match_seeds_mt.restype = C.c_void_p
        L.eg3d_ref_match_seeds_mt.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates), C.c_int]
        L.eg3d_ref_match_refpoints.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.eg3d_ref_epipolar_intersect.argtypes = [C.c_void_p, C.POINTER(A.Seeds), C.POINTER(A.Candidates), A.c_i64p]
        L.eg3d_ref_hits_get.argtypes = [C.c_void_p, C.POINTER(A.c_i64p), C.POINTER(C.POINTER(A.Hit))]
        L.eg3d_ref_points_get.argtypes = [C.c_void_p, C.POINTER(A.PointsView)]
        L.eg3d_ref_points_free.argtypes = [C.c_void_p]
        L.eg3d_ref_dedup_close_points.argtypes = [C.c_void_p, C.POINTER(A.PointsView), A.c_u8p]
        L.eg3d_ref_filter.argtypes = [C.c_void_p, C.c_int64, A.c_f32p, A.c_i64p, A.c_i32p, A.c_f32p, C.c_int64, C.c_float, C.c_int32, A.c_u8p]
        L.eg3d_ref_parse_openmvg.argtypes = [C.c_char_p, C.c_int, A.c_f32p, A.c_i32p, A.c_i32p, A.c_i64p]
        _lib = L
    return _lib


class RefScene:
    """The reference's SfMData + PolyLineGraph2DHMapImpl + Mat** F + plmaps + PLGEdgeManager + PLGPCM3ViewsPLGFollowing built from
    a FlatScene; its methods run the reference's own entry points (default constants only)."""

    def __init__(self, scene):
        self.scene = scene
        self._desc = scene.desc()
        self.h = lib().eg3d_ref_scene_create(C.byref(self._desc))
        err = lib().eg3d_ref_last_error(self.h).decode()
        if err:
            raise RuntimeError(err)

    def _points(self, handle):
        v = A.PointsView()
        lib().eg3d_ref_points_get(handle, C.byref(v))
        ps = PointSet.from_view(v)
        lib().eg3d_ref_points_free(handle)
        return ps

    def match_polyline_sets(self, cands):
        cd = cands.desc()
        return self._points(lib().eg3d_ref_match_polyline_sets(self.h, C.byref(cd)))

    def match_seeds(self, seeds, cands=None, n_threads=1):
        """n_threads > 1: the same per-seed calls from an OpenMP loop over the seeds (the reference's own loop is serial)."""
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        if n_threads > 1:
            return self._points(lib().eg3d_ref_match_seeds_mt(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, n_threads))
        return self._points(lib().eg3d_ref_match_seeds(self.h, C.byref(sd), C.byref(cd) if cd is not None else None))

    def match_refpoints(self, tb=0, te=None):
        te = self.scene.n_tracks if te is None else te
        return self._points(lib().eg3d_ref_match_refpoints(self.h, tb, te))

    def epipolar_intersect(self, seeds, cands=None):
        sd = seeds.desc()
        cd = cands.desc() if cands is not None else None
        n = C.c_int64()
        h = lib().eg3d_ref_epipolar_intersect(self.h, C.byref(sd), C.byref(cd) if cd is not None else None, C.byref(n))
        off = A.c_i64p(); hits = C.POINTER(A.Hit)()
        lib().eg3d_ref_hits_get(h, C.byref(off), C.byref(hits))
        cnt = len(seeds) * self.scene.n_views + 1
        off_np = np.ctypeslib.as_array(off, shape=(cnt,)).copy()
        dt = np.dtype([("polyline", np.uint32), ("segment", np.uint32), ("x", np.float32), ("y", np.float32)])
        nh = int(n.value)
        hn = np.frombuffer((C.c_char * (nh * 16)).from_address(C.addressof(hits.contents)), dtype=dt).copy() if nh else np.zeros(0, dt)
        return off_np, hn

    def dedup_close_points(self, pts):
        keep = np.zeros(pts.n_points, np.uint8)
        v = pts.view_struct()
        lib().eg3d_ref_dedup_close_points(self.h, C.byref(v), A.ptr(keep, A.c_u8p))
        return keep

    def filter(self, xyz, obs_off, obs_view, obs_xy, first_edgepoint, gn_max_mse=2.25, forced_min=-1):
        xyz = np.ascontiguousarray(xyz, np.float32).copy()
        obs_off = np.ascontiguousarray(obs_off, np.int64); obs_view = np.ascontiguousarray(obs_view, np.int32)
        obs_xy = np.ascontiguousarray(obs_xy, np.float32)
        n = len(obs_off) - 1
        inl = np.zeros(n, np.uint8)
        lib().eg3d_ref_filter(self.h, n, A.ptr(xyz, A.c_f32p), A.ptr(obs_off, A.c_i64p), A.ptr(obs_view, A.c_i32p), A.ptr(obs_xy, A.c_f32p),
                              first_edgepoint, gn_max_mse, forced_min, A.ptr(inl, A.c_u8p))
        return xyz, inl


def parse_openmvg(path, max_views=4096):
    cams = np.zeros((max_views, 12), np.float32)
    w = C.c_int32(); h = C.c_int32(); npts = C.c_int64()
    V = lib().eg3d_ref_parse_openmvg(path.encode(), max_views, A.ptr(cams, A.c_f32p), C.byref(w), C.byref(h), C.byref(npts))
    return cams[:V].copy(), int(w.value), int(h.value), int(npts.value)
