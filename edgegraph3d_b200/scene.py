"""Flat host-side containers for the data the hot path reads (numpy arrays laid out as include/eg3d.h expects).

FlatScene is the flattened form of the reference's `SfMData` + `vector<PolyLineGraph2DHMapImpl>` +
`Mat** all_fundamental_matrices`.
"""
from dataclasses import dataclass, field
from typing import Optional
import numpy as np
from . import _abi as A


@dataclass
class FlatScene:
    width: int
    height: int
    cameras: np.ndarray            # [V,12] f32
    fundamental: np.ndarray        # [V,V,9] f64
    fundamental_valid: np.ndarray  # [V,V] u8
    view_poly_off: np.ndarray      # [V+1] i64
    poly_vert_off: np.ndarray      # [NP+1] i64
    verts: np.ndarray              # [NVERT,2] f32
    poly_start: np.ndarray         # [NP] u32
    poly_end: np.ndarray           # [NP] u32
    track_xyz: Optional[np.ndarray] = None   # [NT,3] f32
    track_off: Optional[np.ndarray] = None   # [NT+1] i64
    track_view: Optional[np.ndarray] = None  # [NOBS] i32
    track_xy: Optional[np.ndarray] = None    # [NOBS,2] f32
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        c = np.ascontiguousarray
        self.cameras = c(self.cameras, np.float32).reshape(-1, 12)
        V = self.cameras.shape[0]
        self.fundamental = c(self.fundamental, np.float64).reshape(V, V, 9)
        self.fundamental_valid = c(self.fundamental_valid, np.uint8).reshape(V, V)
        self.view_poly_off = c(self.view_poly_off, np.int64)
        self.poly_vert_off = c(self.poly_vert_off, np.int64)
        self.verts = c(self.verts, np.float32).reshape(-1, 2)
        self.poly_start = c(self.poly_start, np.uint32)
        self.poly_end = c(self.poly_end, np.uint32)
        if self.track_xyz is not None:
            self.track_xyz = c(self.track_xyz, np.float32).reshape(-1, 3)
            self.track_off = c(self.track_off, np.int64)
            self.track_view = c(self.track_view, np.int32)
            self.track_xy = c(self.track_xy, np.float32).reshape(-1, 2)

    @property
    def n_views(self):
        return self.cameras.shape[0]

    @property
    def n_tracks(self):
        return 0 if self.track_xyz is None else self.track_xyz.shape[0]

    def n_polylines(self, view):
        return int(self.view_poly_off[view + 1] - self.view_poly_off[view])

    def polyline(self, view, pl):
        g = int(self.view_poly_off[view]) + int(pl)
        return self.verts[int(self.poly_vert_off[g]):int(self.poly_vert_off[g + 1])]

    def n_segments(self, view):
        g0, g1 = int(self.view_poly_off[view]), int(self.view_poly_off[view + 1])
        nv = np.diff(self.poly_vert_off[g0:g1 + 1])
        return int(np.maximum(nv - 1, 0).sum())

    def input_nbytes(self):
        n = sum(a.nbytes for a in (self.cameras, self.fundamental, self.fundamental_valid, self.view_poly_off,
                                   self.poly_vert_off, self.verts, self.poly_start, self.poly_end))
        if self.track_xyz is not None:
            n += self.track_xyz.nbytes + self.track_off.nbytes + self.track_view.nbytes + self.track_xy.nbytes
        return n

    def desc(self):
        d = A.SceneDesc()
        d.n_views, d.width, d.height = self.n_views, int(self.width), int(self.height)
        d.cameras = A.ptr(self.cameras, A.c_f32p)
        d.fundamental = A.ptr(self.fundamental, A.c_f64p)
        d.fundamental_valid = A.ptr(self.fundamental_valid, A.c_u8p)
        d.view_poly_off = A.ptr(self.view_poly_off, A.c_i64p)
        d.poly_vert_off = A.ptr(self.poly_vert_off, A.c_i64p)
        d.verts = A.ptr(self.verts, A.c_f32p)
        d.poly_start = A.ptr(self.poly_start, A.c_u32p)
        d.poly_end = A.ptr(self.poly_end, A.c_u32p)
        d.n_tracks = self.n_tracks
        d.track_xyz = A.ptr(self.track_xyz, A.c_f32p)
        d.track_off = A.ptr(self.track_off, A.c_i64p)
        d.track_view = A.ptr(self.track_view, A.c_i32p)
        d.track_xy = A.ptr(self.track_xy, A.c_f32p)
        return d


@dataclass
class SeedBatch:
    """(view, plg_point) seeds — what polyline_matching.cpp:168-190 walks out of each candidate polyline."""
    view: np.ndarray      # i32
    polyline: np.ndarray  # u32
    segment: np.ndarray   # u32
    xy: np.ndarray        # [n,2] f32
    cand_set: Optional[np.ndarray] = None  # i32, -1 = sweep

    def __post_init__(self):
        c = np.ascontiguousarray
        self.view = c(self.view, np.int32)
        self.polyline = c(self.polyline, np.uint32)
        self.segment = c(self.segment, np.uint32)
        self.xy = c(self.xy, np.float32).reshape(-1, 2)
        if self.cand_set is not None:
            self.cand_set = c(self.cand_set, np.int32)

    def __len__(self):
        return int(self.view.shape[0])

    def nbytes(self):
        return self.view.nbytes + self.polyline.nbytes + self.segment.nbytes + self.xy.nbytes + \
            (0 if self.cand_set is None else self.cand_set.nbytes)

    def slice(self, lo, hi):
        return SeedBatch(self.view[lo:hi], self.polyline[lo:hi], self.segment[lo:hi], self.xy[lo:hi],
                         None if self.cand_set is None else self.cand_set[lo:hi])

    def take(self, idx):
        return SeedBatch(self.view[idx], self.polyline[idx], self.segment[idx], self.xy[idx],
                         None if self.cand_set is None else self.cand_set[idx])

    def desc(self):
        s = A.Seeds()
        s.n = len(self)
        s.view = A.ptr(self.view, A.c_i32p)
        s.polyline = A.ptr(self.polyline, A.c_u32p)
        s.segment = A.ptr(self.segment, A.c_u32p)
        s.xy = A.ptr(self.xy, A.c_f32p)
        s.cand_set = A.ptr(self.cand_set, A.c_i32p)
        return s


@dataclass
class CandidateSets:
    """`vector<set<ulong>> potentially_compatible_polylines`, one per polyline match (pipelines.cpp:92-100)."""
    n_sets: int
    off: np.ndarray       # [n_sets*V+1] i64
    polyline: np.ndarray  # u32

    def __post_init__(self):
        self.off = np.ascontiguousarray(self.off, np.int64)
        self.polyline = np.ascontiguousarray(self.polyline, np.uint32)

    @staticmethod
    def from_lists(sets, n_views):
        """sets: list (per match) of list (per view) of iterables of polyline ids."""
        off, ids = [0], []
        for s in sets:
            assert len(s) == n_views
            for v in range(n_views):
                ids.extend(sorted(set(int(x) for x in s[v])))
                off.append(len(ids))
        return CandidateSets(len(sets), np.array(off, np.int64), np.array(ids, np.uint32))

    def desc(self):
        c = A.Candidates()
        c.n_sets = int(self.n_sets)
        c.off = A.ptr(self.off, A.c_i64p)
        c.polyline = A.ptr(self.polyline, A.c_u32p)
        return c


def ranges(starts, lens):
    """Concatenation of arange(starts[i], starts[i] + lens[i]) without a Python loop."""
    starts, lens = np.asarray(starts, np.int64), np.asarray(lens, np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    first = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return np.repeat(starts - first, lens) + np.arange(total, dtype=np.int64)


@dataclass
class PointSet:
    """Flattened `vector<new_3dpoint_plgp_matches>` (polyline_graph_2d.hpp:451)."""
    xyz: np.ndarray
    seed: np.ndarray
    chain_pos: np.ndarray
    obs_off: np.ndarray
    obs_view: np.ndarray
    obs_poly: np.ndarray
    obs_seg: np.ndarray
    obs_xy: np.ndarray

    @property
    def n_points(self):
        return int(self.seed.shape[0])

    @property
    def n_obs(self):
        return int(self.obs_view.shape[0])

    def nbytes(self):
        return sum(a.nbytes for a in (self.xyz, self.seed, self.chain_pos, self.obs_off, self.obs_view,
                                      self.obs_poly, self.obs_seg, self.obs_xy))

    @staticmethod
    def from_view(v):
        """Copy out of a C eg3d_points_view (library-owned memory)."""
        n, m = int(v.n_points), int(v.n_obs)

        def arr(p, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(p, shape=(count,)).astype(dt, copy=True)
        return PointSet(arr(v.xyz, 3 * n, np.float32).reshape(n, 3), arr(v.seed, n, np.int32),
                        arr(v.chain_pos, n, np.int32), arr(v.obs_off, n + 1, np.int64) if n else np.zeros(1, np.int64),
                        arr(v.obs_view, m, np.int32), arr(v.obs_poly, m, np.uint32), arr(v.obs_seg, m, np.uint32),
                        arr(v.obs_xy, 2 * m, np.float32).reshape(m, 2))

    def view_struct(self):
        """eg3d_points_view over this set's arrays.  Every field is coerced to the dtype / layout the C struct declares (a merged
        `seed` can be int64: global ordinals; then it is narrowed, which is harmless — the C side never indexes by it) and the
        coerced arrays are kept alive on the view object for as long as the caller holds it."""
        v = A.PointsView()
        v.n_points, v.n_obs = self.n_points, self.n_obs
        keep = []

        def arr(a, dt):
            b = np.ascontiguousarray(a, dt)
            keep.append(b)
            return b
        v.xyz = A.ptr(arr(self.xyz, np.float32), A.c_f32p)
        v.seed = A.ptr(arr(self.seed, np.int32), A.c_i32p)
        v.chain_pos = A.ptr(arr(self.chain_pos, np.int32), A.c_i32p)
        v.obs_off = A.ptr(arr(self.obs_off, np.int64), A.c_i64p)
        v.obs_view = A.ptr(arr(self.obs_view, np.int32), A.c_i32p)
        v.obs_poly = A.ptr(arr(self.obs_poly, np.uint32), A.c_u32p)
        v.obs_seg = A.ptr(arr(self.obs_seg, np.uint32), A.c_u32p)
        v.obs_xy = A.ptr(arr(self.obs_xy, np.float32), A.c_f32p)
        v._keep = keep
        return v

    def identity_keys(self):
        """Per point: (seed, chain_pos, ((view, polyline, segment), ...)) — the parity key of SURVEY §8(d)."""
        keys = []
        for i in range(self.n_points):
            a, b = int(self.obs_off[i]), int(self.obs_off[i + 1])
            keys.append((int(self.seed[i]), int(self.chain_pos[i]),
                         tuple(zip(self.obs_view[a:b].tolist(), self.obs_poly[a:b].tolist(), self.obs_seg[a:b].tolist()))))
        return keys

    def take(self, idx):
        """Points `idx` (in that order) with their observation lists."""
        idx = np.asarray(idx, np.int64)
        lens = (self.obs_off[idx + 1] - self.obs_off[idx]).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        o = ranges(self.obs_off[idx], lens)
        return PointSet(self.xyz[idx], self.seed[idx], self.chain_pos[idx], off, self.obs_view[o], self.obs_poly[o], self.obs_seg[o], self.obs_xy[o])

    @staticmethod
    def concat(parts):
        parts = [p for p in parts]
        if not parts:
            z = np.zeros
            return PointSet(z((0, 3), np.float32), z(0, np.int32), z(0, np.int32), z(1, np.int64), z(0, np.int32),
                            z(0, np.uint32), z(0, np.uint32), z((0, 2), np.float32))
        offs, base = [np.zeros(1, np.int64)], 0
        for p in parts:
            offs.append(p.obs_off[1:] + base)
            base += p.n_obs
        return PointSet(np.concatenate([p.xyz for p in parts]), np.concatenate([p.seed for p in parts]),
                        np.concatenate([p.chain_pos for p in parts]), np.concatenate(offs),
                        np.concatenate([p.obs_view for p in parts]), np.concatenate([p.obs_poly for p in parts]),
                        np.concatenate([p.obs_seg for p in parts]), np.concatenate([p.obs_xy for p in parts]))
