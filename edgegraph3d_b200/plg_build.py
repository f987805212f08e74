"""Row f1 (host, upstream of the hot path): edge images -> optimized 2D polyline graphs, through the C-ABI.

`polyline_graph_from_edge_image` is convertEdgeImagePolyLineGraph_optimized
(src/edgegraph3d/io/input/convert_edge_images_pixel_to_segment.cpp:880-883) and `polyline_graphs_from_edge_images`
is convert_edge_images_to_optimized_polyline_graphs (:885-892); the work is done by eg3d_plg_from_edge_image in
libeg3d.so (edgegraph3d_b200/csrc/eg3d_plg_build.cpp, host C++).  Views are independent, so a thread pool runs them
side by side (ctypes releases the GIL during the call); the reference converts them one after the other.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
import os
import numpy as np
from . import _abi as A
from . import lib as E

EDGE_COLOR = (255, 255, 255)  # include/edgegraph3d/utils/globals/global_defines.hpp:47


@dataclass
class PolylineGraph:
    """What the path reads of a PolyLineGraph2DHMapImpl (polyline_graph_2d.hpp:85-119, 236-249)."""
    poly_vert_off: np.ndarray   # [NP+1] i64; an empty range = a removed polyline (ids are kept, as in the reference)
    verts: np.ndarray           # [NVERT,2] f32
    poly_start: np.ndarray      # [NP] u32
    poly_end: np.ndarray        # [NP] u32
    poly_length: np.ndarray     # [NP] f32
    node_xy: np.ndarray         # [NN,2] f32, (-1,-1) = invalidated
    pixel_node_xy: np.ndarray   # [NPIX,2] f32   (the intermediate pixel graph)
    pixel_adj_off: np.ndarray   # [NPIX+1] i64
    pixel_adj: np.ndarray       # u32

    @property
    def n_polylines(self):
        return len(self.poly_start)

    def polyline(self, i):
        return self.verts[int(self.poly_vert_off[i]):int(self.poly_vert_off[i + 1])]

    def n_valid(self):
        return int((np.diff(self.poly_vert_off) > 1).sum())

    def n_segments(self):
        return int(np.maximum(np.diff(self.poly_vert_off) - 1, 0).sum())


def _np(p, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


def polyline_graph_from_edge_image(img, edge_color=EDGE_COLOR, stop_after=A.PLG_STAGE_FULL):
    """img: [rows, cols] (a mask or single channel) or [rows, cols, ch] uint8 as cv2.imread returns it."""
    L = E.load()
    img = np.ascontiguousarray(img)
    if img.dtype == np.bool_:
        img = img.astype(np.uint8) * 255
    if img.dtype != np.uint8 or img.ndim not in (2, 3):
        raise ValueError("edge image must be uint8 [rows, cols] or [rows, cols, channels]")
    rows, cols = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    color = np.ascontiguousarray(np.asarray(edge_color, np.uint8)[:ch] if np.ndim(edge_color) else np.full(ch, edge_color, np.uint8))
    h = C.c_void_p()
    E._check(L.eg3d_plg_from_edge_image(A.ptr(img, A.c_u8p), rows, cols, ch, A.ptr(color, A.c_u8p), int(stop_after), C.byref(h)))
    try:
        v = A.PlgView()
        E._check(L.eg3d_plg_get(h, C.byref(v)))
        npl, nn, npx = int(v.n_polylines), int(v.n_nodes), int(v.n_pixel_nodes)
        off = _np(v.poly_vert_off, npl + 1, np.int64)
        adj_off = _np(v.pixel_adj_off, npx + 1, np.int64)
        return PolylineGraph(
            poly_vert_off=off, verts=_np(v.verts, 2 * int(off[-1]), np.float32).reshape(-1, 2),
            poly_start=_np(v.poly_start, npl, np.uint32), poly_end=_np(v.poly_end, npl, np.uint32),
            poly_length=_np(v.poly_length, npl, np.float32), node_xy=_np(v.node_xy, 2 * nn, np.float32).reshape(-1, 2),
            pixel_node_xy=_np(v.pixel_node_xy, 2 * npx, np.float32).reshape(-1, 2), pixel_adj_off=adj_off,
            pixel_adj=_np(v.pixel_adj, int(adj_off[-1]), np.uint32))
    finally:
        L.eg3d_plg_free(h)


def polyline_graphs_from_edge_images(imgs, edge_color=EDGE_COLOR, workers=None):
    workers = workers or min(len(imgs), os.cpu_count() or 1) or 1
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(lambda im: polyline_graph_from_edge_image(im, edge_color), imgs))


def scene_polyline_arrays(plgs):
    """Concatenate per-view graphs into the CSR arrays of eg3d_scene_desc / FlatScene."""
    view_poly_off = np.zeros(len(plgs) + 1, np.int64)
    offs, verts, st, en = [np.zeros(1, np.int64)], [], [], []
    base = 0
    for v, g in enumerate(plgs):
        view_poly_off[v + 1] = view_poly_off[v] + g.n_polylines
        offs.append(g.poly_vert_off[1:] + base)
        base += int(g.poly_vert_off[-1])
        verts.append(g.verts); st.append(g.poly_start); en.append(g.poly_end)
    return dict(view_poly_off=view_poly_off, poly_vert_off=np.concatenate(offs),
                verts=np.concatenate(verts).reshape(-1, 2) if verts else np.zeros((0, 2), np.float32),
                poly_start=np.concatenate(st) if st else np.zeros(0, np.uint32),
                poly_end=np.concatenate(en) if en else np.zeros(0, np.uint32))
