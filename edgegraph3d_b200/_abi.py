"""ctypes mirror of include/eg3d.h (plain C structs; no torch types cross the boundary)."""
import ctypes as C

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_i64p = C.POINTER(C.c_int64)

EG3D_OK, EG3D_ERR_INVALID_ARG, EG3D_ERR_NO_DEVICE, EG3D_ERR_CUDA, EG3D_ERR_CAPACITY, EG3D_ERR_OOM = range(6)


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("cameras", c_f32p), ("fundamental", c_f64p), ("fundamental_valid", c_u8p),
        ("view_poly_off", c_i64p), ("poly_vert_off", c_i64p), ("verts", c_f32p),
        ("poly_start", c_u32p), ("poly_end", c_u32p),
        ("n_tracks", C.c_int64), ("track_xyz", c_f32p), ("track_off", c_i64p),
        ("track_view", c_i32p), ("track_xy", c_f32p),
    ]


class Params(C.Structure):
    _fields_ = [
        ("split_interval_distance", C.c_float), ("follow_first_image_distance", C.c_float),
        ("follow_corr_min", C.c_float), ("follow_corr_max", C.c_float),
        ("quasiparallel_cos", C.c_float), ("quasiparallel_dist", C.c_float),
        ("max_proj_distsq_expand", C.c_float), ("expand_grid_cell", C.c_float),
        ("detection_starting_radius", C.c_float), ("detection_mult", C.c_float),
        ("gn_max_iters", C.c_int32),
        ("gn_stop", C.c_double), ("gn_det_min", C.c_double), ("gn_accept_mse", C.c_double),
        ("filter_gn_stop", C.c_double), ("filter_gn_det_min", C.c_double),
        ("filter_gn_max_mse", C.c_float), ("filter_3views_amount", C.c_int32),
        ("dedup_cell", C.c_float), ("dlt_wellposed", C.c_int32), ("filter_abs_int", C.c_int32),
        ("max_chain_points", C.c_int32), ("max_follow_points", C.c_int32),
    ]


class Seeds(C.Structure):
    _fields_ = [("n", C.c_int64), ("view", c_i32p), ("polyline", c_u32p), ("segment", c_u32p),
                ("xy", c_f32p), ("cand_set", c_i32p)]


class Candidates(C.Structure):
    _fields_ = [("n_sets", C.c_int32), ("off", c_i64p), ("polyline", c_u32p)]


class Hit(C.Structure):
    _fields_ = [("polyline", C.c_uint32), ("segment", C.c_uint32), ("x", C.c_float), ("y", C.c_float)]


class PointsView(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_obs", C.c_int64), ("xyz", c_f32p), ("seed", c_i32p),
                ("chain_pos", c_i32p), ("obs_off", c_i64p), ("obs_view", c_i32p), ("obs_poly", c_u32p),
                ("obs_seg", c_u32p), ("obs_xy", c_f32p)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("k1_count_ms", C.c_float), ("k1_fill_ms", C.c_float),
                ("scan_ms", C.c_float), ("k3_ms", C.c_float), ("pack_ms", C.c_float), ("gn_ms", C.c_float),
                ("n_seeds", C.c_int64), ("n_hits", C.c_int64), ("n_segment_tests", C.c_int64),
                ("n_points", C.c_int64), ("n_obs", C.c_int64), ("k1_algorithmic_bytes", C.c_int64),
                ("kernel_launches", C.c_int32), ("n_capacity_overflows", C.c_int32),
                ("k3a_ms", C.c_float), ("k3b_ms", C.c_float), ("n_accepted_seeds", C.c_int64), ("k1_any_ms", C.c_float),
                ("host_wall_ms", C.c_float), ("n_capacity_retries", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PlgView(C.Structure):
    _fields_ = [("n_polylines", C.c_int64), ("poly_vert_off", c_i64p), ("verts", c_f32p), ("poly_start", c_u32p),
                ("poly_end", c_u32p), ("poly_length", c_f32p), ("n_nodes", C.c_int64), ("node_xy", c_f32p),
                ("n_pixel_nodes", C.c_int64), ("pixel_node_xy", c_f32p), ("pixel_adj_off", c_i64p), ("pixel_adj", c_u32p)]


class SimilarityGraphView(C.Structure):
    _fields_ = [("n_views", C.c_int32), ("n_nodes", C.c_int64), ("node_view", c_i32p), ("node_polyline", c_u32p), ("n_edges", C.c_int64),
                ("edge_a", c_i64p), ("edge_b", c_i64p), ("edge_weight", c_f32p), ("dimacs", C.c_char_p), ("dimacs_len", C.c_int64)]


PLG_STAGE_FULL, PLG_STAGE_PIXEL_GRAPH, PLG_STAGE_RAW, PLG_STAGE_MERGED, PLG_STAGE_SIMPLIFIED, PLG_STAGE_CONNECTED = range(6)


def ptr(arr, ctype_ptr):
    """numpy array -> typed ctypes pointer (None -> NULL). The caller keeps `arr` alive."""
    if arr is None:
        return C.cast(None, ctype_ptr)
    return arr.ctypes.data_as(ctype_ptr)


class SfmView(C.Structure):
    _fields_ = [("n_views", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("cameras", c_f32p), ("K", c_f32p), ("R", c_f32p), ("center", c_f32p), ("t", c_f32p), ("view_keys", c_i64p),
                ("n_tracks", C.c_int64), ("track_xyz", c_f32p), ("track_off", c_i64p), ("track_view", c_i32p), ("track_xy", c_f32p)]
