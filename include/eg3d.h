/*
 * eg3d.h — C-ABI of libeg3d.so: the B200-native replacement for EdgeGraph3D's
 * epipolar polyline-matching + Gauss-Newton triangulation hot path.
 *
 * The reference (abignoli/EdgeGraph3D) has no FFI layer: it is one statically
 * linked C++ executable.  The seams this library sits behind are the free
 * functions / abstract classes listed in SURVEY.md §8(b); every entry point
 * below cites the reference interface it replaces (paths relative to the
 * reference root).  include/eg3d_ref_api.hpp is the thin C++ shim that keeps
 * the reference's own signatures on top of this ABI.
 *
 * Conventions: plain C structs, caller-owned inputs (copied during the call),
 * library-owned outputs released with the matching *_free / *_destroy, int
 * status codes, no exceptions across the boundary.  A scene handle is bound to
 * the CUDA device that was current when it was created.  Thread-compatible,
 * not thread-safe.  Every entry point of the hot path (scene creation, matching,
 * Gauss-Newton, filtering) fails with EG3D_ERR_NO_DEVICE when no CUDA device is
 * usable: there is no CPU fallback in this library.  The entry points marked
 * "host" (seed sampler, camera fundamentals, and the stages upstream of the path:
 * polyline graphs from edge images, candidate polyline sets) are order-dependent
 * host code in the reference's design too and need no device.
 */
#ifndef EG3D_H_
#define EG3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum eg3d_status {
  EG3D_OK = 0,
  EG3D_ERR_INVALID_ARG = 1,
  EG3D_ERR_NO_DEVICE = 2,   /* no usable CUDA device: the product path never falls back to the CPU */
  EG3D_ERR_CUDA = 3,        /* a CUDA runtime call failed; see eg3d_last_error() */
  EG3D_ERR_CAPACITY = 4,    /* an output or index capacity was exceeded even after the library's own retries (or a caller buffer is too small) */
  EG3D_ERR_OOM = 5
} eg3d_status;

/* ------------------------------------------------------------------------- */
/* Flat scene: what SfMData + vector<PolyLineGraph2DHMapImpl> + Mat** F hold.  */
/* ------------------------------------------------------------------------- */
/*
 * Replaces, as inputs of every entry point of the path:
 *   SfMData                     external/manifoldReconstructor/include/SfMData.h:16-30
 *   CameraType::cameraMatrix    external/manifoldReconstructor/include/types_reconstructor.hpp:68-82
 *   PolyLineGraph2D::polyline   include/edgegraph3d/plgs/polyline_graph_2d.hpp:85-119
 *   const Mat** all_fundamental_matrices   src/edgegraph3d/utils/edge_graph_3d_utilities.cpp:581-589
 */
typedef struct eg3d_scene_desc {
  int32_t n_views;
  int32_t width, height;            /* image size (all views), SfMData::imageWidth_/imageHeight_ */
  const float*   cameras;           /* [V][12] rows 0..2 of cameraMatrix read as [row][col]: P = K[R|t] */
  const double*  fundamental;       /* [V][V][9] row-major; F[a][b] maps a point of view a to its line in view b */
  const uint8_t* fundamental_valid; /* [V][V] 0 = the reference's 1x1 dummy Mat (geometric_utilities.cpp:780) */
  /* polyline graphs, CSR.  Polyline id inside a view = global index - view_poly_off[view].               */
  const int64_t*  view_poly_off;    /* [V+1] */
  const int64_t*  poly_vert_off;    /* [NP+1]; an empty range is an invalidated polyline (polyline_graph_2d.cpp:1035-1038) */
  const float*    verts;            /* [NVERT][2] polyline_coords */
  const uint32_t* poly_start;       /* [NP] polyline::start node id */
  const uint32_t* poly_end;         /* [NP] polyline::end node id */
  /* SfM tracks: only needed by eg3d_match_refpoints and eg3d_filter (may be NULL with n_tracks = 0).     */
  int64_t        n_tracks;
  const float*   track_xyz;         /* [NT][3] SfMData::points_ */
  const int64_t* track_off;         /* [NT+1] */
  const int32_t* track_view;        /* [NOBS] SfMData::camViewingPointN_ */
  const float*   track_xy;          /* [NOBS][2] SfMData::point2DoncamViewingPoint_ */
} eg3d_scene_desc;

/* Every compile-time constant of the path, defaulted by eg3d_params_default(). */
typedef struct eg3d_params {
  float  split_interval_distance;     /* 20    polyline_matching.hpp:51 */
  float  follow_first_image_distance; /* 10    plg_matching.hpp:39 */
  float  follow_corr_min;             /* 5     plg_matching.hpp:40 */
  float  follow_corr_max;             /* 20    plg_matching.hpp:41 */
  float  quasiparallel_cos;           /* 0.965 polyline_graph_2d.hpp:73 */
  float  quasiparallel_dist;          /* 5     polyline_graph_2d.hpp:74 */
  float  max_proj_distsq_expand;      /* 16    triangulation.hpp:46 */
  float  expand_grid_cell;            /* 4     edge_matcher.cpp:103 */
  float  detection_starting_radius;   /* 10    global_defines.hpp:35 */
  float  detection_mult;              /* 3     global_defines.hpp:37 */
  int32_t gn_max_iters;               /* 30    triangulation.cpp:122, gauss_newton.cpp:97 */
  double gn_stop;                     /* 5e-7  triangulation.cpp:150 */
  double gn_det_min;                  /* 1e-5  triangulation.cpp:97 */
  double gn_accept_mse;               /* 9     triangulation.cpp:168 */
  double filter_gn_stop;              /* 5e-10 gauss_newton.cpp:114 (double literal; the float |diff| is promoted) */
  double filter_gn_det_min;           /* 1e-10 gauss_newton.cpp:70 */
  float  filter_gn_max_mse;           /* 2.25  gauss_newton.hpp:18 */
  int32_t filter_3views_amount;       /* 3     outliers_filtering.hpp:16 */
  float  dedup_cell;                  /* 3     filtering_close_plgps.cpp:75 */
  /* The 2-view DLT initialiser of em_estimate3Dpositions (triangulation.cpp:252-290) gets the camera pair chosen by
   * get_min_max, whose "max" is always the LAST list entry (edge_graph_3d_utilities.hpp:86-88); when that is also the lowest
   * view, cv::triangulatePoints sees the same camera twice and returns whichever point of the back-projected ray its SVD lands on.
   * 2 (default; 0 is an alias): the reference's pair with OpenCV 4.x's own Jacobi SVD restated (bit-identical to
   *    cv2.triangulatePoints incl. the degenerate case) — reproduces a reference linked against OpenCV.
   * 1: use the last list entry whose view differs instead (well-posed; NOT what the reference computes).  Same SVD arithmetic. */
  int32_t dlt_wellposed;
  /* 0 (default): `abs(mse/(2n) - last_mse)` in gauss_newton.cpp:114 is the float overload (GCC >= 6).
   * 1: emulate the truncating `int abs(int)` binding of the author's GCC 5 toolchain.         */
  int32_t filter_abs_int;
  /* capacities of the device path (not reference constants) */
  int32_t max_chain_points;           /* per seed; default 96.  Starting sizes: a batch that needs more is re-run inside the call with doubled  */
  int32_t max_follow_points;          /* per direction during 3-view following; default 160.             capacities (eg3d_timing.n_capacity_retries) */
} eg3d_params;

void eg3d_params_default(eg3d_params* p);

/* Seeds: the (view, plg_point) pairs the matching loop starts from
 * (polyline_matching.cpp:168-190 produces them; plg_edge_manager.cpp:272-288 for pipeline 3). */
typedef struct eg3d_seeds {
  int64_t         n;
  const int32_t*  view;      /* [n] starting_plg_id */
  const uint32_t* polyline;  /* [n] plg_point::polyline_id */
  const uint32_t* segment;   /* [n] pl_point::segment_index */
  const float*    xy;        /* [n][2] pl_point::coords */
  const int32_t*  cand_set;  /* [n] index into eg3d_candidates, or NULL / -1 = sweep all segments of every view */
} eg3d_seeds;

/* Candidate polyline sets: `vector<set<ulong>> potentially_compatible_polylines`, one per polyline match
 * (pipelines.cpp:92-100).  CSR over (set, view) -> ascending polyline ids.                                   */
typedef struct eg3d_candidates {
  int32_t         n_sets;
  const int64_t*  off;       /* [n_sets*V + 1] */
  const uint32_t* polyline;  /* ids, ascending inside each (set, view) range */
} eg3d_candidates;

/* One epipolar hit = PolyLineGraph2D::plg_point (polyline_graph_2d.hpp:278-294) narrowed to 16 bytes. */
typedef struct eg3d_hit {
  uint32_t polyline;
  uint32_t segment;
  float    x, y;
} eg3d_hit;

/* Accepted 3D edge-points: flattened vector<new_3dpoint_plgp_matches> (polyline_graph_2d.hpp:451). */
typedef struct eg3d_points_view {
  int64_t n_points, n_obs;
  const float*    xyz;        /* [n_points][3] */
  const int32_t*  seed;       /* [n_points] ordinal of the producing seed in the call's seed order */
  const int32_t*  chain_pos;  /* [n_points] position inside that seed's chain */
  const int64_t*  obs_off;    /* [n_points+1] */
  const int32_t*  obs_view;   /* [n_obs] */
  const uint32_t* obs_poly;   /* [n_obs] */
  const uint32_t* obs_seg;    /* [n_obs] */
  const float*    obs_xy;     /* [n_obs][2] */
} eg3d_points_view;

/* Per-call device timings (CUDA events on the library's stream), in milliseconds. */
typedef struct eg3d_timing {
  float total_ms;             /* first kernel launch .. last kernel end (no H2D of the scene, no D2H of results) */
  float k1_count_ms;          /* epipolar intersection, counting pass */
  float k1_fill_ms;           /* epipolar intersection, fill pass */
  float scan_ms;              /* prefix sums */
  float k3_ms;                /* k3a_ms + k3b_ms */
  float pack_ms;              /* ordered compaction of accepted points */
  float gn_ms;                /* stand-alone GN kernel (eg3d_gn_*) */
  int64_t n_seeds, n_hits /* hit records materialised */, n_segment_tests /* of the reference's full sweep */, n_points, n_obs;
  int64_t k1_algorithmic_bytes; /* 16 B x segments swept per (seed, view) + 72 + 8 + 16 B x hits */
  int32_t kernel_launches;
  int32_t n_capacity_overflows; /* seeds dropped because a capacity in eg3d_params was exceeded */
  float k3a_ms;               /* K3 phase A: view triples, triple enumeration, PLG following (all seeds) */
  float k3b_ms;               /* K3 phase B: expansion to the remaining views (accepted seeds) */
  int64_t n_accepted_seeds;
  float k1_any_ms;            /* epipolar intersection, any-hit pass of the lazy sweep (eg3d_match_seeds without candidates) */
  float host_wall_ms;         /* wall-clock time of the whole C call on the host (uploads, kernels, syncs; what e2e pays) */
  int32_t n_capacity_retries; /* K3 was re-run with larger per-seed capacities this many times (never truncates, never fails on them) */
} eg3d_timing;

typedef struct eg3d_scene  eg3d_scene;   /* opaque, device resident */
typedef struct eg3d_points eg3d_points;  /* opaque, host+device resident result */
typedef struct eg3d_hits   eg3d_hits;    /* opaque */

const char* eg3d_last_error(void);
/* Fundamental matrices implied by the camera matrices themselves (x_b^T F[a][b] x_a = 0 for the projections of any X):
 * the role of findFundamentalMatrixFromRt (geometric_utilities.cpp:683-710) for rigs with known poses.  Host code.
 * cameras [V][12], out [V][V][9] (row-major, unit Frobenius norm, zero on the diagonal). */
void        eg3d_camera_fundamentals(const float* cameras, int32_t n_views, double* out);
/* Row f4 without OpenCV: generate_all_fundamental_matrices_from_Points (geometric_utilities.cpp:754-820) — for every ordered view
 * pair (i, j) with at least `min_common` (reference: MIN_CORRESPONDENCES_AMOUNT = 10) SfM tracks seen by both, a least-median-of-
 * squares estimate of F[i][j] (x_j^T F x_i = 0) from the tracks' observations in ascending track id; other pairs stay invalid (the
 * reference's 1x1 dummy Mat).  Host code, deterministic.  Same estimator family and constants as cv::findFundamentalMat(FM_LMEDS)
 * (normalised 8-point on minimal samples, confidence 0.99, 2.5 * 1.4826 * (1 + 5/(n-8)) * sqrt(median) inlier band, refit on the
 * inliers, rank 2, F[2][2] = 1) but NOT bit-identical to it: OpenCV's is driven by its own RNG, 7-point solver and SVD.  The path
 * takes F as an input, as the reference's entry points do, so either source can be passed to eg3d_scene_create.
 * Uses n_views and the track arrays of `desc` only.  out_F [V][V][9], out_valid [V][V]. */
eg3d_status eg3d_fundamental_from_tracks(const eg3d_scene_desc* desc, int32_t min_common, double* out_F, uint8_t* out_valid);

/* Row f3 in C++ (host code, no device): the data formats either side of the path.
 *   eg3d_sfm_load   OpenMVG sfm_data JSON -> what SfMData holds (OpenMvgParser.cpp:75-153, 241-301): view index = position
 *                   in `extrinsics`; cameras = rows 0..2 of cameraMatrix = eMatrix * kMatrix with translation = -center * rotation,
 *                   float arithmetic in the operation order of the reference's vendored glm (bit-identical camera matrices);
 *                   radial distortion ignored as the reference does; tracks in file order.  The arrays feed eg3d_scene_desc directly.
 *   eg3d_sfm_save   output_sfm_data (output_sfm_data.cpp:186-229): views / intrinsics / extrinsics of `original_path` kept verbatim,
 *                   `structure` rewritten from the given points (key = running index, id_feat = 0), only those with keep[i] != 0 when
 *                   `keep` is given; obs_view are view INDICES (mapped back to the file's pose keys).
 *   eg3d_write_ply  output_point_cloud.cpp (ascii PLY, optional per-point colour).                                              */
typedef struct eg3d_sfm eg3d_sfm;     /* opaque */
typedef struct eg3d_sfm_view {
  int32_t n_views, width, height;
  const float*   cameras;     /* [V][12] */
  const float*   K;           /* [V][9]  */
  const float*   R;           /* [V][9]  */
  const float*   center;      /* [V][3]  */
  const float*   t;           /* [V][3]  */
  const int64_t* view_keys;   /* [V] pose keys of the file, in view order */
  int64_t        n_tracks;
  const float*   track_xyz;   /* [NT][3] */
  const int64_t* track_off;   /* [NT+1]  */
  const int32_t* track_view;  /* [NOBS]  */
  const float*   track_xy;    /* [NOBS][2] */
} eg3d_sfm_view;
eg3d_status eg3d_sfm_load(const char* path, eg3d_sfm** out);
eg3d_status eg3d_sfm_get(const eg3d_sfm*, eg3d_sfm_view* view);   /* pointers valid until eg3d_sfm_free */
void        eg3d_sfm_free(eg3d_sfm*);
eg3d_status eg3d_sfm_save(const char* path, const char* original_path, int64_t n_points, const float* xyz, const int64_t* obs_off,
                          const int32_t* obs_view, const float* obs_xy, const uint8_t* keep /* may be NULL */, int64_t* n_written /* may be NULL */);
eg3d_status eg3d_write_ply(const char* path, int64_t n_points, const float* xyz, const uint8_t* rgb /* may be NULL */);
int         eg3d_device_count(void);
/* Host evaluation of compute_projection (geometric_utilities.cpp:973-977) as the kernels compute it (tests). */
void        eg3d_project_host(const float* cam12, const float* x3, float* out2);
const char* eg3d_build_info(void);   /* the compile-time switches of this build, "NAME=value ..." */
/* Host evaluation of the 2-view DLT initialiser (cv::triangulatePoints at triangulation.cpp:216,290): opencv_svd = 1 -> what the
 * kernels compute, OpenCV's own Jacobi SVD restated (bit-identical to cv2.triangulatePoints, degenerate inputs included);
 * 0 -> a plain one-sided Jacobi SVD (round 1's initialiser, host only now; kept to show that another SVD returns another point). */
void        eg3d_triangulate_dlt_host(const float* P1, const float* P2, const float* x1, const float* x2,
                                      int32_t opencv_svd, float* out4);

/* Copies the scene to the current CUDA device and builds the derived structures the path reads:
 * per-view segment arrays in the reference's (P[i], P[i-1]) orientation (plg_edge_manager.hpp:92-93,
 * polyline_graph_2d.cpp:318-323) and the uniform polyline grids (polyLine_2d_map.cpp:40-58) for the
 * 4 px expansion lookups and the 30 px refpoint lookups. */
eg3d_status eg3d_scene_create(const eg3d_scene_desc* desc, const eg3d_params* params, eg3d_scene** out);
void        eg3d_scene_destroy(eg3d_scene*);

/* Seed sampler: polyline::next_pl_point_by_distance walked from get_start_plp() towards `end`
 * (polyline_matching.cpp:168-190, polyline_graph_2d.cpp:391-447).  Host code.
 * Writes up to `capacity` seeds for the given (view, polyline) list; returns the count in *n_out. */
eg3d_status eg3d_sample_seeds(const eg3d_scene_desc* desc, const int32_t* views, const uint32_t* polylines,
                              int64_t n_polylines, float spacing, int64_t capacity,
                              int32_t* out_view, uint32_t* out_polyline, uint32_t* out_segment, float* out_xy,
                              int32_t* out_src /* index of the source (view,polyline) pair, may be NULL */,
                              int64_t* n_out);

/* K1 alone: find_epipolar_correspondences (polyline_matching.cpp:45-73) for a batch of seeds.
 * Result: CSR over (seed, view) of eg3d_hit, in the reference's order (polyline id asc, segment asc). */
eg3d_status eg3d_epipolar_intersect(eg3d_scene*, const eg3d_seeds*, const eg3d_candidates* /* may be NULL */,
                                    eg3d_hits** out, eg3d_timing* timing /* may be NULL */);
/* Same computation with the result left on (and released from) the device: what BASELINE configs 2-4 measure, the
 * full sweep itself.  Only `timing` is returned (k1_count_ms, k1_fill_ms, n_hits, n_segment_tests, k1_algorithmic_bytes). */
eg3d_status eg3d_epipolar_intersect_device(eg3d_scene*, const eg3d_seeds*, const eg3d_candidates* /* may be NULL */,
                                           eg3d_timing* timing);
/* host pointers valid until eg3d_hits_free; off has n_seeds*V+1 entries */
eg3d_status eg3d_hits_get(const eg3d_hits*, int64_t* n_seeds, int32_t* n_views, const int64_t** off, const eg3d_hit** hits);
void        eg3d_hits_free(eg3d_hits*);

/* B1/B4: find_new_3d_points_from_compatible_polylines_starting_plgp_expandallviews for a batch of seeds
 * (polyline_matching.cpp:134-144 -> triangulation.cpp:1027-1088): K1 + triple enumeration + PLG following +
 * view expansion.  Output order = seed order, chain order inside a seed. */
eg3d_status eg3d_match_seeds(eg3d_scene*, const eg3d_seeds*, const eg3d_candidates* /* may be NULL */,
                             eg3d_points** out, eg3d_timing* timing);

/* B1 batched over polyline matches: find_new_3d_points_from_compatible_polylines_expandallviews[_parallel]
 * (polyline_matching.hpp:55-56) called once per candidate set as pipelines.cpp:92-100 does; seeds are
 * sampled internally (views ascending, candidate polylines ascending, 20 px steps).
 * view_begin/view_end restrict the STARTING views (multi-GPU shard axis, polyline_matching.cpp:162). */
eg3d_status eg3d_match_polyline_sets(eg3d_scene*, const eg3d_candidates*, int32_t view_begin, int32_t view_end,
                                     eg3d_points** out, eg3d_timing* timing);

/* B2: plg_matching_from_refpoints[_parallel] (plg_matching_from_refpoints.hpp:53-55) over tracks
 * [track_begin, track_end) of the scene (multi-GPU shard axis, plg_matching_from_refpoints.cpp:90). */
eg3d_status eg3d_match_refpoints(eg3d_scene*, int64_t track_begin, int64_t track_end,
                                 eg3d_points** out, eg3d_timing* timing);

/* B4: compute_3D_point_multiple_views_plg_following_expandallviews_vector(sfmd, plgs, F, starting_plg_id,
 * epipolar_correspondences, plmaps) (include/edgegraph3d/utils/geometry/triangulation.hpp:98; triangulation.cpp:1027-1088) for a
 * batch of seeds whose hit lists the CALLER supplies — what PLGPCM3ViewsPLGFollowing::consensus_strategy_single_point_single_
 * intersection hands it (plgpcm_3views_plg_following.cpp:40-50) and what polyline_matching.cpp:140 hands it.  hit_off has
 * n_seeds*V+1 entries (CSR over (seed, view), lists in the caller's order); the starting view's list normally holds the seed
 * itself (polyline_matching.cpp:54-55).  View-triple selection, triple enumeration with the uniqueness test, PLG following and
 * view expansion run on the device; no epipolar search is done here, so a host that keeps its own EdgeManager swaps only the
 * consensus step. */
eg3d_status eg3d_match_correspondences(eg3d_scene*, int64_t n_seeds, const int32_t* start_view, const int64_t* hit_off,
                                       const eg3d_hit* hits, eg3d_points** out, eg3d_timing* timing);

/* B3, EdgeManager side: PLGEdgeManager::detect_nearby_intersections_and_correspondences_plgp(refpoint)
 * (include/edgegraph3d/edge_managers/plg_edge_manager.hpp:74; plg_edge_manager.cpp:261-300) for the SfM points [tb, te): the seeds
 * (view, plg_point, producing track) in the reference's order and their per-view hit lists (CSR over (seed, view), radius
 * filter applied), computed on the device and copied to host memory owned by the handle. */
typedef struct eg3d_corr eg3d_corr;
eg3d_status eg3d_refpoint_correspondences(eg3d_scene*, int64_t track_begin, int64_t track_end, eg3d_corr** out);
eg3d_status eg3d_corr_get(const eg3d_corr*, int64_t* n_seeds, int32_t* n_views, const int32_t** view, const uint32_t** polyline,
                          const uint32_t** segment, const float** xy, const int64_t** track, const int64_t** hit_off,
                          const eg3d_hit** hits);   /* any out pointer may be NULL */
void        eg3d_corr_free(eg3d_corr*);

/* Results stay on the device until asked for: eg3d_points_get copies them (once) into page-locked host memory owned by
 * the handle; eg3d_points_device_get exposes the device-resident arrays (same layout, device pointers on the scene's
 * device) for device-side consumers such as the multi-GPU all-gather. */
eg3d_status eg3d_points_get(const eg3d_points*, eg3d_points_view* view);
eg3d_status eg3d_points_device_get(const eg3d_points*, eg3d_points_view* device_view);
void        eg3d_points_free(eg3d_points*);

/* ------------------------------------------------------------------------- */
/* Multi-GPU exchange: the path's ONE collective step              */
/* ------------------------------------------------------------------------- */
/*
 * The reference is one process; its loops over starting views (polyline_matching.cpp:162) and over SfM points
 * (plg_matching_from_refpoints.cpp:90) are what the ranks of a multi-GPU run split between them, each on a full replica of
 * the read-only scene.  Before the order-dependent density limiter (filtering_close_plgps.cpp:75-124) every rank needs
 * every accepted record, in the reference's loop order.  One process per GPU; the scene handle owns the NCCL communicator
 * (NCCL is loaded at run time by soname, so a host that already runs one — torch.distributed — shares it).
 *   eg3d_comm_unique_id   rank 0 makes the id and hands it to the other ranks by whatever channel the host has
 *   eg3d_comm_create      collective: attaches a communicator of `world` ranks to the scene
 *   eg3d_points_allgather collective: one ncclAllGather of the counts + ONE grouped broadcast of the byte-packed records + a
 *                         device-side merge (radix sort) into ascending (global seed ordinal, chain position).
 *                         seed_global[i] (host, one per seed of the call that produced `mine`) is the ordinal of local seed i
 *                         in the unsharded seed list; NULL = rank-major (shards are contiguous blocks of the loop).  The
 *                         merged result is device resident like any other eg3d_points; its `seed` field holds global ordinals.
 *                         timing: total_ms = exchange + merge, scan_ms = the grouped broadcast alone.
 */
#define EG3D_COMM_ID_BYTES 128
eg3d_status eg3d_comm_unique_id(uint8_t id[EG3D_COMM_ID_BYTES]);
eg3d_status eg3d_comm_create(eg3d_scene*, const uint8_t id[EG3D_COMM_ID_BYTES], int32_t rank, int32_t world);
eg3d_status eg3d_comm_destroy(eg3d_scene*);
eg3d_status eg3d_points_allgather(eg3d_scene*, const eg3d_points* mine, const int64_t* seed_global /* may be NULL */,
                                  eg3d_points** merged, eg3d_timing* timing /* may be NULL */);

/* K2 alone (B5/B6 primitives).  Hypotheses: CSR of (view, xy) observations + initial X.
 * fp64 = 1: em_GaussNewton semantics (triangulation.cpp:105-176), inputs f32, arithmetic f64.
 * fp64 = 0: GaussNewton of the outlier filter (filtering/gauss_newton.cpp:83-134), arithmetic f32.
 * Host-buffer call: copies in, runs, copies out.  out_xyz [n][3], out_mse [n] (last_mse), out_ok [n]. */
eg3d_status eg3d_gn_triangulate(eg3d_scene*, int64_t n_hyp, const int64_t* obs_off, const int32_t* obs_view,
                                const float* obs_xy, const float* init_xyz, int fp64,
                                float* out_xyz, float* out_mse, uint8_t* out_ok, eg3d_timing* timing);

/* Device-resident variant used by the microbenchmark: all pointers are device pointers on the scene's
 * device; fixed `obs_per_hyp` observations per hypothesis (obs arrays are [n][obs_per_hyp]). */
eg3d_status eg3d_gn_triangulate_device(eg3d_scene*, int64_t n_hyp, int32_t obs_per_hyp, const int32_t* d_obs_view,
                                       const float* d_obs_xy, const float* d_init_xyz, int fp64,
                                       float* d_out_xyz, float* d_out_mse, uint8_t* d_out_ok, eg3d_timing* timing);

/* a13: filter_3d_points_close_2d_array (filtering_close_plgps.cpp:75-124).  keep[n_points] on host.
 * Order dependent: points are visited in the order given. */
eg3d_status eg3d_dedup_close_points(eg3d_scene*, const eg3d_points_view* pts, uint8_t* keep);

/* a14 / B6: filter(sfmd, first_edgepoint, gn_max_mse, forced_min_filter) (outliers_filtering.hpp:18-21).
 * Points are tracks given as CSR; xyz is updated in place for GN inliers (gauss_newton.cpp:168-173);
 * inliers[n] receives compute_inliers' bitmap (outliers_filtering.cpp:37-64). forced_min_filter < 0 = off. */
eg3d_status eg3d_filter(eg3d_scene*, int64_t n, float* xyz, const int64_t* obs_off, const int32_t* obs_view,
                        const float* obs_xy, int64_t first_edgepoint, float gn_max_mse, int32_t forced_min_filter,
                        uint8_t* inliers, eg3d_timing* timing);

/* ------------------------------------------------------------------------- */
/* Row f1 (host, upstream of the path): edge image -> optimized polyline graph */
/* ------------------------------------------------------------------------- */
/*
 * Replaces convertEdgeImagePolyLineGraph_optimized(const Mat& img, const Vec3b& edge_color)
 * (src/edgegraph3d/io/input/convert_edge_images_pixel_to_segment.cpp:880-883, called per view by
 * convert_edge_images_to_optimized_polyline_graphs :885-892 from edge_matcher.cpp:83): pixel graph without
 * short cycles (:294-426), polyline extraction (:428-626) and PolyLineGraph2DHMapImpl::optimize
 * (src/edgegraph3d/plgs/polyline_graph_2d_hmap_impl.cpp:255-266).  Order-dependent graph surgery: host code, no
 * device needed.  The result lists EVERY polyline id the reference's graph holds; removed polylines have an empty
 * vertex range, exactly what eg3d_scene_desc expects.
 */
typedef struct eg3d_plg eg3d_plg;  /* opaque, host resident */
typedef enum eg3d_plg_stage {      /* stop_after: intermediate graphs for tests and debugging */
  EG3D_PLG_STAGE_FULL = 0,         /* the reference's result */
  EG3D_PLG_STAGE_PIXEL_GRAPH = 1,  /* convertEdgeImagePixelToGraph_NoCycles only (no polylines) */
  EG3D_PLG_STAGE_RAW = 2,          /* + convert_EdgeGraph_to_PolyLineGraph */
  EG3D_PLG_STAGE_MERGED = 3,       /* + remove_invalid_polylines, remove_degenerate_loops, remove_2connection_nodes */
  EG3D_PLG_STAGE_SIMPLIFIED = 4,   /* + PolyLineGraph2D::optimize (1 px simplification) */
  EG3D_PLG_STAGE_CONNECTED = 5     /* + connect_close_extremes (6 px) and the second simplification */
} eg3d_plg_stage;
typedef struct eg3d_plg_view {
  int64_t         n_polylines;
  const int64_t*  poly_vert_off;   /* [NP+1] */
  const float*    verts;           /* [NVERT][2] polyline_coords */
  const uint32_t* poly_start;      /* [NP] */
  const uint32_t* poly_end;        /* [NP] */
  const float*    poly_length;     /* [NP] polyline::length (-1 = invalidated) */
  int64_t         n_nodes;
  const float*    node_xy;         /* [NN][2] nodes_coords; (-1,-1) = invalidated node */
  int64_t         n_pixel_nodes;   /* the intermediate pixel graph (GraphAdjacencySetUndirectedNoType + node coords) */
  const float*    pixel_node_xy;   /* [NPIX][2] */
  const int64_t*  pixel_adj_off;   /* [NPIX+1] */
  const uint32_t* pixel_adj;       /* ascending neighbour ids */
} eg3d_plg_view;
/* img: rows x cols x channels, continuous (cv::imread layout); a pixel is an edge pixel iff all its channels equal
 * edge_color[c] (EDGE_COLOR = (255,255,255), include/edgegraph3d/utils/globals/global_defines.hpp:47).
 * The caller's image is not modified (the reference clears "useless hub" pixels in place). */
eg3d_status eg3d_plg_from_edge_image(const uint8_t* img, int32_t rows, int32_t cols, int32_t channels,
                                     const uint8_t* edge_color, int32_t stop_after, eg3d_plg** out);
eg3d_status eg3d_plg_get(const eg3d_plg*, eg3d_plg_view* view);   /* pointers valid until eg3d_plg_free */
void        eg3d_plg_free(eg3d_plg*);

/* ------------------------------------------------------------------------- */
/* Row f2 (host, producer of pipeline 2's candidate sets)                    */
/* ------------------------------------------------------------------------- */
/*
 * Replaces polyline_matching_closeness_to_refpoints(plgs, sfmd, img_sz)
 * (src/edgegraph3d/matching/polyline_matching/polyline_matcher.cpp:75-168, called from pipelines.cpp:118): an SfM point
 * whose observations each lie within FIND_WITHIN_DIST (10 px, polyline_matcher.hpp:45) of at most one polyline, in >= 70 %
 * of its views and at comparable distances (ratio <= DETECTION_CORRESPONDENCES_MULTIPLICATION_FACTOR = 3), links those
 * (view, polyline) pairs; every connected component of that graph is one `vector<set<ulong>>` candidate set, i.e. one
 * eg3d_match_polyline_sets input.  Host code over the caller's arrays (the polyline-to-polyline similarity graph +
 * Louvain producer of pipeline 1, :222-336, is not built: its community detection is vendored third-party code).
 */
typedef struct eg3d_polyline_sets eg3d_polyline_sets;  /* opaque, host resident */
eg3d_status eg3d_polyline_sets_from_refpoints(const eg3d_scene_desc* desc, float find_within_dist, float mult,
                                              eg3d_polyline_sets** out);
/* view->off has n_sets*V+1 entries; refpoints = ids of the SfM points that contributed (the function's first result). */
eg3d_status eg3d_polyline_sets_get(const eg3d_polyline_sets*, eg3d_candidates* view, int64_t* n_refpoints,
                                   const int64_t** refpoints);
void        eg3d_polyline_sets_free(eg3d_polyline_sets*);

/*
 * Pipeline 1's producer: polyline_matching_similarity_graph (polyline_matcher.cpp:222-336, called from pipelines.cpp:72).
 * eg3d_polyline_similarity_graph builds the reference's weighted compatibility graph: nodes = (view, polyline) pairs
 * within FIND_WITHIN_DIST of an SfM observation, in first-seen order; an edge joins two pairs close to the same SfM point;
 * its weight is the weighted Jaccard index of the two polylines' close SfM points (compute_compatibility :171-194, point
 * weights compute_refpoint_weight :196-205); edges of weight 0 are dropped.  `dimacs` is, byte for byte, the file
 * GraphAdjacencySetUndirectedNoTypeWeighted::write_to_file (graph_adjacency_set_undirected_no_type_weighted.cpp:54-73)
 * hands to Grappolo.  The reference then takes Grappolo's communities (community_detection_interface.cpp:57-73); that code
 * is multi-threaded only and not reproducible run to run, so eg3d_similarity_graph_communities is this library's own
 * deterministic sequential Louvain, and eg3d_polyline_sets_from_communities
 * (compute_polyline_matches_from_nodes_component_ids :207-219) accepts ids from either.  Host code.
 */
typedef struct eg3d_similarity_graph eg3d_similarity_graph;  /* opaque, host resident */
typedef struct eg3d_similarity_graph_view {
  int32_t         n_views;
  int64_t         n_nodes;
  const int32_t*  node_view;      /* [n_nodes] */
  const uint32_t* node_polyline;  /* [n_nodes] */
  int64_t         n_edges;
  const int64_t*  edge_a;         /* [n_edges] edge_a < edge_b */
  const int64_t*  edge_b;
  const float*    edge_weight;
  const char*     dimacs;         /* NUL-terminated text of the compatibility-graph file */
  int64_t         dimacs_len;
} eg3d_similarity_graph_view;
eg3d_status eg3d_polyline_similarity_graph(const eg3d_scene_desc* desc, float find_within_dist, eg3d_similarity_graph** out);
eg3d_status eg3d_similarity_graph_get(const eg3d_similarity_graph*, eg3d_similarity_graph_view* view);
eg3d_status eg3d_similarity_graph_communities(const eg3d_similarity_graph*, int64_t* community /* [n_nodes] */,
                                              double* modularity /* may be NULL */);
eg3d_status eg3d_polyline_sets_from_communities(const eg3d_similarity_graph*, const int64_t* community /* [n_nodes] */,
                                                eg3d_polyline_sets** out);
void        eg3d_similarity_graph_free(eg3d_similarity_graph*);

#ifdef __cplusplus
}
#endif
#endif /* EG3D_H_ */
