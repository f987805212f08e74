"""Host driver of the path with the reference's top-level shape: edge_matching() (src/edgegraph3d/edge_matcher.cpp:61-146)
and edge_reconstruction_pipeline() (src/edgegraph3d/matching/plg_matching/pipelines.cpp:201-246) as thin glue over the
C-ABI.  Everything that computes lives in libeg3d.so; this file only orders the calls and moves arrays:

  read_sfm_data            -> openmvg_io.load_sfm_data                      (f3)
  edge images -> PLGs      -> eg3d_plg_from_edge_image                      (f1, host)
  fundamental matrices     -> openmvg_io.fundamental_from_tracks           (f4: cv2 LMedS as in the reference, or libeg3d.so's own LMedS without OpenCV)
  pipeline 1 candidates    -> eg3d_polyline_similarity_graph / _communities / eg3d_polyline_sets_from_communities (f2, host)
  pipeline 2 candidates    -> eg3d_polyline_sets_from_refpoints             (f2, host)
  pipelines 1, 2           -> eg3d_match_polyline_sets                      (device)
  pipeline 3               -> eg3d_match_refpoints                          (device)
  density limiter          -> eg3d_dedup_close_points                       (a13)
  add_3dpoints_to_sfmd     -> concatenation after the SfM points            (output_utilities.cpp:96-111)
  filter                   -> eg3d_filter                                   (a14, device)
  output_sfm_data          -> openmvg_io.save_sfm_data                      (f3)
"""
import os
import time
import numpy as np
from . import lib as E
from . import openmvg_io as io
from . import real_scene
from .scene import PointSet, ranges

# chains on real polyline graphs are longer than on the synthetic rigs the defaults are sized for: starting the
# per-seed capacities here saves the library its own retry (eg3d_timing.n_capacity_retries); the defaults work too
REAL_DATA_CAPACITIES = dict(max_chain_points=256, max_follow_points=320)


def candidate_sets(scene):
    """The two producers of `potentially_compatible_polylines` (pipelines.cpp:72, :118) -> (sets for pipeline 1, sets for pipeline 2, info)."""
    t = time.perf_counter()
    graph = E.SimilarityGraph(scene)
    com, q = graph.communities()
    c1 = graph.candidate_sets(com)
    t1 = time.perf_counter() - t
    t = time.perf_counter()
    c2, refpoints = E.polyline_sets_from_refpoints(scene)
    t2 = time.perf_counter() - t
    info = {"compatibility_graph": {"nodes": len(graph.node_view), "edges": len(graph.edge_a), "communities": int(com.max()) + 1 if len(com) else 0,
                                    "modularity": q},
            "candidate_sets_pipeline1": c1.n_sets, "candidate_sets_pipeline2": c2.n_sets, "contributing_sfm_points": len(refpoints),
            "host_s": {"compatibility_graph_and_communities": t1, "sfm_point_components": t2}}
    graph.close()
    return c1, c2, info


def run_pipelines(matcher, scene, cands1, cands2, dist=None, device=None, spacing=20.0):
    """Pipelines 1-3 in the reference's order (pipelines.cpp:217-229).  `matcher` is a lib.DeviceScene (or any object with the
    same match_polyline_sets / match_refpoints methods; the tests pass a stand-in to exercise this glue without a GPU).  -> ([points1, points2, points3], [timing...])

    With `dist` (an initialised torch.distributed, world size N > 1) every rank computes its shard — starting views
    [lo, hi) for pipelines 1-2, SfM points [tb, te) for pipeline 3 — and ONE all-gather per pipeline puts the accepted
    points of all ranks back into the reference's loop order on every rank, so what follows (density limiter, filter) sees
    exactly the single-GPU sequence.  `spacing` = eg3d_params.split_interval_distance of the scene."""
    world = dist.get_world_size() if dist is not None else 1
    out, tms = [], []
    if world == 1:
        for call in (lambda: matcher.match_polyline_sets(cands1), lambda: matcher.match_polyline_sets(cands2), lambda: matcher.match_refpoints(0, scene.n_tracks)):
            r = call()
            pts, tm = r if isinstance(r, tuple) else (r, None)
            out.append(pts); tms.append(tm)
        return out, tms
    from . import multigpu as mg
    rank = dist.get_rank()
    lo, hi = mg.view_block(scene.n_views, world, rank)
    in_library = getattr(matcher, "has_comm", False)      # a lib.DeviceScene after comm_create: the exchange runs inside libeg3d.so (NCCL + device merge)
    for cands in (cands1, cands2):
        views = mg.polyline_set_seed_views(scene, cands, spacing, E.sample_seeds)
        keys = np.where((views >= lo) & (views < hi))[0]
        if in_library:
            dp, tm = matcher.match_polyline_sets(cands, lo, hi, fetch=False)
            merged, _ = matcher.points_allgather(dp, keys, fetch=True)
            dp.free()
        else:
            r = matcher.match_polyline_sets(cands, lo, hi)
            pts, tm = r if isinstance(r, tuple) else (r, None)
            merged, _ = mg.all_gather_points(pts, dist, device=device, order_keys=keys)
        out.append(merged); tms.append(tm)
    tb, te = mg.track_block(scene.n_tracks, world, rank)
    if in_library:
        dp, tm = matcher.match_refpoints(tb, te, fetch=False)
        merged, _ = matcher.points_allgather(dp, None, fetch=True)      # track blocks in rank order ARE the reference's loop order
        dp.free()
    else:
        r = matcher.match_refpoints(tb, te)
        pts, tm = r if isinstance(r, tuple) else (r, None)
        merged, _ = mg.all_gather_points(pts, dist, device=device)
    out.append(merged); tms.append(tm)
    return out, tms


def add_points_to_tracks(scene, pts, keep):
    """add_3dpoints_to_sfmd (output_utilities.cpp:96-111): the kept edge points are appended after the SfM points."""
    kept = np.where(keep)[0]
    xyz = np.concatenate([scene.track_xyz, pts.xyz[kept]])
    lens = pts.obs_off[kept + 1] - pts.obs_off[kept]
    obs_off = np.concatenate([scene.track_off, int(scene.track_off[-1]) + np.cumsum(lens)]).astype(np.int64)
    idx = ranges(pts.obs_off[kept], lens)
    return xyz, obs_off, np.concatenate([scene.track_view, pts.obs_view[idx]]), np.concatenate([scene.track_xy, pts.obs_xy[idx]])


def edge_reconstruction(dev, scene, cands1, cands2, dist=None, device=None, spacing=20.0):
    """edge_reconstruction_pipeline + filter on a device scene -> dict with every intermediate the reference writes out.
    Multi-GPU (`dist`): the matching is sharded (run_pipelines); the density limiter is order dependent and the filter is
    cheap, so every rank runs both on the gathered points and ends with the same result."""
    parts, tms = run_pipelines(dev, scene, cands1, cands2, dist=dist, device=device, spacing=spacing)
    allp = PointSet.concat(parts)
    keep = dev.dedup_close_points(allp)                                   # pipelines.cpp:236
    xyz, obs_off, obs_view, obs_xy = add_points_to_tracks(scene, allp, keep)
    fx, inliers, tmf = dev.filter(xyz, obs_off, obs_view, obs_xy, scene.n_tracks)   # edge_matcher.cpp:132
    return dict(parts=parts, timings=tms, points=allp, keep=keep, xyz=xyz, obs_off=obs_off, obs_view=obs_view, obs_xy=obs_xy,
                filtered_xyz=fx, inliers=inliers, filter_timing=tmf)


def edge_matching(sfm_json, edges_folder, out_folder, params=None, _scene_factory=None):
    """edge_matching(emip) (edge_matcher.hpp:100-102) for an OpenMVG sfm_data JSON + a folder of edge images named like the
    views: writes <out>/before_filtering.json and <out>/output.json (edge_matcher.cpp:129, :135) and returns a summary dict."""
    import cv2
    import json
    doc = json.load(open(sfm_json))
    sfm = io.load_sfm_data(doc)
    names = {v["value"]["ptr_wrapper"]["data"]["id_pose"]: v["value"]["ptr_wrapper"]["data"]["filename"] for v in doc["views"]}
    imgs = []
    for key in sfm["view_keys"]:
        im = cv2.imread(os.path.join(edges_folder, names[key]), cv2.IMREAD_COLOR)
        if im is None:
            raise FileNotFoundError(os.path.join(edges_folder, names[key]))
        imgs.append(im)
    scene, _ = real_scene.scene_from_parts(sfm, imgs)
    cands1, cands2, info = candidate_sets(scene)
    prm = params if params is not None else E.default_params(**REAL_DATA_CAPACITIES)
    for attempt in range(4):
        try:
            with (_scene_factory or E.DeviceScene)(scene, prm) as dev:      # _scene_factory: injection point for the tests only; the product always runs E.DeviceScene (no CPU path here)
                r = edge_reconstruction(dev, scene, cands1, cands2)
            break
        except E.Eg3dError as e:
            # a per-seed capacity was too small for this input (the library reports it, it never truncates): double both and
            # run again on a fresh scene handle
            from . import _abi as A
            # (the device seeding's neighbourhood bound, A6_RAW / A6_IDS, is a different capacity: larger chain slots cannot help it)
            if e.status != A.EG3D_ERR_CAPACITY or attempt == 3 or "A6_" in str(e):
                raise
            prm.max_chain_points *= 2
            prm.max_follow_points *= 2
    info["max_chain_points"], info["max_follow_points"] = int(prm.max_chain_points), int(prm.max_follow_points)
    os.makedirs(out_folder, exist_ok=True)
    n_before = io.save_sfm_data(os.path.join(out_folder, "before_filtering.json"), doc, r["xyz"], r["obs_off"], r["obs_view"], r["obs_xy"], view_keys=sfm["view_keys"])
    n_after = io.save_sfm_data(os.path.join(out_folder, "output.json"), doc, r["filtered_xyz"], r["obs_off"], r["obs_view"], r["obs_xy"], inliers=r["inliers"],
                               view_keys=sfm["view_keys"])
    info.update(points_per_pipeline=[p.n_points for p in r["parts"]], kept_after_density_limiter=int(r["keep"].sum()), points_before_filtering=n_before,
                points_after_filtering=n_after)
    return info


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser(description="edge images + OpenMVG sfm_data JSON -> 3D edge points (the reference's EdgeGraph3D inputs / outputs)")
    ap.add_argument("edges_folder"); ap.add_argument("sfm_json"); ap.add_argument("out_folder")
    a = ap.parse_args()
    print(json.dumps(edge_matching(a.sfm_json, a.edges_folder, a.out_folder)))
