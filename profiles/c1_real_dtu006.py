"""Documentation run (not a bench line): the reference's packaged example end to end on REAL data — the 25 real edge
maps (tests/golden/dtu006_edges.npz) -> polyline graphs (row f1, host) -> candidate sets (row f2, host: communities of
the polyline compatibility graph for pipeline 1, SfM-point components for pipeline 2) -> pipelines 1 + 2
(eg3d_match_polyline_sets) + pipeline 3 (eg3d_match_refpoints) on the device -> density limiter -> outlier filter, every
stage checked against the CPU oracle on the same candidate sets.  (Pipeline 1's communities come from the library's
deterministic Louvain; the reference's Grappolo is not reproducible run to run)
Usage: python profiles/c1_real_dtu006.py [out.json]"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edgegraph3d_b200 import lib as E, real_scene  # noqa: E402
from edgegraph3d_b200.scene import PointSet  # noqa: E402
from tests import oracle_lib as O  # noqa: E402


def same(g, r):
    return bool(g.n_points == r.n_points and np.array_equal(g.obs_off, r.obs_off) and np.array_equal(g.obs_view, r.obs_view)
                and np.array_equal(g.obs_poly, r.obs_poly) and np.array_equal(g.obs_seg, r.obs_seg) and g.obs_xy.tobytes() == r.obs_xy.tobytes())


def run(golden_dir, threads):
    t = time.time(); sc, plgs = real_scene.dtu006_scene(golden_dir); t_plg = time.time() - t
    t = time.time(); graph = E.SimilarityGraph(sc); com, q = graph.communities(); cands1 = graph.candidate_sets(com); t_graph = time.time() - t
    t = time.time(); cands, ref = E.polyline_sets_from_refpoints(sc); t_sets = time.time() - t
    res = {"views": sc.n_views, "tracks": sc.n_tracks, "segments_per_view": [sc.n_segments(v) for v in range(sc.n_views)],
           "host_s": {"polyline_graphs_25_views": t_plg, "compatibility_graph_and_communities": t_graph, "candidate_sets": t_sets},
           "compatibility_graph": {"nodes": len(graph.node_view), "edges": len(graph.edge_a), "communities": int(com.max()) + 1, "modularity": q},
           "candidate_sets_pipeline1": cands1.n_sets, "candidate_sets": cands.n_sets, "contributing_sfm_points": len(ref)}
    prm = E.default_params(max_chain_points=256, max_follow_points=320)   # real chains reach 136 points (default capacity 96)
    osc = O.OracleScene(sc, prm)
    with E.DeviceScene(sc, prm) as dev:
        for _ in range(2):   # second pass = warm
            t = time.time(); g1, tm1 = dev.match_polyline_sets(cands1); g2, tm2 = dev.match_polyline_sets(cands); g3, tm3 = dev.match_refpoints(0, sc.n_tracks); wall = time.time() - t
        t = time.time(); r1 = osc.match_polyline_sets(cands1, n_threads=threads); r2 = osc.match_polyline_sets(cands, n_threads=threads); r3 = osc.match_refpoints(0, sc.n_tracks, n_threads=threads); t_or = time.time() - t
        res["pipeline1"] = {"points": g1.n_points, "obs": g1.n_obs, "device_ms": tm1["total_ms"], "seeds": tm1["n_seeds"], "identical": same(g1, r1),
                            "max_abs_xyz_diff": float(np.abs(g1.xyz - r1.xyz).max()) if same(g1, r1) and g1.n_points else None}
        res["pipeline2"] = {"points": g2.n_points, "obs": g2.n_obs, "device_ms": tm2["total_ms"], "seeds": tm2["n_seeds"], "identical": same(g2, r2),
                            "max_abs_xyz_diff": float(np.abs(g2.xyz - r2.xyz).max()) if same(g2, r2) and g2.n_points else None}
        res["pipeline3"] = {"points": g3.n_points, "obs": g3.n_obs, "device_ms": tm3["total_ms"], "seeds": tm3["n_seeds"], "identical": same(g3, r3),
                            "max_abs_xyz_diff": float(np.abs(g3.xyz - r3.xyz).max()) if same(g3, r3) and g3.n_points else None}
        res["e2e_wall_ms_pipelines_1_2_3"] = wall * 1e3
        res["oracle"] = {"seconds_pipelines_1_2_3": t_or, "threads": threads}
        allp = PointSet.concat([g1, g2, g3])      # the reference's order: pipelines.cpp:217-229
        keep_g = dev.dedup_close_points(allp); keep_o = osc.dedup_close_points(allp)
        kept = np.where(keep_g)[0]
        xyz = np.concatenate([sc.track_xyz, allp.xyz[kept]])
        lens = allp.obs_off[kept + 1] - allp.obs_off[kept]
        obs_off = np.concatenate([sc.track_off, int(sc.track_off[-1]) + np.cumsum(lens)])
        idx = np.concatenate([np.arange(allp.obs_off[i], allp.obs_off[i + 1]) for i in kept])
        obs_view = np.concatenate([sc.track_view, allp.obs_view[idx]]); obs_xy = np.concatenate([sc.track_xy, allp.obs_xy[idx]])
        fx, inl, tmf = dev.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks)
        ox, oinl = osc.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks, n_threads=threads)[:2]
        res["density_limiter"] = {"in": allp.n_points, "kept": int(keep_g.sum()), "identical": bool(np.array_equal(keep_g, keep_o))}
        res["filter"] = {"points_in": len(xyz), "inliers": int(inl.sum()), "edge_point_inliers": int(inl[sc.n_tracks:].sum()),
                         "identical_inlier_sets": bool(np.array_equal(inl, oinl)), "identical_refined_xyz": bool(np.array_equal(fx[inl == 1], ox[oinl == 1]))}
    return res


if __name__ == "__main__":
    res = run(os.path.join(ROOT, "tests", "golden"), os.cpu_count())
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"))
