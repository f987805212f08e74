"""OpenMVG `sfm_data` JSON in / out and PLY out plus the track-based fundamental matrices (row f4):
the data formats either side of the hot path, host-side Python.

  load_sfm_data   external/manifoldReconstructor/src/OpenMvgParser.cpp:75-153, 241-301 (camera conventions: SURVEY A.1)
  save_sfm_data   src/edgegraph3d/io/output/output_sfm_data.cpp:186-229 (+ add_3dpoints_to_sfmd: new points appended)
  write_ply       src/edgegraph3d/io/output/output_point_cloud.cpp
  fundamental_from_tracks  src/edgegraph3d/utils/geometry/geometric_utilities.cpp:754-820 (cv::findFundamentalMat, LMedS)
"""
import json
import numpy as np


def _f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)          # JSON doubles narrowed with GetFloat


def glm_camera_matrix(R, C, K):
    """translation = -center * rotation (OpenMvgParser.cpp:289) and cameraMatrix = eMatrix * kMatrix (:107-125) in float32, in
    the operation order of the reference's vendored glm 0.9.6 (vec3 * mat3: m[i][0] v.x + m[i][1] v.y + m[i][2] v.z;
    mat4 * mat4: column c = m1[0] m2[c][0] + m1[1] m2[c][1] + m1[2] m2[c][2] + m1[3] m2[c][3], left to right, the zero terms
    included).  A BLAS product rounds differently in the last bit of some entries (17 of dtu006's 25 cameras); this form is
    bit-identical to real glm (tests/golden/glm_golden.npz, made by the probe oracle/glm_probe.cpp).  -> (P [3,4], t [3])"""
    f32 = np.float32
    R = np.asarray(R, f32).reshape(3, 3)
    nc = -np.asarray(C, f32)
    t = np.array([f32(f32(f32(R[i, 0] * nc[0]) + f32(R[i, 1] * nc[1])) + f32(R[i, 2] * nc[2])) for i in range(3)], f32)
    E = np.zeros((4, 4), f32); E[:3, :3] = R; E[:3, 3] = t; E[3, 3] = 1        # glm object: E[r][c]
    Km = np.zeros((4, 4), f32); Km[:3, :3] = np.asarray(K, f32).reshape(3, 3)
    res = np.zeros((4, 4), f32)
    for c in range(4):
        acc = (E[0] * Km[c, 0]).astype(f32)
        for k in (1, 2, 3):
            acc = (acc + (E[k] * Km[c, k]).astype(f32)).astype(f32)
        res[c] = acc
    return res[:3].copy(), t


def load_sfm_data(path_or_dict):
    """-> dict(width, height, cameras [V,12] f32 = P = K[R|t] (rows 0..2 of cameraMatrix), K, R, center, t,
    track_xyz [N,3] f32, track_off [N+1] i64, track_view [M] i32, track_xy [M,2] f32, view_keys)."""
    d = path_or_dict if isinstance(path_or_dict, dict) else json.load(open(path_or_dict))
    intr = {}
    for it in d["intrinsics"]:
        data = it["value"]["ptr_wrapper"]["data"]
        f = _f32(data["focal_length"]); pp = _f32(data["principal_point"])
        K = np.zeros((3, 3), np.float32)
        K[0, 0] = f; K[1, 1] = f; K[0, 2] = pp[0]; K[1, 2] = pp[1]; K[2, 2] = 1.0   # radial distortion is ignored (:252-256)
        intr[it["key"]] = (K, int(data["width"]), int(data["height"]))
    views = {v["value"]["ptr_wrapper"]["data"]["id_pose"]: v["value"]["ptr_wrapper"]["data"] for v in d["views"]}
    cams, Ks, Rs, Cs, ts, keys = [], [], [], [], [], []
    pos_of_pose = {}
    for pos, ex in enumerate(d["extrinsics"]):                    # view index = position in `extrinsics` (:292)
        key = ex["key"]
        pos_of_pose[key] = pos
        R = _f32(ex["value"]["rotation"]).reshape(3, 3)
        C = _f32(ex["value"]["center"])
        K, w, h = intr[views[key]["id_intrinsic"]] if key in views else next(iter(intr.values()))
        P, t = glm_camera_matrix(R, C, K)
        cams.append(P.reshape(12)); Ks.append(K); Rs.append(R); Cs.append(C); ts.append(t); keys.append(key)
    xyz, off, tv, txy = [], [0], [], []
    for pt in d["structure"]:
        xyz.append(_f32(pt["value"]["X"]))
        for ob in pt["value"]["observations"]:
            tv.append(pos_of_pose[ob["key"]])
            txy.append(_f32(ob["value"]["x"]))
        off.append(len(tv))
    K0, w, h = next(iter(intr.values()))
    return dict(width=w, height=h, cameras=np.array(cams, np.float32), K=np.array(Ks), R=np.array(Rs), center=np.array(Cs), t=np.array(ts),
                track_xyz=np.array(xyz, np.float32).reshape(-1, 3), track_off=np.array(off, np.int64), track_view=np.array(tv, np.int32),
                track_xy=np.array(txy, np.float32).reshape(-1, 2), view_keys=keys)


def fundamental_from_tracks(n_views, track_off, track_view, track_xy, min_corr=10, engine="auto"):
    """F[i][j] = cv::findFundamentalMat(points_i, points_j, FM_LMEDS) over the tracks seen by both views in ascending
    track id; pairs with < 10 common tracks are invalid (the reference's 1x1 dummy Mat).
    engine: "cv2" = the same third-party call the reference makes (reproduces the committed dtu006 fixture bit for bit);
    "native" = libeg3d.so's own LMedS (eg3d_fundamental_from_tracks, host C++, no OpenCV: same estimator family, not
    bit-identical); "auto" = cv2 when it can be imported, else native."""
    if engine != "native":
        try:
            import cv2
        except ImportError:
            if engine == "cv2":
                raise
            engine = "native"
    if engine == "native":
        from . import lib as E
        return E.fundamental_from_tracks(n_views, track_off, track_view, track_xy, min_corr)
    seen = [dict() for _ in range(n_views)]
    for p in range(len(track_off) - 1):
        for o in range(int(track_off[p]), int(track_off[p + 1])):
            seen[int(track_view[o])][p] = o                       # last observation of a view wins, as get_2d_coordinates does
    F = np.zeros((n_views, n_views, 9), np.float64)
    valid = np.zeros((n_views, n_views), np.uint8)
    for i in range(n_views):
        for j in range(n_views):
            if i == j:
                continue
            common = sorted(set(seen[i]) & set(seen[j]))
            if len(common) < min_corr:
                continue
            pi = np.array([track_xy[seen[i][p]] for p in common], np.float32)
            pj = np.array([track_xy[seen[j][p]] for p in common], np.float32)
            Fm, _ = cv2.findFundamentalMat(pi, pj, cv2.FM_LMEDS)
            if Fm is None or Fm.shape != (3, 3):
                continue
            F[i, j] = Fm.reshape(9); valid[i, j] = 1
    return F, valid


def save_sfm_data(path, original, xyz, obs_off, obs_view, obs_xy, inliers=None, view_keys=None):
    """output_sfm_data: views / intrinsics / extrinsics of `original` (dict or path) kept, `structure` rewritten from the
    given points (key = running id, id_feat = 0 as OUTPUT_SFMD_FEATURE_ID), optionally only the inliers."""
    d = original if isinstance(original, dict) else json.load(open(original))
    keys = view_keys if view_keys is not None else [ex["key"] for ex in d["extrinsics"]]
    st = []
    for i in range(len(obs_off) - 1):
        if inliers is not None and not inliers[i]:
            continue
        obs = [{"key": int(keys[int(obs_view[o])]), "value": {"id_feat": 0, "x": [float(obs_xy[o][0]), float(obs_xy[o][1])]}}
               for o in range(int(obs_off[i]), int(obs_off[i + 1]))]
        st.append({"key": i, "value": {"X": [float(v) for v in xyz[i]], "observations": obs}})
    out = {"sfm_data_version": d.get("sfm_data_version", "0.3"), "root_path": d.get("root_path", ""), "views": d["views"],
           "intrinsics": d["intrinsics"], "extrinsics": d["extrinsics"], "structure": st, "control_points": d.get("control_points", [])}
    with open(path, "w") as f:
        json.dump(out, f)
    return len(st)


def write_ply(path, xyz, rgb=None):
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n" % len(xyz))
        if rgb is not None:
            f.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        f.write("end_header\n")
        for i, p in enumerate(xyz):
            f.write("%g %g %g" % (p[0], p[1], p[2]) + ("" if rgb is None else " %d %d %d" % tuple(rgb[i])) + "\n")
