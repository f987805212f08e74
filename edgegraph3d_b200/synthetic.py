"""Seeded synthetic SfM scenes of the shapes BASELINE.json names (SURVEY.md §8d, C2/C3/C5 and small test rigs).

Cameras sit on rings around the origin and look at it; geometry is a set of random 3D curves (cubic Béziers and a few
closed circles) that every view sees as polylines whose interior vertices are sampled at view-specific curve parameters
(so vertices do not correspond across views, as with real edge maps); tracks are 3D points on the curves with noisy
observations; F is analytic from the float32 camera matrices (the reference's LMedS F is an input artefact, SURVEY
finding 10).  Host-side plumbing only: nothing here is on the timed path.
"""
import numpy as np
from .scene import FlatScene, CandidateSets, SeedBatch


def _look_at(C, target):
    z = target - C
    z /= np.linalg.norm(z)
    up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])  # rows: world -> camera


def make_cameras(n_views, width, height, focal, rng, per_ring=50, radius=4.0):
    n_rings = max(1, (n_views + per_ring - 1) // per_ring)
    P = np.zeros((n_views, 3, 4))
    K = np.array([[focal, 0, width / 2.0], [0, focal, height / 2.0], [0, 0, 1.0]])
    for v in range(n_views):
        ring, k = divmod(v, per_ring)
        n_in_ring = min(per_ring, n_views - ring * per_ring)
        el = 0.0 if n_rings == 1 else np.deg2rad(-25.0 + 50.0 * ring / (n_rings - 1))
        az = 2 * np.pi * (k + 0.37 * ring) / n_in_ring
        r = radius + rng.uniform(-0.2, 0.2)
        C = r * np.array([np.cos(el) * np.cos(az), np.sin(el), np.cos(el) * np.sin(az)])
        R = _look_at(C, rng.normal(0, 0.05, 3))
        t = -R @ C
        P[v] = K @ np.concatenate([R, t[:, None]], axis=1)
    return P.astype(np.float32)


def fundamental_from_cameras(P32):
    """F[a][b] x_a = epipolar line in view b; analytic, from the float32 matrices taken as doubles."""
    P = P32.astype(np.float64).reshape(-1, 3, 4)
    V = P.shape[0]
    M = P[:, :, :3]
    Cc = -np.linalg.solve(M, P[:, :, 3:4])[:, :, 0]            # camera centres [V,3]
    Ch = np.concatenate([Cc, np.ones((V, 1))], axis=1)          # [V,4]
    Pinv = np.linalg.pinv(P)                                   # [V,4,3]
    F = np.zeros((V, V, 3, 3))
    for a in range(V):
        e = P @ Ch[a]                                          # epipoles of a in every b [V,3]
        ex = np.zeros((V, 3, 3))
        ex[:, 0, 1], ex[:, 0, 2] = -e[:, 2], e[:, 1]
        ex[:, 1, 0], ex[:, 1, 2] = e[:, 2], -e[:, 0]
        ex[:, 2, 0], ex[:, 2, 1] = -e[:, 1], e[:, 0]
        Fa = ex @ (P @ Pinv[a])                                # [V,3,3]
        nrm = np.linalg.norm(Fa.reshape(V, -1), axis=1)
        nrm[nrm == 0] = 1
        F[a] = Fa / nrm[:, None, None]
    valid = np.ones((V, V), np.uint8)
    valid[np.arange(V), np.arange(V)] = 0
    return F.reshape(V, V, 9), valid


def _bezier(cp, t):
    t = t[..., None]
    return ((1 - t) ** 3) * cp[0] + 3 * ((1 - t) ** 2) * t * cp[1] + 3 * (1 - t) * t * t * cp[2] + (t ** 3) * cp[3]


def make_scene(n_views=6, width=640, height=480, focal=520.0, n_curves=24, segs_per_curve=24, curve_len=0.9,
               seed=0, n_tracks=0, track_cap=30, closed_frac=0.1, vertex_jitter=0.3, vertex_noise_px=0.03,
               extent=0.75, per_ring=50, track_noise_px=0.3, drop_view_frac=0.0,
               cameras=None, fundamental=None, fundamental_valid=None, real_tracks=None, centers=None):
    """`cameras` ([V,12] f32, e.g. from openmvg_io.load_sfm_data) replaces the synthetic rig; `fundamental` /
    `fundamental_valid` replace the analytic F (e.g. LMedS matrices estimated from the tracks, as the reference does);
    `real_tracks` = (xyz, off, view, xy) replaces the generated tracks; `centers` ([K,3]) are the 3D positions the
    curves are scattered around (the SfM points of a real scene) instead of the uniform box."""
    rng = np.random.default_rng(seed)
    if cameras is not None:
        P32 = np.ascontiguousarray(cameras, np.float32).reshape(-1, 3, 4)
        n_views = P32.shape[0]
    else:
        P32 = make_cameras(n_views, width, height, focal, rng, per_ring=per_ring)
    P = P32.astype(np.float64)
    F, Fv = fundamental_from_cameras(P32)
    if fundamental is not None:
        F, Fv = np.asarray(fundamental, np.float64).reshape(n_views, n_views, 9), np.asarray(fundamental_valid, np.uint8).reshape(n_views, n_views)
    n = segs_per_curve
    curves = []
    for c in range(n_curves):
        if centers is not None:
            ctr = np.asarray(centers[rng.integers(len(centers))], np.float64) + rng.normal(0, 0.3 * curve_len, 3)
        else:
            ctr = rng.uniform(-extent, extent, 3) * np.array([1.0, 0.6, 1.0])
        if rng.uniform() < closed_frac:
            a = rng.normal(size=3); a /= np.linalg.norm(a)
            b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
            curves.append(("circle", ctr, a, b, curve_len / (2 * np.pi)))
        else:
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            cp = np.stack([ctr + curve_len * ((i / 3.0) - 0.5) * d + curve_len * 0.22 * rng.normal(size=3) for i in range(4)])
            curves.append(("bezier", cp))

    def eval_curve(cv, t):
        if cv[0] == "bezier":
            return _bezier(cv[1], t)
        _, ctr, a, b, r = cv
        ang = 2 * np.pi * t[..., None]
        return ctr + r * (np.cos(ang) * a + np.sin(ang) * b)

    view_poly_off = [0]
    poly_vert_off = [0]
    verts = []
    pstart, pend = [], []
    valid_mask = np.zeros((n_views, n_curves), bool)
    margin = 3.0
    for v in range(n_views):
        for c, cv in enumerate(curves):
            t = np.arange(n + 1, dtype=np.float64)
            if vertex_jitter > 0:
                t[1:-1] += rng.uniform(-vertex_jitter, vertex_jitter, n - 1)
            t /= n
            X = eval_curve(cv, t)
            h = X @ P[v, :, :3].T + P[v, :, 3]
            xy = h[:, :2] / h[:, 2:3]
            if vertex_noise_px > 0:
                xy = xy + rng.normal(0, vertex_noise_px, xy.shape)
            closed = cv[0] == "circle"
            if closed:
                xy[-1] = xy[0]
            ok = (h[:, 2] > 0.1).all() and (xy[:, 0] > margin).all() and (xy[:, 0] < width - margin).all() \
                and (xy[:, 1] > margin).all() and (xy[:, 1] < height - margin).all()
            if ok and drop_view_frac > 0 and rng.uniform() < drop_view_frac:
                ok = False
            if ok:
                verts.append(xy.astype(np.float32))
                poly_vert_off.append(poly_vert_off[-1] + n + 1)
                valid_mask[v, c] = True
            else:
                poly_vert_off.append(poly_vert_off[-1])  # invalidated polyline keeps its id
            pstart.append(2 * c)
            pend.append(2 * c if closed else 2 * c + 1)
        view_poly_off.append(len(pstart))
    verts = np.concatenate(verts) if verts else np.zeros((0, 2), np.float32)

    tr = dict(track_xyz=None, track_off=None, track_view=None, track_xy=None)
    if real_tracks is not None:
        tr = dict(track_xyz=real_tracks[0], track_off=real_tracks[1], track_view=real_tracks[2], track_xy=real_tracks[3])
    elif n_tracks > 0:
        xyz, off, tv, txy = [], [0], [], []
        for _ in range(n_tracks):
            c = rng.integers(n_curves)
            X = eval_curve(curves[c], np.array([rng.uniform(0.05, 0.95)]))[0]
            h = X @ P[:, :, :3].transpose(0, 2, 1) + P[:, :, 3]          # [V,3]
            xy = h[:, :2] / h[:, 2:3]
            vis = np.where((h[:, 2] > 0.1) & (xy[:, 0] > margin) & (xy[:, 0] < width - margin) & (xy[:, 1] > margin)
                           & (xy[:, 1] < height - margin) & valid_mask[:, c])[0]
            if len(vis) < 3:
                continue
            if len(vis) > track_cap:
                vis = np.sort(rng.choice(vis, track_cap, replace=False))
            xyz.append(X + rng.normal(0, 0.002, 3))
            tv.extend(vis.tolist())
            txy.append(xy[vis] + rng.normal(0, track_noise_px, (len(vis), 2)))
            off.append(len(tv))
        if xyz:
            tr = dict(track_xyz=np.array(xyz, np.float32), track_off=np.array(off, np.int64),
                      track_view=np.array(tv, np.int32), track_xy=np.concatenate(txy).astype(np.float32))
    sc = FlatScene(width, height, P32.reshape(n_views, 12), F, Fv, np.array(view_poly_off, np.int64),
                   np.array(poly_vert_off, np.int64), verts, np.array(pstart, np.uint32), np.array(pend, np.uint32), **tr)
    sc.meta = dict(valid_mask=valid_mask, n_curves=n_curves, seed=seed, segs_per_curve=n)
    return sc


def curve_candidate_sets(scene, distractor_frac=0.3, seed=0):
    """One candidate set per curve: the curve's own polyline in each view where it is valid, plus now and then a
    distractor — the role `polyline_matching_*` plays upstream."""
    rng = np.random.default_rng(seed + 7)
    vm = scene.meta["valid_mask"]
    V, NC = vm.shape
    sets = []
    for c in range(NC):
        per_view = []
        for v in range(V):
            ids = [c] if vm[v, c] else []
            if rng.uniform() < distractor_frac:
                d = int(rng.integers(NC))
                if vm[v, d]:
                    ids.append(d)
            per_view.append(ids)
        sets.append(per_view)
    return CandidateSets.from_lists(sets, V)


def sample_seeds(sampler, scene, per_view=None, spacing=20.0, views=None, polylines_per_view=None, cand_set=None):
    """Seeds every `spacing` px along polylines, first polylines first, capped at per_view.
    `sampler(scene, views, polylines, spacing)` -> (view, polyline, segment, xy, src) is the a3 seed sampler."""
    V = scene.n_views
    views = range(V) if views is None else views
    vs, pls = [], []
    for v in views:
        npl = scene.n_polylines(v) if polylines_per_view is None else min(polylines_per_view, scene.n_polylines(v))
        vs.extend([v] * npl)
        pls.extend(range(npl))
    view, pl, seg, xy, src = sampler(scene, np.array(vs, np.int32), np.array(pls, np.uint32), spacing)
    if per_view is not None:
        keep = np.zeros(len(view), bool)
        for v in views:
            idx = np.where(view == v)[0][:per_view]
            keep[idx] = True
        view, pl, seg, xy = view[keep], pl[keep], seg[keep], xy[keep]
    cs = None if cand_set is None else np.full(len(view), cand_set, np.int32)
    return SeedBatch(view, pl, seg, xy, cs)


def gn_microbench_inputs(scene, n_hyp, obs_per_hyp=20, seed=99, noise_px=0.5, outlier_frac=0.1):
    """BASELINE config 5: random 3D points, `obs_per_hyp` views each, noisy observations,
    10 % of the hypotheses get one observation displaced by U[5,50] px, init = truth + N(0, 0.02)."""
    rng = np.random.default_rng(seed)
    V = scene.n_views
    P = scene.cameras.astype(np.float64).reshape(V, 3, 4)
    X = rng.uniform(-0.8, 0.8, (n_hyp, 3)) * np.array([1.0, 0.6, 1.0])
    k = min(obs_per_hyp, V)
    views = np.argsort(rng.random((n_hyp, V)), axis=1)[:, :k].astype(np.int32)
    views.sort(axis=1)
    Pv = P[views]                                               # [n,k,3,4]
    h = np.einsum("nkij,nj->nki", Pv[..., :3], X) + Pv[..., 3]
    xy = h[..., :2] / h[..., 2:3] + rng.normal(0, noise_px, (n_hyp, k, 2))
    out = rng.random(n_hyp) < outlier_frac
    which = rng.integers(0, k, n_hyp)
    d = rng.uniform(5, 50, n_hyp) * rng.choice([-1.0, 1.0], n_hyp)
    xy[out, which[out], 0] += d[out]
    init = X + rng.normal(0, 0.02, X.shape)
    return views, xy.astype(np.float32), init.astype(np.float32), X.astype(np.float32)
