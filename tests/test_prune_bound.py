"""Soundness of the exact-safe pruning bound used by K3 (eg3d_dev.cuh: pair_cannot_fit).

Claim: for two observations pa (view a), pb (view b) and ANY 3D point X, the 2-view cost |pa-qa|^2 + |pb-qb|^2 of its
reprojections is >= T^2 whenever  |s0| - g T > 0  and  |s0| - g T >= T (n0 + h T)  with the camera-derived F.  The test
mirrors the formula in numpy (double) and checks it against brute-force minimisation over X: the bound must never fire
for a pair whose best achievable cost is below T^2 (no false rejections => results identical to solving every problem)."""
import numpy as np
from edgegraph3d_b200 import lib as E, synthetic as syn


def cannot_fit(F, pa, pb, T):
    F = F.reshape(3, 3)
    l = F @ np.array([pa[0], pa[1], 1.0])
    s0 = abs(pb[0] * l[0] + pb[1] * l[1] + l[2])
    n0 = np.hypot(l[0], l[1])
    m = F.T @ np.array([pb[0], pb[1], 1.0])
    g = np.hypot(m[0], m[1])
    h = np.sqrt((F[:2, :2] ** 2).sum())
    num = s0 - g * T
    return num > 0 and num >= T * (n0 + h * T)


def best_two_view_cost(Pa, Pb, pa, pb, X0):
    """Gauss-Newton on the 2-view reprojection cost from several starts; returns the smallest cost found."""
    best = np.inf
    for jitter in (0.0, 0.05, 0.3):
        X = X0 + jitter * np.array([0.3, -0.2, 0.1])
        for _ in range(60):
            r, J = [], []
            for P, p in ((Pa, pa), (Pb, pb)):
                h = P[:, :3] @ X + P[:, 3]
                r += [p[0] - h[0] / h[2], p[1] - h[1] / h[2]]
                J += [(P[0, :3] * h[2] - P[2, :3] * h[0]) / h[2] ** 2, (P[1, :3] * h[2] - P[2, :3] * h[1]) / h[2] ** 2]
            r, J = np.array(r), np.array(J)
            best = min(best, float(r @ r))
            try:
                X = X + np.linalg.lstsq(J, r, rcond=None)[0]
            except np.linalg.LinAlgError:
                break
    return best


def test_camera_fundamentals_epipolar_constraint():
    sc = syn.make_scene(n_views=9, n_curves=2, seed=8)
    F = E.camera_fundamentals(sc.cameras).reshape(9, 9, 3, 3)
    P = sc.cameras.astype(np.float64).reshape(9, 3, 4)
    rng = np.random.default_rng(1)
    for _ in range(300):
        a, b = rng.choice(9, 2, replace=False)
        X = np.append(rng.uniform(-1, 1, 3), 1.0)
        xa, xb = P[a] @ X, P[b] @ X
        xa, xb = xa / xa[2], xb / xb[2]
        l = F[a, b] @ xa
        assert abs(xb @ l) / np.hypot(l[0], l[1]) < 1e-8
    assert np.allclose(F[np.arange(9), np.arange(9)], 0)


def test_bound_never_rejects_a_feasible_pair():
    sc = syn.make_scene(n_views=8, n_curves=2, seed=4)
    F = E.camera_fundamentals(sc.cameras).reshape(8, 8, 9)
    P = sc.cameras.astype(np.float64).reshape(8, 3, 4)
    rng = np.random.default_rng(7)
    fired = 0
    for trial in range(400):
        a, b = rng.choice(8, 2, replace=False)
        X = rng.uniform(-0.7, 0.7, 3)
        pa = (lambda h: h[:2] / h[2])(P[a] @ np.append(X, 1))
        pb = (lambda h: h[:2] / h[2])(P[b] @ np.append(X, 1))
        # displace pb so that the pair sits around the decision boundary for this T
        T = float(np.sqrt(9.0 * 2 * rng.integers(3, 60) * 1.02))
        pb = pb + rng.normal(size=2) * rng.uniform(0.2, 3.0) * T
        if cannot_fit(F[a, b], pa, pb, T):
            fired += 1
            cost = best_two_view_cost(P[a], P[b], pa, pb, X)
            assert cost >= T * T * (1 - 1e-9), (cost, T * T)
    assert fired > 40          # the bound does fire on clearly inconsistent pairs
