"""Generate tests/golden/cv2_golden.npz: known-answer vectors for the OpenCV boundary of the hot path, produced by
REAL cv2 calls (run in the build container, cv2 4.13.0).  The reference has no tests or golden vectors of its own
(SURVEY.md §4), and OpenCV is its only third-party arithmetic on this path, so these pin the oracle.

  epiline   : cv2.computeCorrespondEpilines            <- geometric_utilities.cpp:832
  dlt       : cv2.triangulatePoints                    <- triangulation.cpp:216,290
  gn64      : em_GaussNewton rebuilt line by line from cv2.gemm / cv2.determinant / cv2.invert on CV_64F Mats
              <- triangulation.cpp:105-176, 53-103
  gn32      : GaussNewton of the outlier filter, same construction on CV_32F Mats
              <- filtering/gauss_newton.cpp:83-134, 26-75

Usage:  python tests/golden/make_golden_cv2.py      (rewrites cv2_golden.npz next to this file)
"""
import os
import sys
import numpy as np
import cv2

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from edgegraph3d_b200 import synthetic as syn  # noqa: E402

f32 = np.float32


def em_gauss_newton_cv2(cams64, pts32, init):
    """triangulation.cpp:105-176 with the cv::Mat operations done by cv2 (CV_64F)."""
    n = len(pts32)
    X = np.array(init, np.float64).reshape(3, 1)
    last_mse = 0.0
    iters = 0
    for _ in range(30):
        mse = 0.0
        r = np.zeros((2 * n, 1), np.float64)
        XH = np.vstack([X, [[1.0]]])
        for m in range(n):
            h = cv2.gemm(cams64[m], XH, 1, None, 0)
            r[2 * m, 0] = float(pts32[m, 0]) - h[0, 0] / h[2, 0]
            mse += r[2 * m, 0] * r[2 * m, 0]
            r[2 * m + 1, 0] = float(pts32[m, 1]) - h[1, 0] / h[2, 0]
            mse += r[2 * m + 1, 0] * r[2 * m + 1, 0]
        if abs(mse / (n * 2) - last_mse) < 0.0000005:
            break
        last_mse = mse / (n * 2)
        iters += 1
        J = np.zeros((2 * n, 3), np.float64)
        for m in range(n):
            P = cams64[m]
            h = cv2.gemm(P, XH, 1, None, 0)
            xH, yH, zH = h[0, 0], h[1, 0], h[2, 0]
            for c in range(3):
                J[2 * m, c] = (P[0, c] * zH - P[2, c] * xH) / (zH * zH)
                J[2 * m + 1, c] = (P[1, c] * zH - P[2, c] * yH) / (zH * zH)
        H = cv2.gemm(J, J, 1, None, 0, flags=cv2.GEMM_1_T)
        if cv2.determinant(H) < 0.00001:
            return -1, X.ravel().copy(), last_mse, iters
        Hinv = cv2.invert(H)[1]
        M = cv2.gemm(Hinv, J, 1, None, 0, flags=cv2.GEMM_2_T)
        X = X + cv2.gemm(M, r, 1, None, 0)
    return (1 if last_mse < 9 else -1), X.ravel().copy(), last_mse, iters


def gauss_newton_f32_cv2(cams32, pts32, init, gn_max_mse=f32(2.25)):
    """filtering/gauss_newton.cpp:83-134 with the cv::Mat operations done by cv2 (CV_32F); scalar code in float32."""
    n = len(pts32)
    X = np.array(init, f32).reshape(3, 1)
    last_mse = f32(0)
    iters = 0
    for _ in range(30):
        mse = f32(0)
        r = np.zeros((2 * n, 1), f32)
        XH = np.vstack([X, np.array([[1.0]], f32)]).astype(f32)
        for m in range(n):
            h = cv2.gemm(cams32[m], XH, 1, None, 0)
            r[2 * m, 0] = f32(pts32[m, 0] - f32(h[0, 0] / h[2, 0]))
            mse = f32(mse + f32(r[2 * m, 0] * r[2 * m, 0]))
            r[2 * m + 1, 0] = f32(pts32[m, 1] - f32(h[1, 0] / h[2, 0]))
            mse = f32(mse + f32(r[2 * m + 1, 0] * r[2 * m + 1, 0]))
        diff = f32(f32(mse / f32(n * 2)) - last_mse)
        if abs(float(diff)) < 0.0000000005:
            break
        last_mse = f32(mse / f32(n * 2))
        iters += 1
        J = np.zeros((2 * n, 3), f32)
        for m in range(n):
            P = cams32[m]
            h = cv2.gemm(P, XH, 1, None, 0)
            xH, yH, zH = h[0, 0], h[1, 0], h[2, 0]
            zz = f32(zH * zH)
            for c in range(3):
                J[2 * m, c] = f32(f32(f32(P[0, c] * zH) - f32(P[2, c] * xH)) / zz)
                J[2 * m + 1, c] = f32(f32(f32(P[1, c] * zH) - f32(P[2, c] * yH)) / zz)
        H = cv2.gemm(J, J, 1, None, 0, flags=cv2.GEMM_1_T)
        d = f32(cv2.determinant(H))
        if float(d) < 0.0000000001:
            return -1, X.ravel().copy(), last_mse, iters
        Hinv = cv2.invert(H)[1]
        M = cv2.gemm(Hinv, J, 1, None, 0, flags=cv2.GEMM_2_T)
        X = (X + cv2.gemm(M, r, 1, None, 0)).astype(f32)
    return (1 if last_mse < gn_max_mse else -1), X.ravel().copy(), last_mse, iters


def main():
    rng = np.random.default_rng(20260101)
    out = {"cv2_version": np.array(cv2.__version__)}
    sc = syn.make_scene(n_views=12, n_curves=4, seed=5)
    P32 = sc.cameras.reshape(-1, 3, 4)
    V = P32.shape[0]

    # --- epilines
    n = 600
    Fs = sc.fundamental.reshape(V * V, 9)[rng.integers(0, V * V, n)]
    Fs = np.where(np.abs(Fs).sum(1, keepdims=True) == 0, rng.normal(size=(n, 9)), Fs) * rng.choice([1.0, -3.7, 1e3], (n, 1))
    pts = (rng.uniform(0, 1, (n, 2)) * [sc.width, sc.height]).astype(f32)
    lines = np.zeros((n, 3), f32)
    for i in range(n):
        lines[i] = cv2.computeCorrespondEpilines(pts[i].reshape(1, 1, 2), 1, Fs[i].reshape(3, 3)).reshape(3)
    out.update(epi_F=Fs, epi_pts=pts, epi_lines=lines)

    # --- triangulatePoints
    n = 400
    va = rng.integers(0, V, n)
    vb = (va + rng.integers(1, V, n)) % V
    X = rng.uniform(-0.7, 0.7, (n, 3))
    def proj(v, X):
        h = P32[v].astype(np.float64)[:, :3] @ X + P32[v].astype(np.float64)[:, 3]
        return h[:2] / h[2]
    x1 = np.array([proj(va[i], X[i]) for i in range(n)]) + rng.normal(0, 0.7, (n, 2))
    x2 = np.array([proj(vb[i], X[i]) for i in range(n)]) + rng.normal(0, 0.7, (n, 2))
    x1, x2 = x1.astype(f32), x2.astype(f32)
    X4 = np.zeros((n, 4), f32)
    for i in range(n):
        X4[i] = cv2.triangulatePoints(P32[va[i]], P32[vb[i]], x1[i].reshape(2, 1), x2[i].reshape(2, 1)).reshape(4)
    out.update(dlt_cams=P32.reshape(V, 12), dlt_va=va.astype(np.int32), dlt_vb=vb.astype(np.int32), dlt_x1=x1, dlt_x2=x2, dlt_X4=X4)

    # --- GN cases: CSR of observations
    def gn_cases(n_cases, fp64):
        offs, views, xys, inits, oks, Xs, mses, its = [0], [], [], [], [], [], [], []
        for c in range(n_cases):
            k = int(rng.integers(2, 13))
            vs = np.sort(rng.choice(V, k, replace=False))
            if c % 17 == 0 and k >= 3:
                vs[1] = vs[0]  # repeated camera
            Xt = rng.uniform(-0.7, 0.7, 3)
            xy = np.array([proj(v, Xt) for v in vs]) + rng.normal(0, rng.choice([0.3, 1.5, 4.0]), (k, 2))
            if c % 5 == 0:
                xy[rng.integers(k)] += rng.uniform(-60, 60, 2)  # outlier -> mostly rejected
            xy = xy.astype(f32)
            init = Xt + rng.normal(0, rng.choice([0.01, 0.1, 0.6]), 3)
            if fp64:
                init = init.astype(f32).astype(np.float64)
                cams = [np.vstack([P32[v], np.zeros((1, 4), f32)]).astype(np.float64) for v in vs]
                ok, Xo, lm, it = em_gauss_newton_cv2(cams, xy, init)
            else:
                init = init.astype(f32)
                cams = [np.vstack([P32[v], np.zeros((1, 4), f32)]).astype(f32) for v in vs]
                ok, Xo, lm, it = gauss_newton_f32_cv2(cams, xy, init)
            views.extend(vs.tolist()); xys.append(xy); offs.append(len(views)); inits.append(init)
            oks.append(ok == 1); Xs.append(Xo); mses.append(lm); its.append(it)
        return (np.array(offs, np.int64), np.array(views, np.int32), np.concatenate(xys).astype(f32), np.array(inits),
                np.array(oks, np.uint8), np.array(Xs), np.array(mses), np.array(its, np.int32))
    for name, fp64 in (("gn64", True), ("gn32", False)):
        o, v, xy, init, ok, Xo, lm, it = gn_cases(400, fp64)
        out.update({f"{name}_off": o, f"{name}_view": v, f"{name}_xy": xy, f"{name}_init": init, f"{name}_ok": ok,
                    f"{name}_X": Xo, f"{name}_mse": lm, f"{name}_iters": it})
    out["cams"] = P32.reshape(V, 12)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})
    print("gn64 ok", out["gn64_ok"].mean(), "gn32 ok", out["gn32_ok"].mean(), "iters64", np.bincount(out["gn64_iters"]), "iters32", np.bincount(out["gn32_iters"]))


if __name__ == "__main__":
    main()
