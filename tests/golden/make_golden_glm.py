"""Generates tests/golden/glm_golden.npz with the reference's OWN vendored glm (run in the build container):

  make -C oracle ref && python tests/golden/make_golden_glm.py

oracle/_ref/libglm_probe.so (oracle/glm_probe.cpp compiled against /root/reference/external/glm) evaluates, with real glm
types, the camera-matrix construction of OpenMvgParser.cpp:107-125, 252-256, 280-289 and compute_projection
(geometric_utilities.cpp:973-977).  Known answers: 25 dtu006 cameras + 300 random ones (rotation, centre, focal, principal point
-> translation, cameraMatrix rows 0..2) and 2 000 projections (camera, X -> xy)."""
import ctypes as C
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from edgegraph3d_b200 import _abi as A  # noqa: E402

if __name__ == "__main__":
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libglm_probe.so"))
    L.eg3d_ref_glm_camera.argtypes = [A.c_f32p, A.c_f32p, C.c_float, C.c_float, C.c_float, A.c_f32p, A.c_f32p]
    L.eg3d_ref_glm_project.argtypes = [A.c_f32p, A.c_f32p, A.c_f32p]
    f32 = np.float32
    rng = np.random.default_rng(42)
    cases = []
    src = "/root/reference/example/dtu006/input.json"
    if os.path.exists(src):
        doc = json.load(open(src))
        intr = doc["intrinsics"][0]["value"]["ptr_wrapper"]["data"]
        for ex in doc["extrinsics"]:
            cases.append((f32(ex["value"]["rotation"]).reshape(9), f32(ex["value"]["center"]), f32(intr["focal_length"]),
                          f32(intr["principal_point"][0]), f32(intr["principal_point"][1])))
    for _ in range(300):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        cases.append((q.astype(f32).reshape(9), (rng.normal(size=3) * 3).astype(f32), f32(rng.uniform(500, 3000)), f32(rng.uniform(300, 1000)),
                      f32(rng.uniform(200, 700))))
    rot, cen, foc, ppx, ppy, cam, tr = [], [], [], [], [], [], []
    for R, Cc, f, px, py in cases:
        R, Cc = np.ascontiguousarray(R), np.ascontiguousarray(Cc)
        out, t = np.zeros(12, f32), np.zeros(3, f32)
        L.eg3d_ref_glm_camera(A.ptr(R, A.c_f32p), A.ptr(Cc, A.c_f32p), f, px, py, A.ptr(out, A.c_f32p), A.ptr(t, A.c_f32p))
        rot.append(R); cen.append(Cc); foc.append(f); ppx.append(px); ppy.append(py); cam.append(out); tr.append(t)
    cam = np.array(cam, f32)
    pc, px3, pxy = [], [], []
    for _ in range(2000):
        k = int(rng.integers(len(cam)))
        X = (rng.normal(size=3) * 1.5).astype(f32)
        o = np.zeros(2, f32)
        L.eg3d_ref_glm_project(A.ptr(np.ascontiguousarray(cam[k]), A.c_f32p), A.ptr(X, A.c_f32p), A.ptr(o, A.c_f32p))
        pc.append(k); px3.append(X); pxy.append(o)
    out = os.path.join(ROOT, "tests", "golden", "glm_golden.npz")
    np.savez_compressed(out, rot=np.array(rot, f32), center=np.array(cen, f32), focal=np.array(foc, f32), ppx=np.array(ppx, f32), ppy=np.array(ppy, f32),
                        camera=cam, translation=np.array(tr, f32), proj_cam=np.array(pc, np.int32), proj_X=np.array(px3, f32), proj_xy=np.array(pxy, f32))
    print(out, len(cam), "cameras", len(pc), "projections", os.path.getsize(out), "bytes")
