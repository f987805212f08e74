"""Pins the float32 arithmetic that reaches the path through the reference's vendored glm (external/glm 0.9.6.3) against known
answers computed with that glm itself (tests/golden/glm_golden.npz, made by tests/golden/make_golden_glm.py through
oracle/glm_probe.cpp): the camera matrices an OpenMVG JSON turns into (OpenMvgParser.cpp:107-125, 280-289) and
compute_projection (geometric_utilities.cpp:973-977).  Bit for bit: a camera entry that is one ulp off moves every epipolar
hit and every accept / reject decision downstream.  CPU only."""
import os
import numpy as np
from edgegraph3d_b200 import lib as E, openmvg_io as io, _abi as A
from tests import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "glm_golden.npz"))


def test_camera_matrices_match_real_glm():
    for R, C, f, px, py, cam, t in zip(G["rot"], G["center"], G["focal"], G["ppx"], G["ppy"], G["camera"], G["translation"]):
        K = np.zeros((3, 3), np.float32)
        K[0, 0] = f; K[1, 1] = f; K[0, 2] = px; K[1, 2] = py; K[2, 2] = 1
        P, tt = io.glm_camera_matrix(R.reshape(3, 3), C, K)
        assert P.reshape(12).tobytes() == cam.tobytes() and tt.tobytes() == t.tobytes()


def test_dtu006_fixture_cameras_are_the_real_glm_ones():
    z = np.load(os.path.join(HERE, "golden", "dtu006_sfm.npz"))
    assert z["cameras"].astype(np.float32).tobytes() == G["camera"][:25].tobytes()      # the first 25 golden cameras are dtu006's


def test_projection_matches_real_glm_in_oracle_and_product():
    Lo, Lp = O.lib(), E.load()
    out = np.zeros(2, np.float32)
    for k, X, xy in zip(G["proj_cam"], G["proj_X"], G["proj_xy"]):
        cam, X = np.ascontiguousarray(G["camera"][k]), np.ascontiguousarray(X)
        Lo.eg3d_oracle_compute_projection(A.ptr(cam, A.c_f32p), A.ptr(X, A.c_f32p), A.ptr(out, A.c_f32p))
        assert out.tobytes() == xy.tobytes()
        Lp.eg3d_project_host(A.ptr(cam, A.c_f32p), A.ptr(X, A.c_f32p), A.ptr(out, A.c_f32p))
        assert out.tobytes() == xy.tobytes()
