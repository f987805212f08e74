"""Generates tests/golden/dtu006_sfm.npz from the reference's packaged example (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz):

  python tests/golden/make_dtu006_fixture.py [/root/reference/example/dtu006/input.json]

Contents: the 25 camera matrices P = K[R|t] (float32), image size, the 6268 SfM tracks (xyz, CSR of
(view, xy) observations) and the fundamental matrices the reference would use for this input: cv2.findFundamentalMat
(FM_LMEDS) on the common tracks of every ordered view pair (geometric_utilities.cpp:754-820), cv2 version recorded."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from edgegraph3d_b200 import openmvg_io as io  # noqa: E402

if __name__ == "__main__":
    import cv2
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/example/dtu006/input.json"
    d = io.load_sfm_data(src)
    V = d["cameras"].shape[0]
    F, valid = io.fundamental_from_tracks(V, d["track_off"], d["track_view"], d["track_xy"])
    out = os.path.join(ROOT, "tests", "golden", "dtu006_sfm.npz")
    np.savez_compressed(out, cameras=d["cameras"], width=d["width"], height=d["height"], track_xyz=d["track_xyz"], track_off=d["track_off"],
                        track_view=d["track_view"], track_xy=d["track_xy"], fundamental=F, fundamental_valid=valid, cv2_version=cv2.__version__)
    print(out, "views", V, "tracks", len(d["track_xyz"]), "obs", len(d["track_view"]), "valid F pairs", int(valid.sum()), "cv2", cv2.__version__,
          "bytes", os.path.getsize(out))
