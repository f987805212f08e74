"""Generates tests/golden/dtu006_edges.npz from the reference's packaged example (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz):

  python tests/golden/make_dtu006_edges_fixture.py [/root/reference/example/dtu006]

Contents: the 25 edge maps `edges/<filename of view v>.png` of the views of input.json (in view order) as bit-packed
boolean masks (pixel == EDGE_COLOR (255,255,255) after cv2.imread(IMREAD_COLOR), which is how the reference reads them:
edge_graph_3d_utilities.cpp:325-343, global_defines.hpp:47).  This is INPUT data of row f1 (edge image -> polyline
graph); the reference ships no expected outputs for it."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    import cv2
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/example/dtu006"
    d = json.load(open(os.path.join(src, "input.json")))
    views = {v["value"]["ptr_wrapper"]["data"]["id_pose"]: v["value"]["ptr_wrapper"]["data"]["filename"] for v in d["views"]}
    names = [views[ex["key"]] for ex in d["extrinsics"]]            # view index = position in `extrinsics`
    masks = []
    for n in names:
        im = cv2.imread(os.path.join(src, "edges", n), cv2.IMREAD_COLOR)
        masks.append((im == 255).all(axis=2))
    masks = np.stack(masks)
    out = os.path.join(ROOT, "tests", "golden", "dtu006_edges.npz")
    np.savez_compressed(out, packed=np.packbits(masks, axis=2), shape=np.array(masks.shape, np.int64), names=np.array(names))
    print(out, masks.shape, "edge pixels per view", masks.sum(axis=(1, 2)).tolist(), "bytes", os.path.getsize(out))
