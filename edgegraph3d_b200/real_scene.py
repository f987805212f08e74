"""Scene assembly from real inputs: an OpenMVG sfm_data JSON (or the committed dtu006 fixtures) + edge images ->
FlatScene, i.e. what edge_matching() does before the pipelines start (src/edgegraph3d/edge_matcher.cpp:66-115):
read_sfm_data, parse the edge images, convert_edge_images_to_optimized_polyline_graphs, fundamental matrices."""
import os
import numpy as np
from . import plg_build as PB
from .scene import FlatScene


def unpack_edge_masks(npz_path):
    z = np.load(npz_path)
    shape = tuple(int(x) for x in z["shape"])
    return np.unpackbits(z["packed"], axis=2)[:, :, :shape[2]].astype(bool)


def scene_from_parts(sfm, edge_images, edge_color=PB.EDGE_COLOR, fundamental=None, fundamental_valid=None):
    """sfm: dict as openmvg_io.load_sfm_data returns (cameras, width, height, track_*); edge_images: one per view."""
    plgs = PB.polyline_graphs_from_edge_images(edge_images, edge_color)
    if fundamental is None:
        from . import openmvg_io as io
        fundamental, fundamental_valid = io.fundamental_from_tracks(len(edge_images), sfm["track_off"], sfm["track_view"], sfm["track_xy"])
    sc = FlatScene(width=int(sfm["width"]), height=int(sfm["height"]), cameras=sfm["cameras"], fundamental=fundamental,
                   fundamental_valid=fundamental_valid, track_xyz=sfm["track_xyz"], track_off=sfm["track_off"],
                   track_view=sfm["track_view"], track_xy=sfm["track_xy"], **PB.scene_polyline_arrays(plgs))
    return sc, plgs


def dtu006_scene(golden_dir):
    """The reference's packaged example (example/dtu006) from the committed fixtures: real cameras / tracks / LMedS F
    (dtu006_sfm.npz) and real edge maps (dtu006_edges.npz)."""
    z = np.load(os.path.join(golden_dir, "dtu006_sfm.npz"))
    masks = unpack_edge_masks(os.path.join(golden_dir, "dtu006_edges.npz"))
    imgs = [m.astype(np.uint8) * 255 for m in masks]
    return scene_from_parts({k: z[k] for k in ("cameras", "width", "height", "track_xyz", "track_off", "track_view", "track_xy")}, imgs,
                            edge_color=255, fundamental=z["fundamental"], fundamental_valid=z["fundamental_valid"])
