"""Near-plane handling of the software rasterisers: a triangle with a vertex at or behind the camera plane
must be rasterised where it is in front of it, as the reference's hardware raster does after clipping — not dropped.
A hall built from TWO triangles per wall is rendered from inside (every wall crosses the camera plane) and compared with the
same hall tessellated finely, where every visible triangle lies wholly in front of the camera and takes the ordinary path."""
import numpy as np

from vk_voxel_cone_tracing_b200 import raster, synth


def _half(a):
    return a.view(np.float16).astype(np.float32)


def test_walls_that_cross_the_camera_plane_are_drawn():
    coarse, fine = synth.coarse_room(1), synth.coarse_room(96)
    for pos, dirv in (((2.0, 1.7, -3.0), (1.0, -0.15, 0.4)), ((-9.0, 4.5, 8.0), (0.3, -0.6, -1.0)), ((0.0, 3.0, 0.0), (0.0, 0.0, 1.0))):
        d = np.array(dirv) / np.linalg.norm(dirv)
        cam = synth.make_camera(pos, tuple(d), aspect=160 / 90)
        a = raster.gbuffer(coarse, cam, 160, 90)
        b = raster.gbuffer(fine, cam, 160, 90)
        # inside a closed hall every pixel is covered; before the fix the coarse hall lost every wall next to the viewer
        assert (a["depth"] < 1.0).all() and (b["depth"] < 1.0).all()
        assert np.abs(a["depth"] - b["depth"]).max() < 2e-6
        # same wall (material colour) and same normal at every pixel, up to the seams between walls
        same = (a["diffuse"] == b["diffuse"]).all(axis=2)
        assert same.mean() > 0.985
        assert np.abs(_half(a["normal"]) - _half(b["normal"]))[same].max() <= 1e-3


def test_shadow_map_of_the_hall_is_unchanged_by_tessellation():
    # orthographic light: w == 1 everywhere, the ordinary path; guards the refactoring of the projection
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    a = raster.shadow_depth(synth.coarse_room(1), shadow, 128)
    b = raster.shadow_depth(synth.coarse_room(32), shadow, 128)
    assert np.abs(a - b).max() < 1e-5 and (a < 1.0).any()
