"""Shared fixtures for the parity tests: seeded synthetic inputs."""
import functools

import numpy as np

from vk_voxel_cone_tracing_b200 import raster, structs as S, synth


@functools.lru_cache(maxsize=None)
def cornell_inputs(resolution=64, shadow_size=1024, width=128, height=128):
    """Config 1: Cornell box, camera at the origin looking -z, light moved to (0,20,-3.5)."""
    scene = synth.cornell_box()
    cfg = S.default_config(resolution, 6)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = raster.shadow_depth(scene, shadow, shadow_size)
    cam = synth.make_camera((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), aspect=width / height)
    gb = raster.gbuffer(scene, cam, width, height)
    return dict(scene=scene, cfg=cfg, light=light, shadow=shadow, shadow_depth=depth, cam=cam, gbuffer=gb,
                cam_pos=(0.0, 0.0, 0.0))


@functools.lru_cache(maxsize=None)
def atrium_inputs(resolution=256, shadow_size=4096, width=1920, height=1080, levels=6):
    """Config 2: Sponza-scale atrium, camera (-8,3,0) looking +x, reference default light."""
    scene = synth.atrium()
    cfg = S.default_config(resolution, levels)
    light, shadow = synth.make_light()
    depth = raster.shadow_depth(scene, shadow, shadow_size)
    cam = synth.make_camera((-8.0, 3.0, 0.0), (1.0, 0.0, 0.0), aspect=width / height)
    gb = raster.gbuffer(scene, cam, width, height)
    return dict(scene=scene, cfg=cfg, light=light, shadow=shadow, shadow_depth=depth, cam=cam, gbuffer=gb,
                cam_pos=(-8.0, 3.0, 0.0))


def psnr(a, b, peak=1.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


# ---- inputs shared by the GPU parity tests and their CPU twins (tests/test_ref_shaders.py checks on the CPU that the
# oracle and the reference's shaders agree on exactly these inputs; the GPU tests then compare libvgi with both) ----
HELPER_CAM = (1.3, -0.7, 2.9)
HELPER_CLEAR_CASES = (((0, 0, 0), (32, 32, 32), 1), ((5, 30, 17), (9, 4, 20), 2))     # non-negative corners (the reference's only use)


def helper_config(mode_flags=0, levels=3):
    return S.default_config(32, levels, downsample_band=3, mode_flags=mode_flags)


def random_atlas(cfg, seed):
    rng = np.random.RandomState(seed)
    a = rng.randint(0, 256, size=S.atlas_shape(cfg)).astype(np.uint8)
    a[rng.rand(*a.shape[:3]) < 0.5] = 0       # realistic sparsity
    return a


def filter_images(h, w, seed):
    rng = np.random.RandomState(seed)
    dif = rng.rand(h, w, 4).astype(np.float32)
    spc = (rng.rand(h, w, 4) * 3.0).astype(np.float32)    # indirect_specular_intensity = 3
    spc[rng.rand(h, w) < 0.6] = (0.0, 0.0, 0.0, 1.0)      # most pixels have no specular cone
    dif[..., 3] = 1.0
    return dif, spc
