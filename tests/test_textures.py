"""Textured materials (ref: msaaVoxelizer.frag:64, msaaInjectRadiance.frag:73-82,131-136, voxelizer.frag:52-76,
samplers REPEAT + LINEAR with maxLod = 0: GLTFScene.cpp:339, quirk Q23). CPU part: the oracle's texture rule and what the
textures change; the GPU parity lives in tests/test_gpu_textures.py, the pin against the reference's shader text in
tests/test_ref_shaders.py::test_live_textured_injection."""
import numpy as np
import pytest

from oracle import pyoracle as O
from vk_voxel_cone_tracing_b200 import raster, structs as S, synth


@pytest.fixture(scope="module")
def oracle():
    O.build()
    yield O
    O.set_textures([])


def test_bilinear_repeat_known_answers(oracle):
    tex = synth.procedural_textures()
    oracle.set_textures(tex)
    n = tex[3]                                  # 5 rows x 3 columns
    H, W = n.shape[:2]
    # texel centres return the texel; coordinates one period apart return the same value (REPEAT)
    for (x, y) in ((0, 0), (2, 4), (1, 3)):
        u, v = (x + 0.5) / W, (y + 0.5) / H
        want = n[y, x].astype(np.float32) / np.float32(255.0)
        assert np.array_equal(oracle.texture_fetch(3, u, v), want)
        assert np.allclose(oracle.texture_fetch(3, u + 1.0, v - 2.0), want, atol=2e-6)
    # half-way between texel (2, y) and the wrapped texel (0, y): their mean
    got = oracle.texture_fetch(3, 1.0, 0.5 / H)
    want = (n[0, 2].astype(np.float32) / np.float32(255.0) + n[0, 0].astype(np.float32) / np.float32(255.0)) * np.float32(0.5)
    assert np.allclose(got, want, atol=1e-6)


def test_textures_change_occupancy_and_radiance(oracle):
    scene = synth.textured_cornell()
    tex = synth.procedural_textures()
    cfg = S.default_config(32, 2)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = raster.shadow_depth(scene, shadow, 256)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    osc = oracle.OracleScene(scene)
    assert osc.uv is not None and osc.uv.shape == (scene.triangle_count, 3, 2)
    oracle.set_textures(tex)
    op_t, rad_t, pairs_t = oracle.build_clipmap(cfg, regs, osc, light, shadow, depth, 0)
    # the same geometry with the textures switched off (indices -1): factor-only shading
    plain = synth.textured_cornell()
    for k in ("base_color_texture", "emissive_texture", "occlusion_texture"):
        plain.materials[k] = -1
    oracle.set_textures([])
    op_p, rad_p, pairs_p = oracle.build_clipmap(cfg, regs, oracle.OracleScene(plain), light, shadow, depth, 0)
    # alpha-tested materials: holes -> fewer pairs, occupancy is a strict subset
    assert pairs_t < pairs_p
    occ_t, occ_p = op_t[..., 0] > 0, op_p[..., 0] > 0
    assert not (occ_t & ~occ_p).any() and (occ_p & ~occ_t).any()
    # base colour / emissive textures change the injected radiance where both are occupied
    both = occ_t & occ_p
    assert (rad_t[both] != rad_p[both]).any()
