"""GPU parity for textured materials (-m gpu): libvgi against the oracle on the textured Cornell fixture — opacity and
radiance atlases bit for bit (the occlusion-texture alpha test changes the OCCUPANCY, base-colour and emissive textures the
radiance), the SVO fragment list as a multiset, the traced image within 1e-3; and the error paths of vgi_set_textures."""
import numpy as np
import pytest

from tests import common  # noqa: F401  (path set-up)


def _inputs(res=64, levels=3):
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    scene = synth.textured_cornell()
    cfg = S.default_config(res, levels)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = raster.shadow_depth(scene, shadow, 512)
    return scene, synth.procedural_textures(), cfg, light, shadow, depth


@pytest.mark.gpu
def test_textured_clipmap_bit_exact_and_trace():
    import torch
    from oracle import pyoracle as O
    from vk_voxel_cone_tracing_b200 import raster, synth
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    O.build()
    scene, tex, cfg, light, shadow, depth = _inputs()
    cam_pos = (0.4, -0.3, 0.2)
    gi = VoxelGI(cfg)
    gi.set_textures(tex)
    gi.set_scene(scene)
    gi.set_light(light, shadow, depth)
    gi.update_regions(cam_pos)
    gi.build_clipmap(0)
    regs = O.regions(cfg, cam_pos)
    O.set_textures(tex)
    try:
        osc = O.OracleScene(scene)
        op, rad, pairs = O.build_clipmap(cfg, regs, osc, light, shadow, depth, 0)
        assert gi.stats().clip_pairs == pairs
        assert torch.equal(gi.export_atlas(0).cpu(), torch.from_numpy(op)), "opacity atlas differs (alpha test)"
        assert torch.equal(gi.export_atlas(1).cpu(), torch.from_numpy(rad)), "radiance atlas differs (textures)"
        # a second frame with the cadence: levels 1.. keep their radiance, level 0 is re-injected
        gi.build_clipmap(1)
        O.build_clipmap(cfg, regs, osc, light, shadow, depth, 1, op, rad)
        assert torch.equal(gi.export_atlas(1).cpu(), torch.from_numpy(rad))
        # the trace only sees the atlases; its G-buffer comes from the host (factor-only raster of the same geometry)
        cam = synth.make_camera((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), aspect=1.0)
        plain = synth.textured_cornell()
        for k in ("base_color_texture", "emissive_texture", "occlusion_texture"):
            plain.materials[k] = -1
        gb = raster.gbuffer(plain, cam, 96, 96)
        prm = gi.default_vct_params(8)
        hg = O.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
        gi.build_clipmap(0)
        O.build_clipmap(cfg, regs, osc, light, shadow, depth, 0, op, rad)
        ref_d, ref_s, _ = O.cone_trace(cfg, cam, hg, prm, light, shadow, depth, rad)
        d, s = gi.cone_trace(cam, gi.upload_gbuffer(gb), prm)
        cov = gb["depth"] < 1.0
        assert np.abs(d.cpu().numpy() - ref_d)[cov].max() <= 1e-3
        assert np.abs(s.cpu().numpy() - ref_s)[cov].max() <= 1e-3
    finally:
        O.set_textures([])


@pytest.mark.gpu
def test_textured_svo_fragments_match():
    from oracle import pyoracle as O
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    O.build()
    scene, tex, cfg, light, shadow, depth = _inputs(32, 2)
    lo, hi = scene.world_bbox()
    gi = VoxelGI(cfg)
    gi.set_textures(tex)
    gi.set_scene(scene)
    gi.set_light(light, shadow, depth)
    O.set_textures(tex)
    try:
        for level in (5, 6):
            gi.svo_voxelize(level, lo, hi)
            got = gi.svo_fragments().cpu().numpy().view(np.uint32)
            want = O.svo_fragments(level, lo, hi, O.OracleScene(scene), light, shadow, depth)
            a = np.sort(got[:, 0].astype(np.uint64) | (got[:, 1].astype(np.uint64) << 32))
            b = np.sort(want[:, 0].astype(np.uint64) | (want[:, 1].astype(np.uint64) << 32))
            assert a.shape == b.shape and np.array_equal(a, b), f"level {level}"
    finally:
        O.set_textures([])


@pytest.mark.gpu
def test_missing_textures_fail_loudly():
    from vk_voxel_cone_tracing_b200.api import VgiError, VoxelGI
    scene, tex, cfg, light, shadow, depth = _inputs(32, 2)
    gi = VoxelGI(cfg)
    gi.set_scene(scene)                     # materials reference textures 0..3, none provided yet
    gi.set_light(light, shadow, depth)
    gi.update_regions((0.0, 0.0, 0.0))
    with pytest.raises(VgiError):
        gi.build_clipmap(0)
    gi.set_textures(tex[:2])                # still too few
    with pytest.raises(VgiError):
        gi.build_clipmap(0)
    gi.set_textures(tex)
    gi.build_clipmap(0)
    assert gi.stats().clip_pairs > 0
    with pytest.raises(VgiError):           # the G-buffer producer (adjacent pass) stays factor-only
        from vk_voxel_cone_tracing_b200 import synth
        gi.render_gbuffer(synth.make_camera((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), aspect=1.0), 32, 32)
