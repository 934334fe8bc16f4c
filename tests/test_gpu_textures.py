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
    from vk_voxel_cone_tracing_b200 import synth
    cam = synth.make_camera((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), aspect=1.0)
    gi.render_gbuffer(cam, 32, 32)          # textured G-buffer with all textures present
    gi.set_textures(tex[:2])
    with pytest.raises(VgiError):           # ... and not with some missing
        gi.render_gbuffer(cam, 32, 32)
    # a normal-mapped material without tangents: the build stages do not care, the G-buffer producer refuses
    nm = synth.textured_cornell(gbuffer_maps=True)
    nm.tangents = None
    gi.set_textures(synth.procedural_textures())
    gi.set_scene(nm)
    gi.build_clipmap(0)
    with pytest.raises(VgiError):
        gi.render_gbuffer(cam, 32, 32)


def _same(dev, host):
    for k in ("depth", "diffuse", "specular", "normal", "emission"):
        a = dev[k].cpu().numpy()
        b = host[k]
        if b.dtype == np.uint16:
            a = a.view(np.uint16)
        assert a.shape == b.shape and np.array_equal(a, b), (k, int((a != b).sum()))


@pytest.mark.gpu
def test_textured_gbuffer_bit_exact_with_the_host_producer():
    """vgi_render_gbuffer with textured materials (base colour, metallic-roughness, emissive, normal map + tangents, alpha
    cutoff in the visibility pass) against csrc/synth_raster.c, which tests/test_ref_shaders.py::test_live_textured_gbuffer
    pins on the reference's gBufferPass.frag: every attachment bit for bit, from two cameras (one next to the floor, so
    alpha-tested triangles cross the camera plane), and again after vgi_update_nodes rotated the scene (tangents follow)."""
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    scene = synth.textured_cornell(gbuffer_maps=True)
    tex = synth.procedural_textures()
    gi = VoxelGI(S.default_config(32, 2))
    gi.set_textures(tex)
    gi.set_scene(scene)
    w, h = 320, 200
    for eye, look in (((0.0, -1.0, 3.5), (0.0, 0.45, -1.0)), ((-2.0, 1.0, 3.0), (0.9, -0.6, -1.0)), ((0.5, -3.7, 1.0), (0.3, -0.1, -1.0))):
        cam = synth.make_camera(eye, look, aspect=w / h)
        host = raster.gbuffer(scene, cam, w, h, textures=tex)
        assert (host["depth"] < 1.0).mean() > 0.5
        _same(gi.render_gbuffer(cam, w, h), host)
    # the alpha cutoff really removes floor fragments
    opaque = synth.textured_cornell(gbuffer_maps=True)
    opaque.materials["alpha_mode"] = 0
    assert (raster.gbuffer(opaque, cam, w, h, textures=tex)["depth"] != host["depth"]).sum() > 500
    # a factor-only material below its cutoff vanishes as a whole (gBufferPass.frag:96 with no base-colour texture)
    gone = synth.textured_cornell(gbuffer_maps=True)
    back = 1
    gone.materials[back]["alpha_mode"], gone.materials[back]["alpha_cutoff"] = 1, 0.9
    gone.materials[back]["base_color_factor"][3] = 0.5
    gi2 = VoxelGI(S.default_config(32, 2))
    gi2.set_textures(tex)
    gi2.set_scene(gone)
    hg = raster.gbuffer(gone, cam, w, h, textures=tex)
    assert (hg["depth"] != host["depth"]).sum() > 500
    _same(gi2.render_gbuffer(cam, w, h), hg)
    gi2.close()
    # animated node: rotate about y and shift; normals and tangents go through itModel on the device
    a = np.deg2rad(25.0)
    m = np.array([[np.cos(a), 0, np.sin(a), 0.3], [0, 1, 0, 0.1], [-np.sin(a), 0, np.cos(a), -0.2], [0, 0, 0, 1]], np.float64)
    moved = synth.textured_cornell(gbuffer_maps=True)
    moved.nodes[0]["model"] = m.T.astype(np.float32).reshape(-1)
    moved.nodes[0]["it_model"] = np.linalg.inv(m).astype(np.float32).reshape(-1)      # (M^-1)^T column-major = M^-1 row-major
    gi.update_nodes(moved.nodes)
    cam = synth.make_camera((-2.0, 1.0, 3.0), (0.9, -0.6, -1.0), aspect=w / h)
    _same(gi.render_gbuffer(cam, w, h), raster.gbuffer(moved, cam, w, h, textures=tex))


@pytest.mark.gpu
def test_frame_view_host_on_a_textured_scene():
    """The headless view call (device-rendered shadow map + G-buffer, build, trace) on textured materials equals the separate
    calls on host-rendered inputs."""
    import torch
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    scene = synth.textured_cornell(gbuffer_maps=True)
    tex = synth.procedural_textures()
    cfg = S.default_config(64, 3)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = raster.shadow_depth(scene, shadow, 512)
    w, h = 160, 120
    eye = (-2.0, 1.0, 3.0)
    cam = synth.make_camera(eye, (0.9, -0.6, -1.0), aspect=w / h)
    a, b = VoxelGI(cfg), VoxelGI(cfg)
    for gi in (a, b):
        gi.set_textures(tex)
        gi.set_scene(scene)
    a.set_light(light, shadow, depth)
    a.update_regions(eye)
    a.build_clipmap(0)
    prm = a.default_vct_params(8)
    hostgb = raster.gbuffer(scene, cam, w, h, textures=tex)
    want = a.cone_trace(cam, a.upload_gbuffer(hostgb), prm)
    b.set_light(light, shadow, torch.ones((512, 512), dtype=torch.float32, device="cuda"))
    out = (torch.empty((h, w, 4), dtype=torch.float32).pin_memory(), torch.empty((h, w, 4), dtype=torch.float32).pin_memory())
    b.frame_view_host(0, eye, cam, w, h, shadow, prm, out[0], out[1])
    cov = torch.from_numpy(hostgb["depth"] < 1.0)
    assert torch.equal(out[0][cov], want[0].cpu()[cov]) and torch.equal(out[1][cov], want[1].cpu()[cov])
    assert float(out[0][cov][:, :3].max()) > 0.05
