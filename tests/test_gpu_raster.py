"""CUDA producers of the path's image inputs (vgi_render_shadow_map / vgi_render_gbuffer)
against the host-side software rasteriser that pins the rule (csrc/synth_raster.c: pixel-centre sampling, depth test
LESS with ties to the lower triangle index, binary64 edge functions). Bit-exact on every image."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _ctx(inp):
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    return gi


def _same_gbuffer(dev, host):
    for k in ("depth", "diffuse", "specular", "normal", "emission"):
        a = dev[k].cpu().numpy()
        b = host[k]
        if b.dtype == np.uint16:
            a = a.view(np.uint16)
        assert a.shape == b.shape, k
        assert np.array_equal(a, b), (k, int((a != b).sum()))


def test_cornell_shadow_map_and_gbuffer_bit_exact():
    inp = common.cornell_inputs(64, 1024, 128, 128)
    gi = _ctx(inp)
    depth = gi.render_shadow_map(inp["shadow"], 1024).cpu().numpy()
    assert np.array_equal(depth, inp["shadow_depth"])
    assert (depth < 1.0).any() and (depth == 1.0).any()
    _same_gbuffer(gi.render_gbuffer(inp["cam"], 128, 128), inp["gbuffer"])


def test_atrium_1080p_bit_exact_with_large_and_clipped_triangles():
    """Sponza-scale mesh: floor / wall triangles cover thousands of pixels (block-per-triangle queue), triangles
    behind the camera are skipped, fragments beyond the far plane are clipped."""
    inp = common.atrium_inputs(64, 1024, 480, 270, 2)
    gi = _ctx(inp)
    assert np.array_equal(gi.render_shadow_map(inp["shadow"], 1024).cpu().numpy(), inp["shadow_depth"])
    _same_gbuffer(gi.render_gbuffer(inp["cam"], 480, 270), inp["gbuffer"])
    # ragged size, camera elsewhere: compare against a fresh host render
    from vk_voxel_cone_tracing_b200 import raster, synth
    cam = synth.make_camera((3.0, 6.0, -2.0), (-0.6, -0.5, 0.62), aspect=301 / 173)
    _same_gbuffer(gi.render_gbuffer(cam, 301, 173), raster.gbuffer(inp["scene"], cam, 301, 173))


def test_triangles_crossing_the_camera_plane_bit_exact():
    """A hall of two triangles per wall seen from inside: every wall has a vertex behind the camera plane and takes the
    homogeneous edge-function path; device and host rasterisers agree bit for bit, and nothing is dropped."""
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    scene = synth.coarse_room(1)
    gi = VoxelGI(S.default_config(32, 2))
    gi.set_scene(scene)
    for pos, dirv, (w, h) in (((2.0, 1.7, -3.0), (1.0, -0.15, 0.4), (320, 180)), ((-9.0, 4.5, 8.0), (0.3, -0.6, -1.0), (301, 173))):
        d = np.array(dirv) / np.linalg.norm(dirv)
        cam = synth.make_camera(pos, tuple(d), aspect=w / h)
        host = raster.gbuffer(scene, cam, w, h)
        assert (host["depth"] < 1.0).all()
        _same_gbuffer(gi.render_gbuffer(cam, w, h), host)


def test_rendered_inputs_drive_the_path():
    """The device-rendered shadow map and G-buffer feed vgi_set_light / vgi_cone_trace directly (no host copy) and
    give the image the host-rendered inputs give."""
    import torch
    inp = common.cornell_inputs(64, 1024, 128, 128)
    a, b = _ctx(inp), _ctx(inp)
    a.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    b.set_light(inp["light"], inp["shadow"], b.render_shadow_map(inp["shadow"], 1024))
    outs = []
    for gi, gb in ((a, a.upload_gbuffer(inp["gbuffer"])), (b, b.render_gbuffer(inp["cam"], 128, 128))):
        gi.update_regions(inp["cam_pos"])
        gi.build_clipmap(0)
        outs.append(gi.cone_trace(inp["cam"], gb, gi.default_vct_params(8)))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_raster_error_paths():
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VgiError, VoxelGI
    gi = VoxelGI(S.default_config(32, 2))
    inp = common.cornell_inputs(64, 1024, 128, 128)
    with pytest.raises(VgiError):
        gi.render_shadow_map(inp["shadow"], 64)              # no scene yet
    with pytest.raises(VgiError):
        gi.render_gbuffer(inp["cam"], 16, 16)


def test_frame_view_host_equals_the_device_resident_path():
    """vgi_frame_view_host (the batched / headless view: camera up, shadow map + G-buffer rasterised on the device, build,
    trace, both images down) gives exactly what the separate calls give on host-rendered inputs — the e2e call of bench.py."""
    import torch
    inp = common.cornell_inputs(64, 1024, 128, 128)
    a, b = _ctx(inp), _ctx(inp)
    a.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    a.update_regions(inp["cam_pos"])
    a.build_clipmap(0)
    prm = a.default_vct_params(8)
    want = a.cone_trace(inp["cam"], a.upload_gbuffer(inp["gbuffer"]), prm)
    # b never sees a host image: the light's depth buffer is a placeholder that the call replaces with its own render
    b.set_light(inp["light"], inp["shadow"], torch.ones((1024, 1024), dtype=torch.float32, device="cuda"))
    out = (torch.empty((128, 128, 4), dtype=torch.float32).pin_memory(), torch.empty((128, 128, 4), dtype=torch.float32).pin_memory())
    for _ in range(2):      # twice: the second call reuses the staging buffers
        b.frame_view_host(0, inp["cam_pos"], inp["cam"], 128, 128, inp["shadow"], prm, out[0], out[1])
    cov = torch.from_numpy(inp["gbuffer"]["depth"] < 1.0)
    assert torch.equal(out[0][cov], want[0].cpu()[cov]) and torch.equal(out[1][cov], want[1].cpu()[cov])
    assert float(out[0][cov][:, :3].max()) > 0.05
    # shadow = None keeps the current shadow map
    b.frame_view_host(0, inp["cam_pos"], inp["cam"], 128, 128, None, prm, out[0], out[1])
    assert torch.equal(out[0][cov], want[0].cpu()[cov])


def test_pipelined_view_frames_equal_the_synchronous_call():
    """vgi_frame_view_host_begin / _end with two frames in flight over a six-camera sequence: every frame's images equal what
    the synchronous call delivers for the same camera, and the call order is enforced (a third begin, an end with nothing
    in flight, a synchronous call in between fail with VGI_E_STATE)."""
    import torch
    from vk_voxel_cone_tracing_b200 import synth
    from vk_voxel_cone_tracing_b200.api import VgiError
    inp = common.cornell_inputs(64, 1024, 128, 128)
    a, b = _ctx(inp), _ctx(inp)
    for gi in (a, b):
        gi.set_light(inp["light"], inp["shadow"], torch.ones((1024, 1024), dtype=torch.float32, device="cuda"))
    prm = a.default_vct_params(8)
    cams = []
    for i in range(6):
        eye = (0.4 * i - 1.0, 0.2 * i - 0.5, 0.3 * i)
        cams.append((eye, synth.make_camera(eye, (0.1 * i, -0.05 * i, -1.0), aspect=1.0)))
    pin = lambda: (torch.empty((128, 128, 4), dtype=torch.float32).pin_memory(), torch.empty((128, 128, 4), dtype=torch.float32).pin_memory())  # noqa: E731
    want = []
    for i, (eye, cam) in enumerate(cams):
        o = pin()
        a.frame_view_host(i, eye, cam, 128, 128, inp["shadow"], prm, o[0], o[1])
        want.append((o[0].clone(), o[1].clone()))
    outs = [pin() for _ in cams]
    with pytest.raises(VgiError):
        b.frame_view_host_end()
    b.frame_view_host_begin(0, cams[0][0], cams[0][1], 128, 128, inp["shadow"], prm, outs[0][0], outs[0][1])
    for i in range(1, len(cams)):
        b.frame_view_host_begin(i, cams[i][0], cams[i][1], 128, 128, inp["shadow"], prm, outs[i][0], outs[i][1])
        if i == 1:
            with pytest.raises(VgiError):       # two in flight already
                b.frame_view_host_begin(9, cams[0][0], cams[0][1], 128, 128, inp["shadow"], prm, outs[0][0], outs[0][1])
            with pytest.raises(VgiError):
                b.frame_view_host(9, cams[0][0], cams[0][1], 128, 128, inp["shadow"], prm, outs[0][0], outs[0][1])
        b.frame_view_host_end()                 # frame i - 1 is home
        assert torch.equal(outs[i - 1][0], want[i - 1][0]) and torch.equal(outs[i - 1][1], want[i - 1][1]), i - 1
    b.frame_view_host_end()
    assert torch.equal(outs[-1][0], want[-1][0]) and torch.equal(outs[-1][1], want[-1][1])
    assert float(want[0][0][..., :3].max()) > 0.05
