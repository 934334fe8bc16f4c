"""GPU parity: libvgi.so (through the C ABI) vs the CPU oracle on identical seeded inputs.
Bars: occupancy / atlases bit-exact; cone-traced images max abs <= 1e-3, PSNR >= 50 dB."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _build_both(oracle, inp, frame=0):
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    cfg = inp["cfg"]
    gi = VoxelGI(cfg)
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(frame)
    regs = oracle.regions(cfg, inp["cam_pos"])
    got = gi.regions()
    for a, b in zip(regs, got):
        assert list(a.min_corner) == list(b.min_corner) and a.voxel_size == b.voxel_size
    osc = oracle.OracleScene(inp["scene"])
    op, rad, pairs = oracle.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], frame)
    return gi, regs, op, rad, pairs


def test_cornell64_atlases_bit_exact(oracle):
    inp = common.cornell_inputs()
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    st = gi.stats()
    assert st.clip_pairs == pairs
    g_op = gi.export_atlas(0).cpu().numpy()
    g_rad = gi.export_atlas(1).cpu().numpy()
    assert g_op.shape == op.shape
    assert (op[..., 3] > 0).sum() > 10000
    assert np.array_equal(g_op, op), f"opacity atlas differs in {(g_op != op).sum()} bytes"
    assert np.array_equal(g_rad, rad), f"radiance atlas differs in {(g_rad != rad).sum()} bytes"


@pytest.mark.parametrize("mode", [7, 8])
def test_cornell64_cone_trace(oracle, mode):
    from vk_voxel_cone_tracing_b200 import structs as S
    inp = common.cornell_inputs()
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    prm = gi.default_vct_params(mode)
    ref_prm = S.default_vct_params(regs[0], inp["cfg"].resolution, mode)
    assert bytes(memoryview(prm)) == bytes(memoryview(ref_prm))
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s, taps = oracle.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"],
                                           inp["shadow_depth"], rad)
    dgb = gi.upload_gbuffer(gb)
    d, s = gi.cone_trace(inp["cam"], dgb, prm)
    d, s = d.cpu().numpy(), s.cpu().numpy()
    covered = gb["depth"] < 1.0
    assert covered.mean() > 0.5
    err_d = np.abs(d - ref_d)[covered].max()
    err_s = np.abs(s - ref_s)[covered].max()
    assert err_d <= 1e-3, err_d
    assert err_s <= 1e-3, err_s
    assert common.psnr(d[..., :3], ref_d[..., :3]) >= 50.0
    assert common.psnr(s[..., :3], ref_s[..., :3]) >= 50.0


def _gi_for(inp):
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    return gi


def test_cadence_and_moving_camera_bit_exact(oracle):
    """Frames 0..5 with a moving camera: the sparse build only rewrites records its masks know about,
    so stale records (regions that moved away, off-cadence radiance) must still match the oracle's
    dense clear / keep semantics (RadianceInjectionPass.cpp:36-38,75,110,124; VoxelizationPass.cpp:104-126)."""
    inp = common.cornell_inputs()
    cfg = inp["cfg"]
    gi = _gi_for(inp)
    osc = oracle.OracleScene(inp["scene"])
    op, rad = oracle.new_atlas(cfg), oracle.new_atlas(cfg)
    for frame in range(6):
        cam = (0.37 * frame, -0.11 * frame, 0.29 * frame)
        gi.update_regions(cam)
        gi.build_clipmap(frame)
        regs = oracle.regions(cfg, cam)
        oracle.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], frame, op, rad)
        g_op = gi.export_atlas(0).cpu().numpy()
        g_rad = gi.export_atlas(1).cpu().numpy()
        assert np.array_equal(g_op, op), f"frame {frame}: opacity differs in {(g_op != op).sum()} bytes"
        assert np.array_equal(g_rad, rad), f"frame {frame}: radiance differs in {(g_rad != rad).sum()} bytes"


def test_incremental_build_equals_full_rebuild_over_a_camera_path():
    """vgi_build_clipmap_incremental on every frame of a 24-frame camera walk (regions of different levels move on
    different frames, cadence frames come and go) against vgi_build_clipmap on every frame of the same walk: opacity and
    radiance atlases bit for bit on every frame, the traced images at the end, and the call really skips work — frames with
    nothing to rebuild, frames that rebuild only the coarse levels' tail, frames that rebuild everything. The full build is
    compared with the oracle on such a walk by test_cadence_and_moving_camera_bit_exact."""
    import torch
    inp = common.cornell_inputs()
    cfg = inp["cfg"]
    full, inc = _gi_for(inp), _gi_for(inp)
    firsts = []
    for frame in range(24):
        cam = (0.11 * frame, -0.03 * frame, 0.07 * frame)
        for gi in (full, inc):
            gi.update_regions(cam)
        full.build_clipmap(frame)
        firsts.append(inc.build_clipmap_incremental(frame))
        for which in (0, 1):
            a, b = full.export_atlas(which), inc.export_atlas(which)
            assert torch.equal(a, b), f"frame {frame}: atlas {which} differs in {int((a != b).sum())} bytes (first level rebuilt {firsts[-1]})"
    L = cfg.level_count
    assert firsts[0] == 0
    assert firsts.count(L) >= 4, firsts                 # nothing moved, no stale radiance due
    assert any(0 < f < L for f in firsts), firsts       # only coarser levels
    assert firsts.count(0) >= 3, firsts                 # the finest region moved
    prm = full.default_vct_params(8)
    cam = inp["cam"]
    gb = full.upload_gbuffer(inp["gbuffer"])
    d0, s0 = full.cone_trace(cam, gb, prm)
    d1, s1 = inc.cone_trace(cam, inc.upload_gbuffer(inp["gbuffer"]), prm)
    assert torch.equal(d0, d1) and torch.equal(s0, s1)
    # a setter invalidates: the next call rebuilds everything
    inc.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    assert inc.build_clipmap_incremental(24) == 0
    # ... and so does a full build in between
    assert inc.build_clipmap_incremental(25) == L        # odd frame, same regions: only level 0 is on cadence and it is current
    inc.build_clipmap(26)
    assert inc.build_clipmap_incremental(27) == 0


def test_scene_change_clears_stale_records(oracle):
    """Replace the scene by a much smaller one: every record of the old scene must be cleared."""
    from vk_voxel_cone_tracing_b200 import synth
    inp = common.cornell_inputs()
    cfg = inp["cfg"]
    gi = _gi_for(inp)
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(0)
    quad = synth.quad_scene([(0.3, 0.3, 0.3), (0.3, 0.3, 1.7), (1.7, 0.3, 1.7), (1.7, 0.3, 0.3)], normal=(0, 1, 0))
    gi.set_scene(quad)
    gi.build_clipmap(0)
    regs = oracle.regions(cfg, inp["cam_pos"])
    op, rad, pairs = oracle.build_clipmap(cfg, regs, oracle.OracleScene(quad), inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    assert gi.stats().clip_pairs == pairs
    assert np.array_equal(gi.export_atlas(0).cpu().numpy(), op)
    assert np.array_equal(gi.export_atlas(1).cpu().numpy(), rad)


@pytest.mark.parametrize("res,levels", [(32, 3), (128, 2)])
def test_other_resolutions_bit_exact(oracle, res, levels):
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    scene = synth.cornell_box(wall_quads=16, box_quads=6)
    cfg = S.default_config(res, levels)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = raster.shadow_depth(scene, shadow, 512)
    inp = dict(scene=scene, cfg=cfg, light=light, shadow=shadow, shadow_depth=depth)
    gi = _gi_for(inp)
    cam = (0.6, -0.3, 1.1)
    gi.update_regions(cam)
    gi.build_clipmap(0)
    regs = oracle.regions(cfg, cam)
    op, rad, pairs = oracle.build_clipmap(cfg, regs, oracle.OracleScene(scene), light, shadow, depth, 0)
    assert gi.stats().clip_pairs == pairs
    assert np.array_equal(gi.export_atlas(0).cpu().numpy(), op)
    assert np.array_equal(gi.export_atlas(1).cpu().numpy(), rad)


def test_atrium64_off_diagonal_camera(oracle):
    """Config-2 scene and camera (-8,3,0) at 64^3 (the oracle finishes in seconds): regions are off the
    x=y=z diagonal (Q2), big triangles go through the large-triangle queue, the cone trace sees all levels."""
    inp = common.atrium_inputs(64, 1024, 240, 136, 6)
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    assert gi.stats().clip_pairs == pairs
    g_op = gi.export_atlas(0).cpu().numpy()
    g_rad = gi.export_atlas(1).cpu().numpy()
    assert np.array_equal(g_op, op), f"opacity atlas differs in {(g_op != op).sum()} bytes"
    assert np.array_equal(g_rad, rad), f"radiance atlas differs in {(g_rad != rad).sum()} bytes"
    prm = gi.default_vct_params(8)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s, taps = oracle.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"],
                                           inp["shadow_depth"], rad)
    d, s = gi.cone_trace(inp["cam"], gi.upload_gbuffer(gb), prm)
    d, s = d.cpu().numpy(), s.cpu().numpy()
    covered = gb["depth"] < 1.0
    assert np.abs(d - ref_d)[covered].max() <= 1e-3
    assert np.abs(s - ref_s)[covered].max() <= 1e-3
    assert common.psnr(d[..., :3], ref_d[..., :3]) >= 50.0 and common.psnr(s[..., :3], ref_s[..., :3]) >= 50.0


def test_frame_host_matches_device_path(oracle):
    """vgi_frame_host (host buffers in, host images out) == the device-resident call sequence."""
    import torch
    inp = common.cornell_inputs()
    gi = _gi_for(inp)
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(0)
    prm = gi.default_vct_params(8)
    d, s = gi.cone_trace(inp["cam"], gi.upload_gbuffer(inp["gbuffer"]), prm)
    gb = inp["gbuffer"]
    h, w = gb["depth"].shape
    od = np.full((h, w, 4), 7.0, dtype=np.float32)
    os_ = np.full((h, w, 4), 7.0, dtype=np.float32)
    gi.frame_host(0, inp["cam_pos"], inp["cam"], gb, inp["shadow_depth"], None, od, os_)
    covered = gb["depth"] < 1.0
    assert np.array_equal(od[covered], d.cpu().numpy()[covered])
    assert np.array_equal(os_[covered], s.cpu().numpy()[covered])
    assert (od[~covered] == 0).all()   # discarded pixels get a defined value on the host path


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 6])
def test_all_rendering_modes(oracle, mode):
    """The nine output modes of voxelConeTracing.frag:258-293 (7 and 8 are covered above)."""
    inp = common.cornell_inputs(64, 1024, 96, 96)
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    prm = gi.default_vct_params(mode)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s, _ = oracle.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
    d, s = gi.cone_trace(inp["cam"], gi.upload_gbuffer(gb), prm)
    covered = gb["depth"] < 1.0
    assert np.abs(d.cpu().numpy() - ref_d)[covered].max() <= 1e-3
    assert np.abs(s.cpu().numpy() - ref_s)[covered].max() <= 1e-3


def test_32_cones_and_min_step_factor(oracle):
    """enable32Cones (aperture 0.628319, voxelConeTracing.frag:79-114) and a smaller step factor (more steps
    than the tabulated default sequence)."""
    inp = common.cornell_inputs(64, 1024, 96, 96)
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    dgb = gi.upload_gbuffer(gb)
    covered = gb["depth"] < 1.0
    for cones32, step in ((1, 1.0), (0, 0.2), (1, 0.5)):
        prm = gi.default_vct_params(8)
        prm.enable_32_cones = cones32
        prm.min_trace_step_factor = step
        ref_d, ref_s, _ = oracle.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
        d, s = gi.cone_trace(inp["cam"], dgb, prm)
        assert np.abs(d.cpu().numpy() - ref_d)[covered].max() <= 1e-3, (cones32, step)
        assert np.abs(s.cpu().numpy() - ref_s)[covered].max() <= 1e-3, (cones32, step)
        assert common.psnr(d.cpu().numpy()[..., :3], ref_d[..., :3]) >= 50.0


def test_shadow_compare_mode_bit_exact(oracle):
    """VGI_MODE_SHADOW_COMPARE (the fixed Q1: depth comparison instead of averaged raw depth)."""
    from vk_voxel_cone_tracing_b200 import structs as S
    inp = dict(common.cornell_inputs(64, 1024, 96, 96))
    inp["cfg"] = S.default_config(64, 6, mode_flags=S.VGI_MODE_SHADOW_COMPARE)
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    assert np.array_equal(gi.export_atlas(1).cpu().numpy(), rad)
    prm = gi.default_vct_params(8)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s, _ = oracle.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
    d, s = gi.cone_trace(inp["cam"], gi.upload_gbuffer(gb), prm)
    covered = gb["depth"] < 1.0
    assert np.abs(d.cpu().numpy() - ref_d)[covered].max() <= 1e-3


def test_literal_border_export(oracle):
    """VGI_MODE_BORDER_LITERAL: only the low opacity border is wrapped, the radiance border stays zero (Q4/Q5)."""
    from vk_voxel_cone_tracing_b200 import structs as S
    inp = dict(common.cornell_inputs(64, 1024, 96, 96))
    inp["cfg"] = S.default_config(64, 6, mode_flags=S.VGI_MODE_BORDER_LITERAL)
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    assert np.array_equal(gi.export_atlas(0).cpu().numpy(), op)
    assert np.array_equal(gi.export_atlas(1).cpu().numpy(), rad)


def test_error_paths():
    """Call-order and argument errors come back as codes, never as crashes (vgi.h conventions)."""
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VgiError, VoxelGI
    gi = VoxelGI(S.default_config(32, 2))
    with pytest.raises(VgiError) as e:
        gi.voxelize_opacity()
    assert e.value.code == S.VGI_E_STATE
    with pytest.raises(VgiError) as e:
        gi.export_atlas(0)
    assert e.value.code == S.VGI_E_STATE
    with pytest.raises(VgiError) as e:
        gi.set_slab(5, 3)
    assert e.value.code == S.VGI_E_INVALID
    with pytest.raises(VgiError) as e:
        gi.svo_build()
    assert e.value.code == S.VGI_E_STATE
    with pytest.raises(VgiError) as e:
        VoxelGI(S.default_config(100, 2))
    assert e.value.code == S.VGI_E_INVALID


def test_peer_build_single_rank_equals_build_clipmap(oracle):
    """The peer build (vgi_peer_*: kernel-side exchange over mapped peer memory) with a world of one GPU: the same
    kernels, flag barriers and plane-ownership logic as on 8 GPUs (tools/multigpu_check.py checks 2 / 4 / 8), compared
    bit for bit with vgi_build_clipmap over frames with cadence and a moving camera."""
    import socket
    import torch
    import torch.distributed as dist
    from vk_voxel_cone_tracing_b200 import multigpu as M
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    inp = common.cornell_inputs(64, 1024, 128, 128)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        gis = []
        for _ in range(2):
            gi = VoxelGI(inp["cfg"])
            gi.set_scene(inp["scene"])
            gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
            gis.append(gi)
        ref, peer = gis
        pb = M.PeerBuild(peer)
        for frame in range(4):
            cam = tuple(np.array(inp["cam_pos"]) + np.array([0.41, -0.13, 0.27]) * frame)
            ref.update_regions(cam)
            peer.update_regions(cam)
            ref.build_clipmap(frame)
            pb.build(frame)
            for which in (0, 1):
                assert torch.equal(ref.export_atlas(which), peer.export_atlas(which)), (frame, which)
        assert peer.stats().occupied_voxels == ref.stats().occupied_voxels > 0
        pb.close()
        # after detaching, the ctx builds alone again
        peer.build_clipmap(4)
        ref.build_clipmap(4)
        assert torch.equal(ref.export_atlas(1), peer.export_atlas(1))
    finally:
        if own_group:
            dist.destroy_process_group()


@pytest.mark.gpu
def test_update_nodes_equals_a_fresh_upload():
    """vgi_update_nodes (animated nodes: the world transform of msaaVoxelizer.vert:31-36 on the device) gives the triangle
    soup vgi_set_scene computes on the host for the same matrices: same pairs, same atlases, same octree fragments."""
    import copy
    import torch
    from vk_voxel_cone_tracing_b200 import glm, raster, structs as S, synth
    from vk_voxel_cone_tracing_b200.api import VgiError, VoxelGI
    scene = synth.atrium()
    cfg = S.default_config(64, 3)
    light, shadow = synth.make_light()
    depth = raster.shadow_depth(scene, shadow, 512)
    moved = copy.deepcopy(scene)
    # every node gets another rigid motion + non-uniform scale (the inverse-transpose matters for the normals)
    for i in range(moved.nodes.shape[0]):
        m = moved.nodes[i]["model"].reshape(4, 4).T.astype(np.float64)
        a = 0.3 + 0.2 * i
        rot = np.array([[np.cos(a), 0, np.sin(a), 0.4 * i], [0, 1.0 + 0.1 * i, 0, -0.2], [-np.sin(a), 0, np.cos(a), 0.3], [0, 0, 0, 1]])
        new = (rot @ m)
        moved.nodes[i]["model"] = new.T.astype(np.float32).reshape(16)
        moved.nodes[i]["it_model"] = np.linalg.inv(new).astype(np.float32).reshape(16)      # (inverse transpose) stored [col][row]
    cam = (1.0, 3.0, -2.0)
    outs = []
    for mode in ("fresh", "updated"):
        gi = VoxelGI(cfg)
        gi.set_scene(moved if mode == "fresh" else scene)
        if mode == "updated":
            with pytest.raises(VgiError):
                gi.update_nodes(moved.nodes[:1])            # wrong count
            gi.update_nodes(moved.nodes)
        gi.set_light(light, shadow, depth)
        gi.update_regions(cam)
        gi.build_clipmap(0)
        outs.append((gi.stats().clip_pairs, gi.export_atlas(0).cpu(), gi.export_atlas(1).cpu()))
    assert outs[0][0] == outs[1][0] and outs[0][0] > 10000
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_trace_after_the_regions_moved_equals_a_fresh_context():
    """The tracers probe the footprint byte alone, so k_brick_mask has to zero the bytes of bricks that emptied (regions
    that moved away): the images traced after a walk must equal, bit for bit, those of a context that only ever saw the
    last camera position (no stale byte claims a record that is gone)."""
    import torch
    inp = common.cornell_inputs()
    walked, fresh = _gi_for(inp), _gi_for(inp)
    dgb = walked.upload_gbuffer(inp["gbuffer"])
    cams = [(0.0, 0.0, 0.0), (1.9, -0.7, 1.3), (-2.3, 0.9, -1.7), (40.0, 0.0, 0.0), (0.37, -0.11, 0.29)]
    for cam in cams:
        walked.update_regions(cam)
        walked.build_clipmap(0)
        prm = walked.default_vct_params(8)
        walked.cone_trace(inp["cam"], dgb, prm)
    fresh.update_regions(cams[-1])
    fresh.build_clipmap(0)
    prm = fresh.default_vct_params(8)
    a = walked.cone_trace(inp["cam"], dgb, prm)
    b = fresh.cone_trace(inp["cam"], dgb, prm)
    for x, y in zip(a, b):
        assert torch.equal(x.view(torch.int32), y.view(torch.int32))
    assert float(a[0][..., :3].abs().sum()) > 0.0


def test_shaded_pairs_exclude_what_the_mip_overwrites(oracle):
    """Pairs whose voxel lies in the centre half of its level, off the blend band, only set their occupancy bit: the
    canonical pair count still equals the oracle's, the injection's work list is shorter, and the atlases stay bit-exact
    (test_cornell64_atlases_bit_exact and the full-size tests compare them)."""
    inp = common.cornell_inputs()
    gi, regs, op, rad, pairs = _build_both(oracle, inp)
    st = gi.stats()
    assert st.clip_pairs == pairs
    assert 0 < st.shaded_pairs < st.clip_pairs


def test_overlapped_marches_equal_the_serial_order():
    """vgi_set_trace_overlap: the specular march beside the diffuse one (pre-listed pixels, second stream) writes the
    same two images as the serial order, for whole frames, row bands and interleaved tile rows."""
    import torch
    inp = common.cornell_inputs()
    gi = _gi_for(inp)
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(0)
    dgb = gi.upload_gbuffer(inp["gbuffer"])
    prm = gi.default_vct_params(8)
    ref = [t.clone() for t in gi.cone_trace(inp["cam"], dgb, prm)]
    assert float(ref[1][..., :3].abs().sum()) > 0.0
    for blocks in (1, 2, 7):
        gi.set_trace_overlap(blocks)
        got = gi.cone_trace(inp["cam"], dgb, prm)
        torch.cuda.synchronize()
        for x, y in zip(got, ref):
            assert torch.equal(x.view(torch.int32), y.view(torch.int32)), blocks
    gi.set_trace_overlap(2)
    out = [torch.zeros_like(t) for t in ref]
    h = ref[0].shape[0]
    gi.cone_trace(inp["cam"], dgb, prm, out=out, rows=(0, h // 2))       # tile-aligned bands: the same warps as the whole frame
    gi.cone_trace(inp["cam"], dgb, prm, out=out, rows=(h // 2, h))
    torch.cuda.synchronize()
    for x, y in zip(out, ref):
        assert torch.equal(x.view(torch.int32), y.view(torch.int32))
    out = [torch.zeros_like(t) for t in ref]
    for part in range(3):
        gi.cone_trace(inp["cam"], dgb, prm, out=out, part=(part, 3))
    torch.cuda.synchronize()
    for x, y in zip(out, ref):
        assert torch.equal(x.view(torch.int32), y.view(torch.int32))
    gi.set_trace_overlap(0)
