"""GPU parity of libvgi.so (through the C ABI) against the REFERENCE'S OWN SHADERS compiled for the CPU
(oracle/_ref/libvgi_refshaders.so, built by oracle/glsl_shim where /root/reference exists; the library travels with the
snapshot). Same inputs as the oracle-based GPU tests; tests/test_ref_shaders.py::test_live_inputs_of_the_gpu_tests checks
on the CPU that oracle and shaders agree on exactly these inputs. Bars: atlases and node pools bit for bit, images
max abs <= 1e-3 and PSNR >= 50 dB."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refshaders():
    from oracle import refshaders as Rf
    if not Rf.available():
        pytest.skip("oracle/_ref/libvgi_refshaders.so not present (built only where the reference tree exists)")
    Rf.lib()
    return Rf


@pytest.fixture(scope="module")
def ctx():
    import torch
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    cfg = common.helper_config()
    gi = VoxelGI(cfg)
    gi.update_regions(common.HELPER_CAM)
    return gi, cfg, torch


@pytest.mark.parametrize("which", [0, 1])
def test_downsample_vs_reference_shader(ctx, oracle, refshaders, which):
    gi, cfg, torch = ctx
    a = common.random_atlas(cfg, 4 + which)
    regs = oracle.regions(cfg, common.HELPER_CAM)
    ref = a.copy()
    t = torch.from_numpy(a.copy()).cuda()
    for level in (1, 2):
        refshaders.downsample(cfg, regs, level, ref, which)      # opacityDownSample.comp / radianceDownSample.comp
        gi.atlas_downsample(t, which, level)
    got = t.cpu().numpy()
    assert np.array_equal(got, ref), f"{(got != ref).sum()} bytes differ"
    assert (ref != a).any()


def test_copy_alpha_and_clear_vs_reference_shader(ctx, refshaders):
    gi, cfg, torch = ctx
    dst, src = common.random_atlas(cfg, 2), common.random_atlas(cfg, 3)
    for level in range(3):
        ref = dst.copy()
        refshaders.copy_alpha(cfg, level, ref, src)               # copyAlphaImage.comp
        t = torch.from_numpy(dst.copy()).cuda()
        gi.atlas_copy_alpha(t, torch.from_numpy(src).cuda(), level)
        assert np.array_equal(t.cpu().numpy(), ref)
    a = common.random_atlas(cfg, 1)
    for mc, ext, level in common.HELPER_CLEAR_CASES:
        ref = a.copy()
        refshaders.clear_region(cfg, ref, mc, ext, level)         # clipmapCleaning.comp
        t = torch.from_numpy(a.copy()).cuda()
        gi.atlas_clear_region(t, mc, ext, level)
        assert np.array_equal(t.cpu().numpy(), ref)
        assert (ref != a).any()


@pytest.mark.parametrize("literal", [False, True])
def test_wrap_border_vs_reference_shader(refshaders, literal):
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    cfg = S.default_config(32, 2, mode_flags=S.VGI_MODE_BORDER_LITERAL if literal else 0)
    gi = VoxelGI(cfg)
    a = common.random_atlas(cfg, 9)
    ref = a.copy()
    refshaders.wrap_border(cfg, ref, literal)                     # borderWrapping.comp (literal: the host's 16-group dispatch)
    t = torch.from_numpy(a.copy()).cuda()
    gi.atlas_wrap_border(t)
    assert np.array_equal(t.cpu().numpy(), ref)


@pytest.mark.parametrize("mode", [7, 8])
def test_cone_trace_vs_reference_shader(refshaders, oracle, mode):
    """voxelConeTracing.frag marching the atlas the GPU built (exported in the reference layout)."""
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    inp = common.cornell_inputs()
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(0)
    rad = np.ascontiguousarray(gi.export_atlas(1).cpu().numpy())
    prm = gi.default_vct_params(mode)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s, disc = refshaders.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"],
                                               inp["shadow_depth"], rad)
    d, s = gi.cone_trace(inp["cam"], gi.upload_gbuffer(gb), prm)
    d, s = d.cpu().numpy(), s.cpu().numpy()
    covered = gb["depth"] < 1.0
    assert np.array_equal(disc.astype(bool), ~covered) and covered.mean() > 0.5
    assert np.abs(d - ref_d)[covered].max() <= 1e-3
    assert np.abs(s - ref_s)[covered].max() <= 1e-3
    assert common.psnr(d[..., :3], ref_d[..., :3]) >= 50.0
    assert common.psnr(s[..., :3], ref_s[..., :3]) >= 50.0


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("tonemap", [0, 1])
def test_specular_filter_vs_reference_shader(refshaders, method, tonemap):
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    gi = VoxelGI(S.default_config(32, 2))
    dif, spc = common.filter_images(45, 71, 7 + method)
    prm = S.default_filter_params(method, tonemap)
    ref = refshaders.specular_filter(dif, spc, prm)               # specularFilter.frag
    out = gi.specular_filter(torch.from_numpy(dif).cuda(), torch.from_numpy(spc).cuda(), prm).cpu().numpy()
    assert np.abs(out - ref).max() <= 1e-3, np.abs(out - ref).max()
    assert common.psnr(out, ref, peak=max(1.0, float(ref.max()))) >= 50.0


@pytest.mark.parametrize("level", [5, 7])
def test_octree_pool_vs_reference_shaders(refshaders, oracle, level):
    """The scan-based CUDA build against the six octreeNode*.comp programs run in OctreeBuilder::cmdBuild's order on the
    same fragment list (the oracle's sequential order): every topology word of the pool."""
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    inp = common.cornell_inputs()
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    lo, hi = inp["scene"].world_bbox()
    gi.svo_voxelize(level, lo, hi)
    gi.svo_build()
    nodes = gi.svo_nodes().cpu().numpy().view(np.uint32)
    osc = oracle.OracleScene(inp["scene"])
    frags = oracle.svo_fragments(level, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    ref_nodes = refshaders.svo_build(level, frags)
    assert nodes.shape == ref_nodes.shape
    assert np.array_equal(nodes[:, 0], ref_nodes[:, 0])
