"""The stand-alone helper passes on reference-layout atlases (vgi_atlas_*) vs the oracle's restatement of the
reference compute shaders, bit-exact on random RGBA8 atlases."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    cfg = S.default_config(32, 3, downsample_band=3)
    gi = VoxelGI(cfg)
    gi.update_regions((1.3, -0.7, 2.9))
    return gi, cfg, torch


def _random_atlas(cfg, seed):
    from tests import common
    return common.random_atlas(cfg, seed)


def test_clear_region(ctx, oracle):
    gi, cfg, torch = ctx
    a = _random_atlas(cfg, 1)
    for mc, ext, level in (((0, 0, 0), (32, 32, 32), 1), ((5, 30, 17), (9, 4, 20), 2), ((-3, -40, 7), (6, 6, 6), 0)):
        ref = a.copy()
        oracle.clear_region(cfg, ref, mc, ext, level)
        t = torch.from_numpy(a.copy()).cuda()
        gi.atlas_clear_region(t, mc, ext, level)
        assert np.array_equal(t.cpu().numpy(), ref)
        assert (ref != a).any()


def test_copy_alpha(ctx, oracle):
    gi, cfg, torch = ctx
    dst, src = _random_atlas(cfg, 2), _random_atlas(cfg, 3)
    for level in range(3):
        ref = dst.copy()
        oracle.copy_alpha(cfg, level, ref, src)
        t = torch.from_numpy(dst.copy()).cuda()
        gi.atlas_copy_alpha(t, torch.from_numpy(src).cuda(), level)
        assert np.array_equal(t.cpu().numpy(), ref)


@pytest.mark.parametrize("which", [0, 1])
def test_downsample(ctx, oracle, which):
    gi, cfg, torch = ctx
    a = _random_atlas(cfg, 4 + which)
    regs = oracle.regions(cfg, (1.3, -0.7, 2.9))
    ref = a.copy()
    t = torch.from_numpy(a.copy()).cuda()
    for level in (1, 2):
        oracle.downsample(cfg, regs, level, ref, which)
        gi.atlas_downsample(t, which, level)
    got = t.cpu().numpy()
    assert np.array_equal(got, ref), f"{(got != ref).sum()} bytes differ"
    assert (ref != a).any()


@pytest.mark.parametrize("literal", [False, True])
def test_wrap_border(oracle, literal):
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    cfg = S.default_config(32, 2, mode_flags=S.VGI_MODE_BORDER_LITERAL if literal else 0)
    gi = VoxelGI(cfg)
    a = _random_atlas(cfg, 9)
    ref = a.copy()
    oracle.wrap_border(cfg, ref, literal)
    t = torch.from_numpy(a.copy()).cuda()
    gi.atlas_wrap_border(t)
    assert np.array_equal(t.cpu().numpy(), ref)
