"""Specular filter + tonemap (vgi_specular_filter) vs the oracle's literal restatement of
specularFilter.frag / filter.glsl / tonemapping.glsl (257 bilinear taps, 15 x 15 bilateral taps per pixel)."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

TOL = 1e-3   # max abs error on the final float image (north_star tolerance for images)


def _images(h, w, seed):
    rng = np.random.RandomState(seed)
    dif = rng.rand(h, w, 4).astype(np.float32)
    spc = (rng.rand(h, w, 4) * 3.0).astype(np.float32)    # indirect_specular_intensity = 3
    spc[rng.rand(h, w) < 0.6] = (0.0, 0.0, 0.0, 1.0)      # most pixels have no specular cone
    dif[..., 3] = 1.0
    return dif, spc


@pytest.fixture(scope="module")
def gi():
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    return VoxelGI(S.default_config(32, 2))


@pytest.mark.parametrize("method", [0, 1, 5])
@pytest.mark.parametrize("tonemap", [0, 1])
def test_filter_matches_oracle(gi, oracle, method, tonemap):
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    h, w = 45, 71                                           # ragged: not a multiple of the 32 x 8 tile
    dif, spc = _images(h, w, 7 + method)
    prm = S.default_filter_params(method, tonemap)
    ref = oracle.specular_filter(dif, spc, prm)
    out = gi.specular_filter(torch.from_numpy(dif).cuda(), torch.from_numpy(spc).cuda(), prm).cpu().numpy()
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() <= TOL, np.abs(out - ref).max()
    assert common.psnr(out, ref, peak=max(1.0, float(ref.max()))) >= 50.0


def test_filter_on_traced_images_and_defaults(gi, oracle):
    """The pass on real cone-trace outputs (Cornell box, mode 8), default parameters (params = NULL)."""
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    inp = common.cornell_inputs(64, 1024, 128, 128)
    g = VoxelGI(inp["cfg"])
    g.set_scene(inp["scene"])
    g.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    g.update_regions(inp["cam_pos"])
    g.build_clipmap(0)
    d, s = g.cone_trace(inp["cam"], g.upload_gbuffer(inp["gbuffer"]), g.default_vct_params(8))
    out = g.specular_filter(d, s).cpu().numpy()
    ref = oracle.specular_filter(d.cpu().numpy(), s.cpu().numpy(), S.default_filter_params())
    assert np.abs(out - ref).max() <= TOL
    assert (s.cpu().numpy()[..., :3] > 0).any()            # the tall box is metallic: the filter had work to do


def test_filter_error_paths(gi):
    import torch
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VgiError
    a = torch.zeros((8, 8, 4), device="cuda")
    with pytest.raises(VgiError):
        gi.specular_filter(a, a.clone(), out=a)             # aliasing
    with pytest.raises(VgiError):
        gi.specular_filter(a, a.clone(), S.FilterParams(0.0, 0.1, 1, 1))   # gamma 0 with tonemap on
