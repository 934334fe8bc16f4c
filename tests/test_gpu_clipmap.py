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
