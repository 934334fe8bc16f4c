"""GPU parity of the sparse-voxel-octree path vs the CPU oracle: fragment multiset and octree pool
bit-exact (the scan-based build reproduces the sequential allocation order, so the raw pool — not only
its canonical form — must equal the oracle's), SVO cone-traced images within 1e-3 / 50 dB."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _sorted_rows(a):
    a = np.ascontiguousarray(a).view(np.uint32).reshape(-1, 2)
    key = a[:, 0].astype(np.uint64) | (a[:, 1].astype(np.uint64) << np.uint64(32))
    return a[np.argsort(key, kind="stable")]


def _setup(inp):
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    return gi


@pytest.mark.parametrize("level", [5, 7])
def test_cornell_fragments_and_octree_bit_exact(oracle, level):
    inp = common.cornell_inputs()
    gi = _setup(inp)
    lo, hi = inp["scene"].world_bbox()
    gi.svo_voxelize(level, lo, hi)
    frags = gi.svo_fragments().cpu().numpy().view(np.uint32)
    osc = oracle.OracleScene(inp["scene"])
    ref = oracle.svo_fragments(level, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    assert frags.shape == ref.shape and ref.shape[0] > 1000
    assert np.array_equal(_sorted_rows(frags), _sorted_rows(ref))   # atomic append order is free
    gi.svo_build()
    nodes = gi.svo_nodes().cpu().numpy().view(np.uint32)
    ref_nodes = oracle.svo_build(level, ref)
    assert nodes.shape == ref_nodes.shape
    a, b = oracle.svo_canonicalize(nodes), oracle.svo_canonicalize(ref_nodes)
    assert np.array_equal(a, b), "topology / colours differ after canonical child ordering"
    assert np.array_equal(nodes, ref_nodes), "scan-order allocation should reproduce the sequential pool exactly"
    st = gi.stats()
    assert st.svo_fragments == ref.shape[0] and st.svo_nodes == ref_nodes.shape[0]


def test_single_triangle_on_the_max_face(oracle):
    """Geometry on the bbox max faces gets voxel coordinate == resolution (inclusive clamp, Q13), i.e. the
    top octant is not always child 0."""
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    tris = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[1, 1, 1], [0, 1, 1], [1, 0, 1]],
                     [[1, 0, 0], [1, 1, 0], [1, 1, 1]]], dtype=np.float32)
    scene = synth.triangle_soup(tris)
    cfg = S.default_config(32, 2)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = np.ones((64, 64), dtype=np.float32)
    gi = _setup(dict(cfg=cfg, scene=scene, light=light, shadow=shadow, shadow_depth=depth))
    lo, hi = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    osc = oracle.OracleScene(scene)
    for level in (1, 2, 4):
        gi.svo_voxelize(level, lo, hi)
        frags = gi.svo_fragments().cpu().numpy().view(np.uint32)
        ref = oracle.svo_fragments(level, lo, hi, osc, light, shadow, depth)
        assert np.array_equal(_sorted_rows(frags), _sorted_rows(ref))
        assert ((ref[:, 0] & 0xfff) == (1 << level)).any()      # a coordinate equal to the resolution exists
        gi.svo_build()
        assert np.array_equal(gi.svo_nodes().cpu().numpy().view(np.uint32), oracle.svo_build(level, ref))


def test_empty_scene_octree(oracle):
    from vk_voxel_cone_tracing_b200 import structs as S, synth
    # one degenerate triangle: no fragments -> the pool is the 8 zeroed children of the root
    scene = synth.triangle_soup(np.zeros((1, 3, 3), dtype=np.float32))
    cfg = S.default_config(32, 2)
    light, shadow = synth.make_light()
    gi = _setup(dict(cfg=cfg, scene=scene, light=light, shadow=shadow, shadow_depth=np.ones((8, 8), np.float32)))
    gi.svo_voxelize(4, (-1, -1, -1), (1, 1, 1))
    assert gi.svo_fragments().shape[0] == 0
    gi.svo_build()
    nodes = gi.svo_nodes().cpu().numpy().view(np.uint32)
    assert nodes.shape == (8, 2) and not nodes.any()


def test_svo_literal_mode_bit_exact(oracle):
    from vk_voxel_cone_tracing_b200 import structs as S
    inp = dict(common.cornell_inputs())
    cfg = S.default_config(64, 6, mode_flags=S.VGI_MODE_SVO_LITERAL)
    inp["cfg"] = cfg
    gi = _setup(inp)
    lo, hi = inp["scene"].world_bbox()
    gi.svo_voxelize(6, lo, hi)
    frags = gi.svo_fragments().cpu().numpy().view(np.uint32)
    osc = oracle.OracleScene(inp["scene"])
    ref = oracle.svo_fragments(6, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"], S.VGI_MODE_SVO_LITERAL)
    assert np.array_equal(_sorted_rows(frags), _sorted_rows(ref))
    gi.svo_build()
    nodes = gi.svo_nodes().cpu().numpy().view(np.uint32)
    assert np.array_equal(nodes, oracle.svo_build(6, ref, mode_flags=S.VGI_MODE_SVO_LITERAL))


@pytest.mark.parametrize("mode", [7, 8])
def test_cornell_svo_cone_trace(oracle, mode):
    from vk_voxel_cone_tracing_b200 import structs as S
    inp = common.cornell_inputs()
    gi = _setup(inp)
    level = 7
    lo, hi = inp["scene"].world_bbox()
    gi.svo_voxelize(level, lo, hi)
    gi.svo_build()
    nodes = gi.svo_nodes().cpu().numpy().view(np.uint32)
    # reference OctreeVoxelConeTracing defaults (OctreeVoxelConeTracing.h:73-80), volumeDimension = 2^level (Q14)
    prm = gi.default_vct_params(mode)
    prm.volume_dimension = float(1 << level)
    prm.voxel_size = float((np.float32(hi) - np.float32(lo)).max() / np.float32(1 << level))
    prm.indirect_diffuse_intensity = 15.0
    prm.occlusion_decay = 3.0
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s = oracle.svo_cone_trace(inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], nodes, lo, hi,
                                         inp["cfg"].level_count)
    d, s = gi.svo_cone_trace(inp["cam"], gi.upload_gbuffer(gb), prm)
    d, s = d.cpu().numpy(), s.cpu().numpy()
    covered = gb["depth"] < 1.0
    assert np.abs(d - ref_d)[covered].max() <= 1e-3
    assert np.abs(s - ref_s)[covered].max() <= 1e-3
    assert common.psnr(d[..., :3], ref_d[..., :3]) >= 50.0 and common.psnr(s[..., :3], ref_s[..., :3]) >= 50.0
    assert ref_d[covered][:, :3].std() > 0.01


def test_cornell_svo_cone_trace_literal_sampling(oracle):
    """VGI_MODE_SVO_LITERAL end to end: fragments and build as shipped (Q21 / Q22 / Q13) AND the un-halved sample positions
    of voxelConeTracing_Octree.frag:330-333 in the tracer — the variant the oracle compares with the reference's shader
    text bit for bit (tests/test_ref_shaders.py)."""
    from vk_voxel_cone_tracing_b200 import structs as S
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    inp = common.cornell_inputs()
    cfg = S.default_config(inp["cfg"].resolution, inp["cfg"].level_count, mode_flags=S.VGI_MODE_SVO_LITERAL)
    gi = VoxelGI(cfg)
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    level = 6
    lo, hi = inp["scene"].world_bbox()
    gi.svo_voxelize(level, lo, hi)
    gi.svo_build()
    nodes = gi.svo_nodes().cpu().numpy().view(np.uint32)
    prm = gi.default_vct_params(8)
    prm.volume_dimension = float(1 << level)
    prm.voxel_size = float((np.float32(hi) - np.float32(lo)).max() / np.float32(1 << level))
    prm.indirect_diffuse_intensity = 15.0
    prm.occlusion_decay = 3.0
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    ref_d, ref_s = oracle.svo_cone_trace(inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], nodes, lo, hi,
                                         cfg.level_count, mode_flags=S.VGI_MODE_SVO_LITERAL)
    can_d, _ = oracle.svo_cone_trace(inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], nodes, lo, hi,
                                     cfg.level_count)
    d, s = gi.svo_cone_trace(inp["cam"], gi.upload_gbuffer(gb), prm)
    d, s = d.cpu().numpy(), s.cpu().numpy()
    covered = gb["depth"] < 1.0
    assert np.abs(d - ref_d)[covered].max() <= 1e-3 and np.abs(s - ref_s)[covered].max() <= 1e-3
    assert np.abs(ref_d - can_d)[covered].max() > 1.2e-3        # the two sampling rules differ by more than the tolerance
