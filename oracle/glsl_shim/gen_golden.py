"""Golden vectors produced by the REFERENCE'S OWN SHADERS (oracle/_ref/libvgi_refshaders.so, built by build_ref.py from
/root/reference/VFS/Shaders) -> tests/golden/ref_shader_golden.npz.

Run here (the reference tree is not on the GPU box): python oracle/glsl_shim/gen_golden.py
tests/test_ref_shaders.py::test_oracle_matches_reference_shader_golden_* then checks the oracle against these files on
any machine, and (when the library is present) re-checks that the stored outputs are what the shaders produce today.

Every input a test needs is stored next to the outputs, so the fixtures do not depend on the synthetic-scene code or on
the oracle's own clipmap build staying the same."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O, refshaders as Rf  # noqa: E402
from vk_voxel_cone_tracing_b200 import structs as S  # noqa: E402
from tests.common import cornell_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_shader_golden.npz")

DS_R, DS_L, DS_CAM = 16, 3, (3.3, -1.2, 7.9)
TRACE_R, TRACE_W, TRACE_H, TRACE_SHADOW = 16, 24, 24, 128
SVO_LEVEL = 6


def struct_bytes(s):
    return np.frombuffer(bytes(s), dtype=np.uint8).copy()


def sparse_atlas(cfg, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=S.atlas_shape(cfg), dtype=np.uint8)
    a[rng.random(a.shape[:3]) < 0.7] = 0
    return a


def random_fragments(level, n, seed):
    """Fragment words in the packing of voxelizer.frag:99-100 (12-bit x, y, split z; RGB + a count nibble of 1)."""
    rng = np.random.default_rng(seed)
    res = 1 << level
    p = np.clip(rng.normal(res / 2, res / 5, size=(n, 3)).astype(np.int64), 0, res).astype(np.uint32)
    col = rng.integers(0, 256, size=(n, 3), dtype=np.uint32)
    x = p[:, 0] | (p[:, 1] << 12) | ((p[:, 2] & 0xff) << 24)
    y = ((p[:, 2] >> 8) << 28) | col[:, 0] | (col[:, 1] << 8) | (col[:, 2] << 16) | (1 << 24)
    return np.stack([x, y], axis=1).astype(np.uint32)


def voxelizer_test_triangles(seed):
    """Random world-space triangles + the tie cases of msaaVoxelizer.geom:27-32 (axis-aligned planes, |nx| == |ny|)."""
    rng = np.random.default_rng(seed)
    tri = rng.normal(0, 3, size=(1200, 3, 3)).astype(np.float32)
    tri[:100, :, 2] = 1.0
    tri[100:200, :, 0] = -2.0
    tri[200:300, :, 1] = 0.5
    d = rng.normal(0, 1, size=(100, 3, 2)).astype(np.float32)
    tri[300:400, :, 0], tri[300:400, :, 1], tri[300:400, :, 2] = d[..., 0], d[..., 0], d[..., 1]
    return tri


def voxel_centres(region, resolution, n, seed):
    rng = np.random.default_rng(seed)
    v = np.array(list(region.min_corner)) + rng.integers(0, resolution, size=(n, 3))
    return ((v + 0.5) * np.float32(region.voxel_size)).astype(np.float32)


def centre_triangles(ctr, voxel_size):
    """One small triangle strictly inside the voxel around each centre."""
    off = np.float32([[0.1, 0, 0.05], [-0.1, 0.1, 0], [0, -0.1, -0.05]]) * np.float32(voxel_size)
    return ctr[:, None, :] + off[None]


def main():
    O.build()
    Rf.build()
    g = {}

    # ---- A-C: atlas compute passes (opacityDownSample / radianceDownSample / borderWrapping / clipmapCleaning / copyAlphaImage)
    cfg = S.default_config(DS_R, DS_L)
    regs = O.regions(cfg, DS_CAM)
    a = sparse_atlas(cfg, 11)
    g["ds_cfg"] = np.array([DS_R, DS_L, cfg.downsample_band], dtype=np.int32)
    g["ds_min_corners"] = np.array([list(r.min_corner) for r in regs], dtype=np.int32)
    g["ds_in"] = a
    for which, name in ((0, "opacity"), (1, "radiance")):
        x = a.copy()
        for level in range(1, DS_L):
            Rf.downsample(cfg, regs, level, x, which)
        g[f"ds_out_{name}"] = x
    for lit, name in ((True, "literal"), (False, "full")):
        x = a.copy()
        Rf.wrap_border(cfg, x, literal=lit)
        g[f"border_out_{name}"] = x
    x = a.copy()
    g["clear_min_corner"] = np.array([5, 0, 11], dtype=np.int32)     # the reference only passes non-negative corners
    g["clear_extent"] = np.array([DS_R, 7, DS_R], dtype=np.uint32)
    Rf.clear_region(cfg, x, g["clear_min_corner"].tolist(), g["clear_extent"].tolist(), 1)
    g["clear_out"] = x
    x = a.copy()
    Rf.copy_alpha(cfg, 2, x, np.ascontiguousarray(a[::-1]))          # source atlas = the input flipped in z
    g["copy_alpha_out"] = x

    # ---- D: the six octreeNode*.comp programs in OctreeBuilder::cmdBuild's order
    frags = random_fragments(SVO_LEVEL, 2000, 12)
    g["svo_level"] = np.array([SVO_LEVEL], dtype=np.int32)
    g["svo_frags"] = frags
    g["svo_nodes"] = Rf.svo_build(SVO_LEVEL, frags)

    # ---- E: voxelConeTracing.frag on the Cornell box (R = 16, L = 6)
    inp = cornell_inputs(resolution=TRACE_R, shadow_size=TRACE_SHADOW, width=TRACE_W, height=TRACE_H)
    cfg = inp["cfg"]
    regs = O.regions(cfg, inp["cam_pos"])
    osc = O.OracleScene(inp["scene"])
    _, rad, _ = O.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    gb = inp["gbuffer"]
    hg = O.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    g["trace_cfg"] = np.array([TRACE_R, cfg.level_count, TRACE_W, TRACE_H], dtype=np.int32)
    g["trace_radiance"] = rad
    for k in ("diffuse", "normal", "specular", "emission", "depth"):
        g[f"trace_gb_{k}"] = np.ascontiguousarray(gb[k])
    g["trace_shadow_depth"] = np.ascontiguousarray(inp["shadow_depth"])
    g["trace_cam"] = struct_bytes(inp["cam"])
    g["trace_light"] = struct_bytes(inp["light"])
    g["trace_shadow"] = struct_bytes(inp["shadow"])
    prm = S.default_vct_params(regs[0], cfg.resolution, 8)
    g["trace_prm"] = struct_bytes(prm)
    for mode, c32 in ((8, 0), (7, 0), (8, 1), (3, 0)):
        prm.rendering_mode, prm.enable_32_cones = mode, c32
        d, s, disc = Rf.cone_trace(cfg, inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
        g[f"trace_out_m{mode}_c{c32}_diffuse"], g[f"trace_out_m{mode}_c{c32}_specular"] = d, s
    g["trace_discarded"] = disc

    # ---- F: voxelConeTracing_Octree.frag on an octree of the same scene (bbox = the shader's constants, Q14)
    cfrags = O.svo_fragments(SVO_LEVEL, Rf.SPONZA_BB_MIN, Rf.SPONZA_BB_MAX, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    nodes = Rf.svo_build(SVO_LEVEL, cfrags)
    g["svotrace_nodes"] = nodes
    prm = S.default_vct_params(regs[0], cfg.resolution, 8)
    prm.volume_dimension = float(1 << SVO_LEVEL)
    g["svotrace_prm"] = struct_bytes(prm)
    for mode in (8, 7):
        prm.rendering_mode = mode
        d, s, _ = Rf.svo_cone_trace(inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], nodes, 6)
        g[f"svotrace_out_m{mode}_diffuse"], g[f"svotrace_out_m{mode}_specular"] = d, s

    # ---- G: specularFilter.frag on the images of E
    d, s = g["trace_out_m8_c0_diffuse"], g["trace_out_m8_c0_specular"]
    for method in (0, 1):
        for tm in (0, 1):
            g[f"filter_out_f{method}_t{tm}"] = Rf.specular_filter(d, s, S.default_filter_params(method, tm))

    # ---- H: msaaInjectRadiance.frag on every 12th sample the oracle shades at level 1 of the same scene
    fr = O.inject_fragments(cfg, regs, 1, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    sel = np.arange(0, fr["pos"].shape[0], 12)
    cnt, coords, vals = Rf.inject_fragments(cfg, regs, 1, fr["pos"][sel], fr["nrm"][sel], fr["mat"][sel], osc.materials,
                                            inp["light"], inp["shadow"], inp["shadow_depth"])
    g["inject_sel"], g["inject_pos"], g["inject_nrm"], g["inject_mat"] = sel, fr["pos"][sel], fr["nrm"][sel], fr["mat"][sel]
    g["inject_out_count"], g["inject_out_coords"], g["inject_out_values"] = cnt, coords, vals

    # ---- I: msaaVoxelizer.geom (dominant axis) and msaaVoxelizer.frag (region test, toroidal texel addressing, 6 stores)
    tri = voxelizer_test_triangles(21)
    g["geom_tris"] = tri
    g["geom_axis"], _ = Rf.voxelizer_geometry(tri)
    cfgv = S.default_config(16, 4)
    regsv = O.regions(cfgv, (-21.7, -21.7, -21.7))      # camera on the diagonal: Q2's scalar clamp is harmless
    g["vox_cfg"] = np.array([16, 4, 2], dtype=np.int32)
    g["vox_min_corner"] = np.array(list(regsv[2].min_corner), dtype=np.int32)
    g["vox_voxel_size"] = np.array([regsv[2].voxel_size], dtype=np.float32)
    ctr = voxel_centres(regsv[2], 16, 300, 22)
    g["vox_positions"] = ctr
    y = O.new_atlas(cfgv)
    Rf.voxelizer_fragments(cfgv, regsv, 2, ctr, y)
    g["vox_out_texels"] = np.argwhere(y[..., 0] > 0).astype(np.int16)
    assert np.all(y[y[..., 0] > 0] == 255)

    # ---- J: voxelizer.frag (SVO fragment shader) on every 40th sample of the same scene at level 5: as shipped (biased
    #         position in; literal colour, Q21/Q22) and with a 1 x 1 white base-colour texture + world position in
    #         (= the oracle's canonical colour)
    lo, hi = inp["scene"].world_bbox()
    smp = O.svo_fragment_samples(5, lo, hi, osc)
    sel = np.arange(0, smp["world"].shape[0], 40)
    g["svofrag_sel"], g["svofrag_world"], g["svofrag_biased"] = sel, smp["world"][sel], smp["biased"][sel]
    for name, pos, white in (("literal", smp["biased"][sel], False), ("canonical", smp["world"][sel], True)):
        disc, words = Rf.svo_fragments(5, pos, smp["nrm"][sel], smp["mat"][sel], osc.materials, inp["light"], inp["shadow"],
                                       inp["shadow_depth"], white_base_color_texture=white)
        g[f"svofrag_out_{name}_discarded"], g[f"svofrag_out_{name}_words"] = disc, words

    # ---- K: gBufferPass.frag on the fragments of a small view of the atrium (23 materials, smooth normals)
    from tests.common import atrium_inputs
    from vk_voxel_cone_tracing_b200 import raster
    ainp = atrium_inputs(resolution=32, shadow_size=64, width=80, height=45, levels=3)
    mat, nrm = raster.gbuffer_attributes(ainp["scene"], ainp["cam"], 80, 45)
    cov = mat >= 0
    g["gbuf_material"], g["gbuf_normal_in"] = mat, nrm
    d, n, s, e, disc = Rf.gbuffer_fragments(nrm[cov], mat[cov], np.ascontiguousarray(ainp["scene"].materials))
    assert not disc.any()
    g["gbuf_out_diffuse"], g["gbuf_out_normal"], g["gbuf_out_specular"], g["gbuf_out_emission"] = d, n, s, e

    np.savez_compressed(OUT, **g)
    print(OUT, os.path.getsize(OUT), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
