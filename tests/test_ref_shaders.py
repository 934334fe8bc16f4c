"""The oracle against the REFERENCE'S OWN SHADER TEXT.

oracle/glsl_shim compiles the reference's GLSL for the CPU (oracle/_ref/libvgi_refshaders.so; only possible where
/root/reference exists). Two layers:
  * golden: tests/golden/ref_shader_golden.npz holds inputs + the outputs those shaders produced
    (oracle/glsl_shim/gen_golden.py); the oracle must reproduce them — runs anywhere;
  * live: with the library present, the oracle and the shaders run side by side on fresh random inputs, and the
    stored golden outputs are re-derived.
Integer / byte results are compared bit for bit. Float images were bit-identical when generated (both sides evaluate the
same binary32 expressions without contraction); the golden layer allows 2e-6 for a different libm.

What this pins: the vertex stages (world transform; SVO normalisation, bias and dominant axis), opacity + radiance
down-sampling (incl. the blend band), border wrapping (literal and full), region clearing, copy-alpha, the opacity
voxelizer's dominant axis and texel addressing, the injection shading per fragment, the SVO fragment shader (literal and
canonical colour), the octree build (topology word of every node; colours where one fragment lands in a leaf), the
clipmap and octree cone-tracing fragment shaders in their rendering modes, the specular filter + tonemap, the G-buffer
fragment shader.
What it cannot pin (fixed-function rasterisation): which voxels a triangle covers."""
import ctypes as C
import os

import numpy as np
import pytest

from vk_voxel_cone_tracing_b200 import structs as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_shader_golden.npz")
FLOAT_TOL = 2e-6


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def refshaders():
    from oracle import refshaders as Rf
    if not Rf.available():
        pytest.skip("reference tree not present and oracle/_ref/libvgi_refshaders.so not prebuilt")
    Rf.build()
    return Rf


def _struct(cls, raw):
    return cls.from_buffer_copy(bytes(raw))


def _ds_setup(g):
    r, l, band = (int(v) for v in g["ds_cfg"])
    cfg = S.default_config(r, l, downsample_band=band)
    regs = (S.ClipRegion * l)()
    for i in range(l):
        for k in range(3):
            regs[i].min_corner[k] = int(g["ds_min_corners"][i, k])
            regs[i].extent[k] = r
    return cfg, regs


def _trace_setup(g, oracle):
    r, l, w, h = (int(v) for v in g["trace_cfg"])
    cfg = S.default_config(r, l)
    hg = oracle.HostGBuffer(*(np.ascontiguousarray(g[f"trace_gb_{k}"]) for k in ("diffuse", "normal", "specular", "emission", "depth")))
    cam, light, shadow = _struct(S.Camera, g["trace_cam"]), _struct(S.DirLight, g["trace_light"]), _struct(S.DirLightShadow, g["trace_shadow"])
    return cfg, hg, cam, light, shadow, np.ascontiguousarray(g["trace_shadow_depth"])


# ---------------------------------------------------------------------------------------------------
# golden layer (no reference needed)
# ---------------------------------------------------------------------------------------------------

def test_golden_downsample_border_clear_copyalpha(oracle, golden):
    g = golden
    cfg, regs = _ds_setup(g)
    for which, name in ((0, "opacity"), (1, "radiance")):
        x = g["ds_in"].copy()
        for level in range(1, cfg.level_count):
            oracle.downsample(cfg, regs, level, x, which)
        assert np.array_equal(x, g[f"ds_out_{name}"]), f"{name} down-sample differs from the reference shader"
        assert not np.array_equal(x, g["ds_in"])
    for lit, name in ((True, "literal"), (False, "full")):
        x = g["ds_in"].copy()
        oracle.wrap_border(cfg, x, literal=lit)
        assert np.array_equal(x, g[f"border_out_{name}"]), f"border wrap ({name}) differs from the reference shader"
    x = g["ds_in"].copy()
    oracle.clear_region(cfg, x, g["clear_min_corner"].tolist(), g["clear_extent"].tolist(), 1)
    assert np.array_equal(x, g["clear_out"])
    x = g["ds_in"].copy()
    oracle.copy_alpha(cfg, 2, x, np.ascontiguousarray(g["ds_in"][::-1]))
    assert np.array_equal(x, g["copy_alpha_out"])


def test_golden_octree_build(oracle, golden):
    g = golden
    level, frags, want = int(g["svo_level"][0]), np.ascontiguousarray(g["svo_frags"]), g["svo_nodes"]
    for flags in (S.VGI_MODE_SVO_LITERAL, 0):   # the topology is the same in both modes
        got = oracle.svo_build(level, frags, mode_flags=flags)
        assert got.shape == want.shape
        assert np.array_equal(got[:, 0], want[:, 0]), "node pool topology differs from the reference's octreeNode*.comp"
    # colours: leaves that received exactly one fragment hold that fragment's RGB in both (Q10 only changes how several
    # fragments are averaged; the oracle's alpha is 255, the reference keeps its count nibble)
    got = oracle.svo_build(level, frags, mode_flags=S.VGI_MODE_SVO_LITERAL)
    leaf = (want[:, 0] == 0x80000000) & ((want[:, 1] >> 24) == 1)   # flagged, no children, count nibble 1
    assert leaf.sum() > 100
    assert np.array_equal(got[leaf, 1] & 0xffffff, want[leaf, 1] & 0xffffff)


@pytest.mark.parametrize("mode,c32", [(8, 0), (7, 0), (8, 1), (3, 0)])
def test_golden_clipmap_cone_trace(oracle, golden, mode, c32):
    g = golden
    cfg, hg, cam, light, shadow, sd = _trace_setup(g, oracle)
    prm = _struct(S.VctParams, g["trace_prm"])
    prm.rendering_mode, prm.enable_32_cones = mode, c32
    d, s, _ = oracle.cone_trace(cfg, cam, hg, prm, light, shadow, sd, np.ascontiguousarray(g["trace_radiance"]))
    cov = g["trace_discarded"] == 0
    assert cov.sum() > 100
    assert float(np.abs(d - g[f"trace_out_m{mode}_c{c32}_diffuse"])[cov].max()) <= FLOAT_TOL
    assert float(np.abs(s - g[f"trace_out_m{mode}_c{c32}_specular"])[cov].max()) <= FLOAT_TOL
    if mode == 8:   # a real image: direct + indirect + specular all present
        assert float(g[f"trace_out_m8_c{c32}_diffuse"][..., :3].max()) > 0.05
        assert float(g[f"trace_out_m8_c{c32}_specular"][..., :3].max()) > 0.05


@pytest.mark.parametrize("mode", [8, 7])
def test_golden_octree_cone_trace(oracle, golden, mode):
    from oracle import refshaders as Rf     # constants only
    g = golden
    _, hg, cam, light, shadow, sd = _trace_setup(g, oracle)
    prm = _struct(S.VctParams, g["svotrace_prm"])
    prm.rendering_mode = mode
    d, s = oracle.svo_cone_trace(cam, hg, prm, light, shadow, sd, np.ascontiguousarray(g["svotrace_nodes"]),
                                 Rf.SPONZA_BB_MIN, Rf.SPONZA_BB_MAX, 6, mode_flags=S.VGI_MODE_SVO_LITERAL)
    cov = g["trace_discarded"] == 0
    assert float(np.abs(d - g[f"svotrace_out_m{mode}_diffuse"])[cov].max()) <= FLOAT_TOL
    assert float(np.abs(s - g[f"svotrace_out_m{mode}_specular"])[cov].max()) <= FLOAT_TOL


@pytest.mark.parametrize("method,tonemap", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_golden_specular_filter(oracle, golden, method, tonemap):
    g = golden
    out = oracle.specular_filter(g["trace_out_m8_c0_diffuse"], g["trace_out_m8_c0_specular"], S.default_filter_params(method, tonemap))
    assert float(np.abs(out - g[f"filter_out_f{method}_t{tonemap}"]).max()) <= FLOAT_TOL


def _check_injection(cfg, level, fr, sel, cnt, coords, vals):
    """Oracle samples fr[sel] against what msaaInjectRadiance.frag wrote for them (one invocation each, cleared image).
    * the same number of face texels (3, or 6 for emissive materials; 0 = discarded);
    * the same faces, in the same order;
    * every colour byte is the truncation of a value inside the oracle's 16-bit fixed-point cell: with q the oracle's
      round(v * 65536), the shader's uint(v * 255) must lie in [floor((q - 0.5) * 255 / 65536), floor((q + 0.5) * 255 / 65536)]
      - i.e. the shaded radiance agrees to 2^-17 and the quantisation rule (truncate once) is the same;
    * the texel: the shader addresses the voxel that CONTAINS the sample point; the oracle attributes the sample to the
      voxel whose overlap with the triangle produced it (canonical conservative coverage). They coincide when
      the point lies inside that voxel, which is required here; elsewhere the difference is the documented deviation."""
    r, rb = cfg.resolution, cfg.resolution + 2
    assert np.array_equal(cnt, fr["nfaces"][sel])
    assert set(np.unique(cnt).tolist()) <= {0, 3, 6} and (cnt == 3).sum() > 100
    k = np.arange(6)[None, :] < cnt[:, None]
    faces = fr["faces"][sel]
    assert np.array_equal((coords[..., 0] // rb)[k], faces[k])
    q = fr["q"][sel].astype(np.float64)
    lo = np.minimum(np.floor((q - 0.5) * 255.0 / 65536.0), 255)
    hi = np.minimum(np.floor((q + 0.5) * 255.0 / 65536.0), 255)
    for ch in range(3):
        b = ((vals >> (8 * ch)) & 0xff).astype(np.float64)
        assert np.all((b >= lo[..., ch])[k]) and np.all((b <= hi[..., ch])[k]), ch
    assert np.all(((vals >> 24) == 1)[k])     # the CAS average's count
    vox = fr["voxel"][sel]
    want = np.stack([1 + (vox[:, 0] & (r - 1)), 1 + (vox[:, 1] & (r - 1)) + level * rb, 1 + (vox[:, 2] & (r - 1))], axis=1)
    got = np.stack([coords[:, 0, 0] % rb, coords[:, 0, 1], coords[:, 0, 2]], axis=1)
    same = np.all(want == got, axis=1) | (cnt == 0)
    vs = 16.0 * (1 << level) / r
    rel = fr["pos"][sel].astype(np.float64) / vs - vox            # position inside its voxel <=> every component in [0, 1)
    inside = np.all((rel > 1e-4) & (rel < 1.0 - 1e-4), axis=1)
    assert np.all(same[inside]), "a sample inside the oracle's voxel was addressed elsewhere by the shader"
    assert inside.sum() >= 100      # (walls of the Cornell box lie exactly on voxel faces at the coarse levels)


def test_golden_injection_fragments(oracle, golden):
    """Depends on the Cornell generator: if this fails on the position check, re-run oracle/glsl_shim/gen_golden.py."""
    from tests.common import cornell_inputs
    g = golden
    r, l, w, h = (int(v) for v in g["trace_cfg"])
    inp = cornell_inputs(resolution=r, shadow_size=g["trace_shadow_depth"].shape[0], width=w, height=h)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, inp["cam_pos"])
    osc = oracle.OracleScene(inp["scene"])
    fr = oracle.inject_fragments(cfg, regs, 1, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    sel = g["inject_sel"]
    assert np.array_equal(fr["pos"][sel], g["inject_pos"]) and np.array_equal(fr["nrm"][sel], g["inject_nrm"])
    _check_injection(cfg, 1, fr, sel, g["inject_out_count"], g["inject_out_coords"], g["inject_out_values"])


def _gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(GOLDEN), "..", "..", "oracle", "glsl_shim", "gen_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen


def test_golden_voxelizer_axis_and_addressing(oracle, golden):
    g = golden
    assert np.array_equal(oracle.dominant_axis(g["geom_tris"]), g["geom_axis"])
    assert len(np.unique(g["geom_axis"])) == 3
    r, l, level = (int(v) for v in g["vox_cfg"])
    cfg = S.default_config(r, l)
    regs = (S.ClipRegion * l)()
    for k in range(3):
        regs[level].min_corner[k] = int(g["vox_min_corner"][k])
        regs[level].extent[k] = r
    regs[level].voxel_size = float(g["vox_voxel_size"][0])
    x = oracle.new_atlas(cfg)
    oracle.voxelize_level(cfg, regs, level, oracle.TriangleSoup(_gen().centre_triangles(g["vox_positions"], regs[level].voxel_size)), x)
    assert np.array_equal(np.argwhere(x[..., 0] > 0).astype(np.int16), g["vox_out_texels"])
    assert np.all(x[x[..., 0] > 0] == 255)


def _check_svo_fragments(level, smp, sel, lit, can, out_lit, out_can):
    """voxelizer.frag per oracle sample. out_* = (discarded, words) of the shader; lit / can = the oracle's fragment
    lists in literal / canonical mode (one fragment per sample, same order, when nothing is discarded).
    * as shipped (biased position in): nothing the oracle keeps is discarded, the colour word (28 bits incl. the alpha
      nibble) is bit-identical, and the packed voxel is the oracle's whenever the sample lies inside that voxel;
    * with a 1 x 1 white base-colour texture bound and the world position in, the shader's `texture * factor` colour is
      the oracle's canonical colour (Q21 / Q22): bit-identical colour words."""
    n = smp["world"].shape[0]
    assert lit.shape[0] == n and can.shape[0] == n, "a sample was discarded: the fragment lists no longer line up"
    res = 1 << level
    (dl, wl), (dc, wc) = out_lit, out_can
    assert not dl.any() and not dc.any()
    assert np.array_equal(wl[:, 1] & 0x0fffffff, lit[sel, 1] & 0x0fffffff)
    assert np.array_equal(wc[:, 1] & 0x0fffffff, can[sel, 1] & 0x0fffffff)
    assert len(np.unique(can[sel, 1] & 0xffffff)) > 20            # real shading, not a constant
    unpack = lambda w: np.stack([w[:, 0] & 0xfff, (w[:, 0] >> 12) & 0xfff, ((w[:, 0] >> 24) & 0xff) | ((w[:, 1] >> 28) << 8)], axis=1)  # noqa: E731
    same = np.all(unpack(wl) == unpack(lit[sel]), axis=1)
    rel = smp["biased"][sel].astype(np.float64) * res - smp["voxel"][sel]
    inside = np.all((rel > 1e-3) & (rel < 1.0 - 1e-3), axis=1)
    assert inside.sum() >= 100 and np.all(same[inside])


def test_golden_svo_fragment_shader(oracle, golden):
    """Depends on the Cornell generator like test_golden_injection_fragments."""
    from tests.common import cornell_inputs
    g = golden
    r, l, w, h = (int(v) for v in g["trace_cfg"])
    inp = cornell_inputs(resolution=r, shadow_size=g["trace_shadow_depth"].shape[0], width=w, height=h)
    osc = oracle.OracleScene(inp["scene"])
    lo, hi = inp["scene"].world_bbox()
    smp = oracle.svo_fragment_samples(5, lo, hi, osc)
    sel = g["svofrag_sel"]
    assert np.array_equal(smp["world"][sel], g["svofrag_world"]) and np.array_equal(smp["biased"][sel], g["svofrag_biased"])
    args = (5, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    _check_svo_fragments(5, smp, sel, oracle.svo_fragments(*args, S.VGI_MODE_SVO_LITERAL), oracle.svo_fragments(*args, 0),
                         (g["svofrag_out_literal_discarded"], g["svofrag_out_literal_words"]),
                         (g["svofrag_out_canonical_discarded"], g["svofrag_out_canonical_words"]))


def _check_gbuffer(gb, cov, d, n, s, e):
    """The host rasteriser's G-buffer (which vgi_render_gbuffer reproduces bit for bit, tests/test_gpu_raster.py) against
    gBufferPass.frag's four colour outputs after the fixed-function format conversion: RGBA8 UNORM round-to-nearest for
    diffuse (+ roughness) and specular (+ metallic), binary16 round-to-nearest-even for emission and the encoded normal.
    Colours bit-identical; the normal within one binary16 ulp (the rasteriser interpolates and normalises in binary64)."""
    to8 = lambda x: np.floor(np.clip(x, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)  # noqa: E731
    assert cov.sum() > 1000
    assert np.array_equal(to8(d), gb["diffuse"][cov]) and np.array_equal(to8(s), gb["specular"][cov])
    assert np.array_equal(e.astype(np.float16).view(np.uint16), gb["emission"][cov])
    dn = np.abs(n.astype(np.float16).view(np.uint16).astype(np.int32) - gb["normal"][cov].astype(np.int32))
    assert dn.max() <= 1 and (dn == 0).mean() > 0.99
    assert len(np.unique(gb["diffuse"][cov].reshape(-1, 4), axis=0)) > 10       # many materials in view


def test_golden_gbuffer_fragment_shader(golden):
    """Depends on the atrium generator: if the attribute check fails, re-run oracle/glsl_shim/gen_golden.py."""
    from tests.common import atrium_inputs
    from vk_voxel_cone_tracing_b200 import raster
    g = golden
    inp = atrium_inputs(resolution=32, shadow_size=64, width=80, height=45, levels=3)
    mat, nrm = raster.gbuffer_attributes(inp["scene"], inp["cam"], 80, 45)
    assert np.array_equal(mat, g["gbuf_material"]) and np.array_equal(nrm, g["gbuf_normal_in"])
    cov = mat >= 0
    assert np.array_equal(cov, inp["gbuffer"]["depth"] < 1.0)
    _check_gbuffer(inp["gbuffer"], cov, g["gbuf_out_diffuse"], g["gbuf_out_normal"], g["gbuf_out_specular"], g["gbuf_out_emission"])


# ---------------------------------------------------------------------------------------------------
# live layer (needs oracle/_ref/libvgi_refshaders.so, i.e. the reference tree or a prebuilt library)
# ---------------------------------------------------------------------------------------------------

def test_live_golden_outputs_are_what_the_shaders_produce(refshaders, golden):
    g = golden
    cfg, regs = _ds_setup(g)
    x = g["ds_in"].copy()
    for level in range(1, cfg.level_count):
        refshaders.downsample(cfg, regs, level, x, 0)
    assert np.array_equal(x, g["ds_out_opacity"])
    assert np.array_equal(refshaders.svo_build(int(g["svo_level"][0]), g["svo_frags"]), g["svo_nodes"])


@pytest.mark.parametrize("resolution,cam", [(16, (0.0, 0.0, 0.0)), (32, (3.3, -1.2, 7.9)), (32, (-40.0, 2.5, 13.0))])
def test_live_atlas_passes_random(oracle, refshaders, resolution, cam):
    rng = np.random.default_rng(resolution + int(abs(cam[0]) * 10))
    cfg = S.default_config(resolution, 4)
    regs = oracle.regions(cfg, cam)
    for which in (0, 1):
        a = rng.integers(0, 256, size=S.atlas_shape(cfg), dtype=np.uint8)
        a[rng.random(a.shape[:3]) < 0.5] = 0
        for level in range(1, cfg.level_count):
            x, y = a.copy(), a.copy()
            oracle.downsample(cfg, regs, level, x, which)
            refshaders.downsample(cfg, regs, level, y, which)
            assert np.array_equal(x, y), (which, level)
            assert not np.array_equal(x, a)
    a = rng.integers(0, 256, size=S.atlas_shape(cfg), dtype=np.uint8)
    b = rng.integers(0, 256, size=S.atlas_shape(cfg), dtype=np.uint8)
    for lit in (True, False):
        x, y = a.copy(), a.copy()
        oracle.wrap_border(cfg, x, literal=lit)
        refshaders.wrap_border(cfg, y, literal=lit)
        assert np.array_equal(x, y)
    for level in range(cfg.level_count):
        x, y = a.copy(), a.copy()
        oracle.copy_alpha(cfg, level, x, b)
        refshaders.copy_alpha(cfg, level, y, b)
        assert np.array_equal(x, y)
        x, y = a.copy(), a.copy()
        mc, ext = [3, 9, 0], [resolution, resolution // 2, resolution]
        oracle.clear_region(cfg, x, mc, ext, level)
        refshaders.clear_region(cfg, y, mc, ext, level)
        assert np.array_equal(x, y)


@pytest.mark.parametrize("level", [3, 5, 8])
def test_live_octree_build_random(oracle, refshaders, level):
    frags = _gen().random_fragments(level, 3000, 100 + level)
    want = refshaders.svo_build(level, frags)
    got = oracle.svo_build(level, frags, mode_flags=S.VGI_MODE_SVO_LITERAL)
    assert got.shape == want.shape and np.array_equal(got[:, 0], want[:, 0])
    # one fragment per (halved) voxel: every colour word's RGB agrees, leaves and interior means alike
    p = np.stack([frags[:, 0] & 0xfff, (frags[:, 0] >> 12) & 0xfff, ((frags[:, 0] >> 24) & 0xff) | ((frags[:, 1] >> 20) & 0xf00)], axis=1) >> 1
    _, idx = np.unique(p[:, 0].astype(np.uint64) | (p[:, 1].astype(np.uint64) << 16) | (p[:, 2].astype(np.uint64) << 32), return_index=True)
    uniq = np.ascontiguousarray(frags[np.sort(idx)])
    want = refshaders.svo_build(level, uniq)
    got = oracle.svo_build(level, uniq, mode_flags=S.VGI_MODE_SVO_LITERAL)
    assert np.array_equal(got[:, 0], want[:, 0])
    assert np.array_equal(got[:, 1] & 0xffffff, want[:, 1] & 0xffffff)


def test_live_cone_trace_and_filter_cornell(oracle, refshaders):
    from tests.common import cornell_inputs
    inp = cornell_inputs(resolution=32, shadow_size=512, width=40, height=40)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, inp["cam_pos"])
    osc = oracle.OracleScene(inp["scene"])
    _, rad, _ = oracle.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    cov = gb["depth"] < 1.0
    d8 = s8 = None
    for mode in range(9):
        for c32 in (0, 1):
            prm = S.default_vct_params(regs[0], cfg.resolution, mode)
            prm.enable_32_cones = c32
            d0, s0, _ = oracle.cone_trace(cfg, inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
            d1, s1, disc = refshaders.cone_trace(cfg, inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
            assert np.array_equal(disc.astype(bool), ~cov)
            assert np.array_equal(d0[cov], d1[cov]) and np.array_equal(s0[cov], s1[cov]), (mode, c32)
            if mode == 8 and c32 == 0:
                d8, s8 = d1, s1
    assert float(d8[..., :3].max()) > 0.05 and float(s8[..., :3].max()) > 0.05
    for method in (0, 1):
        for tm in (0, 1):
            fp = S.default_filter_params(method, tm)
            assert np.array_equal(oracle.specular_filter(d8, s8, fp), refshaders.specular_filter(d8, s8, fp))
    # octree tracer on the same G-buffer (tree built by the reference's programs from the oracle's canonical fragments)
    frags = oracle.svo_fragments(6, refshaders.SPONZA_BB_MIN, refshaders.SPONZA_BB_MAX, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    nodes = refshaders.svo_build(6, frags)
    for mode in (8, 7, 5, 4, 3):
        prm = S.default_vct_params(regs[0], cfg.resolution, mode)
        prm.volume_dimension = 64.0
        a_d, a_s = oracle.svo_cone_trace(inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], nodes,
                                         refshaders.SPONZA_BB_MIN, refshaders.SPONZA_BB_MAX, 6, mode_flags=S.VGI_MODE_SVO_LITERAL)
        b_d, b_s, _ = refshaders.svo_cone_trace(inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], nodes, 6)
        assert np.array_equal(a_d[cov], b_d[cov]) and np.array_equal(a_s[cov], b_s[cov]), mode


@pytest.mark.parametrize("level", [0, 2, 4])
def test_live_injection_fragments(oracle, refshaders, level):
    from tests.common import cornell_inputs
    inp = cornell_inputs(resolution=32, shadow_size=512, width=16, height=16)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, inp["cam_pos"])
    osc = oracle.OracleScene(inp["scene"])
    fr = oracle.inject_fragments(cfg, regs, level, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    sel = np.arange(fr["pos"].shape[0])
    cnt, coords, vals = refshaders.inject_fragments(cfg, regs, level, fr["pos"], fr["nrm"], fr["mat"], osc.materials,
                                                    inp["light"], inp["shadow"], inp["shadow_depth"])
    _check_injection(cfg, level, fr, sel, cnt, coords, vals)


@pytest.mark.parametrize("level", [0, 1])
def test_live_textured_injection(oracle, refshaders, level):
    """Textured materials (base-colour, emissive, occlusion alpha test; REPEAT coordinates beyond [0, 1]; a 3 x 5 texture):
    every sample the oracle shades, with its interpolated texture coordinate, through the reference's own
    msaaInjectRadiance.frag with the same textures bound — same faces, same colours (the 16-bit fixed-point cell rule of
    _check_injection), and no sample the oracle kept is discarded by the shader's alpha test."""
    from vk_voxel_cone_tracing_b200 import raster, synth
    scene = synth.textured_cornell()
    tex = synth.procedural_textures()
    cfg = S.default_config(32, 2)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = raster.shadow_depth(scene, shadow, 256)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    oracle.set_textures(tex)
    try:
        osc = oracle.OracleScene(scene)
        fr = oracle.inject_fragments(cfg, regs, level, osc, light, shadow, depth)
    finally:
        oracle.set_textures([])
    textured = np.isin(fr["mat"], [i for i, m in enumerate(osc.materials)
                                   if max(int(m["base_color_texture"]), int(m["emissive_texture"]), int(m["occlusion_texture"])) > -1])
    assert textured.sum() > 200 and (fr["uv"][textured] != 0).any()
    sel = np.arange(fr["pos"].shape[0])
    cnt, coords, vals = refshaders.inject_fragments(cfg, regs, level, fr["pos"], fr["nrm"], fr["mat"], osc.materials,
                                                    light, shadow, depth, texcoord=fr["uv"], textures=tex)
    _check_injection(cfg, level, fr, sel, cnt, coords, vals)


def test_live_voxelizer_geometry_axis(oracle, refshaders):
    gen = _gen()
    for seed in (1, 2):
        tri = gen.voxelizer_test_triangles(seed)
        axis, clip = refshaders.voxelizer_geometry(tri)
        assert np.array_equal(oracle.dominant_axis(tri), axis)
        assert np.array_equal(clip[..., :3], tri) and np.all(clip[..., 3] == 1.0)     # identity uViewProj


@pytest.mark.parametrize("cam", [(0.0, 0.0, 0.0), (5.3, 5.3, 5.3), (-21.7, -21.7, -21.7)])
def test_live_voxelizer_fragment_addressing(oracle, refshaders, cam):
    """Opacity stores: region test, fract-wrapped toroidal texel, six faces. Cameras on the x = y = z diagonal keep the
    scalar clamp of msaaVoxelizer.frag:46 (Q2) harmless, so shader and oracle must agree texel for texel."""
    gen = _gen()
    cfg = S.default_config(32, 4)
    regs = oracle.regions(cfg, cam)
    for level in range(4):
        ctr = gen.voxel_centres(regs[level], 32, 300, 7 + level)
        x, y = oracle.new_atlas(cfg), oracle.new_atlas(cfg)
        oracle.voxelize_level(cfg, regs, level, oracle.TriangleSoup(gen.centre_triangles(ctr, regs[level].voxel_size)), x)
        disc = refshaders.voxelizer_fragments(cfg, regs, level, ctr, y)
        assert not disc.any() and np.array_equal(x, y)
        assert int((x[..., 0] > 0).sum()) > 1500


def test_live_q2_scalar_clamp_quantified(oracle, refshaders):
    """SURVEY Q2, measured: with the camera off the diagonal the shader clamps y and z to the X bounds of the region
    (float(vec3) takes .x). It addresses the oracle's (per-axis, canonical) texel exactly when every coordinate of the
    fragment lies inside the x bounds, and a different one otherwise."""
    gen = _gen()
    cfg = S.default_config(32, 3)
    regs = oracle.regions(cfg, (5.3, -1.2, 7.9))
    level = 0
    mc, vs = np.array(list(regs[level].min_corner)), np.float32(regs[level].voxel_size)
    assert len(set(mc.tolist())) == 3
    ctr = gen.voxel_centres(regs[level], 32, 250, 3)
    tris = gen.centre_triangles(ctr, vs)
    inside = np.all((ctr > mc[0] * vs) & (ctr < (mc[0] + 32) * vs), axis=1)
    assert 20 < inside.sum() < 230
    for i in range(ctr.shape[0]):
        x, y = oracle.new_atlas(cfg), oracle.new_atlas(cfg)
        oracle.voxelize_level(cfg, regs, level, oracle.TriangleSoup(tris[i:i + 1]), x)
        refshaders.voxelizer_fragments(cfg, regs, level, ctr[i:i + 1], y)
        assert np.array_equal(x, y) == bool(inside[i]), i


@pytest.mark.parametrize("level", [5, 7])
def test_live_svo_fragment_shader(oracle, refshaders, level):
    from tests.common import cornell_inputs
    inp = cornell_inputs(resolution=32, shadow_size=512, width=16, height=16)
    osc = oracle.OracleScene(inp["scene"])
    lo, hi = inp["scene"].world_bbox()
    smp = oracle.svo_fragment_samples(level, lo, hi, osc)
    sel = np.arange(smp["world"].shape[0])
    args = (level, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    common = (smp["nrm"], smp["mat"], osc.materials, inp["light"], inp["shadow"], inp["shadow_depth"])
    _check_svo_fragments(level, smp, sel, oracle.svo_fragments(*args, S.VGI_MODE_SVO_LITERAL), oracle.svo_fragments(*args, 0),
                         refshaders.svo_fragments(level, smp["biased"], *common),
                         refshaders.svo_fragments(level, smp["world"], *common, white_base_color_texture=True))


def test_live_gbuffer_fragment_shader(refshaders):
    from tests.common import atrium_inputs, cornell_inputs
    from vk_voxel_cone_tracing_b200 import raster
    for inp, w, h in ((atrium_inputs(resolution=32, shadow_size=64, width=160, height=90, levels=3), 160, 90),
                      (cornell_inputs(resolution=32, shadow_size=128, width=96, height=96), 96, 96)):
        mat, nrm = raster.gbuffer_attributes(inp["scene"], inp["cam"], w, h)
        cov = mat >= 0
        d, n, s, e, disc = refshaders.gbuffer_fragments(nrm[cov], mat[cov], np.ascontiguousarray(inp["scene"].materials))
        assert not disc.any() and np.array_equal(cov, inp["gbuffer"]["depth"] < 1.0)
        if w == 160:
            _check_gbuffer(inp["gbuffer"], cov, d, n, s, e)
        else:   # few materials in this view: colours only
            to8 = lambda x: np.floor(np.clip(x, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)  # noqa: E731
            assert np.array_equal(to8(d), inp["gbuffer"]["diffuse"][cov]) and np.array_equal(to8(s), inp["gbuffer"]["specular"][cov])


def test_live_vertex_stage_world_transform(oracle, refshaders):
    """msaaVoxelizer.vert over every triangle vertex in draw order (the atrium has rotated / scaled node matrices) against
    the oracle's world-space triangle soup (vgo_scene_triangles): positions and itModel-transformed normals, bit for bit."""
    from vk_voxel_cone_tracing_b200 import synth
    for scene in (synth.cornell_box(), synth.atrium()):
        osc = oracle.OracleScene(scene)
        pos, nrm = refshaders.voxelizer_vertices(scene)
        assert pos.shape == osc.pos.shape and np.array_equal(pos, osc.pos) and np.array_equal(nrm, osc.nrm)


def test_live_svo_vertex_and_geometry_stages(oracle, refshaders):
    """voxelizer.vert (normalisation by the scene box) chained into voxelizer.geom (dominant axis from the normalised
    positions, bias to [0,1], normal pass-through) against the oracle's svo_to_grid / cross_and_axis, bit for bit."""
    from vk_voxel_cone_tracing_b200 import synth
    for scene in (synth.cornell_box(), synth.atrium()):
        osc = oracle.OracleScene(scene)
        lo, hi = scene.world_bbox()
        ndc0, biased0, axis0 = oracle.svo_vertex_stage(7, lo, hi, osc)
        ndc, biased, axis, nrm = refshaders.svo_vertex_stage(scene, lo, hi)
        assert np.array_equal(ndc, ndc0) and np.array_equal(biased, biased0) and np.array_equal(axis, axis0)
        assert np.array_equal(nrm, osc.nrm) and len(np.unique(axis)) == 3


def test_live_whole_build_with_reference_passes(oracle, refshaders):
    """The pass sequences of VoxelizationPass::render / RadianceInjectionPass::render on the atrium with an off-diagonal
    camera (negative, unequal region corners; six levels), every programmable post-raster pass run by the REFERENCE's shader:
    oracle coverage + injection, then copyAlphaImage / opacity + radiance down-sampling / border wrap from the shader library.
    Both atlases must equal the oracle's own full passes byte for byte; the traced images likewise."""
    from tests.common import atrium_inputs
    inp = atrium_inputs(resolution=32, shadow_size=256, width=96, height=54, levels=6)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, inp["cam_pos"])
    osc = oracle.OracleScene(inp["scene"])
    op, rad, _ = oracle.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    assert len({tuple(r.min_corner) for r in regs}) > 1 and min(regs[0].min_corner) < 0
    op2, rad2 = oracle.new_atlas(cfg), oracle.new_atlas(cfg)
    for l in range(cfg.level_count):
        oracle.voxelize_level(cfg, regs, l, osc, op2)
    for l in range(1, cfg.level_count):
        refshaders.downsample(cfg, regs, l, op2, 0)
    refshaders.wrap_border(cfg, op2, literal=False)
    assert np.array_equal(op, op2)
    for l in range(cfg.level_count):
        oracle.inject_level(cfg, regs, l, osc, inp["light"], inp["shadow"], inp["shadow_depth"], rad2)
    for l in range(cfg.level_count):
        refshaders.copy_alpha(cfg, l, rad2, op2)
    for l in range(1, cfg.level_count):
        refshaders.downsample(cfg, regs, l, rad2, 1)
    refshaders.wrap_border(cfg, rad2, literal=False)
    assert np.array_equal(rad, rad2) and (rad[..., :3] > 0).sum() > 1000
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    cov = gb["depth"] < 1.0
    prm = S.default_vct_params(regs[0], cfg.resolution, 8)
    d0, s0, _ = oracle.cone_trace(cfg, inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
    d1, s1, _ = refshaders.cone_trace(cfg, inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad2)
    assert np.array_equal(d0[cov], d1[cov]) and np.array_equal(s0[cov], s1[cov])


@pytest.mark.parametrize("case", range(6))
def test_live_cone_trace_random_cameras_and_parameters(oracle, refshaders, case):
    """Seeded sweep: random camera pose inside either scene, random frame index (cadence), rendering mode, cone set, start
    offset, step factor, occlusion decay, intensities - oracle and voxelConeTracing.frag bit for bit (NaNs in the same places)."""
    from tests.common import atrium_inputs, cornell_inputs
    from vk_voxel_cone_tracing_b200 import raster, synth
    rng = np.random.default_rng(100 + case)
    r = int(rng.choice([32, 64]))
    if case % 2 == 0:
        inp, (w, h) = cornell_inputs(resolution=r, shadow_size=256, width=40, height=40), (40, 40)
        cam_pos = tuple(rng.uniform(-3, 3, 3).tolist())
    else:
        inp, (w, h) = atrium_inputs(resolution=r, shadow_size=256, width=48, height=27, levels=6), (48, 27)
        cam_pos = (float(rng.uniform(-12, 12)), float(rng.uniform(0, 9)), float(rng.uniform(-7, 7)))
    look = rng.normal(size=3)
    look /= np.linalg.norm(look)
    cam = synth.make_camera(cam_pos, tuple(look.tolist()), aspect=w / h)
    gb = raster.gbuffer(inp["scene"], cam, w, h)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, cam_pos)
    osc = oracle.OracleScene(inp["scene"])
    _, rad, _ = oracle.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], int(rng.integers(0, 4)))
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    cov = gb["depth"] < 1.0
    prm = S.default_vct_params(regs[0], cfg.resolution, int(rng.choice([8, 8, 7, 5, 6, 4])))
    prm.enable_32_cones = int(rng.integers(0, 2))
    prm.trace_start_offset = float(rng.choice([0.0, 0.5, 1.0, 2.0]))
    prm.min_trace_step_factor = float(rng.choice([0.1, 0.2, 0.5, 1.0, 2.0]))
    prm.occlusion_decay = float(rng.choice([0.0, 1.0, 2.0, 5.0]))
    prm.indirect_diffuse_intensity = float(rng.choice([1.0, 8.0, 15.0]))
    prm.ambient_occlusion_factor = float(rng.choice([0.25, 0.5, 1.0]))
    args = (cfg, cam, hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
    d0, s0, _ = oracle.cone_trace(*args)
    d1, s1, disc = refshaders.cone_trace(*args)
    assert np.array_equal(disc.astype(bool), ~cov)
    assert np.array_equal(d0[cov], d1[cov], equal_nan=True) and np.array_equal(s0[cov], s1[cov], equal_nan=True)


def test_live_inputs_of_the_gpu_tests(oracle, refshaders):
    """CPU twin of tests/test_gpu_ref_shaders.py: on exactly the inputs those tests feed libvgi, the oracle and the
    reference's shaders agree (bit for bit), so "libvgi == oracle" and "libvgi == reference shader" are one statement."""
    from tests import common
    cfg = common.helper_config()
    regs = oracle.regions(cfg, common.HELPER_CAM)
    for which in (0, 1):
        a = common.random_atlas(cfg, 4 + which)
        x, y = a.copy(), a.copy()
        for level in (1, 2):
            oracle.downsample(cfg, regs, level, x, which)
            refshaders.downsample(cfg, regs, level, y, which)
        assert np.array_equal(x, y) and (x != a).any()
    dst, src = common.random_atlas(cfg, 2), common.random_atlas(cfg, 3)
    for level in range(3):
        x, y = dst.copy(), dst.copy()
        oracle.copy_alpha(cfg, level, x, src)
        refshaders.copy_alpha(cfg, level, y, src)
        assert np.array_equal(x, y)
    a = common.random_atlas(cfg, 1)
    for mc, ext, level in common.HELPER_CLEAR_CASES:
        x, y = a.copy(), a.copy()
        oracle.clear_region(cfg, x, mc, ext, level)
        refshaders.clear_region(cfg, y, mc, ext, level)
        assert np.array_equal(x, y) and (x != a).any()
    for literal in (False, True):
        bcfg = S.default_config(32, 2, mode_flags=S.VGI_MODE_BORDER_LITERAL if literal else 0)
        a = common.random_atlas(bcfg, 9)
        x, y = a.copy(), a.copy()
        oracle.wrap_border(bcfg, x, literal)
        refshaders.wrap_border(bcfg, y, literal)
        assert np.array_equal(x, y)
    for method in (0, 1):
        for tonemap in (0, 1):
            dif, spc = common.filter_images(45, 71, 7 + method)
            prm = S.default_filter_params(method, tonemap)
            assert np.array_equal(oracle.specular_filter(dif, spc, prm), refshaders.specular_filter(dif, spc, prm))
    inp = common.cornell_inputs()
    osc = oracle.OracleScene(inp["scene"])
    cregs = oracle.regions(inp["cfg"], inp["cam_pos"])
    _, rad, _ = oracle.build_clipmap(inp["cfg"], cregs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    gb = inp["gbuffer"]
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    cov = gb["depth"] < 1.0
    for mode in (7, 8):
        prm = S.default_vct_params(cregs[0], inp["cfg"].resolution, mode)
        d0, s0, _ = oracle.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
        d1, s1, disc = refshaders.cone_trace(inp["cfg"], inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad)
        assert np.array_equal(disc.astype(bool), ~cov)
        assert np.array_equal(d0[cov], d1[cov]) and np.array_equal(s0[cov], s1[cov])
    lo, hi = inp["scene"].world_bbox()
    for level in (5, 7):
        frags = oracle.svo_fragments(level, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
        a, b = oracle.svo_build(level, frags), refshaders.svo_build(level, frags)
        assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])


def test_live_q10_cas_running_average_quantified(oracle, refshaders):
    """SURVEY Q10, measured with the reference's own shader: msaaInjectRadiance.frag run over all fragments of a level on ONE
    image (its CAS running average, sequential order) against the exact mean of the very same per-fragment contributions.
    Facts: a texel with one contribution holds it unchanged; the count lives in the alpha byte and wraps modulo 256; with
    several contributions the per-insert truncation (and, past 255, the wrap) moves the result away from the exact mean -
    which is why the canonical mode (oracle and kernels) accumulates exactly and truncates once."""
    from tests.common import atrium_inputs
    inp = atrium_inputs(resolution=64, shadow_size=1024, width=16, height=9, levels=6)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, inp["cam_pos"])
    osc = oracle.OracleScene(inp["scene"])
    level = 2
    fr = oracle.inject_fragments(cfg, regs, level, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    args = (cfg, regs, level, fr["pos"], fr["nrm"], fr["mat"], osc.materials, inp["light"], inp["shadow"], inp["shadow_depth"])
    cnt, coords, vals = refshaders.inject_fragments(*args)
    cas = refshaders.inject_fragments(*args, accumulate=True)
    d_, h_, w_ = cas.shape[:3]
    k = np.arange(6)[None, :] < cnt[:, None]
    c, v = coords[k], vals[k]
    lin = (c[:, 2].astype(np.int64) * h_ + c[:, 1]) * w_ + c[:, 0]
    order = np.argsort(lin, kind="stable")
    lin, v = lin[order], v[order]
    uniq, start, n = np.unique(lin, return_index=True, return_counts=True)
    exact = np.stack([np.floor(np.add.reduceat(((v >> (8 * ch)) & 0xff).astype(np.float64), start) / n) for ch in range(3)], axis=1)
    got = cas.reshape(-1, 4)[uniq]
    assert np.array_equal(got[:, 3], (n % 256).astype(np.uint8))                 # the count byte wraps
    single = n == 1
    assert single.sum() >= 1 and np.array_equal(got[single, :3].astype(np.float64), exact[single])
    lit = exact.max(axis=1) > 0
    dev = np.abs(got[:, :3].astype(np.float64) - exact)[lit]
    assert n.max() > 255 and 0.5 < dev.mean() < 20.0 and dev.max() >= 8            # round 1: mean 3.4 LSB, max 38 on this input


def test_live_q6_radiance_downsample_does_not_compile_as_shipped(refshaders):
    """`lerpFactor` is declared inside the `if` and used after it. The repaired text (build_ref.py R10) builds;
    the shipped text must fail on exactly that identifier."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "..", "oracle", "glsl_shim"))
    import build_ref as B
    sys.path.pop(0)
    if not B.reference_available():
        pytest.skip("needs the reference tree")
    r = B.compile_unit(B.unit_source("radianceDownSample.comp", repair_q6=False), os.devnull)
    assert r.returncode != 0 and "lerpFactor" in r.stderr


def test_refshader_library_is_test_infrastructure_only():
    """Nothing in the product package may reach into oracle/ (the reference-shader library included)."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(__file__)), "vk_voxel_cone_tracing_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".c")):
                with open(os.path.join(root, f), errors="replace") as fh:
                    text = fh.read()
                assert "refshaders" not in text and "glsl_shim" not in text and "oracle/_ref" not in text, f


def _half_ulps(x, ref16):
    return np.abs(x.astype(np.float16).view(np.uint16).astype(np.int32) - ref16.astype(np.int32))


@pytest.mark.parametrize("eye,look,needs", [
    ((0.0, -1.0, 3.5), (0.0, 0.45, -1.0), ("normal_texture", "metallic_roughness_texture", "emissive_texture")),
    ((-2.0, 1.0, 3.0), (0.9, -0.6, -1.0), ("normal_texture", "metallic_roughness_texture", "base_color_texture", "alpha"))])
def test_live_textured_gbuffer(refshaders, eye, look, needs):
    """The G-buffer producer with textured materials (base colour, metallic-roughness with its unclamped branch, emissive
    through SRGBtoLinear, tangent-space normal map, alpha cutoff) against the reference's own gBufferPass.frag: every covered
    pixel's fs_in block (normal, texture coordinate, tangent as the rasteriser interpolates them) goes through the shader with
    the same textures bound. 8-bit attachments identical, binary16 ones within one ulp; the fragments the shader discards are
    exactly the ones the producer's visibility pass dropped."""
    from vk_voxel_cone_tracing_b200 import raster, synth
    scene = synth.textured_cornell(gbuffer_maps=True)
    tex = synth.procedural_textures()
    w, h = 160, 120
    cam = synth.make_camera(eye, look, aspect=w / h)
    gb = raster.gbuffer(scene, cam, w, h, textures=tex)
    mat, nrm, uv, tg = raster.gbuffer_attributes(scene, cam, w, h, textures=tex, full=True)
    cov = mat >= 0
    assert np.array_equal(cov, gb["depth"] < 1.0) and cov.sum() > 10000
    mats = np.ascontiguousarray(scene.materials)
    d, n, s, e, disc = refshaders.gbuffer_fragments_tex(nrm[cov], uv[cov], tg[cov], mat[cov], mats, tex)
    assert not disc.any()
    to8 = lambda x: np.floor(np.clip(x, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)  # noqa: E731
    assert np.array_equal(to8(d), gb["diffuse"][cov]) and np.array_equal(to8(s), gb["specular"][cov])
    de, dn = _half_ulps(e, gb["emission"][cov]), _half_ulps(n, gb["normal"][cov])
    assert de.max() <= 1 and (de == 0).mean() > 0.99
    assert dn.max() <= 1 and (dn == 0).mean() > 0.97
    # every kind of material is in view: normal-mapped, metallic-roughness-mapped, base-colour, emissive texture
    seen = set(np.unique(mat[cov]).tolist())
    for name in needs:
        assert name == "alpha" or any(mats[i][name] > -1 for i in seen), name
    nm = np.isin(mat, [i for i in seen if mats[i]["normal_texture"] > -1])
    flat = scene.materials.copy()
    flat["normal_texture"] = -1
    scene_flat = synth.textured_cornell(gbuffer_maps=True)
    scene_flat.materials[:] = flat
    gb_flat = raster.gbuffer(scene_flat, cam, w, h, textures=tex)
    assert (gb_flat["normal"][nm] != gb["normal"][nm]).any(axis=-1).mean() > 0.5      # the map really bends the normals
    assert np.array_equal(gb_flat["normal"][~nm], gb["normal"][~nm])

    # alpha cutoff: without it the floor covers more pixels; the shader discards exactly the difference
    opaque = synth.textured_cornell(gbuffer_maps=True)
    opaque.materials["alpha_mode"] = 0
    mat0, nrm0, uv0, tg0 = raster.gbuffer_attributes(opaque, cam, w, h, textures=tex, full=True)
    cov0 = mat0 >= 0
    disc0 = refshaders.gbuffer_fragments_tex(nrm0[cov0], uv0[cov0], tg0[cov0], mat0[cov0], mats, tex)[4].astype(bool)
    dropped = np.zeros((h, w), bool)
    dropped[cov0] = disc0
    assert dropped.sum() > (200 if "alpha" in needs else -1)
    assert np.array_equal(mat[~dropped], mat0[~dropped])            # nothing else changed
    assert (mat[dropped] != mat0[dropped]).all() or (gb["depth"][dropped] > raster.gbuffer(opaque, cam, w, h, textures=tex)["depth"][dropped]).all()
