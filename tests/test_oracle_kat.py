"""Analytic known-answer tests pinning the CPU oracle (the reference ships no tests or golden
vectors for this path — SURVEY.md section 4 — so closed-form cases are the pins)."""
import ctypes as C

import numpy as np
import pytest

from vk_voxel_cone_tracing_b200 import structs as S, synth


def _atlas_voxel(cfg, atlas, level, face, x, y, z):
    rb = cfg.resolution + 2
    return atlas[1 + z, 1 + y + level * rb, 1 + x + face * rb]


def test_regions_origin_and_snapped(oracle):
    cfg = S.default_config(128, 6)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    for l, r in enumerate(regs):
        assert list(r.min_corner) == [-64, -64, -64]
        assert list(r.extent) == [128, 128, 128]
        assert r.voxel_size == np.float32(16.0 * 2 ** l / 128)
    # camera moved +1 in x: level 0 voxel 0.125, snap 2 voxels = 0.25 -> 4 snaps = 8 voxels
    regs = oracle.regions(cfg, (1.0, 0.0, 0.0))
    assert list(regs[0].min_corner) == [-56, -64, -64]
    # level 5: voxel 4.0, snap 1 voxel -> trunc(1/4) = 0
    assert list(regs[5].min_corner) == [-64, -64, -64]
    # negative moves truncate toward zero (VoxelizationPass.cpp:438-448 uses a float->int cast)
    regs = oracle.regions(cfg, (-0.3, 0.0, 0.0))
    assert list(regs[0].min_corner) == [-66, -64, -64]  # trunc(-0.3/0.25) = -1 -> -2 voxels


def test_axis_aligned_quad_occupancy_closed_form(oracle):
    """Quad in the plane z=0.3 covering [0.1,1.9]^2: per level l (voxel v=0.25*2^l) it must occupy
    exactly ceil-span voxels in x,y and one z layer."""
    cfg = S.default_config(64, 6)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    scene = synth.quad_scene([(0.1, 0.1, 0.3), (1.9, 0.1, 0.3), (1.9, 1.9, 0.3), (0.1, 1.9, 0.3)])
    osc = oracle.OracleScene(scene)
    op = oracle.new_atlas(cfg)
    pairs = 0
    for l in range(6):
        pairs += oracle.voxelize_level(cfg, regs, l, osc, op)
    rb = cfg.resolution + 2
    for l in range(6):
        v = 0.25 * 2 ** l
        lo, hi = int(np.floor(0.1 / v)), int(np.floor(1.9 / v))
        n = hi - lo + 1
        lvl = op[1:-1, 1 + l * rb:1 + l * rb + 64, 1:65]  # face 0 interior
        occ = lvl[..., 0] == 255
        assert occ.sum() == n * n, (l, occ.sum(), n * n)
        zz, yy, xx = np.nonzero(occ)
        # toroidal texel = voxel mod R
        assert set(zz) == {int(np.floor(0.3 / v)) % 64}
        assert set(xx) == {i % 64 for i in range(lo, hi + 1)}
        # the voxelizer stores (1,1,1,1) into all six faces (msaaVoxelizer.frag:69-73)
        for f in range(6):
            assert np.array_equal(op[1:-1, 1 + l * rb:1 + l * rb + 64, 1 + f * rb:65 + f * rb][occ],
                                  np.full((n * n, 4), 255, np.uint8))
    # the diagonal-shared voxels are hit by both triangles -> more pairs than voxels
    assert pairs >= sum((int(np.floor(1.9 / (0.25 * 2 ** l))) - int(np.floor(0.1 / (0.25 * 2 ** l))) + 1) ** 2 for l in range(6))


@pytest.mark.parametrize("a8", [255, 128, 51])
def test_uniform_opacity_downsample_closed_form(oracle, a8):
    """opacityDownSample.comp:97-138 on a uniform field: every directional pair gives a+(1-a)a, the
    mean of four pairs is a(2-a); outside the blend band the result is round(255*a(2-a))."""
    cfg = S.default_config(32, 2, downsample_band=2)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    at = oracle.new_atlas(cfg)
    rb = cfg.resolution + 2
    at[1:-1, 1:33, :, :] = 0
    for f in range(6):
        at[1:-1, 1:33, 1 + f * rb:33 + f * rb, :] = a8      # level 0 interior, all channels
    oracle.downsample(cfg, regs, 1, at, 0)
    a = np.float32(a8) / np.float32(255)
    want_center = int(np.float32(a * (np.float32(2) - a)) * np.float32(255) + np.float32(0.5))
    # level 0 (min corner -16, voxel 0.5) covers level-1 voxels -8..7; centre = (prevMin>>1) + R/4 = 0
    for f in range(6):
        t = _atlas_voxel(cfg, at, 1, f, 0, 0, 0)
        assert abs(int(t[3]) - want_center) <= 1 and t[1] == t[3], (f, t, want_center)
    # outermost covered voxel (-8): dist = |-8+0.5-0|-0.5 = 7, thr = R/4 - band = 6,
    # lerp = (7 - 6 + 1) / (band + 1) = 2/3 toward the level's own raw flag (0 here)
    t = _atlas_voxel(cfg, at, 1, 0, (-8) % 32, 0, 0)
    f32 = np.float32
    ds = f32(a * (f32(2) - a))
    lf = f32(f32(2) * (f32(1) / f32(3)))
    want_edge = int(f32(ds * (f32(1) - lf) + f32(0) * lf) * f32(255) + f32(0.5))
    assert abs(int(t[3]) - want_edge) <= 1 and t[1] == t[3], (t, want_edge)
    # one voxel further in (-7): dist 6 -> lerp 1/3
    t = _atlas_voxel(cfg, at, 1, 0, (-7) % 32, 0, 0)
    lf = f32(f32(1) * (f32(1) / f32(3)))
    assert abs(int(t[3]) - int(f32(ds * (f32(1) - lf)) * f32(255) + f32(0.5))) <= 1
    # and -6: dist 5 < thr -> no blending
    assert abs(int(_atlas_voxel(cfg, at, 1, 0, (-6) % 32, 0, 0)[3]) - want_center) <= 1


def test_empty_volume_cone_trace_closed_form(oracle):
    """Empty radiance atlas: every cone returns (0,0,0,1) (voxelConeTracing.frag:391), so mode 7
    (VXAO) is AOfactor * (1 + sum_{n.d>=0} n.d) / 16 (Q15 bias included)."""
    cfg = S.default_config(32, 3)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    rad = oracle.new_atlas(cfg)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    depth = np.ones((64, 64), dtype=np.float32)
    cam = synth.make_camera((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), aspect=1.0)
    n = np.array([0.0, 0.0, 1.0], dtype=np.float32)
    W = H = 4
    gb = dict(diffuse=np.full((H, W, 4), 255, np.uint8), specular=np.zeros((H, W, 4), np.uint8),
              normal=np.tile((np.append(n * 0.5 + 0.5, 1.0)).astype(np.float16).view(np.uint16), (H, W, 1)),
              emission=np.zeros((H, W, 4), np.uint16), depth=np.full((H, W), 0.995, np.float32))
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    prm = S.default_vct_params(regs[0], 32, 7)
    d, s, taps = oracle.cone_trace(cfg, cam, hg, prm, light, shadow, depth, rad)
    cones = np.array([
        [0.57735, 0.57735, 0.57735], [0.57735, -0.57735, -0.57735], [-0.57735, 0.57735, -0.57735],
        [-0.57735, -0.57735, 0.57735], [-0.903007, -0.182696, -0.388844], [-0.903007, 0.182696, 0.388844],
        [0.903007, -0.182696, 0.388844], [0.903007, 0.182696, -0.388844], [-0.388844, -0.903007, -0.182696],
        [0.388844, -0.903007, 0.182696], [0.388844, 0.903007, -0.182696], [-0.388844, 0.903007, 0.182696],
        [-0.182696, -0.388844, -0.903007], [0.182696, 0.388844, -0.903007], [-0.182696, 0.388844, 0.903007],
        [0.182696, -0.388844, 0.903007]])  # voxelConeTracing.frag:118-135
    cos = cones @ n.astype(np.float64)
    want = 0.5 * (1.0 + cos[cos >= 0].sum()) / 16.0
    assert np.allclose(d[..., 0], want, atol=2e-6) and np.allclose(d[..., 3], 1.0)
    assert np.allclose(d[..., 0], d[..., 1]) and np.allclose(d[..., 0], d[..., 2])
    assert taps > 0


def test_single_fragment_svo_path(oracle):
    """One fragment -> exactly one flagged node per level, 8*level nodes, predicted child slots
    (octreeNodeFlag.comp:28-46: position halved, slot = z | x<<1 | y<<2)."""
    level = 5
    px, py, pz = 21, 6, 27
    frag = np.array([[px | (py << 12) | ((pz & 0xff) << 24), ((pz >> 8) << 28) | 0x0a0b0c0d & 0x0fffffff]], dtype=np.uint32)
    nodes = oracle.svo_build(level, frag)
    assert nodes.shape[0] == 8 * level
    flagged = np.nonzero(nodes[:, 0] & 0x80000000)[0]
    assert flagged.shape[0] == level
    # walk the tree with the shader's descent
    res = 1 << level
    pos = np.array([px, py, pz]) >> 1
    idx = 0
    node_pos = np.zeros(3, dtype=np.int64)
    for depth in range(level):
        res >>= 1
        cmp = (pos >= node_pos + res).astype(np.int64)
        slot = int(cmp[2] | (cmp[0] << 1) | (cmp[1] << 2))
        node_pos += cmp * res
        n = idx + slot
        assert nodes[n, 0] & 0x80000000, (depth, n)
        if depth + 1 < level:
            idx = int(nodes[n, 0] & 0x7fffffff)
            assert idx == 8 * (depth + 1)  # children are allocated in level order
    # leaf colour = the fragment's colour with alpha forced to 255 (canonical mean of one sample)
    assert nodes[n, 1] == ((0x0a0b0c0d & 0x00ffffff) | 0xff000000)
    # canonical ordering of an already breadth-first tree is the identity on the topology word
    canon = oracle.svo_canonicalize(nodes)
    assert np.array_equal(canon[:, 0], nodes[:, 0])


def test_svo_canonicalize_is_allocation_order_independent(oracle):
    rng = np.random.RandomState(3)
    level = 6
    p = rng.randint(0, 1 << level, size=(400, 3)).astype(np.uint32)
    col = rng.randint(0, 1 << 24, size=400).astype(np.uint32)
    frags = np.stack([p[:, 0] | (p[:, 1] << 12) | ((p[:, 2] & 0xff) << 24), ((p[:, 2] >> 8) << 28) | col], axis=1).astype(np.uint32)
    a = oracle.svo_build(level, frags)
    b = oracle.svo_build(level, frags[::-1].copy())
    assert a.shape == b.shape
    ca, cb = oracle.svo_canonicalize(a), oracle.svo_canonicalize(b)
    assert np.array_equal(ca, cb)


def test_injection_single_lit_quad(oracle):
    """Upward-facing quad fully lit (shadow map cleared to 1.0 -> literal Q1 visibility = 1):
    radiance = clamp(N.L,0.001,1) * albedo * |n_axis| into the faces selected by -n; for n = +y that is
    face 3 (-Y travel) with weight 1 and faces 0/1, 4/5 with weight 0."""
    cfg = S.default_config(32, 2)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    scene = synth.quad_scene([(0.3, 0.3, 0.3), (0.3, 0.3, 1.7), (1.7, 0.3, 1.7), (1.7, 0.3, 0.3)],
                             normal=(0, 1, 0), base=(0.5, 0.25, 1.0, 1.0))
    osc = oracle.OracleScene(scene)
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5), direction=(0.0, -1.0, 0.2))
    depth = np.ones((256, 256), dtype=np.float32)
    op, rad, pairs = oracle.build_clipmap(cfg, regs, osc, light, shadow, depth, 0)
    # level 0 voxel 0.5: quad at y=0.3 -> y index 0; x,z indices 0..3
    t = _atlas_voxel(cfg, rad, 0, 3, 1, 0, 1)
    f32 = np.float32
    ndl = f32(1.0) / f32(np.sqrt(f32(1.04)))               # n.L for L = normalize(0, 1, -0.2)
    want = []
    for base in (0.5, 0.25, 1.0):
        q = int(f32(f32(ndl * f32(base)) * f32(65536.0)) + f32(0.5))   # 16.16 fixed point contribution
        want.append((q * 255) >> 16)                                # canonical mean of identical samples
    got = tuple(int(c) for c in t)
    assert all(abs(g - w) <= 1 for g, w in zip(got[:3], want)) and got[3] == 255, (got, want)
    assert abs(got[0] - 125) <= 1 and abs(got[2] - 250) <= 1
    assert tuple(int(c) for c in _atlas_voxel(cfg, rad, 0, 2, 1, 0, 1))[:3] == (0, 0, 0)
    t0 = _atlas_voxel(cfg, rad, 0, 0, 1, 0, 1)
    assert tuple(int(c) for c in t0)[:3] == (0, 0, 0)
    assert tuple(int(c) for c in _atlas_voxel(cfg, op, 0, 0, 1, 0, 1)) == (255, 255, 255, 255)


def test_specular_filter_known_answers(oracle):
    """Closed forms of specularFilter.frag on constant images: the gaussian sums 257 taps and divides by
    8 * 32 - 15 = 241 (filter.glsl:11-23), the bilateral filter of a constant is the constant, and the
    Uncharted-2 curve maps its white point 11.2 / exposure to 1 (tonemapping.glsl:21-26)."""
    from vk_voxel_cone_tracing_b200 import structs as S
    h, w = 9, 12
    dif = np.zeros((h, w, 4), np.float32); dif[..., 3] = 1.0
    spc = np.full((h, w, 4), 0.5, np.float32)
    out = oracle.specular_filter(dif, spc, S.default_filter_params(1, 0))
    np.testing.assert_allclose(out[..., :3], 0.5 * 257.0 / 241.0, rtol=2e-6)
    np.testing.assert_allclose(out[..., 3], 1.0)
    out = oracle.specular_filter(dif, spc, S.default_filter_params(0, 0))
    np.testing.assert_allclose(out[..., :3], 0.5, rtol=2e-6)
    dif[..., :3] = 11.2 / 0.1
    out = oracle.specular_filter(dif, np.zeros_like(spc), S.default_filter_params(0, 1))
    np.testing.assert_allclose(out[..., :3], 1.0, rtol=1e-5)
    # a single bright texel: the gaussian spreads < 1 % of it to each neighbour (blurSize 0.01 is in texels / size)
    spc[:] = 0.0; spc[4, 6, :3] = 1.0; dif[..., :3] = 0.0
    out = oracle.specular_filter(dif, spc, S.default_filter_params(1, 0))
    assert abs(out[4, 6, 0] - 257.0 / 241.0) < 0.02 and 0.0 < out[4, 7, 0] < 0.01 and out[4, 9, 0] == 0.0


def _trace_cone_uniform(c, start_pos, direction, aperture, start_level, step_factor, prm, levels):
    """voxelConeTracing.frag:341-392 written out for a volume whose every texel is `c`: the three face taps
    return c each, so the sample is c * (d.x^2 + d.y^2 + d.z^2) at every position and level. Plain float64
    arithmetic, transcribed from the shader text independently of the C oracle."""
    vs0, dim = prm.voxel_size, prm.volume_dimension
    centre = np.array(prm.volume_center[:], dtype=np.float64)
    w = direction * direction
    sample = c * w[0] + c * w[1] + c * w[2]
    cone_coeff = 2.0 * np.tan(aperture * 0.5)
    voxel = vs0 * 2.0 ** start_level
    pos0 = start_pos + direction * voxel * prm.trace_start_offset * 0.5
    step, diameter, occlusion = 0.0, max(0.0, vs0), 0.0
    result = np.zeros(4)
    seg = voxel
    min_radius = vs0 * dim * 0.5
    while step < 30.0 and occlusion < 1.0:
        p = pos0 + direction * step
        dist = np.linalg.norm(centre - p)
        min_level = np.ceil(np.log2(dist / min_radius))
        cur = np.log2(diameter / vs0)
        cur = min(max(max(start_level, cur), min_level), levels - 1)
        voxel = vs0 * 2.0 ** cur
        corr = seg / voxel
        rad = sample[:3] * corr
        op = min(max(1.0 - (1.0 - sample[3]) ** corr, 0.0), 1.0)
        result += min(max(1.0 - result[3], 0.0), 1.0) * np.array([rad[0], rad[1], rad[2], op])
        occlusion += (1.0 - occlusion) * op / (1.0 + (step + voxel) * prm.occlusion_decay)
        prev = step
        step += max(diameter, vs0) * step_factor
        seg = step - prev
        diameter = step * cone_coeff
    return np.array([result[0], result[1], result[2], 1.0 - occlusion])


def test_uniform_volume_cone_accumulation_closed_form(oracle):
    """A radiance atlas filled with one RGBA value: position and level drop out of every texture fetch, which
    leaves the marching schedule (step sequence, level selection, segment correction, front-to-back blend, occlusion
    decay), the 16-cone gather and the mode switch. Those are transcribed here from the shader text in float64 and
    compared with the oracle's image for modes 7 (VXAO), 5 (emissive pixel: emission * AO + indirect) and 6 (specular
    cone with stepFactor = uVoxelSize, Q12)."""
    cfg = S.default_config(32, 3)
    regs = oracle.regions(cfg, (0.0, 0.0, 0.0))
    rad = oracle.new_atlas(cfg)
    texel = np.array([51, 102, 153, 26], dtype=np.uint8)
    rad[...] = texel
    c = texel.astype(np.float64) / 255.0
    light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
    sdepth = np.ones((64, 64), dtype=np.float32)
    cam = synth.make_camera((0.0, 0.0, 0.0), (0.0, 0.0, -1.0), aspect=1.0)
    n = np.array([0.0, 0.6, 0.8])
    W = H = 2
    e = 0.25
    gb = dict(diffuse=np.tile(np.array([255, 255, 255, 51], np.uint8), (H, W, 1)),       # roughness 0.2
              specular=np.tile(np.array([255, 128, 64, 255], np.uint8), (H, W, 1)),      # F0 colour, metallic 1
              normal=np.tile((np.append(n * 0.5 + 0.5, 1.0)).astype(np.float16).view(np.uint16), (H, W, 1)),
              emission=np.tile(np.array([e, e, e, 1.0], np.float16).view(np.uint16), (H, W, 1)),
              depth=np.full((H, W), 0.995, np.float32))
    hg = oracle.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    cones = np.array([
        [0.57735, 0.57735, 0.57735], [0.57735, -0.57735, -0.57735], [-0.57735, 0.57735, -0.57735],
        [-0.57735, -0.57735, 0.57735], [-0.903007, -0.182696, -0.388844], [-0.903007, 0.182696, 0.388844],
        [0.903007, -0.182696, 0.388844], [0.903007, 0.182696, -0.388844], [-0.388844, -0.903007, -0.182696],
        [0.388844, -0.903007, 0.182696], [0.388844, 0.903007, -0.182696], [-0.388844, 0.903007, 0.182696],
        [-0.182696, -0.388844, -0.903007], [0.182696, 0.388844, -0.903007], [-0.182696, 0.388844, 0.903007],
        [0.182696, -0.388844, 0.903007]])
    vpi = np.array(cam.view_proj_inv[:], dtype=np.float64).reshape(4, 4).T       # column-major float[16]
    eye = np.array(cam.eye_pos[:], dtype=np.float64)
    for mode in (7, 5, 6):
        prm = S.default_vct_params(regs[0], 32, mode)
        d, s, _ = oracle.cone_trace(cfg, cam, hg, prm, light, shadow, sdepth, rad)
        for (py, px) in ((0, 0), (1, 1)):
            tc = np.array([(px + 0.5) / W, (py + 0.5) / H])
            clip = vpi @ np.array([tc[0] * 2 - 1, tc[1] * 2 - 1, 0.995, 1.0])       # no y flip (Q18)
            world = clip[:3] / clip[3]
            nn = np.array([np.float16(v * 0.5 + 0.5) for v in n], dtype=np.float64) * 2.0 - 1.0
            nn /= np.linalg.norm(nn)
            # calcMinLevel, voxelConeTracing.frag:394-414
            centre = np.array(prm.volume_center[:], dtype=np.float64)
            dist = np.linalg.norm(centre - world)
            min_radius = prm.voxel_size * prm.volume_dimension * 0.5
            ml = max(np.log2(dist / min_radius), 0.0)
            f = dist / (min_radius * 2.0 ** np.ceil(ml))
            min_level = np.ceil(ml) + ((f - 0.5) * 2.0 if f > 0.5 else 0.0)
            start = world + nn * (prm.voxel_size * 2.0 ** min_level) * prm.trace_start_offset
            ind = np.array([0.0, 0.0, 0.0, 1.0])                                   # Q15
            for cd in cones:
                cos = float(nn @ cd)
                if cos < 0.0:
                    continue
                ind += _trace_cone_uniform(c, start, cd, 0.872665, min_level, max(0.2, prm.min_trace_step_factor), prm, 3) * cos
            ind /= 16.0                                                            # Q11
            ind[3] *= prm.ambient_occlusion_factor
            ind[:3] *= (255.0 / 255.0) * prm.indirect_diffuse_intensity
            if mode == 7:
                want = np.array([ind[3]] * 3 + [1.0])
                got = d[py, px]
            elif mode == 5:
                emis = float(np.float16(e))
                want = np.append(emis * ind[3] + ind[:3], 1.0)
                got = d[py, px]
            else:
                view = eye - world
                view /= np.linalg.norm(view)
                I = -view
                refl = I - 2.0 * (nn @ I) * nn
                spec = _trace_cone_uniform(c, start, refl, max(51.0 / 255.0, 0.05), min_level, prm.voxel_size, prm, 3)
                want = np.append(spec[:3] * (np.array([255, 128, 64]) / 255.0) * prm.indirect_specular_intensity, 1.0)
                got = s[py, px]
            assert np.allclose(got, want, rtol=2e-4, atol=2e-5), (mode, py, px, got, want)
