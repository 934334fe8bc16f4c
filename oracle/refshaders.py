"""ctypes wrapper around oracle/_ref/libvgi_refshaders.so — the reference's own GLSL compiled for the CPU by
oracle/glsl_shim/build_ref.py. TEST INFRASTRUCTURE ONLY: it pins the oracle (tests/test_ref_shaders.py,
oracle/glsl_shim/gen_golden.py) and may serve bench.py's --impl reference leg; the product package never imports it.

Atlases are numpy uint8 arrays of shape (D, H, W, 4) in the reference image layout (structs.atlas_shape)."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "glsl_shim"))
import build_ref  # noqa: E402

sys.path.pop(0)

from vk_voxel_cone_tracing_b200 import structs as S  # noqa: E402

_lib = None


def available():
    """True when the library exists or can be built (the reference tree is present)."""
    return os.path.exists(build_ref.OUT_SO) or build_ref.reference_available()


def build(force=False):
    return build_ref.build(force=force)


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("libvgi_refshaders.so is not built and the reference tree is not present")
        _lib = C.CDLL(path)
        _lib.ref_octree_build.restype = C.c_uint32
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _dims(atlas):
    d, h, w, c = atlas.shape
    assert c == 4 and atlas.dtype == np.uint8 and atlas.flags["C_CONTIGUOUS"]
    return C.c_int(w), C.c_int(h), C.c_int(d)


# ---- clipmap compute passes --------------------------------------------------------------------

def downsample(cfg, regs, level, atlas, which):
    """opacityDownSample.comp (which = 0) / radianceDownSample.comp (which = 1) as DownSampler::cmdDownSample runs them."""
    fn = lib().ref_opacity_downsample if which == 0 else lib().ref_radiance_downsample
    prev = (C.c_int * 3)(*regs[level - 1].min_corner)
    fn(_p(atlas), *_dims(atlas), prev, C.c_int(level), C.c_int(cfg.resolution), C.c_int(cfg.downsample_band))


def wrap_border(cfg, atlas, literal=True):
    """borderWrapping.comp. literal: the host's dispatch of (R + 2) >> 3 groups (BorderWrapper.cpp:139, Q4);
    otherwise enough groups to reach the high border too."""
    r = cfg.resolution
    groups = (r + 2) >> 3 if literal else (r + 2 + 7) >> 3
    lib().ref_border_wrap(_p(atlas), *_dims(atlas), C.c_uint(r), C.c_uint(2), C.c_uint(S.VGI_FACES),
                          C.c_uint(cfg.level_count), C.c_uint(groups))


def clear_region(cfg, atlas, min_corner, extent, level):
    """clipmapCleaning.comp as ClipmapCleaner::cmdClearImageClipmapRegion runs it."""
    lib().ref_clipmap_clean(_p(atlas), *_dims(atlas), (C.c_int * 3)(*min_corner), C.c_int(level),
                            (C.c_uint * 3)(*extent), C.c_int(cfg.resolution), C.c_uint(S.VGI_FACES))


def copy_alpha(cfg, level, dst, src):
    """copyAlphaImage.comp as CopyAlpha::cmdImageCopyAlpha runs it."""
    assert dst.shape == src.shape
    lib().ref_copy_alpha(_p(dst), _p(src), *_dims(dst), C.c_int(level), C.c_int(cfg.resolution), C.c_int(S.VGI_FACES))


# ---- octree ---------------------------------------------------------------------------------------

def svo_build(level, frags, capacity=None, subdivision_only=False):
    """The six octreeNode*.comp programs in OctreeBuilder::cmdBuild's order, invocations executed sequentially."""
    frags = np.ascontiguousarray(frags, dtype=np.uint32)
    n = frags.shape[0]
    if capacity is None:
        capacity = min(max(1000000, n << 3), 500000000)     # OctreeBuilder.cpp:109-112
    nodes = np.zeros((capacity, 2), dtype=np.uint32)
    cnt = lib().ref_octree_build(C.c_uint(level), _p(frags), C.c_uint(n), _p(nodes), C.c_uint(capacity),
                                 C.c_int(1 if subdivision_only else 0))
    assert cnt != 0xffffffff, "node pool too small"
    return nodes[:cnt].copy()


# ---- cone tracing ---------------------------------------------------------------------------------

class _TraceArgs(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("y0", C.c_int), ("y1", C.c_int),
        ("view_proj", C.c_void_p), ("view_proj_inv", C.c_void_p), ("eye", C.c_void_p),
        ("diffuse", C.c_void_p), ("normal", C.c_void_p), ("specular", C.c_void_p), ("emission", C.c_void_p),
        ("depth", C.c_void_p),
        ("radiance", C.c_void_p), ("W", C.c_int), ("H", C.c_int), ("D", C.c_int),
        ("shadow_depth", C.c_void_p), ("sw", C.c_int), ("sh", C.c_int),
        ("light_direction", C.c_void_p), ("light_intensity", C.c_float), ("light_color", C.c_void_p),
        ("shadow_view", C.c_void_p), ("shadow_proj", C.c_void_p), ("z_near", C.c_float), ("z_far", C.c_float),
        ("volume_center", C.c_void_p), ("rendering_mode", C.c_uint), ("voxel_size", C.c_float),
        ("volume_dimension", C.c_float), ("trace_start_offset", C.c_float), ("indirect_diffuse_intensity", C.c_float),
        ("ambient_occlusion_factor", C.c_float), ("min_trace_step_factor", C.c_float),
        ("indirect_specular_intensity", C.c_float), ("occlusion_decay", C.c_float), ("enable_32_cones", C.c_int),
        ("clip_level_count", C.c_int),
        ("out_diffuse", C.c_void_p), ("out_specular", C.c_void_p), ("out_discarded", C.c_void_p),
    ]


class _SvoTraceArgs(C.Structure):
    _fields_ = [f for f in _TraceArgs._fields_ if f[0] not in ("radiance", "W", "H", "D")]
    _fields_.insert([f[0] for f in _fields_].index("shadow_depth"), ("nodes", C.c_void_p))


def _decode_gbuffer(gbuf):
    """What the texture unit does to the G-buffer formats (GBufferPass.cpp:177-194): UNORM8 -> c / 255 in binary32,
    binary16 widened exactly."""
    cached = getattr(gbuf, "_decoded_for_refshaders", None)     # decode once per G-buffer, not once per traced row block
    if cached is not None:
        return cached
    to_f = lambda u8: (u8.astype(np.float32) / np.float32(255.0)).astype(np.float32)  # noqa: E731
    nrm = gbuf.normal.view(np.float16) if gbuf.normal.dtype != np.float16 else gbuf.normal
    emi = gbuf.emission.view(np.float16) if gbuf.emission.dtype != np.float16 else gbuf.emission
    h, w = gbuf.height, gbuf.width
    out = (np.ascontiguousarray(to_f(gbuf.diffuse.reshape(h, w, 4))), np.ascontiguousarray(nrm.reshape(h, w, 4).astype(np.float32)),
           np.ascontiguousarray(to_f(gbuf.specular.reshape(h, w, 4))), np.ascontiguousarray(emi.reshape(h, w, 4).astype(np.float32)),
           np.ascontiguousarray(gbuf.depth.reshape(h, w).astype(np.float32)))
    try:
        gbuf._decoded_for_refshaders = out
    except AttributeError:
        pass
    return out


def _fill_common(a, keep, cam, gbuf, prm, light, shadow, shadow_depth, clip_level_count, rows):
    h, w = gbuf.height, gbuf.width
    dif, nrm, spc, emi, dep = _decode_gbuffer(gbuf)
    out_d = np.zeros((h, w, 4), dtype=np.float32)
    out_s = np.zeros((h, w, 4), dtype=np.float32)
    disc = np.zeros((h, w), dtype=np.uint8)
    sd = np.ascontiguousarray(shadow_depth, dtype=np.float32)
    arrs = dict(
        view_proj=np.array(list(cam.view_proj), dtype=np.float32), view_proj_inv=np.array(list(cam.view_proj_inv), dtype=np.float32),
        eye=np.array(list(cam.eye_pos), dtype=np.float32), diffuse=dif, normal=nrm, specular=spc, emission=emi, depth=dep,
        shadow_depth=sd, light_direction=np.array(list(light.direction), dtype=np.float32),
        light_color=np.array(list(light.color), dtype=np.float32), shadow_view=np.array(list(shadow.view), dtype=np.float32),
        shadow_proj=np.array(list(shadow.proj), dtype=np.float32), volume_center=np.array(list(prm.volume_center), dtype=np.float32),
        out_diffuse=out_d, out_specular=out_s, out_discarded=disc)
    keep.append(arrs)
    for k, v in arrs.items():
        setattr(a, k, v.ctypes.data)
    a.width, a.height = w, h
    a.y0, a.y1 = rows if rows is not None else (0, h)
    a.sh, a.sw = sd.shape
    a.light_intensity = light.intensity
    a.z_near, a.z_far = shadow.z_near, shadow.z_far
    for f in ("rendering_mode", "voxel_size", "volume_dimension", "trace_start_offset", "indirect_diffuse_intensity",
              "ambient_occlusion_factor", "min_trace_step_factor", "indirect_specular_intensity", "occlusion_decay",
              "enable_32_cones"):
        setattr(a, f, getattr(prm, f))
    a.clip_level_count = clip_level_count
    return out_d, out_s, disc


def cone_trace(cfg, cam, gbuf, prm, light, shadow, shadow_depth, radiance, rows=None):
    """voxelConeTracing.frag, one invocation per pixel. gbuf: pyoracle.HostGBuffer. Returns (diffuse, specular, discarded)."""
    a, keep = _TraceArgs(), []
    out = _fill_common(a, keep, cam, gbuf, prm, light, shadow, shadow_depth, cfg.level_count, rows)
    rad = np.ascontiguousarray(radiance)
    a.radiance = rad.ctypes.data
    a.D, a.H, a.W = rad.shape[:3]
    lib().ref_cone_trace(C.byref(a))
    return out


def svo_cone_trace(cam, gbuf, prm, light, shadow, shadow_depth, nodes, clip_level_count=6, rows=None):
    """voxelConeTracing_Octree.frag (the scene bounding box is the constant pair inside the shader, Q14)."""
    a, keep = _SvoTraceArgs(), []
    out = _fill_common(a, keep, cam, gbuf, prm, light, shadow, shadow_depth, clip_level_count, rows)
    nd = np.ascontiguousarray(nodes, dtype=np.uint32)
    a.nodes = nd.ctypes.data
    lib().ref_svo_cone_trace(C.byref(a))
    return out


SPONZA_BB_MIN = (-15.367, -1.011, -9.462)   # voxelConeTracing_Octree.frag:321
SPONZA_BB_MAX = (14.399, 11.43, 8.84)       # voxelConeTracing_Octree.frag:322


def specular_filter(diffuse, specular, prm):
    """specularFilter.frag, one invocation per pixel; prm: structs.FilterParams."""
    h, w = diffuse.shape[:2]
    d = np.ascontiguousarray(diffuse, dtype=np.float32)
    s = np.ascontiguousarray(specular, dtype=np.float32)
    out = np.empty((h, w, 4), dtype=np.float32)
    lib().ref_specular_filter(_p(d), _p(s), C.c_int(w), C.c_int(h), C.c_float(prm.tonemap_gamma),
                              C.c_float(prm.tonemap_exposure), C.c_int(prm.tonemap_enable), C.c_int(prm.filter_method),
                              _p(out))
    return out


# ---- radiance injection, per fragment -------------------------------------------------------------

class _InjectArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("position", C.c_void_p), ("normal", C.c_void_p), ("material_index", C.c_void_p),
        ("materials", C.c_void_p),
        ("region_min_corner", C.c_void_p), ("clip_level", C.c_uint), ("region_max_corner", C.c_void_p),
        ("clip_max_extent", C.c_float), ("voxel_size", C.c_float), ("resolution", C.c_int),
        ("shadow_depth", C.c_void_p), ("sw", C.c_int), ("sh", C.c_int),
        ("light_direction", C.c_void_p), ("light_intensity", C.c_float), ("light_color", C.c_void_p),
        ("shadow_view", C.c_void_p), ("shadow_proj", C.c_void_p), ("z_near", C.c_float), ("z_far", C.c_float),
        ("W", C.c_int), ("H", C.c_int), ("D", C.c_int),
        ("out_count", C.c_void_p), ("out_coords", C.c_void_p), ("out_values", C.c_void_p),
        ("texcoord", C.c_void_p), ("n_textures", C.c_int), ("tex_data", C.c_void_p), ("tex_w", C.c_void_p), ("tex_h", C.c_void_p),
    ]


def voxelization_desc(region):
    """VoxelizationDesc as Voxelizer::cmdVoxelize fills it (Voxelizer.cpp:235-256), in binary32."""
    f = np.float32
    vs = f(region.voxel_size)
    mn = np.array([f(c) * vs - f(1e-6) for c in region.min_corner], dtype=np.float32)
    mx = np.array([f(int(c) + int(e)) * vs + f(1e-6) for c, e in zip(region.min_corner, region.extent)], dtype=np.float32)
    return mn, mx, float(f(region.extent[0]) * vs), float(vs)


def inject_fragments(cfg, regs, level, position, normal, material_index, materials, light, shadow, shadow_depth,
                     accumulate=False, texcoord=None, textures=None):
    """msaaInjectRadiance.frag, one invocation per given fragment on a cleared r32ui image.
    Returns (count[n], coords[n, 6, 3] atlas texel, values[n, 6] packed RGBA8 word).
    accumulate=True: all fragments in the given order on ONE image (the reference's CAS running average for that order, Q10);
    returns the radiance atlas (D, H, W, 4) uint8 with the running count in the alpha byte."""
    n = position.shape[0]
    a = _InjectArgs()
    mn, mx, ext, vs = voxelization_desc(regs[level])
    sd = np.ascontiguousarray(shadow_depth, dtype=np.float32)
    keep = dict(position=np.ascontiguousarray(position, np.float32), normal=np.ascontiguousarray(normal, np.float32),
                material_index=np.ascontiguousarray(material_index, np.int32), materials=np.ascontiguousarray(materials),
                region_min_corner=mn, region_max_corner=mx, shadow_depth=sd,
                light_direction=np.array(list(light.direction), np.float32), light_color=np.array(list(light.color), np.float32),
                shadow_view=np.array(list(shadow.view), np.float32), shadow_proj=np.array(list(shadow.proj), np.float32),
                out_count=np.zeros(n, np.int32), out_coords=np.zeros((n, 6, 3), np.int32), out_values=np.zeros((n, 6), np.uint32))
    for k, v in keep.items():
        setattr(a, k, v.ctypes.data)
    if textures:
        # uTextures[]: float RGBA = byte / 255 in binary32 (what a UNORM8 image returns), REPEAT + LINEAR in the driver
        tex_f = [np.ascontiguousarray(t.astype(np.float32) / np.float32(255.0)) for t in textures]
        ptrs = (C.c_void_p * len(tex_f))(*[t.ctypes.data for t in tex_f])
        tw = np.array([t.shape[1] for t in tex_f], np.int32)
        th = np.array([t.shape[0] for t in tex_f], np.int32)
        tc = np.ascontiguousarray(texcoord, np.float32)
        keep.update(_tex_f=tex_f, _ptrs=ptrs, _tw=tw, _th=th, _tc=tc)
        a.texcoord, a.n_textures, a.tex_data = tc.ctypes.data, len(tex_f), C.cast(ptrs, C.c_void_p).value
        a.tex_w, a.tex_h = tw.ctypes.data, th.ctypes.data
    a.n, a.clip_level, a.clip_max_extent, a.voxel_size, a.resolution = n, level, ext, vs, cfg.resolution
    a.sh, a.sw = sd.shape
    a.light_intensity, a.z_near, a.z_far = light.intensity, shadow.z_near, shadow.z_far
    a.D, a.H, a.W = S.atlas_shape(cfg)[:3]
    if accumulate:
        img = np.zeros(S.atlas_shape(cfg)[:3], dtype=np.uint32)
        lib().ref_inject_accumulate(C.byref(a), _p(img))
        return img.view(np.uint8).reshape(S.atlas_shape(cfg))
    lib().ref_inject_fragments(C.byref(a))
    return keep["out_count"], keep["out_coords"], keep["out_values"]


# ---- opacity voxelization stages ------------------------------------------------------------------

def voxelizer_geometry(tri_pos, view_proj=None):
    """msaaVoxelizer.geom per triangle: (axis[n] = gl_ViewportIndex, clip[n, 3, 4] = uViewProj[axis] * position)."""
    t = np.ascontiguousarray(tri_pos, dtype=np.float32).reshape(-1, 3, 3)
    n = t.shape[0]
    vp = np.ascontiguousarray(view_proj, dtype=np.float32) if view_proj is not None else np.tile(np.eye(4, dtype=np.float32).reshape(16), (3, 1))
    axis = np.zeros(n, dtype=np.int32)
    clip = np.zeros((n, 3, 4), dtype=np.float32)
    lib().ref_voxelizer_geometry(C.c_int(n), _p(t), _p(vp), _p(axis), _p(clip))
    return axis, clip


def voxelizer_fragments(cfg, regs, level, position, atlas):
    """msaaVoxelizer.frag, one invocation per given world position, storing into `atlas`. Returns discarded[n]."""
    pos = np.ascontiguousarray(position, dtype=np.float32)
    n = pos.shape[0]
    mn, mx, ext, vs = voxelization_desc(regs[level])
    disc = np.zeros(n, dtype=np.uint8)
    lib().ref_voxelizer_fragments(C.c_int(n), _p(pos), _p(atlas), *_dims(atlas), _p(mn), C.c_uint(level), _p(mx),
                                  C.c_float(ext), C.c_float(vs), C.c_int(cfg.resolution), _p(disc))
    return disc


# ---- SVO fragment shader, per fragment ------------------------------------------------------------

class _SvoFragArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("position", C.c_void_p), ("normal", C.c_void_p), ("material_index", C.c_void_p),
        ("materials", C.c_void_p), ("material_count", C.c_int), ("white_base_color_texture", C.c_int),
        ("voxel_resolution", C.c_uint),
        ("shadow_depth", C.c_void_p), ("sw", C.c_int), ("sh", C.c_int),
        ("light_direction", C.c_void_p), ("light_intensity", C.c_float), ("light_color", C.c_void_p),
        ("shadow_view", C.c_void_p), ("shadow_proj", C.c_void_p), ("z_near", C.c_float), ("z_far", C.c_float),
        ("out_discarded", C.c_void_p), ("out_words", C.c_void_p),
    ]


def svo_fragments(level, position, normal, material_index, materials, light, shadow, shadow_depth, white_base_color_texture=False):
    """voxelizer.frag (uIsCountMode = 0), one invocation per given fragment: position = the biased [0,1] position the
    geometry shader emits. white_base_color_texture binds a 1 x 1 white texture as every material's base-colour texture,
    which turns `color = texture * pbrBaseColorFactor` into the factor (the oracle's canonical colour, Q21).
    Returns (discarded[n], words[n, 2])."""
    n = position.shape[0]
    a = _SvoFragArgs()
    sd = np.ascontiguousarray(shadow_depth, dtype=np.float32)
    keep = dict(position=np.ascontiguousarray(position, np.float32), normal=np.ascontiguousarray(normal, np.float32),
                material_index=np.ascontiguousarray(material_index, np.int32), materials=np.ascontiguousarray(materials),
                shadow_depth=sd, light_direction=np.array(list(light.direction), np.float32),
                light_color=np.array(list(light.color), np.float32), shadow_view=np.array(list(shadow.view), np.float32),
                shadow_proj=np.array(list(shadow.proj), np.float32),
                out_discarded=np.zeros(n, np.uint8), out_words=np.zeros((n, 2), np.uint32))
    for k, v in keep.items():
        setattr(a, k, v.ctypes.data)
    a.n, a.material_count, a.white_base_color_texture = n, keep["materials"].shape[0], 1 if white_base_color_texture else 0
    a.voxel_resolution = 1 << level
    a.sh, a.sw = sd.shape
    a.light_intensity, a.z_near, a.z_far = light.intensity, shadow.z_near, shadow.z_far
    lib().ref_svo_fragments(C.byref(a))
    return keep["out_discarded"], keep["out_words"]


# ---- G-buffer fragment shader, per fragment -------------------------------------------------------

def gbuffer_fragments(normal, material_index, materials):
    """gBufferPass.frag, one invocation per given fragment (interpolated un-normalised normal, material).
    Returns float32 (n, 4) arrays diffuse, normal, specular, emission (the four colour attachments before format
    conversion) and discarded[n]."""
    nrm = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
    n = nrm.shape[0]
    mi = np.ascontiguousarray(material_index, np.int32)
    mats = np.ascontiguousarray(materials)
    outs = [np.zeros((n, 4), np.float32) for _ in range(4)]
    disc = np.zeros(n, np.uint8)
    lib().ref_gbuffer_fragments(C.c_int(n), _p(nrm), _p(mi), _p(mats), *(_p(o) for o in outs), _p(disc))
    return (*outs, disc)


def gbuffer_fragments_tex(normal, texcoord, tangent, material_index, materials, textures):
    """gBufferPass.frag with the whole fs_in block and the scene's textures (list of (H, W, 4) uint8); same returns."""
    nrm = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
    n = nrm.shape[0]
    tc = np.ascontiguousarray(texcoord, np.float32).reshape(-1, 2)
    tg = np.ascontiguousarray(tangent, np.float32).reshape(-1, 4)
    assert tc.shape[0] == n and tg.shape[0] == n
    mi = np.ascontiguousarray(material_index, np.int32)
    mats = np.ascontiguousarray(materials)
    tex_f = [np.ascontiguousarray(t.astype(np.float32) / np.float32(255.0)) for t in textures]
    ptrs = (C.c_void_p * max(len(tex_f), 1))(*[t.ctypes.data for t in tex_f])
    tw = np.array([t.shape[1] for t in tex_f] or [0], np.int32)
    th = np.array([t.shape[0] for t in tex_f] or [0], np.int32)
    outs = [np.zeros((n, 4), np.float32) for _ in range(4)]
    disc = np.zeros(n, np.uint8)
    lib().ref_gbuffer_fragments_tex(C.c_int(n), _p(nrm), _p(tc), _p(tg), _p(mi), _p(mats), C.c_int(len(tex_f)), ptrs, _p(tw), _p(th),
                                    *(_p(o) for o in outs), _p(disc))
    return (*outs, disc)


# ---- vertex stage of the voxelization passes -------------------------------------------------------

def voxelizer_vertices(scene):
    """msaaVoxelizer.vert over every triangle vertex of a synth.Scene in draw order: (pos[n, 3, 3], nrm[n, 3, 3])."""
    pos = np.ascontiguousarray(scene.positions, np.float32)
    nrm = np.ascontiguousarray(scene.normals, np.float32)
    idx = np.ascontiguousarray(scene.indices, np.uint32)
    prims = np.ascontiguousarray(scene.primitives)
    nodes = np.ascontiguousarray(scene.nodes)
    assert prims.dtype.itemsize == 20 and nodes.dtype.itemsize == 128
    fn = lib().ref_voxelizer_vertices
    fn.restype = C.c_uint
    n = int(fn(_p(pos), _p(nrm), _p(idx), _p(prims), C.c_uint(prims.shape[0]), _p(nodes), C.c_void_p(0), C.c_void_p(0)))
    out_p, out_n = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32)
    fn(_p(pos), _p(nrm), _p(idx), _p(prims), C.c_uint(prims.shape[0]), _p(nodes), _p(out_p), _p(out_n))
    return out_p, out_n


def svo_vertex_stage(scene, bb_min, bb_max):
    """voxelizer.vert then voxelizer.geom over every triangle of a synth.Scene in draw order:
    (ndc[n, 3, 3] = gl_Position.xyz of the vertex stage, biased[n, 3, 3] = gs_out.position, axis[n] = gl_ViewportIndex,
    normal[n, 3, 3] = gs_out.normal)."""
    pos = np.ascontiguousarray(scene.positions, np.float32)
    nrm = np.ascontiguousarray(scene.normals, np.float32)
    idx = np.ascontiguousarray(scene.indices, np.uint32)
    prims = np.ascontiguousarray(scene.primitives)
    nodes = np.ascontiguousarray(scene.nodes)
    bb = np.array(list(map(float, bb_min)) + list(map(float, bb_max)), np.float32)
    fn = lib().ref_svo_vertices
    fn.restype = C.c_uint
    n = int(fn(_p(pos), _p(nrm), _p(idx), _p(prims), C.c_uint(prims.shape[0]), _p(nodes), _p(bb), C.c_void_p(0), C.c_void_p(0)))
    ndc, vnrm = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32)
    fn(_p(pos), _p(nrm), _p(idx), _p(prims), C.c_uint(prims.shape[0]), _p(nodes), _p(bb), _p(ndc), _p(vnrm))
    axis = np.zeros(n, np.int32)
    biased, gnrm = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32)
    lib().ref_svo_geometry(C.c_int(n), _p(ndc), _p(vnrm), _p(axis), _p(biased), _p(gnrm))
    return ndc, biased, axis, gnrm
