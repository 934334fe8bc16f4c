"""ctypes wrapper around oracle/_build/libvgi_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package. Pinned against the reference's shader text (see vgi_oracle.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

from vk_voxel_cone_tracing_b200 import structs as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvgi_oracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("vgi_oracle.c", "vgi_oracle_svo.inc", "vgi_oracle.h", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "vgi.h"))
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Tris(C.Structure):
    _fields_ = [("count", C.c_uint32), ("pos", C.c_void_p), ("nrm", C.c_void_p), ("mat", C.c_void_p),
                ("uv", C.c_void_p), ("materials", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.vgo_atlas_bytes.restype = C.c_size_t
        _lib.vgo_scene_triangle_count.restype = C.c_uint32
        _lib.vgo_voxelize_level.restype = C.c_uint64
        _lib.vgo_voxelization_pass.restype = C.c_uint64
        _lib.vgo_last_specular_taps.restype = C.c_uint64
        _lib.vgo_svo_fragments.restype = C.c_uint32
        _lib.vgo_svo_build.restype = C.c_uint32
        _lib.vgo_svo_canonicalize.restype = C.c_uint32
    return _lib


def set_threads(n=0):
    """Use n OpenMP threads (0 = leave unchanged); returns the thread count in effect."""
    return int(lib().vgo_set_threads(C.c_int(int(n))))


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class OracleScene:
    """World-space triangle soup of a synth.Scene (vgo_scene_triangles)."""

    def __init__(self, scene):
        self.scene = scene
        d = scene.desc()
        n = lib().vgo_scene_triangle_count(C.byref(d))
        self.pos = np.zeros((n, 3, 3), dtype=np.float32)
        self.nrm = np.zeros((n, 3, 3), dtype=np.float32)
        self.mat = np.zeros(n, dtype=np.int32)
        lib().vgo_scene_triangles(C.byref(d), _p(self.pos), _p(self.nrm), _p(self.mat))
        self.materials = np.ascontiguousarray(scene.materials)
        # per-triangle texture coordinates in draw order (GLTFScene.cpp:457-490), only for textured scenes
        self.uv = None
        textured = any(int(m[k]) > -1 for m in self.materials for k in ("base_color_texture", "emissive_texture", "occlusion_texture"))
        if textured:
            uv = []
            for p in scene.primitives:
                idx = scene.indices[p["first_index"]:p["first_index"] + p["index_count"]].astype(np.int64) + int(p["vertex_offset"])
                uv.append(scene.texcoords[idx].reshape(-1, 3, 2))
            self.uv = np.ascontiguousarray(np.concatenate(uv, axis=0), dtype=np.float32)
            assert self.uv.shape[0] == n
        self.tris = Tris(n, self.pos.ctypes.data, self.nrm.ctypes.data, self.mat.ctypes.data,
                         self.uv.ctypes.data if textured else None, self.materials.ctypes.data if textured else None)


_tex_keep = None


def set_textures(images):
    """The scene's texture array for the oracle (vgo_set_textures borrows the pointers: kept alive here)."""
    global _tex_keep
    from vk_voxel_cone_tracing_b200 import structs as S2
    arr, keep = S2.texture_array(images or [])
    _tex_keep = (arr, keep)
    lib().vgo_set_textures(arr, C.c_uint32(len(keep)))


def texture_fetch(texture, u, v):
    out = (C.c_float * 4)()
    lib().vgo_texture_fetch(C.c_uint32(texture), C.c_float(u), C.c_float(v), out)
    return np.array(out[:], dtype=np.float32)


def regions(cfg, cam_pos):
    out = (S.ClipRegion * cfg.level_count)()
    cam = (C.c_float * 3)(*[float(x) for x in cam_pos])
    lib().vgo_regions(C.byref(cfg), cam, out)
    return out


def new_atlas(cfg):
    return np.zeros(S.atlas_shape(cfg), dtype=np.uint8)


def voxelization_pass(cfg, regs, osc, opacity):
    return lib().vgo_voxelization_pass(C.byref(cfg), regs, C.byref(osc.tris), _p(opacity))


def voxelize_level(cfg, regs, level, osc, opacity):
    return lib().vgo_voxelize_level(C.byref(cfg), regs, C.c_uint32(level), C.byref(osc.tris), _p(opacity))


def injection_pass(cfg, regs, osc, light, shadow, shadow_depth, frame_index, opacity, radiance):
    h, w = shadow_depth.shape
    lib().vgo_injection_pass(C.byref(cfg), regs, C.byref(osc.tris), _p(osc.materials), C.byref(light),
                             C.byref(shadow), _p(shadow_depth), C.c_uint32(w), C.c_uint32(h),
                             C.c_uint32(frame_index), _p(opacity), _p(radiance))


def inject_level(cfg, regs, level, osc, light, shadow, shadow_depth, radiance):
    h, w = shadow_depth.shape
    lib().vgo_inject_level(C.byref(cfg), regs, C.c_uint32(level), C.byref(osc.tris), _p(osc.materials),
                           C.byref(light), C.byref(shadow), _p(shadow_depth), C.c_uint32(w), C.c_uint32(h),
                           _p(radiance))


def literal_voxelize_level(cfg, regs, level, osc, opacity, samples=8, shade_at=0, q2_fixed=False):
    """Opacity of one level with the MODELLED raster coverage of the reference (vgi_oracle_literal.inc, quirk Q3)."""
    fn = lib().vgo_literal_voxelize_level
    fn.restype = C.c_uint64
    return int(fn(C.byref(cfg), regs, C.c_uint32(level), C.byref(osc.tris), C.c_int(samples), C.c_int(shade_at),
                  C.c_int(1 if q2_fixed else 0), _p(opacity)))


def literal_inject_level(cfg, regs, level, osc, light, shadow, shadow_depth, radiance, samples=8, shade_at=0, q2_fixed=False):
    """Radiance of one level with the modelled raster coverage (exact-mean accumulation, R <= 128)."""
    h, w = shadow_depth.shape
    fn = lib().vgo_literal_inject_level
    fn.restype = C.c_uint64
    return int(fn(C.byref(cfg), regs, C.c_uint32(level), C.byref(osc.tris), _p(osc.materials), C.byref(light), C.byref(shadow),
                  _p(shadow_depth), C.c_uint32(w), C.c_uint32(h), C.c_int(samples), C.c_int(shade_at),
                  C.c_int(1 if q2_fixed else 0), _p(radiance)))


def inject_fragments(cfg, regs, level, osc, light, shadow, shadow_depth):
    """The samples inject_level shades + their shading results (vgo_inject_fragments); dict of numpy arrays."""
    h, w = shadow_depth.shape
    fn = lib().vgo_inject_fragments
    fn.restype = C.c_uint64
    head = (C.byref(cfg), regs, C.c_uint32(level), C.byref(osc.tris), _p(osc.materials), C.byref(light), C.byref(shadow),
            _p(shadow_depth), C.c_uint32(w), C.c_uint32(h))
    nul = C.c_void_p(0)
    n = int(fn(*head, C.c_uint64(0), nul, nul, nul, nul, nul, nul, nul, nul))
    out = dict(pos=np.zeros((n, 3), np.float32), nrm=np.zeros((n, 3), np.float32), mat=np.zeros(n, np.int32),
               voxel=np.zeros((n, 3), np.int32), nfaces=np.zeros(n, np.int32), faces=np.zeros((n, 6), np.int32),
               q=np.zeros((n, 6, 3), np.uint32), uv=np.zeros((n, 2), np.float32))
    fn(*head, C.c_uint64(n), *(_p(out[k]) for k in ("pos", "nrm", "mat", "voxel", "nfaces", "faces", "q", "uv")))
    return out


def dominant_axis(tri_pos):
    """tri_pos: (n, 3, 3) float32 world-space triangles -> (n,) axis (vgo_dominant_axis)."""
    t = np.ascontiguousarray(tri_pos, dtype=np.float32)
    return np.array([lib().vgo_dominant_axis(C.c_void_p(t[i].ctypes.data)) for i in range(t.shape[0])], dtype=np.int32)


class TriangleSoup:
    """A bare world-space triangle list for the voxelization entry points (what OracleScene provides from a scene)."""

    def __init__(self, pos, nrm=None, mat=None):
        self.pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3, 3)
        n = self.pos.shape[0]
        self.nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3, 3) if nrm is not None else np.tile(np.float32([0, 0, 1]), (n, 3, 1))
        self.mat = np.ascontiguousarray(mat, dtype=np.int32) if mat is not None else np.zeros(n, dtype=np.int32)
        self.tris = Tris(n, self.pos.ctypes.data, self.nrm.ctypes.data, self.mat.ctypes.data)


def clear_region(cfg, atlas, min_corner, extent, level):
    lib().vgo_clear_region(C.byref(cfg), _p(atlas), (C.c_int32 * 3)(*min_corner), (C.c_uint32 * 3)(*extent),
                           C.c_uint32(level))


def copy_alpha(cfg, level, dst, src):
    lib().vgo_copy_alpha(C.byref(cfg), C.c_uint32(level), _p(dst), _p(src))


def downsample(cfg, regs, level, atlas, which):
    lib().vgo_downsample(C.byref(cfg), regs, C.c_uint32(level), _p(atlas), C.c_int(which))


def wrap_border(cfg, atlas, literal=False):
    lib().vgo_wrap_border(C.byref(cfg), _p(atlas), C.c_int(1 if literal else 0))


def build_clipmap(cfg, regs, osc, light, shadow, shadow_depth, frame_index=0, opacity=None, radiance=None):
    """VoxelizationPass::render followed by RadianceInjectionPass::render (Application.cpp:178,192)."""
    opacity = new_atlas(cfg) if opacity is None else opacity
    radiance = new_atlas(cfg) if radiance is None else radiance
    pairs = voxelization_pass(cfg, regs, osc, opacity)
    injection_pass(cfg, regs, osc, light, shadow, shadow_depth, frame_index, opacity, radiance)
    return opacity, radiance, pairs


class HostGBuffer:
    def __init__(self, diffuse, normal, specular, emission, depth):
        self.diffuse, self.normal, self.specular, self.emission, self.depth = diffuse, normal, specular, emission, depth
        self.height, self.width = depth.shape

    def struct(self):
        g = S.GBuffer()
        g.diffuse_rgba8 = self.diffuse.ctypes.data
        g.normal_rgba16f = self.normal.ctypes.data
        g.specular_rgba8 = self.specular.ctypes.data
        g.emission_rgba16f = self.emission.ctypes.data
        g.depth_f32 = self.depth.ctypes.data
        g.width, g.height = self.width, self.height
        return g


def cone_trace(cfg, cam, gbuf, prm, light, shadow, shadow_depth, radiance, rows=None):
    h, w = gbuf.height, gbuf.width
    out_d = np.zeros((h, w, 4), dtype=np.float32)
    out_s = np.zeros((h, w, 4), dtype=np.float32)
    taps = C.c_uint64(0)
    y0, y1 = rows if rows is not None else (0, h)
    sh, sw = shadow_depth.shape
    g = gbuf.struct()
    lib().vgo_cone_trace(C.byref(cfg), C.byref(cam), C.byref(g), C.byref(prm), C.byref(light), C.byref(shadow),
                         _p(shadow_depth), C.c_uint32(sw), C.c_uint32(sh), _p(radiance), _p(out_d), _p(out_s),
                         C.c_uint32(y0), C.c_uint32(y1), C.byref(taps))
    return out_d, out_s, taps.value


# ---- SVO ------------------------------------------------------------------------------------------

def last_specular_taps():
    """Tri-linear taps of the specular cones alone in the last cone_trace call."""
    return int(lib().vgo_last_specular_taps())


def specular_filter(diffuse, specular, prm):
    """diffuse / specular: (H, W, 4) float32 numpy; returns the final (H, W, 4) image (vgo_specular_filter)."""
    h, w = diffuse.shape[:2]
    d = np.ascontiguousarray(diffuse, dtype=np.float32)
    s_ = np.ascontiguousarray(specular, dtype=np.float32)
    out = np.empty((h, w, 4), dtype=np.float32)
    lib().vgo_specular_filter(_p(d), _p(s_), C.c_uint32(w), C.c_uint32(h), C.byref(prm), _p(out))
    return out


def svo_fragments(level, bb_min, bb_max, osc, light, shadow, shadow_depth, mode_flags=0):
    sh, sw = shadow_depth.shape
    args = (C.c_uint32(level), (C.c_float * 3)(*map(float, bb_min)), (C.c_float * 3)(*map(float, bb_max)),
            C.byref(osc.tris), _p(osc.materials), C.byref(light), C.byref(shadow), _p(shadow_depth),
            C.c_uint32(sw), C.c_uint32(sh), C.c_uint32(mode_flags))
    n = lib().vgo_svo_fragments(*args, C.c_void_p(0))
    frags = np.zeros((n, 2), dtype=np.uint32)
    lib().vgo_svo_fragments(*args, _p(frags))
    return frags


def svo_vertex_stage(level, bb_min, bb_max, osc):
    """(ndc[n, 3, 3], biased[n, 3, 3], axis[n]) of every triangle (vgo_svo_vertex_stage)."""
    n = osc.pos.shape[0]
    ndc, biased, axis = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32), np.zeros(n, np.int32)
    lib().vgo_svo_vertex_stage(C.c_uint32(level), (C.c_float * 3)(*map(float, bb_min)), (C.c_float * 3)(*map(float, bb_max)),
                               C.byref(osc.tris), _p(ndc), _p(biased), _p(axis))
    return ndc, biased, axis


def svo_fragment_samples(level, bb_min, bb_max, osc):
    """The sample behind every covered (triangle, voxel) of svo_fragments, before the shading's discard; dict of arrays."""
    fn = lib().vgo_svo_fragment_samples
    fn.restype = C.c_uint32
    head = (C.c_uint32(level), (C.c_float * 3)(*map(float, bb_min)), (C.c_float * 3)(*map(float, bb_max)), C.byref(osc.tris))
    nul = C.c_void_p(0)
    n = int(fn(*head, C.c_uint32(0), nul, nul, nul, nul, nul))
    out = dict(world=np.zeros((n, 3), np.float32), biased=np.zeros((n, 3), np.float32), nrm=np.zeros((n, 3), np.float32),
               mat=np.zeros(n, np.int32), voxel=np.zeros((n, 3), np.int32))
    fn(*head, C.c_uint32(n), *(_p(out[k]) for k in ("world", "biased", "nrm", "mat", "voxel")))
    return out


def svo_build(level, frags, capacity=None, mode_flags=0):
    n = frags.shape[0]
    if capacity is None:
        capacity = min(max(1000000, n << 3), 500000000)
    nodes = np.zeros((capacity, 2), dtype=np.uint32)
    cnt = lib().vgo_svo_build(C.c_uint32(level), _p(frags), C.c_uint32(n), _p(nodes), C.c_uint32(capacity),
                              C.c_uint32(mode_flags))
    return nodes[:cnt].copy()


def svo_canonicalize(nodes):
    out = np.zeros_like(nodes)
    n = lib().vgo_svo_canonicalize(_p(nodes), C.c_uint32(nodes.shape[0]), _p(out))
    return out[:n].copy()


def svo_cone_trace(cam, gbuf, prm, light, shadow, shadow_depth, nodes, bb_min, bb_max, clip_level_count=6,
                   rows=None, mode_flags=0):
    h, w = gbuf.height, gbuf.width
    out_d = np.zeros((h, w, 4), dtype=np.float32)
    out_s = np.zeros((h, w, 4), dtype=np.float32)
    y0, y1 = rows if rows is not None else (0, h)
    sh, sw = shadow_depth.shape
    g = gbuf.struct()
    lib().vgo_svo_cone_trace_mode(C.byref(cam), C.byref(g), C.byref(prm), C.byref(light), C.byref(shadow),
                                  _p(shadow_depth), C.c_uint32(sw), C.c_uint32(sh), _p(nodes),
                                  (C.c_float * 3)(*map(float, bb_min)), (C.c_float * 3)(*map(float, bb_max)),
                                  C.c_uint32(clip_level_count), C.c_uint32(mode_flags), _p(out_d), _p(out_s),
                                  C.c_uint32(y0), C.c_uint32(y1))
    return out_d, out_s
