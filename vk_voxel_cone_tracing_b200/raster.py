"""Synthetic producers of the hot path's image inputs (shadow depth, G-buffer) — host-side C
(csrc/synth_raster.c). Not on the hot path and never timed; see SURVEY.md section 8d "G-buffer for cone
tracing"."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIBSYNTH):
            _build.build_synth()
        _lib = C.CDLL(_build.LIBSYNTH)
    return _lib


def shadow_depth(scene, shadow_desc, size=4096):
    """D32 shadow map of `scene` from the directional light (ref: ReflectiveShadowMapPass, 4096^2)."""
    depth = np.empty((size, size), dtype=np.float32)
    d = scene.desc()
    lib().vgs_shadow_depth(C.byref(d), C.byref(shadow_desc), C.c_uint32(size), C.c_uint32(size),
                           C.c_void_p(depth.ctypes.data))
    return depth


def gbuffer(scene, camera, width, height, textures=None):
    """G-buffer in the reference formats: diffuse RGBA8, normal RGBA16F, specular RGBA8,
    emission RGBA16F, depth D32 (ref: GBufferPass.cpp:177-194). textures: list of (H, W, 4) uint8 images the materials'
    texture indices refer to (gBufferPass.frag: base colour, metallic-roughness, emissive, normal map, alpha cutoff)."""
    from .structs import texture_array
    diffuse = np.empty((height, width, 4), dtype=np.uint8)
    specular = np.empty((height, width, 4), dtype=np.uint8)
    normal = np.empty((height, width, 4), dtype=np.uint16)
    emission = np.empty((height, width, 4), dtype=np.uint16)
    depth = np.empty((height, width), dtype=np.float32)
    d = scene.desc()
    arr, keep = texture_array(textures or [])
    lib().vgs_gbuffer_tex(C.byref(d), arr, C.c_uint32(len(keep)), C.byref(camera), C.c_uint32(width), C.c_uint32(height),
                          C.c_void_p(diffuse.ctypes.data), C.c_void_p(normal.ctypes.data),
                          C.c_void_p(specular.ctypes.data), C.c_void_p(emission.ctypes.data),
                          C.c_void_p(depth.ctypes.data))
    return dict(diffuse=diffuse, normal=normal, specular=specular, emission=emission, depth=depth)


def gbuffer_attributes(scene, camera, width, height, textures=None, full=False):
    """Test aid: per pixel the material index (-1 = not covered) and the interpolated, un-normalised world normal the
    G-buffer fragment stage receives (vgs_gbuffer_attributes_tex); full=True adds the interpolated texture coordinate and
    tangent (the whole `fs_in` block of gBufferPass.frag)."""
    from .structs import texture_array
    material = np.empty((height, width), dtype=np.int32)
    normal = np.empty((height, width, 3), dtype=np.float32)
    uv = np.empty((height, width, 2), dtype=np.float32)
    tangent = np.empty((height, width, 4), dtype=np.float32)
    d = scene.desc()
    arr, keep = texture_array(textures or [])
    lib().vgs_gbuffer_attributes_tex(C.byref(d), arr, C.c_uint32(len(keep)), C.byref(camera), C.c_uint32(width),
                                     C.c_uint32(height), C.c_void_p(material.ctypes.data), C.c_void_p(normal.ctypes.data),
                                     C.c_void_p(uv.ctypes.data), C.c_void_p(tangent.ctypes.data))
    return (material, normal, uv, tangent) if full else (material, normal)
