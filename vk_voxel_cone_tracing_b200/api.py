"""ctypes binding of libvgi.so (include/vgi.h) — the call a Python user makes.

torch is used for device memory and streams only (plumbing); all compute is in the hand-written
CUDA kernels behind the C ABI. There is no CPU fallback: importing the library without the built
extension, or creating a context without a CUDA device, raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from . import structs as S

_lib = None


class VgiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvgi error {code}: {msg}")
        self.code = code


def lib():
    """Load csrc/libvgi.so (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIBVGI):
            raise ImportError(
                f"{_build.LIBVGI} is missing: build it with `python -m vk_voxel_cone_tracing_b200.build` "
                "(libvgi has no CPU fallback)")
        # VGI_LIBVGI_PATH: development override (e.g. the instrumented build made by tools/trace_stats.py)
        _lib = C.CDLL(os.environ.get("VGI_LIBVGI_PATH", _build.LIBVGI))
        _lib.vgi_last_error.restype = C.c_char_p
        _lib.vgi_last_error.argtypes = [C.c_void_p]
        _lib.vgi_atlas_bytes.restype = C.c_size_t
        _lib.vgi_atlas_bytes.argtypes = [C.c_void_p]
    return _lib


class _DevView:
    """Borrowed view of library-owned device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _stream(stream=None):
    if stream is not None:
        return C.c_void_p(int(stream))
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class VoxelGI:
    """One vgi_ctx: the voxel-GI hot path on one GPU."""

    def __init__(self, cfg=None, device=None):
        import torch
        self._torch = torch
        self.cfg = cfg if cfg is not None else S.default_config()
        if device is not None:
            self.cfg.device = int(device)
        self._h = C.c_void_p()
        self._keep = {}
        rc = lib().vgi_create(C.byref(self.cfg), C.byref(self._h))
        if rc != S.VGI_OK:
            raise VgiError(rc, lib().vgi_last_error(None).decode())
        self.device = torch.device("cuda", self.cfg.device if self.cfg.device >= 0 else torch.cuda.current_device())

    # -- plumbing
    def _ck(self, rc):
        if rc != S.VGI_OK:
            raise VgiError(rc, lib().vgi_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().vgi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs
    def set_scene(self, scene):
        d = scene.desc()
        self._ck(lib().vgi_set_scene(self._h, C.byref(d)))
        self.scene = scene

    def update_nodes(self, nodes, stream=None):
        """New node matrices (structured array like Scene.nodes, same count): the scene is re-transformed on the device."""
        arr = np.ascontiguousarray(nodes)
        self._ck(lib().vgi_update_nodes(self._h, C.c_void_p(arr.ctypes.data), C.c_uint32(arr.shape[0]), _stream(stream)))

    def set_textures(self, images):
        """images: list of (H, W, 4) uint8 arrays (the scene's uTextures[]); copied by the library."""
        arr, keep = S.texture_array(images)
        self._ck(lib().vgi_set_textures(self._h, arr, C.c_uint32(len(keep))))

    def set_light(self, light, shadow, shadow_depth):
        """shadow_depth: numpy (H,W) float32 (copied) or a CUDA torch tensor (borrowed)."""
        torch = self._torch
        if isinstance(shadow_depth, np.ndarray):
            shadow_depth = torch.from_numpy(np.ascontiguousarray(shadow_depth, dtype=np.float32)).to(self.device)
        assert shadow_depth.is_cuda and shadow_depth.dtype == torch.float32 and shadow_depth.is_contiguous()
        self._keep["shadow"] = shadow_depth
        h, w = shadow_depth.shape
        self._ck(lib().vgi_set_light(self._h, C.byref(light), C.byref(shadow), C.c_void_p(shadow_depth.data_ptr()),
                                     C.c_uint32(w), C.c_uint32(h), C.c_int(0)))
        self.light, self.shadow = light, shadow

    def update_regions(self, camera_pos):
        self._ck(lib().vgi_update_regions(self._h, _f3(camera_pos)))

    def set_regions(self, regions):
        self._ck(lib().vgi_set_regions(self._h, regions, C.c_uint32(len(regions))))

    def regions(self):
        out = (S.ClipRegion * self.cfg.level_count)()
        self._ck(lib().vgi_get_regions(self._h, out, C.c_uint32(self.cfg.level_count)))
        return out

    # -- clipmap build
    def voxelize_opacity(self, stream=None):
        self._ck(lib().vgi_voxelize_opacity(self._h, _stream(stream)))

    def inject_radiance(self, frame_index=0, stream=None):
        self._ck(lib().vgi_inject_radiance(self._h, C.c_uint32(frame_index), _stream(stream)))

    def build_clipmap(self, frame_index=0, stream=None):
        self._ck(lib().vgi_build_clipmap(self._h, C.c_uint32(frame_index), _stream(stream)))

    def build_clipmap_incremental(self, frame_index=0, stream=None):
        """vgi_build_clipmap_incremental: rebuild only the clip levels a moved region (or a due cadence frame) invalidates.
        Returns the finest level that was rebuilt (level_count: nothing had to be)."""
        first = C.c_uint32(0)
        self._ck(lib().vgi_build_clipmap_incremental(self._h, C.c_uint32(frame_index), C.byref(first), _stream(stream)))
        return int(first.value)

    def export_atlas(self, which, out=None, stream=None):
        """which: 0 opacity, 1 radiance -> uint8 tensor (D, H, W, 4) in the reference image layout."""
        torch = self._torch
        shape = S.atlas_shape(self.cfg)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=self.device)
        self._ck(lib().vgi_export_atlas(self._h, C.c_int(which), C.c_void_p(out.data_ptr()), _stream(stream)))
        return out

    def voxel_store(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(lib().vgi_get_voxel_store(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def bind_voxel_store(self, tensor):
        self._keep["store"] = tensor
        self._ck(lib().vgi_bind_voxel_store(self._h, C.c_void_p(tensor.data_ptr()), C.c_size_t(tensor.numel() * tensor.element_size())))

    def set_slab(self, z0, z1):
        self._ck(lib().vgi_set_slab(self._h, C.c_uint32(z0), C.c_uint32(z1)))

    def stats(self):
        st = S.Stats()
        self._ck(lib().vgi_get_stats(self._h, C.byref(st)))
        return st

    # -- cone tracing
    def default_vct_params(self, rendering_mode=8):
        p = S.VctParams()
        self._ck(lib().vgi_default_vct_params(self._h, C.byref(p)))
        p.rendering_mode = rendering_mode
        return p

    def set_trace_overlap(self, spec_blocks_per_sm):
        """0 = diffuse march, then specular march; n > 0 = the specular march beside the diffuse one, n blocks per SM."""
        self._ck(lib().vgi_set_trace_overlap(self._h, C.c_uint32(spec_blocks_per_sm)))

    def upload_gbuffer(self, gb):
        """dict of numpy arrays (raster.gbuffer) -> dict of CUDA tensors."""
        torch = self._torch
        out = {}
        for k, v in gb.items():
            a = np.ascontiguousarray(v)
            if a.dtype == np.uint16:
                t = torch.from_numpy(a.view(np.int16)).to(self.device)
            else:
                t = torch.from_numpy(a).to(self.device)
            out[k] = t
        return out

    @staticmethod
    def gbuffer_struct(gb):
        g = S.GBuffer()
        g.diffuse_rgba8 = gb["diffuse"].data_ptr()
        g.normal_rgba16f = gb["normal"].data_ptr()
        g.specular_rgba8 = gb["specular"].data_ptr()
        g.emission_rgba16f = gb["emission"].data_ptr()
        g.depth_f32 = gb["depth"].data_ptr()
        g.height, g.width = gb["depth"].shape
        return g

    def cone_trace(self, camera, gbuffer, params, out=None, rows=None, stream=None, part=None):
        """gbuffer: dict of CUDA tensors. Returns (diffuse, specular) float32 (H, W, 4) CUDA tensors.
        rows=(y0, y1): only that row block; part=(i, n): only the 8-row tile rows i, i+n, ... (multi-GPU)."""
        torch = self._torch
        g = self.gbuffer_struct(gbuffer)
        if out is None:
            out = (torch.zeros((g.height, g.width, 4), dtype=torch.float32, device=self.device),
                   torch.zeros((g.height, g.width, 4), dtype=torch.float32, device=self.device))
        if part is not None:
            self._ck(lib().vgi_cone_trace_interleaved(self._h, C.byref(camera), C.byref(g), C.byref(params),
                                                     C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                                     C.c_uint32(part[0]), C.c_uint32(part[1]), _stream(stream)))
            return out
        y0, y1 = rows if rows is not None else (0, g.height)
        self._ck(lib().vgi_cone_trace_rows(self._h, C.byref(camera), C.byref(g), C.byref(params),
                                          C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                          C.c_uint32(y0), C.c_uint32(y1), _stream(stream)))
        return out

    # -- sparse voxel octree
    def svo_voxelize(self, level, bb_min, bb_max, stream=None):
        self._ck(lib().vgi_svo_voxelize(self._h, C.c_uint32(level), _f3(bb_min), _f3(bb_max), _stream(stream)))
        self.svo_level = level

    def svo_build(self, stream=None):
        self._ck(lib().vgi_svo_build(self._h, _stream(stream)))

    def _svo_view(self, getter):
        torch = self._torch
        p, n = C.c_void_p(), C.c_uint32()
        self._ck(getter(self._h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return torch.zeros((0, 2), dtype=torch.int32, device=self.device)
        # library-owned device memory, copied so that the result survives the next build
        return torch.as_tensor(_DevView(p.value, (n.value, 2), "<i4"), device=self.device).clone()

    def svo_fragments(self):
        """(N, 2) int32 CUDA tensor of packed uvec2 fragments (voxelizer.frag:99-100)."""
        return self._svo_view(lib().vgi_svo_get_fragments)

    def svo_nodes(self):
        """(N, 2) int32 CUDA tensor of uvec2 nodes {flag | child index, RGBA8}."""
        return self._svo_view(lib().vgi_svo_get_nodes)

    def svo_cone_trace(self, camera, gbuffer, params, out=None, stream=None):
        torch = self._torch
        g = self.gbuffer_struct(gbuffer)
        if out is None:
            out = (torch.zeros((g.height, g.width, 4), dtype=torch.float32, device=self.device),
                   torch.zeros((g.height, g.width, 4), dtype=torch.float32, device=self.device))
        self._ck(lib().vgi_svo_cone_trace(self._h, C.byref(camera), C.byref(g), C.byref(params),
                                         C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), _stream(stream)))
        return out

    def render_shadow_map(self, shadow, size=4096, out=None, stream=None):
        """D32 shadow depth of the ctx's scene (vgi_render_shadow_map); returns a (size, size) float32 CUDA tensor."""
        torch = self._torch
        if out is None:
            out = torch.empty((size, size), dtype=torch.float32, device=self.device)
        self._ck(lib().vgi_render_shadow_map(self._h, C.byref(shadow), C.c_uint32(out.shape[1]), C.c_uint32(out.shape[0]),
                                            C.c_void_p(out.data_ptr()), _stream(stream)))
        return out

    def render_gbuffer(self, camera, width, height, out=None, stream=None):
        """G-buffer of the ctx's scene in the reference formats (vgi_render_gbuffer); returns a dict of CUDA tensors
        that cone_trace accepts."""
        torch = self._torch
        if out is None:
            out = dict(diffuse=torch.empty((height, width, 4), dtype=torch.uint8, device=self.device),
                       normal=torch.empty((height, width, 4), dtype=torch.int16, device=self.device),
                       specular=torch.empty((height, width, 4), dtype=torch.uint8, device=self.device),
                       emission=torch.empty((height, width, 4), dtype=torch.int16, device=self.device),
                       depth=torch.empty((height, width), dtype=torch.float32, device=self.device))
        g = self.gbuffer_struct(out)
        self._ck(lib().vgi_render_gbuffer(self._h, C.byref(camera), C.byref(g), _stream(stream)))
        return out

    def specular_filter(self, diffuse, specular, params=None, out=None, stream=None):
        """final = diffuse + filtered specular (+ tonemap): the pass after cone tracing (vgi_specular_filter)."""
        torch = self._torch
        h, w = diffuse.shape[0], diffuse.shape[1]
        if out is None:
            out = torch.empty((h, w, 4), dtype=torch.float32, device=self.device)
        self._ck(lib().vgi_specular_filter(self._h, C.c_void_p(diffuse.data_ptr()), C.c_void_p(specular.data_ptr()),
                                          C.c_uint32(w), C.c_uint32(h), C.byref(params) if params is not None else None,
                                          C.c_void_p(out.data_ptr()), _stream(stream)))
        return out

    # -- helper passes on caller-owned reference-layout atlases (uint8 CUDA tensors of atlas_shape)
    def atlas_clear_region(self, atlas, min_corner, extent, level, stream=None):
        self._ck(lib().vgi_atlas_clear_region(self._h, C.c_void_p(atlas.data_ptr()), (C.c_int32 * 3)(*min_corner),
                                              (C.c_uint32 * 3)(*extent), C.c_uint32(level), _stream(stream)))

    def atlas_copy_alpha(self, dst, src, level, stream=None):
        self._ck(lib().vgi_atlas_copy_alpha(self._h, C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()),
                                            C.c_uint32(level), _stream(stream)))

    def atlas_downsample(self, atlas, which, level, stream=None):
        self._ck(lib().vgi_atlas_downsample(self._h, C.c_void_p(atlas.data_ptr()), C.c_int(which), C.c_uint32(level),
                                            _stream(stream)))

    def atlas_wrap_border(self, atlas, stream=None):
        self._ck(lib().vgi_atlas_wrap_border(self._h, C.c_void_p(atlas.data_ptr()), _stream(stream)))

    # -- whole frame with host buffers (the e2e call)
    def frame_host(self, frame_index, camera_pos, camera, host_gbuffer, host_shadow_depth, params,
                   out_diffuse, out_specular, stream=None):
        """host_gbuffer: dict of (pinned) CPU torch tensors or numpy arrays; out_*: (H, W, 4) float32
        (pinned) CPU tensors / numpy arrays. Synchronous."""
        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        g = S.GBuffer()
        g.diffuse_rgba8 = ptr(host_gbuffer["diffuse"])
        g.normal_rgba16f = ptr(host_gbuffer["normal"])
        g.specular_rgba8 = ptr(host_gbuffer["specular"])
        g.emission_rgba16f = ptr(host_gbuffer["emission"])
        g.depth_f32 = ptr(host_gbuffer["depth"])
        g.height, g.width = host_gbuffer["depth"].shape[:2]
        sd = C.c_void_p(ptr(host_shadow_depth)) if host_shadow_depth is not None else C.c_void_p(0)
        prm = C.byref(params) if params is not None else None
        self._ck(lib().vgi_frame_host(self._h, C.c_uint32(frame_index), _f3(camera_pos), C.byref(camera), C.byref(g),
                                      sd, prm, C.c_void_p(ptr(out_diffuse)), C.c_void_p(ptr(out_specular)),
                                      _stream(stream)))

    def frame_view_host(self, frame_index, camera_pos, camera, width, height, shadow, params, out_diffuse, out_specular,
                        stream=None):
        """Batched / headless view: G-buffer (and, with shadow != None, the shadow map) rasterised on the device, only
        the camera goes up; out_*: (H, W, 4) float32 (pinned) CPU tensors / numpy arrays. Synchronous."""
        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        prm = C.byref(params) if params is not None else None
        sh = C.byref(shadow) if shadow is not None else None
        self._ck(lib().vgi_frame_view_host(self._h, C.c_uint32(frame_index), _f3(camera_pos), C.byref(camera),
                                           C.c_uint32(width), C.c_uint32(height), sh, prm,
                                           C.c_void_p(ptr(out_diffuse)), C.c_void_p(ptr(out_specular)), _stream(stream)))

    def frame_view_host_begin(self, frame_index, camera_pos, camera, width, height, shadow, params, out_diffuse, out_specular,
                              stream=None):
        """vgi_frame_view_host_begin: enqueue the frame and return; at most two frames in flight (frame_view_host_end)."""
        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        prm = C.byref(params) if params is not None else None
        sh = C.byref(shadow) if shadow is not None else None
        self._ck(lib().vgi_frame_view_host_begin(self._h, C.c_uint32(frame_index), _f3(camera_pos), C.byref(camera),
                                                 C.c_uint32(width), C.c_uint32(height), sh, prm,
                                                 C.c_void_p(ptr(out_diffuse)), C.c_void_p(ptr(out_specular)), _stream(stream)))

    def frame_view_host_end(self):
        """Wait for the oldest frame begun: its images are in their host buffers on return."""
        self._ck(lib().vgi_frame_view_host_end(self._h))

    # -- per-kernel timing
    def set_timing(self, enable=True):
        self._ck(lib().vgi_set_timing(self._h, C.c_int(1 if enable else 0)))

    def reset_timings(self):
        self._ck(lib().vgi_reset_timings(self._h))

    def timings(self):
        """{kernel name: (total ms, launches)} since the last reset."""
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        n_l = (C.c_uint64 * cap)()
        n = lib().vgi_get_timings(self._h, names, ms, n_l, C.c_uint32(cap))
        if n < 0:
            self._ck(n)
        return {names[i].decode(): (ms[i], int(n_l[i])) for i in range(n)}
