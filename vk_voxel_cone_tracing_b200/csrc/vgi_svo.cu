// vgi_svo.cu — sparse voxel octree path (kernels land in a follow-up commit; the entry points
// report VGI_E_UNSUPPORTED until then so that callers fail loudly).
#include "vgi_internal.h"

extern "C" {
int vgi_svo_voxelize(vgi_ctx*, uint32_t, const float*, const float*, void*) { return VGI_E_UNSUPPORTED; }
int vgi_svo_build(vgi_ctx*, void*) { return VGI_E_UNSUPPORTED; }
int vgi_svo_get_fragments(vgi_ctx*, void**, uint32_t*) { return VGI_E_UNSUPPORTED; }
int vgi_svo_get_nodes(vgi_ctx*, void**, uint32_t*) { return VGI_E_UNSUPPORTED; }
int vgi_svo_cone_trace(vgi_ctx*, const vgi_camera*, const vgi_gbuffer*, const vgi_vct_params*, void*, void*, void*) { return VGI_E_UNSUPPORTED; }
int vgi_atlas_clear_region(vgi_ctx*, void*, const int32_t*, const uint32_t*, uint32_t, void*) { return VGI_E_UNSUPPORTED; }
int vgi_atlas_copy_alpha(vgi_ctx*, void*, const void*, uint32_t, void*) { return VGI_E_UNSUPPORTED; }
int vgi_atlas_downsample(vgi_ctx*, void*, int, uint32_t, void*) { return VGI_E_UNSUPPORTED; }
int vgi_atlas_wrap_border(vgi_ctx*, void*, void*) { return VGI_E_UNSUPPORTED; }
}
