"""vgi_import_vk_memory / vgi_release_vk_memory exercised WITHOUT Vulkan.

What Vulkan's vkGetMemoryFdKHR hands out for an allocation made with VkExportMemoryAllocateInfo{OPAQUE_FD} is, on NVIDIA's
driver, the same kind of object the CUDA driver exports for a cuMemCreate allocation requested with
CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR: an opaque POSIX fd naming device memory. So the import path the patched host uses
(patches/0004: Buffer::getMemoryFd -> vgi_import_vk_memory -> cudaImportExternalMemory(OPAQUE_FD) ->
cudaExternalMemoryGetMappedBuffer) is driven here with an fd from cuMemExportToShareableHandle, and the mapped buffer is used
as the out_diffuse / out_specular target of vgi_cone_trace: the images written through the imported mapping must equal the
images traced into ordinary device memory, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from tests import common


def _check(res, what):
    err = res[0]
    assert int(err) == 0, f"{what}: {err}"
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


@pytest.mark.gpu
def test_cone_trace_into_imported_external_memory():
    torch = pytest.importorskip("torch")
    try:
        from cuda.bindings import driver as cu
    except Exception:
        cu = pytest.importorskip("cuda.cuda")
    from vk_voxel_cone_tracing_b200 import api
    from vk_voxel_cone_tracing_b200.api import VoxelGI

    torch.cuda.init()
    torch.zeros(1, device="cuda")                       # primary context current on this thread
    inp = common.cornell_inputs(32, 512, 96, 64)
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(0)
    gb = gi.upload_gbuffer(inp["gbuffer"])
    prm = gi.default_vct_params(8)
    want_d, want_s = gi.cone_trace(inp["cam"], gb, prm)
    torch.cuda.synchronize()

    h, w = inp["gbuffer"]["depth"].shape
    need = h * w * 16
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = torch.cuda.current_device()
    gran = _check(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM),
                  "cuMemGetAllocationGranularity")
    size = ((need + gran - 1) // gran) * gran

    lib = api.lib()
    lib.vgi_import_vk_memory.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.vgi_release_vk_memory.argtypes = [C.c_void_p, C.c_void_p]
    allocs, ptrs, handles = [], [], []
    for _ in range(2):
        handle = _check(cu.cuMemCreate(size, prop, 0), "cuMemCreate")
        fd = _check(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
                    "cuMemExportToShareableHandle")
        dptr, imp = C.c_void_p(), C.c_void_p()
        rc = lib.vgi_import_vk_memory(gi._h, int(fd), size, C.byref(dptr), C.byref(imp))   # the fd belongs to CUDA afterwards
        assert rc == 0, lib.vgi_last_error(gi._h).decode()
        assert dptr.value
        allocs.append(handle); ptrs.append(dptr.value); handles.append(imp)

    class _Ext:             # what VoxelGI.cone_trace needs of an output tensor: data_ptr()
        def __init__(self, p): self.p = p
        def data_ptr(self): return self.p
    gi.cone_trace(inp["cam"], gb, prm, out=(_Ext(ptrs[0]), _Ext(ptrs[1])))
    torch.cuda.synchronize()
    got = []
    for p in ptrs:
        host = np.empty((h, w, 4), dtype=np.float32)
        _check(cu.cuMemcpyDtoH(host.ctypes.data, p, need), "cuMemcpyDtoH")
        got.append(host)
    cov = inp["gbuffer"]["depth"] < 1.0
    assert np.array_equal(got[0][cov], want_d.cpu().numpy()[cov])
    assert np.array_equal(got[1][cov], want_s.cpu().numpy()[cov])
    assert float(np.abs(got[0][cov]).max()) > 0.0
    for imp, handle in zip(handles, allocs):
        assert lib.vgi_release_vk_memory(gi._h, imp) == 0
        _check(cu.cuMemRelease(handle), "cuMemRelease")
