"""The C-ABI library loads on a CPU-only box and exports every symbol include/vgi.h declares; the
ctypes mirrors have the sizes the header's structs have (checked against a tiny C program).
No compute calls are made here (no GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vgi.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(vgi_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def libvgi():
    from vk_voxel_cone_tracing_b200 import build
    build.build_libvgi()
    return C.CDLL(build.LIBVGI)


def test_every_declared_symbol_is_exported(libvgi):
    names = _declared_functions()
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(libvgi, n)]
    assert not missing, missing


def test_version_and_default_config(libvgi):
    from vk_voxel_cone_tracing_b200 import structs as S
    assert libvgi.vgi_version() == 100
    cfg = S.Config()
    libvgi.vgi_default_config(C.byref(cfg))
    # reference defaults: EngineConfig.h:28-33, VoxelizationPass.h:57
    assert (cfg.resolution, cfg.level_count, cfg.downsample_band, cfg.extent_level0) == (128, 6, 10, 16.0)
    assert list(cfg.clip_min_change)[:6] == [2, 2, 2, 2, 2, 1]
    assert cfg.struct_size == C.sizeof(S.Config)
    ref = S.default_config()
    assert bytes(memoryview(ref)) == bytes(memoryview(cfg))


def test_create_without_gpu_fails_loudly(libvgi):
    """No CPU fallback: on a box without a CUDA device vgi_create must fail with VGI_E_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from vk_voxel_cone_tracing_b200 import structs as S
    cfg = S.default_config(32, 2)
    h = C.c_void_p()
    rc = libvgi.vgi_create(C.byref(cfg), C.byref(h))
    assert rc == S.VGI_E_CUDA
    libvgi.vgi_last_error.restype = C.c_char_p
    assert b"no CUDA device" in libvgi.vgi_last_error(None)
    # argument validation happens before any CUDA call
    cfg.resolution = 100
    assert libvgi.vgi_create(C.byref(cfg), C.byref(h)) == S.VGI_E_INVALID
    assert libvgi.vgi_create(None, C.byref(h)) == S.VGI_E_INVALID


def test_struct_sizes_match_header(tmp_path):
    from vk_voxel_cone_tracing_b200 import structs as S
    prog = tmp_path / "sizes.c"
    names = ["vgi_config", "vgi_clip_region", "vgi_camera", "vgi_dir_light", "vgi_dir_light_shadow", "vgi_material",
             "vgi_primitive", "vgi_node_matrix", "vgi_scene_desc", "vgi_gbuffer", "vgi_vct_params", "vgi_stats",
             "vgi_filter_params"]
    body = "".join(f'printf("%zu\\n", sizeof({n}));' for n in names)
    prog.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void){{{body}return 0;}}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-o", str(exe), str(prog)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirrors = [S.Config, S.ClipRegion, S.Camera, S.DirLight, S.DirLightShadow, S.Material, S.Primitive, S.NodeMatrix,
               S.SceneDesc, S.GBuffer, S.VctParams, S.Stats, S.FilterParams]
    assert sizes == [C.sizeof(m) for m in mirrors]
    assert C.sizeof(S.Material) == 80 and C.sizeof(S.VctParams) == 52   # gltf.glsl:8-26, VoxelConeTracingPass.h:46-59


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "vk_voxel_cone_tracing_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"pyoracle|vgi_oracle|from oracle|import oracle|vgo_", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
