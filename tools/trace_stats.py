"""Development aid: build an instrumented libvgi (-DVGI_TRACE_STATS_BUILD) and print the cone tracer's
work counters for a config (steps, level samples, brick-mask skips, empty footprints, non-zero corners)."""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = os.path.join(ROOT, "tools", "_dev")
LIB = os.path.join(DEV, "libvgi_stats.so")


def build():
    from vk_voxel_cone_tracing_b200 import build as B
    B.build_libvgi()
    os.makedirs(DEV, exist_ok=True)
    obj = os.path.join(DEV, "vgi_trace_stats.o")
    subprocess.check_call([B.NVCC] + B.ARCH + B.COMMON + ["-DVGI_TRACE_STATS_BUILD", "-c", os.path.join(B.CSRC, "vgi_trace.cu"), "-o", obj])
    objs = [obj] + [os.path.join(B.CSRC, f) for f in ("vgi_build.o", "vgi_svo.o", "vgi_atlas.o", "vgi_raster.o", "vgi_post.o", "vgi_api.o")]
    subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", B.GXX, "-lcudart"])
    return LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-only", action="store_true")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    a = ap.parse_args()
    if a.build_only or not os.path.exists(LIB):
        build()
        if a.build_only:
            return
    os.environ["VGI_LIBVGI_PATH"] = LIB
    import torch  # noqa: F401
    from tests import common
    from vk_voxel_cone_tracing_b200 import api
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    inp = common.atrium_inputs(a.res, 4096, a.width, a.height, 6)
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    gi.build_clipmap(0)
    gb = gi.upload_gbuffer(inp["gbuffer"])
    out = (C.c_ulonglong * 8)()
    names = ["steps", "level_samples", "brick_skipped", "loaded_all_zero", "corners_loaded", "corners_nonzero",
             "lane_filtered", "coop_filtered"]
    for mode, label in ((7, "diffuse only (mode 7)"), (6, "specular only (mode 6)")):
        api.lib().vgi_debug_trace_stats(out, 1)
        gi.cone_trace(inp["cam"], gb, gi.default_vct_params(mode))
        api.lib().vgi_debug_trace_stats(out, 1)
        px = a.width * a.height
        print(label, {n: int(out[i]) for i, n in enumerate(names)}, "per pixel:", {n: round(out[i] / px, 1) for i, n in enumerate(names)})


if __name__ == "__main__":
    main()
