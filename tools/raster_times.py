"""Development helper: per-kernel CUDA-event times of the device-side input producers (shadow map + G-buffer)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import common
from vk_voxel_cone_tracing_b200.api import VoxelGI


def main():
    inp = common.atrium_inputs(256, 4096, 1920, 1080, 6)
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    sm = gb = None
    for _ in range(3):
        sm = gi.render_shadow_map(inp["shadow"], 4096, out=sm)
        gb = gi.render_gbuffer(inp["cam"], 1920, 1080, out=gb)
    torch.cuda.synchronize()
    iters = 10
    for what in ("shadow", "gbuffer"):
        gi.set_timing(True)
        gi.reset_timings()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            if what == "shadow":
                gi.render_shadow_map(inp["shadow"], 4096, out=sm)
            else:
                gi.render_gbuffer(inp["cam"], 1920, 1080, out=gb)
        b.record()
        torch.cuda.synchronize()
        print(f"{what}: {a.elapsed_time(b) / iters:.3f} ms per call (with timing events)")
        for k, (ms, n) in gi.timings().items():
            if k.startswith("k_raster"):
                print(f"  {k:28s} {ms / iters * 1e3:9.1f} us  ({n // iters} launches)")
        gi.set_timing(False)
    ref = torch.from_numpy(inp["shadow_depth"]).cuda()
    print("shadow equals host-rendered:", bool(torch.equal(sm, ref)))


if __name__ == "__main__":
    main()
