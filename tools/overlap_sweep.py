"""Development helper: the cone trace with the specular march beside the diffuse one (vgi_set_trace_overlap), swept over
the resident specular blocks per SM; images compared bit for bit with the serial order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import common
from vk_voxel_cone_tracing_b200.api import VoxelGI


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    inp = common.atrium_inputs(256, 4096, W, H, 6)
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    gb = gi.upload_gbuffer(inp["gbuffer"])
    prm = gi.default_vct_params(8)
    gi.build_clipmap(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ref = None
    for blocks in (0, 1, 2, 3, 4, 6, 0):
        gi.set_trace_overlap(blocks)
        ts = []
        out = None
        for i in range(13):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = gi.cone_trace(inp["cam"], gb, prm, out=out)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ts.sort()
        if ref is None:
            ref = [o.clone() for o in out]
        same = all(torch.equal(o.view(torch.int32), r.view(torch.int32)) for o, r in zip(out, ref))
        gi.set_timing(True)
        gi.reset_timings()
        for i in range(5):
            gi.cone_trace(inp["cam"], gb, prm, out=out)
        torch.cuda.synchronize()
        tm = {k: round(ms / 5, 3) for k, (ms, n) in gi.timings().items() if k.startswith(("k_trace", "cone_trace"))}
        gi.set_timing(False)
        print(f"spec blocks beside = {blocks}: trace median {ts[len(ts)//2]:.3f} ms, min {ts[0]:.3f}; identical to serial: {same}; {tm}", flush=True)


if __name__ == "__main__":
    main()
