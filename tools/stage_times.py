"""Development helper: CUDA-event timings of the clipmap build and the cone trace on a config."""
import argparse
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import common
from vk_voxel_cone_tracing_b200.api import VoxelGI


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="atrium")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--mode", type=int, default=8)
    ap.add_argument("--svo", type=int, default=0, help="also time the SVO path at this octree level")
    a = ap.parse_args()
    t0 = time.time()
    if a.scene == "atrium":
        inp = common.atrium_inputs(a.res, 4096, a.width, a.height, a.levels)
    else:
        inp = common.cornell_inputs(a.res, 1024, a.width, a.height)
    print(f"inputs {time.time()-t0:.2f}s tris={inp['scene'].triangle_count}", flush=True)
    gi = VoxelGI(inp["cfg"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    gb = gi.upload_gbuffer(inp["gbuffer"])
    prm = gi.default_vct_params(a.mode)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    vox, inj, trc = [], [], []
    out = None
    for i in range(a.iters + 3):
        ev[0].record()
        gi.voxelize_opacity()
        ev[1].record()
        gi.inject_radiance(0)
        ev[2].record()
        out = gi.cone_trace(inp["cam"], gb, prm, out=out)
        ev[3].record()
        torch.cuda.synchronize()
        if i >= 3:
            vox.append(ev[0].elapsed_time(ev[1]))
            inj.append(ev[1].elapsed_time(ev[2]))
            trc.append(ev[2].elapsed_time(ev[3]))
    st = gi.stats()
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"pairs={st.clip_pairs} occupied={st.occupied_voxels} launches={st.kernel_launches}")
    print(f"voxelize {med(vox):.3f} ms | inject+finalize+mip {med(inj):.3f} ms | trace {med(trc):.3f} ms")
    gi.set_timing(True)
    gi.reset_timings()
    for i in range(a.iters):
        gi.voxelize_opacity()
        gi.inject_radiance(0)
        gi.cone_trace(inp["cam"], gb, prm, out=out)
    torch.cuda.synchronize()
    for k, (ms, n) in gi.timings().items():
        print(f"  {k:20s} {ms / a.iters * 1e3:9.1f} us/frame  ({n // a.iters} launches)")
    gi.set_timing(False)
    if a.svo:
        lo, hi = inp["scene"].world_bbox()
        sprm = gi.default_vct_params(a.mode)
        sprm.volume_dimension = float(1 << a.svo)
        sprm.voxel_size = float((hi - lo).max() / (1 << a.svo))
        sprm.indirect_diffuse_intensity = 15.0
        sprm.occlusion_decay = 3.0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        sv, sb, stc = [], [], []
        for i in range(a.iters + 2):
            ev[0].record()
            gi.svo_voxelize(a.svo, lo, hi)
            ev[1].record()
            gi.svo_build()
            ev[2].record()
            so = gi.svo_cone_trace(inp["cam"], gb, sprm)
            ev[3].record()
            torch.cuda.synchronize()
            if i >= 2:
                sv.append(ev[0].elapsed_time(ev[1])); sb.append(ev[1].elapsed_time(ev[2])); stc.append(ev[2].elapsed_time(ev[3]))
        st = gi.stats()
        print(f"SVO level {a.svo}: fragments={st.svo_fragments} nodes={st.svo_nodes} | voxelize {med(sv):.3f} ms | build {med(sb):.3f} ms | trace {med(stc):.3f} ms")
    d = out[0]
    print("diffuse mean", float(d[..., :3].mean()), "spec mean", float(out[1][..., :3].mean()))


if __name__ == "__main__":
    main()
