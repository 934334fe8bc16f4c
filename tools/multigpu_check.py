"""Run under torchrun (one rank per GPU): verifies on real GPUs that
  1. the slab-sharded clipmap build (NCCL all-gather of occupancy words + packed records) equals the
     replicated single-GPU build bit for bit over several frames with a moving camera and cadence,
  2. row-sharded cone tracing (+ all-gather of the row blocks) equals the full-image trace,
and times both formulations. Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--time-iters", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tests import common
    from vk_voxel_cone_tracing_b200 import multigpu as M
    from vk_voxel_cone_tracing_b200.api import VoxelGI
    if a.scene == "cornell":
        inp = common.cornell_inputs(a.res, 1024, a.width, a.height)
    else:
        inp = common.atrium_inputs(a.res, 4096, a.width, a.height, 6)

    def make():
        gi = VoxelGI(inp["cfg"], device=local)
        gi.set_scene(inp["scene"])
        gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
        return gi

    ref, shard, peer = make(), make(), make()
    sb = M.SlabBuild(shard)
    pb = M.PeerBuild(peer)
    ok_build, ok_peer = True, True
    for frame in range(a.frames):
        cam = tuple(np.array(inp["cam_pos"]) + np.array([0.37, -0.11, 0.29]) * frame)
        ref.update_regions(cam)
        shard.update_regions(cam)
        peer.update_regions(cam)
        ref.build_clipmap(frame)
        sb.build(frame)
        pb.build(frame)
        for which in (0, 1):
            want = ref.export_atlas(which)
            ok_build = ok_build and bool(torch.equal(want, shard.export_atlas(which)))
            ok_peer = ok_peer and bool(torch.equal(want, peer.export_atlas(which)))
        if peer.stats().occupied_voxels == 0:
            ok_peer = False
    # row-sharded trace on the sharded store vs full trace on the reference store
    gb = ref.upload_gbuffer(inp["gbuffer"])
    prm = ref.default_vct_params(8)
    full = ref.cone_trace(inp["cam"], gb, prm)
    y0, y1 = M.rows_for_rank(a.height, rank, world)
    part = shard.cone_trace(inp["cam"], gb, prm, rows=(y0, y1))
    ok_trace = True
    for k in range(2):
        ok_trace = ok_trace and torch.equal(part[k][y0:y1], full[k][y0:y1])
    # interleaved tile rows: the balanced single-view sharding
    inter = shard.cone_trace(inp["cam"], gb, prm, part=(rank, world))
    tile_rows = torch.arange(a.height, device="cuda") // 8
    mine = (tile_rows % world) == rank
    for k in range(2):
        ok_trace = ok_trace and torch.equal(inter[k][mine], full[k][mine])
    # the peer-built store traces to the same image
    ptrace = peer.cone_trace(inp["cam"], gb, prm)
    ok_peer = ok_peer and all(torch.equal(ptrace[k], full[k]) for k in range(2))
    flags = torch.tensor([int(ok_build), int(ok_trace), int(ok_peer)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)

    timing = {}
    if a.time_iters:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for name, fn in (("replicated_build_ms", lambda: ref.build_clipmap(0)), ("slab_build_ms", lambda: sb.build(0)),
                         ("peer_build_ms", lambda: pb.build(0)),
                         ("full_trace_ms", lambda: ref.cone_trace(inp["cam"], gb, prm, out=full)),
                         ("row_sharded_trace_ms", lambda: shard.cone_trace(inp["cam"], gb, prm, out=part, rows=(y0, y1))),
                         ("tile_interleaved_trace_ms", lambda: shard.cone_trace(inp["cam"], gb, prm, out=inter, part=(rank, world)))):
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(a.time_iters):
                fn()
            ev[1].record()
            torch.cuda.synchronize()
            t = torch.tensor([ev[0].elapsed_time(ev[1]) / a.time_iters], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            timing[name] = float(t)
    if a.time_iters:
        # per-kernel times of the peer build (CUDA events on this rank's stream; barriers include the wait for the slowest GPU)
        peer.set_timing(True)
        peer.reset_timings()
        dist.barrier()
        for _ in range(a.time_iters):
            pb.build(0)
        torch.cuda.synchronize()
        timing["peer_build_kernels_us_rank0"] = {k: round(v[0] / a.time_iters * 1e3, 1) for k, v in peer.timings().items()}
        peer.set_timing(False)
    if rank == 0:
        print(json.dumps({"world": world, "scene": a.scene, "res": a.res, "slab_build_bit_exact": bool(flags[0]),
                          "row_sharded_trace_bit_exact": bool(flags[1]), "peer_build_bit_exact": bool(flags[2]),
                          **timing}), flush=True)
    pb.close()
    dist.destroy_process_group()
    if not (flags[0] and flags[1] and flags[2]):
        sys.exit(1)


if __name__ == "__main__":
    main()
