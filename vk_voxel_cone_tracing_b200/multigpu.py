"""Multi-GPU plumbing for the voxel-GI hot path: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests) for the collectives; all compute stays in libvgi.

SURVEY.md section 8(e):
  * cone tracing shards with no collective — by camera view (`views_for_rank`) or by image rows
    (`rows_for_rank` + VoxelGI.cone_trace(rows=...));
  * the clipmap build shards by z slabs with ONE exchange step: `SlabBuild.build()` voxelizes and injects
    the rank's texel planes, all-gathers the occupancy words (a slab is one contiguous range per level because
    z is the slowest axis), finalizes and packs its records, all-gathers the packed records (variable length:
    counts first, then buffers padded to the longest), scatters the other ranks' records into the local store
    and runs the mips replicated. The result equals the single-GPU build bit for bit
    (tools/multigpu_check.py verifies that on the GPU box).

  * `PeerBuild` is the same sharding with the exchange done by the kernels themselves over NVLink peer memory
    (CUDA IPC mapped stores, direct stores into every GPU, flag barriers in the stream; vgi_peer_*).

Measured (B200, configs[1], profiles/r1_peer_build.json): replicated build 0.74 ms per GPU; NCCL slab build
0.80-0.86 ms (two all-gathers + one host read of the record counts cost what the sharded voxelize + inject
saves); peer build 0.58 / 0.47 / 0.44 ms on 2 / 4 / 8 GPUs. bench.py renders N different views (different clip
regions), so it replicates the build and shards the views."""
import ctypes as C

import torch
import torch.distributed as dist


def slab_range(resolution, rank, world):
    """Texel planes [z0, z1) owned by `rank` (resolution is a power of two, world divides it)."""
    if resolution % world:
        raise ValueError(f"world size {world} must divide the resolution {resolution}")
    n = resolution // world
    return rank * n, (rank + 1) * n


def rows_for_rank(height, rank, world, tile=8):
    """Contiguous block of image rows for `rank`, aligned to the tracer's 8-row tiles."""
    tiles = (height + tile - 1) // tile
    t0 = (tiles * rank) // world
    t1 = (tiles * (rank + 1)) // world
    return min(t0 * tile, height), min(t1 * tile, height)


def views_for_rank(n_views, rank, world):
    """Round-robin view assignment (64 views over 8 GPUs -> 8 each)."""
    return list(range(rank, n_views, world))


def all_gather_varlen(ids, recs, count, group=None):
    """All-gather `count` leading rows of (ids[cap], recs[cap, k]) from every rank.
    Returns (counts list, ids_all [world, maxc], recs_all [world, maxc, k]). One small collective for the
    counts (host read), one for each padded buffer."""
    world = dist.get_world_size(group)
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=ids.device)
    cnts = torch.empty(world, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = [int(x) for x in cnts.tolist()]
    maxc = max(max(counts), 1)
    pad_ids = torch.zeros(maxc, dtype=ids.dtype, device=ids.device)
    pad_recs = torch.zeros((maxc,) + tuple(recs.shape[1:]), dtype=recs.dtype, device=recs.device)
    pad_ids[:count] = ids[:count]
    pad_recs[:count] = recs[:count]
    # flat 1-D buffers: every backend (nccl, gloo) accepts the concatenated form
    ids_all = torch.empty(world * maxc, dtype=ids.dtype, device=ids.device)
    recs_all = torch.empty(world * pad_recs.numel(), dtype=recs.dtype, device=recs.device)
    dist.all_gather_into_tensor(ids_all, pad_ids, group=group)
    dist.all_gather_into_tensor(recs_all, pad_recs.reshape(-1), group=group)
    return counts, ids_all.view(world, maxc), recs_all.view((world, maxc) + tuple(recs.shape[1:]))


def all_gather_slabs(words, rank, world, group=None):
    """In-place all-gather of per-level slabs: `words` is (L, n) with level rows split in `world` equal
    contiguous chunks, chunk `rank` valid locally. One collective for all levels."""
    L, n = words.shape
    chunk = n // world
    view = words.view(L, world, chunk)
    mine = view[:, rank, :].contiguous().reshape(-1)
    out = torch.empty(world * L * chunk, dtype=words.dtype, device=words.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    view.copy_(out.view(world, L, chunk).permute(1, 0, 2))
    return words


class SlabBuild:
    """Slab-sharded clipmap build on top of one VoxelGI ctx per rank."""

    def __init__(self, gi, group=None):
        from . import api
        self.gi, self.group, self._api = gi, group, api
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.z0, self.z1 = slab_range(gi.cfg.resolution, self.rank, self.world)
        gi.set_slab(self.z0, self.z1)

    def _occupancy(self):
        gi, api = self.gi, self._api
        p, n = C.c_void_p(), C.c_size_t()
        gi._ck(api.lib().vgi_get_occupancy(gi._h, C.byref(p), C.byref(n)))
        L = gi.cfg.level_count
        return torch.as_tensor(api._DevView(p.value, (L, n.value // 4), "<i4"), device=gi.device)

    def build(self, frame_index=0):
        gi, api, S = self.gi, self._api, self._api.S
        lib = api.lib()
        st = api._stream()
        gi._ck(lib.vgi_slab_build_begin(gi._h, C.c_uint32(frame_index), st))
        all_gather_slabs(self._occupancy(), self.rank, self.world, self.group)
        gi._ck(lib.vgi_slab_finalize(gi._h, C.c_uint32(frame_index), st))
        ids_p, recs_p, cnt = C.c_void_p(), C.c_void_p(), C.c_uint32()
        gi._ck(lib.vgi_get_slab_pack(gi._h, C.byref(ids_p), C.byref(recs_p), C.byref(cnt)))
        n = max(cnt.value, 1)
        ids = torch.as_tensor(api._DevView(ids_p.value, (n,), "<i4"), device=gi.device)
        recs = torch.as_tensor(api._DevView(recs_p.value, (n, 8), "<i4"), device=gi.device)
        counts, ids_all, recs_all = all_gather_varlen(ids, recs, cnt.value, self.group)
        for r in range(self.world):
            if r == self.rank or counts[r] == 0:
                continue
            gi._ck(lib.vgi_slab_unpack(gi._h, C.c_void_p(ids_all[r].data_ptr()), C.c_void_p(recs_all[r].data_ptr()),
                                      C.c_uint32(counts[r]), st))
        gi._ck(lib.vgi_slab_build_end(gi._h, C.c_uint32(frame_index), st))
        self._keep = (ids_all, recs_all)   # alive until the unpack kernels ran
        return sum(counts)


def exchange_handles(blob, group=None):
    """All ranks contribute one bytes blob (their IPC handles); returns the concatenation in rank order."""
    world = dist.get_world_size(group)
    blobs = [None] * world
    dist.all_gather_object(blobs, bytes(blob), group=group)
    assert all(len(b) == len(blob) for b in blobs)
    return b"".join(blobs)


class PeerBuild:
    """Slab-sharded clipmap build whose exchange is done by the kernels themselves over NVLink peer memory
    (vgi_peer_*): every GPU maps every other GPU's voxel store / occupancy words / arrival flags through CUDA IPC, writes
    its slab's results into all stores and meets the others at flag barriers inside the stream. One process per GPU of
    one node; torch.distributed only carries the 192-byte handle blobs at set-up."""

    HANDLE_BYTES = 3 * 64

    def __init__(self, gi, group=None):
        from . import api
        self.gi, self.group, self._api = gi, group, api
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = api.lib()
        mine = (C.c_byte * self.HANDLE_BYTES)()
        gi._ck(lib.vgi_peer_export(gi._h, mine))
        everyone = exchange_handles(bytes(mine), group)
        buf = (C.c_byte * len(everyone)).from_buffer_copy(everyone)
        gi._ck(lib.vgi_peer_attach(gi._h, C.c_uint32(self.rank), C.c_uint32(self.world), buf))
        dist.barrier(group)            # every rank has mapped every buffer before the first build

    def build(self, frame_index=0, stream=None):
        gi, api = self.gi, self._api
        gi._ck(api.lib().vgi_peer_build_clipmap(gi._h, C.c_uint32(frame_index), api._stream(stream)))

    def close(self):
        gi, api = self.gi, self._api
        dist.barrier(self.group)       # nobody unmaps while a peer may still write
        gi._ck(api.lib().vgi_peer_detach(gi._h))
