"""Host-side logic of the N > 1 path with world_size-2 gloo process groups on CPU: slab / row / view
partitions and the variable-length record exchange used by the slab-sharded build."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vk_voxel_cone_tracing_b200 import multigpu as M


def test_partitions_cover_exactly():
    for R in (32, 64, 256, 512):
        for world in (1, 2, 4, 8):
            spans = [M.slab_range(R, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        M.slab_range(64, 0, 3)
    for h in (1080, 2160, 136, 7):
        for world in (1, 2, 3, 8):
            rows = [M.rows_for_rank(h, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            assert all(a[0] % 8 == 0 for a in rows)
    views = [M.views_for_rank(64, r, 8) for r in range(8)]
    assert sorted(sum(views, [])) == list(range(64)) and all(len(v) == 8 for v in views)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. slab all-gather of per-level mask words: every rank owns one contiguous chunk per level
        L, n = 3, 64
        full = torch.arange(L * n, dtype=torch.int32).reshape(L, n) * 7 + 3
        words = torch.zeros(L, n, dtype=torch.int32)
        chunk = n // world
        words[:, rank * chunk:(rank + 1) * chunk] = full[:, rank * chunk:(rank + 1) * chunk]
        M.all_gather_slabs(words, rank, world)
        ok1 = torch.equal(words, full)
        # 2. variable-length record exchange (rank r contributes 5 + 11 r records)
        cnt = 5 + 11 * rank
        rng = np.random.RandomState(100 + rank)
        ids = torch.from_numpy(rng.randint(0, 1 << 30, size=64).astype(np.int32))
        recs = torch.from_numpy(rng.randint(-2 ** 31, 2 ** 31 - 1, size=(64, 8)).astype(np.int32))
        counts, ids_all, recs_all = M.all_gather_varlen(ids, recs, cnt)
        ok2 = counts == [5 + 11 * r for r in range(world)]
        for r in range(world):
            rr = np.random.RandomState(100 + r)
            eid = rr.randint(0, 1 << 30, size=64).astype(np.int32)
            erec = rr.randint(-2 ** 31, 2 ** 31 - 1, size=(64, 8)).astype(np.int32)
            ok2 = ok2 and np.array_equal(ids_all[r, :counts[r]].numpy(), eid[:counts[r]])
            ok2 = ok2 and np.array_equal(recs_all[r, :counts[r]].numpy(), erec[:counts[r]])
        # 3. an empty contribution from one rank
        counts0, _, _ = M.all_gather_varlen(ids, recs, 0 if rank == 0 else 3)
        ok3 = counts0 == [0] + [3] * (world - 1)
        # 4. IPC handle blobs of the peer build come back concatenated in rank order on every rank
        blob = bytes([rank * 16 + k for k in range(16)]) * 12          # 192 bytes, distinct per rank
        everyone = M.exchange_handles(blob)
        ok3 = ok3 and everyone == b"".join(bytes([r * 16 + k for k in range(16)]) * 12 for r in range(world))
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_exchange_protocol_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
