"""The on-disk dump format (vk_voxel_cone_tracing_b200/dump.py): round trip, per-level/face atlas diff, fragment multisets,
canonical node ordering (checked against the oracle's canonicaliser and against a permuted allocation order)."""
import numpy as np

from tests import common
from vk_voxel_cone_tracing_b200 import dump, structs as S


def _permute_blocks(nodes, seed):
    """The same tree with its child blocks (all but the root block) allocated in another order."""
    nb = nodes.shape[0] // 8
    perm = np.concatenate([[0], 1 + np.random.RandomState(seed).permutation(nb - 1)])      # new position of block i
    out = np.zeros_like(nodes)
    for i in range(nb):
        blk = nodes[8 * i:8 * i + 8].copy()
        child = blk[:, 0] & 0x7fffffff
        has = child != 0
        blk[has, 0] = (blk[has, 0] & 0x80000000) | (8 * perm[child[has] // 8]).astype(np.uint32)
        out[8 * perm[i]:8 * perm[i] + 8] = blk
    return out


def test_round_trip_and_diff(oracle, tmp_path):
    inp = common.cornell_inputs(resolution=16, shadow_size=128, width=8, height=8)
    cfg = inp["cfg"]
    regs = oracle.regions(cfg, inp["cam_pos"])
    osc = oracle.OracleScene(inp["scene"])
    op, rad, _ = oracle.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    lo, hi = inp["scene"].world_bbox()
    frags = oracle.svo_fragments(5, lo, hi, osc, inp["light"], inp["shadow"], inp["shadow_depth"])
    nodes = oracle.svo_build(5, frags)
    a = str(tmp_path / "a.npz")
    dump.save(a, cfg=cfg, regions=regs, opacity=op, radiance=rad, svo_level=5, svo_fragments=frags, svo_nodes=nodes)
    d = dump.load(a)
    assert np.array_equal(d["opacity"], op) and np.array_equal(d["radiance"], rad)
    assert np.array_equal(d["svo_fragments"], frags) and np.array_equal(d["svo_nodes"], nodes)
    assert d["config"].tolist() == [16, 6, 10, 0] and d["regions"].shape == (6, 3)
    assert dump.diff(d, d) == []

    # same content, different allocation / append order: still identical
    b = str(tmp_path / "b.npz")
    dump.save(b, cfg=cfg, regions=regs, opacity=op, radiance=rad, svo_level=5,
              svo_fragments=frags[np.random.RandomState(1).permutation(frags.shape[0])], svo_nodes=_permute_blocks(nodes, 2))
    assert dump.diff(d, dump.load(b)) == []
    assert np.array_equal(dump.canonical_nodes(nodes), oracle.svo_canonicalize(nodes))
    assert np.array_equal(dump.canonical_nodes(_permute_blocks(nodes, 3)), dump.canonical_nodes(nodes))

    # real differences are located
    rad2 = rad.copy()
    rb = cfg.resolution + 2
    rad2[3, 2 * rb + 4, 1 * rb + 5, 0] ^= 0x10           # level 2, face 1
    nodes2 = nodes.copy()
    nodes2[9, 1] ^= 1
    c = str(tmp_path / "c.npz")
    dump.save(c, cfg=cfg, regions=regs, opacity=op, radiance=rad2, svo_level=5, svo_fragments=frags[:-1], svo_nodes=nodes2)
    lines = dump.diff(d, dump.load(c))
    assert any(l.startswith("radiance: 1 bytes differ") and "L2F1:1" in l for l in lines)
    assert any(l.startswith("svo_fragments:") for l in lines) and any("1 colour words" in l for l in lines)
    assert not any(l.startswith("opacity") for l in lines)
