"""On-disk dump of the voxel-GI state for diffing: the two atlases in the reference image layout
(Voxelizer.h:40-52: RGBA8, W = (R+2)*6, H = (R+2)*L, D = R+2), the clip regions (ClipmapRegion.h:8-18), the SVO fragment
list (voxelizer.frag:99-100 packing) and node pool (octreeNodeAlloc.comp:30-32 words).

One `.npz` (zip of .npy, numpy's documented format) with these arrays, any of which may be absent:
    format          "vgi-dump-1"
    config          uint32 [resolution, level_count, downsample_band, mode_flags], float32 extent_level0 in `extent_level0`
    regions         int32 (L, 3) min corners; `region_voxel_size` float32 (L,)
    opacity         uint8 (D, H, W, 4)         radiance   uint8 (D, H, W, 4)
    svo_level       uint32 scalar              svo_fragments uint32 (N, 2)        svo_nodes uint32 (M, 2)
A dump written from the GPU context, from the CPU oracle or from a Vulkan capture of the reference's images can be
compared with `diff()` / tools/diff_dump.py: atlases byte for byte (reported per level and face), fragment lists as
multisets (the append order of an atomic counter is free), node pools after canonical child ordering."""
import numpy as np

FORMAT = "vgi-dump-1"


def save(path, cfg=None, regions=None, opacity=None, radiance=None, svo_level=None, svo_fragments=None, svo_nodes=None):
    out = {"format": np.array(FORMAT)}
    if cfg is not None:
        out["config"] = np.array([cfg.resolution, cfg.level_count, cfg.downsample_band, cfg.mode_flags], dtype=np.uint32)
        out["extent_level0"] = np.array(cfg.extent_level0, dtype=np.float32)
    if regions is not None:
        out["regions"] = np.array([list(r.min_corner) for r in regions], dtype=np.int32)
        out["region_voxel_size"] = np.array([r.voxel_size for r in regions], dtype=np.float32)
    for name, a, dt in (("opacity", opacity, np.uint8), ("radiance", radiance, np.uint8),
                        ("svo_fragments", svo_fragments, np.uint32), ("svo_nodes", svo_nodes, np.uint32)):
        if a is not None:
            a = np.ascontiguousarray(_to_numpy(a))
            out[name] = a.view(dt).reshape(a.shape) if a.dtype.itemsize == np.dtype(dt).itemsize else a.astype(dt)
    if svo_level is not None:
        out["svo_level"] = np.array(svo_level, dtype=np.uint32)
    np.savez_compressed(path, **out)


def load(path):
    with np.load(path) as z:
        d = {k: z[k] for k in z.files}
    if str(d.get("format")) != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} file")
    return d


def from_context(gi, path, svo_level=None):
    """Dump what a VoxelGI context holds (atlases exported in the reference layout; SVO buffers if built)."""
    frags = nodes = None
    if svo_level is not None:
        frags, nodes = gi.svo_fragments(), gi.svo_nodes()
    save(path, cfg=gi.cfg, regions=gi.regions(), opacity=gi.export_atlas(0), radiance=gi.export_atlas(1),
         svo_level=svo_level, svo_fragments=frags, svo_nodes=nodes)


def _to_numpy(a):
    if hasattr(a, "detach"):        # a torch tensor (device or host)
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def canonical_nodes(nodes):
    """Node pool rewritten breadth-first from the root block (children visited in slot order), so that equal trees compare
    equal whatever order an atomic counter handed out the child blocks in. Returns (M', 2) uint32."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 2)
    if nodes.shape[0] < 8:
        return nodes.copy()
    out, queue, head = [], [0], 0          # queue of block base indices in BFS order
    new_base = {0: 0}
    while head < len(queue):
        base = queue[head]
        head += 1
        for j in range(8):
            x, y = int(nodes[base + j, 0]), int(nodes[base + j, 1])
            child = x & 0x7fffffff
            if child and child + 8 <= nodes.shape[0] and child not in new_base:
                new_base[child] = 8 * len(queue)
                queue.append(child)
            out.append([(x & 0x80000000) | (new_base.get(child, 0) if child else 0), y])
    return np.array(out, dtype=np.uint32)


def _sorted_rows(a):
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 2)
    key = a[:, 0].astype(np.uint64) | (a[:, 1].astype(np.uint64) << np.uint64(32))
    return a[np.argsort(key, kind="stable")]


def diff(a, b):
    """Differences between two loaded dumps as human-readable lines (empty list = identical where both have data)."""
    out = []
    for k in ("config", "extent_level0", "regions", "region_voxel_size", "svo_level"):
        if k in a and k in b and not np.array_equal(a[k], b[k]):
            out.append(f"{k}: {a[k].tolist()} vs {b[k].tolist()}")
    for k in ("opacity", "radiance"):
        if k not in a or k not in b:
            continue
        if a[k].shape != b[k].shape:
            out.append(f"{k}: shape {a[k].shape} vs {b[k].shape}")
            continue
        ne = a[k] != b[k]
        if ne.any():
            rb = a[k].shape[0]                       # R + 2
            levels, faces = a[k].shape[1] // rb, a[k].shape[2] // rb
            per = ne.reshape(rb, levels, rb, faces, rb, 4).sum(axis=(0, 2, 4, 5))
            worst = int(np.abs(a[k].astype(np.int16) - b[k].astype(np.int16)).max())
            cells = ", ".join(f"L{l}F{f}:{int(per[l, f])}" for l in range(levels) for f in range(faces) if per[l, f])
            out.append(f"{k}: {int(ne.sum())} bytes differ (max |delta| {worst}); per level/face: {cells}")
    if "svo_fragments" in a and "svo_fragments" in b:
        fa, fb = _sorted_rows(a["svo_fragments"]), _sorted_rows(b["svo_fragments"])
        if fa.shape != fb.shape:
            out.append(f"svo_fragments: {fa.shape[0]} vs {fb.shape[0]} fragments")
        elif not np.array_equal(fa, fb):
            out.append(f"svo_fragments: {int((fa != fb).any(axis=1).sum())} fragments differ (as multisets)")
    if "svo_nodes" in a and "svo_nodes" in b:
        na, nb = canonical_nodes(a["svo_nodes"]), canonical_nodes(b["svo_nodes"])
        if na.shape != nb.shape:
            out.append(f"svo_nodes: {na.shape[0]} vs {nb.shape[0]} nodes after canonical ordering")
        else:
            topo = int((na[:, 0] != nb[:, 0]).sum())
            col = int((na[:, 1] != nb[:, 1]).sum())
            if topo or col:
                out.append(f"svo_nodes: {topo} topology words, {col} colour words differ after canonical ordering")
    return out
