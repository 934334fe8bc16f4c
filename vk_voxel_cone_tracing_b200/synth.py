"""Deterministic synthetic inputs for the voxel-GI hot path: procedural scenes in
the exact SoA buffer layout GLTFScene uploads (ref: VFS/GLTFScene.cpp:55-93, 398-411, 457-490),
reference-default camera and light (ref: VFS/Camera.h:48-56, VFS/Application.cpp:386-392,
VFS/DirectionalLight.cpp:21-47), and helpers that pack them into the include/vgi.h structs.

There is no network and Sponza.bin is missing from the reference checkout (.MISSING_LARGE_BLOBS), so
every benchmark and test runs on these generators."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import glm
from .structs import (MATERIAL_DTYPE, NODE_DTYPE, PRIMITIVE_DTYPE, Camera, DirLight, DirLightShadow,
                      SceneDesc)


@dataclass
class Scene:
    positions: np.ndarray
    normals: np.ndarray
    texcoords: np.ndarray
    indices: np.ndarray
    primitives: np.ndarray
    nodes: np.ndarray
    materials: np.ndarray
    name: str = "scene"
    _keep: list = field(default_factory=list)
    tangents: np.ndarray = None     # vec4 per vertex (xyz, handedness), only the G-buffer's normal mapping reads them

    @property
    def triangle_count(self):
        return int(self.primitives["index_count"].sum() // 3)

    @property
    def vertex_count(self):
        return int(self.positions.shape[0])

    def desc(self):
        """vgi_scene_desc pointing at this scene's (contiguous) host arrays."""
        d = SceneDesc()
        arrs = [np.ascontiguousarray(a) for a in (self.positions, self.normals, self.texcoords, self.indices,
                                                  self.primitives, self.nodes, self.materials)]
        self._keep = arrs
        (d.positions, d.normals, d.texcoords, d.indices, d.primitives, d.nodes, d.materials) = [
            a.ctypes.data for a in arrs]
        d.vertex_count = self.positions.shape[0]
        d.index_count = self.indices.shape[0]
        d.primitive_count = self.primitives.shape[0]
        d.node_count = self.nodes.shape[0]
        d.material_count = self.materials.shape[0]
        if self.tangents is not None:
            tg = np.ascontiguousarray(self.tangents, np.float32)
            self._keep.append(tg)
            d.tangents = tg.ctypes.data
        return d

    def world_bbox(self):
        lo = np.full(3, np.inf)
        hi = np.full(3, -np.inf)
        for p in self.primitives:
            idx = self.indices[p["first_index"]:p["first_index"] + p["index_count"]] + p["vertex_offset"]
            m = self.nodes[p["node_index"]]["model"].reshape(4, 4).T.astype(np.float64)
            v = self.positions[np.unique(idx)].astype(np.float64)
            w = v @ m[:3, :3].T + m[:3, 3]
            lo = np.minimum(lo, w.min(0))
            hi = np.maximum(hi, w.max(0))
        return lo.astype(np.float32), hi.astype(np.float32)


class _Builder:
    def __init__(self):
        self.pos, self.nrm, self.uv, self.idx = [], [], [], []
        self.prims, self.nodes, self.mats = [], [], []
        self.nv = 0
        self.ni = 0

    def node(self, model):
        model = np.asarray(model, dtype=np.float32)
        it = np.linalg.inv(model.T.astype(np.float64)).T  # inverse of the maths matrix ...
        it = it.T.astype(np.float32)                      # ... transposed (stored [col][row])
        self.nodes.append((model.reshape(16), it.reshape(16)))
        return len(self.nodes) - 1

    def material(self, base=(1, 1, 1, 1), metallic=0.0, roughness=1.0, emissive=(0, 0, 0)):
        m = np.zeros((), dtype=MATERIAL_DTYPE)
        m["base_color_factor"] = base
        m["base_color_texture"] = -1
        m["metallic_factor"] = metallic
        m["roughness_factor"] = roughness
        m["metallic_roughness_texture"] = -1
        m["emissive_texture"] = -1
        m["alpha_mode"] = 0
        m["alpha_cutoff"] = 0.5
        m["double_sided"] = 0
        m["emissive_factor"] = emissive
        m["normal_texture"] = -1
        m["normal_texture_scale"] = 1.0
        m["occlusion_texture"] = -1
        m["occlusion_texture_strength"] = 1.0
        self.mats.append(m)
        return len(self.mats) - 1

    def grid(self, origin, du, dv, nu, nv, normal, material, node=0, displace=None):
        """(nu x nv)-quad grid spanning origin + s*du + t*dv, s,t in [0,1]; 2*nu*nv triangles."""
        origin, du, dv = (np.asarray(a, dtype=np.float64) for a in (origin, du, dv))
        s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="ij")
        p = origin + s[..., None] * du + t[..., None] * dv
        n = np.broadcast_to(np.asarray(normal, dtype=np.float64), p.shape).copy()
        if displace is not None:
            p, n = displace(p, n, s, t)
        self._add_grid(p, n, s, t, nu, nv, material, node)

    def _add_grid(self, p, n, s, t, nu, nv, material, node):
        i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
        a = (i * (nv + 1) + j).ravel()
        b = a + (nv + 1)
        tri = np.stack([a, b, b + 1, a, b + 1, a + 1], axis=1).ravel().astype(np.uint32)
        self.pos.append(p.reshape(-1, 3).astype(np.float32))
        self.nrm.append(n.reshape(-1, 3).astype(np.float32))
        self.uv.append(np.stack([s.ravel(), t.ravel()], axis=1).astype(np.float32))
        self.idx.append(tri)
        self.prims.append((self.ni, tri.size, self.nv, material, node))
        self.nv += p.reshape(-1, 3).shape[0]
        self.ni += tri.size

    def cylinder(self, base, radius, height, nseg, nstack, material, node=0):
        """Open cylinder around +y; 2*nseg*nstack triangles."""
        s, t = np.meshgrid(np.linspace(0, 1, nseg + 1), np.linspace(0, 1, nstack + 1), indexing="ij")
        ang = 2 * np.pi * s
        n = np.stack([np.cos(ang), np.zeros_like(ang), np.sin(ang)], axis=-1)
        p = np.asarray(base, dtype=np.float64) + radius * n + np.stack(
            [np.zeros_like(t), t * height, np.zeros_like(t)], axis=-1)
        self._add_grid(p, n, s, t, nseg, nstack, material, node)

    def arch(self, c0, span_dir, span, rise, depth_dir, depth, nseg, nacross, material, node=0):
        """Half-elliptical band from c0 to c0+span*span_dir rising `rise` in +y; 2*nseg*nacross tris."""
        s, t = np.meshgrid(np.linspace(0, 1, nseg + 1), np.linspace(0, 1, nacross + 1), indexing="ij")
        ang = np.pi * s
        sd, dd = np.asarray(span_dir, dtype=np.float64), np.asarray(depth_dir, dtype=np.float64)
        up = np.array([0.0, 1.0, 0.0])
        p = (np.asarray(c0, dtype=np.float64) + (0.5 * span * (1 - np.cos(ang)))[..., None] * sd
             + (rise * np.sin(ang))[..., None] * up + ((t - 0.5) * depth)[..., None] * dd)
        n = -(np.cos(ang) * rise)[..., None] * sd * -1.0 - (np.sin(ang) * 0.5 * span)[..., None] * up
        n = n / np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-12)
        self._add_grid(p, n, s, t, nseg, nacross, material, node)

    def build(self, name):
        return Scene(
            positions=np.concatenate(self.pos), normals=np.concatenate(self.nrm),
            texcoords=np.concatenate(self.uv), indices=np.concatenate(self.idx),
            primitives=np.array(self.prims, dtype=PRIMITIVE_DTYPE),
            nodes=np.array(self.nodes, dtype=NODE_DTYPE),
            materials=np.array(self.mats, dtype=MATERIAL_DTYPE), name=name)


def _box(b, center, half, rot_y_deg, n, material, node):
    """Axis box rotated about y, 6 faces of n x n quads (12 n^2 triangles)."""
    c = np.asarray(center, dtype=np.float64)
    h = np.asarray(half, dtype=np.float64)
    a = np.radians(rot_y_deg)
    rot = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    faces = [  # (normal, u axis, v axis) with u x v = normal
        ((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
        ((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
        ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0)),
    ]
    for nrm, ua, va in faces:
        nrm, ua, va = (np.asarray(x, dtype=np.float64) for x in (nrm, ua, va))
        o = nrm * h - ua * h - va * h
        b.grid(c + rot @ o, rot @ (2 * ua * h), rot @ (2 * va * h), n, n, rot @ nrm, material, node)


def cornell_box(wall_quads=32, box_quads=10):
    """Config 1: open-top room, interior [-4,4]^3, red left / green right / white floor and
    back wall, short box rotated +18 deg, tall metallic box rotated -18 deg, one emissive 2x2-quad
    light near the top. ~10.6k triangles, factor-only materials."""
    b = _Builder()
    node = b.node(np.eye(4, dtype=np.float32).reshape(4, 4))
    white = b.material((0.725, 0.725, 0.725, 1.0))
    red = b.material((0.63, 0.065, 0.05, 1.0))
    green = b.material((0.14, 0.45, 0.091, 1.0))
    metal = b.material((0.8, 0.8, 0.8, 1.0), metallic=1.0, roughness=0.2)
    light = b.material((1.0, 1.0, 1.0, 1.0), emissive=(1.0, 1.0, 1.0))
    n = wall_quads
    b.grid((-4, -4, -4), (8, 0, 0), (0, 0, 8), n, n, (0, 1, 0), white, node)      # floor (y = -4)
    b.grid((-4, -4, -4), (8, 0, 0), (0, 8, 0), n, n, (0, 0, 1), white, node)      # back wall (z = -4)
    b.grid((-4, -4, -4), (0, 0, 8), (0, 8, 0), n, n, (1, 0, 0), red, node)        # left wall (x = -4)
    b.grid((4, -4, -4), (0, 8, 0), (0, 0, 8), n, n, (-1, 0, 0), green, node)      # right wall (x = 4)
    _box(b, (1.7, -2.8, -1.7), (1.2, 1.2, 1.2), 18.0, box_quads, white, node)
    _box(b, (-1.9, -1.6, -2.5), (1.2, 2.4, 1.2), -18.0, box_quads, metal, node)
    b.grid((-1, 3.6, -2.5), (2, 0, 0), (0, 0, 2), 2, 2, (0, -1, 0), light, node)  # emissive quad
    return b.build("cornell")


def coarse_room(quads=1, half=12.0, height=6.0):
    """A hall whose floor, ceiling and four walls are `quads` x `quads` grids (2 triangles each at quads = 1): for a camera
    inside, most of these triangles have a vertex BEHIND the camera plane — the case a hardware rasteriser clips and the
    software rasterisers must not drop (near-plane handling)."""
    b = _Builder()
    node = b.node(np.eye(4, dtype=np.float32))
    cols = [(0.8, 0.8, 0.8, 1.0), (0.7, 0.3, 0.2, 1.0), (0.2, 0.6, 0.3, 1.0), (0.3, 0.3, 0.8, 1.0), (0.8, 0.7, 0.2, 1.0),
            (0.6, 0.6, 0.6, 1.0)]
    m = [b.material(c) for c in cols]
    h, y0, y1, n = half, 0.0, height, quads
    b.grid((-h, y0, -h), (2 * h, 0, 0), (0, 0, 2 * h), n, n, (0, 1, 0), m[0], node)       # floor
    b.grid((-h, y1, -h), (0, 0, 2 * h), (2 * h, 0, 0), n, n, (0, -1, 0), m[5], node)      # ceiling
    b.grid((-h, y0, -h), (2 * h, 0, 0), (0, y1 - y0, 0), n, n, (0, 0, 1), m[1], node)     # z = -h
    b.grid((-h, y0, h), (0, y1 - y0, 0), (2 * h, 0, 0), n, n, (0, 0, -1), m[2], node)     # z = +h
    b.grid((-h, y0, -h), (0, y1 - y0, 0), (0, 0, 2 * h), n, n, (1, 0, 0), m[3], node)     # x = -h
    b.grid((h, y0, -h), (0, 0, 2 * h), (0, y1 - y0, 0), n, n, (-1, 0, 0), m[4], node)     # x = +h
    return b.build("coarse_room")


def procedural_textures(seed=7):
    """Four small RGBA8 textures for the textured fixtures: [0] base colour (coloured checker with a smooth gradient, alpha
    0.6-1), [1] emissive pattern (dim stripes), [2] occlusion mask (.r = 0 inside round holes, 1 elsewhere: alpha test),
    [3] a 3 x 5 non-power-of-two noise image (REPEAT wrap on awkward sizes)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:64, 0:64]
    base = np.zeros((64, 64, 4), np.uint8)
    chk = ((xx // 8 + yy // 8) % 2).astype(np.float64)
    base[..., 0] = (90 + 150 * chk).astype(np.uint8)
    base[..., 1] = (40 + 3 * xx).astype(np.uint8)
    base[..., 2] = (250 - 3 * yy).astype(np.uint8)
    base[..., 3] = (153 + 102 * (1 - chk)).astype(np.uint8)
    emis = np.zeros((32, 16, 4), np.uint8)
    yy2, xx2 = np.mgrid[0:32, 0:16]
    emis[..., 0] = (40 * ((yy2 // 4) % 2)).astype(np.uint8)
    emis[..., 1] = (25 * ((xx2 // 2) % 2)).astype(np.uint8)
    emis[..., 2] = 10
    emis[..., 3] = 255
    occ = np.full((48, 48, 4), 255, np.uint8)
    yy3, xx3 = np.mgrid[0:48, 0:48]
    holes = (((xx3 % 16) - 8) ** 2 + ((yy3 % 16) - 8) ** 2) < 20
    occ[holes, 0] = 0
    noise = rng.randint(0, 256, (5, 3, 4)).astype(np.uint8)
    # [4] tangent-space normal map (egg-crate bumps), [5] metallic-roughness map (g = roughness ramp, b = metallic patches):
    # read by the G-buffer producer only (gBufferPass.frag)
    yy4, xx4 = np.mgrid[0:32, 0:32]
    sx = 0.6 * np.cos(2 * np.pi * xx4 / 16.0)
    sy = 0.6 * np.cos(2 * np.pi * yy4 / 8.0)
    nz = np.sqrt(np.maximum(1.0 - 0.5 * (sx * sx + sy * sy), 0.0))
    nmap = np.zeros((32, 32, 4), np.uint8)
    nmap[..., 0] = np.round((sx * 0.5 + 0.5) * 255)
    nmap[..., 1] = np.round((sy * 0.5 + 0.5) * 255)
    nmap[..., 2] = np.round((nz * 0.5 + 0.5) * 255)
    nmap[..., 3] = 255
    mr = np.zeros((16, 24, 4), np.uint8)
    yy5, xx5 = np.mgrid[0:16, 0:24]
    mr[..., 0] = 255
    mr[..., 1] = (20 + 9 * xx5).astype(np.uint8)
    mr[..., 2] = (255 * ((xx5 // 6 + yy5 // 4) % 2)).astype(np.uint8)
    mr[..., 3] = 255
    return [base, emis, occ, noise, nmap, mr]


def compute_tangents(scene):
    """Per-vertex tangents (xyz, handedness) of a Scene from its positions, normals and texture coordinates: the usual
    accumulation of per-triangle dP/du, Gram-Schmidt against the normal, handedness from dP/dv."""
    pos = scene.positions.astype(np.float64)
    uv = scene.texcoords.astype(np.float64)
    nrm = scene.normals.astype(np.float64)
    tan = np.zeros_like(pos)
    bit = np.zeros_like(pos)
    for pr in scene.primitives:
        idx = scene.indices[int(pr["first_index"]):int(pr["first_index"]) + int(pr["index_count"])].astype(np.int64)
        idx = idx.reshape(-1, 3) + int(pr["vertex_offset"])
        e1, e2 = pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]]
        d1, d2 = uv[idx[:, 1]] - uv[idx[:, 0]], uv[idx[:, 2]] - uv[idx[:, 0]]
        det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
        det = np.where(np.abs(det) < 1e-20, 1.0, det)
        tt = (e1 * d2[:, 1:2] - e2 * d1[:, 1:2]) / det[:, None]
        bb = (e2 * d1[:, 0:1] - e1 * d2[:, 0:1]) / det[:, None]
        for k in range(3):
            np.add.at(tan, idx[:, k], tt)
            np.add.at(bit, idx[:, k], bb)
    tan = tan - nrm * np.sum(tan * nrm, axis=1, keepdims=True)
    ln = np.linalg.norm(tan, axis=1, keepdims=True)
    fallback = np.cross(nrm, np.array([0.0, 1.0, 0.0])[None, :])
    fb2 = np.cross(nrm, np.array([1.0, 0.0, 0.0])[None, :])
    fallback = np.where(np.linalg.norm(fallback, axis=1, keepdims=True) < 1e-6, fb2, fallback)
    tan = np.where(ln < 1e-12, fallback, tan)
    tan = tan / np.maximum(np.linalg.norm(tan, axis=1, keepdims=True), 1e-30)
    hand = np.where(np.sum(np.cross(nrm, tan) * bit, axis=1) < 0.0, -1.0, 1.0)
    return np.concatenate([tan, hand[:, None]], axis=1).astype(np.float32)


def textured_cornell(wall_quads=12, box_quads=4, gbuffer_maps=False):
    """The Cornell box with textured materials (texture indices into procedural_textures()): textured floor (base colour),
    alpha-tested back wall (occlusion mask: holes), a wall whose base colour uses the 3 x 5 texture with repeated
    coordinates, an emissive quad with an emissive texture, one box that combines all three. gbuffer_maps=True adds what only
    the G-buffer pass reads: a normal map + metallic-roughness map on the right wall, a normal map on the box, an alpha cutoff
    on the floor (its base-colour texture has alpha 0.6 / 1 in a checker) and per-vertex tangents."""
    b = _Builder()
    node = b.node(np.eye(4, dtype=np.float32).reshape(4, 4))

    def mat(base=(1, 1, 1, 1), emissive=(0, 0, 0), bt=-1, et=-1, ot=-1, **kw):
        i = b.material(base, emissive=emissive, **kw)
        b.mats[i]["base_color_texture"] = bt
        b.mats[i]["emissive_texture"] = et
        b.mats[i]["occlusion_texture"] = ot
        return i

    floor = mat((0.9, 0.9, 0.9, 1.0), bt=0)
    back = mat((0.725, 0.725, 0.725, 1.0), ot=2)
    left = mat((0.63, 0.3, 0.25, 0.9), bt=3)
    right = mat((0.14, 0.45, 0.091, 1.0))
    combo = mat((0.8, 0.8, 0.8, 1.0), bt=0, ot=2)
    light = mat((1.0, 1.0, 1.0, 1.0), emissive=(0.6, 0.7, 0.8), et=1)
    n = wall_quads
    b.grid((-4, -4, -4), (8, 0, 0), (0, 0, 8), n, n, (0, 1, 0), floor, node)
    b.grid((-4, -4, -4), (8, 0, 0), (0, 8, 0), n, n, (0, 0, 1), back, node)
    b.grid((-4, -4, -4), (0, 0, 8), (0, 8, 0), n, n, (1, 0, 0), left, node)
    b.grid((4, -4, -4), (0, 8, 0), (0, 0, 8), n, n, (-1, 0, 0), right, node)
    _box(b, (1.7, -2.8, -1.7), (1.2, 1.2, 1.2), 18.0, box_quads, combo, node)
    b.grid((-1, 3.6, -2.5), (2, 0, 0), (0, 0, 2), 2, 2, (0, -1, 0), light, node)
    sc = b.build("textured_cornell")
    # texture coordinates beyond [0, 1] and negative on the left wall (primitive 2): REPEAT addressing
    v0 = int(sc.primitives[2]["vertex_offset"])
    v1 = int(sc.primitives[3]["vertex_offset"])
    sc.texcoords[v0:v1] = (sc.texcoords[v0:v1] * np.float32(3.3) - np.float32(1.2)).astype(np.float32)
    if gbuffer_maps:
        sc.materials[right]["normal_texture"] = 4
        sc.materials[right]["metallic_roughness_texture"] = 5
        sc.materials[right]["metallic_factor"] = 0.9
        sc.materials[right]["roughness_factor"] = 0.8
        sc.materials[combo]["normal_texture"] = 4
        sc.materials[floor]["alpha_mode"] = 1
        sc.materials[floor]["alpha_cutoff"] = 0.8
        sc.tangents = compute_tangents(sc)
    return sc


def atrium(seed=1234):
    """Config 2: "Sponza-scale" procedural atrium, exactly 262 144 triangles inside the
    Sponza world bounding box [-15.4,-1.0,-9.5]..[14.4,11.4,8.8]: floor, walls,
    two storeys of colonnades with arches, balconies, two barrel vaults, hanging banners; 25
    factor-only materials (base colour ~ U(0.2,0.9)^3), 2 of them emissive."""
    rng = np.random.RandomState(seed)
    b = _Builder()
    ident = b.node(np.eye(4, dtype=np.float32))
    # a Sponza-like uniformly scaled node (glTF node scale 0.008): geometry authored x125
    scaled = b.node(glm.scale(0.008))
    banner_node = b.node(glm.mul(glm.translate((0.3, 0.0, -0.2)), glm.rotate_y(np.radians(7.0))))
    mats = []
    for i in range(25):
        col = rng.uniform(0.2, 0.9, 3)
        if i in (23, 24):
            mats.append(b.material((*col, 1.0), emissive=tuple(rng.uniform(0.6, 1.0, 3))))
        elif i % 6 == 5:
            mats.append(b.material((*col, 1.0), metallic=1.0, roughness=float(rng.uniform(0.15, 0.5))))
        else:
            mats.append(b.material((*col, 1.0), metallic=0.0, roughness=float(rng.uniform(0.5, 1.0))))
    x0, x1, y0, y1, z0, z1 = -15.4, 14.4, -1.0, 11.4, -9.5, 8.8
    # floor 128x64 quads (16 384 tris)
    b.grid((x0, y0, z0), (x1 - x0, 0, 0), (0, 0, z1 - z0), 128, 64, (0, 1, 0), mats[0], ident)
    # 4 walls 64x32 quads each (16 384 tris)
    b.grid((x0, y0, z0), (x1 - x0, 0, 0), (0, y1 - y0, 0), 64, 32, (0, 0, 1), mats[1], ident)
    b.grid((x0, y0, z1), (0, y1 - y0, 0), (x1 - x0, 0, 0), 32, 64, (0, 0, -1), mats[2], ident)
    b.grid((x0, y0, z0), (0, 0, z1 - z0), (0, y1 - y0, 0), 64, 32, (1, 0, 0), mats[3], ident)
    b.grid((x1, y0, z0), (0, y1 - y0, 0), (0, 0, z1 - z0), 32, 64, (-1, 0, 0), mats[4], ident)
    # colonnades: 2 storeys x 2 rows x 12 columns, 32 seg x 32 stacks (98 304 tris)
    xs = np.linspace(-12.5, 11.5, 12)
    for storey, (yb, hgt) in enumerate(((y0, 4.4), (y0 + 5.0, 4.0))):
        for row, zc in enumerate((-4.6, 3.9)):
            for ci, xc in enumerate(xs):
                b.cylinder((xc, yb, zc), 0.32 - 0.06 * storey, hgt, 32, 32, mats[5 + (ci + row + storey) % 6], ident)
    # arches between columns: 2 x 2 x 11, 32 seg x 16 across (45 056 tris)
    for storey, yb in enumerate((y0 + 3.2, y0 + 8.0)):
        for row, zc in enumerate((-4.6, 3.9)):
            for ci in range(11):
                b.arch((xs[ci], yb, zc), (1, 0, 0), xs[ci + 1] - xs[ci], 1.15, (0, 0, 1), 0.7, 32, 16,
                       mats[11 + (ci + row) % 4], ident)
    # balconies: 2 slabs 128x16 quads (8 192 tris)
    b.grid((x0, y0 + 4.7, z0), (x1 - x0, 0, 0), (0, 0, 4.9), 128, 16, (0, -1, 0), mats[15], ident)
    b.grid((x0, y0 + 4.7, 3.9), (x1 - x0, 0, 0), (0, 0, z1 - 3.9), 128, 16, (0, -1, 0), mats[16], ident)
    # banners: 16 hanging quads 16x32 (16 384 tris), two emissive; authored in a rotated node
    for i in range(16):
        xc = -13.0 + i * 1.7
        zc = -1.2 + 1.9 * np.sin(i * 1.3)
        m = mats[23 + (i // 8)] if i in (3, 11) else mats[17 + i % 6]
        b.grid((xc, 3.0 + 0.5 * np.cos(i), zc), (0.9, 0, 0.25), (0, 3.4, 0), 16, 32,
               glm.normalize((-0.25, 0.0, 0.9)), m, banner_node)
    # two barrel vaults 128 x 120 quads each (61 440 tris), authored x125 in the scaled node

    def vault(zc, half, mat):
        def disp(p, n, s, t):
            ang = np.pi * t
            p = p.copy()
            p[..., 2] = (zc - half * np.cos(ang)) * 125.0
            p[..., 1] = (y1 - 2.6 + 2.5 * np.sin(ang)) * 125.0
            n = np.stack([np.zeros_like(ang), -np.sin(ang), np.cos(ang)], axis=-1)
            return p, n
        b.grid((x0 * 125.0, 0, 0), ((x1 - x0) * 125.0, 0, 0), (0, 0, 1), 128, 120, (0, -1, 0), mat, scaled, displace=disp)
    vault(-4.9, 4.5, mats[21])
    vault(4.3, 4.4, mats[22])
    scene = b.build("atrium")
    assert scene.triangle_count == 262144, scene.triangle_count
    return scene


def textured_atrium(seed=1234):
    """The atrium with textured materials (texture indices into procedural_textures()): the same 262 144 triangles, texture
    coordinates repeated 6-24 times per surface; floor, two walls, half of the columns and the arches take a base-colour
    texture, the banners an occlusion mask (alpha test: the holes change the occupancy), the emissive banners an emissive
    texture; for the G-buffer pass the side walls carry the normal map, the metallic columns the metallic-roughness map and
    the floor an alpha cutoff-free base colour; per-vertex tangents included."""
    sc = atrium(seed)
    m = sc.materials
    rep = np.ones(len(m), np.float32)

    def put(i, r, **kw):
        for k, v in kw.items():
            m[i][k] = v
        rep[i] = r
    put(0, 24.0, base_color_texture=0)
    put(1, 8.0, base_color_texture=3)
    put(2, 8.0, base_color_texture=0)
    put(3, 6.0, normal_texture=4)
    put(4, 6.0, normal_texture=4, base_color_texture=3)
    for i in range(5, 11):
        if i % 2:
            put(i, 6.0, base_color_texture=0)
        if m[i]["metallic_factor"] > 0.5:
            put(i, 6.0, metallic_roughness_texture=5)
    for i in range(11, 15):
        put(i, 4.0, base_color_texture=0)
    for i in range(17, 23):
        put(i, 3.0, occlusion_texture=2)
    put(23, 2.0, emissive_texture=1)
    put(24, 2.0, emissive_texture=1)
    for pr in sc.primitives:
        v0 = int(pr["vertex_offset"])
        idx = sc.indices[int(pr["first_index"]):int(pr["first_index"]) + int(pr["index_count"])]
        v1 = v0 + int(idx.max()) + 1
        sc.texcoords[v0:v1] = (sc.texcoords[v0:v1] * rep[int(pr["material_index"])]).astype(np.float32)
    sc.tangents = compute_tangents(sc)
    sc.name = "textured_atrium"
    return sc


def quad_scene(corners, normal=(0, 0, 1), base=(1, 1, 1, 1), emissive=(0, 0, 0)):
    """Two-triangle quad for known-answer tests."""
    b = _Builder()
    node = b.node(np.eye(4, dtype=np.float32))
    m = b.material(base, emissive=emissive)
    c = np.asarray(corners, dtype=np.float64)
    b.grid(c[0], c[1] - c[0], c[3] - c[0], 1, 1, normal, m, node)
    return b.build("quad")


def triangle_soup(tris, normals=None, base=(1, 1, 1, 1)):
    """Scene from an explicit (T,3,3) array of world triangles (identity node)."""
    tris = np.asarray(tris, dtype=np.float32)
    T = tris.shape[0]
    pos = tris.reshape(-1, 3)
    if normals is None:
        n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
        ln = np.linalg.norm(n, axis=1, keepdims=True)
        n = np.where(ln > 0, n / np.maximum(ln, 1e-30), np.array([0, 0, 1.0]))
        normals = np.repeat(n, 3, axis=0)
    b = _Builder()
    node = b.node(np.eye(4, dtype=np.float32))
    m = b.material(base)
    b.pos.append(pos)
    b.nrm.append(np.asarray(normals, dtype=np.float32).reshape(-1, 3))
    b.uv.append(np.zeros((3 * T, 2), dtype=np.float32))
    b.idx.append(np.arange(3 * T, dtype=np.uint32))
    b.prims.append((0, 3 * T, 0, m, node))
    return b.build("soup")


# ---- camera / light -------------------------------------------------------------------------------

def make_camera(position=(0.0, 0.0, 0.0), direction=(0.0, 0.0, -1.0), aspect=1.0, fovy_deg=60.0):
    """Reference camera (VFS/Camera.h:48-56, Camera.cpp:110-121): up (0,-1,0), near 0.01, far 5000."""
    vp, vpi, eye = glm.camera_ubo(position, direction, (0.0, -1.0, 0.0), fovy_deg, aspect)
    cam = Camera()
    cam.view_proj[:] = vp.reshape(16).tolist()
    cam.view_proj_inv[:] = vpi.reshape(16).tolist()
    cam.eye_pos[:] = eye.tolist()
    cam.padding = 0
    return cam


def make_light(origin=(0.0, 30.0, -5.3), direction=(0.0, -1.0, 0.2), color=(1.0, 1.0, 1.0), intensity=1.0,
               half=16.0, z_near=0.1, z_far=30.0):
    """Reference directional light (VFS/Application.cpp:386-392, DirectionalLight.cpp:21-47)."""
    view, proj, d = glm.light_shadow_desc(origin, direction, half, z_near, z_far)
    light = DirLight()
    light.direction[:] = d.tolist()
    light.intensity = intensity
    light.color[:] = list(color)
    light.padding = 0
    sh = DirLightShadow()
    sh.view[:] = view.reshape(16).tolist()
    sh.proj[:] = proj.reshape(16).tolist()
    sh.z_near = z_near
    sh.z_far = z_far
    return light, sh


def struct_bytes(s):
    return bytes(memoryview(s))


def orbit_cameras(n, center, aspect, seed=7):
    """Config 5: n cameras on a seeded orbit (radius 6-12, height 2-8) looking at `center`."""
    rng = np.random.RandomState(seed)
    cams = []
    for i in range(n):
        ang = 2 * np.pi * i / n + rng.uniform(-0.05, 0.05)
        r = rng.uniform(6.0, 12.0)
        pos = np.array([center[0] + r * np.cos(ang), rng.uniform(2.0, 8.0), center[2] + 0.55 * r * np.sin(ang)])
        d = np.asarray(center, dtype=np.float64) - pos
        cams.append(make_camera(pos, d / np.linalg.norm(d), aspect))
    return cams
