// vgi_svo.cu — sparse voxel octree path for sm_100a: one-pass fragment list and a scan-based octree build.
//
// ref: SparseVoxelizer::preVoxelize / cmdVoxelize (VFS/RenderPass/Octree/SparseVoxelizer.cpp:248-326) with
// voxelizer.{vert,geom,frag}, and OctreeBuilder::cmdBuild (VFS/RenderPass/Octree/OctreeBuilder.cpp:205-345)
// with octreeNode{Init,Flag,Alloc,ModifyArg,LeafWrite,MipmapWrite}.comp.
//
// The reference rasterises the scene twice (count, then fill, with a host read-back in between) and builds
// the tree with ~4 dependent dispatches per level that chase pointers from the root for every fragment.
// Here:
//   fragments  k_svo_voxelize emits (triangle, voxel) pairs with the same conservative overlap test as the
//              clipmap voxelizer; k_svo_shade shades them (quad-cooperative shadow taps) and appends the
//              packed uvec2 fragments with a warp-ballot compaction. One pass, no host round trip.
//   octree     every fragment sets one bit in a dense mask indexed by its descent path (3 bits per level in
//              the shader's child-slot order z | x<<1 | y<<2, octreeNodeFlag.comp:40). OR-reducing bytes to
//              bits gives the occupied cells of every depth; ONE exclusive popcount scan over the
//              concatenated masks (depth 1 first) is exactly the order in which the reference's atomic
//              counter would hand out child blocks if its invocations ran sequentially, so child pointers
//              are closed-form and the pool equals the sequential oracle's bit for bit (stronger than the
//              "equal after canonical child ordering" bar). Colours: integer sums per leaf, then one
//              bottom-up pass per depth; every parent writes its 8 children as one 64-byte store.
// Compiled with -fmad=false (fragment positions and colours are quantised results).
#include "vgi_device.cuh"

struct SvoGrid {
    float center[3];
    float extentValue;
    uint32_t res;
    int level;
};

typedef unsigned long long svo_pair_t; // tri << 36 | z << 24 | y << 12 | x

// ref: voxelizer.vert:47 ((world - center) / extentValue), voxelizer.geom:27-30 ((v + 1) * 0.5),
// voxelizer.frag:95 (position * resolution)
DEVFN void svo_setup(const SvoGrid& g, const float p[9], TriSetup& ts, float N[3], int* axis)
{
    float ndc[9], q[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ndc[i * 3 + k] = (p[i * 3 + k] - g.center[k]) / g.extentValue;
            q[i][k] = ((ndc[i * 3 + k] + 1.0f) * 0.5f) * (float)g.res;
        }
    float Nndc[3];
    *axis = cross_and_axis(ndc, Nndc);  // voxelizer.geom:41-43: dominant axis from the normalised positions
    cross_and_axis(p, N);               // world-space plane for the sample point
    const int lo[3] = { 0, 0, 0 };
    const int hi[3] = { (int)g.res, (int)g.res, (int)g.res }; // voxelizer.frag:95-97 clamps to [0, res] INCLUSIVE (Q13)
    tri_setup_grid(ts, q, N, lo, hi);
}

DEVFN void svo_emit_pair(svo_pair_t* __restrict__ pairs, Counters* __restrict__ cnt, uint32_t max_pairs, uint32_t tri, int x, int y, int z)
{
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    const unsigned rank = __popc(m & ((1u << lane_id()) - 1u));
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(&cnt->pairs, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t slot = base + rank;
    if (slot < max_pairs)
        pairs[slot] = ((svo_pair_t)tri << 36) | ((svo_pair_t)(uint32_t)z << 24) | ((svo_pair_t)(uint32_t)y << 12) | (uint32_t)x;
    else
        atomicOr(&cnt->overflow, 1u);
}

#define SVO_SMALL_BOX_MAX 64

__global__ void __launch_bounds__(128) k_svo_voxelize(SvoGrid g, uint32_t ntri, const float4* __restrict__ tri_pos,
                                                       svo_pair_t* __restrict__ pairs, uint32_t max_pairs,
                                                       uint2* __restrict__ large, uint32_t max_large, Counters* __restrict__ cnt)
{
    // same scheme as k_voxelize: 64-bit hit mask per thread, one reservation per warp
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long hits = 0ull;
    int lo0 = 0, lo1 = 0, lo2 = 0, nx = 1, ny = 1;
    if (t < ntri) {
        float p[9], N[3];
        int axis;
        load_tri(tri_pos, t, p, nullptr);
        TriSetup ts;
        svo_setup(g, p, ts, N, &axis);
        if (ts.valid && ts.lo[0] <= ts.hi[0] && ts.lo[1] <= ts.hi[1] && ts.lo[2] <= ts.hi[2]) {
            nx = ts.hi[0] - ts.lo[0] + 1;
            ny = ts.hi[1] - ts.lo[1] + 1;
            const long long vol = (long long)nx * ny * (ts.hi[2] - ts.lo[2] + 1);
            if (vol > SVO_SMALL_BOX_MAX) {
                const uint32_t slot = atomicAdd(&cnt->large, 1u);
                if (slot < max_large) large[slot] = make_uint2(t, 0u);
                else atomicOr(&cnt->overflow, 2u);
            } else {
                lo0 = ts.lo[0]; lo1 = ts.lo[1]; lo2 = ts.lo[2];
                int i = 0;
                for (int z = ts.lo[2]; z <= ts.hi[2]; ++z)
                    for (int y = ts.lo[1]; y <= ts.hi[1]; ++y)
                        for (int x = ts.lo[0]; x <= ts.hi[0]; ++x, ++i)
                            if (tri_overlaps_voxel(ts, x, y, z)) hits |= 1ull << i;
            }
        }
    }
    const uint32_t n = (uint32_t)__popcll(hits);
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (!total) return;
    uint32_t base = 0u;
    if (lane_id() == 31u) base = atomicAdd(&cnt->pairs, total);
    uint32_t slot = __shfl_sync(0xffffffffu, base, 31) + (incl - n);
    for (unsigned long long h = hits; h; h &= h - 1, ++slot) {
        const int i = __ffsll((long long)h) - 1;
        const uint32_t x = (uint32_t)(lo0 + i % nx), y = (uint32_t)(lo1 + (i / nx) % ny), z = (uint32_t)(lo2 + i / (nx * ny));
        if (slot < max_pairs) pairs[slot] = ((svo_pair_t)t << 36) | ((svo_pair_t)z << 24) | ((svo_pair_t)y << 12) | x;
        else atomicOr(&cnt->overflow, 1u);
    }
}

__global__ void __launch_bounds__(256) k_svo_voxelize_large(SvoGrid g, const float4* __restrict__ tri_pos,
                                                             svo_pair_t* __restrict__ pairs, uint32_t max_pairs,
                                                             const uint2* __restrict__ large, uint32_t max_large,
                                                             Counters* __restrict__ cnt)
{
    const uint32_t nitems = min(cnt->large, max_large);
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < nitems; item += warpsPerGrid) {
        const uint32_t t = large[item].x;
        float p[9], N[3];
        int axis;
        load_tri(tri_pos, t, p, nullptr);
        TriSetup ts;
        svo_setup(g, p, ts, N, &axis);
        const int nx = ts.hi[0] - ts.lo[0] + 1, ny = ts.hi[1] - ts.lo[1] + 1, nz = ts.hi[2] - ts.lo[2] + 1;
        const long long vol = (long long)nx * ny * nz;
        for (long long base = 0; base < vol; base += 32) {
            const long long i = base + lane_id();
            if (i < vol) {
                const int x = ts.lo[0] + (int)(i % nx);
                const int y = ts.lo[1] + (int)((i / nx) % ny);
                const int z = ts.lo[2] + (int)(i / ((long long)nx * ny));
                if (tri_overlaps_voxel(ts, x, y, z)) svo_emit_pair(pairs, cnt, max_pairs, t, x, y, z);
            }
        }
    }
}

// ref: voxelizer.frag:48-103 (canonical Q21/Q22; literal behind VGI_MODE_SVO_LITERAL)
__global__ void __launch_bounds__(256) k_svo_shade(SvoGrid g, LightParams lp, int compare_i, int literal,
                                                    const float4* __restrict__ tri_pos, const float4* __restrict__ tri_nrm,
                                                    const vgi_material* __restrict__ materials,
                                                    const svo_pair_t* __restrict__ pairs, uint32_t max_pairs,
                                                    uint2* __restrict__ frags, uint32_t max_frags, Counters* __restrict__ cnt,
                                                    TexSet tex)
{
    const uint32_t npairs = min(cnt->pairs, max_pairs);
    const uint32_t stride = gridDim.x * blockDim.x;
    const unsigned lane = lane_id();
    const bool compare = compare_i != 0;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < npairs; base += stride) {
        const uint32_t i = base + lane;
        int kind = 0; // 0 nothing, 1 emissive, 2 lit
        float spx = 0.0f, spy = 0.0f, scz = 0.0f, NdotL = 0.0f;
        int mat = 0;
        uint32_t vx = 0, vy = 0, vz = 0;
        float uv[2] = { 0.0f, 0.0f };
        if (i < npairs) {
            const svo_pair_t pr = pairs[i];
            const uint32_t tri = (uint32_t)(pr >> 36);
            vx = (uint32_t)(pr & 0xfffu); vy = (uint32_t)((pr >> 12) & 0xfffu); vz = (uint32_t)((pr >> 24) & 0xfffu);
            float p[9], n9[9], N[3];
            load_tri(tri_pos, tri, p, &mat);
            {
                const float4 a = __ldg(tri_nrm + 3 * (size_t)tri), b = __ldg(tri_nrm + 3 * (size_t)tri + 1), c = __ldg(tri_nrm + 3 * (size_t)tri + 2);
                n9[0] = a.x; n9[1] = a.y; n9[2] = a.z; n9[3] = b.x; n9[4] = b.y; n9[5] = b.z; n9[6] = c.x; n9[7] = c.y; n9[8] = c.z;
            }
            float ndc[9], Nndc[3];
#pragma unroll
            for (int v = 0; v < 3; ++v)
#pragma unroll
                for (int k = 0; k < 3; ++k) ndc[v * 3 + k] = (p[v * 3 + k] - g.center[k]) / g.extentValue;
            const int axis = cross_and_axis(ndc, Nndc);
            cross_and_axis(p, N);
            // voxel centre back in world space
            const uint32_t v3[3] = { vx, vy, vz };
            float c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = ((((float)v3[k] + 0.5f) / (float)g.res) * 2.0f - 1.0f) * g.extentValue + g.center[k];
            float pos[3], nrm[3], bary[3];
            bool sampled = inject_sample_at(axis, N, p, n9, c, pos, nrm, bary);
            if (sampled && tex.count) {
                const vgi_material* mt = materials + mat;
                if (mt->base_color_texture > -1 || mt->emissive_texture > -1 || mt->occlusion_texture > -1) {
                    tri_uv_at(tex, tri, bary, uv);
                    if (mt->occlusion_texture > -1) {       // ref: voxelizer.frag:52
                        float t[4];
                        tex_fetch(tex, mt->occlusion_texture, uv[0], uv[1], t);
                        if (t[0] < 0.1f) sampled = false;
                    }
                }
            }
            if (sampled) {
                const vgi_material* m = materials + mat;
                if (m->emissive_factor[0] > 0.0f || m->emissive_factor[1] > 0.0f || m->emissive_factor[2] > 0.0f) {
                    kind = 1;
                } else {
                    const float len2 = dot3(nrm, nrm);
                    if (len2 > 0.0f) {
                        const float len = sqrtf(len2);
                        const float n[3] = { nrm[0] / len, nrm[1] / len, nrm[2] / len };
                        NdotL = f_clamp(dot3(n, lp.dir_to_light), 0.001f, 1.0f);
                        float vpos[3] = { pos[0], pos[1], pos[2] };
                        if (literal) { // Q22: calcVisibility is fed the normalised [0,1] position
#pragma unroll
                            for (int k = 0; k < 3; ++k) vpos[k] = (((pos[k] - g.center[k]) / g.extentValue) + 1.0f) * 0.5f;
                        }
                        float l[3];
                        xform_point(lp.view, vpos, l);
                        const float* P = lp.proj;
                        const float qx = ((P[0] * l[0] + P[4] * l[1]) + P[8] * 0.0f) + P[12];
                        const float qy = ((P[1] * l[0] + P[5] * l[1]) + P[9] * 0.0f) + P[13];
                        spx = qx * 0.5f + 0.5f;
                        spy = qy * 0.5f + 0.5f;
                        if (compare) scz = (P[10] * l[2] + P[14]) - 0.002f;
                        kind = 2;
                    }
                }
            }
        }
        const unsigned lit = __ballot_sync(0xffffffffu, kind == 2);
        (void)lit;
        const float vis = kind == 2 ? lane_visibility(lp, spx, spy, scz, compare) : 0.0f;
        float radiance[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
        bool emit = kind != 0;
        if (kind == 1) {
            const vgi_material* m = materials + mat;
            float em[3] = { m->emissive_factor[0], m->emissive_factor[1], m->emissive_factor[2] };
            if (tex.count && m->emissive_texture > -1) {    // ref: voxelizer.frag:63-67
                float t[4];
                tex_fetch(tex, m->emissive_texture, uv[0], uv[1], t);
                em[0] = em[0] + t[0]; em[1] = em[1] + t[1]; em[2] = em[2] + t[2];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) radiance[k] = f_clamp(em[k], 0.0f, 1.0f);
        } else if (kind == 2) {
            const vgi_material* m = materials + mat;
            float color[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
            if (tex.count && m->base_color_texture > -1) {  // ref: voxelizer.frag:72-77 — texture * factor (both modes)
                float t[4];
                tex_fetch(tex, m->base_color_texture, uv[0], uv[1], t);
#pragma unroll
                for (int k = 0; k < 4; ++k) color[k] = t[k] * m->base_color_factor[k];
            } else if (!literal) { color[0] = m->base_color_factor[0]; color[1] = m->base_color_factor[1]; color[2] = m->base_color_factor[2]; color[3] = m->base_color_factor[3]; } // Q21
            float lc[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) lc[k] = ((NdotL * vis) * lp.color[k]) * lp.intensity;
            if (lc[0] == 0.0f && lc[1] == 0.0f && lc[2] == 0.0f) emit = false; // discard, voxelizer.frag:89-90
#pragma unroll
            for (int k = 0; k < 3; ++k) radiance[k] = f_clamp((lc[k] * color[k]) * color[3], 0.0f, 1.0f);
        }
        // warp-ballot compaction of the surviving fragments
        const unsigned em = __ballot_sync(0xffffffffu, emit);
        if (!em) continue;
        uint32_t slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&cnt->svo_frags, (uint32_t)__popc(em));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (emit) {
            const uint32_t slot = slot0 + __popc(em & ((1u << lane) - 1u));
            if (slot < max_frags) {
                // ref: voxelizer.frag:99-110 — uint() truncation of radiance * 255, & 0xff per channel
                const uint32_t rgba = (((uint32_t)(radiance[3] * 255.0f) & 0xffu) << 24) | (((uint32_t)(radiance[2] * 255.0f) & 0xffu) << 16) |
                                      (((uint32_t)(radiance[1] * 255.0f) & 0xffu) << 8) | ((uint32_t)(radiance[0] * 255.0f) & 0xffu);
                frags[slot] = make_uint2(vx | (vy << 12) | ((vz & 0xffu) << 24), ((vz >> 8) << 28) | (rgba & 0x0fffffffu));
            } else {
                atomicOr(&cnt->overflow, 8u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// octree build
// ---------------------------------------------------------------------------------------------------
// descent path of a fragment (octreeNodeFlag.comp:28-43): position halved (Q13), then one child slot
// z | x<<1 | y<<2 per level, most significant level first -> bit b of (fz, fx, fy) lands at key bit 3b+{0,1,2}
DEVFN uint32_t svo_path_key(uint2 frag, int level)
{
    const uint32_t fx = (frag.x & 0xfffu) >> 1, fy = ((frag.x >> 12) & 0xfffu) >> 1;
    const uint32_t fz = (((frag.x >> 24) & 0xffu) | ((frag.y >> 20) & 0xf00u)) >> 1;
    uint32_t key = 0u;
    for (int b = 0; b < level; ++b)
        key |= (((fz >> b) & 1u) << (3 * b)) | (((fx >> b) & 1u) << (3 * b + 1)) | (((fy >> b) & 1u) << (3 * b + 2));
    return key;
}

struct SvoLayout {
    // masks of depths 1..level concatenated in one word array; depth d has 8^d bits starting at word off[d]
    uint32_t off[13];
    uint32_t total_words;
    int level;
};

__global__ void __launch_bounds__(256) k_svo_mark(SvoLayout lay, const uint2* __restrict__ frags, const Counters* __restrict__ cnt,
                                                   uint32_t max_frags, uint32_t* __restrict__ mask)
{
    const uint32_t n = min(cnt->svo_frags, max_frags);
    uint32_t* leaf = mask + lay.off[lay.level];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = svo_path_key(frags[i], lay.level);
        const uint32_t bit = 1u << (key & 31u);
        if (!(leaf[key >> 5] & bit)) atomicOr(&leaf[key >> 5], bit);
    }
}

// depth d from depth d+1: bit j of M_d = (byte j of M_{d+1} != 0)
__global__ void __launch_bounds__(256) k_svo_pyramid(SvoLayout lay, int d, uint32_t* __restrict__ mask)
{
    const uint32_t nbits = 1u << (3 * d);
    const uint32_t nwords = (nbits + 31u) >> 5;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwords) return;
    const uint4* src = reinterpret_cast<const uint4*>(mask + lay.off[d + 1]) + 2 * (size_t)j; // 32 bytes = 32 child bytes
    const uint4 a = src[0], b = src[1];
    const uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
    uint32_t out = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if ((w[k] >> (8 * t)) & 0xffu) out |= 1u << (4 * k + t);
    }
    if (nbits < 32u) out &= (1u << nbits) - 1u;
    mask[lay.off[d] + j] = out;
}

DEVFN uint32_t svo_rank(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ prefix, uint32_t word, uint32_t bit)
{
    return prefix[word] + __popc(mask[word] & ((1u << bit) - 1u));
}

// per-leaf integer colour sums (canonical exact mean, Q10): acc[leaf rank] = {sum r, sum g, sum b, count}
__global__ void __launch_bounds__(256) k_svo_leaf_acc(SvoLayout lay, const uint2* __restrict__ frags, const Counters* __restrict__ cnt,
                                                       uint32_t max_frags, const uint32_t* __restrict__ mask,
                                                       const uint32_t* __restrict__ prefix, uint32_t* __restrict__ acc)
{
    const uint32_t n = min(cnt->svo_frags, max_frags);
    const uint32_t leafBase = prefix[lay.off[lay.level]];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 f = frags[i];
        const uint32_t key = svo_path_key(f, lay.level);
        const uint32_t r = svo_rank(mask, prefix, lay.off[lay.level] + (key >> 5), key & 31u) - leafBase;
        const uint32_t c = f.y & 0x0fffffffu;
        uint32_t* a = acc + (size_t)r * 4;
        atomicAdd(a + 0, c & 0xffu);
        atomicAdd(a + 1, (c >> 8) & 0xffu);
        atomicAdd(a + 2, (c >> 16) & 0xffu);
        atomicAdd(a + 3, 1u);
    }
}

__global__ void k_svo_zero_acc(uint32_t* __restrict__ acc, SvoLayout lay, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ total)
{
    const size_t n = (size_t)(*total - prefix[lay.off[lay.level]]); // uint4 per occupied leaf
    uint4* a4 = reinterpret_cast<uint4*>(acc);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a4[i] = make_uint4(0, 0, 0, 0);
}

// colour of every occupied cell of depth d into col[global rank]; leaves from the sums, interior cells from
// their children: node.y = uint(sum_j children[j].y / 8) per channel (octreeNodeMipmapWrite.comp:54-60) — the
// float sum of eight bytes / 8 is exact, so it is an integer shift.
__global__ void __launch_bounds__(256) k_svo_colors(SvoLayout lay, int d, int min_mip_depth, const uint32_t* __restrict__ mask,
                                                     const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ acc,
                                                     uint32_t* __restrict__ col)
{
    const uint32_t nbits = 1u << (3 * d);
    const uint32_t nwords = (nbits + 31u) >> 5;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwords) return;
    uint32_t w = mask[lay.off[d] + j];
    if (!w) return;
    const uint32_t base = prefix[lay.off[d] + j];
    const uint32_t leafBase = prefix[lay.off[lay.level]];
    uint32_t k = 0;
    while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        const uint32_t rank = base + k++;
        uint32_t out = 0u;
        if (d == lay.level) {
            const uint4 s = reinterpret_cast<const uint4*>(acc)[rank - leafBase];
            if (s.w) out = (s.x / s.w) | ((s.y / s.w) << 8) | ((s.z / s.w) << 16) | 0xff000000u;
        } else if (d >= min_mip_depth) {
            const uint32_t cell = j * 32u + (uint32_t)bit;
            // children = byte `cell` of the depth d+1 mask
            const uint32_t cw = lay.off[d + 1] + (cell >> 2);
            const uint32_t cbits = (mask[cw] >> (8 * (cell & 3u))) & 0xffu;
            uint32_t crank = prefix[cw] + __popc(mask[cw] & ((1u << (8 * (cell & 3u))) - 1u));
            uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (uint32_t m = cbits; m; m &= m - 1) {
                const uint32_t y = col[crank++];
                s0 += y & 0xffu; s1 += (y >> 8) & 0xffu; s2 += (y >> 16) & 0xffu; s3 += y >> 24;
            }
            out = (s0 >> 3) | ((s1 >> 3) << 8) | ((s2 >> 3) << 16) | ((s3 >> 3) << 24);
        }
        col[rank] = out;
    }
}

// every occupied cell of depth d-1 writes its block of 8 children (depth d) as one 64-byte store; the 8 nodes
// of depth 1 are the root's children (nodes 0..7). Node = {MSB flag | index of the first child, RGBA8}
// (octreeNodeInit.comp:4-7, octreeNodeAlloc.comp:30-32): child block of the flagged node with global rank r
// (flagged nodes counted depth by depth in node order = the sequential atomic counter) starts at (r + 1) << 3.
__global__ void __launch_bounds__(256) k_svo_emit(SvoLayout lay, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ prefix,
                                                   const uint32_t* __restrict__ col, uint2* __restrict__ nodes, uint32_t max_nodes,
                                                   Counters* __restrict__ cnt)
{
    // work item = one word of the parent masks (depths 0..level-1, depth 0 = the root = one virtual item)
    const uint32_t parentWords = lay.off[lay.level] - lay.off[1];
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item > parentWords) return;
    const uint32_t nnodes = 8u * (1u + prefix[lay.off[lay.level]]);
    if (item == 0) {
        cnt->svo_counter = nnodes;
        if (nnodes > max_nodes) atomicOr(&cnt->overflow, 16u);
    }
    if (nnodes > max_nodes) return;
    uint32_t w, parentRankBase, childMaskWordBase;
    int d; // depth of the children written by this item
    uint32_t j = 0;
    if (item == 0) {
        w = 1u; parentRankBase = 0u; d = 1; childMaskWordBase = lay.off[1];
    } else {
        const uint32_t word = lay.off[1] + (item - 1u);
        int pd = 1;
        while (pd + 1 < lay.level && word >= lay.off[pd + 1]) ++pd;
        d = pd + 1;
        j = word - lay.off[pd];
        w = mask[word];
        parentRankBase = prefix[word];
        childMaskWordBase = lay.off[d];
    }
    uint32_t k = 0;
    while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        const uint32_t cell = j * 32u + (uint32_t)bit;                       // parent cell key
        const uint32_t blockStart = item == 0 ? 0u : 8u * (1u + parentRankBase + k);
        ++k;
        const uint32_t cw = childMaskWordBase + (cell >> 2);
        const uint32_t sh = 8u * (cell & 3u);
        const uint32_t mw = mask[cw];
        const uint32_t cbits = (mw >> sh) & 0xffu;
        uint32_t crank = prefix[cw] + __popc(mw & ((1u << sh) - 1u));
        uint2 out[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            out[s] = make_uint2(0u, 0u);
            if ((cbits >> s) & 1u) {
                out[s].x = 0x80000000u | (d < lay.level ? (crank + 1u) << 3 : 0u);
                out[s].y = col[crank];
                ++crank;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(nodes + blockStart);
#pragma unroll
        for (int s = 0; s < 4; ++s) dst[s] = make_uint4(out[2 * s].x, out[2 * s].y, out[2 * s + 1].x, out[2 * s + 1].y);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int svo_fail(vgi_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    vgi_set_thread_error(msg);  // vgi_last_error(NULL) reads the string vgi_api.cpp owns
    return code;
}
#define SCK(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return svo_fail(c, e_ == cudaErrorMemoryAllocation ? VGI_E_NOMEM : VGI_E_CUDA,                   \
                            std::string(#call) + ": " + cudaGetErrorString(e_));                              \
    } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
#define LAUNCH(name, ...) do { c->timer.begin(name, s); __VA_ARGS__; c->timer.end(s); ++c->launches; } while (0)

static SvoLayout make_layout(int level)
{
    SvoLayout lay;
    memset(&lay, 0, sizeof lay);
    lay.level = level;
    uint32_t off = 0;
    for (int d = 1; d <= level; ++d) {
        lay.off[d] = off;
        const uint64_t words = ((1ull << (3 * d)) + 31ull) >> 5;
        off += (uint32_t)((words + 7ull) & ~7ull); // 32-byte granules (k_svo_pyramid reads 32 bytes per output word)
    }
    lay.off[level + 1] = off;
    lay.total_words = off;
    return lay;
}

static SvoGrid make_grid(const vgi_ctx* c)
{
    // ref: voxelizer.vert:42-47
    SvoGrid g;
    const float ex = c->svo_bb_max[0] - c->svo_bb_min[0], ey = c->svo_bb_max[1] - c->svo_bb_min[1], ez = c->svo_bb_max[2] - c->svo_bb_min[2];
    const float m = ex > ey ? (ex > ez ? ex : ez) : (ey > ez ? ey : ez);
    g.extentValue = m * 0.5f;
    for (int k = 0; k < 3; ++k) g.center[k] = (c->svo_bb_min[k] + c->svo_bb_max[k]) * 0.5f;
    g.level = (int)c->svo_level;
    g.res = 1u << c->svo_level;
    return g;
}

extern "C" {

int vgi_svo_voxelize(vgi_ctx* c, uint32_t level, const float bb_min[3], const float bb_max[3], void* stream)
{
    if (!c || !bb_min || !bb_max) return svo_fail(c, VGI_E_INVALID, "vgi_svo_voxelize: null argument");
    if (level < 1 || level > 10) return svo_fail(c, VGI_E_INVALID, "vgi_svo_voxelize: level must be in [1,10]"); // EngineConfig.h:19-21 allows 11; the dense path masks stop at 10
    if (!c->tri_pos && c->ntri) return svo_fail(c, VGI_E_STATE, "vgi_svo_voxelize: call vgi_set_scene first");
    if (!c->pairs) return svo_fail(c, VGI_E_STATE, "vgi_svo_voxelize: call vgi_set_scene first");
    if (c->scene_max_texture >= (int32_t)c->ntex)
        return svo_fail(c, VGI_E_STATE, "vgi_svo_voxelize: a material references a texture that vgi_set_textures has not provided");
    if (!c->light_set) return svo_fail(c, VGI_E_STATE, "vgi_svo_voxelize: call vgi_set_light first");
    for (int k = 0; k < 3; ++k)
        if (!(bb_max[k] >= bb_min[k])) return svo_fail(c, VGI_E_INVALID, "vgi_svo_voxelize: empty bounding box");
    SCK(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    c->svo_level = level;
    memcpy(c->svo_bb_min, bb_min, sizeof c->svo_bb_min);
    memcpy(c->svo_bb_max, bb_max, sizeof c->svo_bb_max);
    // the fragment list shares its capacity rule with the clipmap pair list (vgi_config.max_fragments)
    if (c->svo_frag_capacity < c->max_pairs) {
        SCK(cudaStreamSynchronize(c->last_stream));
        cudaFree(c->svo_frags);
        c->svo_frags = nullptr;
        SCK(cudaMalloc(&c->svo_frags, (size_t)c->max_pairs * sizeof(uint2)));
        c->svo_frag_capacity = c->max_pairs;
    }
    const SvoGrid g = make_grid(c);
    SCK(cudaMemsetAsync(c->counters, 0, sizeof(Counters), s));
    svo_pair_t* pairs = reinterpret_cast<svo_pair_t*>(c->pairs); // the clipmap pair buffer is free between builds
    if (c->ntri) {
        LAUNCH("k_svo_voxelize", k_svo_voxelize<<<cdiv(c->ntri, 128), 128, 0, s>>>(g, c->ntri, c->tri_pos, pairs, c->max_pairs, c->large, c->max_large, c->counters));
        LAUNCH("k_svo_voxelize_large", k_svo_voxelize_large<<<148 * 4, 256, 0, s>>>(g, c->tri_pos, pairs, c->max_pairs, c->large, c->max_large, c->counters));
        LAUNCH("k_svo_shade", k_svo_shade<<<148 * 8, 256, 0, s>>>(g, c->light, (c->cfg.mode_flags & VGI_MODE_SHADOW_COMPARE) ? 1 : 0,
                                                                   (c->cfg.mode_flags & VGI_MODE_SVO_LITERAL) ? 1 : 0, c->tri_pos, c->tri_nrm,
                                                                   c->materials, pairs, c->max_pairs, c->svo_frags, c->svo_frag_capacity, c->counters, c->texset()));
    }
    SCK(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    c->last_stream = s;
    c->voxelized = false; // the pair buffer was reused
    c->svo_voxelized = true;
    c->svo_built = false;
    c->svo_counters_fresh = true;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return svo_fail(c, VGI_E_CUDA, std::string("vgi_svo_voxelize: ") + cudaGetErrorString(e));
    return VGI_OK;
}

int vgi_svo_build(vgi_ctx* c, void* stream)
{
    if (!c) return svo_fail(c, VGI_E_INVALID, "vgi_svo_build: null ctx");
    if (!c->svo_voxelized) return svo_fail(c, VGI_E_STATE, "vgi_svo_build: call vgi_svo_voxelize first");
    SCK(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int level = (int)c->svo_level;
    const SvoLayout lay = make_layout(level);
    // node pool capacity: clamp(8 * Nfrag, 1e6, 5e8) (OctreeBuilder.cpp:110-113) needs the fragment count on the
    // host; the copy queued by vgi_svo_voxelize is complete once its stream is idle
    SCK(cudaStreamSynchronize(c->last_stream));
    if (c->h_counters->overflow) return svo_fail(c, VGI_E_OVERFLOW, "vgi_svo_build: fragment list overflow — raise vgi_config.max_fragments");
    c->svo_nfrag = c->h_counters->svo_frags;
    uint64_t cap = c->cfg.svo_max_nodes;
    if (!cap) {
        cap = (uint64_t)c->svo_nfrag << 3;
        if (cap < 1000000ull) cap = 1000000ull;
        if (cap > 500000000ull) cap = 500000000ull;
    }
    if (cap < 8) cap = 8;
    if (c->svo_node_capacity < cap) {
        cudaFree(c->svo_nodes);
        c->svo_nodes = nullptr;
        SCK(cudaMalloc(&c->svo_nodes, cap * sizeof(uint2)));
        c->svo_node_capacity = (uint32_t)cap;
    }
    // scratch: masks | prefix | block sums | leaf sums (4 u32 per fragment at most) | colours (1 u32 per occupied cell)
    const size_t nblk = (lay.total_words + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS);
    const size_t maxCells = (size_t)c->svo_nfrag * (size_t)level + 64;
    const size_t need = (size_t)lay.total_words * 2 + nblk + 64 + (size_t)c->svo_nfrag * 4 + 64 + maxCells;
    if (c->svo_scratch_words < need) {
        cudaFree(c->svo_scratch);
        c->svo_scratch = nullptr;
        SCK(cudaMalloc(&c->svo_scratch, need * sizeof(uint32_t)));
        c->svo_scratch_words = need;
    }
    uint32_t* mask = c->svo_scratch;
    uint32_t* prefix = mask + lay.total_words;
    uint32_t* bsums = prefix + lay.total_words;
    uint32_t* acc = bsums + ((nblk + 64 + 3) & ~(size_t)3);
    uint32_t* col = acc + (size_t)c->svo_nfrag * 4 + 64;
    SCK(cudaMemsetAsync(mask, 0, (size_t)lay.total_words * sizeof(uint32_t), s));
    const uint32_t nf = c->svo_nfrag;
    const unsigned fgrid = nf ? (cdiv(nf, 256) < 148u * 16u ? cdiv(nf, 256) : 148u * 16u) : 1u;
    LAUNCH("k_svo_mark", k_svo_mark<<<fgrid, 256, 0, s>>>(lay, c->svo_frags, c->counters, c->svo_frag_capacity, mask));
    for (int d = level - 1; d >= 1; --d) {
        const uint32_t nwords = (uint32_t)(((1ull << (3 * d)) + 31ull) >> 5);
        LAUNCH("k_svo_pyramid", k_svo_pyramid<<<cdiv(nwords, 256), 256, 0, s>>>(lay, d, mask));
    }
    LAUNCH("k_scan_block_sums", k_scan_block_sums<<<(unsigned)nblk, SCAN_BLOCK, 0, s>>>(mask, lay.total_words, bsums));
    LAUNCH("k_scan_sums", k_scan_sums<<<1, 1024, 0, s>>>(bsums, (uint32_t)nblk, &c->counters->svo_alloc_num));
    LAUNCH("k_scan_final", k_scan_final<<<(unsigned)nblk, SCAN_BLOCK, 0, s>>>(mask, lay.total_words, bsums, prefix));
    LAUNCH("k_svo_zero_acc", k_svo_zero_acc<<<148 * 4, 256, 0, s>>>(acc, lay, prefix, &c->counters->svo_alloc_num));
    LAUNCH("k_svo_leaf_acc", k_svo_leaf_acc<<<fgrid, 256, 0, s>>>(lay, c->svo_frags, c->counters, c->svo_frag_capacity, mask, prefix, acc));
    // literal Q13: octreeNodeMipmapWrite compares the resolution with the level INDEX, so only the rounds whose
    // index is a power of two write: depths level-1, level-2, level-3 (indices 2, 4, 8)
    int min_mip_depth = 1;
    if (c->cfg.mode_flags & VGI_MODE_SVO_LITERAL) {
        int rounds = 0;
        for (int i = 2; i <= level; i <<= 1) ++rounds;
        min_mip_depth = level - rounds;
        if (min_mip_depth < 1) min_mip_depth = 1;
    }
    for (int d = level; d >= 1; --d) {
        const uint32_t nwords = (uint32_t)(((1ull << (3 * d)) + 31ull) >> 5);
        LAUNCH("k_svo_colors", k_svo_colors<<<cdiv(nwords, 256), 256, 0, s>>>(lay, d, min_mip_depth, mask, prefix, acc, col));
    }
    const uint32_t items = lay.off[level] - lay.off[1] + 1u;
    LAUNCH("k_svo_emit", k_svo_emit<<<cdiv(items, 256), 256, 0, s>>>(lay, mask, prefix, col, c->svo_nodes, c->svo_node_capacity, c->counters));
    SCK(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    c->last_stream = s;
    c->svo_built = true;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return svo_fail(c, VGI_E_CUDA, std::string("vgi_svo_build: ") + cudaGetErrorString(e));
    return VGI_OK;
}

int vgi_svo_get_fragments(vgi_ctx* c, void** dev_ptr, uint32_t* count)
{
    if (!c) return svo_fail(c, VGI_E_INVALID, "vgi_svo_get_fragments: null ctx");
    if (!c->svo_voxelized) return svo_fail(c, VGI_E_STATE, "vgi_svo_get_fragments: call vgi_svo_voxelize first");
    SCK(cudaStreamSynchronize(c->last_stream));
    if (c->h_counters->overflow & (1u | 2u | 8u))
        return svo_fail(c, VGI_E_OVERFLOW, "vgi_svo_get_fragments: fragment list overflow — raise vgi_config.max_fragments");
    c->svo_nfrag = c->h_counters->svo_frags;
    if (dev_ptr) *dev_ptr = c->svo_frags;
    if (count) *count = c->svo_nfrag;
    return VGI_OK;
}

int vgi_svo_get_nodes(vgi_ctx* c, void** dev_ptr, uint32_t* count)
{
    if (!c) return svo_fail(c, VGI_E_INVALID, "vgi_svo_get_nodes: null ctx");
    if (!c->svo_built) return svo_fail(c, VGI_E_STATE, "vgi_svo_get_nodes: call vgi_svo_build first");
    SCK(cudaStreamSynchronize(c->last_stream));
    if (c->svo_counters_fresh) {    // otherwise a clipmap build has reused the counters since: the count cached by vgi_svo_build stands
        if (c->h_counters->overflow & 16u)
            return svo_fail(c, VGI_E_OVERFLOW, "vgi_svo_get_nodes: node pool overflow — raise vgi_config.svo_max_nodes");
        c->svo_nnodes = c->h_counters->svo_counter;
    }
    if (dev_ptr) *dev_ptr = c->svo_nodes;
    if (count) *count = c->svo_nnodes;
    return VGI_OK;
}

} // extern "C"
