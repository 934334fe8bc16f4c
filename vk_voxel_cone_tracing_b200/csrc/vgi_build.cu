// vgi_build.cu — clipmap build kernels for sm_100a: conservative voxelization, radiance injection,
// record finalisation, opacity/radiance down-sampling, atlas export.
//
// Compiled with -fmad=false: everything here that feeds a quantised result (occupancy bits, RGBA8
// texels) follows the IEEE binary32, no-FMA, left-to-right contract of DESIGN.md "numerics", so the
// results are bit-identical to the CPU oracle. Reference citations ("ref:") are relative to the
// reference checkout; this file restates behaviour, it does not share code with the GLSL.
#include "vgi_device.cuh"

// shadow taps of the injection: 1 = one lane per pair with aligned float4 row loads (lane_visibility), 0 = the quad-
// cooperative version of round 1 (warp_visibility). Measured on B200, 3.4 M pairs: see DESIGN.md section 4.
#ifndef VGI_INJECT_LANE_VIS
#define VGI_INJECT_LANE_VIS 1
#endif

// ---------------------------------------------------------------------------------------------------
// K1: voxelize — occupancy bits + (triangle, level, texel) pair list
// ---------------------------------------------------------------------------------------------------
#define SMALL_BOX_MAX 64

// The blend weight of a level's OWN texel inside its centre half (opacityDownSample.comp / radianceDownSample.comp: the
// texel is mix(down-sample of the finer level, own value, lerpFactor)); g = the voxel's offset from pm = min_corner of the
// finer level >> 1, every component < R/2. 0 everywhere but in the band of `band` texels along the centre region's rim.
DEVFN float mip_lerp_factor(const int* pm, const int* g, int half, int band)
{
    float dist[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int cur = pm[k] + g[k];
        const float center = (float)pm[k] + (float)((uint32_t)half >> 1);
        dist[k] = fabsf(((float)cur + 0.5f) - center) - 0.5f;
    }
    const uint32_t thrU = ((uint32_t)half >> 1) - (uint32_t)band;
    const float thr = (float)thrU;
    const float invBand = 1.0f / ((float)band + 1.0f);
    float lerpFactor = 0.0f;
    if (dist[0] >= thr || dist[1] >= thr || dist[2] >= thr) {
        lerpFactor = (f_max(dist[0], f_max(dist[1], dist[2])) - thr) + 1.0f;
        lerpFactor = lerpFactor * invBand;
    }
    return lerpFactor;
}

// Is the injected radiance of voxel (x, y, z) of `level` dead on arrival? Inside the centre half and off its blend band
// the record pass writes mix(down-sample, own, 0) = the down-sample, exactly, whatever the own value: such a (triangle,
// voxel) pair sets its occupancy bit (the raw opacity flag survives the mip) but is never shaded, so it is counted
// (Counters::pairs_unlisted) instead of listed. On the bench scene that is every pair of the levels whose finer level
// already spans the scene: 3.4 M canonical pairs, of which k_inject shades the rest.
DEVFN bool mip_interior(const BuildParams& bp, int level, int x, int y, int z)
{
    if (level == 0) return false;
    const int Rm = bp.R - 1, half = bp.R >> 1;
    const int pm[3] = { bp.lv[level - 1].min_corner[0] >> 1, bp.lv[level - 1].min_corner[1] >> 1, bp.lv[level - 1].min_corner[2] >> 1 };
    const int g[3] = { (x - pm[0]) & Rm, (y - pm[1]) & Rm, (z - pm[2]) & Rm };
    if (!(g[0] < half && g[1] < half && g[2] < half)) return false;
    return mip_lerp_factor(pm, g, half, bp.band) == 0.0f;
}

DEVFN void emit_pair(const BuildParams& bp, uint32_t* __restrict__ occ, vgi_pair_t* __restrict__ pairs,
                     Counters* __restrict__ cnt, uint32_t tri, int level, int vx, int vy, int vz)
{
    const int Rm = bp.R - 1;
    const uint32_t x = vx & Rm, y = vy & Rm, z = vz & Rm;
    if (!owns_plane(bp, (int)z)) return;   // slab-sharded build: the texel planes this GPU owns
    const size_t wordsPerLevel = ((size_t)bp.R * bp.R * bp.R) >> 5;
    const size_t w = (size_t)level * wordsPerLevel + (((((size_t)z << bp.logR) + y) << bp.logR) + x) / 32;
    const uint32_t bit = 1u << (x & 31u);
    if (!(occ[w] & bit)) atomicOr(&occ[w], bit);
    // warp-aggregated: one count of the unlisted pairs and one append of the listed ones per warp
    const bool dead = mip_interior(bp, level, (int)x, (int)y, (int)z);
    const unsigned all = __activemask();
    const unsigned md = __ballot_sync(all, dead);
    if (dead) {
        if ((int)lane_id() == __ffs(md) - 1) atomicAdd(&cnt->pairs_unlisted, (uint32_t)__popc(md));
        return;
    }
    const unsigned m = all & ~md;
    const int leader = __ffs(m) - 1;
    const unsigned rank = __popc(m & ((1u << lane_id()) - 1u));
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(&cnt->pairs, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t slot = base + rank;
    if (slot < bp.max_pairs)
        pairs[slot] = ((vgi_pair_t)tri << 32) | ((vgi_pair_t)level << 27) | ((vgi_pair_t)z << 18) | ((vgi_pair_t)y << 9) | x;
    else
        atomicOr(&cnt->overflow, 1u);
}

// TEX = the scene has textured materials (occlusion-texture alpha test per candidate pair); the untextured instantiation
// carries none of that code (measured: the shared kernel cost the factor-only scene 0.111 -> 0.138 ms)
template <bool TEX>
__global__ void __launch_bounds__(128) k_voxelize(BuildParams bp, const float4* __restrict__ tri_pos,
                                                   uint32_t* __restrict__ occ, vgi_pair_t* __restrict__ pairs,
                                                   uint2* __restrict__ large, Counters* __restrict__ cnt,
                                                   const vgi_material* __restrict__ materials, TexSet tex)
{
    // One thread per (level, triangle), level-major: neighbouring lanes are neighbouring triangles at the same
    // level. A thread first tests its (at most 64) candidate voxels into a 64-bit hit mask; then the warp
    // reserves the space for all its pairs with ONE atomic (warp prefix sum of the popcounts) and every lane
    // writes its pairs to consecutive slots. The single global pair counter is no longer hit once per voxel row.
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inRange = gid < bp.ntri * (uint32_t)bp.vox_nlev;
    const int lrel = inRange ? (int)(gid / bp.ntri) : 0;
    const int l = (int)((bp.vox_levels >> (4 * lrel)) & 15u);   // incremental build: only the levels whose occupancy or pairs are needed
    const uint32_t t = inRange ? gid - (uint32_t)lrel * bp.ntri : 0u;
    unsigned long long hits = 0ull, dead = 0ull;   // dead: occupancy only, never shaded (mip_interior)
    int lo0 = 0, lo1 = 0, lo2 = 0, nx = 1, ny = 1;
    if (inRange) {
        float p[9], N[3];
        int mat = 0;
        load_tri(tri_pos, t, p, &mat);
        const int axis = cross_and_axis(p, N);
        // alpha-tested materials (msaaVoxelizer.frag:64): the pair exists only where the occlusion texture passes
        const int occTex = TEX ? materials[mat].occlusion_texture : -1;
        TriSetup ts;
        tri_setup_level(ts, p, N, bp.lv[l], bp.R);
        if (ts.valid && ts.lo[0] <= ts.hi[0] && ts.lo[1] <= ts.hi[1] && ts.lo[2] <= ts.hi[2]) {
            nx = ts.hi[0] - ts.lo[0] + 1;
            ny = ts.hi[1] - ts.lo[1] + 1;
            const int nz = ts.hi[2] - ts.lo[2] + 1;
            const long long vol = (long long)nx * ny * nz;
            // slab-sharded build: this GPU owns the texel planes [z0, z1); planes of other GPUs are skipped before
            // any overlap test (and a big triangle is queued only where it can touch the slab)
            const int Rm = bp.R - 1;
            const bool slab = bp.z0 > 0 || bp.z1 < bp.R || bp.z_mask != 0;
            if (vol > SMALL_BOX_MAX) {
                bool mine = !slab;
                for (int z = ts.lo[2]; z <= ts.hi[2] && !mine; ++z) mine = owns_plane(bp, z & Rm);
                if (mine) {
                    const uint32_t slot = atomicAdd(&cnt->large, 1u);
                    if (slot < bp.max_large) large[slot] = make_uint2(t, (uint32_t)l);
                    else atomicOr(&cnt->overflow, 2u);
                }
            } else {
                lo0 = ts.lo[0]; lo1 = ts.lo[1]; lo2 = ts.lo[2];
                int i = 0;
                for (int z = ts.lo[2]; z <= ts.hi[2]; ++z) {
                    if (slab && !owns_plane(bp, z & Rm)) { i += nx * ny; continue; }
                    for (int y = ts.lo[1]; y <= ts.hi[1]; ++y)
                        for (int x = ts.lo[0]; x <= ts.hi[0]; ++x, ++i)
                            if (tri_overlaps_voxel(ts, x, y, z)) {
                                if (TEX && occTex > -1) {
                                    const float vs = bp.lv[l].voxel_size;
                                    float c[3] = { ((float)x + 0.5f) * vs, ((float)y + 0.5f) * vs, ((float)z + 0.5f) * vs };
                                    if (!alpha_test_pair(tex, occTex, t, axis, N, p, c)) continue;
                                }
                                hits |= 1ull << i;
                                if (mip_interior(bp, l, x, y, z)) dead |= 1ull << i;
                            }
                }
            }
        }
    }
    const int Rm = bp.R - 1;
    const uint32_t n = (uint32_t)__popcll(hits & ~dead);
    uint32_t incl = n, nDead = (uint32_t)__popcll(dead);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nDead += __shfl_xor_sync(0xffffffffu, nDead, o);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (!(total | nDead)) return;
    uint32_t base = 0u;
    if (lane_id() == 31u) {
        if (total) base = atomicAdd(&cnt->pairs, total);
        if (nDead) atomicAdd(&cnt->pairs_unlisted, nDead);
    }
    uint32_t slot = __shfl_sync(0xffffffffu, base, 31) + (incl - n);
    const size_t wordsPerLevel = ((size_t)bp.R * bp.R * bp.R) >> 5;
    for (unsigned long long h = hits; h; h &= h - 1) {
        const int i = __ffsll((long long)h) - 1;
        const int vx = lo0 + i % nx, vy = lo1 + (i / nx) % ny, vz = lo2 + i / (nx * ny);
        const uint32_t x = vx & Rm, y = vy & Rm, z = vz & Rm;
        const size_t w = (size_t)l * wordsPerLevel + (((((size_t)z << bp.logR) + y) << bp.logR) + x) / 32;
        const uint32_t bit = 1u << (x & 31u);
        if (!(occ[w] & bit)) atomicOr(&occ[w], bit);
        if ((dead >> i) & 1ull) continue;
        if (slot++ < bp.max_pairs)
            pairs[slot - 1u] = ((vgi_pair_t)t << 32) | ((vgi_pair_t)l << 27) | ((vgi_pair_t)z << 18) | ((vgi_pair_t)y << 9) | x;
        else
            atomicOr(&cnt->overflow, 1u);
    }
}

// one warp per big (triangle, level) item; lanes stride over the clipped bounding box
template <bool TEX>
__global__ void __launch_bounds__(256) k_voxelize_large(BuildParams bp, const float4* __restrict__ tri_pos,
                                                         uint32_t* __restrict__ occ, vgi_pair_t* __restrict__ pairs,
                                                         const uint2* __restrict__ large, Counters* __restrict__ cnt,
                                                         const vgi_material* __restrict__ materials, TexSet tex)
{
    const uint32_t nitems = min(cnt->large, bp.max_large);
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < nitems; item += warpsPerGrid) {
        const uint2 it = large[item];
        float p[9], N[3];
        int mat = 0;
        load_tri(tri_pos, it.x, p, &mat);
        const int axis = cross_and_axis(p, N);
        const int occTex = TEX ? materials[mat].occlusion_texture : -1;
        TriSetup ts;
        tri_setup_level(ts, p, N, bp.lv[it.y], bp.R);
        const int nx = ts.hi[0] - ts.lo[0] + 1, ny = ts.hi[1] - ts.lo[1] + 1, nz = ts.hi[2] - ts.lo[2] + 1;
        const long long vol = (long long)nx * ny * nz;
        for (long long base = 0; base < vol; base += 32) {
            const long long i = base + lane_id();
            if (i < vol) {
                const int x = ts.lo[0] + (int)(i % nx);
                const int y = ts.lo[1] + (int)((i / nx) % ny);
                const int z = ts.lo[2] + (int)(i / ((long long)nx * ny));
                bool hit = tri_overlaps_voxel(ts, x, y, z);
                if (TEX && hit && occTex > -1) {
                    const float vs = bp.lv[it.y].voxel_size;
                    float c[3] = { ((float)x + 0.5f) * vs, ((float)y + 0.5f) * vs, ((float)z + 0.5f) * vs };
                    hit = alpha_test_pair(tex, occTex, it.x, axis, N, p, c);
                }
                if (hit) emit_pair(bp, occ, pairs, cnt, it.x, (int)it.y, x, y, z);
            }
        }
    }
}

__global__ void k_zero_acc(uint32_t* __restrict__ acc, const Counters* __restrict__ cnt, uint32_t max_occ)
{
    const size_t n = (size_t)min(cnt->occ_total, max_occ) * 24;
    uint4* a4 = reinterpret_cast<uint4*>(acc);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x)
        a4[i] = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------------------
// K3: radiance injection — one lane per (triangle, voxel) pair for the shading, one quad per pair for
// the 16 shadow taps. ref: msaaInjectRadiance.frag:68-155 (shading), :176-183 (faces from -normal),
// shadow.glsl:8-36. Canonical accumulation (Q10): 16.16 fixed-point integer sums + count, order
// independent.
// ---------------------------------------------------------------------------------------------------
// (measured: 2, 3 or 4 resident blocks per SM give 429 / 421 / 420 us, and plain stores instead of the 64-bit
// reductions 340 us: the kernel is bound by the L2 -> L1 sector traffic of the shadow taps, 25 texels per pair
// with no overlap between neighbouring voxels, not by occupancy or by the atomics)
#ifndef VGI_INJECT_MINBLOCKS
#define VGI_INJECT_MINBLOCKS 4
#endif
template <bool TEX>
__global__ void __launch_bounds__(256, VGI_INJECT_MINBLOCKS) k_inject(BuildParams bp, LightParams lp, const float4* __restrict__ tri_pos,
                                                 const float4* __restrict__ tri_nrm, const vgi_material* __restrict__ materials,
                                                 const vgi_pair_t* __restrict__ pairs, const uint32_t* __restrict__ occ,
                                                 const uint32_t* __restrict__ occ_prefix, uint32_t* __restrict__ acc,
                                                 Counters* __restrict__ cnt, TexSet tex)
{
    const uint32_t npairs = min(cnt->pairs, bp.max_pairs);
    const uint32_t stride = gridDim.x * blockDim.x;
    const unsigned lane = lane_id();
    const bool compare = bp.shadow_compare != 0;
    // warp-uniform trip count: the quad phase needs every lane of the warp
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < npairs; base += stride) {
        const uint32_t i = base + lane;
        PairShade ps;
        ps.kind = 0;
        ps.px = ps.py = ps.cmpz = 0.0f;
        ps.NdotL = 0.0f;
        ps.n[0] = ps.n[1] = ps.n[2] = 0.0f;
        ps.mat = 0;
        ps.cell = 0;
        float uv[2] = { 0.0f, 0.0f };       // texture coordinate of the sample (textured materials only)
        if (i < npairs) {
            const vgi_pair_t pr = pairs[i];
            const uint32_t tri = (uint32_t)(pr >> 32);
            const int level = (int)((pr >> 27) & 7u);
            if ((bp.level_mask >> level) & 1u) {
                const int tx = (int)(pr & 511u), ty = (int)((pr >> 9) & 511u), tz = (int)((pr >> 18) & 511u);
                const LevelParams& lv = bp.lv[level];
                const int Rm = bp.R - 1;
                // unwrap the texel back to the voxel inside the region
                const int vx = lv.min_corner[0] + ((tx - lv.min_corner[0]) & Rm);
                const int vy = lv.min_corner[1] + ((ty - lv.min_corner[1]) & Rm);
                const int vz = lv.min_corner[2] + ((tz - lv.min_corner[2]) & Rm);
                float p[9], n9[9], N[3];
                load_tri(tri_pos, tri, p, &ps.mat);
                {
                    const float4 a = __ldg(tri_nrm + 3 * (size_t)tri), b = __ldg(tri_nrm + 3 * (size_t)tri + 1), c = __ldg(tri_nrm + 3 * (size_t)tri + 2);
                    n9[0] = a.x; n9[1] = a.y; n9[2] = a.z; n9[3] = b.x; n9[4] = b.y; n9[5] = b.z; n9[6] = c.x; n9[7] = c.y; n9[8] = c.z;
                }
                const int axis = cross_and_axis(p, N);
                float c[3] = { ((float)vx + 0.5f) * lv.voxel_size, ((float)vy + 0.5f) * lv.voxel_size, ((float)vz + 0.5f) * lv.voxel_size };
                float pos[3], nrm[3], bary[3];
                bool sampled = inject_sample_at(axis, N, p, n9, c, pos, nrm, bary);
                if (TEX && sampled) {
                    const vgi_material* mt = materials + ps.mat;
                    if (mt->base_color_texture > -1 || mt->emissive_texture > -1 || mt->occlusion_texture > -1) {
                        tri_uv_at(tex, tri, bary, uv);
                        if (mt->occlusion_texture > -1) {       // ref: msaaInjectRadiance.frag:73
                            float t[4];
                            tex_fetch(tex, mt->occlusion_texture, uv[0], uv[1], t);
                            if (t[0] < 0.1f) sampled = false;
                        }
                    }
                }
                if (sampled) {
                    const vgi_material* m = materials + ps.mat;
                    const size_t wordsPerLevel = ((size_t)bp.R * bp.R * bp.R) >> 5;
                    const size_t w = (size_t)level * wordsPerLevel + (((((size_t)tz << bp.logR) + ty) << bp.logR) + tx) / 32;
                    ps.cell = occ_prefix[w] + __popc(occ[w] & ((1u << (tx & 31)) - 1u));
                    if (m->emissive_factor[0] > 0.0f || m->emissive_factor[1] > 0.0f || m->emissive_factor[2] > 0.0f) {
                        ps.kind = 1;
                    } else {
                        const float len2 = dot3(nrm, nrm);
                        if (len2 > 0.0f) {
                            const float len = sqrtf(len2);
                            ps.n[0] = nrm[0] / len; ps.n[1] = nrm[1] / len; ps.n[2] = nrm[2] / len;
                            ps.NdotL = f_clamp(dot3(ps.n, lp.dir_to_light), 0.001f, 1.0f);
                            // ref: shadow.glsl:28-36 — light-space xy of the sample
                            float l[3];
                            xform_point(lp.view, pos, l);
                            const float* P = lp.proj;
                            float qx = ((P[0] * l[0] + P[4] * l[1]) + P[8] * 0.0f) + P[12];
                            float qy = ((P[1] * l[0] + P[5] * l[1]) + P[9] * 0.0f) + P[13];
                            ps.px = qx * 0.5f + 0.5f;
                            ps.py = qy * 0.5f + 0.5f;
                            if (compare) ps.cmpz = (P[10] * l[2] + P[14]) - 0.002f;
                            ps.kind = 2;
                        }
                    }
                    if (ps.kind && ps.cell >= bp.max_occ) { atomicOr(&cnt->overflow, 4u); ps.kind = 0; }
                }
            }
        }
        // quad phase: round r serves the pairs of lanes 8r..8r+7, quad q serves lane 8r+q
        const unsigned lit = __ballot_sync(0xffffffffu, ps.kind == 2);
#if VGI_INJECT_LANE_VIS
        (void)lit;
        const float vis = ps.kind == 2 ? lane_visibility(lp, ps.px, ps.py, ps.cmpz, compare) : 0.0f;
#else
        const float vis = warp_visibility(lp, ps.px, ps.py, ps.cmpz, compare, lit);
#endif
        // contribution of this lane: up to 6 faces x (r, g, b) in 16.16 fixed point
        const vgi_material* m = materials + ps.mat;
        uint32_t q[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
        uint32_t fcode = 0u;            // lit: sign bits of the three faces; emissive: 8
        bool contrib = false;
        if (ps.kind == 1) {
            float em[3] = { m->emissive_factor[0], m->emissive_factor[1], m->emissive_factor[2] };
            if (TEX && m->emissive_texture > -1) {    // ref: msaaInjectRadiance.frag:79-82
                float t[4];
                tex_fetch(tex, m->emissive_texture, uv[0], uv[1], t);
                em[0] = em[0] + t[0]; em[1] = em[1] + t[1]; em[2] = em[2] + t[2];
            }
            q[0][0] = (uint32_t)(f_clamp(em[0], 0.0f, 1.0f) * 65536.0f + 0.5f);
            q[0][1] = (uint32_t)(f_clamp(em[1], 0.0f, 1.0f) * 65536.0f + 0.5f);
            q[0][2] = (uint32_t)(f_clamp(em[2], 0.0f, 1.0f) * 65536.0f + 0.5f);
            fcode = 8u;
            contrib = true;
        } else if (ps.kind == 2) {
            float lc[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) lc[k] = ((ps.NdotL * vis) * lp.color[k]) * lp.intensity;
            if (!(lc[0] == 0.0f && lc[1] == 0.0f && lc[2] == 0.0f)) {
                float col[4] = { m->base_color_factor[0], m->base_color_factor[1], m->base_color_factor[2], m->base_color_factor[3] };
                if (TEX && m->base_color_texture > -1) {  // ref: msaaInjectRadiance.frag:131-136
                    float t[4];
                    tex_fetch(tex, m->base_color_texture, uv[0], uv[1], t);
#pragma unroll
                    for (int k = 0; k < 4; ++k) col[k] = col[k] * t[k];
                }
                float rad[3];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    rad[k] = f_clamp((lc[k] * col[k]) * col[3], 0.0f, 1.0f);
                // faces selected by -normal (ref: msaaInjectRadiance.frag:152-153): 0/1 = +X/-X, ...
                fcode = ((-ps.n[0] > 0.0f) ? 0u : 1u) | ((-ps.n[1] > 0.0f) ? 0u : 2u) | ((-ps.n[2] > 0.0f) ? 0u : 4u);
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    const float w = fabsf(ps.n[f]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) q[f][k] = (uint32_t)((rad[k] * w) * 65536.0f + 0.5f);
                }
                contrib = true;
            }
        }
        if (!contrib) continue;
        // two 32-bit sums per 64-bit reduction: {r, g} and {b, count} of a face are adjacent words, and a sum
        // cannot carry into its neighbour before it would overflow its own 32 bits
        unsigned long long* a = reinterpret_cast<unsigned long long*>(acc + (size_t)ps.cell * 24);
        if (fcode == 8u) {
            const unsigned long long rg = ((unsigned long long)q[0][1] << 32) | q[0][0];
            const unsigned long long bn = (1ull << 32) | q[0][2];
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                atomicAdd(a + f * 2 + 0, rg);
                atomicAdd(a + f * 2 + 1, bn);
            }
        } else {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                unsigned long long* af = a + (2 * f + ((fcode >> f) & 1u)) * 2;
                atomicAdd(af + 0, ((unsigned long long)q[f][1] << 32) | q[f][0]);
                atomicAdd(af + 1, (1ull << 32) | q[f][2]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K4: k_level_masks + k_level_records — per clip level, write every record of the level that is, or was
// last frame, non-zero; all other records are already zero and stay untouched. k_level_masks derives the
// masks below and a compact work list of the visited voxels; k_level_records (one thread per list entry, so
// every lane works) fuses per visited voxel
//   the reference's clears (A4 VoxelizationPass.cpp:104-126, A7 clipmapCleaning.comp),
//   the raw opacity store (msaaVoxelizer.frag:69-73),
//   the radiance average (msaaInjectRadiance.frag:185-201, canonical mean) and copy-alpha (A9),
//   the opacity and radiance down-sample of level-1 into the centre half
//     (opacityDownSample.comp:28-138, radianceDownSample.comp:28-139 with Q6 repaired),
// and writes the 32-byte record exactly once. The visit set comes from bit masks:
//   occ[l]  raw occupancy of this frame (k_voxelize)
//   nz[l]   superset of the non-zero records = occ[l] | centre-half(OR of the 2x2x2 children in nz[l-1])
//   nzPrev  nz of the previous frame (records that may need clearing)
// k_level_masks: lane = mask word (32 voxels along x); the levels run fine to coarse (nz[l] needs nz[l-1]).
// k_level_records: levels run fine to coarse as well (the children of level l are the records of l-1).
// ---------------------------------------------------------------------------------------------------
// Child pair pr (0..3) of face f in the shader's OFFSETS order (child index = dx + 2 dy + 4 dz):
// face 2a composites the low-side child over the high-side child along axis a, face 2a+1 the reverse
// (ref: opacityDownSample.comp:97-138). Pure functions of compile-time loop indices, so the child
// arrays stay in registers.
DEVFN constexpr int ds_pair_base(int f, int pr)
{
    return (f >> 1) == 0 ? (pr << 1) : ((f >> 1) == 1 ? ((pr & 1) | ((pr & 2) << 1)) : pr);
}
DEVFN constexpr int ds_pair_first(int f, int pr) { return ds_pair_base(f, pr) | ((f & 1) ? (1 << (f >> 1)) : 0); }
DEVFN constexpr int ds_pair_second(int f, int pr) { return ds_pair_base(f, pr) | ((f & 1) ? 0 : (1 << (f >> 1))); }

DEVFN uint32_t low_bits(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)); }

// bits b of a 32-texel word starting at texel x0 with ((x0 + b - start) mod R) < half
DEVFN uint32_t centre_mask(int x0, int start, int R, int half)
{
    const int d0 = (x0 - start) & (R - 1);
    const int e = R - d0;                       // bits until the wrap
    const uint32_t first = low_bits(min(e, half - d0));
    const uint32_t second = low_bits(e + half) & ~low_bits(e);
    return first | second;
}

// keep the even bits of a 64-bit value: bit b of the result = bit 2b of v
DEVFN uint32_t even_bits(unsigned long long v)
{
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
    v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
    v = (v | (v >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)v;
}

struct Rec { uint4 lo, hi; };

DEVFN uint32_t rec_radiance(const Rec& r, int f)
{
    return f == 0 ? r.lo.x : f == 1 ? r.lo.y : f == 2 ? r.lo.z : f == 3 ? r.lo.w : f == 4 ? r.hi.x : r.hi.y;
}

__global__ void __launch_bounds__(256) k_level_masks(BuildParams bp, int level, const uint32_t* __restrict__ occ,
                                                      const uint32_t* __restrict__ nzPrev, uint32_t* __restrict__ nzCur,
                                                      uint32_t* __restrict__ visitList, uint32_t listCap, Counters* __restrict__ cnt)
{
    const int R = bp.R, Rm = R - 1, half = R >> 1, logR = bp.logR;
    const int wpr = R >> 5;                     // mask words per x row
    const size_t nvox = (size_t)R * R * R;
    const uint32_t wordsPerLevel = (uint32_t)(nvox >> 5);
    const uint32_t* occL = occ + (size_t)level * wordsPerLevel;
    const uint32_t* prevL = nzPrev + (size_t)level * wordsPerLevel;
    uint32_t* curL = nzCur + (size_t)level * wordsPerLevel;
    const uint32_t* fineL = nzCur + (size_t)(level - 1) * wordsPerLevel;   // valid for level > 0
    const bool inject = (bp.level_mask >> level) & 1u;
    const bool mip = level > 0;
    int pm[3] = { 0, 0, 0 };
    if (mip) { pm[0] = bp.lv[level - 1].min_corner[0] >> 1; pm[1] = bp.lv[level - 1].min_corner[1] >> 1; pm[2] = bp.lv[level - 1].min_corner[2] >> 1; }
    const unsigned lane = lane_id();
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nchunks = wordsPerLevel >> 5;
    uint32_t* list = visitList + (size_t)level * listCap;

    for (uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warpsPerGrid) {
        // ---- phase 1: lane = mask word
        const uint32_t wi = chunk * 32u + lane;
        const int xw = (int)(wi % (uint32_t)wpr);
        const uint32_t row = wi / (uint32_t)wpr;
        const int y = (int)(row & (uint32_t)Rm), z = (int)(row >> logR);
        const uint32_t occw = __ldg(occL + wi);
        const uint32_t prevw = __ldg(prevL + wi);
        uint32_t nzw = occw;
        if (mip) {
            const bool in_yz = (((y - pm[1]) & Rm) < half) && (((z - pm[2]) & Rm) < half);
            if (in_yz) {
                const int fy = (2 * y) & Rm, fz = (2 * z) & Rm;
                const int fxw = ((2 * xw * 32) & Rm) >> 5;
                unsigned long long o = 0ull;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const size_t frow = ((size_t)(fz + (d >> 1)) << logR) + (size_t)(fy + (d & 1));
                    const uint32_t a = fineL[frow * wpr + fxw];
                    const uint32_t b = wpr > 1 ? fineL[frow * wpr + fxw + 1] : a;
                    o |= ((unsigned long long)b << 32) | a;
                }
                nzw |= even_bits(o | (o >> 1)) & centre_mask(xw * 32, pm[0] & Rm, R, half);
            }
        }
        if (!inject) nzw |= prevw;              // off-cadence: last frame's radiance stays in the records
        curL[wi] = nzw;
        const uint32_t visit = nzw | prevw;

        // ---- append the visited voxels (texel index inside the level) to this level's work list
        const uint32_t n = (uint32_t)__popc(visit);
        uint32_t incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (!total) continue;
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(&cnt->visit[level], total);
        base = __shfl_sync(0xffffffffu, base, 31) + (incl - n);
        for (uint32_t v = visit; v; v &= v - 1) {
            if (base < listCap) list[base] = wi * 32u + (uint32_t)(__ffs(v) - 1);
            else atomicOr(&cnt->overflow, 4u);
            ++base;
        }
    }
}

// MODE 0: fused finalize + mip (single GPU). Slab-sharded build: MODE 1 finalizes the
// records of this GPU's z slab only and appends them to the exchange buffer; MODE 2 runs the mips on the
// gathered store (own record read back instead of recomputed).
struct SlabPack {
    uint32_t* ids;      // level << 27 | texel index
    uint4*    recs;     // 2 x uint4 per record
    uint32_t* count;
    uint32_t  cap;
    // MODE 3 (peer build): the voxel stores of every GPU, this one included, mapped into this address space
    VoxelRecord* peer_store[VGI_MAX_PEERS];
    int       npeers;
};

template <int MODE>
__global__ void __launch_bounds__(128) k_level_records(BuildParams bp, int level_arg, const uint32_t* __restrict__ occ,
                                                        const uint32_t* __restrict__ occ_prefix,
                                                        const uint32_t* __restrict__ acc, const uint32_t* __restrict__ nzCur,
                                                        const uint32_t* __restrict__ visitList, uint32_t listCap,
                                                        Counters* __restrict__ cnt, VoxelRecord* __restrict__ store, SlabPack pack)
{
    __shared__ float s_unorm[256];              // (float)c / 255.0f, exact
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = (float)i / 255.0f;
    __syncthreads();

    // peer build: the levels are independent before the mip pass, one launch covers them all (blockIdx.y = level)
    const int level = MODE == 3 ? (int)blockIdx.y : level_arg;
    const int R = bp.R, Rm = R - 1, half = R >> 1, logR = bp.logR;
    const int wpr = R >> 5;
    const size_t nvox = (size_t)R * R * R;
    const uint32_t wordsPerLevel = (uint32_t)(nvox >> 5);
    const uint32_t* occL = occ + (size_t)level * wordsPerLevel;
    const uint32_t* prefL = occ_prefix + (size_t)level * wordsPerLevel;
    const uint32_t* fineL = nzCur + (size_t)(level - 1) * wordsPerLevel;   // valid for level > 0
    const bool inject = (bp.level_mask >> level) & 1u;
    const bool mip = level > 0;
    int pm[3] = { 0, 0, 0 };
    if (mip) { pm[0] = bp.lv[level - 1].min_corner[0] >> 1; pm[1] = bp.lv[level - 1].min_corner[1] >> 1; pm[2] = bp.lv[level - 1].min_corner[2] >> 1; }
    VoxelRecord* storeL = store + (size_t)level * nvox;
    const VoxelRecord* storeF = store + (size_t)(level - 1) * nvox;
    const uint32_t* list = visitList + (size_t)level * listCap;
    const uint32_t n = min(cnt->visit[level], listCap);

    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t vid = list[i];
        const int vx = (int)(vid & (uint32_t)Rm), vy = (int)((vid >> logR) & (uint32_t)Rm), vz = (int)(vid >> (2 * logR));
        const uint32_t wis = vid >> 5;
        const uint32_t lane = vid & 31u;
        if ((MODE == 1 || MODE == 3) && !owns_plane(bp, vz)) continue;
        const uint32_t oword = __ldg(occL + wis);
        VoxelRecord* dstp = storeL + ((((size_t)vz << logR) + vy) << logR) + vx;
        uint4* dst = reinterpret_cast<uint4*>(dstp);
        const bool raw = (oword >> lane) & 1u;
        const int g0 = (vx - pm[0]) & Rm, g1 = (vy - pm[1]) & Rm, g2 = (vz - pm[2]) & Rm;
        const bool centre = mip && g0 < half && g1 < half && g2 < half;
        if (MODE == 2 && !centre) continue;

        // -- finalize (k_finalize semantics)
        Rec own;
        if (MODE == 2) {
            own.lo = dst[0];
            own.hi = dst[1];
        } else if (inject) {
            uint32_t rad[6] = { 0, 0, 0, 0, 0, 0 };
            if (raw) {
                const uint32_t idx = __ldg(prefL + wis) + __popc(oword & ((1u << lane) - 1u));
                if (idx < bp.max_occ) {
                    const uint4* a = reinterpret_cast<const uint4*>(acc + (size_t)idx * 24);
#pragma unroll
                    for (int f = 0; f < 6; ++f) {
                        const uint4 s = a[f];
                        uint32_t r = 0, g = 0, b = 0;
                        if (s.w) {
                            r = (uint32_t)(((unsigned long long)s.x * 255ull) >> 16) / s.w;
                            g = (uint32_t)(((unsigned long long)s.y * 255ull) >> 16) / s.w;
                            b = (uint32_t)(((unsigned long long)s.z * 255ull) >> 16) / s.w;
                            r = r > 255u ? 255u : r; g = g > 255u ? 255u : g; b = b > 255u ? 255u : b;
                        }
                        rad[f] = r | (g << 8) | (b << 16) | 0xff000000u; // copy-alpha: opacity.a of an occupied voxel = 1
                    }
                } else {
#pragma unroll
                    for (int f = 0; f < 6; ++f) rad[f] = 0xff000000u;
                }
            }
            own.lo = make_uint4(rad[0], rad[1], rad[2], rad[3]);
            own.hi.x = rad[4];
            own.hi.y = rad[5];
        } else {
            // off-cadence level: the radiance texels (including their alpha) keep last frame's values
            own.lo = dst[0];
            const uint4 old = dst[1];
            own.hi.x = old.x;
            own.hi.y = old.y;
        }
        if (MODE != 2) {
            own.hi.z = raw ? 0xffffffffu : 0u;          // opacity alpha faces 0..3
            own.hi.w = raw ? 0x0001ffffu : 0u;          // opacity alpha faces 4,5 ; raw flag ; pad
        }
        if (MODE == 3) {
            // the record goes straight into every GPU's store over NVLink (two 16-byte stores per peer)
            const size_t off = (size_t)level * nvox + ((((size_t)vz << logR) + vy) << logR) + vx;
            for (int r = 0; r < pack.npeers; ++r) {
                uint4* pd = reinterpret_cast<uint4*>(pack.peer_store[r] + off);
                pd[0] = own.lo;
                pd[1] = own.hi;
            }
            continue;
        }
        if (MODE == 1) {
            dst[0] = own.lo;
            dst[1] = own.hi;
            const uint32_t slot = atomicAdd(pack.count, 1u);
            if (slot < pack.cap) {
                pack.ids[slot] = ((uint32_t)level << 27) | vid;
                pack.recs[2 * (size_t)slot] = own.lo;
                pack.recs[2 * (size_t)slot + 1] = own.hi;
            } else {
                atomicOr(&cnt->overflow, 4u);
            }
            continue;
        }

        // -- down-sample of level-1 into the centre half (k_downsample semantics)
        if (centre) {
            const int g[3] = { g0, g1, g2 };
            int pstart[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) pstart[k] = ((pm[k] + g[k]) << 1) & Rm;
            const float lerpFactor = mip_lerp_factor(pm, g, half, bp.band);   // 0: the own radiance is not read (mip_interior)
            // children: index i = dx + 2*dy + 4*dz; records whose nz bit is clear are zero
            Rec ch[8];
            uint32_t anyChild = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cx = pstart[0] + (i & 1), cy = pstart[1] + ((i >> 1) & 1), cz = pstart[2] + (i >> 2);
                const size_t crow = ((size_t)cz << logR) + (size_t)cy;
                const uint32_t fw = fineL[crow * wpr + (cx >> 5)];
                ch[i].lo = make_uint4(0, 0, 0, 0);
                ch[i].hi = make_uint4(0, 0, 0, 0);
                if ((fw >> (cx & 31)) & 1u) {
                    const uint4* srcp = reinterpret_cast<const uint4*>(storeF + (crow << logR) + cx);
                    ch[i].lo = srcp[0];
                    ch[i].hi = srcp[1];
                    anyChild = 1u;
                }
            }
            const float ownRaw = raw ? 1.0f : 0.0f; // opacity.r of this level
            uint32_t newOp[6], newRad[6];
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                float s = 0.0f;
                if (anyChild) {
#pragma unroll
                    for (int pr = 0; pr < 4; ++pr) {
                        const int i0 = ds_pair_first(f, pr), i1 = ds_pair_second(f, pr);
                        const uint32_t w0 = f < 4 ? ch[i0].hi.z : ch[i0].hi.w;
                        const uint32_t w1 = f < 4 ? ch[i1].hi.z : ch[i1].hi.w;
                        const float a0 = s_unorm[(w0 >> (8 * (f & 3))) & 0xffu];
                        const float a1 = s_unorm[(w1 >> (8 * (f & 3))) & 0xffu];
                        s = s + a0;
                        s = s + (1.0f - a0) * a1;
                    }
                }
                const float dsOp = s * 0.25f;
                const uint32_t aOp = f_to_unorm8(f_mix(dsOp, ownRaw, lerpFactor));
                newOp[f] = aOp;
                if (inject) {
                    const uint32_t ow = rec_radiance(own, f);
                    // own texel: rgb as injected, alpha = this level's final opacity alpha (copy-alpha ran before the radiance mips)
                    const uint32_t ownTexel = (ow & 0x00ffffffu) | (aOp << 24);
                    float sc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                    if (anyChild) {
#pragma unroll
                        for (int pr = 0; pr < 4; ++pr) {
                            const uint32_t w0 = rec_radiance(ch[ds_pair_first(f, pr)], f);
                            const uint32_t w1 = rec_radiance(ch[ds_pair_second(f, pr)], f);
                            if ((w0 | w1) == 0u) continue;      // adds exact zeros
                            const float k1 = 1.0f - s_unorm[w0 >> 24];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                sc[c] = sc[c] + s_unorm[(w0 >> (8 * c)) & 0xffu];
                                sc[c] = sc[c] + k1 * s_unorm[(w1 >> (8 * c)) & 0xffu];
                            }
                        }
                    }
                    uint32_t out = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float ds = sc[c] * 0.25f;
                        out |= f_to_unorm8(f_mix(ds, s_unorm[(ownTexel >> (8 * c)) & 0xffu], lerpFactor)) << (8 * c);
                    }
                    newRad[f] = out;
                }
            }
            if (inject) {
                own.lo = make_uint4(newRad[0], newRad[1], newRad[2], newRad[3]);
                own.hi.x = newRad[4];
                own.hi.y = newRad[5];
            }
            own.hi.z = newOp[0] | (newOp[1] << 8) | (newOp[2] << 16) | (newOp[3] << 24);
            own.hi.w = (own.hi.w & 0xffff0000u) | newOp[4] | (newOp[5] << 8);
        }
        dst[0] = own.lo;
        dst[1] = own.hi;
    }
}

// ---------------------------------------------------------------------------------------------------
// K5: empty-space masks for the cone tracer, derived from the nz bits of this frame.
//  brick     one bit per 4^3 brick, set when any record in voxels [4b, 4b+4] (per axis, toroidal; the +1
//            covers the high corner of a tri-linear footprint whose low corner lies in the brick) may be
//            non-zero. One byte per (brick row, nz word): bit k of byte [(bz*(R/4) + by)*wpr + xw] is
//            brick bx = 8*xw + k.
//  footprint one byte per voxel, written for the voxels of non-empty bricks and zeroed for those of bricks that were
//            non-empty in the last build (zero-initialised: valid everywhere): bit c = dx + 2 dy + 4 dz is
//            the nz bit of record (x+dx, y+dy, z+dz), i.e. which of the 8 records of the footprint whose
//            low corner is this voxel have to be fetched.
// One thread per brick byte: 25 rows of 33 nz bits stay in registers.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_brick_mask(int R, int L, int logR, const uint32_t* __restrict__ nz,
                                                     uint8_t* __restrict__ brick, uint8_t* __restrict__ footprint)
{
    const int wpr = R >> 5, nb = R >> 2, Rm = R - 1;
    const size_t perLevel = (size_t)nb * nb * wpr;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= perLevel * L) return;
    const int level = (int)(i / perLevel);
    const size_t j = i % perLevel;
    const int xw = (int)(j % wpr), by = (int)((j / wpr) % nb), bz = (int)(j / ((size_t)wpr * nb));
    const size_t nvox = (size_t)R * R * R;
    const uint32_t* nzL = nz + (size_t)level * (nvox >> 5);
    unsigned long long rows[5][5];
    unsigned long long any = 0ull;
#pragma unroll
    for (int dz = 0; dz < 5; ++dz)
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const size_t row = ((size_t)((bz * 4 + dz) & Rm) << logR) + (size_t)((by * 4 + dy) & Rm);
            const uint32_t lo = __ldg(nzL + row * wpr + xw);
            const uint32_t hi = __ldg(nzL + row * wpr + ((xw + 1) % wpr)) & 1u;
            rows[dz][dy] = ((unsigned long long)hi << 32) | lo;
            any |= rows[dz][dy];
        }
    uint32_t out = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if ((any >> (4 * k)) & 0x1full) out |= 1u << k;
    // bricks that emptied since the last build get their footprint bytes zeroed (their nz rows are all zero, so the
    // words below come out 0): the footprint byte is valid for EVERY voxel and the tracer probes it alone
    const uint32_t touch = out | brick[i];
    brick[i] = (uint8_t)out;
    if (!touch) return;
    uint8_t* fpL = footprint + (size_t)level * nvox;
#pragma unroll
    for (int z = 0; z < 4; ++z)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const unsigned long long r00 = rows[z][y], r10 = rows[z][y + 1], r01 = rows[z + 1][y], r11 = rows[z + 1][y + 1];
            uint32_t* dst = reinterpret_cast<uint32_t*>(fpL + ((((size_t)(bz * 4 + z) << logR) + (size_t)(by * 4 + y)) << logR) + (size_t)xw * 32);
            for (int k = 0; k < 8; ++k) {
                if (!((touch >> k) & 1u)) continue;
                const uint32_t a = (uint32_t)(r00 >> (4 * k)) & 0x1fu, b = (uint32_t)(r10 >> (4 * k)) & 0x1fu;
                const uint32_t c = (uint32_t)(r01 >> (4 * k)) & 0x1fu, d = (uint32_t)(r11 >> (4 * k)) & 0x1fu;
                uint32_t word = 0u;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const uint32_t byte = ((a >> x) & 3u) | (((b >> x) & 3u) << 2) | (((c >> x) & 3u) << 4) | (((d >> x) & 3u) << 6);
                    word |= byte << (8 * x);
                }
                dst[k] = word;
            }
        }
}

// ---------------------------------------------------------------------------------------------------
// export to the reference atlas layout (ref: Voxelizer.h:40-52; texel addressing msaaVoxelizer.frag:43-55;
// border: borderWrapping.comp:14-37, canonical = both sides of both atlases, literal = Q4/Q5)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_export(int R, int L, int logR, int which, int literal,
                                                 const VoxelRecord* __restrict__ store, uint32_t* __restrict__ dst)
{
    const int rb = R + 2;
    const size_t W = (size_t)rb * 6, H = (size_t)rb * L, D = rb;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H * D) return;
    const int ax = (int)(i % W), ay = (int)((i / W) % H), az = (int)(i / (W * H));
    const int f = ax / rb, bx = ax % rb;
    const int l = ay / rb, by = ay % rb;
    const int bz = az;
    const bool border = bx == 0 || by == 0 || bz == 0 || bx == R + 1 || by == R + 1 || bz == R + 1;
    if (border && literal) {
        const int lim = ((R + 2) >> 3) << 3; // BorderWrapper.cpp:139 dispatches (R+2)>>3 groups of 8
        const bool written = which == 0 && bx < lim && by < lim && bz < lim && (bx == 0 || by == 0 || bz == 0);
        if (!written) { dst[i] = 0u; return; }
    }
    const int x = (bx + R - 1) & (R - 1), y = (by + R - 1) & (R - 1), z = (bz + R - 1) & (R - 1);
    const VoxelRecord* rec = store + (size_t)l * R * R * R + ((((size_t)z << logR) + y) << logR) + x;
    uint32_t out;
    if (which == 1) out = rec->radiance[f];
    else {
        const uint32_t raw = rec->raw ? 255u : 0u, a = rec->opacity[f];
        out = raw | (a << 8) | (raw << 16) | (a << 24);
    }
    dst[i] = out;
}

// ---------------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------------
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
// launch + optional per-kernel event timing (vgi_set_timing)
#define LAUNCH(name, ...) do { c->timer.begin(name, s); __VA_ARGS__; c->timer.end(s); ++n; } while (0)

// world transform of the scene on the device (vgi_update_nodes). ref: msaaVoxelizer.vert:31-36 — position = model * p,
// normal = itModel * n; same binary32 expression order as vgi_set_scene's host loop, so the triangle soup is identical
__global__ void __launch_bounds__(256) k_transform_scene(uint32_t nvert, const float4* __restrict__ obj_pos, const float4* __restrict__ obj_nrm,
                                                          const vgi_node_matrix* __restrict__ nodes, float4* __restrict__ tri_pos,
                                                          float4* __restrict__ tri_nrm, const float4* __restrict__ obj_tan,
                                                          float4* __restrict__ tri_tan)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvert) return;
    const float4 p = obj_pos[i], n = obj_nrm[i];
    const vgi_node_matrix& nm = nodes[__float_as_uint(p.w)];
    const float* m = nm.model;
    const float* it = nm.it_model;
    float w[3], d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        w[r] = ((m[r] * p.x + m[4 + r] * p.y) + m[8 + r] * p.z) + m[12 + r];
        d[r] = (it[r] * n.x + it[4 + r] * n.y) + it[8 + r] * n.z;
    }
    tri_pos[i] = make_float4(w[0], w[1], w[2], n.w);     // w keeps the material index
    tri_nrm[i] = make_float4(d[0], d[1], d[2], 0.0f);
    if (obj_tan) {      // ref: gBufferPass.vert:40
        const float4 t = obj_tan[i];
        float g[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) g[r] = (it[r] * t.x + it[4 + r] * t.y) + it[8 + r] * t.z;
        tri_tan[i] = make_float4(g[0], g[1], g[2], t.w);
    }
}

int vgi_launch_transform_scene(vgi_ctx* c, cudaStream_t s)
{
    int n = 0;
    const uint32_t nvert = c->ntri * 3u;
    LAUNCH("k_transform_scene", k_transform_scene<<<cdiv(nvert, 256), 256, 0, s>>>(nvert, c->obj_pos, c->obj_nrm, c->d_nodes, c->tri_pos, c->tri_nrm, c->obj_tan, c->tri_tan));
    return n;
}

int vgi_launch_voxelize(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    int n = 0;
    const size_t wordsPerLevel = (size_t)bp.R * bp.R * bp.R >> 5;
    const size_t nwords = wordsPerLevel * bp.L;
    const int nlev = bp.vox_nlev;
    cudaMemsetAsync(c->counters, 0, sizeof(Counters), s);
    if (nlev == bp.L) cudaMemsetAsync(c->occ, 0, nwords * sizeof(uint32_t), s);
    else
        for (int i = 0; i < nlev; ++i)
            cudaMemsetAsync(c->occ + wordsPerLevel * ((bp.vox_levels >> (4 * i)) & 15u), 0, wordsPerLevel * sizeof(uint32_t), s);
    if (bp.ntri && nlev > 0) {
        if (c->scene_max_texture > -1) {
            LAUNCH("k_voxelize", k_voxelize<true><<<cdiv((size_t)bp.ntri * nlev, 128), 128, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
            LAUNCH("k_voxelize_large", k_voxelize_large<true><<<148 * 4, 256, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
        } else {
            LAUNCH("k_voxelize", k_voxelize<false><<<cdiv((size_t)bp.ntri * nlev, 128), 128, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
            LAUNCH("k_voxelize_large", k_voxelize_large<false><<<148 * 4, 256, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
        }
    }
    const unsigned nblk = cdiv(nwords, SCAN_BLOCK * SCAN_ITEMS);
    LAUNCH("k_scan_block_sums", k_scan_block_sums<<<nblk, SCAN_BLOCK, 0, s>>>(c->occ, nwords, c->block_sums));
    LAUNCH("k_scan_sums", k_scan_sums<<<1, 1024, 0, s>>>(c->block_sums, nblk, &c->counters->occ_total));
    LAUNCH("k_scan_final", k_scan_final<<<nblk, SCAN_BLOCK, 0, s>>>(c->occ, nwords, c->block_sums, c->occ_prefix));
    return n;
}

static void launch_masks(vgi_ctx* c, const BuildParams& bp, int cur, cudaStream_t s, int& n)
{
    const uint32_t chunks = (uint32_t)((((size_t)bp.R * bp.R * bp.R) >> 5) >> 5);
    const unsigned grid = min(cdiv(chunks, 8), 148u * 16u);
    if (bp.level_first > 0) {   // untouched levels: their masks are last build's
        const size_t bytes = (((size_t)bp.R * bp.R * bp.R) >> 5) * bp.level_first * sizeof(uint32_t);
        cudaMemcpyAsync(c->nz[cur], c->nz[cur ^ 1], bytes, cudaMemcpyDeviceToDevice, s);
    }
    for (int l = bp.level_first; l < bp.L; ++l)
        LAUNCH("k_level_masks", k_level_masks<<<grid, 256, 0, s>>>(bp, l, c->occ, c->nz[cur ^ 1], c->nz[cur], c->visit_list, c->visit_cap, c->counters));
}

static void launch_brick(vgi_ctx* c, const BuildParams& bp, int cur, cudaStream_t s, int& n)
{
    const size_t nbytes = (size_t)(bp.R >> 2) * (bp.R >> 2) * (bp.R >> 5) * bp.L;
    LAUNCH("k_brick_mask", k_brick_mask<<<cdiv(nbytes, 128), 128, 0, s>>>(bp.R, bp.L, bp.logR, c->nz[cur], c->brick_mask, c->footprint));
}

#ifdef VGI_INJECT_SORT_EXPERIMENT
#include <cub/device/device_radix_sort.cuh>
// development experiment (tools/build_variant.py -DVGI_INJECT_SORT_EXPERIMENT): pairs sorted by (level, z, y, x) with a
// library sort and a host read-back of the count, ONLY to measure what k_inject gains from voxel-ordered pairs
static void experiment_sort_pairs(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    static vgi_pair_t* alt = nullptr;
    static void* tmp = nullptr;
    static size_t tmpBytes = 0;
    uint32_t count = 0;
    cudaMemcpyAsync(&count, &c->counters->pairs, 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (count > bp.max_pairs) count = bp.max_pairs;
    if (!count) return;
    if (!alt) cudaMalloc(&alt, (size_t)bp.max_pairs * sizeof(vgi_pair_t));
    size_t need = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, c->pairs, alt, (int)count, 0, 30, s);
    if (need > tmpBytes) { cudaFree(tmp); cudaMalloc(&tmp, need); tmpBytes = need; }
    cub::DeviceRadixSort::SortKeys(tmp, need, c->pairs, alt, (int)count, 0, 30, s);
    cudaMemcpyAsync(c->pairs, alt, (size_t)count * sizeof(vgi_pair_t), cudaMemcpyDeviceToDevice, s);
}
#endif

static int launch_inject(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    int n = 0;
#ifdef VGI_INJECT_SORT_EXPERIMENT
    experiment_sort_pairs(c, bp, s);
#endif
    LAUNCH("k_zero_acc", k_zero_acc<<<148 * 8, 256, 0, s>>>(c->acc, c->counters, bp.max_occ));
    if (bp.ntri && bp.level_mask) {
        if (c->scene_max_texture > -1) {
            LAUNCH("k_inject", k_inject<true><<<148 * VGI_INJECT_MINBLOCKS * 2, 256, 0, s>>>(bp, c->light, c->tri_pos, c->tri_nrm, c->materials, c->pairs,
                                                                                        c->occ, c->occ_prefix, c->acc, c->counters, c->texset()));
        } else {
            LAUNCH("k_inject", k_inject<false><<<148 * VGI_INJECT_MINBLOCKS * 2, 256, 0, s>>>(bp, c->light, c->tri_pos, c->tri_nrm, c->materials, c->pairs,
                                                                                         c->occ, c->occ_prefix, c->acc, c->counters, c->texset()));
        }
    }
    return n;
}

// side stream for the mask kernels (fork: after the occupancy is final; join: before the records need the visit lists)
static cudaStream_t side_fork(vgi_ctx* c, cudaStream_t s)
{
    if (c->timer.enabled) return s;   // per-kernel timing (vgi_set_timing): one stream, so that no kernel's time includes a neighbour's
    if (!c->side_stream) {
        if (cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess) { c->side_stream = nullptr; return s; }
        cudaEventCreateWithFlags(&c->ev_side_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_side_masks, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_side_done, cudaEventDisableTiming);
    }
    cudaEventRecord(c->ev_side_fork, s);
    cudaStreamWaitEvent(c->side_stream, c->ev_side_fork, 0);
    return c->side_stream;
}

int vgi_launch_inject_finalize(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    int n = 0;
    // nz masks ping-pong between frames: nz[cur] is written by this build, nz[cur ^ 1] is last frame's
    const int cur = c->nz_cur ^ 1;
    // k_level_masks and k_brick_mask need the occupancy bits only: they run beside k_inject / k_level_records
    const cudaStream_t side = side_fork(c, s);
    launch_masks(c, bp, cur, side, n);
    if (side != s) cudaEventRecord(c->ev_side_masks, side);
    launch_brick(c, bp, cur, side, n);
    if (side != s) cudaEventRecord(c->ev_side_done, side);
    n += launch_inject(c, bp, s);
    if (side != s) cudaStreamWaitEvent(s, c->ev_side_masks, 0);
    const SlabPack none = { nullptr, nullptr, nullptr, 0u };
    for (int l = bp.level_first; l < bp.L; ++l)
        LAUNCH("k_level_records", k_level_records<0><<<148 * 8, 128, 0, s>>>(bp, l, c->occ, c->occ_prefix, c->acc, c->nz[cur], c->visit_list, c->visit_cap, c->counters, c->store, none));
    if (side != s) cudaStreamWaitEvent(s, c->ev_side_done, 0);
    c->nz_cur = cur;
    cudaMemcpyAsync(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, s);
    return n;
}

// ---- slab-sharded build (multi-GPU) --------------------------------------------
// records received from the other GPUs -> store
__global__ void __launch_bounds__(256) k_slab_unpack(int R, const uint32_t* __restrict__ ids, const uint4* __restrict__ recs, uint32_t count,
                                                      VoxelRecord* __restrict__ store)
{
    const size_t nvox = (size_t)R * R * R;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t id = ids[i];
        uint4* dst = reinterpret_cast<uint4*>(store + (size_t)(id >> 27) * nvox + (id & 0x07ffffffu));
        dst[0] = recs[2 * (size_t)i];
        dst[1] = recs[2 * (size_t)i + 1];
    }
}

int vgi_launch_slab_begin(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    int n = vgi_launch_voxelize(c, bp, s);
    n += launch_inject(c, bp, s);
    return n;
}

int vgi_launch_slab_finalize(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    int n = 0;
    const int cur = c->nz_cur ^ 1;
    cudaMemsetAsync(c->slab_count, 0, sizeof(uint32_t), s);
    launch_masks(c, bp, cur, s, n);
    const SlabPack pack = { c->slab_ids, c->slab_recs, c->slab_count, c->slab_cap };
    for (int l = 0; l < bp.L; ++l)
        LAUNCH("k_level_records_slab", k_level_records<1><<<148 * 8, 128, 0, s>>>(bp, l, c->occ, c->occ_prefix, c->acc, c->nz[cur], c->visit_list, c->visit_cap, c->counters, c->store, pack));
    cudaMemcpyAsync(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, s);
    return n;
}

int vgi_launch_slab_unpack(vgi_ctx* c, const uint32_t* ids, const uint4* recs, uint32_t count, cudaStream_t s)
{
    int n = 0;
    if (!count) return 0;
    LAUNCH("k_slab_unpack", k_slab_unpack<<<min(cdiv(count, 256), 148u * 8u), 256, 0, s>>>((int)c->cfg.resolution, ids, recs, count, c->store));
    return n;
}

int vgi_launch_slab_end(vgi_ctx* c, const BuildParams& bp, cudaStream_t s)
{
    int n = 0;
    const int cur = c->nz_cur ^ 1;
    const SlabPack none = { nullptr, nullptr, nullptr, 0u };
    for (int l = 1; l < bp.L; ++l)
        LAUNCH("k_level_records_mip", k_level_records<2><<<148 * 8, 128, 0, s>>>(bp, l, c->occ, c->occ_prefix, c->acc, c->nz[cur], c->visit_list, c->visit_cap, c->counters, c->store, none));
    launch_brick(c, bp, cur, s, n);
    c->nz_cur = cur;
    return n;
}

// ---- peer build: the slab-sharded build with the exchange done by the kernels themselves ------------
// Every GPU maps the voxel store, the occupancy words and a row of sync flags of every other GPU (CUDA IPC over
// NVLink / NVSwitch). A GPU voxelizes and injects its own slab of texel planes, stores its occupancy words and its
// finalized records directly into all stores, and the GPUs meet at flag barriers: no NCCL call, no host round trip,
// no staging buffers. The result equals vgi_build_clipmap on one GPU bit for bit.

// All GPUs arrive: slot [rank] of every GPU's flag row receives this barrier's epoch (release, system scope), then
// every GPU waits for all slots of its own row. Preceded in stream order by the kernels whose stores it publishes.
__global__ void k_peer_barrier(PeerSet ps, uint32_t epoch, Counters* cnt)
{
    __threadfence_system();
    const int r = threadIdx.x;
    if (r < ps.n) {
        uint32_t* remote = ps.flags[r] + ps.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(remote), "r"(epoch) : "memory");
        const uint32_t* mine = ps.flags[ps.rank] + r;
        const long long t0 = clock64();
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (clock64() - t0 > 6000000000ll) { atomicOr(&cnt->overflow, 32u); break; } // ~3 s: a peer never arrived
        }
    }
    __syncthreads();
    __threadfence_system();
}

// The occupancy words of the texel planes this GPU owns (a plane of R x R bits = R * R / 32 consecutive words, a
// multiple of four for every supported R): cleared before the voxelizer, then stored into every other GPU's occupancy
// array after it (whole words, zeros included) with 16-byte stores.
template <bool PUSH>
__global__ void __launch_bounds__(256) k_peer_own_planes(PeerSet ps, BuildParams bp, uint32_t* __restrict__ occ)
{
    const uint32_t plane4 = (uint32_t)(((size_t)bp.R * bp.R) >> 7);         // uint4 per plane
    const uint32_t wordsPerLevel = (uint32_t)(((size_t)bp.R * bp.R * bp.R) >> 5);
    const size_t total = (size_t)plane4 * bp.R * bp.L;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = (uint32_t)(i % plane4);
        const int z = (int)((i / plane4) % bp.R), l = (int)(i / ((size_t)plane4 * bp.R));
        if (!owns_plane(bp, z)) continue;
        const size_t idx = (size_t)l * wordsPerLevel + (((size_t)z * plane4 + q) << 2);
        if (PUSH) {
            const uint4 v = *reinterpret_cast<const uint4*>(occ + idx);
            for (int r = 0; r < ps.n; ++r)
                if (r != ps.rank) *reinterpret_cast<uint4*>(ps.occ[r] + idx) = v;
        } else {
            *reinterpret_cast<uint4*>(occ + idx) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

int vgi_launch_peer_build(vgi_ctx* c, const BuildParams& bp, const PeerSet& ps, uint32_t* epoch, cudaStream_t s)
{
    int n = 0;
    const uint32_t wordsPerLevel = (uint32_t)(((size_t)bp.R * bp.R * bp.R) >> 5);
    const size_t nwords = (size_t)wordsPerLevel * bp.L;
    // own slab: voxelize, scan, inject. Only the own planes are cleared: the other planes of the occupancy array are
    // overwritten word by word by their owners every frame.
    cudaMemsetAsync(c->counters, 0, sizeof(Counters), s);
    LAUNCH("k_peer_clear_occ", k_peer_own_planes<false><<<148 * 4, 256, 0, s>>>(ps, bp, c->occ));
    if (bp.ntri) {
        if (c->scene_max_texture > -1) {
            LAUNCH("k_voxelize", k_voxelize<true><<<cdiv((size_t)bp.ntri * bp.L, 128), 128, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
            LAUNCH("k_voxelize_large", k_voxelize_large<true><<<148 * 4, 256, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
        } else {
            LAUNCH("k_voxelize", k_voxelize<false><<<cdiv((size_t)bp.ntri * bp.L, 128), 128, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
            LAUNCH("k_voxelize_large", k_voxelize_large<false><<<148 * 4, 256, 0, s>>>(bp, c->tri_pos, c->occ, c->pairs, c->large, c->counters, c->materials, c->texset()));
        }
    }
    const unsigned nblk = cdiv(nwords, SCAN_BLOCK * SCAN_ITEMS);
    LAUNCH("k_scan_block_sums", k_scan_block_sums<<<nblk, SCAN_BLOCK, 0, s>>>(c->occ, nwords, c->block_sums));
    LAUNCH("k_scan_sums", k_scan_sums<<<1, 1024, 0, s>>>(c->block_sums, nblk, &c->counters->occ_total));
    LAUNCH("k_scan_final", k_scan_final<<<nblk, SCAN_BLOCK, 0, s>>>(c->occ, nwords, c->block_sums, c->occ_prefix));
    // barrier A: every GPU is done with last frame's volume and with its own scan -> occupancy words may travel
    LAUNCH("k_peer_barrier", k_peer_barrier<<<1, 32, 0, s>>>(ps, ++*epoch, c->counters));
    LAUNCH("k_peer_push_occ", k_peer_own_planes<true><<<148 * 4, 256, 0, s>>>(ps, bp, c->occ));
    // barrier B: the occupancy of the whole volume is in place on every GPU
    LAUNCH("k_peer_barrier", k_peer_barrier<<<1, 32, 0, s>>>(ps, ++*epoch, c->counters));
    // the masks of the whole volume (replicated bit arithmetic) run on the side stream beside the own slab's injection
    const int cur = c->nz_cur ^ 1;
    const cudaStream_t side = side_fork(c, s);
    launch_masks(c, bp, cur, side, n);
    if (side != s) cudaEventRecord(c->ev_side_masks, side);
    launch_brick(c, bp, cur, side, n);
    if (side != s) cudaEventRecord(c->ev_side_done, side);
    n += launch_inject(c, bp, s);
    if (side != s) cudaStreamWaitEvent(s, c->ev_side_masks, 0);
    SlabPack pack = { nullptr, nullptr, nullptr, 0u };
    for (int r = 0; r < ps.n; ++r) pack.peer_store[r] = ps.store[r];
    pack.npeers = ps.n;
    LAUNCH("k_level_records_peer", k_level_records<3><<<dim3(148 * 2, bp.L), 128, 0, s>>>(bp, 0, c->occ, c->occ_prefix, c->acc, c->nz[cur], c->visit_list, c->visit_cap, c->counters, c->store, pack));
    // barrier C: every record of every slab has reached every store
    LAUNCH("k_peer_barrier", k_peer_barrier<<<1, 32, 0, s>>>(ps, ++*epoch, c->counters));
    const SlabPack none = { nullptr, nullptr, nullptr, 0u };
    for (int l = 1; l < bp.L; ++l)
        LAUNCH("k_level_records_mip", k_level_records<2><<<148 * 8, 128, 0, s>>>(bp, l, c->occ, c->occ_prefix, c->acc, c->nz[cur], c->visit_list, c->visit_cap, c->counters, c->store, none));
    if (side != s) cudaStreamWaitEvent(s, c->ev_side_done, 0);
    c->nz_cur = cur;
    cudaMemcpyAsync(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, s);
    return n;
}

int vgi_launch_export(vgi_ctx* c, int which, uint8_t* dst, int literal_border, cudaStream_t s)
{
    const int R = (int)c->cfg.resolution, L = (int)c->cfg.level_count;
    int logR = 0;
    while ((1 << logR) < R) ++logR;
    const size_t n = (size_t)(R + 2) * 6 * (size_t)(R + 2) * L * (R + 2);
    k_export<<<cdiv(n, 256), 256, 0, s>>>(R, L, logR, which, literal_border, c->store, reinterpret_cast<uint32_t*>(dst));
    return 1;
}
