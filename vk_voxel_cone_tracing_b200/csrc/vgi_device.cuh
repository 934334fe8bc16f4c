// vgi_device.cuh — device helpers shared by the strict (-fmad=false) translation units vgi_build.cu and
// vgi_svo.cu: the conservative triangle/voxel test, the injection sample point, the literal shadow
// visibility (scalar and quad-cooperative) and the popcount prefix scan. Everything that feeds a
// quantised result follows the IEEE binary32 / no-FMA / left-to-right contract of DESIGN.md "numerics".
#pragma once
#include "vgi_internal.h"

#define DEVFN static __device__ __forceinline__

// ---------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------
DEVFN float f_min(float a, float b) { return a < b ? a : b; }
DEVFN float f_max(float a, float b) { return a > b ? a : b; }
DEVFN float f_clamp(float x, float lo, float hi) { return f_min(f_max(x, lo), hi); }
DEVFN float f_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
DEVFN float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
DEVFN float unorm8_to_f(uint32_t c) { return (float)c / 255.0f; }
DEVFN uint32_t f_to_unorm8(float x)
{
    if (!(x > 0.0f)) return 0u;
    if (x > 1.0f) x = 1.0f;
    return (uint32_t)(x * 255.0f + 0.5f);
}
DEVFN unsigned lane_id() { return threadIdx.x & 31u; }

// Conservative triangle/voxel coverage (DESIGN.md "canonical coverage"; Schwarz-Seidel test in voxel
// units). ref: msaaVoxelizer.geom:27-49 / msaaVoxelizer.frag:43-73 with the raster coverage (Q3)
// replaced by exact overlap.
struct TriSetup {
    float n[3], d1, d2;
    float ne[3][3][2];
    float de[3][3];
    int lo[3], hi[3];
    bool valid;
};

// cross(p1-p0, p2-p0) and dominant axis (ref: msaaVoxelizer.geom:27-32: ties -> z, then y)
DEVFN int cross_and_axis(const float p[9], float N[3])
{
    const float a0 = p[3] - p[0], a1 = p[4] - p[1], a2 = p[5] - p[2];
    const float b0 = p[6] - p[0], b1 = p[7] - p[1], b2 = p[8] - p[2];
    N[0] = a1 * b2 - a2 * b1;
    N[1] = a2 * b0 - a0 * b2;
    N[2] = a0 * b1 - a1 * b0;
    const float ax = fabsf(N[0]), ay = fabsf(N[1]), az = fabsf(N[2]);
    return (ax > ay && ax > az) ? 0 : ((ay > az) ? 1 : 2);
}

DEVFN void tri_setup_grid(TriSetup& ts, const float q[3][3], const float N[3], const int clipLo[3], const int clipHi[3])
{
    float e[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) e[i][k] = q[(i + 1) % 3][k] - q[i][k];
    float* n = ts.n;
    n[0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    n[1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    n[2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    bool valid = !((n[0] == 0.0f && n[1] == 0.0f && n[2] == 0.0f) || (N[0] == 0.0f && N[1] == 0.0f && N[2] == 0.0f));
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (!(fabsf(q[i][k]) < 1.0e9f)) valid = false;
    ts.valid = valid;
    if (!valid) return;

    const float c0 = n[0] > 0.0f ? 1.0f : 0.0f, c1 = n[1] > 0.0f ? 1.0f : 0.0f, c2 = n[2] > 0.0f ? 1.0f : 0.0f;
    ts.d1 = (n[0] * (c0 - q[0][0]) + n[1] * (c1 - q[0][1])) + n[2] * (c2 - q[0][2]);
    ts.d2 = (n[0] * ((1.0f - c0) - q[0][0]) + n[1] * ((1.0f - c1) - q[0][1])) + n[2] * ((1.0f - c2) - q[0][2]);
    // plane 0: xy (sign n.z), plane 1: yz (sign n.x), plane 2: zx (sign n.y)
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const int U = pl, V = (pl + 1) % 3, S = (pl + 2) % 3;
        const float s = n[S] >= 0.0f ? 1.0f : -1.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float nx = -e[i][V] * s;
            const float ny = e[i][U] * s;
            ts.ne[pl][i][0] = nx;
            ts.ne[pl][i][1] = ny;
            ts.de[pl][i] = (-(nx * q[i][U] + ny * q[i][V]) + f_max(0.0f, nx)) + f_max(0.0f, ny);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float mn = f_min(q[0][k], f_min(q[1][k], q[2][k]));
        const float mx = f_max(q[0][k], f_max(q[1][k], q[2][k]));
        const int lo = (int)floorf(mn), hi = (int)floorf(mx);
        ts.lo[k] = lo > clipLo[k] ? lo : clipLo[k];
        ts.hi[k] = hi < clipHi[k] ? hi : clipHi[k];
    }
}

DEVFN void tri_setup_level(TriSetup& ts, const float p[9], const float N[3], const LevelParams& lv, int R)
{
    float q[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) q[i][k] = p[i * 3 + k] / lv.voxel_size;
    const int lo[3] = { lv.min_corner[0], lv.min_corner[1], lv.min_corner[2] };
    const int hi[3] = { lv.min_corner[0] + R - 1, lv.min_corner[1] + R - 1, lv.min_corner[2] + R - 1 };
    tri_setup_grid(ts, q, N, lo, hi);
}

DEVFN bool tri_overlaps_voxel(const TriSetup& ts, int vx, int vy, int vz)
{
    const float f[3] = { (float)vx, (float)vy, (float)vz };
    const float np = (ts.n[0] * f[0] + ts.n[1] * f[1]) + ts.n[2] * f[2];
    if ((np + ts.d1) * (np + ts.d2) > 0.0f) return false;
    bool ok = true;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const int U = pl, V = (pl + 1) % 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if ((ts.ne[pl][i][0] * f[U] + ts.ne[pl][i][1] * f[V]) + ts.de[pl][i] < 0.0f) ok = false;
    }
    return ok;
}

DEVFN void load_tri(const float4* __restrict__ tri_pos, uint32_t t, float p[9], int* mat)
{
    const float4 a = __ldg(tri_pos + 3 * (size_t)t), b = __ldg(tri_pos + 3 * (size_t)t + 1), c = __ldg(tri_pos + 3 * (size_t)t + 2);
    p[0] = a.x; p[1] = a.y; p[2] = a.z;
    p[3] = b.x; p[4] = b.y; p[5] = b.z;
    p[6] = c.x; p[7] = c.y; p[8] = c.z;
    if (mat) *mat = __float_as_int(a.w);
}

// ---------------------------------------------------------------------------------------------------
// shading helpers shared by the injection kernel (ref: msaaInjectRadiance.frag:68-155, shadow.glsl:8-36)
// ---------------------------------------------------------------------------------------------------
DEVFN void xform_point(const float* m, const float* v, float* o)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = ((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r];
}

DEVFN float shadow_texel(const LightParams& lp, int x, int y)
{
    if (x < 0 || y < 0 || x >= lp.sw || y >= lp.sh) return 0.0f; // CLAMP_TO_BORDER, opaque black
    return __ldg(lp.depth + (size_t)y * lp.sw + x);
}

DEVFN float shadow_bilinear(const LightParams& lp, float u, float v, bool compare, float cmpz)
{
    const float x = u * (float)lp.sw - 0.5f, y = v * (float)lp.sh - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    const int ix = (int)f_clamp(fx, -4.0f, (float)lp.sw + 4.0f), iy = (int)f_clamp(fy, -4.0f, (float)lp.sh + 4.0f);
    float t00 = shadow_texel(lp, ix, iy), t10 = shadow_texel(lp, ix + 1, iy);
    float t01 = shadow_texel(lp, ix, iy + 1), t11 = shadow_texel(lp, ix + 1, iy + 1);
    if (compare) {
        t00 = t00 >= cmpz ? 1.0f : 0.0f; t10 = t10 >= cmpz ? 1.0f : 0.0f;
        t01 = t01 >= cmpz ? 1.0f : 0.0f; t11 = t11 >= cmpz ? 1.0f : 0.0f;
    }
    return (t00 * (1.0f - a) + t10 * a) * (1.0f - b) + (t01 * (1.0f - a) + t11 * a) * b;
}

// ref: shadow.glsl:28-36 + :14-26 (literal Q1: mean of 16 bilinear raw-depth taps)
DEVFN float calc_visibility(const LightParams& lp, const float* worldPos, bool compare)
{
    float l[3];
    xform_point(lp.view, worldPos, l);
    const float* P = lp.proj;
    float px = ((P[0] * l[0] + P[4] * l[1]) + P[8] * 0.0f) + P[12];
    float py = ((P[1] * l[0] + P[5] * l[1]) + P[9] * 0.0f) + P[13];
    px = px * 0.5f + 0.5f;
    py = py * 0.5f + 0.5f;
    float cmpz = 0.0f;
    if (compare) cmpz = (P[10] * l[2] + P[14]) - 0.002f;
    const float sx = 1.0f / (float)lp.sw, sy = 1.0f / (float)lp.sh;
    float sum = 0.0f;
    for (int j = 0; j < 4; ++j) {
        const float oy = -1.5f + (float)j;
        for (int i = 0; i < 4; ++i) {
            const float ox = -1.5f + (float)i;
            sum += shadow_bilinear(lp, px + ox * sx, py + oy * sy, compare, cmpz);
        }
    }
    return sum * 0.0625f;
}

// voxel centre projected along the dominant axis onto the triangle plane, clamped into the triangle
DEVFN bool inject_sample_at(int a, const float N[3], const float p[9], const float n9[9], float c[3],
                            float pos[3], float nrm[3], float* bary = nullptr)
{
    // component selects instead of dynamic indexing keep the vertex arrays in registers
#define SEL3(arr, base, i) ((i) == 0 ? (arr)[(base)] : ((i) == 1 ? (arr)[(base) + 1] : (arr)[(base) + 2]))
    const float Na = SEL3(N, 0, a);
    if (Na == 0.0f) return false;
    const float d[3] = { c[0] - p[0], c[1] - p[1], c[2] - p[2] };
    const float t = dot3(N, d) / Na;
    if (a == 0) c[0] = c[0] - t; else if (a == 1) c[1] = c[1] - t; else c[2] = c[2] - t;
    const int u = (a == 0) ? 1 : 0;
    const int v = (a == 2) ? 1 : 2;
    const float p0u = SEL3(p, 0, u), p0v = SEL3(p, 0, v);
    const float e1u = SEL3(p, 3, u) - p0u, e1v = SEL3(p, 3, v) - p0v;
    const float e2u = SEL3(p, 6, u) - p0u, e2v = SEL3(p, 6, v) - p0v;
    const float cu = SEL3(c, 0, u) - p0u, cv = SEL3(c, 0, v) - p0v;
#undef SEL3
    const float den = e1u * e2v - e2u * e1v;
    if (den == 0.0f) return false;
    float b1 = (cu * e2v - e2u * cv) / den;
    float b2 = (e1u * cv - cu * e1v) / den;
    float b0 = (1.0f - b1) - b2;
    b0 = f_max(b0, 0.0f); b1 = f_max(b1, 0.0f); b2 = f_max(b2, 0.0f);
    const float sum = (b0 + b1) + b2;
    b0 = b0 / sum; b1 = b1 / sum; b2 = b2 / sum;
    if (bary) { bary[0] = b0; bary[1] = b1; bary[2] = b2; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pos[k] = (p[k] * b0 + p[3 + k] * b1) + p[6 + k] * b2;
        if (n9) nrm[k] = (n9[k] * b0 + n9[3 + k] * b1) + n9[6 + k] * b2;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Material textures. ref: texture() / textureLod() of msaaVoxelizer.frag:64, msaaInjectRadiance.frag:73-82,131-136 and
// voxelizer.frag:52-76 on samplers created with REPEAT + LINEAR and maxLod = 0 (GLTFScene.cpp:339): a bi-linear read
// of level 0 with exact binary32 weights (unnormalised coordinate u * W - 0.5, texel = byte / 255).
// ---------------------------------------------------------------------------------------------------
DEVFN int tex_wrap(long long i, int n)
{
    if (i >= 0 && i < n) return (int)i;
    const long long m = i % n;
    return (int)(m < 0 ? m + n : m);
}

DEVFN void tex_fetch(const TexSet& ts, int tex, float u, float v, float o[4])
{
    const uint4 t = __ldg(ts.table + tex);
    const int W = (int)t.y, H = (int)t.z;
    const uint32_t* d = ts.data + t.x;
    const float ux = u * (float)W - 0.5f, uy = v * (float)H - 0.5f;
    const float fx = floorf(ux), fy = floorf(uy);
    const float wx = ux - fx, wy = uy - fy;
    const long long ix = (long long)fx, iy = (long long)fy;
    const int x0 = tex_wrap(ix, W), x1 = tex_wrap(ix + 1, W), y0 = tex_wrap(iy, H), y1 = tex_wrap(iy + 1, H);
    const uint32_t t00 = __ldg(d + (size_t)y0 * W + x0), t10 = __ldg(d + (size_t)y0 * W + x1);
    const uint32_t t01 = __ldg(d + (size_t)y1 * W + x0), t11 = __ldg(d + (size_t)y1 * W + x1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = (float)((t00 >> (8 * k)) & 0xffu) / 255.0f, b = (float)((t10 >> (8 * k)) & 0xffu) / 255.0f;
        const float c = (float)((t01 >> (8 * k)) & 0xffu) / 255.0f, e = (float)((t11 >> (8 * k)) & 0xffu) / 255.0f;
        const float r0 = a * (1.0f - wx) + b * wx;
        const float r1 = c * (1.0f - wx) + e * wx;
        o[k] = r0 * (1.0f - wy) + r1 * wy;
    }
}

// texture coordinate of a canonical sample: the vertices' coordinates weighted like its position
DEVFN void tri_uv_at(const TexSet& ts, uint32_t tri, const float* bary, float* uv)
{
    const float2 a = __ldg(ts.tri_uv + 3 * (size_t)tri), b = __ldg(ts.tri_uv + 3 * (size_t)tri + 1), c = __ldg(ts.tri_uv + 3 * (size_t)tri + 2);
    uv[0] = (a.x * bary[0] + b.x * bary[1]) + c.x * bary[2];
    uv[1] = (a.y * bary[0] + b.y * bary[1]) + c.y * bary[2];
}

// ref: msaaVoxelizer.frag:64 / msaaInjectRadiance.frag:73 / voxelizer.frag:52 — "occlusion texture .r < 0.1 -> discard",
// evaluated at the canonical sample of the (triangle, voxel) pair: the voxel centre c projected onto the triangle.
// false = the pair is discarded (also when the sample does not exist: degenerate projection).
DEVFN bool alpha_test_pair(const TexSet& ts, int occlusionTexture, uint32_t tri, int axis, const float N[3], const float p[9],
                           float c[3])
{
    float pos[3], bary[3], uv[2], t[4];
    if (!inject_sample_at(axis, N, p, nullptr, c, pos, nullptr, bary)) return true;    // no sample: nothing to test, the injection skips it too
    tri_uv_at(ts, tri, bary, uv);
    tex_fetch(ts, occlusionTexture, uv[0], uv[1], t);
    return !(t[0] < 0.1f);
}

// Per-pair shading state produced by the lane-per-pair phase of k_inject.
struct PairShade {
    float px, py, cmpz;     // shadow-map uv of the sample, compare depth (shadow_compare mode)
    float NdotL;
    float n[3];
    int   mat;
    uint32_t cell;          // accumulator index of the voxel
    int   kind;             // 0 = nothing, 1 = emissive (no visibility), 2 = lit (needs visibility)
};

// One bilinear tap exactly as shadow_bilinear() evaluates it, texels supplied by the caller.
DEVFN float bilinear_mix(float t00, float t10, float t01, float t11, float a, float b)
{
    return (t00 * (1.0f - a) + t10 * a) * (1.0f - b) + (t01 * (1.0f - a) + t11 * a) * b;
}

// Visibility of 8 pairs per warp round: the four lanes of a quad own the four tap columns of one
// pair (ref: shadow.glsl:14-26, 4x4 taps at -1.5..1.5 texels). The 16 bilinear footprints of a pair
// tile a 5x5 texel block: lane i loads column i (5 texels) and one texel of column 4, neighbours are
// exchanged by shuffles, so a pair costs 7 load instructions whose quad lanes share a 128-byte line
// instead of 64 scattered loads. Every tap keeps its own (ix, iy, a, b) and the taps are summed in
// the shader's order, so the result is bit-identical to calc_visibility(); pairs whose taps do not
// tile the block (float rounding at a texel boundary) take the scalar path.
// Split in two so that a caller can issue the loads of several rounds before consuming any of them.
struct QuadTexels {
    float c[5], e, e4;
    bool quad_ok;
};

struct QuadCoords {
    float a, b[4];
    int ix, iy[4];
};

DEVFN void quad_coords(const LightParams& lp, float px, float py, QuadCoords& q)
{
    const int i = (int)(lane_id() & 3u);
    const float sx = 1.0f / (float)lp.sw, sy = 1.0f / (float)lp.sh;
    const float ox = -1.5f + (float)i;
    const float x = (px + ox * sx) * (float)lp.sw - 0.5f;
    const float fx = floorf(x);
    q.a = x - fx;
    q.ix = (int)f_clamp(fx, -4.0f, (float)lp.sw + 4.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float oy = -1.5f + (float)j;
        const float y = (py + oy * sy) * (float)lp.sh - 0.5f;
        const float fy = floorf(y);
        q.b[j] = y - fy;
        q.iy[j] = (int)f_clamp(fy, -4.0f, (float)lp.sh + 4.0f);
    }
}

// stage 1: issue the block loads of this quad's pair (all lanes of the warp must call)
DEVFN void quad_visibility_load(const LightParams& lp, float px, float py, bool need, QuadTexels& t)
{
    const unsigned lane = lane_id();
    const int i = (int)(lane & 3u);
    const unsigned qbase = lane & ~3u;
    QuadCoords q;
    quad_coords(lp, px, py, q);
    const int ix0 = __shfl_sync(0xffffffffu, q.ix, qbase);
    const bool tiles = (q.ix == ix0 + i) && (q.iy[1] == q.iy[0] + 1) && (q.iy[2] == q.iy[0] + 2) && (q.iy[3] == q.iy[0] + 3);
    const unsigned okmask = __ballot_sync(0xffffffffu, tiles || !need);
    t.quad_ok = ((okmask >> qbase) & 0xfu) == 0xfu;
    const bool ld = need && t.quad_ok;
#pragma unroll
    for (int k = 0; k < 5; ++k) t.c[k] = ld ? shadow_texel(lp, ix0 + i, q.iy[0] + k) : 0.0f;
    t.e = ld ? shadow_texel(lp, ix0 + 4, q.iy[0] + i) : 0.0f;
    t.e4 = ld ? shadow_texel(lp, ix0 + 4, q.iy[0] + 4) : 0.0f;
}

// stage 2: exchange, filter and sum (all lanes of the warp must call)
DEVFN float quad_visibility_finish(const LightParams& lp, float px, float py, bool compare, float cmpz, bool need, QuadTexels& t)
{
    const unsigned lane = lane_id();
    const int i = (int)(lane & 3u);
    const unsigned qbase = lane & ~3u;
    QuadCoords q;
    quad_coords(lp, px, py, q);
    if (compare) {
#pragma unroll
        for (int k = 0; k < 5; ++k) t.c[k] = t.c[k] >= cmpz ? 1.0f : 0.0f;
        t.e = t.e >= cmpz ? 1.0f : 0.0f;
        t.e4 = t.e4 >= cmpz ? 1.0f : 0.0f;
    }
    float r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float right = __shfl_sync(0xffffffffu, t.c[k], (lane + 1u) & 31u);
        const float edge = __shfl_sync(0xffffffffu, t.e, qbase + (unsigned)(k < 4 ? k : 3));
        r[k] = (i < 3) ? right : (k < 4 ? edge : t.e4);
    }
    float tap[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tap[j] = bilinear_mix(t.c[j], r[j], t.c[j + 1], r[j + 1], q.a, q.b[j]);
    if (need && !t.quad_ok) {
        // scalar path for this column's four taps
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t00 = shadow_texel(lp, q.ix, q.iy[j]), t10 = shadow_texel(lp, q.ix + 1, q.iy[j]);
            float t01 = shadow_texel(lp, q.ix, q.iy[j] + 1), t11 = shadow_texel(lp, q.ix + 1, q.iy[j] + 1);
            if (compare) {
                t00 = t00 >= cmpz ? 1.0f : 0.0f; t10 = t10 >= cmpz ? 1.0f : 0.0f;
                t01 = t01 >= cmpz ? 1.0f : 0.0f; t11 = t11 >= cmpz ? 1.0f : 0.0f;
            }
            tap[j] = bilinear_mix(t00, t10, t01, t11, q.a, q.b[j]);
        }
    }
    // sum in the shader's order: rows j outer, columns i inner
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) sum += __shfl_sync(0xffffffffu, tap[j], qbase + (unsigned)k);
    return sum * 0.0625f;
}

DEVFN float quad_visibility(const LightParams& lp, float px, float py, bool compare, float cmpz, bool need)
{
    QuadTexels t;
    quad_visibility_load(lp, px, py, need, t);
    return quad_visibility_finish(lp, px, py, compare, cmpz, need, t);
}

// Visibility of the 32 pairs of a warp (one per lane, kind == lit where `lit` has the lane's bit): four
// rounds of eight quads; the loads of all rounds are issued before the first is consumed.
DEVFN float warp_visibility(const LightParams& lp, float px, float py, float cmpz, bool compare, unsigned lit)
{
    const unsigned lane = lane_id();
    float vis = 0.0f;
    if (!lit) return vis;
    QuadTexels t[4];
    float qpx[4], qpy[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned src = 8u * (unsigned)r + (lane >> 2);
        qpx[r] = __shfl_sync(0xffffffffu, px, src);
        qpy[r] = __shfl_sync(0xffffffffu, py, src);
        quad_visibility_load(lp, qpx[r], qpy[r], (lit >> src) & 1u, t[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned src = 8u * (unsigned)r + (lane >> 2);
        const float qcz = __shfl_sync(0xffffffffu, cmpz, src);
        const float v = quad_visibility_finish(lp, qpx[r], qpy[r], compare, qcz, (lit >> src) & 1u, t[r]);
        // hand the result back to the owning lane: lane l (in round l/8) reads quad l%8
        const float back = __shfl_sync(0xffffffffu, v, (lane & 7u) * 4u);
        if ((int)(lane >> 3) == r) vis = back;
    }
    return vis;
}

// Visibility of ONE pair on ONE lane (shipped since round 2; the quad-cooperative version above is kept for reference and
// A/B: it needs 45 warp instructions per pair, two thirds of them shuffles and replicated coordinate arithmetic, and was
// 67 % of k_inject). The 16 bilinear footprints of shadow.glsl:14-26 tile a 5 x 5 texel block whenever the tap coordinates
// round consistently; its rows are fetched as two ALIGNED float4 loads each (a row of five starting at x0 lies inside the
// eight texels from x0 & ~3: ten 16-byte loads instead of twenty-five scalar ones, one or two sectors per row) and the five
// texels are picked out of the eight registers with two rounds of selects on x0 & 3. Every tap keeps its own (a, b) and the
// taps are summed in the shader's order, so the result is bit-identical to calc_visibility(); pairs whose taps do not tile
// (rounding at a texel boundary) or maps whose width is not a multiple of four take the scalar path.
DEVFN float4 shadow_row4(const LightParams& lp, int x, int y)
{
    if (y < 0 || y >= lp.sh || x < 0 || x >= lp.sw) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // aligned chunk: all in or all out
    return __ldg(reinterpret_cast<const float4*>(lp.depth + (size_t)y * lp.sw + x));
}

DEVFN float lane_visibility(const LightParams& lp, float px, float py, float cmpz, bool compare)
{
    const float sx = 1.0f / (float)lp.sw, sy = 1.0f / (float)lp.sh;
    float a[4], b[4];
    int ix[4], iy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float o = -1.5f + (float)k;
        const float x = (px + o * sx) * (float)lp.sw - 0.5f;
        const float fx = floorf(x);
        a[k] = x - fx;
        ix[k] = (int)f_clamp(fx, -4.0f, (float)lp.sw + 4.0f);
        const float y = (py + o * sy) * (float)lp.sh - 0.5f;
        const float fy = floorf(y);
        b[k] = y - fy;
        iy[k] = (int)f_clamp(fy, -4.0f, (float)lp.sh + 4.0f);
    }
    const bool tiles = ix[1] == ix[0] + 1 && ix[2] == ix[0] + 2 && ix[3] == ix[0] + 3 &&
                       iy[1] == iy[0] + 1 && iy[2] == iy[0] + 2 && iy[3] == iy[0] + 3 && (lp.sw & 3) == 0 &&
                       ((reinterpret_cast<size_t>(lp.depth) & 15) == 0);
    float sum = 0.0f;
    if (tiles) {
        const int cx = ix[0] & ~3, sft = ix[0] - cx;
        float t[5][5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const float4 lo = shadow_row4(lp, cx, iy[0] + r), hi = shadow_row4(lp, cx + 4, iy[0] + r);
            const float v[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
            float w[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = (sft & 1) ? v[k + 1] : v[k];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                float x = (sft & 2) ? w[k + 2] : w[k];
                if (compare) x = x >= cmpz ? 1.0f : 0.0f;
                t[r][k] = x;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sum += bilinear_mix(t[j][i], t[j][i + 1], t[j + 1][i], t[j + 1][i + 1], a[i], b[j]);
    } else {
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) {
                float t00 = shadow_texel(lp, ix[i], iy[j]), t10 = shadow_texel(lp, ix[i] + 1, iy[j]);
                float t01 = shadow_texel(lp, ix[i], iy[j] + 1), t11 = shadow_texel(lp, ix[i] + 1, iy[j] + 1);
                if (compare) {
                    t00 = t00 >= cmpz ? 1.0f : 0.0f; t10 = t10 >= cmpz ? 1.0f : 0.0f;
                    t01 = t01 >= cmpz ? 1.0f : 0.0f; t11 = t11 >= cmpz ? 1.0f : 0.0f;
                }
                sum += bilinear_mix(t00, t10, t01, t11, a[i], b[j]);
            }
    }
    return sum * 0.0625f;
}

// ---------------------------------------------------------------------------------------------------
// K2: exclusive prefix sum of the occupancy popcounts (compact accumulator index per occupied voxel)
// ---------------------------------------------------------------------------------------------------
#define SCAN_BLOCK 1024
#define SCAN_ITEMS 4  // words per thread -> 4096 words per block

static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block_sums(const uint32_t* __restrict__ occ, size_t nwords,
                                                                 uint32_t* __restrict__ block_sums)
{
    __shared__ uint32_t warp_sums[SCAN_BLOCK / 32];
    const size_t base = ((size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    uint32_t s = 0;
    if (base + SCAN_ITEMS <= nwords) {
        const uint4 v = *reinterpret_cast<const uint4*>(occ + base);
        s = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    } else {
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < nwords) s += __popc(occ[base + i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane_id() == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = warp_sums[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

// single block: exclusive scan of block sums in place, grand total -> *total
static __global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ block_sums, uint32_t nblocks, uint32_t* __restrict__ total)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane_id() >= o) inc += n;
        }
        if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_tot[threadIdx.x], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
                if ((int)lane_id() >= o) winc += n;
            }
            warp_tot[threadIdx.x] = winc - w; // exclusive
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = carry + warp_tot[threadIdx.x >> 5] + (inc - v);
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_final(const uint32_t* __restrict__ occ, size_t nwords,
                                                            const uint32_t* __restrict__ block_sums,
                                                            uint32_t* __restrict__ prefix)
{
    __shared__ uint32_t warp_tot[SCAN_BLOCK / 32];
    const size_t base = ((size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        c[i] = (base + i < nwords) ? __popc(occ[base + i]) : 0u;
        s += c[i];
    }
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane_id() >= o) inc += n;
    }
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = warp_tot[threadIdx.x], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if ((int)lane_id() >= o) winc += n;
        }
        warp_tot[threadIdx.x] = winc - w;
    }
    __syncthreads();
    uint32_t run = block_sums[blockIdx.x] + warp_tot[threadIdx.x >> 5] + (inc - s);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < nwords) prefix[base + i] = run;
        run += c[i];
    }
}

