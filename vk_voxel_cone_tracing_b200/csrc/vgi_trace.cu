// vgi_trace.cu — per-pixel voxel cone tracing through the clipmap voxel store (sm_100a).
// ref: VFS/Shaders/voxelConeTracing.frag:143-414, brdf.glsl:30-78, shadow.glsl:8-36.
//
// Floating-point results here are judged with a tolerance (max abs 1e-3, PSNR >= 50 dB), so this
// translation unit is compiled with FMA contraction on. The hardware texture unit is NOT used:
// its 8-bit interpolation weights would break the tolerance against exact-weight filtering, and a
// manual tri-linear fetch lets all six face texels of a voxel live in one 32-byte record
// (DESIGN.md "data layout") so the three face taps of a cone share their sectors.
#include <cuda_fp16.h>

#include "vgi_internal.h"

#define DEVFN static __device__ __forceinline__

__constant__ float c_cones16[16][3] = { // ref: voxelConeTracing.frag:118-135
    { 0.57735f, 0.57735f, 0.57735f }, { 0.57735f, -0.57735f, -0.57735f },
    { -0.57735f, 0.57735f, -0.57735f }, { -0.57735f, -0.57735f, 0.57735f },
    { -0.903007f, -0.182696f, -0.388844f }, { -0.903007f, 0.182696f, 0.388844f },
    { 0.903007f, -0.182696f, 0.388844f }, { 0.903007f, 0.182696f, -0.388844f },
    { -0.388844f, -0.903007f, -0.182696f }, { 0.388844f, -0.903007f, 0.182696f },
    { 0.388844f, 0.903007f, -0.182696f }, { -0.388844f, 0.903007f, 0.182696f },
    { -0.182696f, -0.388844f, -0.903007f }, { 0.182696f, 0.388844f, -0.903007f },
    { -0.182696f, 0.388844f, 0.903007f }, { 0.182696f, -0.388844f, 0.903007f } };
__constant__ float c_cones32[32][3] = { // ref: voxelConeTracing.frag:81-114
    { 0.898904f, 0.435512f, 0.0479745f }, { 0.898904f, -0.435512f, -0.0479745f },
    { 0.898904f, 0.0479745f, -0.435512f }, { 0.898904f, -0.0479745f, 0.435512f },
    { -0.898904f, 0.435512f, -0.0479745f }, { -0.898904f, -0.435512f, 0.0479745f },
    { -0.898904f, 0.0479745f, 0.435512f }, { -0.898904f, -0.0479745f, -0.435512f },
    { 0.0479745f, 0.898904f, 0.435512f }, { -0.0479745f, 0.898904f, -0.435512f },
    { -0.435512f, 0.898904f, 0.0479745f }, { 0.435512f, 0.898904f, -0.0479745f },
    { -0.0479745f, -0.898904f, 0.435512f }, { 0.0479745f, -0.898904f, -0.435512f },
    { 0.435512f, -0.898904f, 0.0479745f }, { -0.435512f, -0.898904f, -0.0479745f },
    { 0.435512f, 0.0479745f, 0.898904f }, { -0.435512f, -0.0479745f, 0.898904f },
    { 0.0479745f, -0.435512f, 0.898904f }, { -0.0479745f, 0.435512f, 0.898904f },
    { 0.435512f, -0.0479745f, -0.898904f }, { -0.435512f, 0.0479745f, -0.898904f },
    { 0.0479745f, 0.435512f, -0.898904f }, { -0.0479745f, -0.435512f, -0.898904f },
    { 0.57735f, 0.57735f, 0.57735f }, { 0.57735f, 0.57735f, -0.57735f },
    { 0.57735f, -0.57735f, 0.57735f }, { 0.57735f, -0.57735f, -0.57735f },
    { -0.57735f, 0.57735f, 0.57735f }, { -0.57735f, 0.57735f, -0.57735f },
    { -0.57735f, -0.57735f, 0.57735f }, { -0.57735f, -0.57735f, -0.57735f } };

// development aid (VGI_TRACE_STATS=1): [0] steps, [1] level samples, [2] skipped by brick mask,
// [3] footprints loaded but all-zero, [4] corners loaded, [5] corners with any of the 3 face texels non-zero
__device__ unsigned long long g_trace_stats[8];
#ifdef VGI_TRACE_STATS_BUILD
#define STAT(i, n) atomicAdd(&g_trace_stats[i], (unsigned long long)(n))
extern "C" void vgi_debug_trace_stats(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_trace_stats, sizeof(g_trace_stats));
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(g_trace_stats, z, sizeof z); }
}
#else
#define STAT(i, n) ((void)0)
#endif

__constant__ float c_level_colors[7][4] = { {1,0,0,1},{0,1,0,1},{0,0,1,1},{1,1,0,1},{0,1,1,1},{1,0,1,1},{1,1,1,1} }; // ref: voxelConeTracing.frag:270-278

#define MIN_TRACE_STEP_FACTOR 0.2f
#define MAX_TRACE_DISTANCE 30.0f
#define MIN_SPECULAR_APERTURE 0.05f


DEVFN float f_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
DEVFN float f_fract(float x) { return x - floorf(x); }
DEVFN float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DEVFN void normalize3(const float* v, float* o)
{
    const float inv = 1.0f / sqrtf(dot3(v, v));
    o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv;
}

// byte k of a packed RGBA8 texel as an exact float: PRMT builds the bit pattern of 2^23 + byte, the
// packed add removes the 2^23 (both on the full-rate pipes; an I2F would go through the quarter-rate one)
DEVFN float2 unpack2(uint32_t t, uint32_t selLo, uint32_t selHi)
{
    const float2 m = make_float2(__uint_as_float(__byte_perm(t, 0x4B000000u, selLo)),
                                 __uint_as_float(__byte_perm(t, 0x4B000000u, selHi)));
    return __fadd2_rn(m, make_float2(-8388608.0f, -8388608.0f));
}

// one clipmap level, three face-weighted tri-linear taps (ref: voxelConeTracing.frag:313-327).
// Texel coordinate = fract(p / extent) * R - 0.5, indices wrapped modulo R: the toroidal addressing
// that the reference obtains from REPEAT + wrapped border texels.
// Empty space is skipped through the footprint byte (k_brick_mask keeps, for EVERY voxel, which of the 8 records of
// the footprint whose low corner is that voxel may be non-zero; round 1 tested a 4^3 brick bit first,
// VGI_TRACE_FP_ONLY = 0); only those records are loaded and filtered.
// Returns false (out = 0) when all eight records of the footprint are zero.
// Per-cone constants of the three face taps (ref: voxelConeTracing.frag:296-303, 318-326): which face of
// each +/- pair the cone reads and the squared direction weights (pre-divided by 255).
struct ConeFaces {
    bool negX, negY, negZ;      // travel direction negative -> odd face of the pair
    float kx, ky, kz;
};

DEVFN ConeFaces cone_faces(const float* dir)
{
    ConeFaces f;
    f.negX = !(dir[0] > 0.0f);
    f.negY = !(dir[1] > 0.0f);
    f.negZ = !(dir[2] > 0.0f);
    f.kx = (dir[0] * dir[0]) * (1.0f / 255.0f);
    f.ky = (dir[1] * dir[1]) * (1.0f / 255.0f);
    f.kz = (dir[2] * dir[2]) * (1.0f / 255.0f);
    return f;
}

// Layout experiment (development builds only, tools/build_variant.py -DVGI_TRACE_BRICK_STORE=1|2): the tracer reads a copy of
// the voxel store whose records are grouped in 4^3 bricks (2 KB, brick-major; inside a brick 1 = z, y, x order, 2 = Morton
// order so that an aligned 2x2x2 block is 256 contiguous bytes). The copy is made by k_brick_copy before every trace and is
// not part of any reported time; DESIGN.md section 4 has the A/B. 0 = the shipped linear (level, z, y, x) layout.
#ifndef VGI_TRACE_BRICK_STORE
#define VGI_TRACE_BRICK_STORE 0
#endif
DEVFN uint32_t rec_index(uint32_t lin, int logR)
{
#if VGI_TRACE_BRICK_STORE
    const uint32_t Rm = (1u << logR) - 1u, nb = (uint32_t)logR - 2u;
    const uint32_t x = lin & Rm, y = (lin >> logR) & Rm, zl = lin >> (2 * logR);    // zl = level << logR | z
    const uint32_t brick = ((((zl >> 2) << nb) + (y >> 2)) << nb) + (x >> 2);
#if VGI_TRACE_BRICK_STORE == 2
    const uint32_t in = (x & 1u) | ((y & 1u) << 1) | ((zl & 1u) << 2) | ((x & 2u) << 2) | ((y & 2u) << 3) | ((zl & 2u) << 4);
#else
    const uint32_t in = ((zl & 3u) << 4) | ((y & 3u) << 2) | (x & 3u);
#endif
    return (brick << 6) | in;
#else
    (void)logR;
    return lin;
#endif
}

#ifndef VGI_TRACE_GATHER_ALL
#define VGI_TRACE_GATHER_ALL 0
#endif
#ifndef VGI_TRACE_SETUP_ALL
#define VGI_TRACE_SETUP_ALL 0
#endif
#ifndef VGI_TRACE_SPEC_GATHER_ALL
#define VGI_TRACE_SPEC_GATHER_ALL 0
#endif
#ifndef VGI_TRACE_FP_ONLY
#define VGI_TRACE_FP_ONLY 1     // 1: the footprint byte is the only emptiness probe; 0: brick bit + footprint byte (round 1)
#endif
// Where a tri-linear footprint lies: texel index of its low corner, the fractional weights and the mask of
// the records that can be non-zero (0 = nothing to fetch).
struct Footprint {
    uint32_t vox;       // level << 3 logR | z << 2 logR | y << logR | x
    uint32_t mask;
    float    w[3];
};

// coordinates + the emptiness test (per-voxel footprint byte; with VGI_TRACE_FP_ONLY = 0 the 4^3 brick bit before it).
// posV = world position in level-0 voxels (pos * R / extent0); the texel coordinate of voxelConeTracing.frag:315-317,
// fract(pos / extent_l) * R - 0.5 = posV * 2^-level - 0.5 (mod R), is split into cell index and weight without
// FRND / F2I (quarter-rate pipe): adding 1.5 * 2^23 rounds (t - 0.5) to the nearest integer r, which is floor(t)
// except at exact ties, where (r, w = t - r) = (floor(t) - 1, 1) names the same tri-linear sample; the low mantissa bits
// of the sum are r in two's complement, so "& (R - 1)" is the toroidal wrap.
DEVFN void probe_level(const TraceParams& tp, const float* posV, int level, Footprint& fp)
{
    const int R = tp.R, Rm = R - 1, logR = tp.logR;
    const float sc = tp.level_scale[level]; // 2^-level
    const float MAGIC = 12582912.0f;        // 1.5 * 2^23
    uint32_t i0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float f = fmaf(posV[k], sc, MAGIC - 1.0f);
        i0[k] = __float_as_uint(f) & (uint32_t)Rm;
        const float r = f - MAGIC;
        fp.w[k] = fmaf(posV[k], sc, -r) - 0.5f;
    }
    STAT(1, 1);
    // 32-bit index arithmetic: L * R^3 <= 8 * 512^3 = 2^30
    const uint32_t nbShift = (uint32_t)logR - 2u, wprShift = (uint32_t)logR - 5u;
    const uint32_t bidx = (((((uint32_t)level << nbShift) + (i0[2] >> 2)) << nbShift) + (i0[1] >> 2) << wprShift) + (i0[0] >> 5);
    fp.vox = ((((((uint32_t)level << logR) + i0[2]) << logR) + i0[1]) << logR) + i0[0];
#if VGI_TRACE_FP_ONLY
    (void)bidx;
    fp.mask = __ldg(tp.footprint + fp.vox);     // valid for every voxel (k_brick_mask zeroes the bytes of emptied bricks)
    if (!fp.mask) STAT(3, 1);
#else
    // both lookups are issued together (one memory latency instead of two in the dependent chain); the
    // footprint byte is only meaningful where the brick bit is set
    const uint32_t bbyte = __ldg(tp.brick_mask + bidx);
    const uint32_t m = __ldg(tp.footprint + fp.vox);
    const bool brick = (bbyte >> ((i0[0] >> 2) & 7u)) & 1u;
    if (!brick) STAT(2, 1);
    else if (!m) STAT(3, 1);
    fp.mask = brick ? m : 0u;
#endif
}

// the filtered fetch of the non-zero records of a footprint: three face-weighted tri-linear taps
DEVFN void filter_footprint(const TraceParams& tp, const Footprint& fp, const ConeFaces& cf, float* out)
{
    const int R = tp.R, Rm = R - 1, logR = tp.logR;
    const uint32_t m = fp.mask;
    const uint32_t ix = fp.vox & (uint32_t)Rm, iy = (fp.vox >> logR) & (uint32_t)Rm, iz = (fp.vox >> (2 * logR)) & (uint32_t)Rm;
#if !VGI_TRACE_BRICK_STORE
    const VoxelRecord* base = tp.store + fp.vox;
#endif
    // record offsets of the +1 neighbours (toroidal); in records, not bytes: -(R-1) * R^2 * 32 overflows int at R = 512
    const int dx = (ix == (uint32_t)Rm) ? -Rm : 1;
    const int dy = ((iy == (uint32_t)Rm) ? -Rm : 1) << logR;
    const int dz = ((iz == (uint32_t)Rm) ? -Rm : 1) << (2 * logR);
    float2 aX0 = make_float2(0.f, 0.f), aX1 = aX0, aY0 = aX0, aY1 = aX0, aZ0 = aX0, aZ1 = aX0;
    STAT(4, __popc(m));
    // eight statically addressed corner blocks (offsets and weights are compile-time combinations), each
    // skipped when no lane of the warp needs it; the face pairs are fetched as three 8-byte loads at
    // immediate offsets from one address and the cone's sign picks the half
    const float* w = fp.w;
    const float wx0 = 1.0f - w[0], wy0 = 1.0f - w[1], wz0 = 1.0f - w[2];
    const float wxy[4] = { wx0 * wy0, w[0] * wy0, wx0 * w[1], w[0] * w[1] };
    const int oxy[4] = { 0, dx, dy, dx + dy };
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (!((m >> c) & 1u)) continue;
        const int off = oxy[c & 3] + ((c & 4) ? dz : 0);
        const float wc = wxy[c & 3] * ((c & 4) ? w[2] : wz0);
#if VGI_TRACE_BRICK_STORE
        const uint2* rec = reinterpret_cast<const uint2*>(tp.store + rec_index(fp.vox + (uint32_t)off, logR));
#else
        const uint2* rec = reinterpret_cast<const uint2*>(base + off);   // base is a VoxelRecord*: off counts records
#endif
        const uint2 fx = __ldg(rec), fy = __ldg(rec + 1), fz = __ldg(rec + 2);
        const uint32_t tx = cf.negX ? fx.y : fx.x;
        const uint32_t ty = cf.negY ? fy.y : fy.x;
        const uint32_t tz = cf.negZ ? fz.y : fz.x;
        STAT(5, (tx | ty | tz) != 0u);
        const float2 w2 = make_float2(wc, wc);
        aX0 = __ffma2_rn(w2, unpack2(tx, 0x7540u, 0x7541u), aX0);
        aX1 = __ffma2_rn(w2, unpack2(tx, 0x7542u, 0x7543u), aX1);
        aY0 = __ffma2_rn(w2, unpack2(ty, 0x7540u, 0x7541u), aY0);
        aY1 = __ffma2_rn(w2, unpack2(ty, 0x7542u, 0x7543u), aY1);
        aZ0 = __ffma2_rn(w2, unpack2(tz, 0x7540u, 0x7541u), aZ0);
        aZ1 = __ffma2_rn(w2, unpack2(tz, 0x7542u, 0x7543u), aZ1);
    }
    out[0] = aX0.x * cf.kx + aY0.x * cf.ky + aZ0.x * cf.kz;
    out[1] = aX0.y * cf.kx + aY0.y * cf.ky + aZ0.y * cf.kz;
    out[2] = aX1.x * cf.kx + aY1.x * cf.ky + aZ1.x * cf.kz;
    out[3] = aX1.y * cf.kx + aY1.y * cf.ky + aZ1.y * cf.kz;
}

// ceil(log2(dist / minRadius)) of voxelConeTracing.frag:367 clamped to [0, L-1], without sqrt / divide /
// log2: tp.min_level_dd[k] is the largest squared distance whose sqrtf(dd) / minRadius is still <= 2^k
// (bisected on the host over the same binary32 operations), so the level is a count of passed thresholds.
DEVFN float min_level_from_dd(const TraceParams& tp, float dd)
{
    // the thresholds ascend (+inf from index L-1 on): count the passed ones by bisection over t[0..6]
    const float* t = tp.min_level_dd;
    const bool b2 = dd > t[3];
    const bool b1 = dd > (b2 ? t[5] : t[1]);
    const bool b0 = dd > (b2 ? (b1 ? t[6] : t[4]) : (b1 ? t[2] : t[0]));
    return ((b2 ? 4.0f : 0.0f) + (b1 ? 2.0f : 0.0f)) + (b0 ? 1.0f : 0.0f);
}

// One marching step of voxelConeTracing.frag:361-389 at `step` (position, level selection, one or two
// level samples, front-to-back accumulation). lod = log2(diameter / voxelSize0).
struct ConeState {
    float result[4];
    float occlusion;
};

DEVFN void cone_step(const TraceParams& tp, ConeState& cs, const float* startPos, const float* dir, const ConeFaces& cf,
                     float startLevel, float step, float lod, float curSegmentLength)
{
    const vgi_vct_params& p = tp.p;
    STAT(0, 1);
    float position[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        position[k] = startPos[k] + dir[k] * step;
        d[k] = p.volume_center[k] - position[k];
    }
    const float minLevel = min_level_from_dd(tp, dot3(d, d));
    const float curLevel = fminf(fmaxf(fmaxf(startLevel, lod), minLevel), (float)(tp.L - 1));
    const float fl = floorf(curLevel);
    const float fr = curLevel - fl;
    const float posV[3] = { position[0] * tp.vox_scale0, position[1] * tp.vox_scale0, position[2] * tp.vox_scale0 };
    // probe both levels first (their mask lookups overlap), then fetch and filter what is not empty
    Footprint f0, f1;
    probe_level(tp, posV, (int)fl, f0);
    f1.mask = 0u;
    if (fr > 0.0f) probe_level(tp, posV, (int)fl + 1, f1); // Q17: floor == ceil when the level is integral
    const bool any = (f0.mask | f1.mask) != 0u;
    float smp[4] = { 0.f, 0.f, 0.f, 0.f };
    if (f0.mask) filter_footprint(tp, f0, cf, smp);
    if (fr > 0.0f && any) {
        float up[4] = { 0.f, 0.f, 0.f, 0.f };
        if (f1.mask) filter_footprint(tp, f1, cf, up);
#pragma unroll
        for (int c = 0; c < 4; ++c) smp[c] = smp[c] * (1.0f - fr) + up[c] * fr;
    }
    if (!any) return; // empty footprints: the accumulators would receive exact zeros
    const float voxelSize = p.voxel_size * exp2f(curLevel);
    const float correction = __fdividef(curSegmentLength, voxelSize);
    float opacity = 0.0f;
    if (smp[3] > 0.0f) // 1 - pow(1 - a, correction)
        opacity = f_clamp(1.0f - exp2f(correction * __log2f(1.0f - smp[3])), 0.0f, 1.0f);
    const float k1 = f_clamp(1.0f - cs.result[3], 0.0f, 1.0f) ;
    cs.result[0] += k1 * (smp[0] * correction);
    cs.result[1] += k1 * (smp[1] * correction);
    cs.result[2] += k1 * (smp[2] * correction);
    cs.result[3] += k1 * opacity;
    cs.occlusion += __fdividef((1.0f - cs.occlusion) * opacity, 1.0f + (step + voxelSize) * p.occlusion_decay);
}

// ref: voxelConeTracing.frag:341-392 — generic march (specular cone: per-pixel aperture)
DEVFN void trace_cone(const TraceParams& tp, const float* startPos_, const float* dir, float coneCoefficient, float maxDistance,
                      float startLevel, float stepFactor, float* out)
{
    const vgi_vct_params& p = tp.p;
    ConeState cs = { { 0.f, 0.f, 0.f, 0.f }, 0.0f };
    const float voxelSize0 = p.voxel_size * exp2f(startLevel);
    float startPos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) startPos[k] = startPos_[k] + dir[k] * voxelSize0 * p.trace_start_offset * 0.5f;
    float step = 0.0f;
    float diameter = fmaxf(step * coneCoefficient, p.voxel_size);
    const ConeFaces cf = cone_faces(dir);
    float curSegmentLength = voxelSize0;
    const float invVoxel = 1.0f / p.voxel_size;
    while (step < maxDistance && cs.occlusion < 1.0f) {
        cone_step(tp, cs, startPos, dir, cf, startLevel, step, __log2f(diameter * invVoxel), curSegmentLength);
        const float prevStep = step;
        step += fmaxf(diameter, p.voxel_size) * stepFactor;
        curSegmentLength = step - prevStep;
        diameter = step * coneCoefficient;
    }
    out[0] = cs.result[0]; out[1] = cs.result[1]; out[2] = cs.result[2];
    out[3] = 1.0f - cs.occlusion;
}

// Diffuse cones share aperture and step factor, so their step / diameter sequence is the same for every
// pixel and cone (it does not depend on the position): the block tabulates it once.
#define MAX_TABLE_STEPS 160   // 30 / (0.2 * voxelSize0-limited start) never needs more for stepFactor >= 0.2
struct StepTable {
    float step[MAX_TABLE_STEPS];
    float lod[MAX_TABLE_STEPS];
    int   n;
};

DEVFN void build_step_table(const TraceParams& tp, StepTable& t, float coneCoefficient, float stepFactor)
{
    const vgi_vct_params& p = tp.p;
    const float invVoxel = 1.0f / p.voxel_size;
    float step = 0.0f;
    float diameter = fmaxf(step * coneCoefficient, p.voxel_size);
    int n = 0;
    while (step < MAX_TRACE_DISTANCE && n < MAX_TABLE_STEPS) {
        t.step[n] = step;
        t.lod[n] = __log2f(diameter * invVoxel);
        ++n;
        step += fmaxf(diameter, p.voxel_size) * stepFactor;
        diameter = step * coneCoefficient;
    }
    t.n = (step < MAX_TRACE_DISTANCE) ? -1 : n; // -1: table too small, callers fall back to trace_cone
}

DEVFN void trace_cone_table(const TraceParams& tp, const StepTable& t, const float* startPos_, const float* dir, float startLevel, float* out)
{
    const vgi_vct_params& p = tp.p;
    ConeState cs = { { 0.f, 0.f, 0.f, 0.f }, 0.0f };
    const float voxelSize0 = p.voxel_size * exp2f(startLevel);
    float startPos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) startPos[k] = startPos_[k] + dir[k] * voxelSize0 * p.trace_start_offset * 0.5f;
    const ConeFaces cf = cone_faces(dir);
    float prevStep = 0.0f;
    for (int k = 0; k < t.n && cs.occlusion < 1.0f; ++k) {
        const float step = t.step[k];
        const float seg = k == 0 ? voxelSize0 : step - prevStep;
        cone_step(tp, cs, startPos, dir, cf, startLevel, step, t.lod[k], seg);
        prevStep = step;
    }
    out[0] = cs.result[0]; out[1] = cs.result[1]; out[2] = cs.result[2];
    out[3] = 1.0f - cs.occlusion;
}

// ---------------------------------------------------------------------------------------------------
// Warp-cooperative filtering. Lanes of a warp march the same cone direction from neighbouring pixels, and at
// the coarse levels (voxels of 0.5 .. 2 world units, 79 % of all corner fetches) their tri-linear footprints
// usually fall into ONE cell: every lane would fetch and unpack the same eight records and differ only in its
// weights. When all lanes that have something to fetch agree on (cell, cone), lanes 0..7 each fetch one corner
// record, blend its three face texels with the cone's direction weights into one float4 and park it in shared
// memory; every lane then forms its own weighted sum from eight broadcast LDS.128. Otherwise each lane filters
// its own footprint as before. Same arithmetic up to re-association (float tolerance work).
// ---------------------------------------------------------------------------------------------------
#define FULL_MASK 0xffffffffu
#ifndef VGI_TRACE_COOP
#define VGI_TRACE_COOP 1
#endif
#ifndef VGI_TRACE_COOP_MIN_LOD
#define VGI_TRACE_COOP_MIN_LOD 0.0f   // vote at every step (measured: 3.24 ms vs 3.27 ms when the finest steps skip the vote)
#endif

// one corner record of the shared cell, its three face texels blended with the cone's direction weights
DEVFN float4 coop_corner(const TraceParams& tp, uint32_t vox, uint32_t mask, unsigned corner, const ConeFaces& cf)
{
    if (!((mask >> corner) & 1u)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int R = tp.R, Rm = R - 1, logR = tp.logR;
    const uint32_t ix = vox & (uint32_t)Rm, iy = (vox >> logR) & (uint32_t)Rm, iz = (vox >> (2 * logR)) & (uint32_t)Rm;
    int off = 0;
    if (corner & 1u) off += (ix == (uint32_t)Rm) ? -Rm : 1;
    if (corner & 2u) off += ((iy == (uint32_t)Rm) ? -Rm : 1) << logR;
    if (corner & 4u) off += ((iz == (uint32_t)Rm) ? -Rm : 1) << (2 * logR);
    const uint2* rec = reinterpret_cast<const uint2*>(tp.store + rec_index(vox + (uint32_t)off, logR));
    const uint2 fx = __ldg(rec), fy = __ldg(rec + 1), fz = __ldg(rec + 2);
    const uint32_t tx = cf.negX ? fx.y : fx.x;
    const uint32_t ty = cf.negY ? fy.y : fy.x;
    const uint32_t tz = cf.negZ ? fz.y : fz.x;
    const float2 kx2 = make_float2(cf.kx, cf.kx), ky2 = make_float2(cf.ky, cf.ky), kz2 = make_float2(cf.kz, cf.kz);
    float2 lo = __fmul2_rn(kx2, unpack2(tx, 0x7540u, 0x7541u)), hi = __fmul2_rn(kx2, unpack2(tx, 0x7542u, 0x7543u));
    lo = __ffma2_rn(ky2, unpack2(ty, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(ky2, unpack2(ty, 0x7542u, 0x7543u), hi);
    lo = __ffma2_rn(kz2, unpack2(tz, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(kz2, unpack2(tz, 0x7542u, 0x7543u), hi);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// a lane's own weighted sum over the parked corner values of its cell
DEVFN void coop_gather(const Footprint& fp, uint32_t m, const float4* s_corner, float* out)
{
    const float* w = fp.w;
    const float wx0 = 1.0f - w[0], wy0 = 1.0f - w[1], wz0 = 1.0f - w[2];
    const float wxy[4] = { wx0 * wy0, w[0] * wy0, wx0 * w[1], w[0] * w[1] };
    float2 lo = make_float2(0.f, 0.f), hi = lo;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#if !VGI_TRACE_GATHER_ALL
        if (!((m >> c) & 1u)) continue;
#endif
        const float wc = wxy[c & 3] * ((c & 4) ? w[2] : wz0);
        const float4 v = s_corner[c];
        const float2 w2 = make_float2(wc, wc);
        lo = __ffma2_rn(w2, make_float2(v.x, v.y), lo);
        hi = __ffma2_rn(w2, make_float2(v.z, v.w), hi);
    }
    out[0] = lo.x; out[1] = lo.y; out[2] = hi.x; out[3] = hi.y;
}

// One level sample of a step. A vote decides whether the lanes with something to fetch fall into at most TWO
// (cell, cone) groups — those of the first and of the last such lane: a warp straddles at most one boundary of the
// cone-major work list, or one cell boundary — then lanes 0..7 fetch and blend one corner record of the first
// group's cell each and lanes 8..15 one of the second group's. Returns false when the lanes are more scattered
// (the caller filters per lane).
DEVFN bool coop_try(const TraceParams& tp, const Footprint& fp, int cone, const float (*cones)[3], float4* s_corner /* 16 */,
                    unsigned lane, float* out)
{
    const bool wanted = fp.mask != 0u;
    const unsigned want = __ballot_sync(FULL_MASK, wanted);
    if (!want) return true;
    const int la = __ffs(want) - 1, lb = 31 - __clz(want);
    const uint32_t voxA = __shfl_sync(FULL_MASK, fp.vox, la), voxB = __shfl_sync(FULL_MASK, fp.vox, lb);
    const int coneA = __shfl_sync(FULL_MASK, cone, la), coneB = __shfl_sync(FULL_MASK, cone, lb);
    const bool inA = fp.vox == voxA && cone == coneA;
    if (__ballot_sync(FULL_MASK, wanted && !(inA || (fp.vox == voxB && cone == coneB)))) return false;
    STAT(7, 1);
    const uint32_t mA = __shfl_sync(FULL_MASK, fp.mask, la), mB = __shfl_sync(FULL_MASK, fp.mask, lb); // mask = f(cell)
    const bool two = voxA != voxB || coneA != coneB;
    if (lane < (two ? 16u : 8u)) {
        const bool second = lane >= 8u;
        const int sc = second ? coneB : coneA;
        const float sdir[3] = { cones[sc][0], cones[sc][1], cones[sc][2] };
        s_corner[lane] = coop_corner(tp, second ? voxB : voxA, second ? mB : mA, lane & 7u, cone_faces(sdir));
    }
    __syncwarp();
    if (wanted) coop_gather(fp, fp.mask, s_corner + (inA ? 0 : 8), out);
    __syncwarp(); // the slots are rewritten by the next sample
    return true;
}

// The diffuse march of one warp: 32 (cone, pixel) items advance through the tabulated steps together so that the
// two level samples of a step can be filtered cooperatively. Per-lane arithmetic is that of cone_step.
DEVFN bool coop_try3(const TraceParams& tp, const Footprint& fp, uint32_t secBit, int c0, int c1, const float4* s_face,
                     float4* s_corner, unsigned lane, float* out);

// v3cone: c0 / c1 = first / last cone of the warp's slice of the work list, multi = more than two cones in it,
// s_face = per-block face table (nullptr: the original vote, coop_try)
DEVFN void march_warp_table(const TraceParams& tp, const StepTable& t, bool have, int cone, const float (*cones)[3], const float* startPos_,
                            const float* dir, float startLevel, float4* s_corner, unsigned lane, float* out,
                            int c0 = 0, int c1 = 0, bool multi = true, const float4* s_face = nullptr)
{
    const vgi_vct_params& p = tp.p;
    ConeState cs = { { 0.f, 0.f, 0.f, 0.f }, 0.0f };
    const float voxelSize0 = p.voxel_size * exp2f(startLevel);
    float startPos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) startPos[k] = startPos_[k] + dir[k] * voxelSize0 * p.trace_start_offset * 0.5f;
    const ConeFaces cf = cone_faces(dir);
    float prevStep = 0.0f;
    bool alive = have;
    const float topLevel = (float)(tp.L - 1);
    for (int k = 0; k < t.n; ++k) {
        if (!__any_sync(FULL_MASK, alive)) break;
        const float step = t.step[k];
        const float seg = k == 0 ? voxelSize0 : step - prevStep;
        prevStep = step;
        Footprint f0, f1;
        f0.mask = 0u; f1.mask = 0u; f0.vox = 0u; f1.vox = 0u; // the weights are read only where the mask is set
        float curLevel = 0.0f, fr = 0.0f;
#if VGI_TRACE_SETUP_ALL
        {   // finished lanes compute along (their indices stay in range) and drop their masks below: no divergent region
#else
        if (alive) {
#endif
            STAT(0, 1);
            float position[3], d[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                position[a] = startPos[a] + dir[a] * step;
                d[a] = p.volume_center[a] - position[a];
            }
            // the distance term can only raise the level: once the cone diameter alone selects the coarsest level
            // (a property of the step, the same for every lane) it is not evaluated
            const float lodk = t.lod[k];
            const float minLevel = lodk >= topLevel ? 0.0f : min_level_from_dd(tp, dot3(d, d));
            curLevel = fminf(fmaxf(fmaxf(startLevel, lodk), minLevel), topLevel);
            const float fl = floorf(curLevel);
            fr = curLevel - fl;
            const float posV[3] = { position[0] * tp.vox_scale0, position[1] * tp.vox_scale0, position[2] * tp.vox_scale0 };
            probe_level(tp, posV, (int)fl, f0);
            if (fr > 0.0f) probe_level(tp, posV, (int)fl + 1, f1); // Q17
        }
#if VGI_TRACE_SETUP_ALL
        if (!alive) { f0.mask = 0u; f1.mask = 0u; }
#endif
        float smp[4] = { 0.f, 0.f, 0.f, 0.f }, up[4] = { 0.f, 0.f, 0.f, 0.f };
        const bool vote = t.lod[k] >= VGI_TRACE_COOP_MIN_LOD;
        if (s_face) {
            const uint32_t secBit = (cone != c0) ? 0x80000000u : 0u;
            if (!(vote && !multi && coop_try3(tp, f0, secBit, c0, c1, s_face, s_corner, lane, smp)) && f0.mask) filter_footprint(tp, f0, cf, smp);
            if (!(vote && !multi && coop_try3(tp, f1, secBit, c0, c1, s_face, s_corner, lane, up)) && f1.mask) filter_footprint(tp, f1, cf, up);
        } else {
        if (!(vote && coop_try(tp, f0, cone, cones, s_corner, lane, smp)) && f0.mask) filter_footprint(tp, f0, cf, smp);
        if (!(vote && coop_try(tp, f1, cone, cones, s_corner, lane, up)) && f1.mask) filter_footprint(tp, f1, cf, up);
        }
        if ((f0.mask | f1.mask) != 0u) { // implies alive
            if (fr > 0.0f) {
#pragma unroll
                for (int c = 0; c < 4; ++c) smp[c] = smp[c] * (1.0f - fr) + up[c] * fr;
            }
            const float voxelSize = p.voxel_size * exp2f(curLevel);
            const float correction = __fdividef(seg, voxelSize);
            float opacity = 0.0f;
            if (smp[3] > 0.0f) opacity = f_clamp(1.0f - exp2f(correction * __log2f(1.0f - smp[3])), 0.0f, 1.0f);
            const float k1 = f_clamp(1.0f - cs.result[3], 0.0f, 1.0f);
            cs.result[0] += k1 * (smp[0] * correction);
            cs.result[1] += k1 * (smp[1] * correction);
            cs.result[2] += k1 * (smp[2] * correction);
            cs.result[3] += k1 * opacity;
            cs.occlusion += __fdividef((1.0f - cs.occlusion) * opacity, 1.0f + (step + voxelSize) * p.occlusion_decay);
            alive = cs.occlusion < 1.0f;
        }
    }
    out[0] = cs.result[0]; out[1] = cs.result[1]; out[2] = cs.result[2];
    out[3] = 1.0f - cs.occlusion;
}

// ---------------------------------------------------------------------------------------------------
// v2 march: no per-lane emptiness probes on the cooperative path. Per step every lane only computes where its
// one or two level samples lie (cell key + weights); a vote checks that the lanes' cells form at most two groups
// per level sample (first / last live lane: a warp straddles one cell boundary or one boundary of the cone-major
// work list), and ONE fetch pass serves both samples: lane j = 8 * group + corner tests the cell's brick bit and
// footprint byte, fetches its corner record, blends the three face texels with its group's cone weights and parks
// the float4 in shared memory (double-buffered per step). The ballot of that pass is the emptiness test of the
// whole step for the whole warp: all-zero steps cost no gather and no accumulation, and every lane then sums only
// the non-zero corners of its cell. Steps whose lanes are more scattered take the per-lane path of v1.
// ---------------------------------------------------------------------------------------------------
#ifndef VGI_TRACE_V2
#define VGI_TRACE_V2 3      // measured on B200, ms per 1080p frame: 0 = round-1 march 2.845, 1 = v2 2.94, 3 = v3 (round-1 march, one-key vote + face table) 2.825; with the footprint-only probe: 0 / 3 = 2.66, and a v5 (v3 + ONE fetch pass per step for both level samples, lanes 16-31 serving the high sample) 2.85 - removed
#endif

// per-cone constants of the fetch lanes, two float4 per cone in shared memory:
// [0] = word offsets of the three face texels inside a record (as bit patterns) , [1] = kx, ky, kz
DEVFN void cone_face_table(const float* dir, float4* dst)
{
    const ConeFaces f = cone_faces(dir);
    dst[0] = make_float4(__uint_as_float(f.negX ? 1u : 0u), __uint_as_float(f.negY ? 3u : 2u), __uint_as_float(f.negZ ? 5u : 4u), 0.0f);
    dst[1] = make_float4(f.kx, f.ky, f.kz, 0.0f);
}

// cell index + weights of a level sample (probe_level without the two mask loads)
DEVFN uint32_t level_cell(const TraceParams& tp, const float* posV, int level, float* w)
{
    const int Rm = tp.R - 1, logR = tp.logR;
    const float sc = tp.level_scale[level];
    const float MAGIC = 12582912.0f;
    uint32_t i0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float f = fmaf(posV[k], sc, MAGIC - 1.0f);
        i0[k] = __float_as_uint(f) & (uint32_t)Rm;
        const float r = f - MAGIC;
        w[k] = fmaf(posV[k], sc, -r) - 0.5f;
    }
    STAT(1, 1);
    return ((((((uint32_t)level << logR) + i0[2]) << logR) + i0[1]) << logR) + i0[0];
}

// fetch lane: corner `corner` of the cell `vox` for cone `cone`; false when that record is known to be zero
DEVFN bool coop_fetch(const TraceParams& tp, uint32_t vox, unsigned corner, const float4* s_face, int cone, float4& out)
{
    const int R = tp.R, Rm = R - 1, logR = tp.logR;
    const uint32_t ix = vox & (uint32_t)Rm, iy = (vox >> logR) & (uint32_t)Rm, iz = (vox >> (2 * logR)) & (uint32_t)Rm;
    const uint32_t level = vox >> (3 * logR);
    const uint32_t nbShift = (uint32_t)logR - 2u, wprShift = (uint32_t)logR - 5u;
    const uint32_t bidx = ((((level << nbShift) + (iz >> 2)) << nbShift) + (iy >> 2) << wprShift) + (ix >> 5);
    const uint32_t bbyte = __ldg(tp.brick_mask + bidx);
    const uint32_t m = __ldg(tp.footprint + vox);   // meaningful only where the brick bit is set
    const bool brick = (bbyte >> ((ix >> 2) & 7u)) & 1u;
    if (!(brick && ((m >> corner) & 1u))) return false;
    int off = 0;
    if (corner & 1u) off += (ix == (uint32_t)Rm) ? -Rm : 1;
    if (corner & 2u) off += ((iy == (uint32_t)Rm) ? -Rm : 1) << logR;
    if (corner & 4u) off += ((iz == (uint32_t)Rm) ? -Rm : 1) << (2 * logR);
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(tp.store + rec_index(vox + (uint32_t)off, logR));
    const float4 fo = s_face[2 * cone], fk = s_face[2 * cone + 1];
    const uint32_t tx = __ldg(rec + __float_as_uint(fo.x));
    const uint32_t ty = __ldg(rec + __float_as_uint(fo.y));
    const uint32_t tz = __ldg(rec + __float_as_uint(fo.z));
    STAT(4, 1);
    const float2 kx2 = make_float2(fk.x, fk.x), ky2 = make_float2(fk.y, fk.y), kz2 = make_float2(fk.z, fk.z);
    float2 lo = __fmul2_rn(kx2, unpack2(tx, 0x7540u, 0x7541u)), hi = __fmul2_rn(kx2, unpack2(tx, 0x7542u, 0x7543u));
    lo = __ffma2_rn(ky2, unpack2(ty, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(ky2, unpack2(ty, 0x7542u, 0x7543u), hi);
    lo = __ffma2_rn(kz2, unpack2(tz, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(kz2, unpack2(tz, 0x7542u, 0x7543u), hi);
    out = make_float4(lo.x, lo.y, hi.x, hi.y);
    return true;
}

// a lane's weighted sum over the parked non-zero corners (bits of m) of its cell
// ALL: every one of the eight parked slots is summed (the absent corners hold zeros): no branch per corner
template <bool ALL = false>
DEVFN void coop_gather_w(const float* w, uint32_t m, const float4* s_corner, float* out)
{
    const float wx0 = 1.0f - w[0], wy0 = 1.0f - w[1], wz0 = 1.0f - w[2];
    const float wxy[4] = { wx0 * wy0, w[0] * wy0, wx0 * w[1], w[0] * w[1] };
    float2 lo = make_float2(0.f, 0.f), hi = lo;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (!ALL && !((m >> c) & 1u)) continue;
        const float wc = wxy[c & 3] * ((c & 4) ? w[2] : wz0);
        const float4 v = s_corner[c];
        const float2 w2 = make_float2(wc, wc);
        lo = __ffma2_rn(w2, make_float2(v.x, v.y), lo);
        hi = __ffma2_rn(w2, make_float2(v.z, v.w), hi);
    }
    out[0] = lo.x; out[1] = lo.y; out[2] = hi.x; out[3] = hi.y;
}

// per-lane filter of a level sample from its cell index and weights (the v1 path, for scattered steps)
DEVFN bool lane_filter(const TraceParams& tp, uint32_t vox, const float* w, const ConeFaces& cf, float* out)
{
    const int Rm = tp.R - 1, logR = tp.logR;
    const uint32_t ix = vox & (uint32_t)Rm, iy = (vox >> logR) & (uint32_t)Rm, iz = (vox >> (2 * logR)) & (uint32_t)Rm;
    const uint32_t level = vox >> (3 * logR);
    const uint32_t nbShift = (uint32_t)logR - 2u, wprShift = (uint32_t)logR - 5u;
    const uint32_t bidx = ((((level << nbShift) + (iz >> 2)) << nbShift) + (iy >> 2) << wprShift) + (ix >> 5);
    const uint32_t bbyte = __ldg(tp.brick_mask + bidx);
    const uint32_t m = __ldg(tp.footprint + vox);
    Footprint fp;
    fp.vox = vox;
    fp.mask = ((bbyte >> ((ix >> 2) & 7u)) & 1u) ? m : 0u;
    fp.w[0] = w[0]; fp.w[1] = w[1]; fp.w[2] = w[2];
    if (!fp.mask) return false;
    filter_footprint(tp, fp, cf, out);
    return true;
}

// The diffuse march of one warp, v2. `sec` = this lane's cone is the second cone of the warp's slice of the work
// list (c0 = first, c1 = last); `multi` = the slice holds more than two cones (per-lane path throughout).
DEVFN void march_warp_v2(const TraceParams& tp, const StepTable& t, bool have, int cone, int c0, int c1, bool multi,
                         const float* startPos_, const float* dir, float startLevel, const float4* s_face, float4* s_coop /* 64 */,
                         unsigned lane, float* out)
{
    const vgi_vct_params& p = tp.p;
    ConeState cs = { { 0.f, 0.f, 0.f, 0.f }, 0.0f };
    const float voxelSize0 = p.voxel_size * exp2f(startLevel);
    float startPos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) startPos[k] = startPos_[k] + dir[k] * voxelSize0 * p.trace_start_offset * 0.5f;
    const uint32_t secBit = (cone != c0) ? 0x80000000u : 0u;
    float prevStep = 0.0f;
    bool alive = have;
    const float topLevel = (float)(tp.L - 1);
    const unsigned g = lane >> 3, corner = lane & 7u;
    for (int k = 0; k < t.n; ++k) {
        const unsigned aliveMask = __ballot_sync(FULL_MASK, alive);
        if (!aliveMask) break;
        const float step = t.step[k];
        const float seg = k == 0 ? voxelSize0 : step - prevStep;
        prevStep = step;
        uint32_t keyLo = 0u, keyHi = 0u;
        float wLo[3] = { 0.f, 0.f, 0.f }, wHi[3] = { 0.f, 0.f, 0.f };
        float curLevel = 0.0f, fr = 0.0f;
        if (alive) {
            STAT(0, 1);
            float position[3], d[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                position[a] = startPos[a] + dir[a] * step;
                d[a] = p.volume_center[a] - position[a];
            }
            const float lodk = t.lod[k];
            const float minLevel = lodk >= topLevel ? 0.0f : min_level_from_dd(tp, dot3(d, d));
            curLevel = fminf(fmaxf(fmaxf(startLevel, lodk), minLevel), topLevel);
            const float fl = floorf(curLevel);
            fr = curLevel - fl;
            const float posV[3] = { position[0] * tp.vox_scale0, position[1] * tp.vox_scale0, position[2] * tp.vox_scale0 };
            keyLo = level_cell(tp, posV, (int)fl, wLo) | secBit;
            if (fr > 0.0f) keyHi = level_cell(tp, posV, (int)fl + 1, wHi) | secBit; // Q17
        }
        const bool wantHi = alive && fr > 0.0f;
        const unsigned hiMask = __ballot_sync(FULL_MASK, wantHi);
        const int la = __ffs(aliveMask) - 1, lb = 31 - __clz(aliveMask);
        const uint32_t keyA = __shfl_sync(FULL_MASK, keyLo, la), keyB = __shfl_sync(FULL_MASK, keyLo, lb);
        const bool inA = keyLo == keyA;
        bool scattered = alive && !(inA || keyLo == keyB);
        uint32_t keyC = 0u, keyD = 0u;
        bool inC = false;
        if (hiMask) {
            const int ha = __ffs(hiMask) - 1, hb = 31 - __clz(hiMask);
            keyC = __shfl_sync(FULL_MASK, keyHi, ha);
            keyD = __shfl_sync(FULL_MASK, keyHi, hb);
            inC = keyHi == keyC;
            scattered = scattered || (wantHi && !(inC || keyHi == keyD));
        }
        float smp[4] = { 0.f, 0.f, 0.f, 0.f }, up[4] = { 0.f, 0.f, 0.f, 0.f };
        bool any = false;
        if (multi || __any_sync(FULL_MASK, scattered)) {
            STAT(6, 1);
            if (alive) {
                const float sdir[3] = { dir[0], dir[1], dir[2] };
                const ConeFaces cf = cone_faces(sdir);
                any = lane_filter(tp, keyLo & 0x7fffffffu, wLo, cf, smp);
                if (wantHi) any = lane_filter(tp, keyHi & 0x7fffffffu, wHi, cf, up) || any;
            }
        } else {
            STAT(7, 1);
            const uint32_t keyG = g == 0u ? keyA : (g == 1u ? keyB : (g == 2u ? keyC : keyD));
            const bool validG = g == 0u ? true : (g == 1u ? keyB != keyA : (g == 2u ? hiMask != 0u : (hiMask != 0u && keyD != keyC)));
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool nz = validG && coop_fetch(tp, keyG & 0x7fffffffu, corner, s_face, (keyG >> 31) ? c1 : c0, v);
            const unsigned nzMask = __ballot_sync(FULL_MASK, nz);
            if (nzMask) {
                float4* buf = s_coop + ((k & 1) << 5);
                if (nz) buf[lane] = v;
                __syncwarp();
                if (alive) {
                    const uint32_t mLo = (nzMask >> (inA ? 0 : 8)) & 0xffu;
                    if (mLo) { coop_gather_w(wLo, mLo, buf + (inA ? 0 : 8), smp); any = true; }
                    if (wantHi) {
                        const uint32_t mHi = (nzMask >> (inC ? 16 : 24)) & 0xffu;
                        if (mHi) { coop_gather_w(wHi, mHi, buf + (inC ? 16 : 24), up); any = true; }
                    }
                }
            } else {
                STAT(2, 1);
            }
        }
        if (any) {
            if (fr > 0.0f) {
#pragma unroll
                for (int c = 0; c < 4; ++c) smp[c] = smp[c] * (1.0f - fr) + up[c] * fr;
            }
            const float voxelSize = p.voxel_size * exp2f(curLevel);
            const float correction = __fdividef(seg, voxelSize);
            float opacity = 0.0f;
            if (smp[3] > 0.0f) opacity = f_clamp(1.0f - exp2f(correction * __log2f(1.0f - smp[3])), 0.0f, 1.0f);
            const float k1 = f_clamp(1.0f - cs.result[3], 0.0f, 1.0f);
            cs.result[0] += k1 * (smp[0] * correction);
            cs.result[1] += k1 * (smp[1] * correction);
            cs.result[2] += k1 * (smp[2] * correction);
            cs.result[3] += k1 * opacity;
            cs.occlusion += __fdividef((1.0f - cs.occlusion) * opacity, 1.0f + (step + voxelSize) * p.occlusion_decay);
            alive = cs.occlusion < 1.0f;
        }
    }
    out[0] = cs.result[0]; out[1] = cs.result[1]; out[2] = cs.result[2];
    out[3] = 1.0f - cs.occlusion;
}

// ---------------------------------------------------------------------------------------------------
// v3 = the shipped march (per-lane probes, vote among the lanes with a non-empty footprint) with a cheaper vote and fetch:
// (cell, "second cone of the warp's slice") travel as ONE 32-bit key (two shuffles instead of four), and the fetch lanes
// read their cone's face-word offsets and direction weights from the per-block table of v2 instead of re-deriving them
// (cone_faces was 2.5 % of the kernel's instructions, the vote 12 %). Slices with more than two cones filter per lane.
// ---------------------------------------------------------------------------------------------------
DEVFN bool coop_try3(const TraceParams& tp, const Footprint& fp, uint32_t secBit, int c0, int c1, const float4* s_face,
                     float4* s_corner /* 16 */, unsigned lane, float* out)
{
    const bool wanted = fp.mask != 0u;
    const unsigned want = __ballot_sync(FULL_MASK, wanted);
    if (!want) return true;
    const int la = __ffs(want) - 1, lb = 31 - __clz(want);
    const uint32_t key = fp.vox | secBit;
    const uint32_t keyA = __shfl_sync(FULL_MASK, key, la), keyB = __shfl_sync(FULL_MASK, key, lb);
    const bool inA = key == keyA;
    if (__ballot_sync(FULL_MASK, wanted && !(inA || key == keyB))) return false;
    STAT(7, 1);
    const uint32_t mA = __shfl_sync(FULL_MASK, fp.mask, la), mB = __shfl_sync(FULL_MASK, fp.mask, lb); // mask = f(cell)
    if (lane < (keyA != keyB ? 16u : 8u)) {
        const bool second = lane >= 8u;
        const uint32_t k = second ? keyB : keyA, m = second ? mB : mA;
        const unsigned corner = lane & 7u;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((m >> corner) & 1u) {
            const uint32_t vox = k & 0x7fffffffu;
            const int R = tp.R, Rm = R - 1, logR = tp.logR;
            const uint32_t ix = vox & (uint32_t)Rm, iy = (vox >> logR) & (uint32_t)Rm, iz = (vox >> (2 * logR)) & (uint32_t)Rm;
            int off = 0;
            if (corner & 1u) off += (ix == (uint32_t)Rm) ? -Rm : 1;
            if (corner & 2u) off += ((iy == (uint32_t)Rm) ? -Rm : 1) << logR;
            if (corner & 4u) off += ((iz == (uint32_t)Rm) ? -Rm : 1) << (2 * logR);
            const uint32_t* rec = reinterpret_cast<const uint32_t*>(tp.store + rec_index(vox + (uint32_t)off, logR));
            const int sc = (k >> 31) ? c1 : c0;
            const float4 fo = s_face[2 * sc], fk = s_face[2 * sc + 1];
            const uint32_t tx = __ldg(rec + __float_as_uint(fo.x)), ty = __ldg(rec + __float_as_uint(fo.y)), tz = __ldg(rec + __float_as_uint(fo.z));
            const float2 kx2 = make_float2(fk.x, fk.x), ky2 = make_float2(fk.y, fk.y), kz2 = make_float2(fk.z, fk.z);
            float2 lo = __fmul2_rn(kx2, unpack2(tx, 0x7540u, 0x7541u)), hi = __fmul2_rn(kx2, unpack2(tx, 0x7542u, 0x7543u));
            lo = __ffma2_rn(ky2, unpack2(ty, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(ky2, unpack2(ty, 0x7542u, 0x7543u), hi);
            lo = __ffma2_rn(kz2, unpack2(tz, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(kz2, unpack2(tz, 0x7542u, 0x7543u), hi);
            v = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        s_corner[lane] = v;
    }
    __syncwarp();
    if (wanted) coop_gather(fp, fp.mask, s_corner + (inA ? 0 : 8), out);
    __syncwarp(); // the slots are rewritten by the next sample
    return true;
}

// ref: voxelConeTracing.frag:394-414
DEVFN float calc_min_level(const TraceParams& tp, const float* worldPos)
{
    const vgi_vct_params& p = tp.p;
    const float d[3] = { p.volume_center[0] - worldPos[0], p.volume_center[1] - worldPos[1], p.volume_center[2] - worldPos[2] };
    const float dist = sqrtf(dot3(d, d));
    const float minRadius = p.voxel_size * p.volume_dimension * 0.5f;
    const float minLevel = fmaxf(log2f(dist / minRadius), 0.0f);
    const float radius = minRadius * exp2f(ceilf(minLevel));
    const float f = dist / radius;
    const float transitionStart = 0.5f;
    const float c = 1.0f / (1.0f - transitionStart);
    if (f > transitionStart) return ceilf(minLevel) + (f - transitionStart) * c;
    return ceilf(minLevel);
}

// shadow.glsl:8-36 (literal Q1: mean of 16 bilinear raw-depth taps, CLAMP_TO_BORDER black)
DEVFN float shadow_texel(const LightParams& lp, int x, int y)
{
    if (x < 0 || y < 0 || x >= lp.sw || y >= lp.sh) return 0.0f;
    return __ldg(lp.depth + (size_t)y * lp.sw + x);
}
DEVFN float calc_visibility(const TraceParams& tp, const float* worldPos)
{
    const LightParams& lp = tp.light;
    const float* V = lp.view;
    const float* P = lp.proj;
    float l[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) l[r] = V[r] * worldPos[0] + V[4 + r] * worldPos[1] + V[8 + r] * worldPos[2] + V[12 + r];
    float px = P[0] * l[0] + P[4] * l[1] + P[12];
    float py = P[1] * l[0] + P[5] * l[1] + P[13];
    px = px * 0.5f + 0.5f;
    py = py * 0.5f + 0.5f;
    const bool compare = tp.shadow_compare != 0;
    const float cmpz = (P[10] * l[2] + P[14]) - 0.002f;
    const float sx = 1.0f / (float)lp.sw, sy = 1.0f / (float)lp.sh;
    float sum = 0.0f;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            const float u = px + (-1.5f + (float)i) * sx, v = py + (-1.5f + (float)j) * sy;
            const float x = u * (float)lp.sw - 0.5f, y = v * (float)lp.sh - 0.5f;
            const float fx = floorf(x), fy = floorf(y);
            const float a = x - fx, b = y - fy;
            const int ix = (int)f_clamp(fx, -4.0f, (float)lp.sw + 4.0f), iy = (int)f_clamp(fy, -4.0f, (float)lp.sh + 4.0f);
            float t00 = shadow_texel(lp, ix, iy), t10 = shadow_texel(lp, ix + 1, iy);
            float t01 = shadow_texel(lp, ix, iy + 1), t11 = shadow_texel(lp, ix + 1, iy + 1);
            if (compare) {
                t00 = t00 >= cmpz ? 1.f : 0.f; t10 = t10 >= cmpz ? 1.f : 0.f;
                t01 = t01 >= cmpz ? 1.f : 0.f; t11 = t11 >= cmpz ? 1.f : 0.f;
            }
            sum += (t00 * (1.0f - a) + t10 * a) * (1.0f - b) + (t01 * (1.0f - a) + t11 * a) * b;
        }
    return sum * 0.0625f;
}

// ref: brdf.glsl:30-78
DEVFN void microfacet_brdf(float NdotL, float NdotV, float NdotH, float VdotH, float alphaRoughness,
                           const float* r0, float r90, const float* diffuseColor, float* o)
{
    const float PI_REF = 3.141592f;
    const float x = f_clamp(1.0f - VdotH, 0.0f, 1.0f);
    const float x2 = x * x;
    const float fw = x2 * x2 * x;
    const float r = alphaRoughness;
    const float attL = 2.0f * NdotL / (NdotL + sqrtf(r * r + (1.0f - r * r) * (NdotL * NdotL)));
    const float attV = 2.0f * NdotV / (NdotV + sqrtf(r * r + (1.0f - r * r) * (NdotV * NdotV)));
    const float G = attL * attV;
    const float rsq = r * r;
    const float ff = (NdotH * rsq - NdotH) * NdotH + 1.0f;
    const float Dm = rsq / (PI_REF * ff * ff);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float F = r0[k] + (r90 - r0[k]) * fw;
        const float diffuseContrib = (1.0f - F) * (diffuseColor[k] / PI_REF);
        const float specContrib = F * G * Dm / (4.0f * NdotL * NdotV);
        o[k] = NdotL * (diffuseContrib + specContrib);
    }
}

struct PixelSetup {
    float worldPos[3], view[3], normal[3];
    float diffuseColor[3], specularColor[3], emission[3];
    float perceptualRoughness, metallic;
    float minLevel;
    float startPos[3];
};

DEVFN bool pixel_setup(const TraceParams& tp, int px, int py, PixelSetup& s)
{
    const size_t pi = (size_t)py * tp.width + px;
    const float depth = __ldg(tp.depth + pi);
    if (depth == 1.0f) return false; // discard
    const float tcx = ((float)px + 0.5f) / (float)tp.width, tcy = ((float)py + 0.5f) / (float)tp.height;
    {   // worldPosFromDepth :305-311 (no y flip, Q18)
        const float v[4] = { tcx * 2.0f - 1.0f, tcy * 2.0f - 1.0f, depth, 1.0f };
        const float* M = tp.view_proj_inv;
        float o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r] = M[r] * v[0] + M[4 + r] * v[1] + M[8 + r] * v[2] + M[12 + r] * v[3];
        s.worldPos[0] = o[0] / o[3]; s.worldPos[1] = o[1] / o[3]; s.worldPos[2] = o[2] / o[3];
    }
    {
        const float d[3] = { tp.eye[0] - s.worldPos[0], tp.eye[1] - s.worldPos[1], tp.eye[2] - s.worldPos[2] };
        normalize3(d, s.view);
    }
    const uchar4 dif = __ldg(reinterpret_cast<const uchar4*>(tp.diffuse) + pi);
    const uchar4 spc = __ldg(reinterpret_cast<const uchar4*>(tp.specular) + pi);
    const uint2 nrmw = __ldg(reinterpret_cast<const uint2*>(tp.normal) + pi);
    const uint2 emiw = __ldg(reinterpret_cast<const uint2*>(tp.emission) + pi);
    s.diffuseColor[0] = dif.x / 255.0f; s.diffuseColor[1] = dif.y / 255.0f; s.diffuseColor[2] = dif.z / 255.0f;
    s.perceptualRoughness = dif.w / 255.0f;
    s.specularColor[0] = spc.x / 255.0f; s.specularColor[1] = spc.y / 255.0f; s.specularColor[2] = spc.z / 255.0f;
    s.metallic = spc.w / 255.0f;
    {
        const __half2 a = *reinterpret_cast<const __half2*>(&nrmw.x), b = *reinterpret_cast<const __half2*>(&nrmw.y);
        const float n[3] = { __low2float(a) * 2.0f - 1.0f, __high2float(a) * 2.0f - 1.0f, __low2float(b) * 2.0f - 1.0f };
        normalize3(n, s.normal);
        const __half2 e0 = *reinterpret_cast<const __half2*>(&emiw.x), e1 = *reinterpret_cast<const __half2*>(&emiw.y);
        s.emission[0] = __low2float(e0); s.emission[1] = __high2float(e0); s.emission[2] = __low2float(e1);
    }
    s.minLevel = calc_min_level(tp, s.worldPos);
    const float voxelSize = tp.p.voxel_size * exp2f(s.minLevel);
#pragma unroll
    for (int k = 0; k < 3; ++k) s.startPos[k] = s.worldPos[k] + s.normal[k] * voxelSize * tp.p.trace_start_offset;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// SVO cone tracing. ref: voxelConeTracing_Octree.frag:148-409 (bbox parameterised, Q14; same >>1 as the
// build, Q13). Samples are point descents through the node pool (no filtering), two per step.
// ---------------------------------------------------------------------------------------------------
DEVFN void svo_unpack(uint32_t y, float* o)
{
    o[0] = (float)(y & 0xffu) * (1.0f / 255.0f);
    o[1] = (float)((y >> 8) & 0xffu) * (1.0f / 255.0f);
    o[2] = (float)((y >> 16) & 0xffu) * (1.0f / 255.0f);
    o[3] = (float)(y >> 24) * (1.0f / 255.0f);
}

// Point samples at `level` (lo) and, when two = true, also at level + 1 (up) with ONE descent: the coarser
// sample is the node where the descent of sampleSVO (voxelConeTracing_Octree.frag:319-346) would stop for the
// coarser target resolution, the finer one is at most one step further down the same path.
DEVFN void sample_svo(const TraceParams& tp, const float* position, uint32_t level, bool two, float* lo, float* up)
{
    uint32_t resolution = (uint32_t)tp.p.volume_dimension;
    uint32_t fp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float scaled = (position[k] - tp.svo_center[k]) / tp.svo_extent;
        float g = ((scaled + 1.0f) * 0.5f) * (float)resolution;
        g = f_clamp(g, 0.0f, (float)resolution);
        fp[k] = tp.svo_literal ? (uint32_t)g : ((uint32_t)g) >> 1;     // canonical: the same halving as the build (Q13)
    }
    const uint32_t targetLo = 1u << level;
    const uint32_t targetFirst = two ? targetLo << 1 : targetLo;
    uint32_t idx = 0, cur = 0;
    do {
        resolution >>= 1;
        const uint32_t cx = fp[0] >= resolution, cy = fp[1] >= resolution, cz = fp[2] >= resolution;
        idx = cur + (cz | (cx << 1) | (cy << 2));
        cur = __ldg(&tp.svo_nodes[idx].x) & 0x7fffffffu;
        fp[0] -= cx * resolution; fp[1] -= cy * resolution; fp[2] -= cz * resolution;
    } while (cur != 0u && resolution > targetFirst);
    const uint32_t yFirst = __ldg(&tp.svo_nodes[idx].y);
    if (!two) {
        svo_unpack(yFirst, lo);
        return;
    }
    svo_unpack(yFirst, up);
    uint32_t yLo = yFirst;
    // the finer descent continues while a child block exists and the resolution is above its target
    while (cur != 0u && resolution > targetLo) {
        resolution >>= 1;
        const uint32_t cx = fp[0] >= resolution, cy = fp[1] >= resolution, cz = fp[2] >= resolution;
        idx = cur + (cz | (cx << 1) | (cy << 2));
        cur = __ldg(&tp.svo_nodes[idx].x) & 0x7fffffffu;
        fp[0] -= cx * resolution; fp[1] -= cy * resolution; fp[2] -= cz * resolution;
        yLo = __ldg(&tp.svo_nodes[idx].y);
    }
    svo_unpack(yLo, lo);
}

DEVFN void svo_trace_cone(const TraceParams& tp, const float* startPos_, const float* dir, float coneCoefficient,
                          float startLevel, float stepFactor, float* out)
{
    const vgi_vct_params& p = tp.p;
    float result[4] = { 0.f, 0.f, 0.f, 0.f };
    float voxelSize = p.voxel_size * exp2f(startLevel);
    float startPos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) startPos[k] = startPos_[k] + dir[k] * voxelSize * p.trace_start_offset * 0.5f;
    float step = 0.0f;
    float diameter = fmaxf(step * coneCoefficient, p.voxel_size);
    float occlusion = 0.0f;
    float curSegmentLength = voxelSize;
    const float invVoxel = 1.0f / p.voxel_size;
    while (step < MAX_TRACE_DISTANCE && occlusion < 1.0f) {
        float position[3], d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            position[k] = startPos[k] + dir[k] * step;
            d[k] = p.volume_center[k] - position[k];
        }
        // level selection as in the clipmap tracer: exact threshold count for the distance term, fast log2 for
        // the (continuously blended) diameter term
        const float minLevel = min_level_from_dd(tp, dot3(d, d));
        const float curLevel = fminf(fmaxf(fmaxf(startLevel, __log2f(diameter * invVoxel)), minLevel), tp.svo_max_level);
        float lo[4], smp[4];
        const float fl = floorf(curLevel);
        const float fr = curLevel - fl;
        float up[4];
        sample_svo(tp, position, (uint32_t)fl, fr > 0.0f, lo, up);
        if (fr > 0.0f) {
#pragma unroll
            for (int c = 0; c < 4; ++c) smp[c] = lo[c] * (1.0f - fr) + up[c] * fr;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) smp[c] = lo[c];
        }
        voxelSize = p.voxel_size * exp2f(curLevel);
        if (smp[0] != 0.0f || smp[1] != 0.0f || smp[2] != 0.0f || smp[3] != 0.0f) { // empty nodes add exact zeros
            const float correction = __fdividef(curSegmentLength, voxelSize);
            float opacity = 0.0f;
            if (smp[3] > 0.0f) opacity = f_clamp(1.0f - exp2f(correction * __log2f(1.0f - smp[3])), 0.0f, 1.0f);
            const float k1 = f_clamp(1.0f - result[3], 0.0f, 1.0f);
#pragma unroll
            for (int c = 0; c < 3; ++c) result[c] += k1 * (smp[c] * correction);
            result[3] += k1 * opacity;
            occlusion += __fdividef((1.0f - occlusion) * opacity, 1.0f + (step + voxelSize) * p.occlusion_decay);
        }
        const float prevStep = step;
        step += fmaxf(diameter, p.voxel_size) * stepFactor;
        curSegmentLength = step - prevStep;
        diameter = step * coneCoefficient;
    }
    out[0] = f_clamp(result[0], 0.0f, 1.0f);
    out[1] = f_clamp(result[1], 0.0f, 1.0f);
    out[2] = f_clamp(result[2], 0.0f, 1.0f);
    out[3] = f_clamp(1.0f - occlusion, 0.0f, 1.0f);
}

// per-pixel combine of the octree tracer (voxelConeTracing_Octree.frag:196-311)
template <int NCONES, int TILE_PIX>
DEVFN void svo_finish_pixel(const TraceParams& tp, const PixelSetup& s, size_t pi, const float (*cones)[3],
                            const float4 (*s_res)[TILE_PIX], int tid, bool needCones, bool needDirect, bool needSpec)
{
    const uint32_t mode = tp.p.rendering_mode;
    float indirect[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
    if (needCones) {
        float validConeCount = 0.0f;
        for (int i = 0; i < NCONES; ++i) {
            const float cosTheta = s.normal[0] * cones[i][0] + s.normal[1] * cones[i][1] + s.normal[2] * cones[i][2];
            if (cosTheta < 0.0f) continue;
            const float4 r = s_res[i][tid];
            indirect[0] += r.x; indirect[1] += r.y; indirect[2] += r.z; indirect[3] += r.w;
            validConeCount += cosTheta;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) indirect[k] /= validConeCount; // Q11: normalised by the sum of cosines
        indirect[3] *= tp.p.ambient_occlusion_factor;
#pragma unroll
        for (int k = 0; k < 3; ++k) indirect[k] *= s.diffuseColor[k] * tp.p.indirect_diffuse_intensity;
#pragma unroll
        for (int k = 0; k < 4; ++k) indirect[k] = f_clamp(indirect[k], 0.0f, 1.0f);
    }

    float spec[3] = { 0.f, 0.f, 0.f };
    if (needSpec && (s.specularColor[0] > 1e-6f || s.specularColor[1] > 1e-6f || s.specularColor[2] > 1e-6f) && s.metallic > 1e-6f) {
        const float I[3] = { -s.view[0], -s.view[1], -s.view[2] };
        const float dn = dot3(s.normal, I);
        const float r[3] = { I[0] - 2.0f * dn * s.normal[0], I[1] - 2.0f * dn * s.normal[1], I[2] - 2.0f * dn * s.normal[2] };
        float sdir[3];
        normalize3(r, sdir);
        const float aperture = fmaxf(s.perceptualRoughness, MIN_SPECULAR_APERTURE);
        float c[4];
        // ref: voxelConeTracing_Octree.frag:215-219 — stepFactor = max(0.2, uVoxelSize) (Q12)
        svo_trace_cone(tp, s.startPos, sdir, 2.0f * tanf(aperture * 0.5f), s.minLevel, fmaxf(MIN_TRACE_STEP_FACTOR, tp.p.voxel_size), c);
#pragma unroll
        for (int k = 0; k < 3; ++k) spec[k] = (c[k] * s.specularColor[k]) * tp.p.indirect_specular_intensity;
    }

    float direct[3] = { 0.f, 0.f, 0.f };
    if (needDirect) {
        if (s.emission[0] > 0.0f || s.emission[1] > 0.0f || s.emission[2] > 0.0f) {
            direct[0] = s.emission[0]; direct[1] = s.emission[1]; direct[2] = s.emission[2];
        } else {
            const float alphaRoughness = s.perceptualRoughness * s.perceptualRoughness;
            const float reflectance = fmaxf(fmaxf(s.specularColor[0], s.specularColor[1]), s.specularColor[2]);
            const float r90 = f_clamp(reflectance * 50.0f, 0.0f, 1.0f);
            const float* L = tp.light.dir_to_light;
            float h[3];
            {
                const float t[3] = { L[0] + s.view[0], L[1] + s.view[1], L[2] + s.view[2] };
                normalize3(t, h);
            }
            const float NdotL = f_clamp(dot3(s.normal, L), 0.001f, 1.0f);
            const float NdotV = f_clamp(fabsf(dot3(s.normal, s.view)), 0.001f, 1.0f);
            const float NdotH = f_clamp(dot3(s.normal, h), 0.0f, 1.0f);
            const float VdotH = f_clamp(dot3(s.view, h), 0.0f, 1.0f);
            float brdf[3];
            microfacet_brdf(NdotL, NdotV, NdotH, VdotH, alphaRoughness, s.specularColor, r90, s.diffuseColor, brdf);
            const float vis = calc_visibility(tp, s.worldPos);
            direct[0] = brdf[0] * vis; direct[1] = brdf[1] * vis; direct[2] = brdf[2] * vis;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) direct[k] = f_clamp(direct[k], 0.0f, 1.0f);
    }

    float4 dc = make_float4(0.f, 0.f, 0.f, 1.f), sc = make_float4(0.f, 0.f, 0.f, 1.f);
    switch (mode) {
    case 0: dc = make_float4(s.diffuseColor[0], s.diffuseColor[1], s.diffuseColor[2], 1.f); break;
    case 1: dc = make_float4(s.specularColor[0], s.specularColor[1], s.specularColor[2], 1.f); break;
    case 2: dc = make_float4(s.normal[0] * 0.5f + 0.5f, s.normal[1] * 0.5f + 0.5f, s.normal[2] * 0.5f + 0.5f, 1.f); break;
    case 3: { // minLevelToColor, voxelConeTracing_Octree.frag:433-458
        float col[4] = { 0.f, 0.f, 0.f, 0.f };
        if (s.minLevel < 6.0f) {
            int lower = (int)floorf(s.minLevel);
            lower = lower < 0 ? 0 : lower;
            const float fr = f_fract(s.minLevel);
            for (int k = 0; k < 4; ++k) col[k] = c_level_colors[lower][k] * (1.0f - fr) + c_level_colors[lower + 1][k] * fr;
        }
        dc = make_float4(col[0] * 0.5f, col[1] * 0.5f, col[2] * 0.5f, col[3] * 0.5f);
        break;
    }
    case 4: dc = make_float4(direct[0] * indirect[3], direct[1] * indirect[3], direct[2] * indirect[3], 1.f); break;
    case 5: dc = make_float4(indirect[0], indirect[1], indirect[2], 1.f); break; // direct not added (_Octree.frag:283-287)
    case 6: sc = make_float4(spec[0], spec[1], spec[2], 1.f); break;
    case 7: dc = make_float4(indirect[3], indirect[3], indirect[3], 1.f); break;
    case 8:
        dc = make_float4(direct[0] * indirect[3] + indirect[0], direct[1] * indirect[3] + indirect[1],
                         direct[2] * indirect[3] + indirect[2], 1.f);
        sc = make_float4(spec[0], spec[1], spec[2], 1.f);
        break;
    default: break;
    }
    tp.out_diffuse[pi] = dc;
    tp.out_specular[pi] = sc;
}

// main pass: diffuse cones + direct term + mode switch; pixels that need a specular cone are
// appended to a compact list and finished by k_trace_specular (the specular march is up to two
// orders of magnitude longer than a diffuse cone, Q12, and would stall whole warps).
//
// One block = one 8x8 pixel tile. Phase 1: one thread per pixel reconstructs the surface point.
// Phase 2: the (cone, pixel) pairs that pass the hemisphere test (voxelConeTracing.frag:175-184) are
// compacted cone-major, so every lane of a warp marches a real cone and neighbouring lanes march the
// same direction from neighbouring pixels. Phase 3: one thread per pixel adds its cones in the
// shader's order and finishes the pixel.
// One block = one TILE_W x 8 pixel tile. The clipmap march runs 8 x 8 tiles (tighter footprints per warp beat the
// better round filling of 16 x 8: 3.37 vs 3.67 ms); the latency-bound octree march runs 16 x 8 for the 16-cone set
// (every thread owns a pixel in phases 1 and 3; 3.05 -> 2.86 ms).
#define TILE_H 8

// Resident blocks per SM the register allocation aims at. Measured on B200 (main / specular, us per 1080p frame):
// unconstrained (96 / 86 registers, 5 blocks) 3086 / 1395; 6 blocks (80 / 85 registers, no spills) 2842 / 1372;
// 7 blocks (72 registers, spills) 2922 for the main kernel.
// With the footprint-only probe the clipmap march fits 72 registers with 12 bytes of spill: 7 blocks 2.58 ms against 2.66 ms
// at 6 (8 blocks / 64 registers: 2.59). The octree march keeps 6 (its tiles need 40 KB of shared memory: 5 blocks fit anyway).
#ifndef VGI_TRACE_MAIN_MINBLOCKS
#define VGI_TRACE_MAIN_MINBLOCKS 7
#endif
#ifndef VGI_TRACE_SPEC_MINBLOCKS
#define VGI_TRACE_SPEC_MINBLOCKS 6
#endif
// SVO = true: the same tile / compaction machinery marching the octree (voxelConeTracing_Octree.frag); the
// per-pixel combine follows that shader (normalisation by the sum of cosines, clamps, specular cone inline).
template <int NCONES, bool SVO, int TILE_W>
__global__ void __launch_bounds__(128, SVO ? 6 : VGI_TRACE_MAIN_MINBLOCKS) k_trace_main(const __grid_constant__ TraceParams tp)
{
    constexpr int TILE_PIX = TILE_W * TILE_H;
    __shared__ float s_pix[8][TILE_PIX];            // startPos xyz, normal xyz, minLevel, valid
    __shared__ float4 s_res[NCONES][TILE_PIX];      // cone result * cos(theta)
    __shared__ uint16_t s_list[NCONES * TILE_PIX];
    __shared__ int s_warp_count[4];
    __shared__ StepTable s_table;
#if VGI_TRACE_V2 == 1
    __shared__ float4 s_coop[4][64];                // per warp, double-buffered: 4 groups x 8 pre-blended corner records
    __shared__ float4 s_face[2 * NCONES];           // per cone: face-texel word offsets, direction weights (fetch lanes)
#elif VGI_TRACE_V2 == 3
    __shared__ float4 s_coop[4][16];
    __shared__ float4 s_face[2 * NCONES];
#else
    __shared__ float4 s_coop[4][16];                // per warp: the eight pre-blended corner records of a shared cell
#endif

    const int tid = threadIdx.x;
    const int tilesX = (tp.width + TILE_W - 1) / TILE_W;
    // tile rows are dealt round-robin when the image is sharded across GPUs (tile_stride > 1)
    const int tx0 = (blockIdx.x % tilesX) * TILE_W, ty0 = tp.y0 + (tp.tile_phase + (int)(blockIdx.x / tilesX) * tp.tile_stride) * TILE_H;
    const uint32_t mode = tp.p.rendering_mode;
    const bool needCones = mode == 4 || mode == 5 || mode == 7 || mode == 8;
    const bool needDirect = SVO ? (mode == 4 || mode == 8) : (mode == 4 || mode == 5 || mode == 8);
    const bool needSpec = mode == 6 || mode == 8;
    const float (*cones)[3] = NCONES == 32 ? c_cones32 : c_cones16;

    // ---- phase 1 (the full PixelSetup is rebuilt in phase 3 rather than kept live across the march)
    const int px = tx0 + (tid & (TILE_W - 1)), py = ty0 + ((tid / TILE_W) & (TILE_H - 1));
    if (tid < TILE_PIX) {
        PixelSetup s;
        bool valid = false;
        if (px < tp.width && py < tp.y1) valid = pixel_setup(tp, px, py, s);
        s_pix[0][tid] = s.startPos[0]; s_pix[1][tid] = s.startPos[1]; s_pix[2][tid] = s.startPos[2];
        s_pix[3][tid] = s.normal[0]; s_pix[4][tid] = s.normal[1]; s_pix[5][tid] = s.normal[2];
        s_pix[6][tid] = s.minLevel;
        s_pix[7][tid] = valid ? 1.0f : 0.0f;
    }
    if (tid == 127 && needCones && !SVO)
        build_step_table(tp, s_table, tp.cone_coeff_diffuse, fmaxf(MIN_TRACE_STEP_FACTOR, tp.p.min_trace_step_factor));
#if VGI_TRACE_V2 == 1 || VGI_TRACE_V2 == 3
    if (!SVO && needCones && tid >= 64 && tid < 64 + NCONES) {
        const float fdir[3] = { cones[tid - 64][0], cones[tid - 64][1], cones[tid - 64][2] };
        cone_face_table(fdir, s_face + 2 * (tid - 64));
    }
#endif
    __syncthreads();

    if (needCones) {
        // ---- phase 2a: cone-major compaction; warp w owns slots [w * S/4, (w+1) * S/4)
        constexpr int SLOTS = NCONES * TILE_PIX, PER_WARP = SLOTS / 4;
        const int warp = tid >> 5, lane = tid & 31;
        uint32_t actBits = 0u;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < PER_WARP / 32; ++k) {
            const int slot = warp * PER_WARP + k * 32 + lane;
            const int cone = slot / TILE_PIX, pix = slot % TILE_PIX;
            const float c = s_pix[3][pix] * cones[cone][0] + s_pix[4][pix] * cones[cone][1] + s_pix[5][pix] * cones[cone][2];
            // one predicate for both passes (a NaN cosine - zero-length G-buffer normal - traces like the shader: NaN colours)
            const bool act = s_pix[7][pix] != 0.0f && !(c < 0.0f);
            actBits |= act ? (1u << k) : 0u;
            cnt += __popc(__ballot_sync(0xffffffffu, act));
        }
        if (lane == 0) s_warp_count[warp] = cnt;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (w < warp) base += s_warp_count[w];
            total += s_warp_count[w];
        }
#pragma unroll
        for (int k = 0; k < PER_WARP / 32; ++k) {
            const int slot = warp * PER_WARP + k * 32 + lane;
            const bool act = (actBits >> k) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, act);
            if (act) s_list[base + __popc(b & ((1u << lane) - 1u))] = (uint16_t)slot;
            else s_res[slot / TILE_PIX][slot % TILE_PIX] = make_float4(0.f, 0.f, 0.f, 0.f);
            base += __popc(b);
        }
        __syncthreads();
        // ---- phase 2b: march
        if (!SVO && VGI_TRACE_COOP && s_table.n >= 0) {
            for (int base = 0; base < total; base += 128) {
                const int i = base + tid;
                const bool have = i < total;
                const int slot = have ? s_list[i] : 0;
                const int cone = slot / TILE_PIX, pix = slot % TILE_PIX;
                const float dir[3] = { cones[cone][0], cones[cone][1], cones[cone][2] };
                const float sp[3] = { s_pix[0][pix], s_pix[1][pix], s_pix[2][pix] };
                const float cosTheta = s_pix[3][pix] * dir[0] + s_pix[4][pix] * dir[1] + s_pix[5][pix] * dir[2];
                float c[4];
#if VGI_TRACE_V2 == 1 || VGI_TRACE_V2 == 3
                // the warp's slice of the cone-major list: first and last cone; more than two -> per-lane path
                const unsigned haveMask = __ballot_sync(FULL_MASK, have);
                const int c0 = __shfl_sync(FULL_MASK, cone, 0), c1 = __shfl_sync(FULL_MASK, cone, 31 - __clz(haveMask | 1u));
                const bool multi = __any_sync(FULL_MASK, have && cone != c0 && cone != c1);
#if VGI_TRACE_V2 == 1
                march_warp_v2(tp, s_table, have, cone, c0, c1, multi, sp, dir, s_pix[6][pix], s_face, s_coop[warp], (unsigned)lane, c);
#else
                march_warp_table(tp, s_table, have, cone, cones, sp, dir, s_pix[6][pix], s_coop[warp], (unsigned)lane, c, c0, c1, multi, s_face);
#endif
#else
                march_warp_table(tp, s_table, have, cone, cones, sp, dir, s_pix[6][pix], s_coop[warp], (unsigned)lane, c);
#endif
                if (have) s_res[cone][pix] = make_float4(c[0] * cosTheta, c[1] * cosTheta, c[2] * cosTheta, c[3] * cosTheta);
            }
        } else
        for (int i = tid; i < total; i += 128) {
            const int slot = s_list[i];
            const int cone = slot / TILE_PIX, pix = slot % TILE_PIX;
            const float dir[3] = { cones[cone][0], cones[cone][1], cones[cone][2] };
            const float sp[3] = { s_pix[0][pix], s_pix[1][pix], s_pix[2][pix] };
            const float cosTheta = s_pix[3][pix] * dir[0] + s_pix[4][pix] * dir[1] + s_pix[5][pix] * dir[2];
            float c[4];
            if (SVO) svo_trace_cone(tp, sp, dir, tp.cone_coeff_diffuse, s_pix[6][pix], fmaxf(MIN_TRACE_STEP_FACTOR, tp.p.min_trace_step_factor), c);
            else if (!VGI_TRACE_COOP && s_table.n >= 0) trace_cone_table(tp, s_table, sp, dir, s_pix[6][pix], c);
            else trace_cone(tp, sp, dir, tp.cone_coeff_diffuse, MAX_TRACE_DISTANCE, s_pix[6][pix],
                            fmaxf(MIN_TRACE_STEP_FACTOR, tp.p.min_trace_step_factor), c);
            s_res[cone][pix] = make_float4(c[0] * cosTheta, c[1] * cosTheta, c[2] * cosTheta, c[3] * cosTheta);
        }
        __syncthreads();
    }

    // ---- phase 3
    if (tid >= TILE_PIX || s_pix[7][tid] == 0.0f) return;
    PixelSetup s;
    pixel_setup(tp, px, py, s);
    const size_t pi = (size_t)py * tp.width + px;
    float indirect[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
    if (SVO) {
        svo_finish_pixel<NCONES, TILE_PIX>(tp, s, pi, cones, s_res, tid, needCones, needDirect, needSpec);
        return;
    }
    if (needCones) {
#pragma unroll 4
        for (int i = 0; i < NCONES; ++i) {
            const float4 r = s_res[i][tid];
            indirect[0] += r.x; indirect[1] += r.y; indirect[2] += r.z; indirect[3] += r.w;
        }
        const float invN = 1.0f / (float)NCONES;
#pragma unroll
        for (int k = 0; k < 4; ++k) indirect[k] *= invN;
        indirect[3] *= tp.p.ambient_occlusion_factor;
#pragma unroll
        for (int k = 0; k < 3; ++k) indirect[k] *= s.diffuseColor[k] * tp.p.indirect_diffuse_intensity;
    }

    const bool wantsSpec = needSpec && (s.specularColor[0] > 1e-6f || s.specularColor[1] > 1e-6f || s.specularColor[2] > 1e-6f) && s.metallic > 1e-6f;
    if (wantsSpec && !tp.spec_presplit) {
        const uint32_t slot = atomicAdd(tp.spec_count, 1u);
        tp.spec_list[slot] = (uint32_t)pi;
    }

    float direct[3] = { 0.f, 0.f, 0.f };
    if (needDirect) {
        if (s.emission[0] > 0.0f || s.emission[1] > 0.0f || s.emission[2] > 0.0f) {
            direct[0] = s.emission[0]; direct[1] = s.emission[1]; direct[2] = s.emission[2];
        } else {
            const float alphaRoughness = s.perceptualRoughness * s.perceptualRoughness;
            const float reflectance = fmaxf(fmaxf(s.specularColor[0], s.specularColor[1]), s.specularColor[2]);
            const float r90 = f_clamp(reflectance * 50.0f, 0.0f, 1.0f);
            const float* L = tp.light.dir_to_light;
            float h[3];
            {
                const float t[3] = { L[0] + s.view[0], L[1] + s.view[1], L[2] + s.view[2] };
                normalize3(t, h);
            }
            const float NdotL = f_clamp(dot3(s.normal, L), 0.001f, 1.0f);
            const float NdotV = f_clamp(fabsf(dot3(s.normal, s.view)), 0.001f, 1.0f);
            const float NdotH = f_clamp(dot3(s.normal, h), 0.0f, 1.0f);
            const float VdotH = f_clamp(dot3(s.view, h), 0.0f, 1.0f);
            float brdf[3];
            microfacet_brdf(NdotL, NdotV, NdotH, VdotH, alphaRoughness, s.specularColor, r90, s.diffuseColor, brdf);
            const float vis = calc_visibility(tp, s.worldPos);
            direct[0] = brdf[0] * vis; direct[1] = brdf[1] * vis; direct[2] = brdf[2] * vis;
        }
    }

    float4 dc = make_float4(0.f, 0.f, 0.f, 1.f);
    switch (mode) {
    case 0: dc = make_float4(s.diffuseColor[0], s.diffuseColor[1], s.diffuseColor[2], 1.f); break;
    case 1: dc = make_float4(s.specularColor[0], s.specularColor[1], s.specularColor[2], 1.f); break;
    case 2: dc = make_float4(s.normal[0] * 0.5f + 0.5f, s.normal[1] * 0.5f + 0.5f, s.normal[2] * 0.5f + 0.5f, 1.f); break;
    case 3: {
        int lower = (int)floorf(s.minLevel);
        lower = lower < 0 ? 0 : (lower > 5 ? 5 : lower);
        const float fr = f_fract(s.minLevel);
        float o[4];
        for (int k = 0; k < 4; ++k) o[k] = (c_level_colors[lower][k] * (1.0f - fr) + c_level_colors[lower + 1][k] * fr) * 0.5f;
        dc = make_float4(o[0], o[1], o[2], o[3]);
        break;
    }
    case 4: dc = make_float4(direct[0] * indirect[3], direct[1] * indirect[3], direct[2] * indirect[3], 1.f); break;
    case 5:
    case 8:
        dc = make_float4(direct[0] * indirect[3] + indirect[0], direct[1] * indirect[3] + indirect[1],
                         direct[2] * indirect[3] + indirect[2], 1.f);
        break;
    case 7: dc = make_float4(indirect[3], indirect[3], indirect[3], 1.f); break;
    default: break;
    }
    tp.out_diffuse[pi] = dc;
    // a listed pixel's specular texel belongs to the specular march, which may already have written it
    if (!(wantsSpec && tp.spec_presplit)) tp.out_specular[pi] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// The pixels of this call's tiles that need a specular cone (voxelConeTracing.frag:146-149, 208), listed ahead of both
// marches so that they can run side by side. Same tile walk as k_trace_main (four 8 x 8 tiles per block, a warp = half
// a tile) and the same predicate as its phase 3; warp-aggregated append, so the list stays in tile order.
__global__ void __launch_bounds__(256) k_spec_classify(const __grid_constant__ TraceParams tp)
{
    const int tilesX4 = (tp.width + 31) / 32;
    const int tid = threadIdx.x;
    const int tx0 = ((int)(blockIdx.x % tilesX4) * 4 + (tid >> 6)) * 8;
    const int ty0 = tp.y0 + (tp.tile_phase + (int)(blockIdx.x / tilesX4) * tp.tile_stride) * TILE_H;
    const int px = tx0 + (tid & 7), py = ty0 + ((tid >> 3) & 7);
    bool want = false;
    size_t pi = 0;
    if (px < tp.width && py < tp.y1) {
        pi = (size_t)py * tp.width + px;
        if (__ldg(tp.depth + pi) != 1.0f) {
            const uchar4 spc = __ldg(reinterpret_cast<const uchar4*>(tp.specular) + pi);
            want = (spc.x / 255.0f > 1e-6f || spc.y / 255.0f > 1e-6f || spc.z / 255.0f > 1e-6f) && spc.w / 255.0f > 1e-6f;
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, want);
    if (!b) return;
    const unsigned lane = tid & 31u;
    uint32_t base = 0u;
    if (lane == 0u) base = atomicAdd(tp.spec_count, (uint32_t)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want) tp.spec_list[base + __popc(b & ((1u << lane) - 1u))] = (uint32_t)pi;
}

// ref: voxelConeTracing.frag:205-216 (stepFactor = uVoxelSize, Q12)
// Specular marches differ by two orders of magnitude in length (aperture from the roughness, early exit on
// full occlusion), so pixels are not bound to lanes: a lane that finishes its cone immediately claims the
// next pixel of the compacted list from a global cursor (warp-aggregated atomic), and every lane of a warp
// executes one marching step per loop iteration.
__global__ void __launch_bounds__(128, VGI_TRACE_SPEC_MINBLOCKS) k_trace_specular(const __grid_constant__ TraceParams tp)
{
    const uint32_t n = *tp.spec_count;
    const unsigned lane = threadIdx.x & 31u;
    const vgi_vct_params& p = tp.p;
    const float invVoxel = 1.0f / p.voxel_size;
    const float stepFactor = p.voxel_size;
    bool active = false, exhausted = false;
    // per-lane cone state
    ConeState cs = { { 0.f, 0.f, 0.f, 0.f }, 0.0f };
    ConeFaces cf = { false, false, false, 0.f, 0.f, 0.f };
    float startPos[3] = { 0.f, 0.f, 0.f }, dir[3] = { 0.f, 0.f, 1.f }, spec[3] = { 0.f, 0.f, 0.f };
    float step = 0.0f, diameter = 0.0f, seg = 0.0f, startLevel = 0.0f, coneCoefficient = 0.0f;
    uint32_t pi = 0u;
    for (;;) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(0xffffffffu, !active && !exhausted);
        if (idle) {
            uint32_t base = 0u;
            if (lane == (unsigned)(__ffs(idle) - 1)) base = atomicAdd(tp.spec_cursor, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, __ffs(idle) - 1);
            if (!active && !exhausted) {
                const uint32_t item = base + __popc(idle & ((1u << lane) - 1u));
                if (item >= n) {
                    exhausted = true;
                } else {
                    pi = tp.spec_list[item];
                    const int px = (int)(pi % (uint32_t)tp.width), py = (int)(pi / (uint32_t)tp.width);
                    PixelSetup s;
                    if (pixel_setup(tp, px, py, s)) {
                        // reflect(-view, normal) = I - 2 dot(N, I) N
                        const float I[3] = { -s.view[0], -s.view[1], -s.view[2] };
                        const float dn = dot3(s.normal, I);
#pragma unroll
                        for (int k = 0; k < 3; ++k) dir[k] = I[k] - 2.0f * dn * s.normal[k];
                        const float aperture = fmaxf(s.perceptualRoughness, MIN_SPECULAR_APERTURE);
                        coneCoefficient = 2.0f * tanf(aperture * 0.5f);
                        startLevel = s.minLevel;
                        const float voxelSize0 = p.voxel_size * exp2f(startLevel);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            startPos[k] = s.startPos[k] + dir[k] * voxelSize0 * p.trace_start_offset * 0.5f;
                            spec[k] = s.specularColor[k] * p.indirect_specular_intensity;
                        }
                        cf = cone_faces(dir);
                        cs.result[0] = cs.result[1] = cs.result[2] = cs.result[3] = 0.0f;
                        cs.occlusion = 0.0f;
                        step = 0.0f;
                        diameter = fmaxf(step * coneCoefficient, p.voxel_size);
                        seg = voxelSize0;
                        active = true;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        // ---- one marching step per active lane
        if (active) {
            cone_step(tp, cs, startPos, dir, cf, startLevel, step, __log2f(diameter * invVoxel), seg);
            const float prevStep = step;
            step += fmaxf(diameter, p.voxel_size) * stepFactor;
            seg = step - prevStep;
            diameter = step * coneCoefficient;
            if (!(step < MAX_TRACE_DISTANCE && cs.occlusion < 1.0f)) {
                tp.out_specular[pi] = make_float4(cs.result[0] * spec[0], cs.result[1] * spec[1], cs.result[2] * spec[2], 1.0f);
                active = false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Specular cones, one WARP per cone (ref: voxelConeTracing.frag:205-216, 341-392; Q12: stepFactor = uVoxelSize, so a
// cone advances 1/16 of a voxel of its level per step and ~8 consecutive samples fall into the same cell).
// The step sequence does not depend on what is sampled and takes one of 256 forms (8-bit roughness): the host
// tabulates it (TraceParams::spec_tab) and the 32 lanes of a warp evaluate 32 CONSECUTIVE steps of one cone:
//   * every lane computes where its two level samples lie (cell key + tri-linear weights);
//   * consecutive lanes with equal keys form runs (about four per level and batch); lane 8 r + c tests the
//     footprint byte of run r's cell and fetches corner record c, blended with the cone's face weights,
//     into shared memory — once per cell instead of once per step; a batch whose cells are all empty ends here;
//   * every lane sums the non-zero corners of its cell with its own weights;
//   * front-to-back accumulation is a prefix product over the batch: 1 - alpha' = (1 - alpha)(1 - o) and
//     1 - occ' = (1 - occ)(1 - o / (1 + (step + voxelSize) decay)) are the shader's updates :380-388 in product form.
// The march ends like the shader's when the occlusion saturates (binary32: 1 - occ below 2^-25), and also once the
// transmittance 1 - alpha is below 1e-10: only rgb leaves this pass, and every later contribution is scaled by it.
// ---------------------------------------------------------------------------------------------------
#ifndef VGI_TRACE_SPEC_WARP
#define VGI_TRACE_SPEC_WARP 1
#endif
#ifndef VGI_TRACE_SPECW_MINBLOCKS
#define VGI_TRACE_SPECW_MINBLOCKS 7      // measured (first warp marcher): 8 (64 registers, spills) 1.70 ms, 6 (80) 1.59 ms, 4 1.69 ms; current kernel: 6 / 7 / 8 = 1.203 / 1.167 / 1.164 ms (72 registers, no spills at 7)
#endif

struct SpecWarpShared {
    float4 val[32];     // 4 cells x 8 pre-blended corner records
    uint2  cell[64];    // non-empty cells of the current batch: (key, mask of the records that can be non-zero)
};

#ifndef VGI_TRACE_SPEC_MERGED
#define VGI_TRACE_SPEC_MERGED 1
#endif

// One level sample of the batch for every lane that wants one; false when every cell is empty.
//  A. run heads (first lane of each group of consecutive lanes with the same cell) test their cell's
//     footprint byte: one probe per CELL, not per step; a pass whose cells are all empty ends with one ballot;
//  B. the non-empty cells are numbered compactly; lane 8 c + corner fetches corner `corner` of compact cell c (its three
//     face texels blended with the cone's weights) into shared memory, four cells per round;
//  C. every lane sums the non-zero corners of its cell with its own tri-linear weights.
DEVFN bool spec_level_pass(const TraceParams& tp, bool want, uint32_t key, const float* w, unsigned lane,
                           uint32_t ox, uint32_t oy, uint32_t oz, float kx, float ky, float kz,
                           SpecWarpShared& sh, float* out)
{
    const unsigned wantMask = __ballot_sync(FULL_MASK, want);
    if (!wantMask) return false;
    const uint32_t k2 = want ? key : 0xffffffffu;
    const uint32_t prev = __shfl_up_sync(FULL_MASK, k2, 1);
    const bool head = want && (lane == 0u || k2 != prev);
    const unsigned heads = __ballot_sync(FULL_MASK, head);
    const int R = tp.R, Rm = R - 1, logR = tp.logR;
    // ---- A
    uint32_t cellMask = 0u;
    if (head) {
        const uint32_t ix = key & (uint32_t)Rm, iy = (key >> logR) & (uint32_t)Rm, iz = (key >> (2 * logR)) & (uint32_t)Rm;
        const uint32_t level = key >> (3 * logR);
        const uint32_t nbShift = (uint32_t)logR - 2u, wprShift = (uint32_t)logR - 5u;
        const uint32_t bidx = ((((level << nbShift) + (iz >> 2)) << nbShift) + (iy >> 2) << wprShift) + (ix >> 5);
        const uint32_t bbyte = __ldg(tp.brick_mask + bidx);
        const uint32_t m = __ldg(tp.footprint + key);   // meaningful only where the brick bit is set
        STAT(1, 1);
        cellMask = ((bbyte >> ((ix >> 2) & 7u)) & 1u) ? m : 0u;
    }
    const unsigned live = __ballot_sync(FULL_MASK, cellMask != 0u);
    if (!live) return false;
    // ---- B: my run's head lane, its mask and its compact number
    const int hl = (31 - __clz((int)(heads & (0xffffffffu >> (31u - lane))))) & 31;
    const uint32_t headMask = __shfl_sync(FULL_MASK, cellMask, hl);
    const uint32_t myMask = want ? headMask : 0u;
    const int myCell = __popc(live & ((1u << hl) - 1u));
    const int nCells = __popc(live);
    if (cellMask) sh.cell[myCell] = make_uint2(key, cellMask);     // head lanes of non-empty cells (hl == lane)
    __syncwarp();
    const unsigned corner = lane & 7u;
    for (int c0 = 0; c0 < nCells; c0 += 4) {
        const int c = c0 + (int)(lane >> 3);
        if (c < nCells) {
            const uint2 cm = sh.cell[c];
#if VGI_TRACE_SPEC_GATHER_ALL
            if (!((cm.y >> corner) & 1u)) sh.val[lane] = make_float4(0.f, 0.f, 0.f, 0.f);   // every parked slot is gathered
#endif
            if ((cm.y >> corner) & 1u) {
                const uint32_t vox = cm.x;
                const uint32_t ix = vox & (uint32_t)Rm, iy = (vox >> logR) & (uint32_t)Rm, iz = (vox >> (2 * logR)) & (uint32_t)Rm;
                int off = 0;
                if (corner & 1u) off += (ix == (uint32_t)Rm) ? -Rm : 1;
                if (corner & 2u) off += ((iy == (uint32_t)Rm) ? -Rm : 1) << logR;
                if (corner & 4u) off += ((iz == (uint32_t)Rm) ? -Rm : 1) << (2 * logR);
                const uint32_t* rec = reinterpret_cast<const uint32_t*>(tp.store + rec_index(vox + (uint32_t)off, logR));
                const uint32_t tx = __ldg(rec + ox), ty = __ldg(rec + oy), tz = __ldg(rec + oz);
                STAT(4, 1);
                const float2 kx2 = make_float2(kx, kx), ky2 = make_float2(ky, ky), kz2 = make_float2(kz, kz);
                float2 lo = __fmul2_rn(kx2, unpack2(tx, 0x7540u, 0x7541u)), hi = __fmul2_rn(kx2, unpack2(tx, 0x7542u, 0x7543u));
                lo = __ffma2_rn(ky2, unpack2(ty, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(ky2, unpack2(ty, 0x7542u, 0x7543u), hi);
                lo = __ffma2_rn(kz2, unpack2(tz, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(kz2, unpack2(tz, 0x7542u, 0x7543u), hi);
                sh.val[lane] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        }
        __syncwarp();
        // ---- C
        const int q = myCell - c0;
        if (myMask && q >= 0 && q < 4) coop_gather_w<VGI_TRACE_SPEC_GATHER_ALL != 0>(w, myMask, sh.val + 8 * q, out);
        __syncwarp();
    }
    return myMask != 0u;
}

// footprint byte of a cell (0 = every record of its footprint is zero)
DEVFN uint32_t spec_probe_cell(const TraceParams& tp, uint32_t key)
{
    STAT(1, 1);
    return __ldg(tp.footprint + key);   // valid for every voxel (k_brick_mask)
}

// Both level samples of the batch in ONE pass (spec_level_pass twice, merged): the run heads of both levels probe their
// cells together (four mask loads in flight instead of two dependent pairs), the non-empty cells of both levels share one
// compact numbering (lo cells first) and therefore the fetch rounds: typically one round of four cells instead of one per
// level. Returns per lane whether its lo / hi sample is non-zero.
DEVFN void spec_both_levels(const TraceParams& tp, bool wantLo, uint32_t keyLo, const float* wLo, bool wantHi, uint32_t keyHi,
                            const float* wHi, unsigned lane, uint32_t ox, uint32_t oy, uint32_t oz, float kx, float ky, float kz,
                            SpecWarpShared& sh, float* smpLo, float* smpHi, bool& anyLo, bool& anyHi)
{
    anyLo = anyHi = false;
    if (!__ballot_sync(FULL_MASK, wantLo || wantHi)) return;
    const uint32_t kL = wantLo ? keyLo : 0xffffffffu, kH = wantHi ? keyHi : 0xffffffffu;
    const uint32_t pL = __shfl_up_sync(FULL_MASK, kL, 1), pH = __shfl_up_sync(FULL_MASK, kH, 1);
    const bool headL = wantLo && (lane == 0u || kL != pL), headH = wantHi && (lane == 0u || kH != pH);
    const unsigned headsL = __ballot_sync(FULL_MASK, headL), headsH = __ballot_sync(FULL_MASK, headH);
    uint32_t maskL = 0u, maskH = 0u;
    if (headL) maskL = spec_probe_cell(tp, keyLo);
    if (headH) maskH = spec_probe_cell(tp, keyHi);
    const unsigned liveL = __ballot_sync(FULL_MASK, maskL != 0u), liveH = __ballot_sync(FULL_MASK, maskH != 0u);
    if (!(liveL | liveH)) return;
    const unsigned le = 0xffffffffu >> (31u - lane);
    const int hlL = (31 - __clz((int)(headsL & le))) & 31, hlH = (31 - __clz((int)(headsH & le))) & 31;
    const uint32_t mL = __shfl_sync(FULL_MASK, maskL, hlL), mH = __shfl_sync(FULL_MASK, maskH, hlH);
    const uint32_t myL = wantLo ? mL : 0u, myH = wantHi ? mH : 0u;
    const int nL = __popc(liveL), nCells = nL + __popc(liveH);
    const int cellL = __popc(liveL & ((1u << hlL) - 1u)), cellH = nL + __popc(liveH & ((1u << hlH) - 1u));
    if (maskL) sh.cell[cellL] = make_uint2(keyLo, maskL);      // head lanes of non-empty cells (their hl is their own lane)
    if (maskH) sh.cell[cellH] = make_uint2(keyHi, maskH);
    __syncwarp();
    const int R = tp.R, Rm = R - 1, logR = tp.logR;
    const unsigned corner = lane & 7u;
    for (int c0 = 0; c0 < nCells; c0 += 4) {
        const int c = c0 + (int)(lane >> 3);
        if (c < nCells) {
            const uint2 cm = sh.cell[c];
#if VGI_TRACE_SPEC_GATHER_ALL
            if (!((cm.y >> corner) & 1u)) sh.val[lane] = make_float4(0.f, 0.f, 0.f, 0.f);   // every parked slot is gathered
#endif
            if ((cm.y >> corner) & 1u) {
                const uint32_t vox = cm.x;
                const uint32_t ix = vox & (uint32_t)Rm, iy = (vox >> logR) & (uint32_t)Rm, iz = (vox >> (2 * logR)) & (uint32_t)Rm;
                int off = 0;
                if (corner & 1u) off += (ix == (uint32_t)Rm) ? -Rm : 1;
                if (corner & 2u) off += ((iy == (uint32_t)Rm) ? -Rm : 1) << logR;
                if (corner & 4u) off += ((iz == (uint32_t)Rm) ? -Rm : 1) << (2 * logR);
                const uint32_t* rec = reinterpret_cast<const uint32_t*>(tp.store + rec_index(vox + (uint32_t)off, logR));
                const uint32_t tx = __ldg(rec + ox), ty = __ldg(rec + oy), tz = __ldg(rec + oz);
                STAT(4, 1);
                const float2 kx2 = make_float2(kx, kx), ky2 = make_float2(ky, ky), kz2 = make_float2(kz, kz);
                float2 lo = __fmul2_rn(kx2, unpack2(tx, 0x7540u, 0x7541u)), hi = __fmul2_rn(kx2, unpack2(tx, 0x7542u, 0x7543u));
                lo = __ffma2_rn(ky2, unpack2(ty, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(ky2, unpack2(ty, 0x7542u, 0x7543u), hi);
                lo = __ffma2_rn(kz2, unpack2(tz, 0x7540u, 0x7541u), lo); hi = __ffma2_rn(kz2, unpack2(tz, 0x7542u, 0x7543u), hi);
                sh.val[lane] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        }
        __syncwarp();
        const int qL = cellL - c0, qH = cellH - c0;
        if (myL && qL >= 0 && qL < 4) coop_gather_w<VGI_TRACE_SPEC_GATHER_ALL != 0>(wLo, myL, sh.val + 8 * qL, smpLo);
        if (myH && qH >= 0 && qH < 4) coop_gather_w<VGI_TRACE_SPEC_GATHER_ALL != 0>(wHi, myH, sh.val + 8 * qH, smpHi);
        __syncwarp();
    }
    anyLo = myL != 0u;
    anyHi = myH != 0u;
}

__global__ void __launch_bounds__(128, VGI_TRACE_SPECW_MINBLOCKS) k_trace_specular_warp(const __grid_constant__ TraceParams tp)
{
    __shared__ SpecWarpShared s_sh[4];
    const uint32_t n = *tp.spec_count;
    const unsigned lane = threadIdx.x & 31u;
    SpecWarpShared& sh = s_sh[threadIdx.x >> 5];
    const vgi_vct_params& p = tp.p;
    const float topLevel = (float)(tp.L - 1);
    for (;;) {
        uint32_t item = 0u;
        if (lane == 0u) item = atomicAdd(tp.spec_cursor, 1u);
        item = __shfl_sync(FULL_MASK, item, 0);
        if (item >= n) break;
        const uint32_t pi = tp.spec_list[item];
        const int px = (int)(pi % (uint32_t)tp.width), py = (int)(pi / (uint32_t)tp.width);
        PixelSetup s;
        if (!pixel_setup(tp, px, py, s)) continue;       // cannot happen: listed pixels are covered
        float dir[3], startPos[3], spec[3];
        {
            const float I[3] = { -s.view[0], -s.view[1], -s.view[2] };   // reflect(-view, normal) = I - 2 dot(N, I) N
            const float dn = dot3(s.normal, I);
#pragma unroll
            for (int k = 0; k < 3; ++k) dir[k] = I[k] - 2.0f * dn * s.normal[k];
        }
        const uint32_t rb = (uint32_t)(s.perceptualRoughness * 255.0f + 0.5f) & 255u;
        const float coneCoefficient = __ldg(tp.spec_coeff + rb);
        const uint32_t nSteps = __ldg(tp.spec_cnt + rb);
        const float2* tab = tp.spec_tab + (size_t)rb * tp.spec_stride;
        const float startLevel = s.minLevel;
        const float voxelSize0 = p.voxel_size * exp2f(startLevel);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            startPos[k] = s.startPos[k] + dir[k] * voxelSize0 * p.trace_start_offset * 0.5f;
            spec[k] = s.specularColor[k] * p.indirect_specular_intensity;
        }
        (void)coneCoefficient;
        const ConeFaces cf = cone_faces(dir);
        const uint32_t ox = cf.negX ? 1u : 0u, oy = cf.negY ? 3u : 2u, oz = cf.negZ ? 5u : 4u;
        float acc[3] = { 0.f, 0.f, 0.f };
        float carryA = 1.0f, carryO = 1.0f;       // 1 - result.a, 1 - occlusion before the batch
        for (uint32_t base = 0u; base < nSteps; base += 32u) {
            const uint32_t k = base + lane;
            const bool valid = k < nSteps;
            const float2 sl = __ldg(tab + k);       // rows are padded to a multiple of 32
            const float step = sl.x, lodk = sl.y;
            float prevStep = __shfl_up_sync(FULL_MASK, step, 1);
            if (lane == 0u) prevStep = base ? __ldg(&tab[base - 1u].x) : 0.0f;
            const float seg = k == 0u ? voxelSize0 : step - prevStep;
            uint32_t keyLo = 0u, keyHi = 0u;
            float wLo[3] = { 0.f, 0.f, 0.f }, wHi[3] = { 0.f, 0.f, 0.f };
            float curLevel = 0.0f, fr = 0.0f;
            if (valid) {
                STAT(0, 1);
                float position[3], d[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    position[a] = startPos[a] + dir[a] * step;
                    d[a] = p.volume_center[a] - position[a];
                }
                const float minLevel = min_level_from_dd(tp, dot3(d, d));
                curLevel = fminf(fmaxf(fmaxf(startLevel, lodk), minLevel), topLevel);
                const float fl = floorf(curLevel);
                fr = curLevel - fl;
                const float posV[3] = { position[0] * tp.vox_scale0, position[1] * tp.vox_scale0, position[2] * tp.vox_scale0 };
                keyLo = level_cell(tp, posV, (int)fl, wLo);
                // Q17: no second sample when the level is integral (computed unconditionally: no divergent branch)
                keyHi = level_cell(tp, posV, min((int)fl + 1, tp.L - 1), wHi);
            }
            float smp[4] = { 0.f, 0.f, 0.f, 0.f }, up[4] = { 0.f, 0.f, 0.f, 0.f };
#if VGI_TRACE_SPEC_MERGED
            bool anyLo, anyHi;
            spec_both_levels(tp, valid, keyLo, wLo, valid && fr > 0.0f, keyHi, wHi, lane, ox, oy, oz, cf.kx, cf.ky, cf.kz, sh, smp, up,
                             anyLo, anyHi);
#else
            const bool anyLo = spec_level_pass(tp, valid, keyLo, wLo, lane, ox, oy, oz, cf.kx, cf.ky, cf.kz, sh, smp);
            const bool anyHi = spec_level_pass(tp, valid && fr > 0.0f, keyHi, wHi, lane, ox, oy, oz, cf.kx, cf.ky, cf.kz, sh, up);
#endif
            const bool any = anyLo || anyHi;
            if (!__any_sync(FULL_MASK, any)) { STAT(2, 1); continue; }        // the whole batch is empty space
            float fa = 1.0f, fo = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
            if (any) {
                if (fr > 0.0f) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) smp[c] = smp[c] * (1.0f - fr) + up[c] * fr;
                }
                const float voxelSize = p.voxel_size * exp2f(curLevel);
                const float correction = __fdividef(seg, voxelSize);
                float opacity = 0.0f;
                if (smp[3] > 0.0f) opacity = f_clamp(1.0f - exp2f(correction * __log2f(1.0f - smp[3])), 0.0f, 1.0f);
                cr = smp[0] * correction; cg = smp[1] * correction; cb = smp[2] * correction;
                fa = 1.0f - opacity;
                fo = 1.0f - __fdividef(opacity, 1.0f + (step + voxelSize) * p.occlusion_decay);
            }
            // inclusive prefix products over the batch
            float pa = fa, po = fo;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const float ta = __shfl_up_sync(FULL_MASK, pa, dlt), to = __shfl_up_sync(FULL_MASK, po, dlt);
                if (lane >= (unsigned)dlt) { pa *= ta; po *= to; }
            }
            float ea = __shfl_up_sync(FULL_MASK, pa, 1);
            if (lane == 0u) ea = 1.0f;
            const float k1 = carryA * ea;               // clamp(1 - result.a, 0, 1) before this lane's step
            acc[0] += k1 * cr; acc[1] += k1 * cg; acc[2] += k1 * cb;
            carryA *= __shfl_sync(FULL_MASK, pa, 31);
            carryO *= __shfl_sync(FULL_MASK, po, 31);
            if (carryO < 2.98023224e-8f || carryA < 1e-10f) break;
        }
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) {
            acc[0] += __shfl_xor_sync(FULL_MASK, acc[0], dlt);
            acc[1] += __shfl_xor_sync(FULL_MASK, acc[1], dlt);
            acc[2] += __shfl_xor_sync(FULL_MASK, acc[2], dlt);
        }
        if (lane == 0u) tp.out_specular[pi] = make_float4(acc[0] * spec[0], acc[1] * spec[1], acc[2] * spec[2], 1.0f);
    }
}

int vgi_launch_trace_svo(vgi_ctx* c, const TraceParams& tp, cudaStream_t s)
{
    const int rows = tp.y1 - tp.y0;
    if (rows <= 0 || tp.width <= 0) return 0;
    const int tileW = tp.p.enable_32_cones ? 8 : 16;
    const unsigned grid = (unsigned)(((tp.width + tileW - 1) / tileW) * ((rows + TILE_H - 1) / TILE_H));
    c->timer.begin("k_trace_svo", s);
    if (tp.p.enable_32_cones) k_trace_main<32, true, 8><<<grid, 128, 0, s>>>(tp);
    else k_trace_main<16, true, 16><<<grid, 128, 0, s>>>(tp);
    c->timer.end(s);
    return 1;
}

#if VGI_TRACE_BRICK_STORE
__global__ void __launch_bounds__(256) k_brick_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nrec, int logR)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += (size_t)gridDim.x * blockDim.x) {
        const size_t j = rec_index((uint32_t)i, logR);
        dst[2 * j] = src[2 * i];
        dst[2 * j + 1] = src[2 * i + 1];
    }
}
static VoxelRecord* g_brick_store = nullptr;
static size_t g_brick_records = 0;
#endif

int vgi_launch_trace(vgi_ctx* c, const TraceParams& tp_in, cudaStream_t s)
{
    int n = 0;
#if VGI_TRACE_BRICK_STORE
    TraceParams tp = tp_in;
    {
        const size_t nrec = ((size_t)tp.R * tp.R * tp.R) * tp.L;
        if (g_brick_records != nrec) {
            cudaFree(g_brick_store);
            cudaMalloc(&g_brick_store, nrec * sizeof(VoxelRecord));
            g_brick_records = nrec;
        }
        k_brick_copy<<<148 * 16, 256, 0, s>>>(reinterpret_cast<const uint4*>(tp.store), reinterpret_cast<uint4*>(g_brick_store), nrec, tp.logR);
        tp.store = g_brick_store;
    }
#else
    const TraceParams& tp = tp_in;
#endif
    cudaMemsetAsync(tp.spec_count, 0, 2 * sizeof(uint32_t), s); // list length + work cursor
    const int rows = tp.y1 - tp.y0;
    if (rows <= 0 || tp.width <= 0) return 0;
    const int tileRows = (rows + TILE_H - 1) / TILE_H;
    const int myTileRows = tileRows > tp.tile_phase ? (tileRows - tp.tile_phase + tp.tile_stride - 1) / tp.tile_stride : 0;
    if (myTileRows <= 0) return 0;
    const int tileW = 8; // 8 x 8 tiles: measured faster than 16 x 8 for the clipmap march (tighter footprints per warp)
    const unsigned grid = (unsigned)(((tp.width + tileW - 1) / tileW) * myTileRows);
    const uint32_t mode = tp.p.rendering_mode;
    if ((mode == 6 || mode == 8) && VGI_TRACE_SPEC_WARP && tp.spec_tab && c->trace_spec_blocks > 0u) {
        // Both marches at once. The diffuse march leaves a third of the issue slots idle and the specular one ends in a
        // tail of a few long cones; a few resident specular blocks per SM beside the diffuse blocks fill the former, and
        // a second wave of specular workers behind k_trace_main takes over the SMs for whatever is left of the list.
        if (!c->spec_stream) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = served first
            if (cudaStreamCreateWithPriority(&c->spec_stream, cudaStreamNonBlocking, hi) != cudaSuccess) c->spec_stream = nullptr;
            else {
                cudaEventCreateWithFlags(&c->ev_spec_fork, cudaEventDisableTiming);
                cudaEventCreateWithFlags(&c->ev_spec_done, cudaEventDisableTiming);
            }
        }
        if (c->spec_stream) {
            static int perSmW = 0, sms = 148;
            if (!perSmW) {
                int dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmW, k_trace_specular_warp, 128, 0);
                if (perSmW <= 0) perSmW = 4;
            }
            TraceParams tq = tp;
            tq.spec_presplit = 1;
            c->timer.begin("cone_trace(main||specular)", s);
            cudaEvent_t joint_b = c->timer.enabled && !c->timer.pending.empty() ? c->timer.pending.back().b : nullptr;
            k_spec_classify<<<(unsigned)(((tp.width + 31) / 32) * myTileRows), 256, 0, s>>>(tq); ++n;
            cudaEventRecord(c->ev_spec_fork, s);
            cudaStreamWaitEvent(c->spec_stream, c->ev_spec_fork, 0);
            const int beside = (int)c->trace_spec_blocks < perSmW ? (int)c->trace_spec_blocks : perSmW;
            c->timer.begin("k_trace_specular_warp", c->spec_stream);
            k_trace_specular_warp<<<sms * beside, 128, 0, c->spec_stream>>>(tq); ++n;
            c->timer.end(c->spec_stream);
            cudaEventRecord(c->ev_spec_done, c->spec_stream);
            c->timer.begin("k_trace_main", s);
            if (tp.p.enable_32_cones) k_trace_main<32, false, 8><<<grid, 128, 0, s>>>(tq);
            else k_trace_main<16, false, 8><<<grid, 128, 0, s>>>(tq);
            ++n;
            c->timer.end(s);
            if (c->mark_main_done) cudaEventRecord(c->mark_main_done, s); // the diffuse image is complete here
            if (perSmW > beside) {      // second wave: exits at once when the list is already exhausted
                k_trace_specular_warp<<<sms * (perSmW - beside), 128, 0, s>>>(tq); ++n;
            }
            cudaStreamWaitEvent(s, c->ev_spec_done, 0);
            if (joint_b) cudaEventRecord(joint_b, s);
            return n;
        }
    }
    c->timer.begin("k_trace_main", s);
    if (tp.p.enable_32_cones) k_trace_main<32, false, 8><<<grid, 128, 0, s>>>(tp);
    else k_trace_main<16, false, 8><<<grid, 128, 0, s>>>(tp);
    ++n;
    c->timer.end(s);
    if (c->mark_main_done) cudaEventRecord(c->mark_main_done, s); // the diffuse image is complete here
    if (mode == 6 || mode == 8) {
        // persistent kernels: exactly as many blocks as fit on the device at once
        static int specBlocks = 0, specWarpBlocks = 0;
        if (!specBlocks) {
            int perSm = 0, perSmW = 0, dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace_specular, 128, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmW, k_trace_specular_warp, 128, 0);
            specBlocks = sms * (perSm > 0 ? perSm : 4);
            specWarpBlocks = sms * (perSmW > 0 ? perSmW : 4);
        }
        if (VGI_TRACE_SPEC_WARP && tp.spec_tab) {
            c->timer.begin("k_trace_specular_warp", s);
            k_trace_specular_warp<<<specWarpBlocks, 128, 0, s>>>(tp); ++n;
        } else {
            c->timer.begin("k_trace_specular", s);
            k_trace_specular<<<specBlocks, 128, 0, s>>>(tp); ++n;
        }
        c->timer.end(s);
    }
    return n;
}
