// vgi_atlas.cu — stand-alone equivalents of the reference's clipmap compute helpers, operating on
// caller-owned atlases in the REFERENCE image layout (RGBA8, x fastest, W=(R+2)*6, H=(R+2)*L, D=R+2;
// ref: Voxelizer.h:40-52). A host that keeps its own Vulkan images (imported through
// vgi_import_vk_memory) can swap the four helper passes one at a time:
//   ClipmapCleaner::cmdClear*ClipRegion  (ClipmapCleaner.cpp:94-134, clipmapCleaning.comp:17-31)
//   CopyAlpha::cmdImageCopyAlpha         (CopyAlpha.cpp:93-146, copyAlphaImage.comp:16-29)
//   DownSampler::cmdDownSample*          (DownSampler.cpp:115-168, opacityDownSample.comp / radianceDownSample.comp)
//   BorderWrapper::cmdWrappingOpacityBorder (BorderWrapper.cpp:123-151, borderWrapping.comp:14-37)
// The fused pipeline (vgi_build.cu) does not use these; they exist for drop-in parity with the helper
// entry points the passes look up from the RenderPassManager blackboard. Compiled with -fmad=false.
#include "vgi_device.cuh"

struct AtlasDim {
    int R, L, rb;
    size_t W, H;
};

DEVFN AtlasDim atlas_dim(int R, int L)
{
    AtlasDim d;
    d.R = R; d.L = L; d.rb = R + 2;
    d.W = (size_t)d.rb * 6;
    d.H = (size_t)d.rb * L;
    return d;
}

DEVFN uint32_t* atlas_px(const AtlasDim& d, uint8_t* atlas, size_t x, size_t y, size_t z)
{
    return reinterpret_cast<uint32_t*>(atlas) + (z * d.H + y) * d.W + x;
}

__global__ void __launch_bounds__(256) k_atlas_clear_region(uint8_t* __restrict__ atlas, int R, int L, int mx, int my, int mz,
                                                             uint32_t ex, uint32_t ey, uint32_t ez, int level)
{
    const AtlasDim d = atlas_dim(R, L);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)ex * ey * ez) return;
    const int x = (int)(i % ex), y = (int)((i / ex) % ey), z = (int)(i / ((size_t)ex * ey));
    // clipmapCleaning.comp:22-24: (pos + minCorner) % resolution, made well defined for negative corners
    const int px = ((x + mx) % R + R) % R + 1;
    const int py = ((y + my) % R + R) % R + 1 + level * d.rb;
    const int pz = ((z + mz) % R + R) % R + 1;
#pragma unroll
    for (int f = 0; f < 6; ++f) *atlas_px(d, atlas, (size_t)px + (size_t)f * d.rb, py, pz) = 0u;
}

__global__ void __launch_bounds__(256) k_atlas_copy_alpha(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int R, int L, int level)
{
    const AtlasDim d = atlas_dim(R, L);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)R * R * R) return;
    const int x = (int)(i % R), y = (int)((i / R) % R), z = (int)(i / ((size_t)R * R));
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        uint32_t* o = atlas_px(d, dst, (size_t)x + 1 + (size_t)f * d.rb, (size_t)y + 1 + (size_t)level * d.rb, (size_t)z + 1);
        const uint32_t s = *atlas_px(d, const_cast<uint8_t*>(src), (size_t)x + 1 + (size_t)f * d.rb, (size_t)y + 1 + (size_t)level * d.rb, (size_t)z + 1);
        *o = (*o & 0x00ffffffu) | (s & 0xff000000u);
    }
}

// One thread per coarse voxel of the centre half; the children live in the rows of level-1, the output in
// the rows of `level`, so the in-place update has no hazard inside a launch.
__global__ void __launch_bounds__(128) k_atlas_downsample(uint8_t* __restrict__ atlas, int R, int L, int band, int pmx, int pmy, int pmz,
                                                           int level, int which)
{
    const AtlasDim d = atlas_dim(R, L);
    const int half = R >> 1;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y, gz = blockIdx.z;
    if (gx >= half) return;
    const int g[3] = { gx, gy, gz };
    const int prevMin[3] = { pmx, pmy, pmz };
    int wpos[3], pstart[3];
    float dist[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int cur = (prevMin[k] >> 1) + g[k];
        wpos[k] = cur & (R - 1);
        pstart[k] = (cur << 1) & (R - 1);
        const float center = (float)(prevMin[k] >> 1) + (float)((uint32_t)half >> 1);
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
    for (int f = 0; f < 6; ++f) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            v[i] = *atlas_px(d, atlas, (size_t)(pstart[0] + (i & 1) + 1 + f * d.rb), (size_t)(pstart[1] + ((i >> 1) & 1) + 1 + d.rb * (level - 1)),
                             (size_t)(pstart[2] + (i >> 2) + 1));
        uint32_t* outp = atlas_px(d, atlas, (size_t)(wpos[0] + 1 + f * d.rb), (size_t)(wpos[1] + 1 + d.rb * level), (size_t)(wpos[2] + 1));
        const uint32_t own = *outp;
        const int axis = f >> 1, abit = 1 << axis;
        uint32_t result;
        if (which == 0) {
            float s = 0.0f;
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
                const int base = axis == 0 ? (pr << 1) : (axis == 1 ? ((pr & 1) | ((pr & 2) << 1)) : pr);
                const int i0 = base | ((f & 1) ? abit : 0), i1 = base | ((f & 1) ? 0 : abit);
                const float a0 = unorm8_to_f(v[i0] >> 24), a1 = unorm8_to_f(v[i1] >> 24);
                s = s + a0;
                s = s + (1.0f - a0) * a1;
            }
            const float ds = s * 0.25f;
            // opacityDownSample.comp:66-73: a = mix(ds, own.r, lerp); g = a; r and b unchanged
            const uint32_t a = f_to_unorm8(f_mix(ds, unorm8_to_f(own & 0xffu), lerpFactor));
            result = (own & 0x00ff00ffu) | (a << 8) | (a << 24);
        } else {
            result = 0u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float s = 0.0f;
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const int base = axis == 0 ? (pr << 1) : (axis == 1 ? ((pr & 1) | ((pr & 2) << 1)) : pr);
                    const int i0 = base | ((f & 1) ? abit : 0), i1 = base | ((f & 1) ? 0 : abit);
                    s = s + unorm8_to_f((v[i0] >> (8 * c)) & 0xffu);
                    s = s + (1.0f - unorm8_to_f(v[i0] >> 24)) * unorm8_to_f((v[i1] >> (8 * c)) & 0xffu);
                }
                const float ds = s * 0.25f;
                result |= f_to_unorm8(f_mix(ds, unorm8_to_f((own >> (8 * c)) & 0xffu), lerpFactor)) << (8 * c);
            }
        }
        *outp = result;
    }
}

__global__ void __launch_bounds__(256) k_atlas_wrap(uint8_t* __restrict__ atlas, int R, int L, int literal)
{
    const AtlasDim d = atlas_dim(R, L);
    const int rb = d.rb;
    const int lim = literal ? (((R + 2) >> 3) << 3) : rb; // BorderWrapper.cpp:139 dispatches (R+2)>>3 groups of 8 (Q4)
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rb * rb * rb) return;
    const int x = (int)(i % rb), y = (int)((i / rb) % rb), z = (int)(i / ((size_t)rb * rb));
    if (x >= lim || y >= lim || z >= lim) return;
    if (x < R + 1 && y < R + 1 && z < R + 1 && x > 0 && y > 0 && z > 0) return;
    const int rx = ((x + R - 1) & (R - 1)) + 1, ry = ((y + R - 1) & (R - 1)) + 1, rz = ((z + R - 1) & (R - 1)) + 1;
    for (int l = 0; l < L; ++l)
#pragma unroll
        for (int f = 0; f < 6; ++f)
            *atlas_px(d, atlas, (size_t)(x + rb * f), (size_t)(y + rb * l), (size_t)z) =
                *atlas_px(d, atlas, (size_t)(rx + rb * f), (size_t)(ry + rb * l), (size_t)rz);
}

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

int vgi_launch_atlas_clear(uint8_t* atlas, int R, int L, const int32_t* mc, const uint32_t* ext, int level, cudaStream_t s)
{
    const size_t n = (size_t)ext[0] * ext[1] * ext[2];
    if (!n) return 0;
    k_atlas_clear_region<<<cdiv(n, 256), 256, 0, s>>>(atlas, R, L, mc[0], mc[1], mc[2], ext[0], ext[1], ext[2], level);
    return 1;
}

int vgi_launch_atlas_copy_alpha(uint8_t* dst, const uint8_t* src, int R, int L, int level, cudaStream_t s)
{
    k_atlas_copy_alpha<<<cdiv((size_t)R * R * R, 256), 256, 0, s>>>(dst, src, R, L, level);
    return 1;
}

int vgi_launch_atlas_downsample(uint8_t* atlas, int R, int L, int band, const int32_t* prev_min, int level, int which, cudaStream_t s)
{
    const int half = R >> 1;
    dim3 grid(cdiv(half, 128), half, half);
    k_atlas_downsample<<<grid, half < 128 ? half : 128, 0, s>>>(atlas, R, L, band, prev_min[0], prev_min[1], prev_min[2], level, which);
    return 1;
}

int vgi_launch_atlas_wrap(uint8_t* atlas, int R, int L, int literal, cudaStream_t s)
{
    const size_t rb = (size_t)R + 2;
    k_atlas_wrap<<<cdiv(rb * rb * rb, 256), 256, 0, s>>>(atlas, R, L, literal);
    return 1;
}
