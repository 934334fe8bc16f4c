// vgi_post.cu — the pass that follows cone tracing: final = diffuse + filtered specular,
// optional Uncharted-2 tonemap. ref: VFS/RenderPass/SpecularFilterPass.cpp:73-91, VFS/Shaders/specularFilter.frag:25-53,
// filter.glsl:9-24 (gaussian), :27-63 (bilateral), tonemapping.glsl:4-26.
//
// Both filters read the specular image through "VCTSampler" (LINEAR, CLAMP_TO_EDGE) at the texCoord of a fullscreen quad
// of the image's own size, so every tap lands on a texel centre plus a fixed offset:
//  * gaussian: blurSize = 0.01 is divided by the texture size (filter.glsl:11), i.e. the 256 taps lie within 0.01 TEXEL
//    of the centre. Their bilinear footprints therefore stay inside the 3x3 neighbourhood and the whole sum collapses to
//    nine weights that depend on nothing but the tap pattern: the host folds the 257 taps once (same binary32 loop
//    counters as the shader) and the kernel does 9 loads per pixel instead of 1028 — HBM-bound (3 float4 images).
//  * bilateral: 15 x 15 taps at whole-texel offsets = plain texel fetches; a block stages its tile + 7-texel apron in
//    shared memory and every thread walks the 225 taps from there (MUFU.EX2-bound).
// Floating-point tolerance work (compared with the oracle at 1e-3): FMA contraction on.
#include "vgi_internal.h"

#define DEVFN static __device__ __forceinline__

struct PostParams {
    const float4* diffuse;
    const float4* specular;
    float4*       out;
    int           width, height;
    float         gauss_w[9];       // 3x3 weights of the folded gaussian taps, row-major (dy, dx)
    float         inv_gamma, exposure, white;
    int           tonemap;
};

__constant__ float c_bilateral_kernel[15] = { // ref: filter.glsl:37-41
    0.031225216f, 0.033322271f, 0.035206333f, 0.036826804f, 0.038138565f, 0.039104044f, 0.039695028f, 0.039894000f,
    0.039695028f, 0.039104044f, 0.038138565f, 0.036826804f, 0.035206333f, 0.033322271f, 0.031225216f };

DEVFN float uncharted2(float c) // ref: tonemapping.glsl:9-19
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((c * (A * c + C * B) + D * E) / (c * (A * c + B) + D * F)) - E / F;
}

DEVFN float4 finish_pixel(const PostParams& p, float4 fc, float sr, float sg, float sb)
{
    fc.x += sr; fc.y += sg; fc.z += sb;
    if (p.tonemap) { // ref: tonemapping.glsl:21-26
        fc.x = powf(uncharted2(fc.x * p.exposure) * p.white, p.inv_gamma);
        fc.y = powf(uncharted2(fc.y * p.exposure) * p.white, p.inv_gamma);
        fc.z = powf(uncharted2(fc.z * p.exposure) * p.white, p.inv_gamma);
    }
    return fc;
}

__global__ void __launch_bounds__(256) k_filter_gaussian(const __grid_constant__ PostParams p)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.width || y >= p.height) return;
    const int xs[3] = { max(x - 1, 0), x, min(x + 1, p.width - 1) };
    const int ys[3] = { max(y - 1, 0), y, min(y + 1, p.height - 1) };
    float sr = 0.f, sg = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float4 t = __ldg(p.specular + (size_t)ys[j] * p.width + xs[i]);
            const float w = p.gauss_w[j * 3 + i];
            sr += w * t.x; sg += w * t.y; sb += w * t.z;
        }
    const size_t pi = (size_t)y * p.width + x;
    p.out[pi] = finish_pixel(p, __ldg(p.diffuse + pi), sr, sg, sb);
}

#define BIL_R 7
#define BIL_TW 32
#define BIL_TH 8
__global__ void __launch_bounds__(BIL_TW * BIL_TH) k_filter_bilateral(const __grid_constant__ PostParams p)
{
    constexpr int SW = BIL_TW + 2 * BIL_R, SH = BIL_TH + 2 * BIL_R;
    __shared__ float4 s_tile[SH][SW];
    const int x0 = blockIdx.x * BIL_TW - BIL_R, y0 = blockIdx.y * BIL_TH - BIL_R;
    for (int i = threadIdx.x; i < SW * SH; i += BIL_TW * BIL_TH) {
        const int sx = i % SW, sy = i / SW;
        const int gx = min(max(x0 + sx, 0), p.width - 1), gy = min(max(y0 + sy, 0), p.height - 1); // CLAMP_TO_EDGE
        s_tile[sy][sx] = __ldg(p.specular + (size_t)gy * p.width + gx);
    }
    __syncthreads();
    const int lx = threadIdx.x & (BIL_TW - 1), ly = threadIdx.x / BIL_TW;
    const int x = blockIdx.x * BIL_TW + lx, y = blockIdx.y * BIL_TH + ly;
    if (x >= p.width || y >= p.height) return;
    const float4 c = s_tile[ly + BIL_R][lx + BIL_R];
    // normpdf3(v, 10) / normpdf(0, 10) = exp(-0.5 |v|^2 / 100) (filter.glsl:32-35,50,58); in base 2 for MUFU.EX2
    const float k2 = -0.5f / 100.0f * 1.4426950408889634f;
    float Z = 0.f, fr = 0.f, fg = 0.f, fb = 0.f;
#pragma unroll 1
    for (int j = -BIL_R; j <= BIL_R; ++j) {
        const float kj = c_bilateral_kernel[BIL_R + j];
#pragma unroll
        for (int i = -BIL_R; i <= BIL_R; ++i) {
            const float4 s = s_tile[ly + BIL_R + j][lx + BIL_R + i];
            const float dr = s.x - c.x, dg = s.y - c.y, db = s.z - c.z;
            const float f = exp2f(k2 * (dr * dr + dg * dg + db * db)) * (kj * c_bilateral_kernel[BIL_R + i]);
            Z += f;
            fr += f * s.x; fg += f * s.y; fb += f * s.z;
        }
    }
    const float iz = 1.0f / Z;
    const size_t pi = (size_t)y * p.width + x;
    p.out[pi] = finish_pixel(p, __ldg(p.diffuse + pi), fr * iz, fg * iz, fb * iz);
}

// Folds the gaussian tap pattern of filter.glsl:9-24 into 3x3 texel weights (see the header comment).
static void fold_gaussian_taps(float* w9)
{
    const float DOUBLE_PI = 6.28318530718f, DIRECTIONS = 32.0f, QUALITY = 8.0f, blurSize = 0.01f;
    double acc[3][3] = {};
    acc[1][1] = 1.0; // the centre tap
    for (float d = 0.0f; d < DOUBLE_PI; d += DOUBLE_PI / DIRECTIONS)
        for (float i = 1.0f / QUALITY; i <= 1.0f; i += 1.0f / QUALITY) {
            const double ox = (double)(cosf(d) * blurSize) * i, oy = (double)(sinf(d) * blurSize) * i; // texel units
            double wx[3] = { 0, 0, 0 }, wy[3] = { 0, 0, 0 };
            if (ox >= 0) { wx[1] = 1.0 - ox; wx[2] = ox; } else { wx[0] = -ox; wx[1] = 1.0 + ox; }
            if (oy >= 0) { wy[1] = 1.0 - oy; wy[2] = oy; } else { wy[0] = -oy; wy[1] = 1.0 + oy; }
            for (int b = 0; b < 3; ++b)
                for (int a = 0; a < 3; ++a) acc[b][a] += wy[b] * wx[a];
        }
    const double norm = (double)(QUALITY * DIRECTIONS - 15.0f);
    for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) w9[b * 3 + a] = (float)(acc[b][a] / norm);
}

int vgi_launch_specular_filter(vgi_ctx* c, const void* diffuse, const void* specular, uint32_t width, uint32_t height,
                               const vgi_filter_params* prm, void* out, cudaStream_t s)
{
    PostParams p;
    p.diffuse = (const float4*)diffuse;
    p.specular = (const float4*)specular;
    p.out = (float4*)out;
    p.width = (int)width; p.height = (int)height;
    fold_gaussian_taps(p.gauss_w);
    p.tonemap = prm->tonemap_enable == 1;
    p.inv_gamma = 1.0f / prm->tonemap_gamma;
    p.exposure = prm->tonemap_exposure;
    {
        const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f, W = 11.2f;
        p.white = 1.0f / (((W * (A * W + C * B) + D * E) / (W * (A * W + B) + D * F)) - E / F);
    }
    const dim3 grid((width + 31) / 32, (height + 7) / 8);
    if (prm->filter_method == 1) {
        c->timer.begin("k_filter_gaussian", s);
        k_filter_gaussian<<<grid, 256, 0, s>>>(p);
    } else { // 0 and every other value: bilateral (specularFilter.frag:35-46)
        c->timer.begin("k_filter_bilateral", s);
        k_filter_bilateral<<<grid, BIL_TW * BIL_TH, 0, s>>>(p);
    }
    c->timer.end(s);
    return 1;
}
