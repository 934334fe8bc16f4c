// glsl_shim.h — just enough of GLSL 4.50 in C++20 to compile the reference's shader TEXT on the CPU.
//
// TEST INFRASTRUCTURE ONLY (part of oracle/). oracle/glsl_shim/build_ref.py reads the shaders where they lie
// under /root/reference/VFS/Shaders, applies a short list of purely syntactic rewrites (float-literal suffixes,
// `layout(...)` removal, interface blocks -> structs / globals, `out`/`inout` -> references, swizzles -> calls)
// and pipes the result, between this header and a small driver, into g++; only the resulting
// oracle/_ref/libvgi_refshaders.so is kept. The arithmetic that runs is therefore the reference's own shader
// code; what this header supplies is what the Vulkan implementation would: vector types, built-in functions
// and the fixed-function image / sampler semantics (Vulkan 1.2 spec chapters 16 "Image Operations"):
//   * RGBA8 UNORM load = c / 255, store = clamp to [0,1], * 255, round to nearest (same rule as the oracle);
//   * LINEAR filtering: unnormalised coordinate c = s * size - 0.5, i0 = floor(c), weights (1 - w, w),
//     lerp in x, then y, then z; REPEAT wraps texel indices modulo size, CLAMP_TO_EDGE clamps them,
//     CLAMP_TO_BORDER returns the border colour (opaque / transparent black: 0 in the channels used);
//   * built-ins computed in binary32 with the C library's float functions; normalize(v) = v / length(v),
//     mix(x, y, a) = x * (1 - a) + y * a, dot products summed left to right.
// No glm here on purpose: a closed, non-template overload set (like GLSL's own) keeps implicit conversions
// predictable, and -Werror=float-conversion turns any silent float -> int narrowing into a build error.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

// light.glsl is shared with the reference's C++ host and says `using namespace glm;` under __cplusplus
namespace glm {}

namespace glsl {

typedef unsigned int uint;

template <class T, int N> struct tvec;
template <class V> struct swz;

// ---- storage (anonymous structs in unions: GNU extension, fine for g++) ----
template <class T> struct tvec<T, 2> {
    typedef T elem; enum { dim = 2 };
    union { struct { T x, y; }; struct { T r, g; }; struct { T s, t; }; T d[2]; };
    tvec() : x(), y() {}
    explicit tvec(T v) : x(v), y(v) {}
    tvec(T a, T b) : x(a), y(b) {}
    template <class U> explicit tvec(const tvec<U, 2>& o) : x((T)o.x), y((T)o.y) {}
    template <class U> explicit tvec(const tvec<U, 3>& o) : x((T)o.x), y((T)o.y) {}
    template <class U> explicit tvec(const tvec<U, 4>& o) : x((T)o.x), y((T)o.y) {}
#include "glsl_vec_ops.inl"
};
template <class T> struct tvec<T, 3> {
    typedef T elem; enum { dim = 3 };
    union { struct { T x, y, z; }; struct { T r, g, b; }; struct { T s, t, p; }; T d[3]; };
    tvec() : x(), y(), z() {}
    explicit tvec(T v) : x(v), y(v), z(v) {}
    tvec(T a, T b, T c) : x(a), y(b), z(c) {}
    tvec(const tvec<T, 2>& a, T c) : x(a.x), y(a.y), z(c) {}
    tvec(T a, const tvec<T, 2>& b) : x(a), y(b.x), z(b.y) {}
    template <class U> explicit tvec(const tvec<U, 3>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U> explicit tvec(const tvec<U, 4>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
#include "glsl_vec_ops.inl"
};
template <class T> struct tvec<T, 4> {
    typedef T elem; enum { dim = 4 };
    union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; struct { T s, t, p, q; }; T d[4]; };
    tvec() : x(), y(), z(), w() {}
    explicit tvec(T v) : x(v), y(v), z(v), w(v) {}
    tvec(T a_, T b_, T c_, T d_) : x(a_), y(b_), z(c_), w(d_) {}
    tvec(const tvec<T, 3>& v, T d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    tvec(T a_, const tvec<T, 3>& v) : x(a_), y(v.x), z(v.y), w(v.z) {}
    tvec(const tvec<T, 2>& u, const tvec<T, 2>& v) : x(u.x), y(u.y), z(v.x), w(v.y) {}
    tvec(const tvec<T, 2>& u, T c_, T d_) : x(u.x), y(u.y), z(c_), w(d_) {}
    template <class U> explicit tvec(const tvec<U, 4>& o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
#include "glsl_vec_ops.inl"
};

typedef tvec<float, 2> vec2;   typedef tvec<float, 3> vec3;   typedef tvec<float, 4> vec4;
typedef tvec<int, 2> ivec2;    typedef tvec<int, 3> ivec3;    typedef tvec<int, 4> ivec4;
typedef tvec<uint, 2> uvec2;   typedef tvec<uint, 3> uvec3;   typedef tvec<uint, 4> uvec4;
typedef tvec<bool, 2> bvec2;   typedef tvec<bool, 3> bvec3;   typedef tvec<bool, 4> bvec4;

// ---- l-value swizzle: v.xyz() on a non-const vector. Converts to the vector type, assigns through. ----
template <class V> struct swz {
    typedef typename V::elem T;
    T* p[V::dim];
    operator V() const { V r; for (int i = 0; i < V::dim; ++i) r.d[i] = *p[i]; return r; }
    swz& operator=(const V& v) { for (int i = 0; i < V::dim; ++i) *p[i] = v.d[i]; return *this; }
    swz& operator=(const swz& o) { return *this = (V)o; }
#define GLSL_SWZ_ASSIGN(OP) \
    swz& operator OP(const V& v) { V t = (V)*this; t OP v; return *this = t; } \
    swz& operator OP(T s) { V t = (V)*this; t OP s; return *this = t; }
    GLSL_SWZ_ASSIGN(+=) GLSL_SWZ_ASSIGN(-=) GLSL_SWZ_ASSIGN(*=) GLSL_SWZ_ASSIGN(/=)
#undef GLSL_SWZ_ASSIGN
};

// ---- scalar built-ins (GLSL 8.1 - 8.3), binary32 ----
inline float radians(float d) { return d * 0.017453292519943295f; }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
// pow(x, y) is undefined for x < 0 in GLSL; GPUs evaluate exp2(y * log2(x)) = NaN for every negative base, whereas powf(negative,
// integer) is finite. Follow the hardware (the oracle's glsl_pow does the same; see the note on min / max below).
inline float pow(float x, float y) { return x < 0.0f ? NAN : ::powf(x, y); }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float exp2(float x) { return ::exp2f(x); }
inline float exp2(int x) { return ::exp2f((float)x); }          // GLSL promotes int -> float implicitly
inline float log2(float x) { return ::log2f(x); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float inversesqrt(float x) { return 1.0f / ::sqrtf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline int   abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float round(float x) { return ::roundf(x); }
inline float fract(float x) { return x - ::floorf(x); }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }
// NaN operands: GLSL leaves min / max / clamp undefined. IEEE 754 minNum / maxNum (the non-NaN operand wins) is what
// NVIDIA hardware (FMNMX), CUDA's fminf / fmaxf and the oracle's f_min / f_max composition in f_clamp do; it matters in
// traceCone when the filtered alpha exceeds 1 by an ulp (32-cone weights) and pow(1 - a, c) of a negative base is NaN:
// clamp(1 - NaN, 0, 1) is then 0, not NaN.
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int   min(int a, int b) { return b < a ? b : a; }
inline int   max(int a, int b) { return a < b ? b : a; }
inline uint  min(uint a, uint b) { return b < a ? b : a; }
inline uint  max(uint a, uint b) { return a < b ? b : a; }
inline float min(float a, int b) { return min(a, (float)b); }   // int -> float promotion
inline float max(float a, int b) { return max(a, (float)b); }
inline float min(int a, float b) { return min((float)a, b); }
inline float max(int a, float b) { return max((float)a, b); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int   clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline uint  clamp(uint x, uint lo, uint hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x)
{
    const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float length(float x) { return abs(x); }

// ---- matrices: column major like GLSL ----
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    explicit mat4(const float* m) { for (int i = 0; i < 4; ++i) c[i] = vec4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    friend vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
};
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(const mat4& m) { for (int i = 0; i < 3; ++i) c[i] = vec3(m.c[i]); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
    friend vec3 operator*(const mat3& m, const vec3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
};

// ---- images and samplers: what the fixed-function hardware would do ----
inline float unorm8_to_float(uint8_t c) { return (float)c / 255.0f; }
inline uint8_t float_to_unorm8(float x)
{
    if (!(x > 0.0f)) return 0;      // also NaN
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)(x * 255.0f + 0.5f);
}

struct image3D {        // rgba8 storage image, x fastest
    uint8_t* data; int W, H, D;
    bool inside(const ivec3& p) const { return p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < W && p.y < H && p.z < D; }
    uint8_t* at(const ivec3& p) const { return data + (((size_t)p.z * (size_t)H + (size_t)p.y) * (size_t)W + (size_t)p.x) * 4; }
};
inline vec4 imageLoad(const image3D& img, const ivec3& p)
{
    if (!img.inside(p)) return vec4(0.0f);   // robustBufferAccess-style: out-of-bounds loads return zero
    const uint8_t* t = img.at(p);
    return vec4(unorm8_to_float(t[0]), unorm8_to_float(t[1]), unorm8_to_float(t[2]), unorm8_to_float(t[3]));
}
inline void imageStore(const image3D& img, const ivec3& p, const vec4& v)
{
    if (!img.inside(p)) return;              // out-of-bounds stores are discarded
    uint8_t* t = img.at(p);
    t[0] = float_to_unorm8(v.x); t[1] = float_to_unorm8(v.y); t[2] = float_to_unorm8(v.z); t[3] = float_to_unorm8(v.w);
}
// r32ui view of the same memory (VoxelRadianceR32View)
struct uimage3D { uint32_t* data; int W, H, D; };
// optional per-thread log of the texels an invocation's atomics touched (drivers that inspect single fragments)
struct atomic_log { int n; int xyz[16][3]; };
inline thread_local atomic_log* g_atomic_log = nullptr;
inline uint imageAtomicCompSwap(const uimage3D& img, const ivec3& p, uint compare, uint value)
{
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= img.W || p.y >= img.H || p.z >= img.D) return compare;   // discarded
    uint32_t* t = img.data + ((size_t)p.z * (size_t)img.H + (size_t)p.y) * (size_t)img.W + (size_t)p.x;
    const uint old = *t;
    if (old == compare) {
        *t = value;
        if (g_atomic_log && g_atomic_log->n < 16) {
            int* e = g_atomic_log->xyz[g_atomic_log->n++];
            e[0] = p.x; e[1] = p.y; e[2] = p.z;
        }
    }
    return old;
}
inline uint atomicCompSwap(uint& mem, uint compare, uint value) { const uint old = mem; if (old == compare) mem = value; return old; }
inline uint atomicAdd(uint& mem, uint v) { const uint old = mem; mem = old + v; return old; }
inline uint atomicOr(uint& mem, uint v) { const uint old = mem; mem = old | v; return old; }
inline void barrier() {}
inline void memoryBarrier() {}

enum address_mode { REPEAT = 0, CLAMP_TO_EDGE = 1, CLAMP_TO_BORDER = 2 };
enum filter_mode { NEAREST = 0, LINEAR = 1 };

inline int wrap_index(long i, int n, int mode, bool& border)
{
    if (i >= 0 && i < n) return (int)i;      // in range: every mode agrees (and no integer division)
    if (mode == REPEAT) { long m = i % n; return (int)(m < 0 ? m + n : m); }
    if (mode == CLAMP_TO_EDGE) return (int)(i < 0 ? 0 : (i >= n ? n - 1 : i));
    if (i < 0 || i >= n) { border = true; return 0; }
    return (int)i;
}

struct sampler3D {      // rgba8 texels
    const uint8_t* data; int W, H, D; int address; int filter;
    vec4 texel(long x, long y, long z) const
    {
        bool border = false;
        const int ix = wrap_index(x, W, address, border), iy = wrap_index(y, H, address, border), iz = wrap_index(z, D, address, border);
        if (border) return vec4(0.0f);
        const uint8_t* t = data + (((size_t)iz * (size_t)H + (size_t)iy) * (size_t)W + (size_t)ix) * 4;
        return vec4(unorm8_to_float(t[0]), unorm8_to_float(t[1]), unorm8_to_float(t[2]), unorm8_to_float(t[3]));
    }
};
inline vec4 texelFetch(const sampler3D& s, const ivec3& p, int /*lod*/) { return s.texel(p.x, p.y, p.z); }
inline ivec3 textureSize(const sampler3D& s, int) { return ivec3(s.W, s.H, s.D); }
inline vec4 texture(const sampler3D& s, const vec3& c)
{
    if (s.filter == NEAREST)
        return s.texel((long)::floorf(c.x * (float)s.W), (long)::floorf(c.y * (float)s.H), (long)::floorf(c.z * (float)s.D));
    const float ux = c.x * (float)s.W - 0.5f, uy = c.y * (float)s.H - 0.5f, uz = c.z * (float)s.D - 0.5f;
    const float fx = ::floorf(ux), fy = ::floorf(uy), fz = ::floorf(uz);
    const float wx = ux - fx, wy = uy - fy, wz = uz - fz;
    const long ix = (long)fx, iy = (long)fy, iz = (long)fz;
    const vec4 x00 = s.texel(ix, iy, iz) * (1.0f - wx) + s.texel(ix + 1, iy, iz) * wx;
    const vec4 x10 = s.texel(ix, iy + 1, iz) * (1.0f - wx) + s.texel(ix + 1, iy + 1, iz) * wx;
    const vec4 x01 = s.texel(ix, iy, iz + 1) * (1.0f - wx) + s.texel(ix + 1, iy, iz + 1) * wx;
    const vec4 x11 = s.texel(ix, iy + 1, iz + 1) * (1.0f - wx) + s.texel(ix + 1, iy + 1, iz + 1) * wx;
    const vec4 y0 = x00 * (1.0f - wy) + x10 * wy;
    const vec4 y1 = x01 * (1.0f - wy) + x11 * wy;
    return y0 * (1.0f - wz) + y1 * wz;
}

struct sampler2D {      // float texels with `channels` components per texel (missing: 0, 0, 0, 1)
    const float* data; int W, H, channels; int address; int filter;
    vec4 texel(long x, long y) const
    {
        bool border = false;
        const int ix = wrap_index(x, W, address, border), iy = wrap_index(y, H, address, border);
        if (border) return vec4(0.0f, 0.0f, 0.0f, channels == 1 ? 1.0f : 0.0f);   // opaque black for depth, transparent otherwise
        const float* t = data + ((size_t)iy * (size_t)W + (size_t)ix) * (size_t)channels;
        vec4 r(0.0f, 0.0f, 0.0f, 1.0f);
        for (int k = 0; k < channels; ++k) r.d[k] = t[k];
        return r;
    }
};
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.W, s.H); }
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int) { return s.texel(p.x, p.y); }
inline vec4 texture(const sampler2D& s, const vec2& c)
{
    if (s.filter == NEAREST) return s.texel((long)::floorf(c.x * (float)s.W), (long)::floorf(c.y * (float)s.H));
    const float ux = c.x * (float)s.W - 0.5f, uy = c.y * (float)s.H - 0.5f;
    const float fx = ::floorf(ux), fy = ::floorf(uy);
    const float wx = ux - fx, wy = uy - fy;
    const long ix = (long)fx, iy = (long)fy;
    const vec4 r0 = s.texel(ix, iy) * (1.0f - wx) + s.texel(ix + 1, iy) * wx;
    const vec4 r1 = s.texel(ix, iy + 1) * (1.0f - wx) + s.texel(ix + 1, iy + 1) * wx;
    return r0 * (1.0f - wy) + r1 * wy;
}
inline vec4 textureLod(const sampler2D& s, const vec2& c, float /*lod: samplers are created with maxLod = 0*/) { return texture(s, c); }
// non-shadow sampler: projective divide of the coordinate, the third component is ignored (GLSL 8.9.2)
inline vec4 textureProj(const sampler2D& s, const vec4& c) { return texture(s, vec2(c.x / c.w, c.y / c.w)); }

// ---- per-invocation built-in variables ----
struct per_vertex { vec4 cur_position; };   // `gl_in[i].gl_Position` (gl_Position is a macro below)
struct invocation {
    uvec3 gl_GlobalInvocationID;
    vec4  gl_FragCoord;
    bool  discarded;
    // geometry stage: inputs, the current output vertex, and what EmitVertex() recorded
    per_vertex in_vertices[3];
    vec4 cur_position; int cur_viewport;
    int emitted; vec4 out_position[8]; int out_viewport[8];
    void (*emit_hook)(int vertex);      // driver callback: capture the shader's own per-vertex outputs at EmitVertex()
};
inline thread_local invocation g_inv;

// vkCmdDispatch(gx, gy, gz) of a shader with local size (lx, ly, lz): every invocation, sequentially, in
// gl_GlobalInvocationID order (x fastest). Sequential execution is one legal schedule of the dispatch.
template <class F> inline void dispatch(uint gx, uint gy, uint gz, uint lx, uint ly, uint lz, F&& shader_main, bool independent = false)
{
    if (independent) {
        // invocations that neither read what another one writes nor use atomics (the atlas passes): z planes across the
        // host threads - any schedule gives the same image, and the CPU baseline may use every core
#pragma omp parallel for schedule(static)
        for (long z = 0; z < (long)(gz * lz); ++z)
            for (uint y = 0; y < gy * ly; ++y)
                for (uint x = 0; x < gx * lx; ++x) {
                    g_inv.gl_GlobalInvocationID = uvec3(x, y, (uint)z);
                    shader_main();
                }
        return;
    }
    for (uint z = 0; z < gz * lz; ++z)
        for (uint y = 0; y < gy * ly; ++y)
            for (uint x = 0; x < gx * lx; ++x) {
                g_inv.gl_GlobalInvocationID = uvec3(x, y, z);
                shader_main();
            }
}

} // namespace glsl

// ---- driver support: the buffers OctreeBuilder binds to its six compute programs (OctreeBuilder.cpp:84-127) ----
struct ref_octree_state {
    glsl::uvec2*       nodes;          // "OctreeStorageBuffer"  (binding 1)
    const glsl::uvec2* fragments;      // "FragmentListStorageBuffer" (binding 2)
    unsigned counter;                  // Counter (binding 0)
    unsigned alloc_begin, alloc_num;   // _infoBuffer (binding 3), initialised { 0, 8 } (OctreeBuilder.cpp:96)
    unsigned indirect[3];              // _indirectBuffer (binding 4), initialised { 1, 1, 1 } (OctreeBuilder.cpp:107)
    unsigned fragment_num, voxel_resolution, target_level;   // push constants
};

// Names the shader text uses unqualified. Using-DECLARATIONS (not a using-directive) so that they hide the C
// library's ::exp2(double) etc. during unqualified lookup inside the shader's namespace.
#define GLSL_USING_BUILTINS \
    using glsl::uint; using glsl::vec2; using glsl::vec3; using glsl::vec4; using glsl::ivec2; using glsl::ivec3; using glsl::ivec4; \
    using glsl::uvec2; using glsl::uvec3; using glsl::uvec4; using glsl::bvec2; using glsl::bvec3; using glsl::bvec4; \
    using glsl::mat3; using glsl::mat4; using glsl::image3D; using glsl::uimage3D; using glsl::sampler2D; using glsl::sampler3D; \
    using glsl::radians; using glsl::sin; using glsl::cos; using glsl::tan; using glsl::pow; using glsl::exp; using glsl::log; \
    using glsl::exp2; using glsl::log2; using glsl::sqrt; using glsl::inversesqrt; using glsl::abs; using glsl::sign; \
    using glsl::floor; using glsl::ceil; using glsl::round; using glsl::fract; using glsl::mod; using glsl::min; using glsl::max; \
    using glsl::clamp; using glsl::mix; using glsl::step; using glsl::smoothstep; using glsl::length; \
    using glsl::imageLoad; using glsl::imageStore; using glsl::imageAtomicCompSwap; using glsl::atomicCompSwap; \
    using glsl::atomicAdd; using glsl::atomicOr; using glsl::barrier; using glsl::memoryBarrier; \
    using glsl::texture; using glsl::textureLod; using glsl::textureProj; using glsl::texelFetch; using glsl::textureSize;

inline void EmitVertex();
inline void EndPrimitive() {}
inline void EmitVertex()
{
    glsl::invocation& v = glsl::g_inv;
    if (v.emitted < 8) { v.out_position[v.emitted] = v.cur_position; v.out_viewport[v.emitted] = v.cur_viewport; }
    if (v.emit_hook) v.emit_hook(v.emitted);
    ++v.emitted;
}
#define gl_in (glsl::g_inv.in_vertices)
#define gl_Position (glsl::g_inv.cur_position)
#define gl_ViewportIndex (glsl::g_inv.cur_viewport)
#define gl_GlobalInvocationID (glsl::g_inv.gl_GlobalInvocationID)
#define gl_FragCoord (glsl::g_inv.gl_FragCoord)
#define discard do { glsl::g_inv.discarded = true; return; } while (0)
