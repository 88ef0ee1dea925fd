// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of kajiya's shader numeric core.
// Nothing under kajiya_b200/ may include, link or call this file (see DESIGN.md "oracle").
// PARITY UNPINNED: the reference ships no golden vectors for this path;
// this restatement is pinned only by self-made known-answer tests (tests/test_oracle_kat.py).
//
// Follows (paths relative to /root/reference/assets/shaders/inc/):
//   hash.hlsl, math.hlsl, math_const.hlsl, uv.hlsl, pack_unpack.hlsl, reservoir.hlsl, gbuffer.hlsl,
//   quasi_random.hlsl, color/srgb.hlsl, color/ycbcr.hlsl, working_color_space.hlsl, bilinear.hlsl, sh.hlsl
// Transcendentals come from include/kjb_numeric.h (the ABI's numeric contract).
#pragma once
#include "../include/kjb_numeric.h"
#include "../include/kjb.h"
#include <cstdint>
#include <cmath>

namespace kjo {

typedef uint32_t uint;

// ---------------------------------------------------------------- vectors (HLSL-like)
struct float2 { float x, y; float2() : x(0), y(0) {} float2(float a) : x(a), y(a) {} float2(float a, float b) : x(a), y(b) {} };
struct float3 { float x, y, z; float3() : x(0), y(0), z(0) {} float3(float a) : x(a), y(a), z(a) {} float3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
struct float4 { float x, y, z, w; float4() : x(0), y(0), z(0), w(0) {} float4(float a) : x(a), y(a), z(a), w(a) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} float4(float3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float4(float2 a, float2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    float3 xyz() const { return float3(x, y, z); } float2 xy() const { return float2(x, y); } };
struct int2 { int x, y; int2() : x(0), y(0) {} int2(int a) : x(a), y(a) {} int2(int a, int b) : x(a), y(b) {} };
struct uint2 { uint x, y; uint2() : x(0), y(0) {} uint2(uint a) : x(a), y(a) {} uint2(uint a, uint b) : x(a), y(b) {} };
struct uint4 { uint x, y, z, w; uint4() : x(0), y(0), z(0), w(0) {} uint4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {} };

#define KJO_V2(op) \
    inline float2 operator op(float2 a, float2 b) { return float2(a.x op b.x, a.y op b.y); } \
    inline float2 operator op(float2 a, float b) { return float2(a.x op b, a.y op b); } \
    inline float2 operator op(float a, float2 b) { return float2(a op b.x, a op b.y); }
#define KJO_V3(op) \
    inline float3 operator op(float3 a, float3 b) { return float3(a.x op b.x, a.y op b.y, a.z op b.z); } \
    inline float3 operator op(float3 a, float b) { return float3(a.x op b, a.y op b, a.z op b); } \
    inline float3 operator op(float a, float3 b) { return float3(a op b.x, a op b.y, a op b.z); }
#define KJO_V4(op) \
    inline float4 operator op(float4 a, float4 b) { return float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    inline float4 operator op(float4 a, float b) { return float4(a.x op b, a.y op b, a.z op b, a.w op b); } \
    inline float4 operator op(float a, float4 b) { return float4(a op b.x, a op b.y, a op b.z, a op b.w); }
KJO_V2(+) KJO_V2(-) KJO_V2(*) KJO_V2(/) KJO_V3(+) KJO_V3(-) KJO_V3(*) KJO_V3(/) KJO_V4(+) KJO_V4(-) KJO_V4(*) KJO_V4(/)
inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float b) { a = a * b; return a; }
inline float3& operator/=(float3& a, float b) { a = a / b; return a; }
inline float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
inline float2& operator+=(float2& a, float2 b) { a = a + b; return a; }
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator-(int2 a, int2 b) { return int2(a.x - b.x, a.y - b.y); }
inline int2 operator*(int2 a, int b) { return int2(a.x * b, a.y * b); }
inline bool operator==(int2 a, int2 b) { return a.x == b.x && a.y == b.y; }

inline float min(float a, float b) { return kjb_min(a, b); }
inline float max(float a, float b) { return kjb_max(a, b); }
inline float abs(float a) { return kjb_abs(a); }
inline float sqrt(float a) { return kjb_sqrt(a); }
inline float rsqrt(float a) { return kjb_rsqrt(a); }
inline float rcp(float a) { return kjb_rcp(a); }
inline float floor(float a) { return kjb_floor(a); }
inline float frac(float a) { return kjb_frac(a); }
inline float saturate(float a) { return kjb_saturate(a); }
inline float clamp(float x, float a, float b) { return kjb_clamp(x, a, b); }
inline float lerp(float a, float b, float t) { return kjb_lerp(a, b, t); }
inline float step(float e, float x) { return kjb_step(e, x); }
inline float smoothstep(float a, float b, float x) { return kjb_smoothstep(a, b, x); }
inline float sin(float a) { return kjb_sin(a); }
inline float cos(float a) { return kjb_cos(a); }
inline float exp2(float a) { return kjb_exp2(a); }
inline float log2(float a) { return kjb_log2(a); }
inline float exp(float a) { return kjb_exp(a); }
inline float log(float a) { return kjb_log(a); }
inline float pow(float a, float b) { return kjb_pow(a, b); }
inline float atan(float a) { return kjb_atan(a); }
inline float atan2(float a, float b) { return kjb_atan2(a, b); }
inline float sign(float a) { return kjb_sign(a); }
inline float asfloat(uint u) { return kjb_u2f(u); }
inline uint asuint(float f) { return kjb_f2u(f); }

inline float2 min(float2 a, float2 b) { return float2(min(a.x, b.x), min(a.y, b.y)); }
inline float2 max(float2 a, float2 b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float4 min(float4 a, float4 b) { return float4(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
inline float4 max(float4 a, float4 b) { return float4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
inline float3 abs(float3 a) { return float3(abs(a.x), abs(a.y), abs(a.z)); }
inline float2 abs(float2 a) { return float2(abs(a.x), abs(a.y)); }
inline float4 abs(float4 a) { return float4(abs(a.x), abs(a.y), abs(a.z), abs(a.w)); }
inline float3 sqrt(float3 a) { return float3(sqrt(a.x), sqrt(a.y), sqrt(a.z)); }
inline float4 sqrt(float4 a) { return float4(sqrt(a.x), sqrt(a.y), sqrt(a.z), sqrt(a.w)); }
inline float2 floor(float2 a) { return float2(floor(a.x), floor(a.y)); }
inline float2 frac(float2 a) { return float2(frac(a.x), frac(a.y)); }
inline float3 exp(float3 a) { return float3(exp(a.x), exp(a.y), exp(a.z)); }
inline float3 clamp(float3 v, float3 a, float3 b) { return min(max(v, a), b); }
inline float2 clamp(float2 v, float2 a, float2 b) { return min(max(v, a), b); }
// lerp = fma(b - a, t, a) per component (numeric contract: explicit FMA, kjb_numeric.h)
inline float3 lerp(float3 a, float3 b, float t) { return float3(kjb_lerp(a.x, b.x, t), kjb_lerp(a.y, b.y, t), kjb_lerp(a.z, b.z, t)); }
inline float3 lerp(float3 a, float3 b, float3 t) { return float3(kjb_lerp(a.x, b.x, t.x), kjb_lerp(a.y, b.y, t.y), kjb_lerp(a.z, b.z, t.z)); }
inline float4 lerp(float4 a, float4 b, float t) { return float4(kjb_lerp(a.x, b.x, t), kjb_lerp(a.y, b.y, t), kjb_lerp(a.z, b.z, t), kjb_lerp(a.w, b.w, t)); }
inline float2 lerp(float2 a, float2 b, float t) { return float2(kjb_lerp(a.x, b.x, t), kjb_lerp(a.y, b.y, t)); }
inline float2 saturate(float2 a) { return float2(saturate(a.x), saturate(a.y)); }
inline float3 saturate(float3 a) { return float3(saturate(a.x), saturate(a.y), saturate(a.z)); }
// dot products: left-to-right FMA chains (the evaluation order both sides of the parity agree on)
// HLSL mad(): a * s + c with one rounding per component (the filters' weighted-sum taps; mirrors kjb_device.cuh)
inline float  mad(float a, float s, float c) { return kjb_fma(a, s, c); }
inline float2 mad(float2 a, float s, float2 c) { return float2(kjb_fma(a.x, s, c.x), kjb_fma(a.y, s, c.y)); }
inline float3 mad(float3 a, float s, float3 c) { return float3(kjb_fma(a.x, s, c.x), kjb_fma(a.y, s, c.y), kjb_fma(a.z, s, c.z)); }
inline float4 mad(float4 a, float s, float4 c) { return float4(kjb_fma(a.x, s, c.x), kjb_fma(a.y, s, c.y), kjb_fma(a.z, s, c.z), kjb_fma(a.w, s, c.w)); }
inline float dot(float2 a, float2 b) { return kjb_fma(a.y, b.y, a.x * b.x); }
inline float dot(float3 a, float3 b) { return kjb_fma(a.z, b.z, kjb_fma(a.y, b.y, a.x * b.x)); }
inline float dot(float4 a, float4 b) { return kjb_fma(a.w, b.w, kjb_fma(a.z, b.z, kjb_fma(a.y, b.y, a.x * b.x))); }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(float2 a) { return sqrt(dot(a, a)); }
inline float length(float3 a) { return sqrt(dot(a, a)); }
inline float3 normalize(float3 a) { return a * rsqrt(dot(a, a)); }
inline float3 reflect(float3 i, float3 n) { return i - 2.0f * dot(n, i) * n; }
inline bool any_nonzero(float3 a) { return a.x != 0.0f || a.y != 0.0f || a.z != 0.0f; }

// ---------------------------------------------------------------- matrices
// glam column-major Mat4 as uploaded by the reference; HLSL `mul(M, v)`.
inline float4 mul(const kjb_mat4& M, float4 v) {
    const float* m = M.m;
    return float4(
        kjb_fma(m[12], v.w, kjb_fma(m[8], v.z, kjb_fma(m[4], v.y, m[0] * v.x))),
        kjb_fma(m[13], v.w, kjb_fma(m[9], v.z, kjb_fma(m[5], v.y, m[1] * v.x))),
        kjb_fma(m[14], v.w, kjb_fma(m[10], v.z, kjb_fma(m[6], v.y, m[2] * v.x))),
        kjb_fma(m[15], v.w, kjb_fma(m[11], v.z, kjb_fma(m[7], v.y, m[3] * v.x))));
}
// float3x3 stored as rows (HLSL float3x3(r0, r1, r2) constructor order)
struct float3x3 { float3 r0, r1, r2; };
inline float3 mul(const float3x3& M, float3 v) { return float3(dot(M.r0, v), dot(M.r1, v), dot(M.r2, v)); }
// mul(v, M): row vector times matrix
inline float3 mul(float3 v, const float3x3& M) {
    return float3(kjb_fma(v.z, M.r2.x, kjb_fma(v.y, M.r1.x, v.x * M.r0.x)),
                  kjb_fma(v.z, M.r2.y, kjb_fma(v.y, M.r1.y, v.x * M.r0.y)),
                  kjb_fma(v.z, M.r2.z, kjb_fma(v.y, M.r1.z, v.x * M.r0.z)));
}

// ---------------------------------------------------------------- math_const.hlsl
static const float M_PI_F = 3.14159265358979323846f;
static const float M_TAU_F = 6.28318530717958647692f;
static const float M_FRAC_1_PI_F = 0.318309886183790671537767526745028724f;
static const float M_PLASTIC_F = 1.32471795724474602596f;
static const float GOLDEN_ANGLE = 2.39996323f;       // math_const.hlsl:31
static const float FLT_MAX_F = 3.402823466e+38f;

// ---------------------------------------------------------------- hash.hlsl:7-55 (all u32: bit-exact)
inline uint hash1(uint x) { x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u); return x; }
inline uint hash1_mut(uint& h) { uint res = h; h = hash1(h); return res; }
inline uint hash_combine2(uint x, uint y) {
    const uint M = 1664525u, C = 1013904223u;
    uint seed = (x * M + y + C) * M;
    seed ^= (seed >> 11u); seed ^= (seed << 7u) & 0x9d2c5680u; seed ^= (seed << 15u) & 0xefc60000u; seed ^= (seed >> 18u);
    return seed;
}
inline uint hash2(uint2 v) { return hash_combine2(v.x, hash1(v.y)); }
inline uint hash3(uint x, uint y, uint z) { return hash_combine2(x, hash2(uint2(y, z))); }
inline uint hash4(uint x, uint y, uint z, uint w) { return hash_combine2(x, hash3(y, z, w)); }
inline float uint_to_u01_float(uint h) { h &= 0x007FFFFFu; h |= 0x3F800000u; return asfloat(h) - 1.0f; }
inline float interleaved_gradient_noise(uint2 px) {   // hash.hlsl:57-59
    return frac(52.9829189f * frac(0.06711056f * float(px.x) + 0.00583715f * float(px.y)));
}

// ---------------------------------------------------------------- quasi_random.hlsl
inline float radical_inverse_vdc(uint bits) {
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return float(bits) * 2.3283064365386963e-10f;
}
inline float2 hammersley(uint i, uint n) { return float2(float(i + 1) / float(n), radical_inverse_vdc(i + 1)); }
inline float2 r2_sequence(uint i) {                   // quasi_random.hlsl:19-24
    const float a1 = 1.0f / M_PLASTIC_F;
    const float a2 = 1.0f / (M_PLASTIC_F * M_PLASTIC_F);
    return frac(float2(a1, a2) * float(i) + 0.5f);
}

// ---------------------------------------------------------------- math.hlsl
inline float max3(float x, float y, float z) { return max(x, max(y, z)); }
inline float square(float x) { return x * x; }
inline float3x3 build_orthonormal_basis(float3 n) {   // math.hlsl:21-43
    float3 b1, b2;
    if (n.z < 0.0f) {
        const float a = 1.0f / (1.0f - n.z);
        const float b = n.x * n.y * a;
        b1 = float3(1.0f - n.x * n.x * a, -b, n.x);
        b2 = float3(b, n.y * n.y * a - 1.0f, -n.y);
    } else {
        const float a = 1.0f / (1.0f + n.z);
        const float b = -n.x * n.y * a;
        b1 = float3(1.0f - n.x * n.x * a, b, -n.x);
        b2 = float3(b, 1.0f - n.y * n.y * a, -n.y);
    }
    float3x3 m; m.r0 = float3(b1.x, b2.x, n.x); m.r1 = float3(b1.y, b2.y, n.y); m.r2 = float3(b1.z, b2.z, n.z);
    return m;
}
inline float3 uniform_sample_cone(float2 urand, float cos_theta_max) {   // math.hlsl:45-50
    float cos_theta = (1.0f - urand.x) + urand.x * cos_theta_max;
    float sin_theta = sqrt(saturate(1.0f - cos_theta * cos_theta));
    float phi = urand.y * M_TAU_F;
    return float3(sin_theta * cos(phi), sin_theta * sin(phi), cos_theta);
}
inline float inverse_depth_relative_diff(float primary_depth, float secondary_depth) {   // math.hlsl:65-67
    return abs(max(1e-20f, primary_depth) / max(1e-20f, secondary_depth) - 1.0f);
}
inline float exponential_squish(float len, float squish_scale) { return exp2(-clamp(squish_scale * len, 0.0f, 100.0f)); }
inline float exponential_unsquish(float len, float squish_scale) { return max(0.0f, -1.0f / squish_scale * log2(1e-30f + len)); }
inline float3 uniform_sample_hemisphere(float2 urand) {                  // math.hlsl:78-83
    float phi = urand.y * M_TAU_F;
    float cos_theta = 1.0f - urand.x;
    float sin_theta = sqrt(1.0f - cos_theta * cos_theta);
    return float3(cos(phi) * sin_theta, sin(phi) * sin_theta, cos_theta);
}
inline float3 uniform_sample_sphere(float2 urand) {                      // math.hlsl:85-91
    float z = 1.0f - 2.0f * urand.x;
    float xy = sqrt(max(0.0f, 1.0f - z * z));
    float sn = sin(M_TAU_F * urand.y);
    float cs = cos(M_TAU_F * urand.y);
    return float3(cs * xy, sn * xy, z);
}

// ---------------------------------------------------------------- uv.hlsl
inline float2 get_uv(int2 pix, float4 texSize) { return (float2(float(pix.x), float(pix.y)) + 0.5f) * float2(texSize.z, texSize.w); }
inline float2 get_uv(float2 pix, float4 texSize) { return (pix + 0.5f) * float2(texSize.z, texSize.w); }
inline float2 cs_to_uv(float2 cs) { return cs * float2(0.5f, -0.5f) + float2(0.5f, 0.5f); }
inline float2 uv_to_cs(float2 uv) { return (uv - float2(0.5f)) * float2(2.0f, -2.0f); }

// ---------------------------------------------------------------- pack_unpack.hlsl
inline float unpack_unorm(uint pckd, uint bitCount) { uint maxVal = (1u << bitCount) - 1; return float(pckd & maxVal) / float(maxVal); }
inline uint pack_unorm(float val, uint bitCount) { uint maxVal = (1u << bitCount) - 1; return uint(clamp(val, 0.0f, 1.0f) * float(maxVal) + 0.5f); }
inline float pack_normal_11_10_11(float3 n) {          // pack_unpack.hlsl:14-20
    uint pckd = 0;
    pckd += pack_unorm(n.x * 0.5f + 0.5f, 11);
    pckd += pack_unorm(n.y * 0.5f + 0.5f, 10) << 11;
    pckd += pack_unorm(n.z * 0.5f + 0.5f, 11) << 21;
    return asfloat(pckd);
}
inline float3 unpack_normal_11_10_11_no_normalize(float pckd) {
    uint p = asuint(pckd);
    return float3(unpack_unorm(p, 11), unpack_unorm(p >> 11, 10), unpack_unorm(p >> 21, 11)) * 2.0f - 1.0f;
}
inline float3 unpack_normal_11_10_11(float pckd) { return normalize(unpack_normal_11_10_11_no_normalize(pckd)); }
inline uint pack_color_888(float3 color) {
    color = sqrt(color);
    uint pckd = 0;
    pckd += pack_unorm(color.x, 8); pckd += pack_unorm(color.y, 8) << 8; pckd += pack_unorm(color.z, 8) << 16;
    return pckd;
}
inline float3 unpack_color_888(uint p) {
    float3 color = float3(unpack_unorm(p, 8), unpack_unorm(p >> 8, 8), unpack_unorm(p >> 16, 8));
    return color * color;
}
inline uint pack_2x16f_uint(float2 f) { return kjb_f32_to_f16(f.x) | (kjb_f32_to_f16(f.y) << 16u); }
inline float2 unpack_2x16f_uint(uint u) { return float2(kjb_f16_to_f32(u & 0xffff), kjb_f16_to_f32((u >> 16) & 0xffff)); }
// octahedral (pack_unpack.hlsl:66-88)
inline float2 octa_wrap(float2 v) { return (1.0f - abs(float2(v.y, v.x))) * (float2(step(0.0f, v.x), step(0.0f, v.y)) * 2.0f - 1.0f); }
inline float2 octa_encode(float3 n) {
    n = n / (abs(n.x) + abs(n.y) + abs(n.z));
    float2 nxy(n.x, n.y);
    if (n.z < 0.0f) nxy = octa_wrap(nxy);
    return nxy * 0.5f + 0.5f;
}
inline float3 octa_decode(float2 f) {
    f = f * 2.0f - 1.0f;
    float3 n = float3(f.x, f.y, 1.0f - abs(f.x) - abs(f.y));
    float t = clamp(-n.z, 0.0f, 1.0f);
    n.x -= (step(0.0f, n.x) * 2 - 1) * t;
    n.y -= (step(0.0f, n.y) * 2 - 1) * t;
    return normalize(n);
}
// RGB9E5 (pack_unpack.hlsl:102-164)
inline int floor_log2(float x) { uint f = asuint(x); uint be = (f & 0x7F800000u) >> 23; return int(be) - 127; }
inline uint float3_to_rgb9e5(float3 rgb) {
    const float MAX_RGB9E5 = (511.0f / 512.0f) * 65536.0f;
    float rc = clamp(rgb.x, 0.0f, MAX_RGB9E5), gc = clamp(rgb.y, 0.0f, MAX_RGB9E5), bc = clamp(rgb.z, 0.0f, MAX_RGB9E5);
    float maxrgb = max(rc, max(gc, bc));
    int fl = floor_log2(maxrgb);
    int exp_shared = (fl > -16 ? fl : -16) + 1 + 15;
    float denom = exp2(float(exp_shared - 15 - 9));
    int maxm = int(floor(maxrgb / denom + 0.5f));
    if (maxm == 512) { denom *= 2; exp_shared += 1; }
    int rm = int(floor(rc / denom + 0.5f)), gm = int(floor(gc / denom + 0.5f)), bm = int(floor(bc / denom + 0.5f));
    return (uint(rm) << 23) | (uint(gm) << 14) | (uint(bm) << 5) | uint(exp_shared);
}
inline float3 rgb9e5_to_float3(uint v) {
    int exponent = int(v & 31u) - 15 - 9;
    float scale = exp2(float(exponent));
    return float3(float((v >> 23) & 511u) * scale, float((v >> 14) & 511u) * scale, float((v >> 5) & 511u) * scale);
}

// ---------------------------------------------------------------- color
inline float sRGB_to_luminance(float3 col) { return dot(col, float3(0.2126f, 0.7152f, 0.0722f)); }   // color/srgb.hlsl:4-6
inline float3 sRGB_to_YCbCr(float3 col) {             // color/ycbcr.hlsl:4-6
    return float3(dot(float3(0.2126f, 0.7152f, 0.0722f), col), dot(float3(-0.1146f, -0.3854f, 0.5f), col), dot(float3(0.5f, -0.4542f, -0.0458f), col));
}
inline float3 YCbCr_to_sRGB(float3 col) {             // color/ycbcr.hlsl:8-10
    return max(float3(0.0f), float3(dot(float3(1.0f, 0.0f, 1.5748f), col), dot(float3(1.0f, -0.1873f, -.4681f), col), dot(float3(1.0f, 1.8556f, 0.0f), col)));
}
inline float4 linear_rgb_to_crunched_luma_chroma(float4 v) {   // working_color_space.hlsl:9-13
    float3 c = sRGB_to_YCbCr(v.xyz());
    float k = sqrt(c.x) / max(1e-8f, c.x);
    return float4(c * k, v.w);
}
inline float4 crunched_luma_chroma_to_linear_rgb(float4 v) {   // working_color_space.hlsl:14-18
    float3 c = v.xyz() * v.x;
    c = YCbCr_to_sRGB(c);
    return float4(c, v.w);
}

// ---------------------------------------------------------------- gbuffer.hlsl
struct GbufferData { float3 albedo, emissive, normal; float roughness, metalness;
    GbufferData() : roughness(0), metalness(0) {} };
inline float roughness_to_perceptual_roughness(float r) { return sqrt(r); }
inline float perceptual_roughness_to_roughness(float r) { return r * r; }
inline uint4 gbuffer_pack(const GbufferData& g) {      // gbuffer.hlsl:51-63
    uint4 r;
    r.x = pack_color_888(g.albedo);
    r.y = asuint(pack_normal_11_10_11(g.normal));
    r.z = pack_2x16f_uint(float2(roughness_to_perceptual_roughness(g.roughness), g.metalness));
    r.w = float3_to_rgb9e5(g.emissive);
    return r;
}
inline GbufferData gbuffer_unpack(uint4 d) {           // gbuffer.hlsl:65-76
    GbufferData res;
    res.albedo = unpack_color_888(d.x);
    res.normal = unpack_normal_11_10_11(asfloat(d.y));
    float2 rm = unpack_2x16f_uint(d.z);
    res.roughness = perceptual_roughness_to_roughness(rm.x);
    res.metalness = rm.y;
    res.emissive = rgb9e5_to_float3(d.w);
    return res;
}

// ---------------------------------------------------------------- reservoir.hlsl:6-98
struct Reservoir1sppStreamState { float p_q_sel = 0, M_sum = 0; };
struct Reservoir1spp {
    float w_sum = 0; uint payload = 0; float M = 0, W = 0;
    static Reservoir1spp from_raw(uint2 raw) {
        Reservoir1spp res; res.payload = raw.x;
        float2 MW = unpack_2x16f_uint(raw.y); res.M = MW.x; res.W = MW.y; return res;
    }
    uint2 as_raw() const { return uint2(payload, pack_2x16f_uint(float2(M, max(0.0f, W)))); }
    bool update(float w, uint sample_payload, uint& rng) {
        w_sum += w; M += 1;
        const float dart = uint_to_u01_float(hash1_mut(rng));
        const float prob = w / w_sum;
        if (prob >= dart) { payload = sample_payload; return true; }
        return false;
    }
    bool update_with_stream(Reservoir1spp r, float p_q, float weight, Reservoir1sppStreamState& st, uint sample_payload, uint& rng) {
        st.M_sum += r.M;
        if (update(p_q * weight * r.W * r.M, sample_payload, rng)) { st.p_q_sel = p_q; return true; }
        return false;
    }
    void init_with_stream(float p_q, float weight, Reservoir1sppStreamState& st, uint sample_payload) {
        payload = sample_payload; w_sum = p_q * weight; M = (weight != 0) ? 1.0f : 0.0f; W = weight;
        st.p_q_sel = p_q; st.M_sum = M;
    }
    void finish_stream(const Reservoir1sppStreamState& st) { M = st.M_sum; W = w_sum / (max(1e-8f, M * st.p_q_sel)); }
};

// ---------------------------------------------------------------- bilinear.hlsl
struct Bilinear { float2 origin, weights; };
inline Bilinear get_bilinear_filter(float2 uv, float2 tex_size) {
    Bilinear r;
    float2 p = uv * tex_size - 0.5f;
    r.origin = float2(kjb_trunc(p.x), kjb_trunc(p.y));
    r.weights = frac(p);
    return r;
}
inline float4 get_bilinear_custom_weights(Bilinear f, float4 cw) {
    float4 w;
    w.x = (1.0f - f.weights.x) * (1.0f - f.weights.y);
    w.y = f.weights.x * (1.0f - f.weights.y);
    w.z = (1.0f - f.weights.x) * f.weights.y;
    w.w = f.weights.x * f.weights.y;
    return w * cw;
}
inline float4 apply_bilinear_custom_weights(float4 s00, float4 s10, float4 s01, float4 s11, float4 w) {
    float4 r = s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w;
    return r * rcp(dot(w, float4(1.0f)));
}

// ---------------------------------------------------------------- sh.hlsl
inline float4 sh_eval(float3 dir) {
    return float4(0.28209479177387814347403972578039f, -0.48860251190291992158638462283836f * dir.y,
                  0.48860251190291992158638462283836f * dir.z, -0.48860251190291992158638462283836f * dir.x);
}

}  // namespace kjo
