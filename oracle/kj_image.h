// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).
// Typed access to kjb_image storage with the quantisation of the Vulkan format the reference
// creates each image with (image formats are part of the algorithm).
// Out-of-bounds loads return 0 and out-of-bounds stores are dropped (D3D/Vulkan robust image access),
// which the shaders rely on at screen edges (e.g. temporal_filter.hlsl:78 `input_tex[px + int2(x,y)]`).
#pragma once
#include "kj_math.h"
#include <vector>
#include <cstring>

namespace kjo {

inline uint32_t format_texel_bytes(uint32_t f) {
    switch (f) {
        case KJB_FMT_R32_FLOAT: case KJB_FMT_RG16_FLOAT: case KJB_FMT_RGBA8_UNORM: case KJB_FMT_RGBA8_SNORM:
        case KJB_FMT_A2R10G10B10_UNORM: case KJB_FMT_R11G11B10_UFLOAT: case KJB_FMT_R32_UINT: return 4;
        case KJB_FMT_RG32_UINT: case KJB_FMT_RGBA16_FLOAT: case KJB_FMT_RGBA16_SNORM: case KJB_FMT_RG32_FLOAT: return 8;
        case KJB_FMT_RGBA32_FLOAT: case KJB_FMT_RGBA32_UINT: return 16;
        case KJB_FMT_R8_UNORM: case KJB_FMT_R8_SNORM: return 1;
        case KJB_FMT_R16_FLOAT: return 2;
        default: return 0;
    }
}

inline int snorm_enc(float v, float scale) {   // round half away from zero (DESIGN.md "storage rounding")
    v = clamp(v, -1.0f, 1.0f) * scale;
    return v >= 0.0f ? int(v + 0.5f) : -int(-v + 0.5f);
}
inline uint unorm_enc(float v, float scale) { return uint(clamp(v, 0.0f, 1.0f) * scale + 0.5f); }

// B10G11R11_UFLOAT: truncating conversion (pack_unpack.hlsl:166-177 comment "GPU will convert ... by trimming")
inline uint f32_to_uf(float v, int mant_bits) {
    if (!(v > 0.0f)) return 0;
    uint h = kjb_f32_to_f16(v);          // note: RN to half first would double-round; do exact truncation instead
    (void)h;
    uint u = asuint(v);
    int e = int(u >> 23) - 127 + 15;
    uint m = (u & 0x7fffffu) >> (23 - mant_bits);
    if (e >= 31) return (30u << mant_bits) | ((1u << mant_bits) - 1u);   // clamp to max finite
    if (e <= 0) {                                                          // denormal
        if (e < -mant_bits) return 0;
        m = ((u & 0x7fffffu) | 0x800000u) >> (23 - mant_bits + 1 - e);
        return m;
    }
    return (uint(e) << mant_bits) | m;
}
inline float uf_to_f32(uint v, int mant_bits) {
    uint e = v >> mant_bits, m = v & ((1u << mant_bits) - 1u);
    if (e == 0) return float(m) * exp2(float(-14 - mant_bits));
    if (e == 31) return asfloat(0x7f800000u);
    return asfloat(((e + 112u) << 23) | (m << (23 - mant_bits)));
}

struct Img {
    kjb_image d;
    Img() { memset(&d, 0, sizeof(d)); }
    Img(const kjb_image& i) : d(i) {}
    int w() const { return int(d.width); }
    int h() const { return int(d.height); }
    bool inb(int x, int y) const { return x >= 0 && y >= 0 && x < int(d.width) && y < int(d.height); }
    uint8_t* at(int x, int y, int layer = 0) const {
        return (uint8_t*)d.data + (size_t(layer) * d.height * d.width + size_t(y) * d.width + x) * format_texel_bytes(d.format);
    }

    float4 load(int x, int y, int layer = 0) const {
        if (!inb(x, y) || !d.data) return float4(0.0f);
        const uint8_t* p = at(x, y, layer);
        switch (d.format) {
            case KJB_FMT_R32_FLOAT: { float v; memcpy(&v, p, 4); return float4(v, 0, 0, 1); }
            case KJB_FMT_RG32_FLOAT: { float v[2]; memcpy(v, p, 8); return float4(v[0], v[1], 0, 1); }
            case KJB_FMT_RGBA32_FLOAT: { float v[4]; memcpy(v, p, 16); return float4(v[0], v[1], v[2], v[3]); }
            case KJB_FMT_RGBA16_FLOAT: { uint16_t v[4]; memcpy(v, p, 8);
                return float4(kjb_f16_to_f32(v[0]), kjb_f16_to_f32(v[1]), kjb_f16_to_f32(v[2]), kjb_f16_to_f32(v[3])); }
            case KJB_FMT_RG16_FLOAT: { uint16_t v[2]; memcpy(v, p, 4); return float4(kjb_f16_to_f32(v[0]), kjb_f16_to_f32(v[1]), 0, 1); }
            case KJB_FMT_R16_FLOAT: { uint16_t v; memcpy(&v, p, 2); return float4(kjb_f16_to_f32(v), 0, 0, 1); }
            case KJB_FMT_RGBA8_UNORM: return float4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f);
            case KJB_FMT_RGBA8_SNORM: { const int8_t* s = (const int8_t*)p;
                return float4(max(s[0] / 127.0f, -1.0f), max(s[1] / 127.0f, -1.0f), max(s[2] / 127.0f, -1.0f), max(s[3] / 127.0f, -1.0f)); }
            case KJB_FMT_R8_UNORM: return float4(p[0] / 255.0f, 0, 0, 1);
            case KJB_FMT_R8_SNORM: return float4(max(((const int8_t*)p)[0] / 127.0f, -1.0f), 0, 0, 1);
            case KJB_FMT_RGBA16_SNORM: { int16_t v[4]; memcpy(v, p, 8);
                return float4(max(v[0] / 32767.0f, -1.0f), max(v[1] / 32767.0f, -1.0f), max(v[2] / 32767.0f, -1.0f), max(v[3] / 32767.0f, -1.0f)); }
            case KJB_FMT_A2R10G10B10_UNORM: { uint v; memcpy(&v, p, 4);
                return float4(((v >> 20) & 1023u) / 1023.0f, ((v >> 10) & 1023u) / 1023.0f, (v & 1023u) / 1023.0f, (v >> 30) / 3.0f); }
            case KJB_FMT_R11G11B10_UFLOAT: { uint v; memcpy(&v, p, 4);
                return float4(uf_to_f32(v & 2047u, 6), uf_to_f32((v >> 11) & 2047u, 6), uf_to_f32(v >> 22, 5), 1); }
            default: return float4(0.0f);
        }
    }
    uint4 load_u(int x, int y) const {
        if (!inb(x, y) || !d.data) return uint4(0, 0, 0, 0);
        const uint8_t* p = at(x, y);
        uint v[4] = {0, 0, 0, 0};
        memcpy(v, p, format_texel_bytes(d.format) < 16 ? format_texel_bytes(d.format) : 16);
        return uint4(v[0], v[1], v[2], v[3]);
    }
    void store(int x, int y, float4 c, int layer = 0) const {
        if (!inb(x, y) || !d.data) return;
        uint8_t* p = at(x, y, layer);
        switch (d.format) {
            case KJB_FMT_R32_FLOAT: memcpy(p, &c.x, 4); break;
            case KJB_FMT_RG32_FLOAT: { float v[2] = {c.x, c.y}; memcpy(p, v, 8); break; }
            case KJB_FMT_RGBA32_FLOAT: { float v[4] = {c.x, c.y, c.z, c.w}; memcpy(p, v, 16); break; }
            case KJB_FMT_RGBA16_FLOAT: { uint16_t v[4] = {(uint16_t)kjb_f32_to_f16(c.x), (uint16_t)kjb_f32_to_f16(c.y), (uint16_t)kjb_f32_to_f16(c.z), (uint16_t)kjb_f32_to_f16(c.w)}; memcpy(p, v, 8); break; }
            case KJB_FMT_RG16_FLOAT: { uint16_t v[2] = {(uint16_t)kjb_f32_to_f16(c.x), (uint16_t)kjb_f32_to_f16(c.y)}; memcpy(p, v, 4); break; }
            case KJB_FMT_R16_FLOAT: { uint16_t v = (uint16_t)kjb_f32_to_f16(c.x); memcpy(p, &v, 2); break; }
            case KJB_FMT_RGBA8_UNORM: p[0] = (uint8_t)unorm_enc(c.x, 255.0f); p[1] = (uint8_t)unorm_enc(c.y, 255.0f); p[2] = (uint8_t)unorm_enc(c.z, 255.0f); p[3] = (uint8_t)unorm_enc(c.w, 255.0f); break;
            case KJB_FMT_RGBA8_SNORM: { int8_t* s = (int8_t*)p; s[0] = (int8_t)snorm_enc(c.x, 127.0f); s[1] = (int8_t)snorm_enc(c.y, 127.0f); s[2] = (int8_t)snorm_enc(c.z, 127.0f); s[3] = (int8_t)snorm_enc(c.w, 127.0f); break; }
            case KJB_FMT_R8_UNORM: p[0] = (uint8_t)unorm_enc(c.x, 255.0f); break;
            case KJB_FMT_R8_SNORM: ((int8_t*)p)[0] = (int8_t)snorm_enc(c.x, 127.0f); break;
            case KJB_FMT_RGBA16_SNORM: { int16_t v[4] = {(int16_t)snorm_enc(c.x, 32767.0f), (int16_t)snorm_enc(c.y, 32767.0f), (int16_t)snorm_enc(c.z, 32767.0f), (int16_t)snorm_enc(c.w, 32767.0f)}; memcpy(p, v, 8); break; }
            case KJB_FMT_A2R10G10B10_UNORM: { uint v = (unorm_enc(c.w, 3.0f) << 30) | (unorm_enc(c.x, 1023.0f) << 20) | (unorm_enc(c.y, 1023.0f) << 10) | unorm_enc(c.z, 1023.0f); memcpy(p, &v, 4); break; }
            case KJB_FMT_R11G11B10_UFLOAT: { uint v = f32_to_uf(c.x, 6) | (f32_to_uf(c.y, 6) << 11) | (f32_to_uf(c.z, 5) << 22); memcpy(p, &v, 4); break; }
            default: break;
        }
    }
    void store_u(int x, int y, uint4 c) const {
        if (!inb(x, y) || !d.data) return;
        uint v[4] = {c.x, c.y, c.z, c.w};
        memcpy(at(x, y), v, format_texel_bytes(d.format) < 16 ? format_texel_bytes(d.format) : 16);
    }
    float4 load(int2 p) const { return load(p.x, p.y); }
    uint4 load_u(int2 p) const { return load_u(p.x, p.y); }
    void store(int2 p, float4 c) const { store(p.x, p.y, c); }

    // SampleLevel(sampler_nnc, uv, 0): nearest, clamp-to-edge
    float4 sample_nearest_clamp(float2 uv) const {
        int x = int(floor(uv.x * float(w()))), y = int(floor(uv.y * float(h())));
        x = x < 0 ? 0 : (x >= w() ? w() - 1 : x); y = y < 0 ? 0 : (y >= h() ? h() - 1 : y);
        return load(x, y);
    }
    // SampleLevel(sampler_lnc, uv, 0): bilinear, clamp-to-edge
    float4 sample_bilinear_clamp(float2 uv, int layer = 0) const {
        float fx = uv.x * float(w()) - 0.5f, fy = uv.y * float(h()) - 0.5f;
        float x0f = floor(fx), y0f = floor(fy);
        float tx = fx - x0f, ty = fy - y0f;
        int x0 = int(x0f), y0 = int(y0f), x1 = x0 + 1, y1 = y0 + 1;
        auto cl = [](int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); };
        x0 = cl(x0, w()); x1 = cl(x1, w()); y0 = cl(y0, h()); y1 = cl(y1, h());
        float4 a = load(x0, y0, layer), b = load(x1, y0, layer), c = load(x0, y1, layer), dd = load(x1, y1, layer);
        float4 top = lerp(a, b, tx), bot = lerp(c, dd, tx);   // fma(b - a, t, a) per component
        return lerp(top, bot, ty);
    }
    // TextureCube.SampleLevel(sampler_llr, dir, 0): face select (Vulkan spec table), bilinear inside the face,
    // clamp at face edges (no seamless filtering: documented deviation).
    float4 sample_cube(float3 dir) const {
        float ax = abs(dir.x), ay = abs(dir.y), az = abs(dir.z);
        int face; float sc, tc, ma;
        if (ax >= ay && ax >= az) { ma = ax; if (dir.x >= 0) { face = 0; sc = -dir.z; tc = -dir.y; } else { face = 1; sc = dir.z; tc = -dir.y; } }
        else if (ay >= az) { ma = ay; if (dir.y >= 0) { face = 2; sc = dir.x; tc = dir.z; } else { face = 3; sc = dir.x; tc = -dir.z; } }
        else { ma = az; if (dir.z >= 0) { face = 4; sc = dir.x; tc = -dir.y; } else { face = 5; sc = -dir.x; tc = -dir.y; } }
        float2 uv(0.5f * (sc / ma + 1.0f), 0.5f * (tc / ma + 1.0f));
        return sample_bilinear_clamp(uv, face);
    }
};

}  // namespace kjo
