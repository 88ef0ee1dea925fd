// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).
// Restates (relative to /root/reference/assets/shaders/inc/):
//   frame_constants.hlsl:92-224 (ViewRayContext & transforms), brdf.hlsl, brdf_lut.hlsl, layered_brdf.hlsl,
//   sun.hlsl, atmosphere.hlsl, atmosphere_felix.hlsl, blue_noise.hlsl:8-15, lights/triangle.hlsl, ray_cone.hlsl
#pragma once
#include "kj_image.h"

namespace kjo {

// Global per-"context" state standing in for descriptor sets 1/2 (bindless LUTs, FrameConstants, lights).
struct Globals {
    kjb_frame_constants fc;
    std::vector<kjb_triangle_light> lights;
    Img brdf_fg_lut;      // 64x64 RGBA16F
    Img blue_noise;       // 256x256 RGBA8 (LDR_RGBA_0)
};

// ---------------------------------------------------------------- frame_constants.hlsl
struct ViewRayContext {
    float4 ray_dir_cs, ray_dir_vs_h, ray_dir_ws_h;
    float4 ray_origin_cs, ray_origin_vs_h, ray_origin_ws_h;
    float4 ray_hit_cs, ray_hit_vs_h, ray_hit_ws_h;

    float3 ray_dir_vs() const { return normalize(ray_dir_vs_h.xyz()); }
    float3 ray_dir_ws() const { return normalize(ray_dir_ws_h.xyz()); }
    float3 ray_origin_vs() const { return ray_origin_vs_h.xyz() / ray_origin_vs_h.w; }
    float3 ray_origin_ws() const { return ray_origin_ws_h.xyz() / ray_origin_ws_h.w; }
    float3 ray_hit_vs() const { return ray_hit_vs_h.xyz() / ray_hit_vs_h.w; }
    float3 ray_hit_ws() const { return ray_hit_ws_h.xyz() / ray_hit_ws_h.w; }

    float3 biased_secondary_ray_origin_ws() const { return ray_hit_ws() - ray_dir_ws() * (length(ray_hit_vs()) + length(ray_hit_ws())) * 1e-4f; }   // frame_constants.hlsl:133-135
    float3 biased_secondary_ray_origin_ws_with_normal(float3 normal) const {   // frame_constants.hlsl:140-144
        float3 ws_abs = abs(ray_hit_ws());
        float max_comp = max(max(ws_abs.x, ws_abs.y), max(ws_abs.z, -ray_hit_vs().z));
        return ray_hit_ws() + (normal - ray_dir_ws()) * max(1e-4f, max_comp * 1e-6f);
    }
    static ViewRayContext from_uv(const kjb_view_constants& vc, float2 uv) {    // :146-159
        ViewRayContext res;
        float2 cs = uv_to_cs(uv);
        res.ray_dir_cs = float4(cs.x, cs.y, 0.0f, 1.0f);
        res.ray_dir_vs_h = mul(vc.sample_to_view, res.ray_dir_cs);
        res.ray_dir_ws_h = mul(vc.view_to_world, res.ray_dir_vs_h);
        res.ray_origin_cs = float4(cs.x, cs.y, 1.0f, 1.0f);
        res.ray_origin_vs_h = mul(vc.sample_to_view, res.ray_origin_cs);
        res.ray_origin_ws_h = mul(vc.view_to_world, res.ray_origin_vs_h);
        return res;
    }
    static ViewRayContext from_uv_and_depth(const kjb_view_constants& vc, float2 uv, float depth) {   // :161-178
        ViewRayContext res = from_uv(vc, uv);
        float2 cs = uv_to_cs(uv);
        res.ray_hit_cs = float4(cs.x, cs.y, depth, 1.0f);
        res.ray_hit_vs_h = mul(vc.sample_to_view, res.ray_hit_cs);
        res.ray_hit_ws_h = mul(vc.view_to_world, res.ray_hit_vs_h);
        return res;
    }
    static ViewRayContext from_uv_and_biased_depth(const kjb_view_constants& vc, float2 uv, float depth) {   // :180-182
        return from_uv_and_depth(vc, uv, min(1.0f, depth * asfloat(0x3f800040u)));
    }
};

inline float3 get_eye_position(const kjb_view_constants& vc) {
    float4 e = mul(vc.view_to_world, float4(0, 0, 0, 1));
    return e.xyz() / e.w;
}
inline float3 direction_view_to_world(const kjb_view_constants& vc, float3 v) { return mul(vc.view_to_world, float4(v, 0)).xyz(); }
inline float3 direction_world_to_view(const kjb_view_constants& vc, float3 v) { return mul(vc.world_to_view, float4(v, 0)).xyz(); }
inline float3 position_world_to_clip(const kjb_view_constants& vc, float3 v) {
    float4 p = mul(vc.world_to_view, float4(v, 1)); p = mul(vc.view_to_clip, p); return p.xyz() / p.w;
}
inline float3 position_world_to_sample(const kjb_view_constants& vc, float3 v) {
    float4 p = mul(vc.world_to_view, float4(v, 1)); p = mul(vc.view_to_sample, p); return p.xyz() / p.w;
}
// clip_to_view._11 (HLSL row 1, col 1) = m[0]; ._43 (row 4, col 3) = m[2*4+3]
inline float pixel_cone_spread_angle_from_image_height(const kjb_view_constants& vc, float image_height) {
    return atan(2.0f * vc.clip_to_view.m[0] / image_height);
}
static const int hi_px_subpixels[4][2] = {{1, 1}, {1, 0}, {0, 0}, {0, 1}};   // frame_constants.hlsl:235-240
inline int2 halfres_subsample_offset(uint frame_index) { return int2(hi_px_subpixels[frame_index & 3][0], hi_px_subpixels[frame_index & 3][1]); }

// ---------------------------------------------------------------- ray_cone.hlsl
struct RayCone {
    float width, spread_angle;
    static RayCone from_spread_angle(float a) { RayCone r; r.width = 0; r.spread_angle = a; return r; }
    RayCone propagate(float surface_spread_angle, float hit_t) const {
        RayCone r; r.width = spread_angle * hit_t + width; r.spread_angle = spread_angle + surface_spread_angle; return r;
    }
    float width_at_t(float hit_t) const { return width + spread_angle * hit_t; }
};

// ---------------------------------------------------------------- blue_noise.hlsl:8-15
inline float4 blue_noise_for_pixel(const Globals& g, uint2 px, uint n) {
    float2 r2 = r2_sequence(n);
    uint2 offset(uint(r2.x * 256.0f), uint(r2.y * 256.0f));
    return g.blue_noise.load(int((px.x + offset.x) % 256u), int((px.y + offset.y) % 256u)) * 255.0f / 256.0f + 0.5f / 256.0f;
}

// ---------------------------------------------------------------- atmosphere_felix.hlsl
namespace atm {
static const float PLANET_RADIUS = 6371000.0f;
static const float ATMOSPHERE_HEIGHT = 100000.0f;
static const float RAYLEIGH_HEIGHT = ATMOSPHERE_HEIGHT * 0.08f;
static const float MIE_HEIGHT = ATMOSPHERE_HEIGHT * 0.012f;
inline float3 planet_center() { return float3(0, -PLANET_RADIUS, 0); }
inline float3 C_RAYLEIGH() { return float3(5.802f, 13.558f, 33.100f) * 1e-6f; }
inline float3 C_MIE() { return float3(3.996f, 3.996f, 3.996f) * 1e-6f; }
inline float3 C_OZONE() { return float3(0.650f, 1.881f, 0.085f) * 1e-6f; }

inline float2 SphereIntersection(float3 rayStart, float3 rayDir, float3 sphereCenter, float sphereRadius) {   // :50-66
    rayStart = rayStart - sphereCenter;
    float a = dot(rayDir, rayDir);
    float b = 2.0f * dot(rayStart, rayDir);
    float c = dot(rayStart, rayStart) - (sphereRadius * sphereRadius);
    float d = b * b - 4 * a * c;
    if (d < 0) return float2(-1.0f);
    d = sqrt(d);
    return float2(-b - d, -b + d) / (2 * a);
}
inline float2 AtmosphereIntersection(float3 s, float3 d) { return SphereIntersection(s, d, planet_center(), PLANET_RADIUS + ATMOSPHERE_HEIGHT); }
inline float PhaseRayleigh(float costh) { return 3 * (1 + costh * costh) / (16 * 3.14159265359f); }
inline float PhaseMie(float costh, float g = 0.85f) {
    g = min(g, 0.9381f);
    float k = 1.55f * g - 0.55f * g * g * g;
    float kcosth = k * costh;
    return (1 - k * k) / ((4 * 3.14159265359f) * (1 - kcosth) * (1 - kcosth));
}
inline float AtmosphereHeight(float3 positionWS) { return length(positionWS - planet_center()) - PLANET_RADIUS; }
inline float3 AtmosphereDensity(float h) {
    return float3(exp(-max(0.0f, h / RAYLEIGH_HEIGHT)), exp(-max(0.0f, h / MIE_HEIGHT)), max(0.0f, 1 - abs(h - 25000.0f) / 15000.0f));
}
inline float3 IntegrateOpticalDepth(float3 rayStart, float3 rayDir) {   // :124-144
    float2 intersection = AtmosphereIntersection(rayStart, rayDir);
    float rayLength = intersection.y;
    int sampleCount = 8;
    float stepSize = rayLength / float(sampleCount);
    float3 opticalDepth(0.0f);
    for (int i = 0; i < sampleCount; i++) {
        float3 localPosition = rayStart + rayDir * (float(i) + 0.5f) * stepSize;
        float localHeight = AtmosphereHeight(localPosition);
        float3 localDensity = AtmosphereDensity(localHeight);
        opticalDepth += localDensity * stepSize;
    }
    return opticalDepth;
}
inline float3 Absorb(float3 od) {   // :176-181
    return exp(-(od.x * C_RAYLEIGH() + od.y * C_MIE() * 1.1f + od.z * C_OZONE()) * 1.0f);
}
inline float3 IntegrateScattering(float3 rayStart, float3 rayDir, float rayLength, float3 lightDir, float3 lightColor, float3& transmittance) {   // :185-243
    const float sampleDistributionExponent = 5;
    float2 intersection = AtmosphereIntersection(rayStart, rayDir);
    rayLength = min(rayLength, intersection.y);
    if (intersection.x > 0) { rayStart = rayStart + rayDir * intersection.x; rayLength -= intersection.x; }
    float costh = dot(rayDir, lightDir);
    float phaseR = PhaseRayleigh(costh);
    float phaseM = PhaseMie(costh);
    int sampleCount = 16;
    float3 opticalDepth(0.0f), rayleigh(0.0f), mie(0.0f);
    float prevRayTime = 0;
    for (int i = 1; i <= sampleCount; i++) {
        float rayTime = pow(float(i) / float(sampleCount), sampleDistributionExponent) * rayLength;
        float stepSize = (rayTime - prevRayTime);
        float3 localPosition = rayStart + rayDir * lerp(prevRayTime, rayTime, 0.5f);
        float localHeight = AtmosphereHeight(localPosition);
        float3 localDensity = AtmosphereDensity(localHeight);
        opticalDepth += localDensity * stepSize;
        float3 viewTransmittance = Absorb(opticalDepth);
        float3 opticalDepthlight = IntegrateOpticalDepth(localPosition, lightDir);
        float3 lightTransmittance = Absorb(opticalDepthlight);
        rayleigh += viewTransmittance * lightTransmittance * phaseR * localDensity.x * stepSize;
        mie += viewTransmittance * lightTransmittance * phaseM * localDensity.y * stepSize;
        prevRayTime = rayTime;
    }
    transmittance = Absorb(opticalDepth);
    return (rayleigh * C_RAYLEIGH() + mie * C_MIE()) * lightColor * 20.0f;
}
}  // namespace atm

inline float3 sun_direction(const Globals& g) { return float3(g.fc.sun_direction[0], g.fc.sun_direction[1], g.fc.sun_direction[2]); }
inline float3 atmosphere_default(const Globals& g, float3 wi, float3 light_dir) {   // atmosphere.hlsl:7-24
    float3 transmittance;
    const float INF = asfloat(0x7f800000u);
    float3 sky_ambient(g.fc.sky_ambient[0], g.fc.sky_ambient[1], g.fc.sky_ambient[2]);
    float3 sun_mult(g.fc.sun_color_multiplier[0], g.fc.sun_color_multiplier[1], g.fc.sun_color_multiplier[2]);
    return (sky_ambient + sun_mult * atm::IntegrateScattering(float3(0.0f), wi, INF, light_dir, float3(1.0f), transmittance)) * g.fc.pre_exposure;
}
inline float3 sun_color_in_direction(const Globals& g, float3 dir) {   // sun.hlsl:21-27
    float3 sun_mult(g.fc.sun_color_multiplier[0], g.fc.sun_color_multiplier[1], g.fc.sun_color_multiplier[2]);
    return 20.0f * sun_mult * g.fc.pre_exposure * atm::Absorb(atm::IntegrateOpticalDepth(float3(0.0f), dir));
}
inline float3 sample_sun_direction(const Globals& g, float2 urand, bool soft) {   // sun.hlsl:33-42
    if (soft) {
        if (g.fc.sun_angular_radius_cos < 1.0f) {
            const float3x3 basis = build_orthonormal_basis(normalize(sun_direction(g)));
            return mul(basis, uniform_sample_cone(urand, g.fc.sun_angular_radius_cos));
        }
    }
    return sun_direction(g);
}

// ---------------------------------------------------------------- brdf.hlsl
static const float BRDF_SAMPLING_MIN_COS = 1e-5f;
struct BrdfValue { float3 value_over_pdf, value; float pdf = 0; float3 transmission_fraction;
    static BrdfValue invalid() { return BrdfValue(); } };
struct BrdfSample : BrdfValue { float3 wi; float approx_roughness = 0;
    static BrdfSample invalid() { BrdfSample r; r.wi = float3(0, 0, -1); return r; }
    bool is_valid() const { return wi.z > 1e-6f; } };

struct DiffuseBrdf {
    float3 albedo;
    BrdfSample sample(float3, float2 urand) const {   // brdf.hlsl:54-72
        float phi = urand.x * M_TAU_F;
        float cos_theta = sqrt(max(0.0f, 1.0f - urand.y));
        float sin_theta = sqrt(max(0.0f, 1.0f - cos_theta * cos_theta));
        BrdfSample res;
        float sin_phi = sin(phi), cos_phi = cos(phi);
        res.wi = float3(cos_phi * sin_theta, sin_phi * sin_theta, cos_theta);
        res.pdf = M_FRAC_1_PI_F;
        res.value_over_pdf = albedo;
        res.value = res.value_over_pdf * res.pdf;
        res.transmission_fraction = float3(0.0f);
        res.approx_roughness = 1.0f;
        return res;
    }
    BrdfValue evaluate(float3, float3 wi) const {     // :74-81
        BrdfValue res;
        res.pdf = wi.z > 0.0f ? M_FRAC_1_PI_F : 0.0f;
        res.value_over_pdf = wi.z > 0.0f ? albedo : float3(0.0f);
        res.value = res.value_over_pdf * res.pdf;
        res.transmission_fraction = float3(0.0f);
        return res;
    }
};
inline float3 eval_fresnel_schlick(float3 f0, float3 f90, float cos_theta) {   // :95-97
    return lerp(f0, f90, pow(max(0.0f, 1.0f - cos_theta), 5.0f));
}
inline float g_smith_ggx_correlated(float ndotv, float ndotl, float a2) {      // :107-112
    float lambda_v = ndotl * sqrt((-ndotv * a2 + ndotv) * ndotv + a2);
    float lambda_l = ndotv * sqrt((-ndotl * a2 + ndotl) * ndotl + a2);
    return 2.0f * ndotl * ndotv / (lambda_v + lambda_l);
}
inline float g_smith_ggx1(float ndotv, float a2) {                             // :114-117
    float tan2_v = (1.0f - ndotv * ndotv) / (ndotv * ndotv);
    return 2.0f / (1.0f + sqrt(1.0f + a2 * tan2_v));
}
struct SmithShadowingMasking { float g, g_over_g1_wo;
    static SmithShadowingMasking eval(float ndotv, float ndotl, float a2) {    // :127-137
        SmithShadowingMasking r; r.g = g_smith_ggx_correlated(ndotv, ndotl, a2); r.g_over_g1_wo = r.g / g_smith_ggx1(ndotv, a2); return r; } };
struct NdfSample { float3 m; float pdf; };
struct SpecularBrdf {
    float roughness; float3 albedo;
    static float ggx_ndf(float a2, float cos_theta) { float ds = cos_theta * cos_theta * (a2 - 1.0f) + 1.0f; return a2 / (M_PI_F * ds * ds); }
    static float ggx_ndf_0_1(float a2, float cos_theta) { float ds = cos_theta * cos_theta * (a2 - 1.0f) + 1.0f; return a2 * a2 / (ds * ds); }
    static float pdf_ggx_vn(float a2, float3 wo, float3 h) {                   // :162-166
        float g1 = g_smith_ggx1(wo.z, a2);
        float d = ggx_ndf(a2, h.z);
        return g1 * d * max(0.f, dot(wo, h)) / wo.z;
    }
    NdfSample sample_vndf(float alpha, float3 wo, float2 urand) const {        // :186-214
        float alpha_x = alpha, alpha_y = alpha;
        float a2 = alpha_x * alpha_y;
        float3 Vh = normalize(float3(alpha_x * wo.x, alpha_y * wo.y, wo.z));
        float3 T1 = (Vh.z < 0.9999f) ? normalize(cross(float3(0, 0, 1), Vh)) : float3(1, 0, 0);
        float3 T2 = cross(Vh, T1);
        float r = sqrt(urand.x);
        float phi = (2.f * M_PI_F) * urand.y;
        float t1 = r * cos(phi);
        float t2 = r * sin(phi);
        float s = 0.5f * (1.f + Vh.z);
        t2 = (1.f - s) * sqrt(1.f - t1 * t1) + s * t2;
        float3 Nh = t1 * T1 + t2 * T2 + sqrt(max(0.f, 1.f - t1 * t1 - t2 * t2)) * Vh;
        float3 h = normalize(float3(alpha_x * Nh.x, alpha_y * Nh.y, max(0.f, Nh.z)));
        NdfSample res; res.m = h; res.pdf = pdf_ggx_vn(a2, wo, h);
        return res;
    }
    BrdfSample sample(float3 wo, float2 urand) const {                         // :216-262
        NdfSample ndf_sample = sample_vndf(roughness, wo, urand);
        const float3 wi = reflect(-wo, ndf_sample.m);
        if (ndf_sample.m.z <= BRDF_SAMPLING_MIN_COS || wi.z <= BRDF_SAMPLING_MIN_COS || wo.z <= BRDF_SAMPLING_MIN_COS) return BrdfSample::invalid();
        const float jacobian = 1.0f / (4.0f * dot(wi, ndf_sample.m));
        const float3 fresnel = eval_fresnel_schlick(albedo, float3(1.0f), dot(ndf_sample.m, wi));
        const float a2 = roughness * roughness;
        const float cos_theta = ndf_sample.m.z;
        SmithShadowingMasking sm = SmithShadowingMasking::eval(wo.z, wi.z, a2);
        BrdfSample res;
        res.pdf = ndf_sample.pdf * jacobian / wi.z;
        res.wi = wi;
        res.transmission_fraction = float3(1.0f) - fresnel;
        res.approx_roughness = roughness;
        res.value_over_pdf = fresnel * sm.g_over_g1_wo;
        res.value = fresnel * sm.g * ggx_ndf(a2, cos_theta) / (4 * wo.z * wi.z);
        return res;
    }
    BrdfValue evaluate(float3 wo, float3 wi) const {                           // :264-306
        if (wi.z <= 0.0f || wo.z <= 0.0f) return BrdfValue::invalid();
        const float a2 = roughness * roughness;
        const float3 m = normalize(wo + wi);
        const float cos_theta = m.z;
        const float pdf_h = pdf_ggx_vn(a2, wo, m);
        const float jacobian = 1.0f / (4.0f * dot(wi, m));
        const float3 fresnel = eval_fresnel_schlick(albedo, float3(1.0f), dot(m, wi));
        SmithShadowingMasking sm = SmithShadowingMasking::eval(wo.z, wi.z, a2);
        BrdfValue res;
        res.pdf = pdf_h * jacobian / wi.z;
        res.transmission_fraction = float3(1.0f) - fresnel;
        res.value_over_pdf = fresnel * sm.g_over_g1_wo;
        res.value = fresnel * sm.g * ggx_ndf(a2, cos_theta) / (4 * wo.z * wi.z);
        return res;
    }
};

// ---------------------------------------------------------------- brdf_lut.hlsl
struct SpecularBrdfEnergyPreservation {
    float3 preintegrated_reflection, preintegrated_reflection_mult, preintegrated_transmission_fraction;
    float valid_sample_fraction;
    static float3 sample_fg_lut(const Globals& g, float ndotv, float roughness) {   // :10-13
        const float2 scale((64.0f - 1.0f) / 64.0f), bias(0.5f / 64.0f);
        float2 uv = float2(ndotv, roughness) * scale + bias;
        return g.brdf_fg_lut.sample_bilinear_clamp(uv).xyz();
    }
    static SpecularBrdfEnergyPreservation from_brdf_ndotv(const Globals& g, const SpecularBrdf& brdf, float ndotv) {   // :15-93 (the `#elif 1` branch :57-77)
        const float roughness = brdf.roughness;
        const float3 specular_albedo = brdf.albedo;
        float3 fg = sample_fg_lut(g, ndotv, roughness);
        float3 single_scatter = specular_albedo * fg.x + fg.y;
        SpecularBrdfEnergyPreservation res;
        res.valid_sample_fraction = fg.z;
        float e_ss = fg.x + fg.y;
        float3 f_ss = single_scatter / e_ss;
        float3 f_ss_tail = lerp(f_ss, float3(1.0f), 0.4f);
        float3 bounce_radiance = (1.0f - e_ss) * f_ss_tail;
        float3 mult = 1.0f + bounce_radiance / (1.0f - bounce_radiance);
        res.preintegrated_reflection = single_scatter * mult;
        res.preintegrated_reflection_mult = mult;
        res.preintegrated_transmission_fraction = 1.0f - res.preintegrated_reflection;
        return res;
    }
};

// ---------------------------------------------------------------- layered_brdf.hlsl
inline float3 metalness_albedo_boost(float metalness, float3 diffuse_albedo) {   // :11-22
    const float a0 = 1.749f, a1 = -1.61f, e1 = 0.5555f, e3 = 0.8244f;
    const float x = metalness;
    const float3 y = diffuse_albedo;
    const float3 y3 = y * y * y;
    return 1.0f + (0.25f - (x - 0.5f) * (x - 0.5f)) * (a0 + a1 * abs(x - 0.5f)) * (e1 * y + e3 * y3);
}
inline void apply_metalness_to_brdfs(SpecularBrdf& s, DiffuseBrdf& d, float metalness) {   // :24-38
    const float3 albedo = d.albedo;
    s.albedo = lerp(s.albedo, albedo, metalness);
    d.albedo = max(0.0f, 1.0f - metalness) * albedo;
    const float3 boost = metalness_albedo_boost(metalness, albedo);
    s.albedo = min(float3(1.0f), s.albedo * boost);
    d.albedo = min(float3(1.0f), d.albedo * boost);
}
struct LayeredBrdf {
    SpecularBrdf specular_brdf; DiffuseBrdf diffuse_brdf; SpecularBrdfEnergyPreservation energy_preservation;
    static LayeredBrdf from_gbuffer_ndotv(const Globals& g, const GbufferData& gb, float ndotv) {   // :45-66
        SpecularBrdf s; s.albedo = float3(0.04f); s.roughness = gb.roughness;
        DiffuseBrdf d; d.albedo = gb.albedo;
        apply_metalness_to_brdfs(s, d, gb.metalness);
        LayeredBrdf res;
        res.energy_preservation = SpecularBrdfEnergyPreservation::from_brdf_ndotv(g, s, ndotv);
        res.specular_brdf = s; res.diffuse_brdf = d;
        return res;
    }
    float3 evaluate(float3 wo, float3 wi) const {   // :68-89
        if (wo.z <= 0 || wi.z <= 0) return float3(0.0f);
        const BrdfValue diff = diffuse_brdf.evaluate(wo, wi);
        const BrdfValue spec = specular_brdf.evaluate(wo, wi);
        return spec.value * energy_preservation.preintegrated_reflection_mult + diff.value * spec.transmission_fraction;
    }
    float3 evaluate_directional_light(float3 wo, float3 wi) const {   // :91-121
        if (wo.z <= 0 || wi.z <= 0) return float3(0.0f);
        const BrdfValue diff = diffuse_brdf.evaluate(wo, wi);
        const BrdfValue spec = specular_brdf.evaluate(wo, wi);
        const float3 mult_dir = lerp(float3(1.0f), energy_preservation.preintegrated_reflection_mult, sqrt(abs(wi.z)));
        return spec.value * mult_dir + diff.value * spec.transmission_fraction;
    }
    BrdfSample sample(float3 wo, float3 urand) const {   // :123-169
        BrdfSample bs;
        const float spec_wt = sRGB_to_luminance(energy_preservation.preintegrated_reflection);
        const float diffuse_wt = sRGB_to_luminance(energy_preservation.preintegrated_transmission_fraction * diffuse_brdf.albedo);
        const float transmission_p = diffuse_wt / (spec_wt + diffuse_wt);
        const float lobe_xi = urand.z;
        if (lobe_xi < transmission_p) {
            bs = diffuse_brdf.sample(wo, float2(urand.x, urand.y));
            const float lobe_pdf = transmission_p;
            bs.value_over_pdf = bs.value_over_pdf / lobe_pdf;
            bs.pdf *= lobe_pdf;
            bs.value_over_pdf *= energy_preservation.preintegrated_transmission_fraction;
            bs.value *= energy_preservation.preintegrated_transmission_fraction;
        } else {
            bs = specular_brdf.sample(wo, float2(urand.x, urand.y));
            const float lobe_pdf = (1.0f - transmission_p);
            bs.value_over_pdf = bs.value_over_pdf / lobe_pdf;
            bs.pdf *= lobe_pdf;
            bs.value_over_pdf *= energy_preservation.preintegrated_reflection_mult;
            bs.value *= energy_preservation.preintegrated_reflection_mult;
        }
        return bs;
    }
};

// ---------------------------------------------------------------- lights/triangle.hlsl
struct LightSampleResultArea { float3 pos, normal; float pdf; };
inline LightSampleResultArea sample_triangle_light(float3 v, float3 e0, float3 e1, float2 urand) {   // :36-42, :71-80
    float3 perp = cross(e0, e1);
    float perp_inv_len = rsqrt(dot(perp, perp));
    LightSampleResultArea res;
    float su0 = sqrt(urand.x);
    float b0 = 1.0f - su0;
    float b1 = urand.y * su0;
    res.pos = v + b0 * e0 + b1 * e1;
    res.normal = perp * perp_inv_len;
    res.pdf = 2.0f * perp_inv_len;
    return res;
}

}  // namespace kjo
