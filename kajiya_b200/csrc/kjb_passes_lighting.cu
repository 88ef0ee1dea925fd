// Lighting composite as sm_100a kernels: "trace shadow mask" (renderers/shadows.rs:10-35,
// rt/trace_sun_shadow_mask.rgen.hlsl) and "light gbuffer" (renderers/deferred.rs:8-43, light_gbuffer.hlsl).
#include "kjb_context.h"

using namespace kjb;

// ------------------------------------------------------------------ rt/trace_sun_shadow_mask.rgen.hlsl:19-60
KJB_KERNEL(128) k_trace_sun_shadow_mask(Globals g, Img depth_tex, Img geometric_normal_tex, ImgW output_tex, Rows kjb_rows) {
    KJB_PX; if (x >= output_tex.w || y >= output_tex.h) return;
    const kjb_view_constants& vc = g.fc.view_constants;
    const float2 uv = (f2(float(x), float(y)) + 0.5f) / f2(float(output_tex.w), float(output_tex.h));
    const float z_over_w = ld_r32f(depth_tex, x, y);
    if (0.0f == z_over_w) { st_r8u(output_tex, x, y, 1.0f); return; }
    const float2 cs = uv_to_cs(uv);
    float4 pt_vs = mul(vc.sample_to_view, f4(cs.x, cs.y, z_over_w, 1.0f));
    float4 pt_ws = mul(vc.view_to_world, pt_vs);
    pt_ws = pt_ws / pt_ws.w; pt_vs = pt_vs / pt_vs.w;
    const float3 normal_vs = ld_a2r10g10b10(geometric_normal_tex, x, y) * 2.0f - 1.0f;
    const float3 normal_ws = xyz(mul(vc.view_to_world, f4(normal_vs, 0.0f)));
    const float bias_amount = (-pt_vs.z + length(xyz(pt_ws))) * 1e-5f;
    const float3 ray_origin = xyz(pt_ws) + normal_ws * bias_amount;
    const float4 bn = blue_noise_for_pixel(g, uint32_t(x), uint32_t(y), g.fc.frame_index);
    const bool is_shadowed = rt_is_shadowed(g, ray_origin, sample_sun_direction(g.fc, f2(bn.x, bn.y), true), 0.0f, KJB_FLT_MAX);
    st_r8u(output_tex, x, y, is_shadowed ? 0.0f : 1.0f);
}

// ------------------------------------------------------------------ light_gbuffer.hlsl:60-260 (debug_shading_mode 0, 2, 3, 4)
struct LightGbufferImgs { Img gbuffer_tex, depth_tex, shadow_mask_tex, rtr_tex, rtdgi_tex, unconvolved_sky_cube_tex; ImgW temporal_output_tex, output_tex; };
KJB_KERNEL(256) k_light_gbuffer(Globals g, LightGbufferImgs t, float4 ots, uint32_t mode, float real_sun_radius_cos, Rows kjb_rows) {
    KJB_PX; if (x >= t.output_tex.w || y >= t.output_tex.h) return;
    const kjb_view_constants& vc = g.fc.view_constants;
    const float s4[4] = {ots.x, ots.y, ots.z, ots.w};
    const float2 uv = get_uv(x, y, s4);
    const ViewRayContext vrc = ViewRayContext::from_uv(vc, uv);
    const float3 ray_dir = vrc.ray_dir_ws();
    const float depth = ld_r32f(t.depth_tex, x, y);
    if (depth == 0.0f) {   // sky + sun disk
        const float real_sun_angular_radius = 0.53f * 0.5f * KJB_PI_F / 180.0f;
        const float sun_angular_radius_cos = kjb_min(real_sun_radius_cos, g.fc.sun_angular_radius_cos);
        const float current_sun_angular_radius = kjb_acos(sun_angular_radius_cos);
        const float sun_radius_ratio = real_sun_angular_radius / current_sun_angular_radius;
        float3 output = xyz(sample_cube_rgba16f(t.unconvolved_sky_cube_tex, ray_dir));
        if (dot(ray_dir, sun_direction(g.fc)) > sun_angular_radius_cos) output += 800.0f * sun_color_in_direction(g.fc, ray_dir) * sun_radius_ratio * sun_radius_ratio;
        st_rgba16f(t.temporal_output_tex, x, y, f4(output, 1)); st_rgba16f(t.output_tex, x, y, f4(output, 1));
        return;
    }
    const float3 to_light_norm = sun_direction(g.fc);
    float shadow_mask = ld_r8u(t.shadow_mask_tex, x, y);
    if (mode == 4u) shadow_mask = 1;
    const GbufferData gbuffer = gbuffer_unpack(ld_rgba32u(t.gbuffer_tex, x, y));
    const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
    const float3 wi = mul(to_light_norm, tangent_to_world);
    float3 wo = mul(-ray_dir, tangent_to_world);
    if (wo.z < 0.0f) { wo.z *= -0.25f; wo = normalize(wo); }
    const LayeredBrdf brdf = layered_brdf_from_gbuffer_ndotv(g, gbuffer, wo.z);
    const float3 brdf_value = layered_evaluate_directional_light(brdf, wo, wi) * kjb_max(0.0f, wi.z);
    const float3 light_radiance = shadow_mask * f3(g.sun_color[0], g.sun_color[1], g.sun_color[2]);
    float3 total_radiance = brdf_value * light_radiance;
    total_radiance += gbuffer.emissive;
    float3 gi_irradiance = f3(0.0f);
    if (mode != 4u) gi_irradiance = xyz(ld_rgba16f(t.rtdgi_tex, x, y));
    total_radiance += gi_irradiance * brdf.diffuse_brdf.albedo * brdf.ep.preintegrated_transmission_fraction;
    const float3 rtr = ld_r11g11b10(t.rtr_tex, x, y);
    if (mode != 4u) total_radiance += rtr * brdf.ep.preintegrated_reflection;   // !RTR_RENDER_SCALED_BY_FG
    st_rgba16f(t.temporal_output_tex, x, y, f4(total_radiance, 1.0f));
    float3 output = total_radiance;
    if (mode == 3u) { output = rtr * brdf.ep.preintegrated_reflection; output = output / brdf.ep.preintegrated_reflection; }
    if (mode == 2u) output = gi_irradiance;
    st_rgba16f(t.output_tex, x, y, f4(output, 1.0f));
}

#define CHK(img, fmt, name) if (!check_img(c, (img), (fmt), P, name)) return 1
#define CHKE(img, fmt, name, w, h) if (!check_img(c, (img), (fmt), P, name, (w), (h))) return 1

extern "C" {

int kjb_pass_trace_sun_shadow_mask(kjb_context* c, const kjb_trace_sun_shadow_mask_args* a) {
    const char* P = "trace shadow mask"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_R8_UNORM, "output_tex"); CHKE(a->depth_tex, KJB_FMT_R32_FLOAT, "depth_tex", W, H); CHKE(a->geometric_normal_tex, KJB_FMT_A2R10G10B10_UNORM, "geometric_normal_tex", W, H);
    if (!c->tlas_valid) return c->fail("trace shadow mask: no acceleration structure (call kjb_rebuild_tlas)");
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_trace_sun_shadow_mask, KJB_GRID2D(W, H, 16, 8), c->g, img_ro(a->depth_tex), img_ro(a->geometric_normal_tex), img_rw(a->output_tex));
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_light_gbuffer(kjb_context* c, const kjb_light_gbuffer_args* a) {
    const char* P = "light gbuffer"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    if (a->debug_show_wrc || a->debug_shading_mode == 1 || a->debug_shading_mode > 4) return c->fail("light gbuffer: unsupported debug mode");
    CHK(a->output_tex, KJB_FMT_RGBA16_FLOAT, "output_tex"); CHKE(a->temporal_output_tex, KJB_FMT_RGBA16_FLOAT, "temporal_output_tex", W, H); CHKE(a->gbuffer_tex, KJB_FMT_RGBA32_FLOAT, "gbuffer_tex", W, H);
    CHKE(a->depth_tex, KJB_FMT_R32_FLOAT, "depth_tex", W, H); CHKE(a->shadow_mask_tex, KJB_FMT_R8_UNORM, "shadow_mask_tex", W, H); CHKE(a->rtr_tex, KJB_FMT_R11G11B10_UFLOAT, "rtr_tex", W, H);
    CHKE(a->rtdgi_tex, KJB_FMT_RGBA16_FLOAT, "rtdgi_tex", W, H); CHK(a->unconvolved_sky_cube_tex, KJB_FMT_RGBA16_FLOAT, "unconvolved_sky_cube_tex");
    LightGbufferImgs t{img_ro(a->gbuffer_tex), img_ro(a->depth_tex), img_ro(a->shadow_mask_tex), img_ro(a->rtr_tex), img_ro(a->rtdgi_tex), img_ro(a->unconvolved_sky_cube_tex),
                       img_rw(a->temporal_output_tex), img_rw(a->output_tex)};
    const float real_sun_radius_cos = kjb_cos(0.53f * 0.5f * KJB_PI_F / 180.0f);
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_light_gbuffer, KJB_GRID2D(W, H, 32, 8), c->g, t, f4(a->output_tex_size[0], a->output_tex_size[1], a->output_tex_size[2], a->output_tex_size[3]), a->debug_shading_mode, real_sun_radius_cos);
    KJB_PASS_EPILOGUE(c, P);
}

}  // extern "C"
