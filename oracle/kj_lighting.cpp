// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// "trace shadow mask" (rt/trace_sun_shadow_mask.rgen.hlsl) and "light gbuffer" (light_gbuffer.hlsl).
#include "kj_ctx.h"

namespace kjo {
extern "C" {

// ------------------------------------------------------------------ rt/trace_sun_shadow_mask.rgen.hlsl:19-60
int kjb_pass_trace_sun_shadow_mask(kjb_context* ctx, const kjb_trace_sun_shadow_mask_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img depth_tex(a->depth_tex), geometric_normal_tex(a->geometric_normal_tex), output_tex(a->output_tex);
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 uv = (float2(float(x), float(y)) + 0.5f) / float2(float(W), float(H));
        const float z_over_w = depth_tex.load(x, y).x;
        if (0.0f == z_over_w) { output_tex.store(x, y, float4(1.0f)); continue; }
        const float2 cs = uv_to_cs(uv);
        float4 pt_vs = mul(vc.sample_to_view, float4(cs.x, cs.y, z_over_w, 1.0f));
        float4 pt_ws = mul(vc.view_to_world, pt_vs);
        pt_ws = pt_ws / pt_ws.w; pt_vs = pt_vs / pt_vs.w;
        const float3 normal_vs = geometric_normal_tex.load(x, y).xyz() * 2.0f - 1.0f;
        const float3 normal_ws = mul(vc.view_to_world, float4(normal_vs, 0.0f)).xyz();
        const float bias_amount = (-pt_vs.z + length(pt_ws.xyz())) * 1e-5f;
        const float3 ray_origin = pt_ws.xyz() + normal_ws * bias_amount;
        const float4 bn = blue_noise_for_pixel(g, uint2(uint(x), uint(y)), g.fc.frame_index);
        const bool is_shadowed = rt_is_shadowed(ctx->scene, ray_origin, sample_sun_direction(g, float2(bn.x, bn.y), true), 0.0f, FLT_MAX_F);
        output_tex.store(x, y, float4(is_shadowed ? 0.0f : 1.0f));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ light_gbuffer.hlsl:60-260 (modes 0, 2, 3, 4)
int kjb_pass_light_gbuffer(kjb_context* ctx, const kjb_light_gbuffer_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    if (a->debug_show_wrc || a->debug_shading_mode == 1 || a->debug_shading_mode > 4) { ctx->last_error = "light gbuffer: unsupported debug mode"; return 1; }
    Img gbuffer_tex(a->gbuffer_tex), depth_tex(a->depth_tex), shadow_mask_tex(a->shadow_mask_tex), rtr_tex(a->rtr_tex), rtdgi_tex(a->rtdgi_tex),
        temporal_output_tex(a->temporal_output_tex), output_tex(a->output_tex), unconvolved_sky_cube_tex(a->unconvolved_sky_cube_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const uint mode = a->debug_shading_mode;
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 uv = get_uv(int2(x, y), output_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv(vc, uv);
        const float3 ray_dir = view_ray_context.ray_dir_ws();
        const float depth = depth_tex.load(x, y).x;
        if (depth == 0.0f) {   // sky + sun disk
            const float real_sun_angular_radius = 0.53f * 0.5f * M_PI_F / 180.0f;
            const float sun_angular_radius_cos = min(cos(real_sun_angular_radius), g.fc.sun_angular_radius_cos);
            const float current_sun_angular_radius = acos(sun_angular_radius_cos);
            const float sun_radius_ratio = real_sun_angular_radius / current_sun_angular_radius;
            float3 output = unconvolved_sky_cube_tex.sample_cube(ray_dir).xyz();
            if (dot(ray_dir, sun_direction(g)) > sun_angular_radius_cos) output += 800.0f * sun_color_in_direction(g, ray_dir) * sun_radius_ratio * sun_radius_ratio;
            temporal_output_tex.store(x, y, float4(output, 1)); output_tex.store(x, y, float4(output, 1));
            continue;
        }
        const float3 to_light_norm = sun_direction(g);
        float shadow_mask = shadow_mask_tex.load(x, y).x;
        if (mode == 4) shadow_mask = 1;
        const GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(x, y));
        const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
        const float3 wi = mul(to_light_norm, tangent_to_world);
        float3 wo = mul(-ray_dir, tangent_to_world);
        if (wo.z < 0.0f) { wo.z *= -0.25f; wo = normalize(wo); }
        const LayeredBrdf brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z);
        const float3 brdf_value = brdf.evaluate_directional_light(wo, wi) * max(0.0f, wi.z);
        const float3 light_radiance = shadow_mask * sun_color_in_direction(g, sun_direction(g));
        float3 total_radiance = brdf_value * light_radiance;
        total_radiance += gbuffer.emissive;
        float3 gi_irradiance(0.0f);
        if (mode != 4) gi_irradiance = rtdgi_tex.load(x, y).xyz();
        total_radiance += gi_irradiance * brdf.diffuse_brdf.albedo * brdf.energy_preservation.preintegrated_transmission_fraction;
        if (mode != 4) total_radiance += rtr_tex.load(x, y).xyz() * brdf.energy_preservation.preintegrated_reflection;   // !RTR_RENDER_SCALED_BY_FG
        temporal_output_tex.store(x, y, float4(total_radiance, 1.0f));
        float3 output = total_radiance;
        if (mode == 3) {
            output = rtr_tex.load(x, y).xyz() * brdf.energy_preservation.preintegrated_reflection;
            output = output / brdf.energy_preservation.preintegrated_reflection;   // true_brdf == brdf (textures on)
        }
        if (mode == 2) output = gi_irradiance;
        output_tex.store(x, y, float4(output, 1.0f));
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"
}  // namespace kjo
