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


// ------------------------------------------------------------------ lighting/sample_lights.rgen.hlsl:18-63
int kjb_pass_sample_lights(kjb_context* ctx, const kjb_sample_lights_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img depth_tex(a->depth_tex), out0_tex(a->out0_tex), out1_tex(a->out1_tex), out2_tex(a->out2_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = out0_tex.w(), H = out0_tex.h();
    if (g.fc.triangle_light_count == 0) { ctx->last_error = "sample lights: the scene has no triangle lights"; return 1; }
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 hi_px(x * 2 + hso.x, y * 2 + hso.y);
        const float depth = depth_tex.load(hi_px).x;
        if (0.0f == depth) { out0_tex.store(x, y, float4(0.0f)); continue; }
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_depth(vc, uv, depth);
        const float3 shadow_ray_origin = view_ray_context.biased_secondary_ray_origin_ws();
        const float4 urand4 = blue_noise_for_pixel(g, uint2(uint(x), uint(y)), g.fc.frame_index);
        const float2 urand(urand4.x, urand4.y);
        const uint light_count = g.fc.triangle_light_count;
        const uint light_idx = kjb_cvt_u32(urand4.z * float(light_count)) % light_count;
        const float light_choice_pmf = 1.0f / float(light_count);
        const kjb_triangle_light& tl = g.lights[light_idx];
        const float3 v0(tl.verts[0][0], tl.verts[0][1], tl.verts[0][2]), v1(tl.verts[1][0], tl.verts[1][1], tl.verts[1][2]), v2(tl.verts[2][0], tl.verts[2][1], tl.verts[2][2]);
        const LightSampleResultArea light_sample = sample_triangle_light(v0, v1 - v0, v2 - v0, urand);
        const float3 to_light_ws = light_sample.pos - shadow_ray_origin;
        const float dist_to_light = length(to_light_ws);
        const bool is_shadowed = rt_is_shadowed(ctx->scene, shadow_ray_origin, to_light_ws / max(1e-8f, dist_to_light), 0.0f, dist_to_light - 1e-4f);
        const float3 radiance(tl.radiance[0], tl.radiance[1], tl.radiance[2]);
        out0_tex.store(x, y, float4(is_shadowed ? float3(0.0f) : radiance, 1.0f));
        out1_tex.store(x, y, float4(view_ray_context.ray_hit_vs() + direction_world_to_view(vc, to_light_ws), light_sample.pdf * light_choice_pmf));
        out2_tex.store(x, y, float4(direction_world_to_view(vc, light_sample.normal), 0.0f));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ lighting/spatial_reuse_lights.hlsl:33-168 (SHUFFLE_SUBPIXELS, BORROW_SAMPLES, USE_APPROXIMATE_SAMPLE_SHADOWING, RENDER_INTO_RTR)
int kjb_pass_spatial_reuse_lights(kjb_context* ctx, const kjb_spatial_reuse_lights_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    if (!a->spatial_resolve_offsets) { ctx->last_error = "spatial reuse lights: spatial_resolve_offsets is null"; return 1; }
    Img gbuffer_tex(a->gbuffer_tex), depth_tex(a->depth_tex), hit0_tex(a->hit0_tex), hit1_tex(a->hit1_tex), hit2_tex(a->hit2_tex), half_view_normal_tex(a->half_view_normal_tex),
        half_depth_tex(a->half_depth_tex), output_tex(a->output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int32_t* offs = a->spatial_resolve_offsets;
    const int W = output_tex.w(), H = output_tex.h();
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 uv = get_uv(int2(x, y), output_tex_size);
        const float depth = depth_tex.load(x, y).x;
        if (0.0f == depth) continue;
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_depth(vc, uv, depth);
        GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(x, y));
        gbuffer.roughness = max(gbuffer.roughness, 3e-4f);
        const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
        float3 wo = mul(-normalize(view_ray_context.ray_dir_ws()), tangent_to_world);
        if (wo.z < 0.0f) { wo.z *= -0.25f; wo = normalize(wo); }
        const LayeredBrdf layered_brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z);
        const SpecularBrdf specular_brdf = layered_brdf.specular_brdf;
        const float3 energy_preservation_mult = layered_brdf.energy_preservation.preintegrated_reflection_mult;
        const uint px_idx_in_quad = (((uint(x) & 1u) | (uint(y) & 1u) * 2u) + g.fc.frame_index) & 3u;
        const uint sample_count = 8, filter_idx = 3;
        float4 contrib_accum(0.0f);
        const float3 normal_vs = direction_world_to_view(vc, gbuffer.normal);
        for (uint sample_i = 0; sample_i < sample_count; ++sample_i) {
            const int32_t* o = offs + 4 * ((px_idx_in_quad * 16 + sample_i) + 64 * filter_idx);
            const int2 sample_px(x / 2 + o[0], y / 2 + o[1]);
            const float sample_depth = half_depth_tex.load(sample_px).x;
            const float4 packed0 = hit0_tex.load(sample_px);
            if (packed0.w != 0 && sample_depth != 0) {
                const float2 sample_uv = get_uv(int2(sample_px.x * 2 + hso.x, sample_px.y * 2 + hso.y), output_tex_size);
                const ViewRayContext sample_ray_ctx = ViewRayContext::from_uv_and_depth(vc, sample_uv, sample_depth);
                const float3 sample_origin_vs = sample_ray_ctx.ray_hit_vs();
                const float4 packed1 = hit1_tex.load(sample_px);
                float neighbor_sampling_pdf = packed1.w;
                const float3 sample_hit_normal_vs = hit2_tex.load(sample_px).xyz();
                const float3 center_to_hit_vs = packed1.xyz() - lerp(view_ray_context.ray_hit_vs(), sample_origin_vs, 0.5f);
                const float3 wi = normalize(mul(direction_view_to_world(vc, center_to_hit_vs), tangent_to_world));
                const float3 sample_normal_vs = half_view_normal_tex.load(sample_px).xyz();
                float rejection_bias = 1;
                rejection_bias *= saturate((dot(normal_vs, sample_normal_vs) - 0.9f) / (0.999f - 0.9f));   // inverse_lerp(0.9, 0.999, .), math.hlsl
                rejection_bias *= exp2(-10.0f * abs(depth / sample_depth - 1.0f));
                {
                    const float3 surface_offset = sample_origin_vs - view_ray_context.ray_hit_vs();
                    const float fraction_of_normal_direction_as_offset = dot(surface_offset, normal_vs) / length(surface_offset);
                    if (wi.z > 0 && wi.z * 0.2f < fraction_of_normal_direction_as_offset) rejection_bias *= sample_i == 0 ? 1.0f : 0.0f;
                }
                const BrdfValue spec = specular_brdf.evaluate(wo, wi);
                const float center_to_hit_dist2 = dot(center_to_hit_vs, center_to_hit_vs);
                const float to_psa_metric = max(0.0f, wi.z) * max(0.0f, dot(sample_hit_normal_vs, -normalize(center_to_hit_vs))) / center_to_hit_dist2;
                neighbor_sampling_pdf /= to_psa_metric;
                const float3 contrib_rgb = packed0.xyz() * spec.value * energy_preservation_mult * step(0.0f, wi.z) * (neighbor_sampling_pdf > 0 ? (1 / neighbor_sampling_pdf) : 0.0f);
                const float contrib_wt = rejection_bias;
                contrib_accum = contrib_accum + float4(contrib_rgb, 1) * contrib_wt;
            }
        }
        const float contrib_norm_factor = max(1e-8f, contrib_accum.w);
        const float3 out_color = contrib_accum.xyz() / contrib_norm_factor;
        output_tex.store(x, y, float4(output_tex.load(x, y).xyz() + out_color, 1.0f));   // RENDER_INTO_RTR: output_tex[px].rgb += out_color
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"
}  // namespace kjo
