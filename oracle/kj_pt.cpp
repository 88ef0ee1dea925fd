// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// CPU restatement of kajiya's reference path tracer, /root/reference/assets/shaders/rt/reference_path_trace.rgen.hlsl:75-377
// (a Vulkan ray-generation shader in the reference — there is no CPU path tracer upstream).
// Compile-time switches as shipped (:20-43): MAX_EYE_PATH_LENGTH 16, RR from 3, FIREFLY_SUPPRESSION, USE_PIXEL_FILTER,
// USE_SOFT_SHADOWS, USE_LIGHTS, USE_EMISSIVE on; INDIRECT_ONLY is a runtime argument here.
#include "kj_ctx.h"

namespace kjo {

namespace {
const uint MAX_EYE_PATH_LENGTH = 16;
const uint RUSSIAN_ROULETTE_START_PATH_LENGTH = 3;

float inv_error_function(float x, float truncation) {   // :60-68
    const float ALPHA = 0.14f;
    const float INV_ALPHA = 1.0f / ALPHA;
    const float K = 2.0f / (M_PI_F * ALPHA);
    float y = log(max(truncation, 1.0f - x * x));
    float z = K + 0.5f * y;
    return sqrt(max(0.0f, sqrt(z * z - y * INV_ALPHA) - z)) * sign(x);
}
float remap_unorm_to_gaussian(float x, float truncation) { return inv_error_function(x * 2.0f - 1.0f, truncation); }
}

extern "C" int kjb_pass_reference_path_trace(kjb_context* ctx, const kjb_reference_pt_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img output_tex(a->output_tex);
    const int W = output_tex.w(), H = output_tex.h();
    const bool INDIRECT_ONLY = a->indirect_only != 0;
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        float4 prev = output_tex.load(x, y);
        if (!(prev.w < 1000)) continue;
        float4 radiance_sample_count_packed(0.0f);
        uint rng = hash_combine2(hash_combine2(uint(x), hash1(uint(y))), g.fc.frame_index);
        {
            float px_off0 = 0.5f, px_off1 = 0.5f;
            const float psf_scale = 0.4f;
            px_off0 += psf_scale * remap_unorm_to_gaussian(uint_to_u01_float(hash1_mut(rng)), 1e-8f);
            px_off1 += psf_scale * remap_unorm_to_gaussian(uint_to_u01_float(hash1_mut(rng)), 1e-8f);
            const float2 pixel_center = float2(float(x), float(y)) + float2(px_off0, px_off1);
            const float2 uv = pixel_center / float2(float(W), float(H));

            Ray outgoing_ray;
            {
                const ViewRayContext view_ray_context = ViewRayContext::from_uv(vc, uv);
                const float3 ray_dir_ws = view_ray_context.ray_dir_ws();
                outgoing_ray.origin = view_ray_context.ray_origin_ws(); outgoing_ray.dir = normalize(ray_dir_ws); outgoing_ray.tmin = 0.0f; outgoing_ray.tmax = FLT_MAX_F;
            }
            float3 throughput(1.0f), total_radiance(0.0f);
            float roughness_bias = 0.0f;
            RayCone ray_cone = RayCone::from_spread_angle(pixel_cone_spread_angle_from_image_height(vc, float(H)));
            ray_cone.spread_angle *= 0.3f;

            for (uint path_length = 0; path_length < MAX_EYE_PATH_LENGTH; ++path_length) {
                if (path_length == 1) outgoing_ray.tmax = FLT_MAX_F;
                GbufferPathVertex primary_hit = gbuffer_raytrace(ctx->scene, g, outgoing_ray, ray_cone, path_length, false);
                if (primary_hit.is_hit) {
                    ray_cone = ray_cone.propagate(0.0f, primary_hit.ray_t);
                    float2 sun_urand; sun_urand.x = uint_to_u01_float(hash1_mut(rng)); sun_urand.y = uint_to_u01_float(hash1_mut(rng));
                    const float3 to_light_norm = sample_sun_direction(g, sun_urand, true);
                    const bool is_shadowed = (INDIRECT_ONLY && path_length == 0) || rt_is_shadowed(ctx->scene, primary_hit.position, to_light_norm, 1e-4f, FLT_MAX_F);
                    GbufferData gbuffer = gbuffer_unpack(primary_hit.gbuffer_packed);
                    if (dot(gbuffer.normal, outgoing_ray.dir) >= 0.0f) {
                        if (0 == path_length) gbuffer.normal = -gbuffer.normal; else break;
                    }
                    if (INDIRECT_ONLY && path_length == 0) { gbuffer.albedo = float3(1.0f); gbuffer.metalness = 0.0f; }
                    const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
                    const float3 wi = mul(to_light_norm, tangent_to_world);
                    float3 wo = mul(-outgoing_ray.dir, tangent_to_world);
                    if (wo.z < 0.0f) { wo.z *= -0.25f; wo = normalize(wo); }
                    LayeredBrdf brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z);
                    brdf.specular_brdf.roughness = lerp(brdf.specular_brdf.roughness, 1.0f, roughness_bias);   // FIREFLY_SUPPRESSION
                    {
                        const float3 brdf_value = brdf.evaluate_directional_light(wo, wi);
                        const float3 light_radiance = is_shadowed ? float3(0.0f) : sun_color_in_direction(g, sun_direction(g));
                        total_radiance += throughput * brdf_value * light_radiance * max(0.0f, wi.z);
                        total_radiance += gbuffer.emissive * throughput;
                        if (g.fc.triangle_light_count > 0) {
                            const float light_selection_pmf = 1.0f / float(g.fc.triangle_light_count);
                            const uint light_idx = hash1_mut(rng) % g.fc.triangle_light_count;
                            float2 urand; urand.x = uint_to_u01_float(hash1_mut(rng)); urand.y = uint_to_u01_float(hash1_mut(rng));
                            const kjb_triangle_light& tl = g.lights[light_idx];
                            float3 v0(tl.verts[0][0], tl.verts[0][1], tl.verts[0][2]), v1(tl.verts[1][0], tl.verts[1][1], tl.verts[1][2]), v2(tl.verts[2][0], tl.verts[2][1], tl.verts[2][2]);
                            LightSampleResultArea ls = sample_triangle_light(v0, v1 - v0, v2 - v0, urand);
                            const float3 shadow_ray_origin = primary_hit.position;
                            const float3 to_light_ws = ls.pos - primary_hit.position;
                            const float dist_to_light2 = dot(to_light_ws, to_light_ws);
                            const float3 to_light_norm_ws = to_light_ws * rsqrt(dist_to_light2);
                            const float to_psa_metric = max(0.0f, dot(to_light_norm_ws, gbuffer.normal)) * max(0.0f, dot(to_light_norm_ws, -ls.normal)) / dist_to_light2;
                            if (to_psa_metric > 0.0f) {
                                float3 wi2 = mul(to_light_norm_ws, tangent_to_world);
                                const bool sh = rt_is_shadowed(ctx->scene, shadow_ray_origin, to_light_norm_ws, 1e-3f, sqrt(dist_to_light2) - 2e-3f);
                                float3 radiance(tl.radiance[0], tl.radiance[1], tl.radiance[2]);
                                total_radiance += sh ? float3(0.0f) : throughput * radiance * brdf.evaluate(wo, wi2) / ls.pdf * to_psa_metric / light_selection_pmf;
                            }
                        }
                    }
                    float3 urand;
                    urand.x = uint_to_u01_float(hash1_mut(rng)); urand.y = uint_to_u01_float(hash1_mut(rng)); urand.z = uint_to_u01_float(hash1_mut(rng));
                    BrdfSample brdf_sample = brdf.sample(wo, urand);
                    if (brdf_sample.is_valid()) {
                        roughness_bias = lerp(roughness_bias, 1.0f, 0.5f * brdf_sample.approx_roughness);
                        outgoing_ray.origin = primary_hit.position;
                        outgoing_ray.dir = mul(tangent_to_world, brdf_sample.wi);
                        outgoing_ray.tmin = 1e-4f;
                        throughput *= brdf_sample.value_over_pdf;
                    } else break;
                    if (path_length >= RUSSIAN_ROULETTE_START_PATH_LENGTH) {
                        const float rr_coin = uint_to_u01_float(hash1_mut(rng));
                        const float continue_p = max(gbuffer.albedo.x, max(gbuffer.albedo.y, gbuffer.albedo.z));
                        if (rr_coin > continue_p) break; else throughput /= continue_p;
                    }
                } else {
                    total_radiance += throughput * atmosphere_default(g, outgoing_ray.dir, sun_direction(g));
                    break;
                }
            }
            if (total_radiance.x >= 0.0f && total_radiance.y >= 0.0f && total_radiance.z >= 0.0f) radiance_sample_count_packed += float4(total_radiance, 1.0f);
        }
        float4 cur = radiance_sample_count_packed;
        float tsc = cur.w + prev.w;
        float lrp = cur.w / max(1.0f, tsc);
        float3 c = cur.xyz() / max(1.0f, cur.w);
        output_tex.store(x, y, float4(max(float3(0.0f), lerp(prev.xyz(), c, lrp)), max(1.0f, tsc)));
    } }, ctx->num_threads);
    return 0;
}

}  // namespace kjo
