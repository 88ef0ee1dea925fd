// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// CPU restatement of kajiya's ray-traced reflections, one function per render-graph pass of crates/lib/kajiya/src/renderers/rtr.rs.
// Shader paths are relative to /root/reference/assets/shaders/.  Settings frozen to rtr/rtr_settings.hlsl.
#include "kj_ircache_lookup.h"

namespace kjo {
namespace {

const float SKY_DIST = 1e4f;                           // rtr/reflection_trace_common.inc.hlsl:3
const float RTR_ROUGHNESS_CLAMP = 6e-4f;               // rtr/rtr_settings.hlsl:42
const float RTR_RESTIR_TEMPORAL_M_CLAMP = 8.0f;        // :10
const float RTR_RESTIR_MAX_PDF_CLAMP = 200.0f;         // :46
const float RTR_NEIGHBOR_RAY_ORIGIN_CENTER_BIAS = 0.5f;// :20
const float SAMPLING_BIAS = 0.15f;                     // reflection_trace_common.inc.hlsl:37-43 (USE_HEAVY_BIAS)
inline float rtr_encode_cos_theta_for_fp16(float x) { return 1 - x; }
inline float rtr_decode_cos_theta_from_fp16(float x) { return 1 - x; }
inline int2 reservoir_payload_to_px(uint payload) { return int2(int(payload & 0xffff), int(payload >> 16)); }
inline float3 get_prev_eye_position(const kjb_view_constants& vc) { float4 e = mul(vc.prev_view_to_prev_world, float4(0, 0, 0, 1)); return e.xyz() / e.w; }   // frame_constants.hlsl:187-190
inline float3 position_world_to_view(const kjb_view_constants& vc, float3 v) { return mul(vc.world_to_view, float4(v, 1)).xyz(); }                         // :204-206
inline float depth_to_view_z(const kjb_view_constants& vc, float depth) { return rcp(depth * -vc.clip_to_view.m[2 * 4 + 3]); }                                // :192-194 (`_43`)
inline float length_squared(float3 v) { return dot(v, v); }
inline float3 specular_dominant_direction(float3 n, float3 v, float roughness) {   // inc/brdf.hlsl:313-317
    float3 r = reflect(-v, n);
    float f = (1.0f - roughness) * (sqrt(1.0f - roughness) + roughness);
    return normalize(lerp(n, r, f));
}
inline float3 soft_color_clamp(float3 center, float3 history, float3 ex, float3 dev) {   // inc/soft_color_clamp.hlsl
    float3 history_dist = abs(history - ex) / max(abs(history * 0.1f), dev);
    float3 closest_pt = clamp(history, center - dev, center + dev);
    return lerp(history, closest_pt, float3(smoothstep(1.0f, 3.0f, history_dist.x), smoothstep(1.0f, 3.0f, history_dist.y), smoothstep(1.0f, 3.0f, history_dist.z)));
}

// rtr/rtr_restir_pack_unpack.inc.hlsl:1-22
struct RtrRestirRayOrigin {
    float3 ray_origin_eye_offset_ws; float roughness; uint frame_index_mod4;
    static RtrRestirRayOrigin from_raw(float4 raw) {
        RtrRestirRayOrigin r; r.ray_origin_eye_offset_ws = raw.xyz();
        float2 misc = unpack_2x16f_uint(asuint(raw.w));
        r.roughness = misc.x; r.frame_index_mod4 = kjb_cvt_u32(misc.y) & 3u;
        return r;
    }
    float4 to_raw() const { return float4(ray_origin_eye_offset_ws, asfloat(pack_2x16f_uint(float2(roughness, float(frame_index_mod4))))); }
};

// inc/blue_noise.hlsl:28-56 (Heitz/Belcour spp64 tables supplied by the host; see kjb_rtr_trace_args)
inline float blue_noise_sampler(const uint32_t* ranking, const uint32_t* scrambling, const uint32_t* sobol, int pixel_i, int pixel_j, int sampleIndex, int sampleDimension) {
    pixel_i &= 127; pixel_j &= 127; sampleIndex &= 255; sampleDimension &= 255;
    const int rankedSampleIndex = sampleIndex ^ int(ranking[sampleDimension + (pixel_i + pixel_j * 128) * 8]);
    int value = int(sobol[sampleDimension + rankedSampleIndex * 256]);
    value = value ^ int(scrambling[(sampleDimension % 8) + (pixel_i + pixel_j * 128) * 8]);
    return (0.5f + float(value)) / 256.0f;
}

struct RtrTraceResult { float3 total_radiance; float hit_t; float3 hit_normal_vs; };

// rtr/reflection_trace_common.inc.hlsl:49-257 (USE_WORLD_RADIANCE_CACHE 0, USE_HEAVY_BIAS 1)
RtrTraceResult do_the_thing(const kjb_context& ctx, const Img& gbuffer_tex, const Img& depth_tex, const Img& rtdgi_tex, const Img& sky_cube_tex, float4 gbuffer_tex_size,
                            const IrcacheBufs& ircache, uint2 px, float3 normal_ws, float roughness, uint& rng, Ray outgoing_ray) {
    const Globals& g = ctx.g; const kjb_view_constants& vc = g.fc.view_constants;
    const float roughness_bias = roughness;   // USE_AGGRESSIVE_SECONDARY_ROUGHNESS_BIAS
    const float reflected_cone_spread_angle = sqrt(roughness) * 0.05f;
    const RayCone ray_cone = RayCone::from_spread_angle(pixel_cone_spread_angle_from_image_height(vc, gbuffer_tex_size.y))
        .propagate(reflected_cone_spread_angle, length(outgoing_ray.origin - get_eye_position(vc)));
    const GbufferPathVertex primary_hit = gbuffer_raytrace(ctx.scene, g, outgoing_ray, ray_cone, 1, false);
    if (primary_hit.is_hit) {
        GbufferData gbuffer = gbuffer_unpack(primary_hit.gbuffer_packed);
        gbuffer.roughness = lerp(gbuffer.roughness, 1.0f, roughness_bias);
        const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
        const float3 wo = mul(-outgoing_ray.dir, tangent_to_world);
        const LayeredBrdf brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z);

        const float3 primary_hit_cs = position_world_to_sample(vc, primary_hit.position);
        const float2 primary_hit_uv = cs_to_uv(float2(primary_hit_cs.x, primary_hit_cs.y));
        const float primary_hit_screen_depth = depth_tex.sample_nearest_clamp(primary_hit_uv).x;
        const uint4 screen_gb = gbuffer_tex.load_u(kjb_cvt_i32(primary_hit_uv.x * gbuffer_tex_size.x), kjb_cvt_i32(primary_hit_uv.y * gbuffer_tex_size.y));
        const float3 primary_hit_screen_normal_ws = unpack_normal_11_10_11(asfloat(screen_gb.y));
        const bool is_on_screen = abs(primary_hit_cs.x) < 1.0f && abs(primary_hit_cs.y) < 1.0f
            && inverse_depth_relative_diff(primary_hit_cs.z, primary_hit_screen_depth) < 5e-3f
            && dot(primary_hit_screen_normal_ws, -outgoing_ray.dir) > 0.0f
            && dot(primary_hit_screen_normal_ws, gbuffer.normal) > 0.7f;

        float3 total_radiance(0.0f);
        {   // Sun
            float2 urand; urand.x = uint_to_u01_float(hash1_mut(rng)); urand.y = uint_to_u01_float(hash1_mut(rng));
            const float3 to_light_norm = sample_sun_direction(g, urand, true);
            const bool is_shadowed = rt_is_shadowed(ctx.scene, primary_hit.position, to_light_norm, 1e-4f, SKY_DIST);
            const float3 wi = mul(to_light_norm, tangent_to_world);
            const float3 brdf_value = brdf.evaluate(wo, wi) * max(0.0f, wi.z);
            const float3 light_radiance = is_shadowed ? float3(0.0f) : sun_color_in_direction(g, sun_direction(g));
            total_radiance += brdf_value * light_radiance;
        }
        const float3 reflected_normal_vs = direction_world_to_view(vc, gbuffer.normal);
        total_radiance += gbuffer.emissive;
        if (is_on_screen) {   // USE_SCREEN_GI_REPROJECTION
            const float3 reprojected_radiance = rtdgi_tex.sample_nearest_clamp(primary_hit_uv).xyz() * g.fc.pre_exposure_delta;
            total_radiance += reprojected_radiance * gbuffer.albedo;
        } else {
            {   // USE_LIGHTS
                float2 urand; urand.x = uint_to_u01_float(hash1_mut(rng)); urand.y = uint_to_u01_float(hash1_mut(rng));
                for (uint light_idx = 0; light_idx < g.fc.triangle_light_count; light_idx += 1) {
                    const kjb_triangle_light& tl = g.lights[light_idx];
                    float3 v0(tl.verts[0][0], tl.verts[0][1], tl.verts[0][2]), v1(tl.verts[1][0], tl.verts[1][1], tl.verts[1][2]), v2(tl.verts[2][0], tl.verts[2][1], tl.verts[2][2]);
                    LightSampleResultArea ls = sample_triangle_light(v0, v1 - v0, v2 - v0, urand);
                    const float3 shadow_ray_origin = primary_hit.position;
                    const float3 to_light_ws = ls.pos - shadow_ray_origin;
                    const float dist_to_light2 = dot(to_light_ws, to_light_ws);
                    const float3 to_light_norm_ws = to_light_ws * rsqrt(dist_to_light2);
                    const float to_psa_metric = max(0.0f, dot(to_light_norm_ws, gbuffer.normal)) * max(0.0f, dot(to_light_norm_ws, -ls.normal)) / dist_to_light2;
                    if (to_psa_metric > 0.0f) {
                        const bool is_shadowed = rt_is_shadowed(ctx.scene, shadow_ray_origin, to_light_norm_ws, 1e-4f, sqrt(dist_to_light2) - 2e-4f);
                        const float3 bounce_albedo = lerp(gbuffer.albedo, float3(1.0f), 0.04f);
                        const float3 brdf_value = bounce_albedo * to_psa_metric / M_PI_F;
                        float3 radiance(tl.radiance[0], tl.radiance[1], tl.radiance[2]);
                        total_radiance += !is_shadowed ? (radiance * brdf_value / ls.pdf) : float3(0.0f);
                    }
                }
            }
            {   // USE_IRCACHE
                const float cone_width = ray_cone.propagate(0, primary_hit.ray_t).width;
                const float3 gi = ircache_lookup(g, ircache, outgoing_ray.origin, primary_hit.position, gbuffer.normal, 1, rng, false, cone_width < 0.1f);
                total_radiance += gi * gbuffer.albedo;
            }
        }
        RtrTraceResult result; result.total_radiance = total_radiance; result.hit_t = primary_hit.ray_t; result.hit_normal_vs = reflected_normal_vs;
        return result;
    }
    RtrTraceResult result;
    result.total_radiance = sky_cube_tex.sample_cube(outgoing_ray.dir).xyz();
    result.hit_t = SKY_DIST;
    result.hit_normal_vs = -direction_world_to_view(vc, outgoing_ray.dir);
    return result;
}

inline float3 flip_wo(float3 wo) { if (wo.z < 0.0f) { wo.z *= -0.25f; wo = normalize(wo); } return wo; }   // "shading normals facing away" hack, all rtr passes

}  // namespace

extern "C" {

// ------------------------------------------------------------------ R1: rtr/reflection.rgen.hlsl:41-169
int kjb_pass_rtr_trace(kjb_context* ctx, const kjb_rtr_trace_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img gbuffer_tex(a->gbuffer_tex), depth_tex(a->depth_tex), rtdgi_tex(a->rtdgi_tex), sky_cube_tex(a->sky_cube_tex), out0_tex(a->out0_tex), out1_tex(a->out1_tex), out2_tex(a->out2_tex), rng_out_tex(a->rng_out_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const IrcacheBufs ircache = IrcacheBufs::from(a->ircache);
    const uint32_t *ranking = (const uint32_t*)a->ranking_tile_buf.data, *scrambling = (const uint32_t*)a->scambling_tile_buf.data, *sobol = (const uint32_t*)a->sobol_buf.data;
    const bool have_sampler_tables = ranking && scrambling && sobol;
    const int W = out0_tex.w(), H = out0_tex.h();
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);
    pass_pixels(ctx, W, H, ircache.bound(), [&](int x, int y) {
        const int2 px(x, y); const int2 hi_px = px * 2 + hso;
        const float depth = depth_tex.load(hi_px).x;
        if (0.0f == depth) { out0_tex.store(px, float4(0, 0, 0, -SKY_DIST)); return; }
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(hi_px));
        gbuffer.roughness = max(gbuffer.roughness, RTR_ROUGHNESS_CLAMP);
        if (a->reuse_rtdgi_rays && gbuffer.roughness > 0.6f) return;   // keep the diffuse candidates
        const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_biased_depth(vc, uv, depth);
        const float3 refl_ray_origin_ws = view_ray_context.biased_secondary_ray_origin_ws_with_normal(gbuffer.normal);
        const float3 wo = flip_wo(mul(-view_ray_context.ray_dir_ws(), tangent_to_world));
        SpecularBrdf specular_brdf;
        specular_brdf.albedo = lerp(float3(0.04f), gbuffer.albedo, gbuffer.metalness);
        specular_brdf.roughness = gbuffer.roughness;
        const uint noise_offset = g.fc.frame_index;   // USE_TEMPORAL_JITTER
        uint rng = hash3(uint(x), uint(y), noise_offset);
        float2 urand;
        if (have_sampler_tables) { urand.x = blue_noise_sampler(ranking, scrambling, sobol, x, y, int(noise_offset), 0); urand.y = blue_noise_sampler(ranking, scrambling, sobol, x, y, int(noise_offset), 1); }
        else { const float4 bn = blue_noise_for_pixel(g, uint2(x, y), noise_offset); urand = float2(bn.x, bn.y); }
        urand.x = lerp(urand.x, 0.0f, SAMPLING_BIAS);
        BrdfSample brdf_sample = specular_brdf.sample(wo, urand);
        for (uint retry_i = 0; retry_i < 4 && !brdf_sample.is_valid(); ++retry_i) {
            urand.x = uint_to_u01_float(hash1_mut(rng)); urand.y = uint_to_u01_float(hash1_mut(rng));
            urand.x = lerp(urand.x, 0.0f, SAMPLING_BIAS);
            brdf_sample = specular_brdf.sample(wo, urand);
        }
        const float cos_theta = normalize(wo + brdf_sample.wi).z;
        if (brdf_sample.is_valid()) {
            Ray outgoing_ray; outgoing_ray.dir = mul(tangent_to_world, brdf_sample.wi); outgoing_ray.origin = refl_ray_origin_ws; outgoing_ray.tmin = 0; outgoing_ray.tmax = SKY_DIST;
            rng_out_tex.store_u(x, y, uint4(rng, 0, 0, 0));
            const RtrTraceResult result = do_the_thing(*ctx, gbuffer_tex, depth_tex, rtdgi_tex, sky_cube_tex, gbuffer_tex_size, ircache, uint2(x, y), gbuffer.normal, gbuffer.roughness, rng, outgoing_ray);
            const float3 hit_offset_ws = outgoing_ray.dir * result.hit_t;
            const SpecularBrdfEnergyPreservation brdf_lut = SpecularBrdfEnergyPreservation::from_brdf_ndotv(g, specular_brdf, wo.z);
            const float pdf = brdf_sample.pdf / brdf_lut.valid_sample_fraction;
            out0_tex.store(px, float4(result.total_radiance, rtr_encode_cos_theta_for_fp16(cos_theta)));
            out1_tex.store(px, float4(hit_offset_ws, pdf));
            out2_tex.store(px, float4(result.hit_normal_vs, 0));
        } else {
            out0_tex.store(px, float4(1, 0, 1, 0));
            out1_tex.store(px, float4(0.0f));
        }
    });
    return 0;
}

// ------------------------------------------------------------------ R2: rtr/reflection_validate.rgen.hlsl:42-146
int kjb_pass_rtr_validate(kjb_context* ctx, const kjb_rtr_validate_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img gbuffer_tex(a->gbuffer_tex), depth_tex(a->depth_tex), rtdgi_tex(a->rtdgi_tex), sky_cube_tex(a->sky_cube_tex), invalidity_tex(a->refl_restir_invalidity_tex),
        ray_orig_history_tex(a->ray_orig_history_tex), ray_history_tex(a->ray_history_tex), rng_history_tex(a->rng_history_tex), irradiance_history_tex(a->irradiance_history_tex),
        reservoir_history_tex(a->reservoir_history_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const IrcacheBufs ircache = IrcacheBufs::from(a->ircache);
    // dispatched over the half-res image's half_res() extent (rtr.rs:229): one thread per 2x2 quad of half-res pixels
    const int W = (invalidity_tex.w() + 1) / 2, H = (invalidity_tex.h() + 1) / 2;
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);
    const float ped = g.fc.pre_exposure_delta;
    // Quads are disjoint, each thread touches only its own quad: row-parallel unless the irradiance cache is bound.
    // (The tile scissor addresses half-res rows; a quad row q covers half-res rows 2q, 2q+1.)
    const uint32_t sy0 = ctx->scissor_y0, sy1 = ctx->scissor_y1;
    if (sy1 > sy0) { ctx->scissor_y0 = sy0 / 2; ctx->scissor_y1 = (sy1 + 1) / 2; }   // scissor in quad rows for the loop below
    pass_pixels(ctx, W, H, ircache.bound(), [&](int qx, int qy) {
        const int2 px = int2(qx, qy) * 2 + hso;
        const int2 hi_px = px * 2 + hso;
        const float depth = depth_tex.load(hi_px).x;
        if (0.0f == depth) { invalidity_tex.store(px, float4(1.0f)); return; }
        GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(hi_px));
        gbuffer.roughness = max(gbuffer.roughness, RTR_ROUGHNESS_CLAMP);
        const float3 ray_orig_ws = ray_orig_history_tex.load(px).xyz() + get_prev_eye_position(vc);
        const float3 ray_hit_ws = ray_history_tex.load(px).xyz() + ray_orig_ws;
        Ray outgoing_ray; outgoing_ray.dir = normalize(ray_hit_ws - ray_orig_ws); outgoing_ray.origin = ray_orig_ws; outgoing_ray.tmin = 0; outgoing_ray.tmax = SKY_DIST;
        uint rng = rng_history_tex.load_u(px).x;
        const RtrTraceResult result = do_the_thing(*ctx, gbuffer_tex, depth_tex, rtdgi_tex, sky_cube_tex, gbuffer_tex_size, ircache, uint2(px.x, px.y), gbuffer.normal, gbuffer.roughness, rng, outgoing_ray);
        uint4 rraw = reservoir_history_tex.load_u(px);
        Reservoir1spp r = Reservoir1spp::from_raw(uint2(rraw.x, rraw.y));
        const float4 prev_irradiance_packed = irradiance_history_tex.load(px);
        const float3 prev_irradiance = max(float3(0.0f), prev_irradiance_packed.xyz() * ped);
        const float3 check_radiance = max(float3(0.0f), result.total_radiance);
        const float rad_diff = length(abs(prev_irradiance - check_radiance) / max(float3(1e-3f), prev_irradiance + check_radiance));
        const float invalidity = smoothstep(0.1f, 0.5f, rad_diff / length(float3(1.0f)));
        r.M *= 1 - invalidity;
        irradiance_history_tex.store(px, float4(check_radiance, prev_irradiance_packed.w));
        invalidity_tex.store(px, float4(invalidity));
        { uint2 rr = r.as_raw(); reservoir_history_tex.store_u(px.x, px.y, uint4(rr.x, rr.y, 0, 0)); }
        for (uint i = 1; i <= 3; ++i) {   // also reduce M of the quad neighbours
            const uint k = (g.fc.frame_index + i) & 3;
            const int2 npx = int2(qx, qy) * 2 + int2(hi_px_subpixels[k][0], hi_px_subpixels[k][1]);
            const float4 neighbor_prev_irradiance_packed = irradiance_history_tex.load(npx);
            {
                const float3 av = max(float3(0.0f), neighbor_prev_irradiance_packed.xyz() * ped);
                const float3 bv = prev_irradiance;
                const float neigh_rad_diff = length(abs(av - bv) / max(float3(1e-8f), av + bv));
                if (neigh_rad_diff < 0.2f) irradiance_history_tex.store(npx, float4(check_radiance, neighbor_prev_irradiance_packed.w));
            }
            invalidity_tex.store(npx, float4(invalidity));
            if (invalidity > 0) {
                uint4 nraw = reservoir_history_tex.load_u(npx);
                Reservoir1spp nr = Reservoir1spp::from_raw(uint2(nraw.x, nraw.y));
                nr.M *= 1 - invalidity;
                uint2 rr = nr.as_raw(); reservoir_history_tex.store_u(npx.x, npx.y, uint4(rr.x, rr.y, 0, 0));
            }
        }
    });
    ctx->scissor_y0 = sy0; ctx->scissor_y1 = sy1;
    return 0;
}

// ------------------------------------------------------------------ R3: rtr/rtr_restir_temporal.hlsl:148-533
int kjb_pass_rtr_restir_temporal(kjb_context* ctx, const kjb_rtr_restir_temporal_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img gbuffer_tex(a->gbuffer_tex), half_view_normal_tex(a->half_view_normal_tex), depth_tex(a->depth_tex), candidate0_tex(a->candidate0_tex), candidate1_tex(a->candidate1_tex),
        candidate2_tex(a->candidate2_tex), irradiance_history_tex(a->irradiance_history_tex), ray_orig_history_tex(a->ray_orig_history_tex), ray_history_tex(a->ray_history_tex),
        rng_history_tex(a->rng_history_tex), reservoir_history_tex(a->reservoir_history_tex), reprojection_tex(a->reprojection_tex), hit_normal_history_tex(a->hit_normal_history_tex),
        irradiance_out_tex(a->irradiance_out_tex), ray_orig_output_tex(a->ray_orig_output_tex), ray_output_tex(a->ray_output_tex), rng_output_tex(a->rng_output_tex),
        hit_normal_output_tex(a->hit_normal_output_tex), reservoir_out_tex(a->reservoir_out_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = irradiance_out_tex.w(), H = irradiance_out_tex.h();
    const int2 hi_px_offset = halfres_subsample_offset(g.fc.frame_index);
    const float3 eye = get_eye_position(vc), prev_eye = get_prev_eye_position(vc);
    const float ped = g.fc.pre_exposure_delta;

    // :103-146 find_best_reprojection_in_neighborhood
    auto find_best_reprojection_in_neighborhood = [&](float2 base_px, int2& best_px, float3 refl_ray_origin_ws, bool wide) {
        float best_dist = 1e10f;
        const float2 clip_scale(vc.clip_to_view.m[0], vc.clip_to_view.m[5]);
        const float2 offset_scale = float2(1, -1) * -2.0f * clip_scale * float2(gbuffer_tex_size.z, gbuffer_tex_size.w);
        const float3 look_direction = direction_view_to_world(vc, float3(0, 0, -1));
        {
            const float z_offset = dot(look_direction, refl_ray_origin_ws - eye);
            const float2 o = float2(float(hi_px_offset.x), float(hi_px_offset.y)) * offset_scale * z_offset;
            refl_ray_origin_ws += direction_view_to_world(vc, float3(o.x, o.y, 0));
        }
        const int start_coord = wide ? -1 : 0;
        for (int y = start_coord; y <= 1; ++y) for (int x = start_coord; x <= 1; ++x) {
            const int2 spx(kjb_cvt_i32(floor(base_px.x + float(x))), kjb_cvt_i32(floor(base_px.y + float(y))));
            const RtrRestirRayOrigin ray_orig = RtrRestirRayOrigin::from_raw(ray_orig_history_tex.load(spx));
            float3 orig = ray_orig.ray_origin_eye_offset_ws + prev_eye;
            const float2 orig_jitter(float(hi_px_subpixels[ray_orig.frame_index_mod4][0]), float(hi_px_subpixels[ray_orig.frame_index_mod4][1]));
            {
                const float z_offset = dot(look_direction, orig);
                const float2 o = orig_jitter * offset_scale * z_offset;
                orig += direction_view_to_world(vc, float3(o.x, o.y, 0));
            }
            const float d = length(orig - refl_ray_origin_ws);
            if (d < best_dist) { best_dist = d; best_px = spx; }
        }
    };

    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y); const int2 hi_px = px * 2 + hi_px_offset;
        const float depth = depth_tex.load(hi_px).x;
        if (0.0f == depth) {
            irradiance_out_tex.store(px, float4(0, 0, 0, -SKY_DIST)); hit_normal_output_tex.store(px, float4(0.0f)); reservoir_out_tex.store_u(x, y, uint4(0, 0, 0, 0));
            continue;
        }
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        const float3 normal_vs = half_view_normal_tex.load(px).xyz();
        const float3 normal_ws = direction_view_to_world(vc, normal_vs);
        float local_normal_flatness = 1;
        for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) local_normal_flatness *= saturate(dot(normal_vs, half_view_normal_tex.load(px + int2(xx, yy)).xyz()));
        float reprojection_neighborhood_stability = 1;
        for (int yy = 0; yy <= 1; ++yy) for (int xx = 0; xx <= 1; ++xx) reprojection_neighborhood_stability *= reprojection_tex.load(px * 2 + int2(xx, yy)).z;

        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_biased_depth(vc, uv, depth);
        const float3 refl_ray_origin_ws = view_ray_context.biased_secondary_ray_origin_ws_with_normal(normal_ws);
        const float3 refl_ray_origin_vs = position_world_to_view(vc, refl_ray_origin_ws);
        const float3x3 tangent_to_world = build_orthonormal_basis(normal_ws);
        float3 outgoing_dir(0, 0, 1);
        uint rng = hash3(uint(x), uint(y), g.fc.frame_index);
        const float3 wo = flip_wo(mul(-normalize(view_ray_context.ray_dir_ws()), tangent_to_world));
        const GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(hi_px));
        const float a2 = max(RTR_ROUGHNESS_CLAMP, gbuffer.roughness) * max(RTR_ROUGHNESS_CLAMP, gbuffer.roughness);

        float p_q_sel = 0, pdf_sel = 0, cos_theta = 0;
        float3 irradiance_sel(0.0f); float4 ray_orig_sel(0.0f); float3 ray_hit_sel_ws(1.0f), hit_normal_sel(1.0f);
        uint rng_sel = rng_output_tex.load_u(px).x;
        Reservoir1sppStreamState stream_state; Reservoir1spp reservoir;
        const uint reservoir_payload = uint(x) | (uint(y) << 16);
        reservoir.payload = reservoir_payload;
        {   // :68-83 do_the_thing = load the candidate
            const float4 hit0 = candidate0_tex.load(px), hit1 = candidate1_tex.load(px), hit2 = candidate2_tex.load(px);
            const float3 r_out_value = hit0.xyz(); const float r_pdf = min(hit1.w, RTR_RESTIR_MAX_PDF_CLAMP); const float r_cos_theta = rtr_decode_cos_theta_from_fp16(hit0.w);
            const float3 r_hit_vs = hit1.xyz(); const float3 r_hit_normal_ws = direction_view_to_world(vc, hit2.xyz());
            if (r_pdf > 0) {
                outgoing_dir = normalize(r_hit_vs);
                const float p_q = p_q_sel = 1 * max(1e-3f, sRGB_to_luminance(r_out_value)) * r_pdf;
                const float inv_pdf_q = 1.0f / r_pdf;
                pdf_sel = r_pdf; cos_theta = r_cos_theta; irradiance_sel = r_out_value;
                RtrRestirRayOrigin ray_orig; ray_orig.ray_origin_eye_offset_ws = refl_ray_origin_ws; ray_orig.roughness = gbuffer.roughness; ray_orig.frame_index_mod4 = g.fc.frame_index & 3;
                ray_orig_sel = ray_orig.to_raw();
                ray_hit_sel_ws = r_hit_vs + refl_ray_origin_ws;
                hit_normal_sel = r_hit_normal_ws;
                if (p_q * inv_pdf_q > 0) reservoir.init_with_stream(p_q, inv_pdf_q, stream_state, reservoir_payload);
            }
        }
        const float4 center_reproj = reprojection_tex.load(hi_px);
        {   // USE_RESAMPLING
            const float ang_offset = float(((g.fc.frame_index + 7u) * 11u) % 32u) * M_TAU_F;
            const uint max_samples = center_reproj.z < 1.0f ? 5u : 1u;
            for (uint sample_i = 0; sample_i < max_samples && stream_state.M_sum < RTR_RESTIR_TEMPORAL_M_CLAMP; ++sample_i) {
                const float ang = (float(sample_i) + ang_offset) * GOLDEN_ANGLE;
                const float rpx_offset_radius = sqrt(float(((sample_i - 1u) + g.fc.frame_index) & 3u) + 1.0f) * clamp(8.0f - stream_state.M_sum, 1.0f, 7.0f);
                const float2 reservoir_px_offset_base = float2(cos(ang), sin(ang)) * rpx_offset_radius;
                const int2 rpx_offset = sample_i == 0 ? int2(0, 0) : int2(kjb_cvt_i32(reservoir_px_offset_base.x), kjb_cvt_i32(reservoir_px_offset_base.y));
                const float4 reproj = reprojection_tex.load(hi_px + rpx_offset * 2);
                int2 reproj_px;
                {
                    const float2 base_px = float2(float(x), float(y)) + float2(gbuffer_tex_size.x, gbuffer_tex_size.y) * reproj.xy() / 2.0f;
                    int2 best_px(kjb_cvt_i32(floor(base_px.x + 0.5f)), kjb_cvt_i32(floor(base_px.y + 0.5f)));
                    // USE_REPROJECTION_SEARCH with USE_HALFRES_SUBSAMPLE_JITTERING
                    if (reprojection_neighborhood_stability >= 1) {
                        if (abs(gbuffer_tex_size.x * reproj.x) > 0.1f || abs(gbuffer_tex_size.y * reproj.y) > 0.1f) find_best_reprojection_in_neighborhood(base_px, best_px, refl_ray_origin_ws, false);
                    } else {
                        find_best_reprojection_in_neighborhood(base_px, best_px, refl_ray_origin_ws, true);
                    }
                    reproj_px = best_px;
                }
                const int2 rpx = reproj_px + rpx_offset;
                uint4 rraw = reservoir_history_tex.load_u(rpx);
                Reservoir1spp r = Reservoir1spp::from_raw(uint2(rraw.x, rraw.y));
                const int2 spx = reservoir_payload_to_px(r.payload);
                const float4 prev_ray_orig_and_roughness = ray_orig_history_tex.load(spx) + float4(prev_eye, 0);
                if (length_squared(refl_ray_origin_ws - prev_ray_orig_and_roughness.xyz()) > 0.05f * refl_ray_origin_vs.z * refl_ray_origin_vs.z) continue;   // disocclusion
                const float4 prev_irrad_and_cos_theta = irradiance_history_tex.load(spx) * float4(ped, ped, ped, 1);
                const float3 prev_irrad = prev_irrad_and_cos_theta.xyz();
                const float prev_cos_theta = rtr_decode_cos_theta_from_fp16(prev_irrad_and_cos_theta.w);
                const float4 sample_hit_ws_and_pdf_packed = ray_history_tex.load(spx);
                const float prev_pdf = sample_hit_ws_and_pdf_packed.w;
                const float3 sample_hit_ws = sample_hit_ws_and_pdf_packed.xyz() + prev_ray_orig_and_roughness.xyz();
                const float prev_dist = length(sample_hit_ws_and_pdf_packed.xyz());
                const float4 hn = hit_normal_history_tex.load(spx);
                const float4 sample_hit_normal_ws_dot(hn.x * 2 - 1, hn.y * 2 - 1, hn.z * 2 - 1, hn.w);
                const float3 dir_to_sample_hit_unnorm = sample_hit_ws - refl_ray_origin_ws;
                const float dist_to_sample_hit = length(dir_to_sample_hit_unnorm);
                const float3 dir_to_sample_hit = normalize(dir_to_sample_hit_unnorm);
                r.M = min(r.M, RTR_RESTIR_TEMPORAL_M_CLAMP);
                {   // USE_TRANSLATIONAL_CLAMP
                    const float3 current_wo = normalize(view_ray_context.ray_hit_ws() - eye);
                    const float3 prev_wo = normalize(view_ray_context.ray_hit_ws() - prev_eye);
                    const float wo_dot = saturate(dot(current_wo, prev_wo));
                    const float wo_similarity = pow(saturate(SpecularBrdf::ggx_ndf_0_1(max(3e-5f, a2), wo_dot)), 64.0f);
                    float mult = lerp(wo_similarity, 1.0f, smoothstep(0.05f, 0.5f, sqrt(gbuffer.roughness)));
                    mult = lerp(1.0f, mult, local_normal_flatness);
                    r.M *= mult;
                }
                float p_q = 1;
                p_q *= max(1e-3f, sRGB_to_luminance(prev_irrad));
                p_q *= step(0.0f, dot(dir_to_sample_hit, normal_ws));   // RTR_RESTIR_BRDF_SAMPLING
                p_q *= prev_pdf;
                const float visibility = 1;
                float jacobian = 1;
                jacobian *= clamp(prev_dist / dist_to_sample_hit, 1e-4f, 1e4f);
                jacobian *= jacobian;
                jacobian *= max(0.0f, -dot(sample_hit_normal_ws_dot.xyz(), dir_to_sample_hit)) / max(1e-5f, sample_hit_normal_ws_dot.w);
                {   // USE_JACOBIAN_BASED_REJECTION
                    const float JACOBIAN_REJECT_THRESHOLD = lerp(1.1f, 4.0f, gbuffer.roughness * gbuffer.roughness);
                    if (!(jacobian < JACOBIAN_REJECT_THRESHOLD && jacobian > 1.0f / JACOBIAN_REJECT_THRESHOLD)) continue;
                }
                p_q *= jacobian;
                if (reservoir.update_with_stream(r, p_q, visibility, stream_state, reservoir_payload, rng)) {
                    outgoing_dir = dir_to_sample_hit;
                    p_q_sel = p_q; pdf_sel = prev_pdf; cos_theta = prev_cos_theta; irradiance_sel = prev_irrad;
                    ray_orig_sel = prev_ray_orig_and_roughness;
                    ray_hit_sel_ws = sample_hit_ws;
                    hit_normal_sel = sample_hit_normal_ws_dot.xyz();
                    rng_sel = rng_history_tex.load_u(spx).x;
                }
            }
            reservoir.finish_stream(stream_state);
            reservoir.W = min(reservoir.W, 1e20f);   // RESTIR_RESERVOIR_W_CLAMP
        }
        (void)p_q_sel;
        const float4 hit_normal_ws_dot(hit_normal_sel, -dot(hit_normal_sel, outgoing_dir));
        irradiance_out_tex.store(px, float4(irradiance_sel, rtr_encode_cos_theta_for_fp16(cos_theta)));
        ray_orig_output_tex.store(px, float4(ray_orig_sel.xyz() - eye, ray_orig_sel.w));
        hit_normal_output_tex.store(px, float4(hit_normal_ws_dot.x * 0.5f + 0.5f, hit_normal_ws_dot.y * 0.5f + 0.5f, hit_normal_ws_dot.z * 0.5f + 0.5f, hit_normal_ws_dot.w));
        ray_output_tex.store(px, float4(ray_hit_sel_ws - ray_orig_sel.xyz(), pdf_sel));
        rng_output_tex.store_u(x, y, uint4(rng_sel, 0, 0, 0));
        { uint2 rr = reservoir.as_raw(); reservoir_out_tex.store_u(x, y, uint4(rr.x, rr.y, 0, 0)); }
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ R4: rtr/resolve.hlsl:78-663 (USE_RESTIR, BORROW_SAMPLES, CUT_CORNERS_IN_MATH)
int kjb_pass_rtr_resolve(kjb_context* ctx, const kjb_rtr_resolve_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img gbuffer_tex(a->gbuffer_tex), depth_tex(a->depth_tex), hit1_tex(a->hit1_tex), reprojection_tex(a->reprojection_tex), half_view_normal_tex(a->half_view_normal_tex),
        ray_len_history_tex(a->ray_len_history_tex), restir_irradiance_tex(a->restir_irradiance_tex), restir_ray_tex(a->restir_ray_tex), restir_reservoir_tex(a->restir_reservoir_tex),
        restir_ray_orig_tex(a->restir_ray_orig_tex), restir_hit_normal_tex(a->restir_hit_normal_tex), output_tex(a->output_tex), ray_len_output_tex(a->ray_len_output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);
    const float3 eye = get_eye_position(vc);
    const uint MAX_SAMPLE_COUNT = 8;
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y); const int2 half_px(x / 2, y / 2);
        const float2 uv = get_uv(px, output_tex_size);
        const float depth = depth_tex.load(px).x;
        if (0.0f == depth) { output_tex.store(px, float4(0.0f)); continue; }
        GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(px));
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_biased_depth(vc, uv, depth);
        const float3 refl_ray_origin_ws = view_ray_context.biased_secondary_ray_origin_ws_with_normal(gbuffer.normal);
        const float3 refl_ray_origin_vs = position_world_to_view(vc, refl_ray_origin_ws);
        gbuffer.roughness = max(gbuffer.roughness, RTR_ROUGHNESS_CLAMP);
        const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
        const float3 wo = flip_wo(mul(-normalize(view_ray_context.ray_dir_ws()), tangent_to_world));
        const SpecularBrdf specular_brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z).specular_brdf;
        const uint px_idx_in_quad = (((uint(x) & 1u) | (uint(y) & 1u) * 2u) + g.fc.frame_index) & 3u;   // SHUFFLE_SUBPIXELS
        const float a2 = max(RTR_ROUGHNESS_CLAMP, gbuffer.roughness) * max(RTR_ROUGHNESS_CLAMP, gbuffer.roughness);
        const float surf_to_hit_dist = length(hit1_tex.load(half_px).xyz());
        const float eye_to_surf_dist = length(refl_ray_origin_vs);
        const float eye_ray_z_scale = -view_ray_context.ray_dir_vs().z;
        const float4 reprojection_params = reprojection_tex.load(px);
        const float ray_squish_scale = 16.0f / max(1e-5f, eye_to_surf_dist);
        const float ray_len_avg = exponential_unsquish(lerp(
            exponential_squish(ray_len_history_tex.sample_bilinear_clamp(uv + reprojection_params.xy()).y, ray_squish_scale),
            exponential_squish(surf_to_hit_dist, ray_squish_scale), 0.1f), ray_squish_scale);
        const uint sample_count = MAX_SAMPLE_COUNT;
        float4 contrib_accum(0.0f); float ray_len_accum = 0;
        const float3 normal_vs = direction_world_to_view(vc, gbuffer.normal);
        const float tan_theta = sqrt(gbuffer.roughness) * 0.25f;
        const float c2v11 = vc.clip_to_view.m[5];
        float kernel_size_ws;
        {
            const float clamped_ray_len_avg = max(ray_len_avg, eye_to_surf_dist / eye_ray_z_scale * c2v11 * 0.2f * smoothstep(0.0f, 0.05f * eye_to_surf_dist, ray_len_avg));
            const float kernel_size_vs = clamped_ray_len_avg / (clamped_ray_len_avg + eye_to_surf_dist);
            kernel_size_ws = kernel_size_vs * eye_to_surf_dist * eye_ray_z_scale;
            kernel_size_ws *= tan_theta;
        }
        {
            const float scale_factor = eye_to_surf_dist * eye_ray_z_scale * c2v11;
            kernel_size_ws = min(kernel_size_ws, 0.1f * scale_factor);
            kernel_size_ws = max(kernel_size_ws, output_tex_size.w * 4.0f * scale_factor);
        }
        float3 kernel_t1, kernel_t2;
        {   // get_specular_filter_kernel_basis (:69-76)
            const float3 v = -normalize(view_ray_context.ray_dir_ws());
            const float3 dominant = specular_dominant_direction(gbuffer.normal, v, gbuffer.roughness);
            const float3 reflected = reflect(-dominant, gbuffer.normal);
            kernel_t1 = normalize(cross(gbuffer.normal, reflected)) * kernel_size_ws;
            kernel_t2 = cross(reflected, kernel_t1);
        }
        const float4 blue = blue_noise_for_pixel(g, uint2(uint(half_px.x) + 16u, uint(half_px.y) + 16u), g.fc.frame_index);
        const float KERNEL_SHARPNESS = 0.666f;
        const float RADIUS_SAMPLE_MULT = 1.0f / pow(float(MAX_SAMPLE_COUNT), KERNEL_SHARPNESS);
        const float ang_offset = float(g.fc.frame_index * 59u % 128u) * M_PLASTIC_F;
        const float RADIUS_INC_ON_FAIL = 0.25f;
        float sample_radius_accum = 1;
        for (int sample_i = 1; sample_i <= int(sample_count); ++sample_i, sample_radius_accum += RADIUS_INC_ON_FAIL) {
            const bool is_center_sample = sample_i == int(sample_count);
            int2 sample_offset;
            {
                const float ang = (float(sample_i) + ang_offset) * GOLDEN_ANGLE + (float(px_idx_in_quad) / 4.0f) * M_TAU_F;
                float sample_i_with_jitter = sample_radius_accum;
                if (is_center_sample) sample_i_with_jitter = contrib_accum.w > 1e-8f ? blue.y : 0.0f;
                else sample_i_with_jitter += blue.y;
                const float radius = pow(sample_i_with_jitter, KERNEL_SHARPNESS) * RADIUS_SAMPLE_MULT;
                const float3 offset_ws = (cos(ang) * kernel_t1 + sin(ang) * kernel_t2) * radius;
                const float3 sample_ws = refl_ray_origin_ws + offset_ws;
                const float3 sample_cs = position_world_to_sample(vc, sample_ws);
                const float2 sample_uv = cs_to_uv(float2(sample_cs.x, sample_cs.y));
                const int2 sample_px(kjb_cvt_i32(floor(sample_uv.x * output_tex_size.x / 2.0f)), kjb_cvt_i32(floor(sample_uv.y * output_tex_size.y / 2.0f)));
                sample_offset = int2(sample_px.x - half_px.x, sample_px.y - half_px.y);
            }
            const int2 sample_px = half_px + sample_offset;
            float rejection_bias = 1;
            const float3 sample_normal_vs = half_view_normal_tex.load(sample_px).xyz();
            float pdf0_mult = 1, pdf1_mult = 1;
            const float bent_pdf_ndotl_fix = 1;
            // USE_RESTIR
            const uint4 reservoir_raw = restir_reservoir_tex.load_u(sample_px);
            const Reservoir1spp r = Reservoir1spp::from_raw(uint2(reservoir_raw.x, reservoir_raw.y));
            const int2 spx = reservoir_payload_to_px(r.payload);
            const RtrRestirRayOrigin sample_origin = RtrRestirRayOrigin::from_raw(restir_ray_orig_tex.load(spx));
            const float3 sample_origin_ws = sample_origin.ray_origin_eye_offset_ws + eye;
            const float sample_roughness = sample_origin.roughness;
            if (reservoir_raw.x == 0 || sample_roughness > gbuffer.roughness * 2) continue;
            const float4 restir_ray = restir_ray_tex.load(spx);
            const float3 sample_hit_ws = restir_ray.xyz() + sample_origin_ws;
            const float3 sample_origin_vs = position_world_to_view(vc, sample_origin_ws);
            const float4 restir_irr = restir_irradiance_tex.load(spx);
            const float3 sample_radiance = restir_irr.xyz();
            const float sample_ray_pdf = restir_ray.w;
            const float neighbor_sampling_pdf = 1.0f / r.W;
            const float3 center_to_hit_vs = position_world_to_view(vc, sample_hit_ws) - lerp(refl_ray_origin_vs, sample_origin_vs, RTR_NEIGHBOR_RAY_ORIGIN_CENTER_BIAS);
            const float sample_cos_theta = rtr_decode_cos_theta_from_fp16(restir_irr.w);
            const float center_to_hit_dist = length(center_to_hit_vs);
            const float sample_to_hit_dist = length(sample_hit_ws - sample_origin_ws);
            {   // RTR_USE_BULLSHIT_TO_FIX_EDGE_HALOS
                const float center_to_hit_dist_wat_i_dont_even = length(position_world_to_view(vc, sample_hit_ws)
                    - lerp(refl_ray_origin_vs, sample_origin_vs, lerp(1.0f, RTR_NEIGHBOR_RAY_ORIGIN_CENTER_BIAS, 0.4f * min(1.0f, 3 * sqrt(gbuffer.roughness)))));
                pdf0_mult *= max(1e-5f, pow(center_to_hit_dist_wat_i_dont_even / sample_to_hit_dist, 2.0f));
                pdf1_mult *= max(1.0f, pow(center_to_hit_dist / sample_to_hit_dist, 2.0f));
            }
            const float3 wi = normalize(mul(direction_view_to_world(vc, center_to_hit_vs), tangent_to_world));
            if (wi.z < 1e-5f) continue;
            rejection_bias *= dot(normal_vs, sample_normal_vs) > 0.7f ? 1.0f : 0.0f;
            {   // depth-based rejection
                const float depth_diff = abs(refl_ray_origin_vs.z - sample_origin_vs.z) / max(1e-10f, kernel_size_ws);
                rejection_bias *= exp2(-max(0.3f, normal_vs.z) * depth_diff * depth_diff);
            }
            const float3 surface_offset = sample_origin_vs - refl_ray_origin_vs;
            // USE_APPROXIMATE_SAMPLE_SHADOWING
            if (dot(center_to_hit_vs, normal_vs) * 0.2f / length(center_to_hit_vs) < dot(surface_offset, normal_vs) / length(surface_offset)) rejection_bias *= is_center_sample ? 1.0f : 0.0f;
            const BrdfValue spec = specular_brdf.evaluate(wo, wi);
            const float spec_weight = spec.pdf * step(0.0f, wi.z);
            float contrib_wt = 0;
            {
                const float cos_theta = normalize(wo + wi).z;
                const float bent_cos_theta = min(sample_cos_theta, cos_theta * 1.25f);
                const float sample_ray_ndf = SpecularBrdf::ggx_ndf(a2, bent_cos_theta);
                const float center_ndf = SpecularBrdf::ggx_ndf(a2, cos_theta);
                const float bent_sample_pdf0 = spec.pdf * sample_ray_ndf / center_ndf;
                const float pdf_lerp_t = smoothstep(0.4f, 0.7f, sqrt(gbuffer.roughness)) * smoothstep(0.0f, 0.1f, ray_len_avg / eye_to_surf_dist);
                const float3 pdfs[2] = {
                    float3(min(bent_sample_pdf0, RTR_RESTIR_MAX_PDF_CLAMP) * bent_pdf_ndotl_fix, neighbor_sampling_pdf * pdf0_mult, 1 - pdf_lerp_t),
                    float3(min(spec.pdf, RTR_RESTIR_MAX_PDF_CLAMP), neighbor_sampling_pdf * pdf1_mult, pdf_lerp_t)};
                for (uint pdf_i = 0; pdf_i < 2; ++pdf_i) {
                    const float bent_sample_pdf = pdfs[pdf_i].x, nsp = pdfs[pdf_i].y, pdf_influence = pdfs[pdf_i].z;
                    const float mis_weight = max(1e-4f, spec.pdf / (sample_ray_pdf + spec.pdf));
                    contrib_wt = rejection_bias * mis_weight * max(1e-10f, spec_weight / bent_sample_pdf);
                    contrib_accum += float4(sample_radiance * bent_sample_pdf / nsp * spec.value_over_pdf, 1) * contrib_wt * pdf_influence;
                }
            }
            ray_len_accum += exponential_squish(surf_to_hit_dist, ray_squish_scale) * contrib_wt;
            sample_radius_accum += 1.0f - RADIUS_INC_ON_FAIL;
        }
        const float contrib_norm_factor = max(1e-14f, contrib_accum.w);
        float3 rgb = contrib_accum.xyz() / contrib_norm_factor;
        ray_len_accum /= contrib_norm_factor;
        const SpecularBrdfEnergyPreservation brdf_lut = SpecularBrdfEnergyPreservation::from_brdf_ndotv(g, specular_brdf, wo.z);
        rgb = rgb / brdf_lut.preintegrated_reflection;          // !RTR_RENDER_SCALED_BY_FG
        rgb = rgb * brdf_lut.preintegrated_reflection_mult;
        ray_len_accum = exponential_unsquish(ray_len_accum, ray_squish_scale);
        output_tex.store(px, float4(rgb, 0));
        ray_len_output_tex.store(px, float4(ray_len_accum, ray_len_avg, 0, 0));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ R5: rtr/temporal_filter.hlsl:36-259
int kjb_pass_rtr_temporal(kjb_context* ctx, const kjb_rtr_temporal_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img input_tex(a->input_tex), history_tex(a->history_tex), depth_tex(a->depth_tex), ray_len_tex(a->ray_len_tex), reprojection_tex(a->reprojection_tex),
        invalidity_tex(a->refl_restir_invalidity_tex), gbuffer_tex(a->gbuffer_tex), output_tex(a->output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    const float ped = g.fc.pre_exposure_delta;
    const float4 history_mult(ped, ped, ped, 1);
    const float3 eye = get_eye_position(vc), prev_eye = get_prev_eye_position(vc);
    const float2 texSize(output_tex_size.x, output_tex_size.y);
    // inc/image.hlsl:85-170 image_sample_catmull_rom_5tap(sampler_lnc, IdentityImageRemap)
    auto catmull_rom_5tap = [&](const Img& tex, float2 suv) {
        auto smp = [&](float2 p) { return tex.sample_bilinear_clamp(p); };
        float2 samplePos = suv * texSize;
        float2 texPos1 = floor(samplePos - 0.5f) + 0.5f;
        float2 f = samplePos - texPos1;
        float2 w0 = f * (-0.5f + f * (1.0f - 0.5f * f));
        float2 w1 = 1.0f + f * f * (-2.5f + 1.5f * f);
        float2 w2 = f * (0.5f + f * (2.0f - 1.5f * f));
        float2 w3 = f * f * (-0.5f + 0.5f * f);
        float2 w12 = w1 + w2;
        float2 offset12 = w2 / (w1 + w2);
        float2 texPos0 = texPos1 - 1.0f, texPos3 = texPos1 + 2.0f, texPos12 = texPos1 + offset12;
        texPos0 = texPos0 / texSize; texPos3 = texPos3 / texSize; texPos12 = texPos12 / texSize;
        float4 result(0.0f);
        result += smp(float2(texPos12.x, texPos0.y)) * w12.x * w0.y;
        result += smp(float2(texPos0.x, texPos12.y)) * w0.x * w12.y;
        result += smp(float2(texPos12.x, texPos12.y)) * w12.x * w12.y;
        result += smp(float2(texPos3.x, texPos12.y)) * w3.x * w12.y;
        result += smp(float2(texPos12.x, texPos3.y)) * w12.x * w3.y;
        return result / (w12.x * w0.y + w0.x * w12.y + w12.x * w12.y + w3.x * w12.y + w12.x * w3.y);
    };
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y);
        const float4 center = linear_rgb_to_crunched_luma_chroma(input_tex.load(px));
        const float refl_ray_length = clamp(ray_len_tex.load(px).x, 0.0f, 1e3f);
        const float2 uv = get_uv(px, output_tex_size);
        const float center_depth = depth_tex.load(px).x;
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_depth(vc, uv, center_depth);
        const float3 reflector_vs = view_ray_context.ray_hit_vs();
        const float3 reflection_hit_vs = reflector_vs + view_ray_context.ray_dir_vs() * refl_ray_length;
        const float4 reflection_hit_cs = mul(vc.view_to_sample, float4(reflection_hit_vs, 1));
        const float4 prev_hit_cs = mul(vc.clip_to_prev_clip, reflection_hit_cs);
        float2 hit_prev_uv = cs_to_uv(float2(prev_hit_cs.x, prev_hit_cs.y) / prev_hit_cs.w);
        const float4 prev_reflector_cs = mul(vc.clip_to_prev_clip, view_ray_context.ray_hit_cs);
        const float2 reflector_prev_uv = cs_to_uv(float2(prev_reflector_cs.x, prev_reflector_cs.y) / prev_reflector_cs.w);
        const float4 reproj = reprojection_tex.load(px);
        const float reflector_move_rate = min(1.0f, length(reproj.xy()) / length(reflector_prev_uv - uv));
        hit_prev_uv = lerp(uv, hit_prev_uv, reflector_move_rate);
        const uint quad_reproj_valid_packed = kjb_cvt_u32(reproj.z * 15.0f + 0.5f);
        float4 history0(0.0f); float history0_valid = 1;
        if (0 == quad_reproj_valid_packed) {
            history0_valid = 0;
        } else if (15 == quad_reproj_valid_packed) {
            history0 = max(float4(0.0f), catmull_rom_5tap(history_tex, uv + reproj.xy())) * history_mult;
        } else {
            const float4 quad_reproj_valid((quad_reproj_valid_packed & 1u) ? 1.0f : 0.0f, (quad_reproj_valid_packed & 2u) ? 1.0f : 0.0f, (quad_reproj_valid_packed & 4u) ? 1.0f : 0.0f, (quad_reproj_valid_packed & 8u) ? 1.0f : 0.0f);
            const Bilinear bilinear = get_bilinear_filter(uv + reproj.xy(), texSize);
            const int ox = kjb_cvt_i32(bilinear.origin.x), oy = kjb_cvt_i32(bilinear.origin.y);
            const float4 s00 = history_tex.load(ox, oy) * history_mult, s10 = history_tex.load(ox + 1, oy) * history_mult, s01 = history_tex.load(ox, oy + 1) * history_mult, s11 = history_tex.load(ox + 1, oy + 1) * history_mult;
            const float4 weights = get_bilinear_custom_weights(bilinear, quad_reproj_valid);
            if (dot(weights, float4(1.0f)) > 1e-5f) history0 = apply_bilinear_custom_weights(s00, s10, s01, s11, weights);
            else history0 = (s00 + s10 + s01 + s11) / 4.0f;
        }
        history0 = linear_rgb_to_crunched_luma_chroma(history0);
        const float4 history1 = linear_rgb_to_crunched_luma_chroma(history_tex.sample_bilinear_clamp(hit_prev_uv) * history_mult);
        const float history1_valid = quad_reproj_valid_packed == 15 ? 1.0f : 0.0f;
        float4 vsum(0.0f), vsum2(0.0f); float wsum = 0;
        for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
            const int2 sample_px = px + int2(xx, yy);
            const float sample_depth = depth_tex.load(sample_px).x;
            const float4 neigh = linear_rgb_to_crunched_luma_chroma(input_tex.load(sample_px));
            float w = 1;
            w *= exp2(-200.0f * abs(center_depth / sample_depth - 1.0f));
            vsum = mad(neigh, w, vsum); vsum2 = mad(neigh * neigh, w, vsum2); wsum += w;
        }
        const float4 ex = vsum / wsum, ex2 = vsum2 / wsum;
        const float4 dev = sqrt(max(float4(0.0f), ex2 - ex * ex));
        const GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(px));
        const float restir_invalidity = invalidity_tex.load(x / 2, y / 2).x;
        const float n_deviations = lerp(reproj.z > 0 ? 2.0f : 1.25f, 0.625f, restir_invalidity);
        float wo_similarity;
        {
            const float3 current_wo = normalize(view_ray_context.ray_hit_ws() - eye);
            const float3 prev_wo = normalize(view_ray_context.ray_hit_ws() - prev_eye);
            const float clamped_roughness = max(0.1f, gbuffer.roughness);
            wo_similarity = pow(saturate(SpecularBrdf::ggx_ndf_0_1(clamped_roughness * clamped_roughness, dot(current_wo, prev_wo))), 32.0f);
        }
        const float h0diff = length((history0.xyz() - ex.xyz()) / dev.xyz());
        const float h1diff = length((history1.xyz() - ex.xyz()) / dev.xyz());
        float h0_score = 1.0f * smoothstep(0.0f, 0.5f, sqrt(gbuffer.roughness)) * lerp(wo_similarity, 1.0f, sqrt(gbuffer.roughness));
        float h1_score = (1 - h0_score) * lerp(1.0f, smoothstep(0.0f, 1.0f, h0diff - h1diff), smoothstep(0.0f, 0.15f, sqrt(gbuffer.roughness)));
        h0_score *= history0_valid; h1_score *= history1_valid;
        const float score_sum = h0_score + h1_score;
        h0_score /= score_sum;
        h1_score = 1 - h0_score;
        if (!(h0_score < 1.001f)) { h0_score = 1; h1_score = 0; }
        float4 clamped_history0 = history0, clamped_history1 = history1;
        {
            const float3 c0 = soft_color_clamp(center.xyz(), history0.xyz(), ex.xyz(), dev.xyz() * n_deviations);
            const float3 c1 = soft_color_clamp(center.xyz(), history1.xyz(), ex.xyz(), dev.xyz() * n_deviations);
            clamped_history0 = float4(c0, history0.w); clamped_history1 = float4(c1, history1.w);
        }
        const float4 clamped_history = clamped_history0 * h0_score + clamped_history1 * h1_score;
        const float max_sample_count = 16;
        const float current_sample_count = clamped_history.w * saturate(h0_score * history0_valid + h1_score * history1_valid);
        float4 res = lerp(clamped_history, center, 1.0f / (1.0f + min(max_sample_count, current_sample_count * lerp(wo_similarity, 1.0f, 0.5f))));
        res.w = min(current_sample_count, max_sample_count) + 1;
        res = crunched_luma_chroma_to_linear_rgb(res);
        output_tex.store(px, max(float4(0.0f), res));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ R6: rtr/spatial_cleanup.hlsl:19-65
int kjb_pass_rtr_cleanup(kjb_context* ctx, const kjb_rtr_cleanup_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img input_tex(a->input_tex), depth_tex(a->depth_tex), geometric_normal_tex(a->geometric_normal_tex), output_tex(a->output_tex);
    if (!a->spatial_resolve_offsets) { ctx->last_error = "reflection cleanup: spatial_resolve_offsets is null"; return 1; }
    const int32_t* offs = a->spatial_resolve_offsets;
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y);
        const float4 center = input_tex.load(px);
        const float center_depth = depth_tex.load(px).x;
        const float center_sample_count = center.w;
        const float min_sample_count = 8;
        if (center_sample_count >= min_sample_count || center_depth == 0.0f) { output_tex.store(px, center); continue; }
        const float3 center_normal_vs = geometric_normal_tex.load(px).xyz() * 2.0f - 1.0f;
        const float filter_radius_ss = 0.5f * vc.view_to_clip.m[5] / -depth_to_view_z(vc, center_depth);
        const uint filter_idx = kjb_cvt_u32(clamp(filter_radius_ss * 7.0f, 0.0f, 7.0f));
        float3 vsum(0.0f); float wsum = 0;
        int sc = kjb_cvt_i32(8.0f - center_sample_count / 2.0f); sc = sc < 2 ? 2 : (sc > 8 ? 8 : sc);   // clamp(int(..), min/4, min) with float bounds 2, 8
        const uint sample_count = uint(sc);
        const int kernel_scale = center_sample_count < 4 ? 2 : 1;
        const uint px_idx_in_quad = (((uint(x) & 1u) | (uint(y) & 1u) * 2u) + g.fc.frame_index) & 3u;
        for (uint sample_i = 0; sample_i < sample_count; ++sample_i) {
            const int32_t* o = offs + 4 * ((px_idx_in_quad * 16 + sample_i) + 64 * filter_idx);
            const int2 sample_px(x + kernel_scale * o[0], y + kernel_scale * o[1]);
            const float4 nl = input_tex.load(sample_px);
            const float3 neigh = sqrt(nl.xyz());   // linear_rgb_to_crunched_rgb
            const float sample_depth = depth_tex.load(sample_px).x;
            const float3 sample_normal_vs = geometric_normal_tex.load(sample_px).xyz() * 2.0f - 1.0f;
            float w = 1;
            w *= exp2(-50.0f * abs(center_normal_vs.z * (center_depth / sample_depth - 1.0f)));
            const float dp = saturate(dot(center_normal_vs, sample_normal_vs));
            w *= dp * dp * dp;
            vsum = mad(neigh, w, vsum); wsum += w;
        }
        const float3 v = vsum / wsum;
        output_tex.store(px, float4(v * v, 1));   // crunched_rgb_to_linear_rgb
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"
}  // namespace kjo
