// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// CPU restatement of kajiya's ray-traced diffuse GI passes, one function per render-graph pass
// (crates/lib/kajiya/src/renderers/rtdgi.rs).  Paths below are relative to /root/reference/assets/shaders/.
#include "kj_ircache_lookup.h"

namespace kjo {

namespace {

const float SKY_DIST = 1e4f;                       // rtdgi/diffuse_trace_common.inc.hlsl:16
const float RESTIR_TEMPORAL_M_CLAMP = 20.0f;       // rtdgi/rtdgi_restir_settings.hlsl:2
const float RESTIR_RESERVOIR_W_CLAMP = 10.0f;      // :5
const float SSGI_NEAR_FIELD_RADIUS = 80.0f;        // rtdgi/near_field_settings.hlsl:2

inline bool is_rtdgi_validation_frame(const Globals& g) { return g.fc.frame_index % 3 == 0; }   // rtdgi_restir_settings.hlsl:40-46
inline bool is_rtdgi_tracing_frame(const Globals& g) { return !is_rtdgi_validation_frame(g); }
inline int2 reservoir_payload_to_px(uint payload) { return int2(int(payload & 0xffff), int(payload >> 16)); }
inline float rtr_encode_cos_theta_for_fp16(float x) { return 1 - x; }   // rtr/rtr_settings.hlsl:52-55

// rtdgi/rtdgi_common.hlsl:12-39
struct TemporalReservoirOutput {
    float depth; float3 ray_hit_offset_ws; float luminance; float3 hit_normal_ws;
    static TemporalReservoirOutput from_raw(uint4 raw) {
        float2 a = unpack_2x16f_uint(raw.y), b = unpack_2x16f_uint(raw.z);
        TemporalReservoirOutput r; r.depth = asfloat(raw.x); r.ray_hit_offset_ws = float3(a.x, a.y, b.x); r.luminance = b.y;
        r.hit_normal_ws = unpack_normal_11_10_11(asfloat(raw.w));
        return r;
    }
    uint4 as_raw() const {
        return uint4(asuint(depth), pack_2x16f_uint(float2(ray_hit_offset_ws.x, ray_hit_offset_ws.y)),
                     pack_2x16f_uint(float2(ray_hit_offset_ws.z, luminance)), asuint(pack_normal_11_10_11(hit_normal_ws)));
    }
};

struct TraceResult { float3 out_value, hit_normal_ws; float hit_t, pdf; bool is_hit; };

// rtdgi/diffuse_trace_common.inc.hlsl:38-221  (USE_WORLD_RADIANCE_CACHE 0; ircache lookup handled by the caller-supplied hook)
TraceResult do_the_thing(const kjb_context& ctx, const Img& depth_tex, const Img& reprojected_gi_tex, const Img& sky_cube_tex,
                         float4 gbuffer_tex_size, uint2 px, float3 normal_ws, uint& rng, Ray outgoing_ray, const IrcacheBufs& ircache) {
    const Globals& g = ctx.g; const kjb_view_constants& vc = g.fc.view_constants;
    float3 total_radiance(0.0f);
    float3 hit_normal_ws = -outgoing_ray.dir;
    float hit_t = outgoing_ray.tmax;
    float pdf = max(0.0f, 1.0f / (dot(normal_ws, outgoing_ray.dir) * 2 * M_PI_F));

    const float reflected_cone_spread_angle = 0.03f;
    RayCone ray_cone = RayCone::from_spread_angle(pixel_cone_spread_angle_from_image_height(vc, gbuffer_tex_size.y * 0.5f))
        .propagate(reflected_cone_spread_angle, length(outgoing_ray.origin - get_eye_position(vc)));

    const GbufferPathVertex primary_hit = gbuffer_raytrace(ctx.scene, g, outgoing_ray, ray_cone, 1, false);

    if (primary_hit.is_hit) {
        hit_t = primary_hit.ray_t;
        GbufferData gbuffer = gbuffer_unpack(primary_hit.gbuffer_packed);
        hit_normal_ws = gbuffer.normal;

        const float3 primary_hit_cs = position_world_to_sample(vc, primary_hit.position);
        const float2 primary_hit_uv = cs_to_uv(float2(primary_hit_cs.x, primary_hit_cs.y));
        const float primary_hit_screen_depth = depth_tex.sample_nearest_clamp(primary_hit_uv).x;
        bool is_on_screen = abs(primary_hit_cs.x) < 1.0f && abs(primary_hit_cs.y) < 1.0f
            && inverse_depth_relative_diff(primary_hit_cs.z, primary_hit_screen_depth) < 5e-3f;

        float4 reprojected_radiance(0.0f);
        if (is_on_screen) {
            reprojected_radiance = reprojected_gi_tex.sample_nearest_clamp(primary_hit_uv) * g.fc.pre_exposure_delta;
            is_on_screen = reprojected_radiance.w > 0;
        }

        gbuffer.roughness = lerp(gbuffer.roughness, 1.0f, 0.5f);   // ROUGHNESS_BIAS
        const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
        const float3 wo = mul(-outgoing_ray.dir, tangent_to_world);
        const LayeredBrdf brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z);

        // Sun
        float3 sun_radiance = sun_color_in_direction(g, sun_direction(g));
        if (any_nonzero(sun_radiance)) {
            const float3 to_light_norm = sample_sun_direction(g, blue_noise_for_pixel(g, px, rng).xy(), false);
            const bool is_shadowed = rt_is_shadowed(ctx.scene, primary_hit.position, to_light_norm, 1e-4f, SKY_DIST);
            const float3 wi = mul(to_light_norm, tangent_to_world);
            const float3 brdf_value = brdf.evaluate(wo, wi) * max(0.0f, wi.z);
            const float3 light_radiance = is_shadowed ? float3(0.0f) : sun_radiance;
            total_radiance += brdf_value * light_radiance;
        }

        total_radiance += gbuffer.emissive;   // USE_EMISSIVE

        if (is_on_screen) {                    // USE_SCREEN_GI_REPROJECTION
            total_radiance += reprojected_radiance.xyz() * gbuffer.albedo;
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
                        const bool is_shadowed = rt_is_shadowed(ctx.scene, shadow_ray_origin, to_light_norm_ws, 1e-3f, sqrt(dist_to_light2) - 2e-3f);
                        const float3 bounce_albedo = lerp(gbuffer.albedo, float3(1.0f), 0.04f);
                        const float3 brdf_value = bounce_albedo * to_psa_metric / M_PI_F;
                        float3 radiance(tl.radiance[0], tl.radiance[1], tl.radiance[2]);
                        total_radiance += !is_shadowed ? (radiance * brdf_value / ls.pdf) : float3(0.0f);
                    }
                }
            }
            {   // USE_IRCACHE (contributes 0 when no cache is bound: kjb_ircache_bindings.meta_buf == NULL)
                const float3 gi = ircache_lookup(g, ircache, outgoing_ray.origin, primary_hit.position, gbuffer.normal, 1, rng, false);
                total_radiance += gi * gbuffer.albedo;
            }
        }
    } else {
        total_radiance += sky_cube_tex.sample_cube(outgoing_ray.dir).xyz();
    }

    TraceResult result;
    result.out_value = total_radiance; result.hit_t = hit_t; result.hit_normal_ws = hit_normal_ws; result.pdf = pdf; result.is_hit = primary_hit.is_hit;
    return result;
}

inline float3 rtdgi_candidate_ray_dir(const Globals& g, uint2 px, const float3x3& tangent_to_world) {   // rtdgi/candidate_ray_dir.hlsl:1-24
    float2 urand = blue_noise_for_pixel(g, px, g.fc.frame_index).xy();
    float3 wi = uniform_sample_hemisphere(urand);
    return mul(tangent_to_world, wi);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ D1: rtdgi/fullres_reproject.hlsl:29-77
int kjb_pass_rtdgi_reproject(kjb_context* ctx, const kjb_rtdgi_reproject_args* a) {
    Img input_tex(a->input_tex), reprojection_tex(a->reprojection_tex), output_tex(a->output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        int2 px(x, y);
        float2 uv = get_uv(px, output_tex_size);
        float4 reproj = reprojection_tex.load(px);
        float2 prev_uv = uv + reproj.xy();
        uint quad_reproj_valid_packed = uint(reproj.z * 15.0f + 0.5f);

        // GatherBlue(sampler_nnc, uv + 0.5 * sign(prev_uv) * output_tex_size.zw): the 2x2 footprint of a bilinear
        // fetch at that location; component order of Gather is (0,1),(1,1),(1,0),(0,0) — only all()==15 is used.
        float2 guv = uv + 0.5f * float2(sign(prev_uv.x), sign(prev_uv.y)) * float2(output_tex_size.z, output_tex_size.w);
        float gx = guv.x * float(W) - 0.5f, gy = guv.y * float(H) - 0.5f;
        int gx0 = int(floor(gx)); int gy0 = int(floor(gy));
        bool all_neigh_valid = true;
        for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) {
            int sx = gx0 + i, sy = gy0 + j;
            sx = sx < 0 ? 0 : (sx >= W ? W - 1 : sx); sy = sy < 0 ? 0 : (sy >= H ? H - 1 : sy);
            if (uint(reprojection_tex.load(sx, sy).z * 15.0f + 0.5f) != 15u) all_neigh_valid = false;
        }

        float4 history(0.0f);
        if (0 == quad_reproj_valid_packed) {
        } else if (15 == quad_reproj_valid_packed) {
            if (all_neigh_valid) {
                // image_sample_catmull_rom (inc/image.hlsl:42-79) with identity remap
                float2 pixel = prev_uv * float2(float(W), float(H)) + 0.5f;
                float2 frc = frac(pixel);
                int2 ipixel(kjb_cvt_i32(pixel.x) - 1, kjb_cvt_i32(pixel.y) - 1);
                auto cubic = [](float4 A, float4 B, float4 C, float4 D, float t) {   // inc/curve.hlsl cubic_hermite
                    float t2 = t * t, t3 = t * t * t;
                    float4 aa = -A / 2.0f + (3.0f * B) / 2.0f - (3.0f * C) / 2.0f + D / 2.0f;
                    float4 bb = A - (5.0f * B) / 2.0f + 2.0f * C - D / 2.0f;
                    float4 cc = -A / 2.0f + C / 2.0f;
                    float4 dd = B;
                    return aa * t3 + bb * t2 + cc * t + dd;
                };
                float4 rows[4];
                for (int j = 0; j < 4; ++j) {
                    float4 c0 = input_tex.load(ipixel.x - 1, ipixel.y - 1 + j), c1 = input_tex.load(ipixel.x, ipixel.y - 1 + j),
                           c2 = input_tex.load(ipixel.x + 1, ipixel.y - 1 + j), c3 = input_tex.load(ipixel.x + 2, ipixel.y - 1 + j);
                    rows[j] = cubic(c0, c1, c2, c3, frc.x);
                }
                history = max(float4(0.0f), cubic(rows[0], rows[1], rows[2], rows[3], frc.y));
            } else {
                history = input_tex.sample_bilinear_clamp(prev_uv);
            }
        } else {
            float4 quad_reproj_valid((quad_reproj_valid_packed & 1u) ? 1.0f : 0.0f, (quad_reproj_valid_packed & 2u) ? 1.0f : 0.0f,
                                     (quad_reproj_valid_packed & 4u) ? 1.0f : 0.0f, (quad_reproj_valid_packed & 8u) ? 1.0f : 0.0f);
            const Bilinear bilinear = get_bilinear_filter(prev_uv, float2(output_tex_size.x, output_tex_size.y));
            int ox = kjb_cvt_i32(bilinear.origin.x), oy = kjb_cvt_i32(bilinear.origin.y);
            float4 s00 = input_tex.load(ox, oy), s10 = input_tex.load(ox + 1, oy), s01 = input_tex.load(ox, oy + 1), s11 = input_tex.load(ox + 1, oy + 1);
            float4 weights = get_bilinear_custom_weights(bilinear, quad_reproj_valid);
            if (dot(weights, float4(1.0f)) > 1e-5f) history = apply_bilinear_custom_weights(s00, s10, s01, s11, weights);
        }
        output_tex.store(px, history);
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D3: rtdgi/diffuse_validate.rgen.hlsl:46-111
int kjb_pass_rtdgi_validate(kjb_context* ctx, const kjb_rtdgi_validate_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img half_view_normal_tex(a->half_view_normal_tex), depth_tex(a->depth_tex), reprojected_gi_tex(a->reprojected_gi_tex), reservoir_tex(a->reservoir_tex),
        reservoir_ray_history_tex(a->reservoir_ray_history_tex), sky_cube_tex(a->sky_cube_tex), irradiance_history_tex(a->irradiance_history_tex),
        ray_orig_history_tex(a->ray_orig_history_tex), out_tex(a->rt_history_invalidity_out_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = out_tex.w(), H = out_tex.h();
    const int2 hi_px_offset = halfres_subsample_offset(g.fc.frame_index);
    const IrcacheBufs ircache = IrcacheBufs::from(a->ircache);
    pass_pixels(ctx, W, H, ircache.bound(), [&](int x, int y) { {
        const int2 px(x, y); const int2 hi_px = px * 2 + hi_px_offset;
        if (0.0f == depth_tex.load(hi_px).x) { out_tex.store(px, float4(1.0f)); return; }
        float invalidity = 0.0f;
        if (is_rtdgi_validation_frame(g)) {
            const float3 normal_vs = half_view_normal_tex.load(px).xyz();
            const float3 normal_ws = direction_view_to_world(vc, normal_vs);
            const float3 prev_ray_orig = ray_orig_history_tex.load(px).xyz();
            const float3 prev_hit_pos = reservoir_ray_history_tex.load(px).xyz() + prev_ray_orig;
            const float4 prev_radiance_packed = irradiance_history_tex.load(px);
            const float3 prev_radiance = max(float3(0.0f), prev_radiance_packed.xyz());

            Ray prev_ray; prev_ray.dir = normalize(prev_hit_pos - prev_ray_orig); prev_ray.origin = prev_ray_orig; prev_ray.tmin = 0; prev_ray.tmax = SKY_DIST;
            uint rng = hash3(uint(x), uint(y), 0);
            TraceResult result = do_the_thing(*ctx, depth_tex, reprojected_gi_tex, sky_cube_tex, gbuffer_tex_size, uint2(x, y), normal_ws, rng, prev_ray, ircache);
            const float3 new_radiance = max(float3(0.0f), result.out_value);

            const float rad_diff = length(abs(prev_radiance - new_radiance) / max(float3(1e-3f), prev_radiance + new_radiance));
            invalidity = smoothstep(0.1f, 0.5f, rad_diff / length(float3(1.0f)));
            const float prev_hit_dist = length(prev_hit_pos - prev_ray_orig);
            if (abs(result.hit_t - prev_hit_dist) / (prev_hit_dist + prev_hit_dist) < 0.2f) {
                irradiance_history_tex.store(px, float4(new_radiance, prev_radiance_packed.w));
                uint4 raw = reservoir_tex.load_u(px);
                Reservoir1spp r = Reservoir1spp::from_raw(uint2(raw.x, raw.y));
                const float lum_old = sRGB_to_luminance(prev_radiance);
                const float lum_new = sRGB_to_luminance(new_radiance);
                r.M *= clamp(lum_old / max(1e-8f, lum_new), 0.03f, 1.0f);
                const float allowed_luminance_increment = 10.0f;
                r.W *= clamp(lum_old / max(1e-8f, lum_new) * allowed_luminance_increment, 0.01f, 1.0f);
                uint2 rr = r.as_raw();
                reservoir_tex.store_u(x, y, uint4(rr.x, rr.y, 0, 0));
            }
        }
        out_tex.store(px, float4(invalidity));
    } });
    return 0;
}

// ------------------------------------------------------------------ D4: rtdgi/trace_diffuse.rgen.hlsl:49-120
int kjb_pass_rtdgi_trace(kjb_context* ctx, const kjb_rtdgi_trace_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img half_view_normal_tex(a->half_view_normal_tex), depth_tex(a->depth_tex), reprojected_gi_tex(a->reprojected_gi_tex), reprojection_tex(a->reprojection_tex),
        sky_cube_tex(a->sky_cube_tex), cand_irr(a->candidate_irradiance_out_tex), cand_normal(a->candidate_normal_out_tex), cand_hit(a->candidate_hit_out_tex),
        inv_in(a->rt_history_invalidity_in_tex), inv_out(a->rt_history_invalidity_out_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = cand_irr.w(), H = cand_irr.h();
    const int2 hi_px_offset = halfres_subsample_offset(g.fc.frame_index);
    const IrcacheBufs ircache = IrcacheBufs::from(a->ircache);
    pass_pixels(ctx, W, H, ircache.bound(), [&](int x, int y) { {
        const int2 px(x, y); const int2 hi_px = px * 2 + hi_px_offset;
        float depth = depth_tex.load(hi_px).x;
        if (0.0f == depth) {
            cand_irr.store(px, float4(0.0f)); cand_normal.store(px, float4(0, 0, 1, 0)); inv_out.store(px, float4(0.0f));
            return;
        }
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_biased_depth(vc, uv, depth);
        const float NEAR_FIELD_FADE_OUT_END = -view_ray_context.ray_hit_vs().z * (SSGI_NEAR_FIELD_RADIUS * gbuffer_tex_size.w * 0.5f);
        {   // RTDGI_INTERLEAVED_VALIDATION_ALWAYS_TRACE_NEAR_FIELD
            const float3 normal_vs = half_view_normal_tex.load(px).xyz();
            const float3 normal_ws = direction_view_to_world(vc, normal_vs);
            const float3x3 tangent_to_world = build_orthonormal_basis(normal_ws);
            const float3 outgoing_dir = rtdgi_candidate_ray_dir(g, uint2(x, y), tangent_to_world);
            Ray outgoing_ray; outgoing_ray.dir = outgoing_dir;
            outgoing_ray.origin = view_ray_context.biased_secondary_ray_origin_ws_with_normal(normal_ws);
            outgoing_ray.tmin = 0;
            outgoing_ray.tmax = is_rtdgi_tracing_frame(g) ? SKY_DIST : NEAR_FIELD_FADE_OUT_END;
            uint rng = hash3(uint(x), uint(y), g.fc.frame_index & 31);
            TraceResult result = do_the_thing(*ctx, depth_tex, reprojected_gi_tex, sky_cube_tex, gbuffer_tex_size, uint2(x, y), normal_ws, rng, outgoing_ray, ircache);
            if (!is_rtdgi_tracing_frame(g) && !result.is_hit) { result.out_value = float3(0.0f); result.hit_t = SKY_DIST; }
            const float3 hit_offset_ws = outgoing_ray.dir * result.hit_t;
            const float cos_theta = dot(normalize(outgoing_dir - view_ray_context.ray_dir_ws()), normal_ws);
            cand_irr.store(px, float4(result.out_value, rtr_encode_cos_theta_for_fp16(cos_theta)));
            cand_hit.store(px, float4(hit_offset_ws, result.pdf * (is_rtdgi_tracing_frame(g) ? 1.0f : -1.0f)));
            cand_normal.store(px, float4(direction_world_to_view(vc, result.hit_normal_ws), 0));
        }
        const float4 reproj = reprojection_tex.load(hi_px);
        const int2 reproj_px(kjb_cvt_i32(floor(float(x) + gbuffer_tex_size.x * reproj.x / 2 + 0.5f)), kjb_cvt_i32(floor(float(y) + gbuffer_tex_size.y * reproj.y / 2 + 0.5f)));
        inv_out.store(px, float4(inv_in.load(reproj_px).x));
    } });
    return 0;
}

// ------------------------------------------------------------------ D5: rtdgi/temporal_validity_integrate.hlsl:21-119
int kjb_pass_rtdgi_validity_integrate(kjb_context* ctx, const kjb_rtdgi_validity_integrate_args* a) {
    const Globals& g = ctx->g;
    Img input_tex(a->input_tex), history_tex(a->history_tex), reprojection_tex(a->reprojection_tex), half_depth_tex(a->half_depth_tex), output_tex(a->output_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    // The shader exchanges values between lanes of an 8x8 group (WaveReadLaneAt lane^2, ^16, ^1, ^8 with lane = x + 8*y inside
    // 32-wide waves => partners are pixel (x^2,y), (x,y^2), (x^1,y), (x,y^1); SURVEY.md H5, NVIDIA wave32 mapping).
    // Groups cover the image rounded up to 8, and out-of-image lanes run the same code on zero loads.
    const int PW = (W + 7) & ~7, PH = (H + 7) & ~7;
    std::vector<float> A(size_t(PW) * PH), B(size_t(PW) * PH), E0(size_t(PW) * PH), E1(size_t(PW) * PH);
    parallel_rows(PH, [&](int y) { for (int x = 0; x < PW; ++x) {
        float2 invalid_blurred(0.0f);
        const int k = 2;
        for (int yy = -k; yy <= k; ++yy) for (int xx = -k; xx <= k; ++xx) {
            float w = exp2(-0.1f * float(xx * xx + yy * yy));
            invalid_blurred = mad(float2(input_tex.load(x + xx, y + yy).x, 1), w, invalid_blurred);
        }
        invalid_blurred = invalid_blurred / invalid_blurred.y;
        A[size_t(y) * PW + x] = invalid_blurred.x;

        const float center_depth = half_depth_tex.load(x, y).x;
        float edge = 1;
        for (int yy = 0; yy <= k; ++yy) for (int xx = 1; xx <= k; ++xx) {
            const int2 sample_px(x * 2 + xx, y * 2 + yy);
            const int2 sample_px_half(x + xx / 2, y + yy / 2);
            const float4 reproj = reprojection_tex.load(sample_px);
            const float sample_depth = half_depth_tex.load(sample_px_half).x;
            if (reproj.w < 0 || inverse_depth_relative_diff(center_depth, sample_depth) > 0.1f) { edge = 0; break; }
            edge *= (reproj.z == 0 && sample_depth != 0) ? 1.0f : 0.0f;
        }
        E0[size_t(y) * PW + x] = edge;
    } }, ctx->num_threads);
    for (int y = 0; y < PH; ++y) for (int x = 0; x < PW; ++x) {
        B[size_t(y) * PW + x] = lerp(A[size_t(y) * PW + x], A[size_t(y) * PW + (x ^ 2)], 0.5f);
        E1[size_t(y) * PW + x] = max(E0[size_t(y) * PW + x], E0[size_t(y) * PW + (x ^ 1)]);
    }
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        float inv = lerp(B[size_t(y) * PW + x], B[size_t(y ^ 2) * PW + x], 0.5f);
        inv = smoothstep(0.0f, 1.0f, inv);
        float edge = max(E1[size_t(y) * PW + x], E1[size_t(y ^ 1) * PW + x]);
        inv += edge;
        inv = saturate(inv);

        const float4 reproj = reprojection_tex.load(x * 2, y * 2);
        const float2 reproj_px = float2(float(x), float(y)) + float2(gbuffer_tex_size.x, gbuffer_tex_size.y) * reproj.xy() / 2.0f + 0.5f;
        float history = 0;
        const int sample_count = 8;
        float ang_off = uint_to_u01_float(hash3(uint(x), uint(y), g.fc.frame_index)) * M_PI_F * 2;
        for (uint sample_i = 0; sample_i < uint(sample_count); ++sample_i) {
            float ang = (float(sample_i) + ang_off) * GOLDEN_ANGLE;
            float radius = float(sample_i) * 1.0f;
            float2 sample_offset = float2(cos(ang), sin(ang)) * radius;
            const int2 sample_px(kjb_cvt_i32(reproj_px.x + sample_offset.x), kjb_cvt_i32(reproj_px.y + sample_offset.y));
            history += history_tex.load(sample_px).x;
        }
        history /= float(sample_count);
        output_tex.store(x, y, float4(max(history * 0.75f, inv), input_tex.load(x, y).x, 0, 0));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D6: rtdgi/restir_temporal.hlsl:83-422
static int2 get_rpx_offset(uint sample_i, uint frame_index) {   // :64-81
    const int offsets[4][2] = {{-1, -1}, {1, 1}, {-1, 1}, {1, -1}};
    const int* a = offsets[frame_index & 3]; const int* b = offsets[(sample_i + (frame_index ^ 1)) & 3];
    return sample_i == 0 ? int2(0, 0) : int2(a[0] + b[0], a[1] + b[1]);
}

int kjb_pass_rtdgi_restir_temporal(kjb_context* ctx, const kjb_rtdgi_restir_temporal_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img half_view_normal_tex(a->half_view_normal_tex), depth_tex(a->depth_tex), candidate_radiance_tex(a->candidate_radiance_tex), candidate_normal_tex(a->candidate_normal_tex),
        candidate_hit_tex(a->candidate_hit_tex), radiance_history_tex(a->radiance_history_tex), ray_orig_history_tex(a->ray_orig_history_tex), ray_history_tex(a->ray_history_tex),
        reservoir_history_tex(a->reservoir_history_tex), reprojection_tex(a->reprojection_tex), hit_normal_history_tex(a->hit_normal_history_tex),
        candidate_history_tex(a->candidate_history_tex), rt_invalidity_tex(a->rt_invalidity_tex);
    Img radiance_out_tex(a->radiance_out_tex), ray_orig_output_tex(a->ray_orig_output_tex), ray_output_tex(a->ray_output_tex), hit_normal_output_tex(a->hit_normal_output_tex),
        reservoir_out_tex(a->reservoir_out_tex), candidate_out_tex(a->candidate_out_tex), temporal_reservoir_packed_tex(a->temporal_reservoir_packed_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = radiance_out_tex.w(), H = radiance_out_tex.h();
    const int2 hi_px_offset = halfres_subsample_offset(g.fc.frame_index);
    const uint frame_index = g.fc.frame_index;

    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y); const int2 hi_px = px * 2 + hi_px_offset;
        float depth = depth_tex.load(hi_px).x;
        if (0.0f == depth) {
            radiance_out_tex.store(px, float4(0, 0, 0, -SKY_DIST));
            hit_normal_output_tex.store(px, float4(0.0f));
            reservoir_out_tex.store_u(x, y, uint4(0, 0, 0, 0));
            continue;
        }
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_biased_depth(vc, uv, depth);
        const float3 normal_vs = half_view_normal_tex.load(px).xyz();
        const float3 normal_ws = direction_view_to_world(vc, normal_vs);
        const float3 refl_ray_origin_ws = view_ray_context.biased_secondary_ray_origin_ws_with_normal(normal_ws);

        const float3 hit_offset_ws = candidate_hit_tex.load(px).xyz();
        float3 outgoing_dir = normalize(hit_offset_ws);

        uint rng = hash3(uint(x), uint(y), frame_index);

        int2 src_px_sel = px; (void)src_px_sel;
        float3 radiance_sel(0.0f), ray_orig_sel_ws(0.0f), ray_hit_sel_ws(1.0f), hit_normal_sel(1.0f);

        Reservoir1sppStreamState stream_state;
        Reservoir1spp reservoir;
        const uint reservoir_payload = uint(x) | (uint(y) << 16);

        if (is_rtdgi_tracing_frame(g)) {
            const float hit_t = length(hit_offset_ws);
            // do_the_thing (restir_temporal.hlsl:55-62): read the traced candidate
            const float3 out_value = candidate_radiance_tex.load(px).xyz();
            const float inv_pdf = 1;
            const float3 cand_hit_normal_ws = direction_view_to_world(vc, candidate_normal_tex.load(px).xyz());

            const float p_q = 1.0f * max(0.0f, sRGB_to_luminance(out_value)) * step(0.0f, dot(outgoing_dir, normal_ws));
            radiance_sel = out_value;
            ray_orig_sel_ws = refl_ray_origin_ws;
            ray_hit_sel_ws = refl_ray_origin_ws + outgoing_dir * hit_t;
            hit_normal_sel = cand_hit_normal_ws;
            reservoir.init_with_stream(p_q, inv_pdf, stream_state, reservoir_payload);

            float rl = lerp(candidate_history_tex.load(px).y, sqrt(hit_t), 0.05f);
            candidate_out_tex.store(px, float4(sqrt(hit_t), rl, 0, 0));
        }

        const float rt_invalidity = sqrt(saturate(rt_invalidity_tex.load(px).y));
        const uint MAX_RESOLVE_SAMPLE_COUNT = 5;
        float center_M = 0;

        for (uint sample_i = 0; sample_i < MAX_RESOLVE_SAMPLE_COUNT && stream_state.M_sum < 1.25f * RESTIR_TEMPORAL_M_CLAMP; ++sample_i) {
            const int2 rpx_offset = get_rpx_offset(sample_i, frame_index);
            if (sample_i > 0 && rpx_offset == int2(0, 0)) continue;

            const float4 reproj = reprojection_tex.load(hi_px + rpx_offset * 2);

            const uint xor_seq[4][2] = {{3, 3}, {2, 1}, {1, 2}, {3, 3}};
            const uint pxv[2] = {xor_seq[frame_index & 3][0], xor_seq[frame_index & 3][1]};

            // (px + rpx_offset) is computed in uint (wraps below zero), XORed, converted to float (restir_temporal.hlsl:216-238)
            const uint perm_x = (uint(x) + uint(rpx_offset.x)) ^ pxv[0], perm_y = (uint(y) + uint(rpx_offset.y)) ^ pxv[1];
            const float2 base = sample_i == 0 ? float2(float(uint(x)), float(uint(y))) : float2(float(perm_x), float(perm_y));
            const int2 permuted_reproj_px(
                kjb_cvt_i32(floor(base.x + gbuffer_tex_size.x * reproj.x * 0.5f + 0.0f + 0.5f)),
                kjb_cvt_i32(floor(base.y + gbuffer_tex_size.y * reproj.y * 0.5f + 0.0f + 0.5f)));
            // index arithmetic wraps like the shader's 32-bit ints
            const int2 rpx(int(uint(permuted_reproj_px.x) + uint(rpx_offset.x)), int(uint(permuted_reproj_px.y) + uint(rpx_offset.y)));

            const int2 permuted_neighbor_px(kjb_cvt_i32(floor(base.x + 0.5f)), kjb_cvt_i32(floor(base.y + 0.5f)));
            const int2 neighbor_px(int(uint(permuted_neighbor_px.x) + uint(rpx_offset.x)), int(uint(permuted_neighbor_px.y) + uint(rpx_offset.y)));
            const int2 neighbor_px_hi(int(uint(neighbor_px.x) * 2u + uint(hi_px_offset.x)), int(uint(neighbor_px.y) * 2u + uint(hi_px_offset.y)));

            uint4 rraw = reservoir_history_tex.load_u(rpx);
            Reservoir1spp r = Reservoir1spp::from_raw(uint2(rraw.x, rraw.y));
            const int2 spx = reservoir_payload_to_px(r.payload);

            float visibility = 1;
            float relevance = 1;
            const float sample_depth = depth_tex.load(neighbor_px_hi).x;

            const float3 prev_ray_orig = ray_orig_history_tex.load(spx).xyz();
            if (length(prev_ray_orig - refl_ray_origin_ws) > 0.1f * -view_ray_context.ray_hit_vs().z) continue;
            if (0 == sample_depth) continue;
            if (reproj.z == 0) continue;

            relevance *= 1 - smoothstep(0.0f, 0.1f, inverse_depth_relative_diff(depth, sample_depth));

            const float3 sample_normal_vs = half_view_normal_tex.load(neighbor_px).xyz();
            const float normal_similarity_dot = max(0.0f, dot(sample_normal_vs, normal_vs));
            const float normal_cutoff = 0.2f;
            if (sample_i != 0 && normal_similarity_dot < normal_cutoff) continue;
            relevance *= pow(normal_similarity_dot, 4.0f);

            const float4 sample_hit_ws_and_dist = ray_history_tex.load(spx) + float4(prev_ray_orig, 0.0f);
            const float3 sample_hit_ws = sample_hit_ws_and_dist.xyz();
            const float prev_dist = sample_hit_ws_and_dist.w;

            const float4 hn = hit_normal_history_tex.load(spx);
            const float4 sample_hit_normal_ws_dot(hn.x * 2 - 1, hn.y * 2 - 1, hn.z * 2 - 1, hn.w);   // decode_hit_normal_and_dot

            const float3 dir_to_sample_hit_unnorm = sample_hit_ws - refl_ray_origin_ws;
            const float dist_to_sample_hit = length(dir_to_sample_hit_unnorm);
            const float3 dir_to_sample_hit = normalize(dir_to_sample_hit_unnorm);
            const float center_to_hit_vis = -dot(sample_hit_normal_ws_dot.xyz(), dir_to_sample_hit);

            const float4 prev_rad = radiance_history_tex.load(spx) * float4(g.fc.pre_exposure_delta, g.fc.pre_exposure_delta, g.fc.pre_exposure_delta, 1);

            r.M = max(0.0f, min(r.M, exp2(log2(RESTIR_TEMPORAL_M_CLAMP) * (1.0f - rt_invalidity))));

            const float p_q = 1 * max(0.0f, sRGB_to_luminance(prev_rad.xyz())) * step(0.0f, dot(dir_to_sample_hit, normal_ws));

            float jacobian = 1;
            {
                jacobian *= clamp(prev_dist / dist_to_sample_hit, 1e-4f, 1e4f);
                jacobian *= jacobian;
                jacobian *= clamp(center_to_hit_vis / sample_hit_normal_ws_dot.w, 0.0f, 1e4f);
            }
            r.M *= relevance;
            if (0 == sample_i) center_M = r.M;

            if (reservoir.update_with_stream(r, p_q, jacobian * visibility, stream_state, reservoir_payload, rng)) {
                outgoing_dir = dir_to_sample_hit;
                src_px_sel = rpx;
                radiance_sel = prev_rad.xyz();
                ray_orig_sel_ws = prev_ray_orig;
                ray_hit_sel_ws = sample_hit_ws;
                hit_normal_sel = sample_hit_normal_ws_dot.xyz();
            }
        }
        reservoir.finish_stream(stream_state);
        reservoir.W = min(reservoir.W, RESTIR_RESERVOIR_W_CLAMP);
        reservoir.M = center_M + 0.5f;

        const float4 hit_normal_ws_dot = float4(hit_normal_sel, -dot(hit_normal_sel, outgoing_dir));
        radiance_out_tex.store(px, float4(radiance_sel, dot(normal_ws, outgoing_dir)));
        ray_orig_output_tex.store(px, float4(ray_orig_sel_ws, 0.0f));
        hit_normal_output_tex.store(px, float4(hit_normal_ws_dot.x * 0.5f + 0.5f, hit_normal_ws_dot.y * 0.5f + 0.5f, hit_normal_ws_dot.z * 0.5f + 0.5f, hit_normal_ws_dot.w));
        ray_output_tex.store(px, float4(ray_hit_sel_ws - ray_orig_sel_ws, length(ray_hit_sel_ws - refl_ray_origin_ws)));
        uint2 rr = reservoir.as_raw();
        reservoir_out_tex.store_u(x, y, uint4(rr.x, rr.y, 0, 0));

        TemporalReservoirOutput res_packed;
        res_packed.depth = depth;
        res_packed.ray_hit_offset_ws = ray_hit_sel_ws - view_ray_context.ray_hit_ws();
        res_packed.luminance = max(0.0f, sRGB_to_luminance(radiance_sel));
        res_packed.hit_normal_ws = hit_normal_ws_dot.xyz();
        temporal_reservoir_packed_tex.store_u(x, y, res_packed.as_raw());
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D7: rtdgi/restir_spatial.hlsl:48-372 (+ occlusion_raymarch.hlsl:69-146)
static float normal_inluence_nonlinearity(float x, float b) { return x < -b ? 0.0f : (x + b) * (x + b) / (4 * b); }   // :41-45

int kjb_pass_rtdgi_restir_spatial(kjb_context* ctx, const kjb_rtdgi_restir_spatial_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img reservoir_input_tex(a->reservoir_input_tex), half_view_normal_tex(a->half_view_normal_tex), half_depth_tex(a->half_depth_tex),
        half_ssao_tex(a->half_ssao_tex), temporal_reservoir_packed_tex(a->temporal_reservoir_packed_tex), reservoir_output_tex(a->reservoir_output_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size), output_tex_size = f4(a->output_tex_size);
    const uint spatial_reuse_pass_idx = a->spatial_reuse_pass_idx;
    const int W = reservoir_output_tex.w(), H = reservoir_output_tex.h();
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);

    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y); const int2 hi_px = px * 2 + hso;
        float depth = half_depth_tex.load(px).x;
        const uint seed = g.fc.frame_index + spatial_reuse_pass_idx * 123;
        uint rng = hash3(uint(x), uint(y), seed);
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_depth(vc, uv, depth);
        const float3 center_normal_vs = half_view_normal_tex.load(px).xyz();
        const float3 center_normal_ws = direction_view_to_world(vc, center_normal_vs);
        const float center_depth = half_depth_tex.load(px).x;
        const float center_ssao = half_ssao_tex.load(px).x;

        Reservoir1sppStreamState stream_state;
        Reservoir1spp reservoir;
        float sample_radius_offset = uint_to_u01_float(hash1_mut(rng));
        uint4 craw = reservoir_input_tex.load_u(px);
        Reservoir1spp center_r = Reservoir1spp::from_raw(uint2(craw.x, craw.y));
        float kernel_tightness = 1.0f - center_ssao;
        const uint SAMPLE_COUNT_PASS0 = 8, SAMPLE_COUNT_PASS1 = 5;
        const float MAX_INPUT_M_IN_PASS0 = RESTIR_TEMPORAL_M_CLAMP;
        const float MAX_INPUT_M_IN_PASS1 = MAX_INPUT_M_IN_PASS0 * float(SAMPLE_COUNT_PASS0);
        const float MAX_INPUT_M_IN_PASS = spatial_reuse_pass_idx == 0 ? MAX_INPUT_M_IN_PASS0 : MAX_INPUT_M_IN_PASS1;
        kernel_tightness = lerp(kernel_tightness, 1.0f, 0.5f * smoothstep(MAX_INPUT_M_IN_PASS * 0.5f, MAX_INPUT_M_IN_PASS, center_r.M));
        float max_kernel_radius = spatial_reuse_pass_idx == 0 ? lerp(32.0f, 12.0f, kernel_tightness) : lerp(16.0f, 6.0f, kernel_tightness);
        if (spatial_reuse_pass_idx >= 2) max_kernel_radius = 8;
        const float2 dist_to_edge_xy = min(float2(float(x), float(y)), float2(output_tex_size.x, output_tex_size.y) - float2(float(x), float(y)));
        const float allow_edge_overstep = center_r.M < 10 ? 100.0f : 1.25f;
        const float2 kernel_radius = min(float2(max_kernel_radius), dist_to_edge_xy * allow_edge_overstep);
        uint sample_count = spatial_reuse_pass_idx == 0 ? SAMPLE_COUNT_PASS0 : SAMPLE_COUNT_PASS1;
        const uint2 ang_offset_seed = spatial_reuse_pass_idx == 0 ? uint2(uint(x) >> 3, uint(y) >> 3) : uint2(uint(x) >> 2, uint(y) >> 2);
        float ang_offset = uint_to_u01_float(hash3(ang_offset_seed.x, ang_offset_seed.y, g.fc.frame_index * 2 + spatial_reuse_pass_idx)) * M_PI_F * 2;

        for (uint sample_i = 0; sample_i < sample_count; ++sample_i) {
            float ang = (float(sample_i) + ang_offset) * GOLDEN_ANGLE;
            float2 radius = 0 == sample_i ? float2(0.0f) : (pow((float(sample_i) + sample_radius_offset) / float(sample_count), 0.5f) * kernel_radius);
            float2 off_f = float2(cos(ang), sin(ang)) * radius;
            int2 rpx_offset(kjb_cvt_i32(off_f.x), kjb_cvt_i32(off_f.y));
            const bool is_center_sample = sample_i == 0;
            const int2 rpx = px + rpx_offset;

            const uint4 reservoir_raw = reservoir_input_tex.load_u(rpx);
            if (0 == reservoir_raw.x) continue;
            Reservoir1spp r = Reservoir1spp::from_raw(uint2(reservoir_raw.x, reservoir_raw.y));
            r.M = min(r.M, 500.0f);
            const int2 spx = reservoir_payload_to_px(r.payload);
            const TemporalReservoirOutput spx_packed = TemporalReservoirOutput::from_raw(temporal_reservoir_packed_tex.load_u(spx));
            const float reused_luminance = spx_packed.luminance;

            float visibility = 1;
            float relevance = 1;
            const float3 sample_normal_vs = half_view_normal_tex.load(rpx).xyz();
            const float normal_similarity_dot = dot(sample_normal_vs, center_normal_vs);
            relevance *= normal_inluence_nonlinearity(normal_similarity_dot, 0.5f) / normal_inluence_nonlinearity(1.0f, 0.5f);
            const float sample_ssao = half_ssao_tex.load(rpx).x;
            relevance *= 1 - abs(sample_ssao - center_ssao);

            const float2 rpx_uv = get_uv(rpx * 2 + hso, gbuffer_tex_size);
            const float rpx_depth = half_depth_tex.load(rpx).x;
            if (rpx_depth == 0.0f) continue;
            const ViewRayContext rpx_ray_ctx = ViewRayContext::from_uv_and_depth(vc, rpx_uv, rpx_depth);
            const float2 spx_uv = get_uv(spx * 2 + hso, gbuffer_tex_size);
            const ViewRayContext spx_ray_ctx = ViewRayContext::from_uv_and_depth(vc, spx_uv, spx_packed.depth);
            const float3 sample_hit_ws = spx_packed.ray_hit_offset_ws + spx_ray_ctx.ray_hit_ws();
            const float3 reused_dir_to_sample_hit_unnorm_ws = sample_hit_ws - rpx_ray_ctx.ray_hit_ws();
            const float reused_dist = length(reused_dir_to_sample_hit_unnorm_ws);
            const float3 reused_dir_to_sample_hit_ws = reused_dir_to_sample_hit_unnorm_ws / reused_dist;
            const float3 dir_to_sample_hit_unnorm = sample_hit_ws - view_ray_context.ray_hit_ws();
            const float dist_to_sample_hit = length(dir_to_sample_hit_unnorm);
            const float3 dir_to_sample_hit = normalize(dir_to_sample_hit_unnorm);

            if (!is_center_sample) {
                const float depth_diff = abs(max(0.3f, center_normal_vs.z) * (center_depth / rpx_depth - 1.0f));
                const float depth_threshold = spatial_reuse_pass_idx == 0 ? 0.15f : 0.1f;
                relevance *= 1 - smoothstep(0.0f, depth_threshold, depth_diff);
            }

            if (a->perform_occlusion_raymarch) {   // occlusion_raymarch.hlsl:69-146 with halfres depth, max 6 samples
                const float2 ray_orig_uv = spx_uv;
                const float surface_offset_len = length(ViewRayContext::from_uv_and_depth(vc, ray_orig_uv, depth).ray_hit_vs() - view_ray_context.ray_hit_vs());
                const float MAX_RAYMARCH_DIST_MULT = 3.0f;
                const float3 raymarch_dir_unnorm_ws = sample_hit_ws - view_ray_context.ray_hit_ws();
                const float3 raymarch_end_ws = view_ray_context.ray_hit_ws()
                    + raymarch_dir_unnorm_ws * min(1.0f, MAX_RAYMARCH_DIST_MULT * surface_offset_len / length(raymarch_dir_unnorm_ws));

                const float2 raymarch_start_uv = uv;
                const float3 raymarch_start_cs = view_ray_context.ray_hit_cs.xyz();
                const float2 fullres_depth_tex_size(gbuffer_tex_size.x, gbuffer_tex_size.y);
                const float2 halfres_depth_tex_size(output_tex_size.x, output_tex_size.y);
                const float3 raymarch_end_cs = position_world_to_clip(vc, raymarch_end_ws);
                const float2 raymarch_end_uv = cs_to_uv(float2(raymarch_end_cs.x, raymarch_end_cs.y));
                const float2 raymarch_uv_delta = raymarch_end_uv - raymarch_start_uv;
                const float2 raymarch_len_px = raymarch_uv_delta * halfres_depth_tex_size;
                const uint MIN_PX_PER_STEP = 2;
                int k_count = kjb_cvt_i32(floor(length(raymarch_len_px) / float(MIN_PX_PER_STEP)));
                if (k_count > 6) k_count = 6;
                const float Z_LAYER_THICKNESS = 0.05f;
                const float depth_step_per_z = (raymarch_end_cs.z - raymarch_start_cs.z) / length(float2(raymarch_end_cs.x, raymarch_end_cs.y) - float2(raymarch_start_cs.x, raymarch_start_cs.y));
                float t_step = 1.0f / float(k_count);
                float t = 0.5f * t_step;
                for (int k = 0; k < k_count; ++k) {
                    const float3 interp_pos_cs = lerp(raymarch_start_cs, raymarch_end_cs, t);
                    const float2 uv_at_interp = cs_to_uv(float2(interp_pos_cs.x, interp_pos_cs.y));
                    uint2 px_at_interp(
                        (kjb_cvt_u32(floor(uv_at_interp.x * fullres_depth_tex_size.x - float(hso.x))) & ~1u) + uint(hso.x),
                        (kjb_cvt_u32(floor(uv_at_interp.y * fullres_depth_tex_size.y - float(hso.y))) & ~1u) + uint(hso.y));
                    float depth_at_interp = half_depth_tex.load(int(px_at_interp.x >> 1u), int(px_at_interp.y >> 1u)).x;
                    const float2 quantized_cs_at_interp = uv_to_cs((float2(float(px_at_interp.x), float(px_at_interp.y)) + 0.5f) / fullres_depth_tex_size);
                    const float biased_interp_z = raymarch_start_cs.z + depth_step_per_z * length(quantized_cs_at_interp - float2(raymarch_start_cs.x, raymarch_start_cs.y));
                    if (depth_at_interp > biased_interp_z) {
                        const float depth_diff = inverse_depth_relative_diff(interp_pos_cs.z, depth_at_interp);
                        float hit = smoothstep(Z_LAYER_THICKNESS, Z_LAYER_THICKNESS * 0.5f, depth_diff);
                        visibility *= 1 - hit;
                    }
                    t += t_step;
                }
            }

            const float3 sample_hit_normal_ws = spx_packed.hit_normal_ws;
            const float center_to_hit_vis = -dot(sample_hit_normal_ws, dir_to_sample_hit);
            const float reused_to_hit_vis = -dot(sample_hit_normal_ws, reused_dir_to_sample_hit_ws);
            float p_q = 1;
            p_q *= reused_luminance;
            p_q *= max(0.0f, dot(dir_to_sample_hit, center_normal_ws));
            float jacobian = 1;
            jacobian *= reused_dist / dist_to_sample_hit;
            jacobian *= jacobian;
            jacobian *= clamp(center_to_hit_vis / reused_to_hit_vis, 0.0f, 1e4f);
            jacobian = sqrt(jacobian);
            if (is_center_sample) jacobian = 1;
            if (!(p_q >= 0)) continue;
            r.M *= relevance;
            if (a->occlusion_raymarch_importance_only) { p_q *= lerp(0.25f, 1.0f, visibility); visibility = 1; }
            reservoir.update_with_stream(r, p_q, visibility * jacobian, stream_state, r.payload, rng);
        }
        reservoir.finish_stream(stream_state);
        reservoir.W = min(reservoir.W, RESTIR_RESERVOIR_W_CLAMP);
        uint2 rr = reservoir.as_raw();
        reservoir_output_tex.store_u(x, y, uint4(rr.x, rr.y, 0, 0));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D8: rtdgi/restir_check.rgen.hlsl:21-66 (optional pass)
int kjb_pass_rtdgi_restir_check(kjb_context* ctx, const kjb_rtdgi_restir_check_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img half_depth_tex(a->half_depth_tex), temporal_reservoir_packed_tex(a->temporal_reservoir_packed_tex), reservoir_input_tex(a->reservoir_input_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size);
    const int W = reservoir_input_tex.w(), H = reservoir_input_tex.h();
    const int2 hi_px_offset = halfres_subsample_offset(g.fc.frame_index);
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y); const int2 hi_px = px * 2 + hi_px_offset;
        const float depth = half_depth_tex.load(px).x;
        const float2 uv = get_uv(hi_px, gbuffer_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_biased_depth(vc, uv, depth);
        uint4 raw = reservoir_input_tex.load_u(px);
        Reservoir1spp r = Reservoir1spp::from_raw(uint2(raw.x, raw.y));
        const int2 spx = reservoir_payload_to_px(r.payload);
        const TemporalReservoirOutput spx_packed = TemporalReservoirOutput::from_raw(temporal_reservoir_packed_tex.load_u(spx));
        const float2 spx_uv = get_uv(spx * 2 + hi_px_offset, gbuffer_tex_size);
        const ViewRayContext spx_ray_ctx = ViewRayContext::from_uv_and_depth(vc, spx_uv, spx_packed.depth);
        const float3 hit_ws = spx_packed.ray_hit_offset_ws + spx_ray_ctx.ray_hit_ws();
        const float3 spx_pos_ws = spx_ray_ctx.ray_hit_ws();
        const float3 trace_origin_ws = view_ray_context.biased_secondary_ray_origin_ws();
        const float3 trace_vec = hit_ws - trace_origin_ws;
        if (rt_is_shadowed(ctx->scene, trace_origin_ws, normalize(trace_vec), 0.0f, min(5 * length(spx_pos_ws - trace_origin_ws), length(trace_vec) * 0.999f))) {
            r.W = 0;
            uint2 rr = r.as_raw(); reservoir_input_tex.store_u(x, y, uint4(rr.x, rr.y, 0, 0));
        }
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D9: rtdgi/restir_resolve.hlsl:42-205
static float ggx_ndf_unnorm(float a2, float cos_theta) { float ds = cos_theta * cos_theta * (a2 - 1.0f) + 1.0f; return a2 / (ds * ds); }

int kjb_pass_rtdgi_restir_resolve(kjb_context* ctx, const kjb_rtdgi_restir_resolve_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img radiance_tex(a->radiance_tex), reservoir_input_tex(a->reservoir_input_tex), gbuffer_tex(a->gbuffer_tex), depth_tex(a->depth_tex), half_view_normal_tex(a->half_view_normal_tex),
        half_depth_tex(a->half_depth_tex), ssao_tex(a->ssao_tex), candidate_radiance_tex(a->candidate_radiance_tex), candidate_hit_tex(a->candidate_hit_tex),
        temporal_reservoir_packed_tex(a->temporal_reservoir_packed_tex), irradiance_output_tex(a->irradiance_output_tex);
    const float4 gbuffer_tex_size = f4(a->gbuffer_tex_size), output_tex_size = f4(a->output_tex_size);
    const int W = irradiance_output_tex.w(), H = irradiance_output_tex.h();
    const int2 hso = halfres_subsample_offset(g.fc.frame_index);
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y);
        float depth = depth_tex.load(px).x;
        if (0 == depth) { irradiance_output_tex.store(px, float4(0.0f)); continue; }
        const float2 uv = get_uv(px, gbuffer_tex_size);
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_depth(vc, uv, depth);
        GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(px));
        const float3 center_normal_ws = gbuffer.normal;
        const float3 center_normal_vs = direction_world_to_view(vc, center_normal_ws);
        const float center_depth = depth;
        const float center_ssao = ssao_tex.load(px).x;
        const uint frame_hash = hash1(g.fc.frame_index);
        const uint px_idx_in_quad = (((uint(x) & 1) | (uint(y) & 1) * 2) + frame_hash) & 3;
        const float4 blue = blue_noise_for_pixel(g, uint2(x, y), g.fc.frame_index) * M_TAU_F;
        const float NEAR_FIELD_FADE_OUT_END = -view_ray_context.ray_hit_vs().z * (SSGI_NEAR_FIELD_RADIUS * output_tex_size.w * 0.5f);
        const float NEAR_FIELD_FADE_OUT_START = NEAR_FIELD_FADE_OUT_END * 0.5f;
        const float near_field_influence = center_ssao;

        float3 total_irradiance(0.0f);
        bool sharpen_gi_kernel = false;
        {
            float w_sum = 0; float3 weighted_irradiance(0.0f);
            for (uint sample_i = 0; sample_i < 4; ++sample_i) {
                const float ang = (float(sample_i) + blue.x) * GOLDEN_ANGLE + (float(px_idx_in_quad) / 4.0f) * M_TAU_F;
                const float radius = pow(float(sample_i), 0.666f) * 1.0f + 0.4f;
                const float2 reservoir_px_offset = float2(cos(ang), sin(ang)) * radius;
                const int2 rpx(kjb_cvt_i32(floor(float(x) * 0.5f + reservoir_px_offset.x)), kjb_cvt_i32(floor(float(y) * 0.5f + reservoir_px_offset.y)));
                const float2 rpx_uv = get_uv(rpx * 2 + hso, gbuffer_tex_size);
                const float rpx_depth = half_depth_tex.load(rpx).x;
                const ViewRayContext rpx_ray_ctx = ViewRayContext::from_uv_and_depth(vc, rpx_uv, rpx_depth);
                {
                    const float3 hit_ws = candidate_hit_tex.load(rpx).xyz() + rpx_ray_ctx.ray_hit_ws();
                    const float3 sample_offset = hit_ws - view_ray_context.ray_hit_ws();
                    const float sample_dist = length(sample_offset);
                    const float3 sample_dir = sample_offset / sample_dist;
                    const float geometric_term = 2 * max(0.0f, dot(center_normal_ws, sample_dir));
                    const float atten = smoothstep(NEAR_FIELD_FADE_OUT_END, NEAR_FIELD_FADE_OUT_START, sample_dist);
                    sharpen_gi_kernel |= atten > 0.9f;
                    float3 contribution = candidate_radiance_tex.load(rpx).xyz() * geometric_term;
                    contribution *= lerp(0.0f, atten, near_field_influence);
                    float3 sample_normal_vs = half_view_normal_tex.load(rpx).xyz();
                    float w = 1;
                    w *= ggx_ndf_unnorm(0.01f, saturate(dot(center_normal_vs, sample_normal_vs)));
                    w *= exp2(-200.0f * abs(center_normal_vs.z * (center_depth / rpx_depth - 1.0f)));
                    weighted_irradiance = mad(contribution, w, weighted_irradiance);
                    w_sum += w;
                }
            }
            total_irradiance += weighted_irradiance / max(1e-20f, w_sum);
        }
        {
            float w_sum = 0; float3 weighted_irradiance(0.0f);
            const float kernel_scale = sharpen_gi_kernel ? 0.5f : 1.0f;
            for (uint sample_i = 0; sample_i < 4; ++sample_i) {
                const float ang = (float(sample_i) + blue.x) * GOLDEN_ANGLE + (float(px_idx_in_quad) / 4.0f) * M_TAU_F;
                const float radius = pow(float(sample_i), 0.666f) * 1.0f * kernel_scale + 0.4f * kernel_scale;
                const float2 reservoir_px_offset = float2(cos(ang), sin(ang)) * radius;
                const int2 rpx(kjb_cvt_i32(floor(float(x) * 0.5f + reservoir_px_offset.x)), kjb_cvt_i32(floor(float(y) * 0.5f + reservoir_px_offset.y)));
                uint4 rraw = reservoir_input_tex.load_u(rpx);
                Reservoir1spp r = Reservoir1spp::from_raw(uint2(rraw.x, rraw.y));
                const int2 spx = reservoir_payload_to_px(r.payload);
                const TemporalReservoirOutput spx_packed = TemporalReservoirOutput::from_raw(temporal_reservoir_packed_tex.load_u(spx));
                const float2 spx_uv = get_uv(spx * 2 + hso, gbuffer_tex_size);
                const ViewRayContext spx_ray_ctx = ViewRayContext::from_uv_and_depth(vc, spx_uv, spx_packed.depth);
                {
                    const float rpx_depth = half_depth_tex.load(rpx).x;
                    const float3 hit_ws = spx_packed.ray_hit_offset_ws + spx_ray_ctx.ray_hit_ws();
                    const float3 sample_offset = hit_ws - view_ray_context.ray_hit_ws();
                    const float sample_dist = length(sample_offset);
                    const float3 sample_dir = sample_offset / sample_dist;
                    const float geometric_term = 2 * max(0.0f, dot(center_normal_ws, sample_dir));
                    float3 radiance = radiance_tex.load(spx).xyz();
                    {
                        const float atten = smoothstep(NEAR_FIELD_FADE_OUT_START, NEAR_FIELD_FADE_OUT_END, sample_dist);
                        radiance *= lerp(1.0f, atten, near_field_influence);
                    }
                    const float3 contribution = radiance * geometric_term * r.W;
                    float3 sample_normal_vs = half_view_normal_tex.load(spx).xyz();
                    const float sample_ssao = ssao_tex.load(rpx * 2 + hso).x;
                    float w = 1;
                    w *= ggx_ndf_unnorm(0.01f, saturate(dot(center_normal_vs, sample_normal_vs)));
                    w *= exp2(-200.0f * abs(center_normal_vs.z * (center_depth / rpx_depth - 1.0f)));
                    w *= exp2(-20.0f * abs(center_ssao - sample_ssao));
                    weighted_irradiance = mad(contribution, w, weighted_irradiance);
                    w_sum += w;
                }
            }
            total_irradiance += weighted_irradiance / max(1e-20f, w_sum);
        }
        irradiance_output_tex.store(px, float4(total_irradiance, 1));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D10: rtdgi/temporal_filter.hlsl:39-252
int kjb_pass_rtdgi_temporal(kjb_context* ctx, const kjb_rtdgi_temporal_args* a) {
    const Globals& g = ctx->g;
    Img input_tex(a->input_tex), history_tex(a->history_tex), variance_history_tex(a->variance_history_tex), reprojection_tex(a->reprojection_tex),
        rt_history_invalidity_tex(a->rt_history_invalidity_tex), output_tex(a->output_tex), history_output_tex(a->history_output_tex), variance_history_output_tex(a->variance_history_output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    const float ped = g.fc.pre_exposure_delta;
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y);
        float2 uv = get_uv(px, output_tex_size);
        float4 center = linear_rgb_to_crunched_luma_chroma(input_tex.load(px));
        float4 reproj = reprojection_tex.load(px);
        const float4 history_mult(ped, ped, ped, 1);
        float4 history = linear_rgb_to_crunched_luma_chroma(history_tex.load(px) * history_mult);

        float4 vsum(0.0f), vsum2(0.0f); float wsum = 0, hist_diff = 0, hist_vsum = 0, hist_vsum2 = 0;
        const int k = 2;
        for (int yy = -k; yy <= k; ++yy) for (int xx = -k; xx <= k; ++xx) {
            float4 neigh = linear_rgb_to_crunched_luma_chroma(input_tex.load(x + xx, y + yy));
            float4 hist_neigh = linear_rgb_to_crunched_luma_chroma(history_tex.load(x + xx, y + yy) * history_mult);
            float neigh_luma = neigh.x, hist_luma = hist_neigh.x;
            float w = exp(-3.0f * float(xx * xx + yy * yy) / float((k + 1.) * (k + 1.)));
            vsum = mad(neigh, w, vsum);
            vsum2 = mad(neigh * neigh, w, vsum2);
            wsum += w;
            hist_diff += abs(neigh_luma - hist_luma) / max(1e-5f, neigh_luma + hist_luma) * w;
            hist_vsum = mad(hist_luma, w, hist_vsum);
            hist_vsum2 = mad(hist_luma * hist_luma, w, hist_vsum2);
        }
        float4 ex = vsum / wsum, ex2 = vsum2 / wsum;
        float4 dev = sqrt(max(float4(0.0f), ex2 - ex * ex));
        hist_diff /= wsum; hist_vsum /= wsum; hist_vsum2 /= wsum;

        float4 mh = variance_history_tex.sample_bilinear_clamp(uv + reproj.xy());
        const float2 moments_history = float2(mh.x, mh.y) * float2(ped, ped * ped);
        const float center_luma = center.x + (hist_vsum - ex.x);
        const float2 current_moments(center_luma, center_luma * center_luma);
        float2 vout = max(float2(0.0f), lerp(moments_history, current_moments, 0.25f));
        variance_history_output_tex.store(px, float4(vout.x, vout.y, 0, 0));
        const float center_temporal_dev = sqrt(max(0.0f, moments_history.y - moments_history.x * moments_history.x));

        float temporal_change = abs(hist_vsum - ex.x) / max(1e-8f, hist_vsum + ex.x);
        const float rt_invalid = saturate(sqrt(rt_history_invalidity_tex.load(x / 2, y / 2).x) * 4);
        const float current_sample_count = history.w;
        float clamp_box_size = 1 * lerp(0.25f, 2.0f, 1.0f - rt_invalid) * lerp(0.333f, 1.0f, saturate(reproj.w)) * 2;
        clamp_box_size = max(clamp_box_size, 0.5f);
        float4 nmin = center - dev * clamp_box_size, nmax = center + dev * clamp_box_size;
        float4 clamped_history = float4(clamp(history.xyz(), nmin.xyz(), nmax.xyz()), history.w);
        const float variance_adjusted_temporal_change = smoothstep(0.1f, 1.0f, 0.05f * temporal_change / center_temporal_dev);
        float max_sample_count = 32;
        max_sample_count = lerp(max_sample_count, 4.0f, variance_adjusted_temporal_change);
        max_sample_count *= lerp(1.0f, 0.5f, rt_invalid);
        float3 res = lerp(clamped_history.xyz(), center.xyz(), 1.0f / (1.0f + min(max_sample_count, current_sample_count)));
        const float output_sample_count = min(current_sample_count, max_sample_count) + 1;
        float4 output = crunched_luma_chroma_to_linear_rgb(float4(res, output_sample_count));
        history_output_tex.store(px, output);
        output_tex.store(px, float4(output.xyz(), saturate(output_sample_count * lerp(1.0f, 0.5f, rt_invalid) * smoothstep(0.3f, 0.0f, temporal_change) / 32.0f)));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ D11: rtdgi/spatial_filter.hlsl:33-101
int kjb_pass_rtdgi_spatial(kjb_context* ctx, const kjb_rtdgi_spatial_args* a) {
    const Globals& g = ctx->g;
    Img input_tex(a->input_tex), depth_tex(a->depth_tex), ssao_tex(a->ssao_tex), geometric_normal_tex(a->geometric_normal_tex), output_tex(a->output_tex);
    const int W = output_tex.w(), H = output_tex.h();
    auto crunch = [](float3 v) { return v * rcp(max3(v.x, v.y, v.z) + 1.0f); };
    auto uncrunch = [](float3 v) { return v * rcp(1.0f - max3(v.x, v.y, v.z)); };
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y);
        float4 sum(0.0f);
        const float4 cin = input_tex.load(px);
        const float center_validity = cin.w;
        const float center_depth = depth_tex.load(px).x;
        const float center_ssao = ssao_tex.load(px).x;
        const float3 center_value = cin.xyz();
        const float3 center_normal_vs = geometric_normal_tex.load(px).xyz() * 2.0f - 1.0f;
        if (center_validity == 1) { output_tex.store(px, float4(center_value, 1.0f)); continue; }
        const float ang_off = float((g.fc.frame_index * 23) % 32) * M_TAU_F + interleaved_gradient_noise(uint2(x, y)) * M_PI_F;
        const uint MAX_SAMPLE_COUNT = 8;
        const float MAX_RADIUS_PX = sqrt(lerp(16.0f * 16.0f, 2.0f * 2.0f, center_validity));
        const float KERNEL_SHARPNESS = 0.666f;
        uint sample_count = kjb_cvt_u32(exp2(4.0f * square(1.0f - center_validity)));
        sample_count = sample_count < 2 ? 2 : (sample_count > MAX_SAMPLE_COUNT ? MAX_SAMPLE_COUNT : sample_count);
        sum += float4(crunch(center_value), 1);
        const float RADIUS_SAMPLE_MULT = MAX_RADIUS_PX / pow(float(MAX_SAMPLE_COUNT - 1), KERNEL_SHARPNESS);
        for (uint sample_i = 1; sample_i < MAX_SAMPLE_COUNT; ++sample_i) {
            const float ang = (float(sample_i) + ang_off) * GOLDEN_ANGLE;
            float radius = pow(float(sample_i), KERNEL_SHARPNESS) * RADIUS_SAMPLE_MULT;
            float2 sample_offset = float2(cos(ang), sin(ang)) * radius;
            // `px + sample_offset` is uint2 + float2 -> float2, truncated to int2 (spatial_filter.hlsl:78)
            const int2 sample_px(kjb_cvt_i32(float(x) + sample_offset.x), kjb_cvt_i32(float(y) + sample_offset.y));
            const float sample_depth = depth_tex.load(sample_px).x;
            const float3 sample_val = input_tex.load(sample_px).xyz();
            const float sample_ssao = ssao_tex.load(sample_px).x;
            if (sample_depth != 0 && sample_i < sample_count) {
                float wt = 1;
                wt *= exp2(-100.0f * abs(center_normal_vs.z * (center_depth / sample_depth - 1.0f)));
                wt *= exp2(-20.0f * abs(sample_ssao - center_ssao));
                sum = mad(float4(crunch(sample_val), 1.0f), wt, sum);
            }
        }
        float norm_factor = 1.0f / max(1e-5f, sum.w);
        float3 filtered = uncrunch(sum.xyz() * norm_factor);
        output_tex.store(px, float4(filtered, 1.0f));
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"

}  // namespace kjo
