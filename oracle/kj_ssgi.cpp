// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// Screen-space ambient occlusion guide (crates/lib/kajiya/src/renderers/ssgi.rs, assets/shaders/ssgi/*.hlsl with USE_AO_ONLY 1):
// the lighting gather of the shader only feeds `color_accum`, which the AO-only output never reads, so it is not restated.
#include "kj_ctx.h"

namespace kjo {
namespace {
inline float fast_sqrt(float x) { return asfloat(0x1fbd1df5u + (asuint(x) >> 1u)); }                       // ssgi.hlsl:51-53
inline float fast_acos(float inX) {                                                                         // :56-61
    float x = abs(inX);
    float res = -0.156583f * x + 1.57079632679489661923f;
    res *= fast_sqrt(1.0f - x);
    return (inX >= 0) ? res : M_PI_F - res;
}
inline float integrate_arc(float h1, float h2, float n) {                                                   // :103-107
    float a = -cos(2.0f * h1 - n) + cos(n) + 2.0f * h1 * sin(n);
    float b = -cos(2.0f * h2 - n) + cos(n) + 2.0f * h2 * sin(n);
    return 0.25f * (a + b);
}
inline float update_horizion_angle(float prev, float cur, float blend) { return cur > prev ? lerp(prev, cur, blend) : prev; }   // :109-111
}  // namespace

extern "C" {

// ------------------------------------------------------------------ "ssao": ssgi/ssgi.hlsl:214-341 (half-res, SSGI_HALF_SAMPLE_COUNT 6)
int kjb_pass_ssao(kjb_context* ctx, const kjb_ssao_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img gbuffer_tex(a->gbuffer_tex), depth_tex(a->half_depth_tex), output_tex(a->output_tex);
    const float4 input_tex_size = f4(a->input_tex_size), output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    static const float temporal_rotations[6] = {60.0f, 300.0f, 180.0f, 240.0f, 120.0f, 0.0f};
    static const float temporal_offsets[4] = {0.0f, 0.5f, 0.25f, 0.75f};
    const uint SSGI_HALF_SAMPLE_COUNT = 6;
    const float M_FRAC_PI_2 = 1.57079632679489661923f;
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 uv = get_uv(int2(x, y), output_tex_size);
        const float depth = depth_tex.load(x, y).x;
        if (0.0f == depth) { output_tex.store(x, y, float4(0, 0, 0, 1)); continue; }
        const GbufferData gbuffer = gbuffer_unpack(gbuffer_tex.load_u(x * 2, y * 2));
        const float3 normal_vs = normalize(mul(vc.world_to_view, float4(gbuffer.normal, 0)).xyz());
        const ViewRayContext view_ray_context = ViewRayContext::from_uv_and_depth(vc, uv, depth);
        const float3 v_vs = -normalize(view_ray_context.ray_dir_vs());
        const float4 ray_hit_cs = view_ray_context.ray_hit_cs;
        const float3 ray_hit_vs = view_ray_context.ray_hit_vs();
        const uint ux = uint(x), uy = uint(y);
        const float spatial_direction_noise = 1.0f / 16.0f * float((((ux + uy) & 3u) << 2u) + (ux & 3u));
        const float temporal_direction_noise = temporal_rotations[g.fc.frame_index % 6u] / 360.0f;
        const float spatial_offset_noise = (1.0f / 4.0f) * float((uy - ux) & 3u);
        const float temporal_offset_noise = temporal_offsets[g.fc.frame_index / 6u % 4u];
        const float ss_angle = frac(spatial_direction_noise + temporal_direction_noise) * M_PI_F;
        const float rand_offset = frac(spatial_offset_noise + temporal_offset_noise);
        float2 cs_slice_dir(cos(ss_angle) * input_tex_size.y / input_tex_size.x, sin(ss_angle));
        float kernel_radius_ws, kernel_radius_shrinkage;
        {
            const float ws_to_cs = 0.5f / -ray_hit_vs.z * vc.view_to_clip.m[5];
            const float cs_kernel_radius_scaled = 60.0f * output_tex_size.w;   // SSGI_KERNEL_RADIUS
            kernel_radius_ws = cs_kernel_radius_scaled / ws_to_cs;
            cs_slice_dir = cs_slice_dir * cs_kernel_radius_scaled;
            kernel_radius_shrinkage = min(1.0f, 0.4f / cs_kernel_radius_scaled);   // MAX_KERNEL_RADIUS_CS
        }
        cs_slice_dir = cs_slice_dir * kernel_radius_shrinkage;
        kernel_radius_ws *= kernel_radius_shrinkage;
        const float3 center_vs = ray_hit_vs;
        cs_slice_dir = cs_slice_dir * (1.0f / float(SSGI_HALF_SAMPLE_COUNT));
        // mul(float4(cs_slice_dir, 0, 0), sample_to_view).xy — row vector times matrix
        const float* m = vc.sample_to_view.m;
        const float2 vs_slice_dir(kjb_fma(cs_slice_dir.y, m[1], cs_slice_dir.x * m[0]), kjb_fma(cs_slice_dir.y, m[5], cs_slice_dir.x * m[4]));
        const float3 slice_normal_vs = normalize(cross(v_vs, float3(vs_slice_dir.x, vs_slice_dir.y, 0)));
        float3 proj_normal_vs = normal_vs - slice_normal_vs * dot(slice_normal_vs, normal_vs);
        const float slice_contrib_weight = length(proj_normal_vs);
        proj_normal_vs = proj_normal_vs / slice_contrib_weight;
        const float n_angle = fast_acos(clamp(dot(proj_normal_vs, v_vs), -1.0f, 1.0f)) * sign(dot(vs_slice_dir, float2(proj_normal_vs.x - v_vs.x, proj_normal_vs.y - v_vs.y)));
        float theta_cos_max1 = cos(n_angle - M_FRAC_PI_2), theta_cos_max2 = cos(n_angle + M_FRAC_PI_2);
        int2 prev_sample_coord0(x, y), prev_sample_coord1(x, y);
        // process_sample (:121-206) restricted to what feeds the horizon angles
        auto process_sample = [&](float4 sample_cs, float theta_cos_max) {
            if (sample_cs.z > 0) {
                const float4 sample_vs4 = mul(vc.sample_to_view, sample_cs);
                const float3 sample_vs = sample_vs4.xyz() / sample_vs4.w;
                const float3 sample_vs_offset = sample_vs - center_vs;
                const float sample_vs_offset_len = length(sample_vs_offset);
                const float sample_theta_cos = dot(sample_vs_offset, v_vs) / sample_vs_offset_len;
                const float sample_distance_normalized = sample_vs_offset_len / kernel_radius_ws;
                if (sample_distance_normalized < 1.0f) {
                    const float sample_influence = smoothstep(1.0f, 0.0f, sample_distance_normalized);
                    theta_cos_max = update_horizion_angle(theta_cos_max, sample_theta_cos, sample_influence);
                }
            } else {
                theta_cos_max = update_horizion_angle(theta_cos_max, -1.0f, 1.0f);   // sky: assume no occlusion
            }
            return theta_cos_max;
        };
        for (uint i = 0; i < SSGI_HALF_SAMPLE_COUNT; ++i) {
            {
                const float t = float(i) + rand_offset;
                float4 sample_cs(ray_hit_cs.x - cs_slice_dir.x * t, ray_hit_cs.y - cs_slice_dir.y * t, 0, 1);
                const float2 suv = cs_to_uv(float2(sample_cs.x, sample_cs.y));
                const int2 sample_px(kjb_cvt_i32(output_tex_size.x * suv.x), kjb_cvt_i32(output_tex_size.y * suv.y));
                if (sample_px.x != prev_sample_coord0.x || sample_px.y != prev_sample_coord0.y) {
                    prev_sample_coord0 = sample_px;
                    sample_cs.z = depth_tex.load(sample_px).x;
                    theta_cos_max1 = process_sample(sample_cs, theta_cos_max1);
                }
            }
            {
                const float t = float(i) + (1.0f - rand_offset);
                float4 sample_cs(ray_hit_cs.x + cs_slice_dir.x * t, ray_hit_cs.y + cs_slice_dir.y * t, 0, 1);
                const float2 suv = cs_to_uv(float2(sample_cs.x, sample_cs.y));
                const int2 sample_px(kjb_cvt_i32(output_tex_size.x * suv.x), kjb_cvt_i32(output_tex_size.y * suv.y));
                if (sample_px.x != prev_sample_coord1.x || sample_px.y != prev_sample_coord1.y) {
                    prev_sample_coord1 = sample_px;
                    sample_cs.z = depth_tex.load(sample_px).x;
                    theta_cos_max2 = process_sample(sample_cs, theta_cos_max2);
                }
            }
        }
        const float h1 = -fast_acos(theta_cos_max1), h2 = +fast_acos(theta_cos_max2);
        const float h1p = n_angle + max(h1 - n_angle, -M_FRAC_PI_2), h2p = n_angle + min(h2 - n_angle, M_FRAC_PI_2);
        const float inv_ao = integrate_arc(h1p, h2p, n_angle);
        float4 col(max(0.0f, inv_ao));   // USE_AO_ONLY: rgb = a
        col = col * slice_contrib_weight;
        output_tex.store(x, y, max(float4(0.0f), col));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ "ssao spatial": ssgi/spatial_filter.hlsl:35-73
int kjb_pass_ssao_spatial(kjb_context* ctx, const kjb_ssao_spatial_args* a) {
    Img ssgi_tex(a->ssgi_tex), depth_tex(a->depth_tex), normal_tex(a->normal_tex), output_tex(a->output_tex);
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        float4 result(0.0f); float w_sum = 0.0f;
        const float center_depth = depth_tex.load(x, y).x;
        if (center_depth != 0.0f) {
            const float3 center_normal = normal_tex.load(x, y).xyz();
            w_sum = 1.0f; result = ssgi_tex.load(x, y);
            for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
                if (xx == 0 && yy == 0) continue;
                const float depth = depth_tex.load(x + xx, y + yy).x;
                const float4 ssgi = ssgi_tex.load(x + xx, y + yy);
                const float3 normal = normal_tex.load(x + xx, y + yy).xyz();
                if (depth != 0.0f) {
                    const float depth_diff = 1.0f - (center_depth / depth);
                    const float depth_factor = exp2(-200.0f * abs(depth_diff));
                    float normal_factor = max(0.0f, dot(normal, center_normal));
                    normal_factor *= normal_factor; normal_factor *= normal_factor;
                    float w = 1; w *= depth_factor; w *= normal_factor;
                    w_sum += w; result = mad(ssgi, w, result);
                }
            }
        }
        result = float4(result.x);   // USE_AO_ONLY
        output_tex.store(x, y, result / max(w_sum, 1e-5f));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ "ssao upsample": ssgi/upsample.hlsl:38-75
int kjb_pass_ssao_upsample(kjb_context* ctx, const kjb_ssao_upsample_args* a) {
    Img ssgi_tex(a->ssgi_tex), depth_tex(a->depth_tex), output_tex(a->output_tex);
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        float4 result(0.0f); float w_sum = 0.0f;
        const float center_depth = depth_tex.load(x, y).x;
        if (center_depth != 0.0f) {
            for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
                const int spx = x / 2 + xx, spy = y / 2 + yy;
                const float depth = depth_tex.load(spx * 2, spy * 2).x;
                const float4 ssgi = ssgi_tex.load(spx, spy);
                if (depth != 0.0f) {   // the normal factor is computed but unused by the shader (:26-27)
                    const float depth_diff = 1.0f - (center_depth / depth);
                    const float depth_factor = exp2(-200.0f * abs(depth_diff));
                    float w = 1; w *= depth_factor;
                    w *= exp(-dot(float2(float(xx), float(yy)), float2(float(xx), float(yy))));
                    w_sum += w; result = mad(ssgi, w, result);
                }
            }
        }
        result = float4(result.x);
        if (w_sum > 1e-6f) output_tex.store(x, y, result / w_sum);
        else output_tex.store(x, y, ssgi_tex.load(x / 2, y / 2));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ "ssao temporal": ssgi/temporal_filter.hlsl:19-61
int kjb_pass_ssao_temporal(kjb_context* ctx, const kjb_ssao_temporal_args* a) {
    Img input_tex(a->input_tex), history_tex(a->history_tex), reprojection_tex(a->reprojection_tex), final_output_tex(a->final_output_tex), history_output_tex(a->history_output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int W = history_output_tex.w(), H = history_output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 uv = get_uv(int2(x, y), output_tex_size);
        const float4 center = input_tex.load(x, y);
        const float4 reproj = reprojection_tex.load(x, y);
        const float4 history = history_tex.sample_bilinear_clamp(uv + reproj.xy());
        float4 vsum(0.0f), vsum2(0.0f); float wsum = 0.0f;
        const int k = 2;
        for (int yy = -k; yy <= k; ++yy) for (int xx = -k; xx <= k; ++xx) {
            const float4 neigh = input_tex.load(x + xx * 2, y + yy * 2);
            const float w = exp(-3.0f * float(xx * xx + yy * yy) / float((k + 1.) * (k + 1.)));
            vsum = mad(neigh, w, vsum); vsum2 = mad(neigh * neigh, w, vsum2); wsum += w;
        }
        const float4 ex = vsum / wsum, ex2 = vsum2 / wsum;
        const float4 dev = sqrt(max(float4(0.0f), ex2 - ex * ex));
        const float box_size = 0.5f, n_deviations = 5.0f;
        const float4 mid = lerp(center, ex, box_size * box_size);
        const float4 nmin = mid - dev * box_size * n_deviations, nmax = mid + dev * box_size * n_deviations;
        const float4 clamped_history = min(max(history, nmin), nmax);
        float4 res = lerp(clamped_history, center, 1.0f / 8.0f);
        res = float4(res.x);
        history_output_tex.store(x, y, res);
        final_output_tex.store(x, y, res);
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"
}  // namespace kjo
