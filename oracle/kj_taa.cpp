// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// CPU restatement of kajiya's temporal anti-aliasing / super-resolution, one function per render-graph pass of
// crates/lib/kajiya/src/renderers/taa.rs:41-185.  Shader paths relative to /root/reference/assets/shaders/.
#include "kj_ctx.h"

namespace kjo {

namespace {
// taa/taa_common.hlsl (TAA_NONLINEARITY_TYPE 1, TAA_COLOR_MAPPING_MODE 1)
inline float linear_to_perceptual(float a) { return sqrt(max(0.0f, a)); }
inline float perceptual_to_linear(float a) { return a * a; }
inline float3 decode_rgb(float3 v) { float mc = max3(v.x, v.y, v.z); return v * linear_to_perceptual(mc) / max(1e-20f, mc); }
inline float3 encode_rgb(float3 v) { float mc = max3(v.x, v.y, v.z); return v * perceptual_to_linear(mc) / max(1e-20f, mc); }
inline float4 input_remap(float4 v) { return float4(sRGB_to_YCbCr(decode_rgb(v.xyz())), 1); }

// inc/image.hlsl:85-162 image_sample_catmull_rom_approx(useCornerTaps=false) with sampler_llc and HistoryRemap (reproject_history.hlsl:27-37)
float4 catmull_rom_5tap_history(const Img& tex, float2 uv, float2 texSize, float ped) {
    auto remap = [&](float4 v) { return float4(decode_rgb(v.xyz() * ped), v.w); };
    auto smp = [&](float2 p) { return remap(tex.sample_bilinear_clamp(p)); };
    float2 samplePos = uv * texSize;
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
    result = result / (w12.x * w0.y + w0.x * w12.y + w12.x * w12.y + w3.x * w12.y + w12.x * w3.y);
    return result;
}

struct UnjitteredSampleInfo { float4 color; float coverage; float3 ex, ex2; };
// inc/unjitter_taa.hlsl:58-125 with InputRemap
UnjitteredSampleInfo sample_image_unjitter_taa(const Img& img, int2 output_px, float2 output_tex_size, float2 sample_offset_pixels, float kernel_scale, int k) {
    const float2 input_tex_size(float(img.w()), float(img.h()));
    const float2 irs = input_tex_size / output_tex_size;
    const int2 base_src_px(kjb_cvt_i32((float(output_px.x) + 0.5f) * irs.x), kjb_cvt_i32((float(output_px.y) + 0.5f) * irs.y));
    const float2 dst_sample_loc = float2(float(output_px.x), float(output_px.y)) + 0.5f;
    const float2 base_src_sample_loc = (float2(float(base_src_px.x), float(base_src_px.y)) + 0.5f + sample_offset_pixels * float2(1, -1)) / irs;
    float4 res(0.0f); float3 ex(0.0f), ex2(0.0f); float dev_wt_sum = 0.0f, wt_sum = 0.0f;
    const float kernel_distance_mult = 1.0f * kernel_scale;
    for (int y = -k; y <= k; ++y) for (int x = -k; x <= k; ++x) {
        int2 src_px = base_src_px + int2(x, y);
        float2 src_sample_loc = base_src_sample_loc + float2(float(x), float(y)) / irs;
        float4 col = input_remap(img.load(src_px));
        float2 sco = (src_sample_loc - dst_sample_loc) * kernel_distance_mult;
        float dist2 = dot(sco, sco);
        float dev_wt = exp2(-dist2 * irs.x);
        float wt = exp2(-10 * dist2 * irs.x);
        res = mad(col, wt, res); wt_sum += wt;
        ex = mad(col.xyz(), dev_wt, ex); ex2 = mad(col.xyz() * col.xyz(), dev_wt, ex2); dev_wt_sum += dev_wt;
    }
    UnjitteredSampleInfo info; info.color = res; info.coverage = wt_sum; info.ex = ex / dev_wt_sum; info.ex2 = ex2 / dev_wt_sum;
    return info;
}
}  // namespace

extern "C" {

// ------------------------------------------------------------------ T1 taa/reproject_history.hlsl:38-129
int kjb_pass_taa_reproject(kjb_context* ctx, const kjb_taa_reproject_args* a) {
    Img history_tex(a->history_tex), reprojection_tex(a->reprojection_tex), depth_tex(a->depth_tex), output_tex(a->output_tex), closest_velocity_output(a->closest_velocity_output);
    const float4 input_tex_size = f4(a->input_tex_size), output_tex_size = f4(a->output_tex_size);
    const float ped = ctx->g.fc.pre_exposure_delta;
    const int W = output_tex.w(), H = output_tex.h();
    const float2 irs = float2(input_tex_size.x, input_tex_size.y) / float2(output_tex_size.x, output_tex_size.y);
    auto reproj_px_of = [&](int x, int y) { return int2(int(kjb_cvt_u32((float(x) + 0.5f) * irs.x)), int(kjb_cvt_u32((float(y) + 0.5f) * irs.y))); };
    auto should_dilate0 = [&](int x, int y) {
        const int2 rp = reproj_px_of(x, y);
        float2 v = reprojection_tex.load(rp.x - 1, rp.y - 1).xy(); float2 vel_min = v, vel_max = v;
        v = reprojection_tex.load(rp.x + 1, rp.y - 1).xy(); vel_min = min(vel_min, v); vel_max = max(vel_max, v);
        v = reprojection_tex.load(rp.x - 1, rp.y + 1).xy(); vel_min = min(vel_min, v); vel_max = max(vel_max, v);
        v = reprojection_tex.load(rp.x + 1, rp.y + 1).xy(); vel_min = min(vel_min, v); vel_max = max(vel_max, v);
        const float2 d = vel_max - vel_min, thr = 0.1f * max(float2(input_tex_size.z, input_tex_size.w), abs(vel_max + vel_min));
        return d.x > thr.x || d.y > thr.y;
    };
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 reproj_px = reproj_px_of(x, y);
        float2 uv = get_uv(int2(x, y), output_tex_size);
        int2 closest_px = reproj_px;
        // WaveReadLaneAt(lane ^ 2), (lane ^ 16) inside the 8x8 group = pixels (x^2, y), (x, y^2)
        const bool should_dilate = should_dilate0(x, y) || should_dilate0(x ^ 2, y) || should_dilate0(x, y ^ 2) || should_dilate0(x ^ 2, y ^ 2);
        if (should_dilate) {
            float reproj_depth = depth_tex.load(reproj_px).x;
            for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
                float d = depth_tex.load(reproj_px.x + xx, reproj_px.y + yy).x;
                if (d > reproj_depth) { reproj_depth = d; closest_px = int2(reproj_px.x + xx, reproj_px.y + yy); }
            }
        }
        const float2 reproj_xy = reprojection_tex.load(closest_px).xy();
        closest_velocity_output.store(x, y, float4(reproj_xy.x, reproj_xy.y, 0, 0));
        float2 history_uv = uv + reproj_xy;
        float4 history_packed = catmull_rom_5tap_history(history_tex, history_uv, float2(output_tex_size.x, output_tex_size.y), ped);
        output_tex.store(x, y, float4(history_packed.xyz(), max(0.0f, history_packed.w)));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ T2 taa/filter_input.hlsl:32-89
int kjb_pass_taa_filter_input(kjb_context* ctx, const kjb_taa_filter_input_args* a) {
    Img input_tex(a->input_tex), depth_tex(a->depth_tex), output_tex(a->output_tex), dev_output_tex(a->dev_output_tex);
    const int W = output_tex.w(), H = output_tex.h();
    struct FilteredInput { float3 clamped_ex, var; };
    auto inner = [&](int px, int py, float center_depth, float luma_cutoff, float depth_scale) {
        float3 iex(0.0f), iex2(0.0f), clamped_iex(0.0f); float iwsum = 0, clamped_iwsum = 0;
        const int k = 1;
        for (int y = -k; y <= k; ++y) for (int x = -k; x <= k; ++x) {
            const float distance_w = exp(-(0.8f / float(k * k)) * float(x * x + y * y));
            float3 s = input_remap(input_tex.load(px + x, py + y)).xyz();
            const float depth = depth_tex.load(px + x, py + y).x;
            float w = 1;
            w *= exp2(-min(16.0f, depth_scale * inverse_depth_relative_diff(center_depth, depth)));
            w *= distance_w;
            w *= pow(saturate(luma_cutoff / s.x), 8.0f);
            clamped_iwsum += w; clamped_iex = mad(s, w, clamped_iex);
            iwsum += 1; iex += s; iex2 += s * s;
        }
        clamped_iex = clamped_iex / clamped_iwsum;
        iex = iex / iwsum; iex2 = iex2 / iwsum;
        FilteredInput r; r.clamped_ex = clamped_iex; r.var = max(float3(0.0f), iex2 - iex * iex);
        return r;
    };
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float center_depth = depth_tex.load(x, y).x;
        FilteredInput fi = inner(x, y, center_depth, 1e10f, 200);
        FilteredInput cfi = inner(x, y, center_depth, fi.clamped_ex.x * 1.001f, 200);
        output_tex.store(x, y, float4(cfi.clamped_ex, 0));
        dev_output_tex.store(x, y, float4(sqrt(fi.var), 0));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ T3 taa/filter_history.hlsl:15-62
int kjb_pass_taa_filter_history(kjb_context* ctx, const kjb_taa_filter_history_args* a) {
    Img input_tex(a->input_tex), output_tex(a->output_tex);
    const float4 input_tex_size = f4(a->input_tex_size), output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    auto filter_input = [&](float2 uv, float luma_cutoff, int k) {
        float3 iex(0.0f); float iwsum = 0;
        int2 src_px(kjb_cvt_i32(floor(uv.x * input_tex_size.x + 1e-3f)), kjb_cvt_i32(floor(uv.y * input_tex_size.y + 1e-3f)));
        for (int y = -k; y <= k; ++y) for (int x = -k; x <= k; ++x) {
            const float distance_w = exp(-(0.8f / float(k * k)) * float(x * x + y * y));
            float3 s = sRGB_to_YCbCr(input_tex.load(src_px.x + x, src_px.y + y).xyz());
            float w = 1;
            w *= distance_w;
            w *= pow(saturate(luma_cutoff / s.x), 8.0f);
            iwsum += w; iex = mad(s, w, iex);
        }
        return iex / iwsum;
    };
    const int k = (input_tex_size.x / output_tex_size.x > 1.75f) ? 2 : 1;
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        float2 uv = get_uv(int2(x, y), output_tex_size);
        float filtered_luma = filter_input(uv, 1e10f, k).x;
        output_tex.store(x, y, float4(filter_input(uv, filtered_luma * 1.001f, k), 0));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ T4 taa/input_prob.hlsl:47-109
int kjb_pass_taa_input_prob(kjb_context* ctx, const kjb_taa_input_prob_args* a) {
    const Globals& g = ctx->g;
    Img filtered_input_tex(a->filtered_input_tex), filtered_input_dev_tex(a->filtered_input_dev_tex), filtered_history_tex(a->filtered_history_tex),
        reprojection_tex(a->reprojection_tex), smooth_var_history_tex(a->smooth_var_history_tex), velocity_history_tex(a->velocity_history_tex), output_tex(a->output_tex);
    const float4 input_tex_size = f4(a->input_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    const float2 sop(g.fc.view_constants.sample_offset_pixels[0], g.fc.view_constants.sample_offset_pixels[1]);
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        float input_prob = 0;
        float3 ivar(0.0f);
        for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) ivar = max(ivar, filtered_input_dev_tex.load(x + xx * 2, y + yy * 2).xyz());
        ivar = ivar * ivar;
        const float2 input_uv = (float2(float(x), float(y)) + sop) * float2(input_tex_size.z, input_tex_size.w);
        const float4 closest_history = filtered_history_tex.sample_nearest_clamp(input_uv);
        const float2 rxy = reprojection_tex.load(x, y).xy();
        const float3 closest_smooth_var = smooth_var_history_tex.sample_bilinear_clamp(input_uv + rxy).xyz();
        const float2 closest_vel = velocity_history_tex.sample_bilinear_clamp(input_uv + rxy).xy() * g.fc.delta_time_seconds;
        const float3 combined_var = min(closest_smooth_var, ivar * 10.0f);
        for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
            const float3 s = filtered_input_tex.load(x + xx, y + yy).xyz();
            const float3 idiff = s - closest_history.xyz();
            const float2 vel = reprojection_tex.load(x + xx, y + yy).xy();
            const float vdiff = length((vel - closest_vel) / max(float2(1.0f), abs(vel + closest_vel)));
            float prob = exp2(-1.0f * length(idiff * idiff / max(float3(1e-6f), combined_var)) - 1000 * vdiff);
            input_prob = max(input_prob, prob);
        }
        output_tex.store(x, y, float4(input_prob));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ T5 taa/filter_prob.hlsl, T6 taa/filter_prob2.hlsl
int kjb_pass_taa_prob_filter(kjb_context* ctx, const kjb_taa_prob_filter_args* a) {
    Img input_tex(a->input_tex), output_tex(a->output_tex);
    pass_rows(ctx, output_tex.h(), [&](int y) { for (int x = 0; x < output_tex.w(); ++x) {
        float prob = input_tex.load(x, y).x;
        for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) prob = max(prob, input_tex.load(x + xx, y + yy).x);
        output_tex.store(x, y, float4(prob));
    } }, ctx->num_threads);
    return 0;
}
int kjb_pass_taa_prob_filter2(kjb_context* ctx, const kjb_taa_prob_filter_args* a) {
    Img input_tex(a->input_tex), output_tex(a->output_tex);
    pass_rows(ctx, output_tex.h(), [&](int y) { for (int x = 0; x < output_tex.w(); ++x) {
        float2 weighted_prob(0.0f);
        const float SQUISH_STRENGTH = 10;
        for (int yy = -2; yy <= 2; ++yy) for (int xx = -2; xx <= 2; ++xx) {
            float neighbor_prob = input_tex.load(x + xx * 2, y + yy * 2).x;
            weighted_prob += float2(exponential_squish(neighbor_prob, SQUISH_STRENGTH), 1);
        }
        output_tex.store(x, y, float4(exponential_unsquish(weighted_prob.x / weighted_prob.y, SQUISH_STRENGTH)));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ T7 taa/taa.hlsl:94-338
int kjb_pass_taa(kjb_context* ctx, const kjb_taa_args* a) {
    const Globals& g = ctx->g;
    Img input_tex(a->input_tex), history_tex(a->history_tex), reprojection_tex(a->reprojection_tex), closest_velocity_tex(a->closest_velocity_tex), velocity_history_tex(a->velocity_history_tex),
        smooth_var_history_tex(a->smooth_var_history_tex), input_prob_tex(a->input_prob_tex), temporal_output_tex(a->temporal_output_tex), output_tex(a->output_tex),
        smooth_var_output_tex(a->smooth_var_output_tex), velocity_output_tex(a->velocity_output_tex);
    const float4 input_tex_size = f4(a->input_tex_size), output_tex_size = f4(a->output_tex_size);
    const int W = temporal_output_tex.w(), H = temporal_output_tex.h();
    const float2 sop(g.fc.view_constants.sample_offset_pixels[0], g.fc.view_constants.sample_offset_pixels[1]);
    const float dt = g.fc.delta_time_seconds;
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 irf = float2(input_tex_size.x, input_tex_size.y) / float2(output_tex_size.x, output_tex_size.y);
        const int2 reproj_px(int(kjb_cvt_u32((float(x) + 0.5f) * irf.x)), int(kjb_cvt_u32((float(y) + 0.5f) * irf.y)));
        float2 uv = get_uv(int2(x, y), output_tex_size);
        float4 history_packed = history_tex.load(x, y);
        float3 history = history_packed.xyz();
        float history_coverage = max(0.0f, history_packed.w);
        // fetch_blurred_history(px, 2, 1) (:60-81)
        float4 bhistory_packed;
        {
            float4 csum(0.0f); float wsum = 0;
            for (int yy = -2; yy <= 2; ++yy) for (int xx = -2; xx <= 2; ++xx) {
                float4 c = history_tex.load(x + xx, y + yy);
                float2 offset = float2(float(xx), float(yy)) * 1.0f;
                float w = exp(-dot(offset, offset));
                csum = mad(c, w, csum); wsum += w;
            }
            bhistory_packed = csum / wsum;
        }
        float3 bhistory = bhistory_packed.xyz();
        float3 bhistory_coverage(bhistory_packed.w);
        history = sRGB_to_YCbCr(history); bhistory = sRGB_to_YCbCr(bhistory);
        const float4 reproj = reprojection_tex.load(reproj_px);
        const float2 reproj_xy = closest_velocity_tex.load(x, y).xy();
        UnjitteredSampleInfo center_sample = sample_image_unjitter_taa(input_tex, int2(x, y), float2(output_tex_size.x, output_tex_size.y), sop, 1.0f, 1);
        UnjitteredSampleInfo bcenter_sample = sample_image_unjitter_taa(input_tex, int2(x, y), float2(output_tex_size.x, output_tex_size.y), sop, 0.333f, 1);
        float coverage = 1;
        float3 center = center_sample.color.xyz();
        coverage = center_sample.coverage;
        float3 bcenter = bcenter_sample.color.xyz() / bcenter_sample.coverage;
        history = lerp(history, bcenter, saturate(1.0f - history_coverage));
        bhistory = lerp(bhistory, bcenter, saturate(float3(1.0f) - bhistory_coverage));
        const float input_prob = input_prob_tex.load(reproj_px).x;
        float3 ex = center_sample.ex, ex2 = center_sample.ex2;
        const float3 var = max(float3(0.0f), ex2 - ex * ex);
        const float3 prev_var(smooth_var_history_tex.sample_bilinear_clamp(uv + reproj_xy).x);
        const float2 vel_now = closest_velocity_tex.load(x, y).xy() / dt;
        const float2 vel_prev = velocity_history_tex.sample_bilinear_clamp(uv + closest_velocity_tex.load(x, y).xy()).xy();
        const float vel_diff = length((vel_now - vel_prev) / max(float2(1.0f), abs(vel_now + vel_prev)));
        const float var_blend = saturate(0.3f + 0.7f * (1 - reproj.z) + vel_diff);
        float3 smooth_var = max(var, lerp(prev_var, var, var_blend));
        const float var_prob_blend = saturate(input_prob);
        smooth_var = lerp(var, smooth_var, var_prob_blend);
        const float3 input_dev = sqrt(var);
        float3 clamped_history;
        {
            float box_n_deviations = 0.8f;
            box_n_deviations = lerp(box_n_deviations, 3.0f, input_prob);
            float3 nmin = ex - input_dev * box_n_deviations, nmax = ex + input_dev * box_n_deviations;
            float3 clamped_bhistory = clamp(bhistory, nmin, nmax);
            const float clamping_event = length(max(float3(0.0f), max(bhistory - nmax, nmin - bhistory)) / max(float3(0.01f), ex));
            float3 outlier3 = max(float3(0.0f), (max(nmin - history, history - nmax)) / (0.1f + max(max(abs(history), abs(ex)), float3(1e-5f))));
            float3 boutlier3 = max(float3(0.0f), (max(nmin - bhistory, bhistory - nmax)) / (0.1f + max(max(abs(bhistory), abs(ex)), float3(1e-5f))));
            float outlier = max(outlier3.x, max(outlier3.y, outlier3.z));
            float boutlier = max(boutlier3.x, max(boutlier3.y, boutlier3.z));
            const float2 huv = uv + reproj_xy, hs = saturate(huv);
            const bool history_valid = huv.x == hs.x && huv.y == hs.y;
            if (history_valid) {
                const float non_disoccluding_outliers = max(0.0f, outlier - boutlier) * 10;
                const float3 unclamped_history_detail = history - clamped_bhistory;
                const float temporal_clamping_detail = abs(unclamped_history_detail.x / max(1e-3f, input_dev.x)) * 0.05f;   // length(scalar) = abs
                const float temporal_stability = saturate(1 - temporal_clamping_detail);
                const float allow_unclamped_detail = saturate(non_disoccluding_outliers) * temporal_stability;
                float3 history_detail = history - bhistory;
                history_detail = lerp(history_detail, unclamped_history_detail, allow_unclamped_detail);
                const float initial_bclamp_amount = saturate(dot(clamped_bhistory - bhistory, bcenter - bhistory)
                    / max(1e-5f, length(clamped_bhistory - bhistory) * length(bcenter - bhistory)));
                const float effective_clamp_amount = saturate(initial_bclamp_amount) * (1 - allow_unclamped_detail);
                const float keep_detail = 1 - effective_clamp_amount;
                history_detail *= keep_detail;
                clamped_history = clamped_bhistory + history_detail;
                if (irf.x < 1.0f) history_coverage *= lerp(lerp(0.0f, 0.9f, keep_detail), 1.0f, saturate(10 * clamping_event));
            } else {
                clamped_history = clamped_bhistory; coverage = 1; center = bcenter; history_coverage = 0;
            }
            clamped_history = lerp(clamped_history, history, smoothstep(0.5f, 1.0f, input_prob));
        }
        float total_coverage = max(1e-5f, history_coverage + coverage);
        float3 temporal_result = (clamped_history * history_coverage + center) / total_coverage;
        const float max_coverage = max(2.0f, 8.0f / (irf.x * irf.y));
        total_coverage = min(max_coverage, total_coverage);
        coverage = total_coverage;
        smooth_var_output_tex.store(x, y, float4(smooth_var, 0));
        temporal_result = YCbCr_to_sRGB(temporal_result);
        temporal_result = encode_rgb(temporal_result);
        temporal_result = max(float3(0.0f), temporal_result);
        temporal_output_tex.store(x, y, float4(temporal_result, coverage));
        output_tex.store(x, y, float4(temporal_result, 0));   // this_frame_result = lerp(temporal_result, 0, a = 0): rgb = temporal_result, a = 0
        const float2 vo = closest_velocity_tex.load(x, y).xy() / dt;
        velocity_output_tex.store(x, y, float4(vo.x, vo.y, 0, 0));
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"
}  // namespace kjo
