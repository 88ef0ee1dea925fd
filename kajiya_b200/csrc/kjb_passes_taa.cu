// Temporal anti-aliasing / temporal super-resolution as sm_100a kernels — one per render-graph pass of
// crates/lib/kajiya/src/renderers/taa.rs:41-185 (shaders: /root/reference/assets/shaders/taa/, inc/unjitter_taa.hlsl, inc/image.hlsl).
// All seven passes are small-stencil screen-space filters: 32x8 thread blocks on the pass's grid, 8-byte RGBA16F texels =>
// 256 B per warp-row request; constant filter weights (exp(-r^2 ...)) are evaluated once on the host with the numeric
// contract and passed as kernel parameters.
#include "kjb_context.h"

using namespace kjb;

KJB_DEV float3 taa_decode_rgb(float3 v) { const float mc = max3(v.x, v.y, v.z); return v * kjb_sqrt(kjb_max(0.0f, mc)) / kjb_max(1e-20f, mc); }   // taa_common.hlsl:46-53
KJB_DEV float3 taa_encode_rgb(float3 v) { const float mc = max3(v.x, v.y, v.z); return v * (mc * mc) / kjb_max(1e-20f, mc); }                     // :55-62
KJB_DEV float3 taa_input_remap(float4 v) { return rgb_to_ycbcr(taa_decode_rgb(xyz(v))); }
struct W9 { float w[9]; };
struct W25t { float w[25]; };

// ------------------------------------------------------------------ T1 reproject_history.hlsl:38-129
KJB_DEV bool t1_should_dilate0(const Img& reprojection_tex, int x, int y, float2 irs, float4 its) {
    const int rx = int(kjb_cvt_u32((float(x) + 0.5f) * irs.x)), ry = int(kjb_cvt_u32((float(y) + 0.5f) * irs.y));
    float2 v = xy(ld_rgba16s(reprojection_tex, rx - 1, ry - 1)); float2 vel_min = v, vel_max = v;
    v = xy(ld_rgba16s(reprojection_tex, rx + 1, ry - 1)); vel_min = vmin(vel_min, v); vel_max = vmax(vel_max, v);
    v = xy(ld_rgba16s(reprojection_tex, rx - 1, ry + 1)); vel_min = vmin(vel_min, v); vel_max = vmax(vel_max, v);
    v = xy(ld_rgba16s(reprojection_tex, rx + 1, ry + 1)); vel_min = vmin(vel_min, v); vel_max = vmax(vel_max, v);
    const float2 d = vel_max - vel_min, thr = 0.1f * vmax(f2(its.z, its.w), vabs(vel_max + vel_min));
    return d.x > thr.x || d.y > thr.y;
}
KJB_KERNEL(256) k_taa_reproject(Globals g, Img history_tex, Img reprojection_tex, Img depth_tex, ImgW output_tex, ImgW closest_velocity_output, float4 its, float4 ots, Rows kjb_rows) {
    KJB_PX; const int W = output_tex.w, H = output_tex.h; if (x >= W || y >= H) return;
    const float ped = g.fc.pre_exposure_delta;
    const float2 irs = f2(its.x, its.y) / f2(ots.x, ots.y);
    const int rx = int(kjb_cvt_u32((float(x) + 0.5f) * irs.x)), ry = int(kjb_cvt_u32((float(y) + 0.5f) * irs.y));
    const float s4[4] = {ots.x, ots.y, ots.z, ots.w};
    const float2 uv = get_uv(x, y, s4);
    int cx = rx, cy = ry;
    // lane^2 / lane^16 exchange of the 8x8 group == pixels (x^2,y), (x,y^2) and transitively (x^2,y^2)
    const bool should_dilate = t1_should_dilate0(reprojection_tex, x, y, irs, its) || t1_should_dilate0(reprojection_tex, x ^ 2, y, irs, its)
                            || t1_should_dilate0(reprojection_tex, x, y ^ 2, irs, its) || t1_should_dilate0(reprojection_tex, x ^ 2, y ^ 2, irs, its);
    if (should_dilate) {
        float reproj_depth = ld_r32f(depth_tex, rx, ry);
        for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
            const float d = ld_r32f(depth_tex, rx + xx, ry + yy);
            if (d > reproj_depth) { reproj_depth = d; cx = rx + xx; cy = ry + yy; }
        }
    }
    const float2 reproj_xy = xy(ld_rgba16s(reprojection_tex, cx, cy));
    st_rg16f(closest_velocity_output, x, y, reproj_xy.x, reproj_xy.y);
    const float2 history_uv = uv + reproj_xy;
    // image_sample_catmull_rom_5tap (inc/image.hlsl:85-162, no corner taps), bilinear taps with clamp addressing, HistoryRemap
    const Img& ht = history_tex;
    auto smp = [&](float2 p) {
        const float4 h = bilinear_clamp(ht.w, ht.h, p, [&](int sx, int sy) { return ld_rgba16f(ht, sx, sy); });
        return f4(taa_decode_rgb(xyz(h) * ped), h.w);
    };
    const float2 texSize = f2(ots.x, ots.y);
    const float2 samplePos = history_uv * texSize;
    const float2 texPos1 = vfloor(samplePos - 0.5f) + 0.5f;
    const float2 f = samplePos - texPos1;
    const float2 w0 = f * (-0.5f + f * (1.0f - 0.5f * f));
    const float2 w1 = 1.0f + f * f * (-2.5f + 1.5f * f);
    const float2 w2 = f * (0.5f + f * (2.0f - 1.5f * f));
    const float2 w3 = f * f * (-0.5f + 0.5f * f);
    const float2 w12 = w1 + w2;
    const float2 offset12 = w2 / (w1 + w2);
    const float2 texPos0 = (texPos1 - 1.0f) / texSize, texPos3 = (texPos1 + 2.0f) / texSize, texPos12 = (texPos1 + offset12) / texSize;
    float4 result = f4(0.0f);
    result += smp(f2(texPos12.x, texPos0.y)) * w12.x * w0.y;
    result += smp(f2(texPos0.x, texPos12.y)) * w0.x * w12.y;
    result += smp(f2(texPos12.x, texPos12.y)) * w12.x * w12.y;
    result += smp(f2(texPos3.x, texPos12.y)) * w3.x * w12.y;
    result += smp(f2(texPos12.x, texPos3.y)) * w12.x * w3.y;
    result = result / (w12.x * w0.y + w0.x * w12.y + w12.x * w12.y + w3.x * w12.y + w12.x * w3.y);
    st_rgba16f(output_tex, x, y, f4(xyz(result), kjb_max(0.0f, result.w)));
}

// ------------------------------------------------------------------ T2 filter_input.hlsl:32-89
struct FilteredInput { float3 clamped_ex, var; };
KJB_DEV FilteredInput t2_inner(const Img& input_tex, const Img& depth_tex, int px, int py, float center_depth, float luma_cutoff, float depth_scale, const float* dw) {
    float3 iex = f3(0.0f), iex2 = f3(0.0f), clamped_iex = f3(0.0f); float iwsum = 0, clamped_iwsum = 0;
    for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
        const float3 s = taa_input_remap(ld_rgba16f(input_tex, px + x, py + y));
        const float depth = ld_r32f(depth_tex, px + x, py + y);
        float w = 1;
        w *= kjb_exp2(-kjb_min(16.0f, depth_scale * inverse_depth_relative_diff(center_depth, depth)));
        w *= dw[(y + 1) * 3 + (x + 1)];
        w *= kjb_pow(kjb_saturate(luma_cutoff / s.x), 8.0f);
        clamped_iwsum += w; clamped_iex = mad(s, w, clamped_iex);
        iwsum += 1; iex += s; iex2 += s * s;
    }
    FilteredInput r; r.clamped_ex = clamped_iex / clamped_iwsum;
    iex = iex / iwsum; iex2 = iex2 / iwsum;
    r.var = vmax(f3(0.0f), iex2 - iex * iex);
    return r;
}
KJB_KERNEL(256) k_taa_filter_input(Img input_tex, Img depth_tex, ImgW output_tex, ImgW dev_output_tex, W9 dw, Rows kjb_rows) {
    KJB_PX; if (x >= output_tex.w || y >= output_tex.h) return;
    const float center_depth = ld_r32f(depth_tex, x, y);
    const FilteredInput fi = t2_inner(input_tex, depth_tex, x, y, center_depth, 1e10f, 200.0f, dw.w);
    const FilteredInput cfi = t2_inner(input_tex, depth_tex, x, y, center_depth, fi.clamped_ex.x * 1.001f, 200.0f, dw.w);
    st_rgba16f(output_tex, x, y, f4(cfi.clamped_ex, 0));
    st_rgba16f(dev_output_tex, x, y, f4(vsqrt(fi.var), 0));
}

// ------------------------------------------------------------------ T3 filter_history.hlsl:15-62
KJB_DEV float3 t3_filter(const Img& input_tex, float2 uv, float4 its, float luma_cutoff, int k, const float* dw) {
    float3 iex = f3(0.0f); float iwsum = 0;
    const int sx = kjb_cvt_i32(kjb_floor(uv.x * its.x + 1e-3f)), sy = kjb_cvt_i32(kjb_floor(uv.y * its.y + 1e-3f));
    for (int y = -k; y <= k; ++y) for (int x = -k; x <= k; ++x) {
        const float3 s = rgb_to_ycbcr(xyz(ld_rgba16f(input_tex, sx + x, sy + y)));
        float w = 1;
        w *= dw[(y + 2) * 5 + (x + 2)];
        w *= kjb_pow(kjb_saturate(luma_cutoff / s.x), 8.0f);
        iwsum += w; iex = mad(s, w, iex);
    }
    return iex / iwsum;
}
KJB_KERNEL(256) k_taa_filter_history(Img input_tex, ImgW output_tex, float4 its, float4 ots, int k, W25t dw, Rows kjb_rows) {
    KJB_PX; if (x >= output_tex.w || y >= output_tex.h) return;
    const float s4[4] = {ots.x, ots.y, ots.z, ots.w};
    const float2 uv = get_uv(x, y, s4);
    const float filtered_luma = t3_filter(input_tex, uv, its, 1e10f, k, dw.w).x;
    st_rgba16f(output_tex, x, y, f4(t3_filter(input_tex, uv, its, filtered_luma * 1.001f, k, dw.w), 0));
}

// ------------------------------------------------------------------ T4 input_prob.hlsl:47-109
KJB_KERNEL(256) k_taa_input_prob(Globals g, Img filtered_input_tex, Img filtered_input_dev_tex, Img filtered_history_tex, Img reprojection_tex, Img smooth_var_history_tex,
                                 Img velocity_history_tex, ImgW output_tex, float4 its, Rows kjb_rows) {
    KJB_PX; if (x >= output_tex.w || y >= output_tex.h) return;
    float input_prob = 0;
    float3 ivar = f3(0.0f);
    for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) ivar = vmax(ivar, xyz(ld_rgba16f(filtered_input_dev_tex, x + xx * 2, y + yy * 2)));
    ivar = ivar * ivar;
    const float2 sop = f2(g.fc.view_constants.sample_offset_pixels[0], g.fc.view_constants.sample_offset_pixels[1]);
    const float2 input_uv = (f2(float(x), float(y)) + sop) * f2(its.z, its.w);
    const int2 hp = nearest_clamp_px(filtered_history_tex, input_uv);
    const float3 closest_history = xyz(ld_rgba16f(filtered_history_tex, hp.x, hp.y));
    const float2 rxy = xy(ld_rgba16s(reprojection_tex, x, y));
    const Img& sv = smooth_var_history_tex; const Img& vh = velocity_history_tex;
    const float3 closest_smooth_var = xyz(bilinear_clamp(sv.w, sv.h, input_uv + rxy, [&](int sx, int sy) { return ld_rgba16f(sv, sx, sy); }));
    const float4 cv4 = bilinear_clamp(vh.w, vh.h, input_uv + rxy, [&](int sx, int sy) { const float2 v = ld_rg16f(vh, sx, sy); return f4(v.x, v.y, 0, 0); });
    const float2 closest_vel = f2(cv4.x, cv4.y) * g.fc.delta_time_seconds;
    const float3 combined_var = vmin(closest_smooth_var, ivar * 10.0f);
    for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
        const float3 s = xyz(ld_rgba16f(filtered_input_tex, x + xx, y + yy));
        const float3 idiff = s - closest_history;
        const float2 vel = xy(ld_rgba16s(reprojection_tex, x + xx, y + yy));
        const float vdiff = length((vel - closest_vel) / vmax(f2(1.0f), vabs(vel + closest_vel)));
        const float prob = kjb_exp2(-1.0f * length(idiff * idiff / vmax(f3(1e-6f), combined_var)) - 1000 * vdiff);
        input_prob = kjb_max(input_prob, prob);
    }
    st_raw<uint16_t>(output_tex, x, y, uint16_t(kjb_f32_to_f16(input_prob)));
}

// ------------------------------------------------------------------ T5 filter_prob.hlsl / T6 filter_prob2.hlsl
KJB_KERNEL(256) k_taa_prob_filter(Img input_tex, ImgW output_tex, Rows kjb_rows) {
    KJB_PX; if (x >= output_tex.w || y >= output_tex.h) return;
    float prob = ld_r16f(input_tex, x, y);
    for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) prob = kjb_max(prob, ld_r16f(input_tex, x + xx, y + yy));
    st_raw<uint16_t>(output_tex, x, y, uint16_t(kjb_f32_to_f16(prob)));
}
KJB_KERNEL(256) k_taa_prob_filter2(Img input_tex, ImgW output_tex, Rows kjb_rows) {
    KJB_PX; if (x >= output_tex.w || y >= output_tex.h) return;
    float2 weighted_prob = f2(0.0f);
    for (int yy = -2; yy <= 2; ++yy) for (int xx = -2; xx <= 2; ++xx) {
        const float neighbor_prob = ld_r16f(input_tex, x + xx * 2, y + yy * 2);
        weighted_prob += f2(kjb_exp2(-kjb_clamp(10.0f * neighbor_prob, 0.0f, 100.0f)), 1);     // exponential_squish
    }
    const float prob = kjb_max(0.0f, -1.0f / 10.0f * kjb_log2(1e-30f + weighted_prob.x / weighted_prob.y));   // exponential_unsquish
    st_raw<uint16_t>(output_tex, x, y, uint16_t(kjb_f32_to_f16(prob)));
}

// ------------------------------------------------------------------ T7 taa.hlsl:94-338 (+ inc/unjitter_taa.hlsl:58-125)
struct Unjittered { float4 color; float coverage; float3 ex, ex2; };
KJB_DEV Unjittered sample_image_unjitter_taa(const Img& img, int ox, int oy, float2 output_tex_size, float2 sample_offset_pixels, float kernel_scale) {
    const float2 irs = f2(float(img.w), float(img.h)) / output_tex_size;
    const int bx = kjb_cvt_i32((float(ox) + 0.5f) * irs.x), by = kjb_cvt_i32((float(oy) + 0.5f) * irs.y);
    const float2 dst_sample_loc = f2(float(ox), float(oy)) + 0.5f;
    const float2 base_src_sample_loc = (f2(float(bx), float(by)) + 0.5f + sample_offset_pixels * f2(1, -1)) / irs;
    float4 res = f4(0.0f); float3 ex = f3(0.0f), ex2 = f3(0.0f); float dev_wt_sum = 0.0f, wt_sum = 0.0f;
    const float kdm = 1.0f * kernel_scale;
    for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
        const float2 src_sample_loc = base_src_sample_loc + f2(float(x), float(y)) / irs;
        const float4 col = f4(taa_input_remap(ld_rgba16f(img, bx + x, by + y)), 1);
        const float2 sco = (src_sample_loc - dst_sample_loc) * kdm;
        const float dist2 = dot(sco, sco);
        const float dev_wt = kjb_exp2(-dist2 * irs.x);
        const float wt = kjb_exp2(-10 * dist2 * irs.x);
        res = mad(col, wt, res); wt_sum += wt;
        ex = mad(xyz(col), dev_wt, ex); ex2 = mad(xyz(col) * xyz(col), dev_wt, ex2); dev_wt_sum += dev_wt;
    }
    Unjittered u; u.color = res; u.coverage = wt_sum; u.ex = ex / dev_wt_sum; u.ex2 = ex2 / dev_wt_sum;
    return u;
}
struct TaaImgs { Img input_tex, history_tex, reprojection_tex, closest_velocity_tex, velocity_history_tex, smooth_var_history_tex, input_prob_tex;
                 ImgW temporal_output_tex, output_tex, smooth_var_output_tex, velocity_output_tex; };
KJB_KERNEL(256) k_taa(Globals g, TaaImgs t, float4 its, float4 ots, W25t bw, Rows kjb_rows) {
    KJB_PX; if (x >= t.temporal_output_tex.w || y >= t.temporal_output_tex.h) return;
    const float2 sop = f2(g.fc.view_constants.sample_offset_pixels[0], g.fc.view_constants.sample_offset_pixels[1]);
    const float dt = g.fc.delta_time_seconds;
    const float2 irf = f2(its.x, its.y) / f2(ots.x, ots.y);
    const int rx = int(kjb_cvt_u32((float(x) + 0.5f) * irf.x)), ry = int(kjb_cvt_u32((float(y) + 0.5f) * irf.y));
    const float s4[4] = {ots.x, ots.y, ots.z, ots.w};
    const float2 uv = get_uv(x, y, s4);
    const float4 history_packed = ld_rgba16f(t.history_tex, x, y);
    float3 history = xyz(history_packed);
    float history_coverage = kjb_max(0.0f, history_packed.w);
    float4 bhistory_packed;
    {   // fetch_blurred_history(px, 2, 1): w = exp(-r^2), host-evaluated table
        float4 csum = f4(0.0f); float wsum = 0;
        for (int yy = -2; yy <= 2; ++yy) for (int xx = -2; xx <= 2; ++xx) {
            const float w = bw.w[(yy + 2) * 5 + (xx + 2)];
            csum = mad(ld_rgba16f(t.history_tex, x + xx, y + yy), w, csum); wsum += w;
        }
        bhistory_packed = csum / wsum;
    }
    float3 bhistory = xyz(bhistory_packed);
    const float3 bhistory_coverage = f3(bhistory_packed.w);
    history = rgb_to_ycbcr(history); bhistory = rgb_to_ycbcr(bhistory);
    const float4 reproj = ld_rgba16s(t.reprojection_tex, rx, ry);
    const float2 cvel = ld_rg16f(t.closest_velocity_tex, x, y);
    const float2 reproj_xy = cvel;
    const Unjittered center_sample = sample_image_unjitter_taa(t.input_tex, x, y, f2(ots.x, ots.y), sop, 1.0f);
    const Unjittered bcenter_sample = sample_image_unjitter_taa(t.input_tex, x, y, f2(ots.x, ots.y), sop, 0.333f);
    float coverage = center_sample.coverage;
    float3 center = xyz(center_sample.color);
    const float3 bcenter = xyz(bcenter_sample.color) / bcenter_sample.coverage;
    history = vlerp(history, bcenter, kjb_saturate(1.0f - history_coverage));
    bhistory = vlerp(bhistory, bcenter, f3(kjb_saturate(1.0f - bhistory_coverage.x), kjb_saturate(1.0f - bhistory_coverage.y), kjb_saturate(1.0f - bhistory_coverage.z)));
    const float input_prob = ld_r16f(t.input_prob_tex, rx, ry);
    const float3 ex = center_sample.ex, ex2 = center_sample.ex2;
    const float3 var = vmax(f3(0.0f), ex2 - ex * ex);
    const Img& sv = t.smooth_var_history_tex; const Img& vh = t.velocity_history_tex;
    const float3 prev_var = f3(bilinear_clamp(sv.w, sv.h, uv + reproj_xy, [&](int sx, int sy) { return ld_rgba16f(sv, sx, sy); }).x);
    const float2 vel_now = cvel / dt;
    const float4 vp4 = bilinear_clamp(vh.w, vh.h, uv + cvel, [&](int sx, int sy) { const float2 v = ld_rg16f(vh, sx, sy); return f4(v.x, v.y, 0, 0); });
    const float2 vel_prev = f2(vp4.x, vp4.y);
    const float vel_diff = length((vel_now - vel_prev) / vmax(f2(1.0f), vabs(vel_now + vel_prev)));
    const float var_blend = kjb_saturate(0.3f + 0.7f * (1 - reproj.z) + vel_diff);
    float3 smooth_var = vmax(var, vlerp(prev_var, var, var_blend));
    smooth_var = vlerp(var, smooth_var, kjb_saturate(input_prob));
    const float3 input_dev = vsqrt(var);
    float3 clamped_history;
    {
        const float box_n_deviations = kjb_lerp(0.8f, 3.0f, input_prob);
        const float3 nmin = ex - input_dev * box_n_deviations, nmax = ex + input_dev * box_n_deviations;
        const float3 clamped_bhistory = vclamp(bhistory, nmin, nmax);
        const float clamping_event = length(vmax(f3(0.0f), vmax(bhistory - nmax, nmin - bhistory)) / vmax(f3(0.01f), ex));
        const float3 outlier3 = vmax(f3(0.0f), (vmax(nmin - history, history - nmax)) / (0.1f + vmax(vmax(vabs(history), vabs(ex)), f3(1e-5f))));
        const float3 boutlier3 = vmax(f3(0.0f), (vmax(nmin - bhistory, bhistory - nmax)) / (0.1f + vmax(vmax(vabs(bhistory), vabs(ex)), f3(1e-5f))));
        const float outlier = kjb_max(outlier3.x, kjb_max(outlier3.y, outlier3.z));
        const float boutlier = kjb_max(boutlier3.x, kjb_max(boutlier3.y, boutlier3.z));
        const float2 huv = uv + reproj_xy, hs = vsaturate(huv);
        if (huv.x == hs.x && huv.y == hs.y) {
            const float non_disoccluding_outliers = kjb_max(0.0f, outlier - boutlier) * 10;
            const float3 unclamped_history_detail = history - clamped_bhistory;
            const float temporal_clamping_detail = kjb_abs(unclamped_history_detail.x / kjb_max(1e-3f, input_dev.x)) * 0.05f;
            const float temporal_stability = kjb_saturate(1 - temporal_clamping_detail);
            const float allow_unclamped_detail = kjb_saturate(non_disoccluding_outliers) * temporal_stability;
            float3 history_detail = history - bhistory;
            history_detail = vlerp(history_detail, unclamped_history_detail, allow_unclamped_detail);
            const float initial_bclamp_amount = kjb_saturate(dot(clamped_bhistory - bhistory, bcenter - bhistory)
                / kjb_max(1e-5f, length(clamped_bhistory - bhistory) * length(bcenter - bhistory)));
            const float effective_clamp_amount = kjb_saturate(initial_bclamp_amount) * (1 - allow_unclamped_detail);
            const float keep_detail = 1 - effective_clamp_amount;
            history_detail *= keep_detail;
            clamped_history = clamped_bhistory + history_detail;
            if (irf.x < 1.0f) history_coverage *= kjb_lerp(kjb_lerp(0.0f, 0.9f, keep_detail), 1.0f, kjb_saturate(10 * clamping_event));
        } else {
            clamped_history = clamped_bhistory; coverage = 1; center = bcenter; history_coverage = 0;
        }
        clamped_history = vlerp(clamped_history, history, kjb_smoothstep(0.5f, 1.0f, input_prob));
    }
    float total_coverage = kjb_max(1e-5f, history_coverage + coverage);
    float3 temporal_result = (clamped_history * history_coverage + center) / total_coverage;
    const float max_coverage = kjb_max(2.0f, 8.0f / (irf.x * irf.y));
    total_coverage = kjb_min(max_coverage, total_coverage);
    coverage = total_coverage;
    st_rgba16f(t.smooth_var_output_tex, x, y, f4(smooth_var, 0));
    temporal_result = ycbcr_to_rgb(temporal_result);
    temporal_result = taa_encode_rgb(temporal_result);
    temporal_result = vmax(f3(0.0f), temporal_result);
    st_rgba16f(t.temporal_output_tex, x, y, f4(temporal_result, coverage));
    st_rgba16f(t.output_tex, x, y, f4(temporal_result, 0));
    const float2 vo = cvel / dt;
    st_rg16f(t.velocity_output_tex, x, y, vo.x, vo.y);
}

#define F4A(a) f4((a)[0], (a)[1], (a)[2], (a)[3])
#define CHK(img, fmt, name) if (!check_img(c, (img), (fmt), P, name)) return 1
#define CHKE(img, fmt, name, w, h) if (!check_img(c, (img), (fmt), P, name, (w), (h))) return 1

extern "C" {

int kjb_pass_taa_reproject(kjb_context* c, const kjb_taa_reproject_args* a) {
    const char* P = "reproject taa"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_RGBA16_FLOAT, "output_tex"); CHKE(a->history_tex, KJB_FMT_RGBA16_FLOAT, "history_tex", W, H); CHK(a->reprojection_tex, KJB_FMT_RGBA16_SNORM, "reprojection_tex");
    CHK(a->depth_tex, KJB_FMT_R32_FLOAT, "depth_tex"); CHKE(a->closest_velocity_output, KJB_FMT_RG16_FLOAT, "closest_velocity_output", W, H);
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa_reproject, KJB_GRID2D(W, H, 32, 8), c->g, img_ro(a->history_tex), img_ro(a->reprojection_tex), img_ro(a->depth_tex), img_rw(a->output_tex), img_rw(a->closest_velocity_output),
               F4A(a->input_tex_size), F4A(a->output_tex_size));
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_taa_filter_input(kjb_context* c, const kjb_taa_filter_input_args* a) {
    const char* P = "taa filter input"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_RGBA16_FLOAT, "output_tex"); CHKE(a->input_tex, KJB_FMT_RGBA16_FLOAT, "input_tex", W, H); CHKE(a->depth_tex, KJB_FMT_R32_FLOAT, "depth_tex", W, H);
    CHKE(a->dev_output_tex, KJB_FMT_RGBA16_FLOAT, "dev_output_tex", W, H);
    W9 dw; for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) dw.w[(y + 1) * 3 + (x + 1)] = kjb_exp(-(0.8f / float(1 * 1)) * float(x * x + y * y));
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa_filter_input, KJB_GRID2D(W, H, 32, 8), img_ro(a->input_tex), img_ro(a->depth_tex), img_rw(a->output_tex), img_rw(a->dev_output_tex), dw);
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_taa_filter_history(kjb_context* c, const kjb_taa_filter_history_args* a) {
    const char* P = "taa filter history"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_RGBA16_FLOAT, "output_tex"); CHK(a->input_tex, KJB_FMT_RGBA16_FLOAT, "input_tex");
    const int k = (a->input_tex_size[0] / a->output_tex_size[0] > 1.75f) ? 2 : 1;
    W25t dw; for (int y = -2; y <= 2; ++y) for (int x = -2; x <= 2; ++x) dw.w[(y + 2) * 5 + (x + 2)] = kjb_exp(-(0.8f / float(k * k)) * float(x * x + y * y));
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa_filter_history, KJB_GRID2D(W, H, 32, 8), img_ro(a->input_tex), img_rw(a->output_tex), F4A(a->input_tex_size), F4A(a->output_tex_size), k, dw);
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_taa_input_prob(kjb_context* c, const kjb_taa_input_prob_args* a) {
    const char* P = "taa input prob"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_R16_FLOAT, "output_tex"); CHKE(a->filtered_input_tex, KJB_FMT_RGBA16_FLOAT, "filtered_input_tex", W, H); CHKE(a->filtered_input_dev_tex, KJB_FMT_RGBA16_FLOAT, "filtered_input_dev_tex", W, H);
    CHKE(a->filtered_history_tex, KJB_FMT_RGBA16_FLOAT, "filtered_history_tex", W, H); CHKE(a->reprojection_tex, KJB_FMT_RGBA16_SNORM, "reprojection_tex", W, H);
    CHK(a->smooth_var_history_tex, KJB_FMT_RGBA16_FLOAT, "smooth_var_history_tex"); CHK(a->velocity_history_tex, KJB_FMT_RG16_FLOAT, "velocity_history_tex");
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa_input_prob, KJB_GRID2D(W, H, 32, 8), c->g, img_ro(a->filtered_input_tex), img_ro(a->filtered_input_dev_tex), img_ro(a->filtered_history_tex), img_ro(a->reprojection_tex),
               img_ro(a->smooth_var_history_tex), img_ro(a->velocity_history_tex), img_rw(a->output_tex), F4A(a->input_tex_size));
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_taa_prob_filter(kjb_context* c, const kjb_taa_prob_filter_args* a) {
    const char* P = "taa prob filter"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_R16_FLOAT, "output_tex"); CHKE(a->input_tex, KJB_FMT_R16_FLOAT, "input_tex", W, H);
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa_prob_filter, KJB_GRID2D(W, H, 32, 8), img_ro(a->input_tex), img_rw(a->output_tex));
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_taa_prob_filter2(kjb_context* c, const kjb_taa_prob_filter_args* a) {
    const char* P = "taa prob filter2"; const uint32_t W = a->output_tex.width, H = a->output_tex.height;
    CHK(a->output_tex, KJB_FMT_R16_FLOAT, "output_tex"); CHKE(a->input_tex, KJB_FMT_R16_FLOAT, "input_tex", W, H);
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa_prob_filter2, KJB_GRID2D(W, H, 32, 8), img_ro(a->input_tex), img_rw(a->output_tex));
    KJB_PASS_EPILOGUE(c, P);
}
int kjb_pass_taa(kjb_context* c, const kjb_taa_args* a) {
    const char* P = "taa"; const uint32_t W = a->temporal_output_tex.width, H = a->temporal_output_tex.height;
    CHK(a->temporal_output_tex, KJB_FMT_RGBA16_FLOAT, "temporal_output_tex"); CHK(a->input_tex, KJB_FMT_RGBA16_FLOAT, "input_tex"); CHKE(a->history_tex, KJB_FMT_RGBA16_FLOAT, "history_tex", W, H);
    CHK(a->reprojection_tex, KJB_FMT_RGBA16_SNORM, "reprojection_tex"); CHKE(a->closest_velocity_tex, KJB_FMT_RG16_FLOAT, "closest_velocity_tex", W, H);
    CHKE(a->velocity_history_tex, KJB_FMT_RG16_FLOAT, "velocity_history_tex", W, H); CHKE(a->smooth_var_history_tex, KJB_FMT_RGBA16_FLOAT, "smooth_var_history_tex", W, H);
    CHK(a->input_prob_tex, KJB_FMT_R16_FLOAT, "input_prob_tex"); CHKE(a->output_tex, KJB_FMT_RGBA16_FLOAT, "output_tex", W, H);
    CHKE(a->smooth_var_output_tex, KJB_FMT_RGBA16_FLOAT, "smooth_var_output_tex", W, H); CHKE(a->velocity_output_tex, KJB_FMT_RG16_FLOAT, "velocity_output_tex", W, H);
    TaaImgs t;
    t.input_tex = img_ro(a->input_tex); t.history_tex = img_ro(a->history_tex); t.reprojection_tex = img_ro(a->reprojection_tex); t.closest_velocity_tex = img_ro(a->closest_velocity_tex);
    t.velocity_history_tex = img_ro(a->velocity_history_tex); t.smooth_var_history_tex = img_ro(a->smooth_var_history_tex); t.input_prob_tex = img_ro(a->input_prob_tex);
    t.temporal_output_tex = img_rw(a->temporal_output_tex); t.output_tex = img_rw(a->output_tex); t.smooth_var_output_tex = img_rw(a->smooth_var_output_tex); t.velocity_output_tex = img_rw(a->velocity_output_tex);
    W25t bw; for (int y = -2; y <= 2; ++y) for (int x = -2; x <= 2; ++x) { const float ox = float(x) * 1.0f, oy = float(y) * 1.0f; bw.w[(y + 2) * 5 + (x + 2)] = kjb_exp(-(ox * ox + oy * oy)); }
    KJB_ROWS(c, H);
    KJB_LAUNCH(c, k_taa, KJB_GRID2D(W, H, 32, 8), c->g, t, F4A(a->input_tex_size), F4A(a->output_tex_size), bw);
    KJB_PASS_EPILOGUE(c, P);
}

}  // extern "C"
