// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// The shadow denoiser of renderers/shadow_denoise.rs: "shadow bitpack" (bitpack_shadow_mask.hlsl + ffx_denoiser_shadows_prepare.hlsl),
// "shadow temporal" (megakernel.hlsl + ffx_denoiser_shadows_tileclassification.hlsl), "shadow spatial" (spatial_filter.hlsl +
// ffx_denoiser_shadows_filter.hlsl).  Written per pixel: the shaders' group-shared staging and wave reductions only redistribute values
// that are functions of pixel / group coordinates (the thread remap FFX_DNSR_Shadows_RemapLane8x8 is a permutation inside an 8x8 group).
#include "kj_ctx.h"

namespace kjo {
namespace {
inline uint rounded_divide(uint v, uint d) { return (v + d - 1) / d; }                                       // ffx_denoiser_shadows_util.hlsl:26-29
inline uint linear_tile_index(uint tx, uint ty, uint screen_width) { return ty * rounded_divide(screen_width, 8) + tx; }   // :36-39
struct TileImg {   // WriteMask / ReadRaytracedShadowMask / Write-ReadTileMetaData: linear index -> texel of the extent-sized image
    Img img; uint ext_x;
    uint read(uint linear) const { return img.load_u(int(linear % ext_x), int(linear / ext_x)).x; }
    void write(uint linear, uint v) const { img.store_u(int(linear % ext_x), int(linear / ext_x), uint4(v, 0, 0, 0)); }
};
struct KernelWeights { float k[9]; };
KernelWeights kernel_weights() {   // FFX_DNSR_Shadows_KernelWeight, tileclassification.hlsl:169-184 (KERNEL_RADIUS 8)
    auto W = [](int i) { return exp(-3.0f * float(i * i) / ((8 + 1.0f) * (8 + 1.0f))); };
    float sum = 0; sum += W(0);
    for (int c = 1; c <= 8; ++c) sum += 2 * W(c);
    const float inv = rcp(sum);
    KernelWeights r; for (int i = 0; i <= 8; ++i) r.k[i] = W(i) * inv;
    return r;
}
float4 cubic_hermite(float4 A, float4 B, float4 C, float4 D, float t) {   // inc/curve.hlsl:4-13
    float t2 = t * t, t3 = t * t * t;
    float4 a = -A / 2.0f + (3.0f * B) / 2.0f - (3.0f * C) / 2.0f + D / 2.0f;
    float4 b = A - (5.0f * B) / 2.0f + 2.0f * C - D / 2.0f;
    float4 c = -A / 2.0f + C / 2.0f;
    return a * t3 + b * t2 + c * t + B;
}
float4 sample_catmull_rom(const Img& tex, float2 size, float2 P) {   // inc/image.hlsl:42-79, identity remap, fetch = integer load (0 outside)
    float2 pixel = P * size + 0.5f;
    float2 frc = frac(pixel);
    int2 ip(kjb_cvt_i32(pixel.x) - 1, kjb_cvt_i32(pixel.y) - 1);
    float4 rows[4];
    for (int j = 0; j < 4; ++j)
        rows[j] = cubic_hermite(tex.load(ip.x - 1, ip.y - 1 + j), tex.load(ip.x, ip.y - 1 + j), tex.load(ip.x + 1, ip.y - 1 + j), tex.load(ip.x + 2, ip.y - 1 + j), frc.x);
    return cubic_hermite(rows[0], rows[1], rows[2], rows[3], frc.y);
}
float soft_color_clamp1(float center, float history, float ex, float dev) {   // inc/soft_color_clamp.hlsl:1-14, scalar
    float history_dist = abs(history - ex) / max(abs(history * 0.1f), dev);
    float closest_pt = clamp(history, center - dev, center + dev);
    return lerp(history, closest_pt, smoothstep(1.0f, 3.0f, history_dist));
}
}  // namespace

extern "C" {

int kjb_pass_shadow_bitpack(kjb_context* ctx, const kjb_shadow_bitpack_args* a) {
    Img input_tex(a->input_tex); TileImg out{Img(a->output_tex), a->bitpacked_shadow_mask_extent[0]};
    const uint W = uint(a->input_tex_size[0]), H = uint(a->input_tex_size[1]);
    const uint tiles_x = (W + 7) / 8, tiles_y = (H + 3) / 4;
    pass_rows(ctx, int(tiles_y), [&](int ty) { for (uint tx = 0; tx < tiles_x; ++tx) {
        uint mask = 0;   // WaveActiveBitOr over the tile's 8x4 lanes (prepare.hlsl:28-35)
        for (uint ly = 0; ly < 4; ++ly) for (uint lx = 0; lx < 8; ++lx) {
            const uint px = tx * 8 + lx, py = uint(ty) * 4 + ly;
            if (input_tex.load(int(px), int(py)).x > 0.5f) mask |= 1u << ((py % 4) * 8 + (px % 8));
        }
        out.write(linear_tile_index(tx, uint(ty), W), mask);
    } }, ctx->num_threads);
    return 0;
}

int kjb_pass_shadow_temporal(kjb_context* ctx, const kjb_shadow_temporal_args* a) {
    const Globals& g = ctx->g;
    Img shadow_mask_tex(a->shadow_mask_tex), prev_moments_tex(a->prev_moments_tex), prev_accum_tex(a->prev_accum_tex), reprojection_tex(a->reprojection_tex),
        output_moments_tex(a->output_moments_tex), temporal_output_tex(a->temporal_output_tex);
    const TileImg bitpacked{Img(a->bitpacked_shadow_mask_tex), a->bitpacked_shadow_mask_extent[0]}, meta{Img(a->meta_output_tex), a->bitpacked_shadow_mask_extent[0]};
    const float4 input_tex_size = f4(a->input_tex_size);
    const uint W = uint(input_tex_size.x), H = uint(input_tex_size.y);
    const KernelWeights kw = kernel_weights();
    const int groups_x = int((W + 7) / 8), groups_y = int((H + 7) / 8);

    auto write_moments = [&](int x, int y, float4 m) { m.z = min(m.z, 32.0f); output_moments_tex.store(x, y, m); };   // megakernel.hlsl:93-99
    // FFX_DNSR_Shadows_HorizontalNeighborhood, tileclassification.hlsl:193-252
    auto horizontal_neighborhood = [&](int dx, int dy) -> float {
        if (dy < 0 || dy >= int(H)) return 0.0f;
        const uint tile_x = uint(dx) / 8, tile_y = uint(dy) / 4;
        const uint lin = linear_tile_index(tile_x, tile_y, W);
        const bool first = tile_x == 0, last = tile_x == rounded_divide(W, 8) - 1;
        uint left_tile = 0; if (!first) left_tile = bitpacked.read(uint(int(lin) - 1));
        const uint center_tile = bitpacked.read(lin);
        uint right_tile = 0; if (!last) right_tile = bitpacked.read(uint(int(lin) + 1));
        const uint row_base = (uint(dy) % 4) * 8;
        const uint left = (left_tile >> row_base) & 0xFF, center = (center_tile >> row_base) & 0xFF, right = (right_tile >> row_base) & 0xFF;
        uint neighborhood = left | (center << 8) | (right << 16);
        neighborhood = neighborhood >> (uint(dx) % 8);
        float moment = 0.0f;
        for (int i = 0; i < 8; ++i) moment += (neighborhood & (1u << i)) ? kw.k[8 - i] : 0.0f;
        moment += (neighborhood & (1u << 8)) ? kw.k[0] : 0.0f;
        for (int i = 1; i <= 8; ++i) moment += (neighborhood & (1u << (8 + i))) ? kw.k[i] : 0.0f;
        return moment;
    };

    pass_rows(ctx, groups_y, [&](int gy) { for (int gx = 0; gx < groups_x; ++gx) {
        // FFX_DNSR_Shadows_SearchSpatialRegion (:48-86): IsShadowReciever is the constant true upstream, so only this decides
        const int base_tx = gx, base_ty = gy * 2;
        uint combined_or = 0, combined_and = 0xFFFFFFFFu;
        for (int j = -2; j <= 3; ++j) for (int i = -1; i <= 1; ++i) {
            const int tx = clamp(base_tx + i, 0, int(rounded_divide(W, 8)) - 1), ty = clamp(base_ty + j, 0, int(rounded_divide(H, 4)) - 1);
            const uint m = bitpacked.read(linear_tile_index(uint(tx), uint(ty), W));
            combined_or |= m; combined_and &= m;
        }
        const bool all_in_light = combined_and == 0xFFFFFFFFu, all_in_shadow = combined_or == 0u;
        const uint meta_index = uint(gy) * rounded_divide(W, 8) + uint(gx);
        if (all_in_light || all_in_shadow) {   // FFX_DNSR_Shadows_ClearTargets (:300-313)
            const float shadow_value = all_in_light ? 1.0f : 0.0f;
            meta.write(meta_index, (all_in_light ? 2u : 0u) | 1u);
            for (int ly = 0; ly < 8; ++ly) for (int lx = 0; lx < 8; ++lx) {
                const int x = gx * 8 + lx, y = gy * 8 + ly;
                temporal_output_tex.store(x, y, float4(shadow_value, 0, 0, 0));
                write_moments(x, y, float4(shadow_value, 0, 8.0f, shadow_value));
            }
            continue;
        }
        meta.write(meta_index, 0u);
        for (int ly = 0; ly < 8; ++ly) for (int lx = 0; lx < 8; ++lx) {
            const int x = gx * 8 + lx, y = gy * 8 + ly;
            const float4 reproj = reprojection_tex.load(x, y);
            // FFX_DNSR_Shadows_ComputeLocalNeighborhood (:256-287): smem[gtid.x][8 + gtid.y -+ i] = the horizontal sums of rows y -+ i
            float local_neighborhood = 0;
            const float upper = horizontal_neighborhood(x, y - 8), center = horizontal_neighborhood(x, y), lower = horizontal_neighborhood(x, y + 8);
            local_neighborhood = mad(center, kw.k[0], local_neighborhood);
            local_neighborhood = mad(upper, kw.k[8], local_neighborhood);
            local_neighborhood = mad(lower, kw.k[8], local_neighborhood);
            for (int i = 1; i < 8; ++i) {
                local_neighborhood = mad(horizontal_neighborhood(x, y - i), kw.k[i], local_neighborhood);
                local_neighborhood = mad(horizontal_neighborhood(x, y + i), kw.k[i], local_neighborhood);
            }
            const float2 uv = (float2(float(x), float(y)) + 0.5f) * float2(input_tex_size.z, input_tex_size.w);
            const float2 history_uv = uv + reproj.xy();

            const float shadow_current = shadow_mask_tex.load(x, y).x;
            const uint quad_reproj_valid_packed = kjb_cvt_u32(reproj.z * 15.0f + 0.5f);
            const bool is_disoccluded = (quad_reproj_valid_packed & 15u) != 15u;   // dot(valid4, 1) < 4
            float4 previous_moments(0.0f);
            if (!is_disoccluded) {
                previous_moments = sample_catmull_rom(prev_moments_tex, float2(input_tex_size.x, input_tex_size.y), history_uv);
                previous_moments.y = max(0.0f, previous_moments.y); previous_moments.z = max(0.0f, previous_moments.z);
            }
            const float old_m = previous_moments.x, old_s = previous_moments.y;
            const float sample_count = previous_moments.z + 1.0f;
            const float new_m = lerp(old_m, shadow_current, 1.0f / sample_count);
            const float new_s = lerp(old_s, (shadow_current - old_m) * (shadow_current - new_m), 1.0f / sample_count);
            float variance = new_s;
            float4 moments_current(new_m, new_s, sample_count, local_neighborhood);

            const float mean = local_neighborhood;
            float spatial_variance = local_neighborhood;
            spatial_variance = max(spatial_variance - mean * mean, 0.0f);
            const float std_deviation = sqrt(spatial_variance);
            float shadow_previous = shadow_current;
            if (g.fc.frame_index != 0) shadow_previous = sample_catmull_rom(prev_accum_tex, float2(input_tex_size.x, input_tex_size.y), history_uv).x;
            const float sigma = 2.0f;
            const float temporal_discontinuity = (previous_moments.w - moments_current.w) / max(0.5f * std_deviation, 0.001f);
            const float sample_counter_damper = exp(-temporal_discontinuity * temporal_discontinuity / sigma);
            moments_current.z *= max(0.5f, sample_counter_damper);
            float shadow_clamped = soft_color_clamp1(shadow_current, shadow_previous, mean, std_deviation * 0.5f);
            if (moments_current.z < 16.0f) {
                const float variance_boost = max(16.0f - moments_current.z, 1.0f);
                variance = max(variance, spatial_variance);
                variance *= variance_boost;
            }
            shadow_clamped = lerp(shadow_clamped, shadow_current, 1.0f / max(1.0f, moments_current.z));
            temporal_output_tex.store(x, y, float4(shadow_clamped, variance, 0, 0));
            write_moments(x, y, moments_current);
        }
    } }, ctx->num_threads);
    return 0;
}

int kjb_pass_shadow_spatial(kjb_context* ctx, const kjb_shadow_spatial_args* a) {
    Img input_tex(a->input_tex), geometric_normal_tex(a->geometric_normal_tex), depth_tex(a->depth_tex), output_tex(a->output_tex);
    const TileImg meta{Img(a->meta_tex), a->bitpacked_shadow_mask_extent[0]};
    const uint W = uint(a->input_tex_size[0]), H = uint(a->input_tex_size[1]);
    const int step = int(a->step_size);
    const int groups_x = int((W + 7) / 8), groups_y = int((H + 7) / 8);
    // what the 16x16 group-shared tile holds for a pixel (filter.hlsl:33-92): clamped coordinates, normals and input through f32tof16
    auto q16 = [](float v) { return kjb_f16_to_f32(kjb_f32_to_f16(v)); };
    struct Tap { float3 normal; float2 input; float depth; };
    auto staged = [&](int x, int y) {
        const int px = clamp(x, 0, int(W) - 1), py = clamp(y, 0, int(H) - 1);
        const float3 n = geometric_normal_tex.load(px, py).xyz() * 2.0f - 1.0f;
        const float4 in = input_tex.load(px, py);
        Tap t; t.normal = float3(q16(n.x), q16(n.y), q16(n.z)); t.input = float2(q16(in.x), q16(in.y)); t.depth = depth_tex.load(px, py).x;
        return t;
    };
    pass_rows(ctx, groups_y, [&](int gy) { for (int gx = 0; gx < groups_x; ++gx) {
        const uint m = meta.read(uint(gy) * rounded_divide(W, 8) + uint(gx));
        const bool is_cleared = (m & 1u) != 0, all_in_light = (m & 2u) != 0;
        for (int ly = 0; ly < 8; ++ly) for (int lx = 0; lx < 8; ++lx) {
            const int x = gx * 8 + lx, y = gy * 8 + ly;
            float2 results(0.0f, 0.0f);
            if (is_cleared) {   // pass index is the constant 0 upstream (spatial_filter.hlsl:64): cleared tiles are written
                results.x = all_in_light ? 1.0f : 0.0f;
            } else {            // FFX_DNSR_Shadows_ApplyFilterWithPrecache (:216-236)
                float weight_sum = 1.0f; float2 shadow_sum(0.0f, 0.0f);
                if (depth_tex.load(x, y).x != 0.0f) {
                    const float depth_center = depth_tex.load(x, y).x;
                    const Tap c = staged(x, y);
                    weight_sum = 1.0f; shadow_sum = c.input;
                    const float variance = c.input.y;
                    const float std_deviation = sqrt(max(variance + 1e-9f, 0.0f));
                    const float sharp = max(0.0f, 1.0f - 2.0f * std_deviation);
                    const float kernel_sharpening = max(1e-10f, 1.0f - sharp * sharp);
                    const float kernel[3] = {1.0f, exp2(-0.5849625007211563f / kernel_sharpening), exp2(-2.584962500721156f / kernel_sharpening)};
                    for (int yy = -1; yy <= 1; ++yy) for (int xx = -1; xx <= 1; ++xx) {
                        const Tap n = staged(x + xx * step, y + yy * step);
                        const float sky_pixel_multiplier = ((xx == 0 && yy == 0) || n.depth >= 1.0f || n.depth <= 0.0f) ? 0.0f : 1.0f;
                        float w = kernel[xx < 0 ? -xx : xx] * kernel[yy < 0 ? -yy : yy];
                        w *= exp(-abs(c.input.x - n.input.x) / std_deviation);
                        w *= exp2(-abs(1.0f - (depth_center / n.depth)) / 0.01f);
                        w *= pow(saturate(dot(c.normal, n.normal)), 32.0f);
                        w *= sky_pixel_multiplier;
                        shadow_sum = shadow_sum + float2(w, w * w) * n.input;
                        weight_sum += w;
                    }
                }
                results = float2(shadow_sum.x / weight_sum, shadow_sum.y / (weight_sum * weight_sum));
            }
            output_tex.store(x, y, float4(max(0.0f, results.x), max(0.0f, results.y), 0, 0));
        }
    } }, ctx->num_threads);
    return 0;
}

}  // extern "C"
}  // namespace kjo
