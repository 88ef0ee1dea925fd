// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED (no reference goldens).
// Input producers and small utility passes.  Paths relative to /root/reference/assets/shaders/.
#include "kj_ctx.h"

namespace kjo {

extern "C" {

// ------------------------------------------------------------------ primary-visibility G-buffer by ray casting.
// Stands in for the raster pass (raster_simple_ps.hlsl:39-140): same outputs and encodings, geometry found with
// the ray/triangle contract instead of the rasteriser.  Hit shading = rt/gbuffer.rchit.hlsl (S1).
int kjb_pass_raster_gbuffer(kjb_context* ctx, const kjb_raster_gbuffer_args* a) {
    const Globals& g = ctx->g; const kjb_view_constants& vc = g.fc.view_constants;
    Img gn(a->geometric_normal_out), gb(a->gbuffer_out), dp(a->depth_out), vel(a->velocity_out);
    const int W = gb.w(), H = gb.h();
    const float4 size(float(W), float(H), 1.0f / float(W), 1.0f / float(H));
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const float2 uv = get_uv(int2(x, y), size);
        const ViewRayContext vrc = ViewRayContext::from_uv(vc, uv);
        Ray r; r.origin = vrc.ray_origin_ws(); r.dir = vrc.ray_dir_ws(); r.tmin = 0; r.tmax = FLT_MAX_F;
        Scene::HitInfo h = ctx->scene.closest(r, false);
        if (!h.hit) {
            gn.store(x, y, float4(0.0f)); gb.store_u(x, y, uint4(0, 0, 0, 0)); dp.store(x, y, float4(0.0f)); vel.store(x, y, float4(0.0f));
            continue;
        }
        RayCone cone = RayCone::from_spread_angle(pixel_cone_spread_angle_from_image_height(vc, float(H)));
        uint4 packed = rchit_gbuffer(ctx->scene, g, r, h, cone, 0);
        const float3 pos_ws = r.origin + r.dir * h.t;
        const float3 pos_cs = position_world_to_clip(vc, pos_ws);
        const WorldTri& wt = ctx->scene.tris[h.tri];
        float3 gnorm_ws = normalize(cross(wt.e1, wt.e2));
        if (dot(gnorm_ws, r.dir) > 0) gnorm_ws = -gnorm_ws;
        const float3 gnorm_vs = normalize(direction_world_to_view(vc, gnorm_ws));
        const float3 vs_pos = mul(vc.world_to_view, float4(pos_ws, 1)).xyz();
        float3 prev_pos_ws = pos_ws;
        if (a->prev_instances && wt.instance < a->prev_instance_count) {   // object motion: the same surface point under last frame's transform
            const kjb_instance& inst = ctx->scene.instances[wt.instance];
            const kjb_gpu_mesh& mesh = ctx->scene.meshes[inst.mesh_index];
            float3 p[3];
            for (int k = 0; k < 3; ++k) p[k] = ctx->scene.load_f3(mesh.vertex_core_offset + ctx->scene.load_u32(mesh.index_offset + (wt.prim * 3 + k) * 4) * 16);
            const float3 p_obj = p[0] * (1.0f - h.u - h.v) + p[1] * h.u + p[2] * h.v;
            prev_pos_ws = Scene::xform_point(a->prev_instances[wt.instance].transform, p_obj);
        }
        const float3 prev_vs_pos = mul(vc.prev_world_to_prev_view, float4(prev_pos_ws, 1)).xyz();
        gn.store(x, y, float4(gnorm_vs * 0.5f + 0.5f, 0));
        gb.store_u(x, y, packed);
        dp.store(x, y, float4(pos_cs.z));
        vel.store(x, y, float4(prev_vs_pos - vs_pos, 0));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ calculate_reprojection_map.hlsl:17-142
int kjb_pass_reprojection_map(kjb_context* ctx, const kjb_reprojection_map_args* a) {
    const kjb_view_constants& vc = ctx->g.fc.view_constants;
    Img depth_tex(a->depth_tex), geometric_normal_tex(a->geometric_normal_tex), prev_depth_tex(a->prev_depth_tex), velocity_tex(a->velocity_tex), output_tex(a->output_tex);
    const float4 output_tex_size = f4(a->output_tex_size);
    const int W = output_tex.w(), H = output_tex.h();
    pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) {
        const int2 px(x, y);
        float2 uv = get_uv(px, output_tex_size);
        if (depth_tex.load(px).x == 0.0f) {
            float2 cs = uv_to_cs(uv);
            float4 pos_cs(cs.x, cs.y, 0.0f, 1.0f);
            float4 pos_vs = mul(vc.clip_to_view, pos_cs);
            float4 prev_vs = pos_vs;
            float4 prev_cs = mul(vc.view_to_clip, prev_vs);
            float4 prev_pcs = mul(vc.clip_to_prev_clip, prev_cs);
            float2 prev_uv = cs_to_uv(prev_pcs.xy());
            float2 uv_diff = prev_uv - uv;
            output_tex.store(px, float4(uv_diff.x, uv_diff.y, 0, 0));
            continue;
        }
        float depth = 0.0f;
        { float s_depth = depth_tex.load(px).x; if (s_depth != 0.0f) depth = max(depth, s_depth); }
        float3 normal_vs = geometric_normal_tex.load(px).xyz() * 2.0f - 1.0f;
        float3 normal_pvs = mul(vc.prev_clip_to_prev_view, mul(vc.clip_to_prev_clip, mul(vc.view_to_clip, float4(normal_vs, 0)))).xyz();
        float2 cs = uv_to_cs(uv);
        float4 pos_cs(cs.x, cs.y, depth, 1.0f);
        float4 pos_vs = mul(vc.clip_to_view, pos_cs);
        float dist_to_point = -(pos_vs.z / pos_vs.w);
        float4 prev_vs = pos_vs / pos_vs.w;
        float3 v = velocity_tex.load(px).xyz();
        prev_vs.x += v.x; prev_vs.y += v.y; prev_vs.z += v.z;
        float4 prev_cs = mul(vc.view_to_clip, prev_vs);
        float4 prev_pcs = mul(vc.clip_to_prev_clip, prev_cs);
        float2 prev_uv = cs_to_uv(prev_pcs.xy() / prev_pcs.w);
        float2 uv_diff = prev_uv - uv;
        uv_diff = floor(uv_diff * 32767.0f + 0.5f) / 32767.0f;
        prev_uv = uv + uv_diff;
        float4 prev_pvs = mul(vc.prev_clip_to_prev_view, prev_pcs);
        prev_pvs = prev_pvs / prev_pvs.w;
        float plane_dist_prev_dz = min(-0.2f, normal_vs.z);
        const Bilinear bilinear_at_prev = get_bilinear_filter(prev_uv, float2(output_tex_size.x, output_tex_size.y));
        // GatherRed(...).wzxy at the texel quad whose top-left is `origin`: (x,y),(x+1,y),(x,y+1),(x+1,y+1), clamp addressing
        const int ox = kjb_cvt_i32(bilinear_at_prev.origin.x), oy = kjb_cvt_i32(bilinear_at_prev.origin.y);
        auto cl = [](int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); };
        float4 prev_depth(prev_depth_tex.load(cl(ox, W), cl(oy, H)).x, prev_depth_tex.load(cl(ox + 1, W), cl(oy, H)).x,
                          prev_depth_tex.load(cl(ox, W), cl(oy + 1, H)).x, prev_depth_tex.load(cl(ox + 1, W), cl(oy + 1, H)).x);
        const float k43 = -vc.prev_clip_to_prev_view.m[2 * 4 + 3];
        float4 prev_view_z(rcp(prev_depth.x * k43), rcp(prev_depth.y * k43), rcp(prev_depth.z * k43), rcp(prev_depth.w * k43));
        float4 quad_dists = abs(plane_dist_prev_dz * (prev_view_z - prev_pvs.z));
        const float acceptance_threshold = 0.001f * (1080.0f / output_tex_size.y);
        const float3 pos_vs_norm = normalize(pos_vs.xyz() / pos_vs.w);
        const float ndotv = dot(normal_vs, pos_vs_norm);
        const float prev_ndotv = dot(normal_pvs, normalize(prev_pvs.xyz()));
        const float thr = acceptance_threshold * dist_to_point / -ndotv;
        float4 quad_validity(step(quad_dists.x, thr), step(quad_dists.y, thr), step(quad_dists.z, thr), step(quad_dists.w, thr));
        auto inb = [&](int xx, int yy) { return (xx >= 0 && yy >= 0 && xx < int(kjb_cvt_u32(output_tex_size.x)) && yy < int(kjb_cvt_u32(output_tex_size.y))) ? 1.0f : 0.0f; };
        quad_validity.x *= inb(ox, oy); quad_validity.y *= inb(ox + 1, oy); quad_validity.z *= inb(ox, oy + 1); quad_validity.w *= inb(ox + 1, oy + 1);
        float validity = dot(quad_validity, float4(1, 2, 4, 8)) / 15.0f;
        float accuracy = 1;
        accuracy *= smoothstep(0.8f, 0.95f, prev_ndotv / ndotv);
        float2 sat = saturate(prev_uv);
        if (sat.x != prev_uv.x || sat.y != prev_uv.y) accuracy = -1;
        output_tex.store(px, float4(uv_diff.x, uv_diff.y, validity, accuracy));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ sky/comp_cube.hlsl, convolve_cube.hlsl, inc/cube_map.hlsl
static const float CUBE_ROT[6][9] = {
    {0, 0, -1, 0, -1, 0, -1, 0, 0}, {0, 0, 1, 0, -1, 0, 1, 0, 0}, {1, 0, 0, 0, 0, -1, 0, 1, 0},
    {1, 0, 0, 0, 0, 1, 0, -1, 0}, {1, 0, 0, 0, -1, 0, 0, 0, -1}, {-1, 0, 0, 0, -1, 0, 0, 0, 1}};
static float3 cube_dir(int face, float2 uv) {
    float3 v(uv.x * 2 - 1, uv.y * 2 - 1, -1.0f);
    const float* m = CUBE_ROT[face];
    return normalize(float3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z));
}
int kjb_pass_sky_cube(kjb_context* ctx, const kjb_sky_cube_args* a) {
    Img out(a->output_tex); const int W = out.w();
    for (int face = 0; face < 6; ++face) parallel_rows(W, [&](int y) { for (int x = 0; x < W; ++x) {
        float2 uv = (float2(float(x), float(y)) + 0.5f) / 64.0f;
        float3 dir = cube_dir(face, uv);
        float3 o = atmosphere_default(ctx->g, dir, sun_direction(ctx->g));
        out.store(x, y, float4(o, 1), face);
    } }, ctx->num_threads);
    return 0;
}
int kjb_pass_convolve_sky(kjb_context* ctx, const kjb_convolve_sky_args* a) {
    Img in(a->input_tex), out(a->output_tex); const int W = out.w();
    for (int face = 0; face < 6; ++face) parallel_rows(W, [&](int y) { for (int x = 0; x < W; ++x) {
        float2 uv = (float2(float(x), float(y)) + 0.5f) / float(a->face_width);
        float3 output_dir = cube_dir(face, uv);
        const float3x3 basis = build_orthonormal_basis(output_dir);
        const uint sample_count = 512;
        float4 result(0.0f);
        for (uint i = 0; i < sample_count; ++i) {
            float2 urand = hammersley(i, sample_count);
            float3 input_dir = mul(basis, uniform_sample_cone(urand, 0.99f));
            result += in.sample_cube(input_dir);
        }
        out.store(x, y, result / float(sample_count), face);
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ lut/brdf_fg.hlsl:6-45
int kjb_pass_brdf_fg_lut(kjb_context* ctx, const kjb_brdf_fg_lut_args* a) {
    Img out(a->output_tex);
    parallel_rows(64, [&](int py) { for (int pxx = 0; pxx < 64; ++pxx) {
        float ndotv = (float(pxx) / (64.0f - 1.0f)) * (1.0f - 1e-3f) + 1e-3f;
        float roughness = max(1e-5f, float(py) / (64.0f - 1.0f));
        float3 wo(sqrt(1.0f - ndotv * ndotv), 0, ndotv);
        float aa = 0, bb = 0, valid = 0;
        SpecularBrdf brdf_a; brdf_a.roughness = roughness; brdf_a.albedo = float3(1.0f);
        SpecularBrdf brdf_b = brdf_a; brdf_b.albedo = float3(0.0f);
        const uint num_samples = 1024;
        for (uint i = 0; i < num_samples; ++i) {
            float2 urand = hammersley(i, num_samples);
            BrdfSample v_a = brdf_a.sample(wo, urand);
            if (v_a.is_valid()) {
                BrdfValue v_b = brdf_b.evaluate(wo, v_a.wi);
                aa += (v_a.value_over_pdf.x - v_b.value_over_pdf.x);
                bb += v_b.value_over_pdf.x;
                valid += 1;
            }
        }
        out.store(pxx, py, float4(float3(aa, bb, valid) / float(num_samples), 1.0f));
    } }, ctx->num_threads);
    return 0;
}

// ------------------------------------------------------------------ extract_half_res_*.hlsl
int kjb_pass_extract_half_res_depth(kjb_context* ctx, const kjb_extract_half_res_args* a) {
    Img in(a->input_tex), out(a->output_tex); const int2 o = halfres_subsample_offset(ctx->g.fc.frame_index);
    pass_rows(ctx, out.h(), [&](int y) { for (int x = 0; x < out.w(); ++x) out.store(x, y, float4(in.load(x * 2 + o.x, y * 2 + o.y).x)); }, 1);
    return 0;
}
int kjb_pass_extract_half_res_ssao(kjb_context* ctx, const kjb_extract_half_res_args* a) {
    return kjb_pass_extract_half_res_depth(ctx, a);
}
int kjb_pass_extract_half_res_fused(kjb_context* ctx, const kjb_extract_half_res_fused_args* a) {   // the product's one-launch fusion = the reference's three passes
    kjb_extract_half_res_args d{a->depth_tex, a->half_depth_out}, n{a->gbuffer_tex, a->half_view_normal_out}, s{a->ssao_tex, a->half_ssao_out};
    int rc = kjb_pass_extract_half_res_depth(ctx, &d) | kjb_pass_extract_half_res_view_normal(ctx, &n);
    if (a->ssao_tex.data && a->half_ssao_out.data) rc |= kjb_pass_extract_half_res_ssao(ctx, &s);
    return rc;
}
int kjb_pass_extract_half_res_view_normal(kjb_context* ctx, const kjb_extract_half_res_args* a) {   // extract_half_res_gbuffer_view_normal_rgba8.hlsl:15-53 ("tired" branch)
    Img in(a->input_tex), out(a->output_tex); const int2 o = halfres_subsample_offset(ctx->g.fc.frame_index);
    const kjb_view_constants& vc = ctx->g.fc.view_constants;
    pass_rows(ctx, out.h(), [&](int y) { for (int x = 0; x < out.w(); ++x) {
        uint4 gbt = in.load_u(x * 2 + o.x, y * 2 + o.y);
        float3 normal_ws = unpack_normal_11_10_11_no_normalize(asfloat(gbt.y));
        float3 normal_vs = normalize(mul(vc.world_to_view, float4(normal_ws, 0)).xyz());
        out.store(x, y, float4(normal_vs, 1));
    } }, 1);
    return 0;
}

}  // extern "C"

}  // namespace kjo
