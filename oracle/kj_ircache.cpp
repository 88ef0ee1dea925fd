// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h / kj_ircache_lookup.h headers).  PARITY UNPINNED and, for this subsystem,
// STATISTICAL: the reference's irradiance cache is racy by design.  Every pass below runs single-threaded in index order.
// One function per render-graph pass of crates/lib/kajiya/src/renderers/ircache.rs; shaders under assets/shaders/ircache/.
#include "kj_ircache_lookup.h"

namespace kjo {
namespace {

struct IrcacheTraceResult { float3 incident_radiance, direction, hit_pos; };

// Slot 0 of the indirection table is never written (inclusive scan, ircache_compact_entries.hlsl:17) and keeps naming entry 0, which
// is also slot 1 when alive.  On the reference's GPU the two copies are adjacent lanes of one wave running in lockstep — same reads,
// same writes — so the entry is effectively processed once; a serial loop would process it twice.  Skip the stale duplicate.
inline bool slot_is_stale_duplicate(const uint32_t* ind, uint alloc_count, uint slot) { return slot == 0 && alloc_count > 1 && ind[1] == ind[0]; }

// ircache/ircache_trace_common.inc.hlsl:37-227 (MAX_PATH_LENGTH 1, USE_WORLD_RADIANCE_CACHE 0, IRCACHE_LOOKUP_PRECISE)
IrcacheTraceResult ircache_trace(const kjb_context& ctx, const IrcacheBufs& b, const Img& sky_cube_tex, const Vertex& entry, SampleParams sample_params, uint life) {
    const Globals& g = ctx.g;
    uint rng = sample_params.rng();
    Ray outgoing_ray; outgoing_ray.origin = entry.position; outgoing_ray.dir = sample_params.direction(); outgoing_ray.tmin = 0.0f; outgoing_ray.tmax = FLT_MAX_F;
    IrcacheTraceResult result; result.direction = outgoing_ray.dir; result.hit_pos = float3(0.0f);
    float3 throughput(1.0f); float roughness_bias = 0.5f; float3 irradiance_sum(0.0f);
    for (uint path_length = 0; path_length < 1; ++path_length) {
        const GbufferPathVertex primary_hit = gbuffer_raytrace(ctx.scene, g, outgoing_ray, RayCone::from_spread_angle(0.1f), path_length + 1, false);
        if (primary_hit.is_hit) {
            if (0 == path_length) result.hit_pos = primary_hit.position;
            const float3 to_light_norm = sun_direction(g);
            const bool is_shadowed = rt_is_shadowed(ctx.scene, primary_hit.position, to_light_norm, 1e-4f, FLT_MAX_F);
            GbufferData gbuffer = gbuffer_unpack(primary_hit.gbuffer_packed);
            const float3x3 tangent_to_world = build_orthonormal_basis(gbuffer.normal);
            const float3 wi = mul(to_light_norm, tangent_to_world);
            float3 wo = mul(-outgoing_ray.dir, tangent_to_world);
            if (wo.z < 0.0f) { wo.z *= -0.25f; wo = normalize(wo); }
            LayeredBrdf brdf = LayeredBrdf::from_gbuffer_ndotv(g, gbuffer, wo.z);
            brdf.specular_brdf.roughness = lerp(brdf.specular_brdf.roughness, 1.0f, roughness_bias);
            const float3 brdf_value = brdf.evaluate_directional_light(wo, wi);
            const float3 light_radiance = is_shadowed ? float3(0.0f) : sun_color_in_direction(g, sun_direction(g));
            irradiance_sum += throughput * brdf_value * light_radiance * max(0.0f, wi.z);
            irradiance_sum += gbuffer.emissive * throughput;
            if (g.fc.triangle_light_count > 0) {
                const float light_selection_pmf = 1.0f / float(g.fc.triangle_light_count);
                const uint light_idx = hash1_mut(rng) % g.fc.triangle_light_count;
                float2 urand; urand.x = uint_to_u01_float(hash1_mut(rng)); urand.y = uint_to_u01_float(hash1_mut(rng));
                const kjb_triangle_light& tl = g.lights[light_idx];
                float3 v0(tl.verts[0][0], tl.verts[0][1], tl.verts[0][2]), v1(tl.verts[1][0], tl.verts[1][1], tl.verts[1][2]), v2(tl.verts[2][0], tl.verts[2][1], tl.verts[2][2]);
                LightSampleResultArea ls = sample_triangle_light(v0, v1 - v0, v2 - v0, urand);
                const float3 to_light_ws = ls.pos - primary_hit.position;
                const float dist_to_light2 = dot(to_light_ws, to_light_ws);
                const float3 to_light_norm_ws = to_light_ws * rsqrt(dist_to_light2);
                const float to_psa_metric = max(0.0f, dot(to_light_norm_ws, gbuffer.normal)) * max(0.0f, dot(to_light_norm_ws, -ls.normal)) / dist_to_light2;
                if (to_psa_metric > 0.0f) {
                    float3 wi2 = mul(to_light_norm_ws, tangent_to_world);
                    const bool sh = rt_is_shadowed(ctx.scene, primary_hit.position, to_light_norm_ws, 1e-3f, sqrt(dist_to_light2) - 2e-3f);
                    float3 radiance(tl.radiance[0], tl.radiance[1], tl.radiance[2]);
                    irradiance_sum += sh ? float3(0.0f) : throughput * radiance * brdf.evaluate(wo, wi2) / ls.pdf * to_psa_metric / light_selection_pmf;
                }
            }
            // SAMPLE_IRCACHE_AT_LAST_VERTEX
            irradiance_sum += ircache_lookup(g, b, entry.position, primary_hit.position, gbuffer.normal, 1 + ircache_entry_life_to_rank(life), rng, true) * throughput * gbuffer.albedo;
            break;   // MAX_PATH_LENGTH == 1: the BRDF-sampled continuation ray is never traced
        } else {
            if (0 == path_length) result.hit_pos = outgoing_ray.origin + outgoing_ray.dir * 1000.0f;
            irradiance_sum += throughput * sky_cube_tex.sample_cube(outgoing_ray.dir).xyz();
            break;
        }
    }
    result.incident_radiance = irradiance_sum;
    return result;
}

IrcacheBufs bufs_of(const kjb_ircache_trace_args* a) {
    IrcacheBufs b{}; b.meta = (uint32_t*)a->meta_buf.data; b.pool = (uint32_t*)a->pool_buf.data; b.reposition_proposal = (float4*)a->reposition_proposal_buf.data;
    b.reposition_count = (uint32_t*)a->reposition_proposal_count_buf.data; b.grid_meta = (uint32_t*)a->grid_meta_buf.data; b.entry_cell = (uint32_t*)a->entry_cell_buf.data;
    b.spatial = (const float4*)a->spatial_buf.data; b.irradiance = nullptr; b.life = (uint32_t*)a->life_buf.data; b.aux = (float4*)a->aux_buf.data;
    return b;
}
}  // namespace

extern "C" {

int kjb_pass_ircache_clear_pool(kjb_context*, const kjb_ircache_clear_pool_args* a) {   // clear_ircache_pool.hlsl
    uint32_t *pool = (uint32_t*)a->pool_buf.data, *life = (uint32_t*)a->life_buf.data;
    for (uint idx = 0; idx < KJB_IRCACHE_MAX_ENTRIES; ++idx) { pool[idx] = idx; life[idx] = IRCACHE_ENTRY_LIFE_RECYCLED; }
    return 0;
}

int kjb_pass_ircache_scroll_cascades(kjb_context* ctx, const kjb_ircache_scroll_cascades_args* a) {   // scroll_cascades.hlsl:36-69
    const uint32_t* gm = (const uint32_t*)a->grid_meta_buf.data; uint32_t* gm2 = (uint32_t*)a->grid_meta_buf2.data;
    uint32_t *entry_cell = (uint32_t*)a->entry_cell_buf.data, *life = (uint32_t*)a->life_buf.data, *pool = (uint32_t*)a->pool_buf.data, *meta = (uint32_t*)a->meta_buf.data;
    float4* irradiance = (float4*)a->irradiance_buf.data;
    const kjb_frame_constants& fc = ctx->g.fc;
    for (uint z = 0; z < 32 * 12; ++z) for (uint y = 0; y < 32; ++y) for (uint x = 0; x < 32; ++x) {
        const uint dz = z % 32, cascade = z / 32;
        const uint dst_cell_idx = x + y * 32 + dz * 1024 + cascade * 32768;
        const int* sb = fc.ircache_cascades[cascade].voxels_scrolled_this_frame;
        const uint ox = x - uint(sb[0]), oy = y - uint(sb[1]), oz = dz - uint(sb[2]);
        if (!(ox < 32 && oy < 32 && oz < 32)) {   // deallocate_cell(dst_cell_idx)
            const uint m0 = gm[dst_cell_idx * 2], m1 = gm[dst_cell_idx * 2 + 1];
            if (m1 & IRCACHE_ENTRY_META_OCCUPIED) {
                const uint entry_idx = m0;
                life[entry_idx] = IRCACHE_ENTRY_LIFE_RECYCLED;
                for (uint i = 0; i < 3; ++i) irradiance[entry_idx * 3 + i] = float4(0.0f);
                const uint entry_alloc_count = meta[IRCACHE_META_ALLOC_COUNT]; meta[IRCACHE_META_ALLOC_COUNT] -= 1;
                pool[entry_alloc_count - 1] = entry_idx;
            }
        }
        const uint sx = x + uint(sb[0]), sy = y + uint(sb[1]), sz = dz + uint(sb[2]);
        if (sx < 32 && sy < 32 && sz < 32) {
            const uint src_cell_idx = sx + sy * 32 + sz * 1024 + cascade * 32768;
            const uint m0 = gm[src_cell_idx * 2], m1 = gm[src_cell_idx * 2 + 1];
            gm2[dst_cell_idx * 2] = m0; gm2[dst_cell_idx * 2 + 1] = m1;
            if (m1 & IRCACHE_ENTRY_META_OCCUPIED) entry_cell[m0] = dst_cell_idx;
        } else { gm2[dst_cell_idx * 2] = 0; gm2[dst_cell_idx * 2 + 1] = 0; }
    }
    return 0;
}

int kjb_pass_ircache_prepare_age_dispatch_args(kjb_context*, const kjb_ircache_dispatch_args_args* a) {   // prepare_age_dispatch_args.hlsl
    const uint32_t* meta = (const uint32_t*)a->meta_buf.data; uint32_t* args = (uint32_t*)a->dispatch_args.data;
    args[0] = (meta[IRCACHE_META_ENTRY_COUNT] + 63) / 64; args[1] = 1; args[2] = 1; args[3] = 0;
    return 0;
}
int kjb_pass_ircache_prepare_trace_dispatch_args(kjb_context*, const kjb_ircache_dispatch_args_args* a) {   // prepare_trace_dispatch_args.hlsl
    uint32_t* meta = (uint32_t*)a->meta_buf.data; uint32_t* args = (uint32_t*)a->dispatch_args.data;
    const uint alloc_count = meta[IRCACHE_META_ALLOC_COUNT];
    meta[IRCACHE_META_TRACING_ALLOC_COUNT] = alloc_count;
    args[8] = (alloc_count + 63) / 64; args[9] = 1; args[10] = 1; args[11] = 0;
    const uint mx = std::max(alloc_count * 4, std::max(alloc_count * 16, alloc_count * 4));
    args[0] = mx; args[1] = 1; args[2] = 1; args[3] = 0; args[4] = mx; args[5] = 1; args[6] = 1; args[7] = 0; args[12] = mx; args[13] = 1; args[14] = 1; args[15] = 0;
    return 0;
}

int kjb_pass_ircache_age_entries(kjb_context*, const kjb_ircache_age_args* a) {   // age_ircache_entries.hlsl:55-94 (dispatch_indirect: ceil(entry_count/64) groups of 64)
    uint32_t *meta = (uint32_t*)a->meta_buf.data, *gm = (uint32_t*)a->grid_meta_buf.data, *entry_cell = (uint32_t*)a->entry_cell_buf.data, *life = (uint32_t*)a->life_buf.data,
             *pool = (uint32_t*)a->pool_buf.data, *count = (uint32_t*)a->reposition_proposal_count_buf.data, *occ = (uint32_t*)a->entry_occupancy_buf.data;
    float4 *spatial = (float4*)a->spatial_buf.data, *proposal = (float4*)a->reposition_proposal_buf.data, *irradiance = (float4*)a->irradiance_buf.data;
    const uint total_entry_count = meta[IRCACHE_META_ENTRY_COUNT];
    const uint threads = (total_entry_count + 63) / 64 * 64;
    for (uint entry_idx = 0; entry_idx < threads && entry_idx < KJB_IRCACHE_MAX_ENTRIES; ++entry_idx) {
        if (entry_idx < total_entry_count) {
            const uint l = life[entry_idx];
            if (l != IRCACHE_ENTRY_LIFE_RECYCLED) {
                const uint new_age = l + 1;
                if (is_ircache_entry_life_valid(new_age)) {
                    life[entry_idx] = new_age;
                    gm[entry_cell[entry_idx] * 2 + 1] &= ~IRCACHE_ENTRY_META_JUST_ALLOCATED;
                } else {
                    life[entry_idx] = IRCACHE_ENTRY_LIFE_RECYCLED;
                    for (uint i = 0; i < 3; ++i) irradiance[entry_idx * 3 + i] = float4(0.0f);
                    const uint entry_alloc_count = meta[IRCACHE_META_ALLOC_COUNT]; meta[IRCACHE_META_ALLOC_COUNT] -= 1;
                    pool[entry_alloc_count - 1] = entry_idx;
                    gm[entry_cell[entry_idx] * 2 + 1] &= ~(IRCACHE_ENTRY_META_OCCUPIED | IRCACHE_ENTRY_META_JUST_ALLOCATED);
                }
            }
            spatial[entry_idx] = proposal[entry_idx];   // IRCACHE_USE_POSITION_VOTING: flush the reposition proposal
            count[entry_idx] = 0;
        } else {
            spatial[entry_idx] = float4(0.0f);
        }
        const uint l2 = life[entry_idx];
        occ[entry_idx] = (entry_idx < total_entry_count && is_ircache_entry_life_valid(l2)) ? 1u : 0u;
    }
    // threads beyond the dispatched groups never run: their occupancy stays whatever the (transient, zero-initialised here) buffer held
    return 0;
}

int kjb_pass_inclusive_prefix_scan_u32(kjb_context*, const kjb_prefix_scan_args* a) {   // prefix_scan/*.hlsl: inclusive scan
    uint32_t* d = (uint32_t*)a->inout_buf.data; uint32_t acc = 0;
    for (uint32_t i = 0; i < a->element_count; ++i) { acc += d[i]; d[i] = acc; }
    return 0;
}

int kjb_pass_ircache_compact(kjb_context*, const kjb_ircache_compact_args* a) {   // ircache_compact_entries.hlsl
    const uint32_t *meta = (const uint32_t*)a->meta_buf.data, *life = (const uint32_t*)a->life_buf.data, *occ = (const uint32_t*)a->entry_occupancy_buf.data;
    uint32_t* ind = (uint32_t*)a->entry_indirection_buf.data;
    const uint total = meta[IRCACHE_META_ENTRY_COUNT];
    const uint threads = (total + 63) / 64 * 64;
    for (uint e = 0; e < threads && e < KJB_IRCACHE_MAX_ENTRIES; ++e)
        if (e < total && is_ircache_entry_life_valid(life[e])) ind[occ[e]] = e;   // inclusive prefix => 1-based slot (upstream quirk, kept)
    return 0;
}

int kjb_pass_ircache_reset(kjb_context*, const kjb_ircache_reset_args* a) {   // reset_entry.hlsl
    const uint32_t *meta = (const uint32_t*)a->meta_buf.data, *ind = (const uint32_t*)a->entry_indirection_buf.data;
    const float4* irradiance = (const float4*)a->irradiance_buf.data; float4* aux = (float4*)a->aux_buf.data;
    const uint total_alloc_count = meta[IRCACHE_META_TRACING_ALLOC_COUNT];
    for (uint di = 0; di < total_alloc_count; ++di) {
        if (slot_is_stale_duplicate(ind, total_alloc_count, di)) continue;
        const uint entry_idx = ind[di];
        const float4 v = irradiance[entry_idx * 3];
        if (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f) for (uint i = 0; i < IRCACHE_AUX_STRIDE; ++i) aux[entry_idx * IRCACHE_AUX_STRIDE + i] = float4(0.0f);
    }
    return 0;
}

int kjb_pass_ircache_trace_access(kjb_context* ctx, const kjb_ircache_trace_access_args* a) {   // trace_accessibility.rgen.hlsl:21-66
    const uint32_t *meta = (const uint32_t*)a->meta_buf.data, *ind = (const uint32_t*)a->entry_indirection_buf.data, *life = (const uint32_t*)a->life_buf.data;
    const float4* spatial = (const float4*)a->spatial_buf.data; float4* aux = (float4*)a->aux_buf.data;
    const uint alloc_count = meta[IRCACHE_META_TRACING_ALLOC_COUNT];
    pass_items(ctx, alloc_count * IRCACHE_OCTA_DIMS2, [&](uint di) {
        if (slot_is_stale_duplicate(ind, alloc_count, di / IRCACHE_OCTA_DIMS2)) return;
        const uint entry_idx = ind[di / IRCACHE_OCTA_DIMS2], octa_idx = di % IRCACHE_OCTA_DIMS2;
        if (!is_ircache_entry_life_valid(life[entry_idx])) return;
        const Vertex entry = unpack_vertex(spatial[entry_idx]);
        const uint output_idx = entry_idx * IRCACHE_AUX_STRIDE + octa_idx;
        Reservoir1spp r = Reservoir1spp::from_raw(uint2(asuint(aux[output_idx].x), asuint(aux[output_idx].y)));
        Vertex prev_entry = unpack_vertex(aux[output_idx + IRCACHE_OCTA_DIMS2 * 2]);
        if (rt_is_shadowed(ctx->scene, entry.position, prev_entry.position - entry.position, 0.001f, 0.999f)) {
            r.M *= 0.8f;
            uint2 raw = r.as_raw(); aux[output_idx].x = asfloat(raw.x); aux[output_idx].y = asfloat(raw.y);
        }
    });
    return 0;
}

int kjb_pass_ircache_validate(kjb_context* ctx, const kjb_ircache_trace_args* a) {   // ircache_validate.rgen.hlsl:44-131
    const IrcacheBufs b = bufs_of(a); Img sky(a->sky_cube_tex);
    const uint32_t* ind = (const uint32_t*)a->entry_indirection_buf.data; float4* aux = b.aux;
    const uint alloc_count = b.meta[IRCACHE_META_TRACING_ALLOC_COUNT];
    const float ped = ctx->g.fc.pre_exposure_delta;
    pass_items(ctx, alloc_count * IRCACHE_VALIDATION_SAMPLES_PER_FRAME, [&](uint di) {
        if (slot_is_stale_duplicate(ind, alloc_count, di / 4)) return;
        const uint entry_idx = ind[di / 4], sample_idx = di % 4;
        const uint life = b.life[entry_idx];
        const SampleParams sample_params = SampleParams::from_spf_entry_sample_frame(4, entry_idx, sample_idx, ctx->g.fc.frame_index);
        const uint output_idx = entry_idx * IRCACHE_AUX_STRIDE + sample_params.octa_idx();
        Reservoir1spp r = Reservoir1spp::from_raw(uint2(asuint(aux[output_idx].x), asuint(aux[output_idx].y)));
        if (r.M > 0) {
            float4 prev_value_and_count = aux[output_idx + IRCACHE_OCTA_DIMS2] * float4(ped, ped, ped, 1);
            Vertex prev_entry = unpack_vertex(aux[output_idx + IRCACHE_OCTA_DIMS2 * 2]);
            SampleParams sp; sp.value = r.payload;
            IrcacheTraceResult prev_traced = ircache_trace(*ctx, b, sky, prev_entry, sp, life);
            const float lim = lerp(0.5f, 1.0f, smoothstep(-0.1f, 0.0f, dot(prev_traced.direction, prev_entry.normal)));
            const float3 av = prev_traced.incident_radiance * lim, bv = prev_value_and_count.xyz();
            const float3 dist3 = abs(av - bv) / (av + bv);
            const float dist = max(dist3.x, max(dist3.y, dist3.z));
            const float invalidity = smoothstep(0.1f, 0.5f, dist);
            r.M = max(0.0f, min(r.M, exp2(log2(float(IRCACHE_RESTIR_M_CLAMP)) * (1.0f - invalidity))));
            prev_value_and_count = float4(av, prev_value_and_count.w);
            uint2 raw = r.as_raw(); aux[output_idx].x = asfloat(raw.x); aux[output_idx].y = asfloat(raw.y);
            aux[output_idx + IRCACHE_OCTA_DIMS2] = prev_value_and_count;
        }
    });
    return 0;
}

int kjb_pass_ircache_trace(kjb_context* ctx, const kjb_ircache_trace_args* a) {   // trace_irradiance.rgen.hlsl:44-145
    const IrcacheBufs b = bufs_of(a); Img sky(a->sky_cube_tex);
    const uint32_t* ind = (const uint32_t*)a->entry_indirection_buf.data; float4* aux = b.aux;
    const uint alloc_count = b.meta[IRCACHE_META_TRACING_ALLOC_COUNT];
    const float ped = ctx->g.fc.pre_exposure_delta;
    pass_items(ctx, alloc_count * IRCACHE_SAMPLES_PER_FRAME, [&](uint di) {
        if (slot_is_stale_duplicate(ind, alloc_count, di / 4)) return;
        const uint entry_idx = ind[di / 4], sample_idx = di % 4;
        const uint life = b.life[entry_idx];
        const float4 packed_entry = b.spatial[entry_idx];
        const Vertex entry = unpack_vertex(packed_entry);
        uint rng = hash1(hash1(entry_idx) + ctx->g.fc.frame_index);
        const SampleParams sample_params = SampleParams::from_spf_entry_sample_frame(4, entry_idx, sample_idx, ctx->g.fc.frame_index);
        IrcacheTraceResult traced = ircache_trace(*ctx, b, sky, entry, sample_params, life);
        const float lim = lerp(0.5f, 1.0f, smoothstep(-0.1f, 0.0f, dot(traced.direction, entry.normal)));
        const float3 new_value = traced.incident_radiance * lim;
        const float new_lum = sRGB_to_luminance(new_value);
        Reservoir1sppStreamState stream_state; Reservoir1spp reservoir;
        reservoir.init_with_stream(new_lum, 1.0f, stream_state, sample_params.value);
        const uint output_idx = entry_idx * IRCACHE_AUX_STRIDE + sample_params.octa_idx();
        float4 prev_value_and_count = aux[output_idx + IRCACHE_OCTA_DIMS2] * float4(ped, ped, ped, 1);
        float3 val_sel = new_value; bool selected_new = true;
        {
            Reservoir1spp r = Reservoir1spp::from_raw(uint2(asuint(aux[output_idx].x), asuint(aux[output_idx].y)));
            if (r.M > 0) {
                r.M = min(r.M, 30.0f);
                if (reservoir.update_with_stream(r, sRGB_to_luminance(prev_value_and_count.xyz()), 1.0f, stream_state, r.payload, rng)) { val_sel = prev_value_and_count.xyz(); selected_new = false; }
            }
        }
        reservoir.finish_stream(stream_state);
        uint2 raw = reservoir.as_raw(); aux[output_idx].x = asfloat(raw.x); aux[output_idx].y = asfloat(raw.y);
        aux[output_idx + IRCACHE_OCTA_DIMS2] = float4(val_sel, reservoir.W);
        if (selected_new) aux[output_idx + IRCACHE_OCTA_DIMS2 * 2] = packed_entry;
    });
    return 0;
}

int kjb_pass_ircache_sum(kjb_context* ctx, const kjb_ircache_sum_args* a) {   // sum_up_irradiance.hlsl:34-89
    const uint32_t *meta = (const uint32_t*)a->meta_buf.data, *ind = (const uint32_t*)a->entry_indirection_buf.data;
    float4* irradiance = (float4*)a->irradiance_buf.data; const float4* aux = (const float4*)a->aux_buf.data;
    const uint total_alloc_count = meta[IRCACHE_META_TRACING_ALLOC_COUNT];
    const float ped = ctx->g.fc.pre_exposure_delta;
    for (uint di = 0; di < total_alloc_count; ++di) {
        if (slot_is_stale_duplicate(ind, total_alloc_count, di)) continue;
        const uint entry_idx = ind[di];
        float4 sh_rgb[3] = {float4(0.0f), float4(0.0f), float4(0.0f)};
        float valid_samples = 0;
        for (uint octa_idx = 0; octa_idx < IRCACHE_OCTA_DIMS2; ++octa_idx) {
            const float4 ra = aux[entry_idx * IRCACHE_AUX_STRIDE + octa_idx];
            const Reservoir1spp r = Reservoir1spp::from_raw(uint2(asuint(ra.x), asuint(ra.y)));
            SampleParams sp; sp.value = r.payload;
            const float3 dir = sp.direction();
            const float4 contrib = aux[entry_idx * IRCACHE_AUX_STRIDE + IRCACHE_OCTA_DIMS2 + octa_idx];
            const float3 radiance = contrib.xyz() * contrib.w;
            const float4 sh = float4(0.282095f, dir.x * 0.488603f, dir.y * 0.488603f, dir.z * 0.488603f) * 4.0f;
            sh_rgb[0] += sh * radiance.x; sh_rgb[1] += sh * radiance.y; sh_rgb[2] += sh * radiance.z;
            valid_samples += contrib.w > 0 ? 1.0f : 0.0f;
        }
        const float sc = 1.0f / max(1.0f, valid_samples);
        for (uint basis_i = 0; basis_i < 3; ++basis_i) {
            const float4 new_value = sh_rgb[basis_i] * sc;
            float4 prev_value = irradiance[entry_idx * 3 + basis_i] * ped;
            const bool should_reset = !(prev_value.x != 0.0f || prev_value.y != 0.0f || prev_value.z != 0.0f || prev_value.w != 0.0f);
            if (should_reset) prev_value = new_value;
            irradiance[entry_idx * 3 + basis_i] = lerp(prev_value, new_value, 0.25f);
        }
    }
    return 0;
}

}  // extern "C"
}  // namespace kjo
