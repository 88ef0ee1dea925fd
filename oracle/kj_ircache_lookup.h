// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).  PARITY UNPINNED; the irradiance cache is racy BY DESIGN upstream
// (docs/gi-overview.md:296, ircache.rs:68-76 write_no_sync): the oracle runs cache-touching passes single-threaded in pixel
// order, which is ONE of the orders the reference can exhibit — CUDA-vs-oracle parity for the cache is therefore statistical.
// Restates ircache/{ircache_constants,ircache_grid,ircache_sampler_common.inc,lookup}.hlsl.
#pragma once
#include "kj_ctx.h"

namespace kjo {

static const float IRCACHE_GRID_CELL_DIAMETER = 0.16f * 0.125f;
static const uint IRCACHE_CASCADE_SIZE = 32, IRCACHE_CASCADE_COUNT = 12;
static const uint IRCACHE_META_TRACING_ALLOC_COUNT = 0, IRCACHE_META_ENTRY_COUNT = 2, IRCACHE_META_ALLOC_COUNT = 3;   // u32 slots
static const uint IRCACHE_ENTRY_META_OCCUPIED = 1u, IRCACHE_ENTRY_META_JUST_ALLOCATED = 2u;
static const uint IRCACHE_ENTRY_LIFE_RECYCLE = 0x8000000u, IRCACHE_ENTRY_LIFE_RECYCLED = 0x8000001u;
static const uint IRCACHE_ENTRY_LIFE_PER_RANK = 4, IRCACHE_ENTRY_RANK_COUNT = 3;
static const uint IRCACHE_OCTA_DIMS = 4, IRCACHE_OCTA_DIMS2 = 16, IRCACHE_IRRADIANCE_STRIDE = 3, IRCACHE_AUX_STRIDE = 64;
static const uint IRCACHE_SAMPLES_PER_FRAME = 4, IRCACHE_VALIDATION_SAMPLES_PER_FRAME = 4, IRCACHE_RESTIR_M_CLAMP = 30;
inline bool is_ircache_entry_life_valid(uint life) { return life < IRCACHE_ENTRY_LIFE_PER_RANK * IRCACHE_ENTRY_RANK_COUNT; }
inline uint ircache_entry_life_to_rank(uint life) { return life / IRCACHE_ENTRY_LIFE_PER_RANK; }
inline uint ircache_entry_life_for_rank(uint rank) { return rank * IRCACHE_ENTRY_LIFE_PER_RANK; }

struct IrcacheBufs {   // DEFINE_IRCACHE_BINDINGS (ircache/bindings.hlsl) + aux for IRCACHE_LOOKUP_PRECISE
    uint32_t *meta, *pool, *reposition_count, *grid_meta, *entry_cell, *life;
    float4 *reposition_proposal; const float4 *spatial; const float4 *irradiance; float4* aux;
    bool bound() const { return meta != nullptr; }
    static IrcacheBufs from(const kjb_ircache_bindings& b) {
        IrcacheBufs r{}; r.meta = (uint32_t*)b.meta_buf.data; r.pool = (uint32_t*)b.pool_buf.data; r.reposition_proposal = (float4*)b.reposition_proposal_buf.data;
        r.reposition_count = (uint32_t*)b.reposition_proposal_count_buf.data; r.grid_meta = (uint32_t*)b.grid_meta_buf.data; r.entry_cell = (uint32_t*)b.entry_cell_buf.data;
        r.spatial = (const float4*)b.spatial_buf.data; r.irradiance = (const float4*)b.irradiance_buf.data; r.life = (uint32_t*)b.life_buf.data; r.aux = nullptr;   /* not in the binding block (ircache/bindings.hlsl): only the cache's own passes are IRCACHE_LOOKUP_PRECISE */
        return r;
    }
};

struct IrcacheCoord { uint cx, cy, cz, cascade;
    uint cell_idx() const { return cx + cy * IRCACHE_CASCADE_SIZE + cz * IRCACHE_CASCADE_SIZE * IRCACHE_CASCADE_SIZE + cascade * IRCACHE_CASCADE_SIZE * IRCACHE_CASCADE_SIZE * IRCACHE_CASCADE_SIZE; } };
inline uint ws_local_pos_to_cascade_idx(float3 local_pos, uint reserved_cells) {   // ircache_grid.hlsl:34-39
    const float3 fcoord = local_pos / IRCACHE_GRID_CELL_DIAMETER;
    const float max_coord = max(abs(fcoord.x), max(abs(fcoord.y), abs(fcoord.z)));
    const float cascade_float = log2(max_coord / float(IRCACHE_CASCADE_SIZE / 2 - reserved_cells));
    return kjb_cvt_u32(clamp(kjb_ceil(max(0.0f, cascade_float)), 0.0f, float(IRCACHE_CASCADE_COUNT - 1)));
}
inline IrcacheCoord ws_pos_to_ircache_coord(const kjb_frame_constants& fc, float3 pos, float3 normal, float3 jitter) {   // :41-75
    const float3 center(fc.ircache_grid_center[0], fc.ircache_grid_center[1], fc.ircache_grid_center[2]);
    const uint reserved_cells = 1;
    { const uint cascade = ws_local_pos_to_cascade_idx(pos - center, reserved_cells);
      const float cell_diameter = IRCACHE_GRID_CELL_DIAMETER * float(1u << cascade);
      pos = pos + cell_diameter * jitter; }
    const uint cascade = ws_local_pos_to_cascade_idx(pos - center, reserved_cells);
    const float cell_diameter = IRCACHE_GRID_CELL_DIAMETER * float(1u << cascade);
    const int* co = fc.ircache_cascades[cascade].origin;
    const float3 cell_offset = normal * cell_diameter * 0.5f;
    const float3 q = (pos + cell_offset) / cell_diameter;
    const int ix = kjb_cvt_i32(floor(q.x)) - co[0], iy = kjb_cvt_i32(floor(q.y)) - co[1], iz = kjb_cvt_i32(floor(q.z)) - co[2];
    auto cl = [](int v) { return uint(v < 0 ? 0 : (v > 31 ? 31 : v)); };
    IrcacheCoord r; r.cascade = cascade; r.cx = cl(ix); r.cy = cl(iy); r.cz = cl(iz);
    return r;
}
inline float ircache_grid_cell_diameter_in_cascade(uint cascade) { return IRCACHE_GRID_CELL_DIAMETER * float(1u << cascade); }

// ircache_sampler_common.inc.hlsl
struct SampleParams { uint value;
    static SampleParams from_spf_entry_sample_frame(uint spf, uint entry_idx, uint sample_idx, uint frame_idx) {
        const uint PERIOD = IRCACHE_OCTA_DIMS2 / spf;
        uint xy = sample_idx * PERIOD + (frame_idx % PERIOD);
        xy ^= (xy & 4u) >> 2u;
        SampleParams r; r.value = xy + ((frame_idx << 16u) ^ (entry_idx)) * IRCACHE_OCTA_DIMS2; return r;
    }
    uint octa_idx() const { return value % IRCACHE_OCTA_DIMS2; }
    uint rng() const { return hash1(value >> 4u); }
    float2 octa_uv() const {
        const uint oi = octa_idx();
        const float2 urand = r2_sequence(rng() % 1024u);
        return (float2(float(oi % IRCACHE_OCTA_DIMS), float(oi / IRCACHE_OCTA_DIMS)) + urand) / 4.0f;
    }
    float3 direction() const { return octa_decode(octa_uv()); }
};

struct Vertex { float3 position, normal; };
inline float3 unpack_unit_direction_11_10_11_o(uint pck) {
    return float3(float(pck & 2047u) * (2.0f / 2047.0f) - 1.0f, float((pck >> 11u) & 1023u) * (2.0f / 1023.0f) - 1.0f, float(pck >> 21u) * (2.0f / 2047.0f) - 1.0f);
}
inline Vertex unpack_vertex(float4 p) { Vertex v; v.position = p.xyz(); v.normal = unpack_unit_direction_11_10_11_o(asuint(p.w)); return v; }   // mesh.hlsl:34-39
inline float4 pack_vertex(const Vertex& v) { return float4(v.position, pack_normal_11_10_11(v.normal)); }                                  // mesh.hlsl:41-46

inline float eval_sh_geometrics(float4 sh, float3 normal) {   // lookup.hlsl:197-212
    float R0 = sh.x;
    float3 R1 = 0.5f * float3(sh.y, sh.z, sh.w);
    float lenR1 = length(R1);
    float q = 0.5f * (1.0f + dot(R1 / lenR1, normal));
    float p = 1.0f + 2.0f * lenR1 / R0;
    float a = (1.0f - lenR1 / R0) / (1.0f + lenR1 / R0);
    return R0 * (a + (1.0f - a) * (p + 1.0f) * pow(q, p));
}

// IrcacheLookupParams::lookup (lookup.hlsl:76-311).  `precise` = IRCACHE_LOOKUP_PRECISE (used by the cache's own tracing passes).
inline float3 ircache_lookup(const Globals& g, const IrcacheBufs& b, float3 query_from_ws, float3 pt_ws, float3 normal_ws, uint query_rank, uint& rng, bool precise, bool stochastic_interpolation = false) {
    if (!b.bound()) return float3(0.0f);
    const kjb_frame_constants& fc = g.fc;
    bool allocated_by_us = false, just_allocated = false;
    // lookup.hlsl:80-86: `select(stochastic_interpolation, float3(rand, rand, rand) - 0.5, 0)` — select() is a function, its arguments are
    // evaluated eagerly, so the three hash1_mut(rng) draws happen whether or not interpolation is on
    float3 jr; jr.x = uint_to_u01_float(hash1_mut(rng)); jr.y = uint_to_u01_float(hash1_mut(rng)); jr.z = uint_to_u01_float(hash1_mut(rng));
    const float3 jitter = stochastic_interpolation ? jr - 0.5f : float3(0.0f);
    {
        const IrcacheCoord rcoord = ws_pos_to_ircache_coord(fc, pt_ws, normal_ws, jitter);
        const int* so = fc.ircache_cascades[rcoord.cascade].voxels_scrolled_this_frame;
        auto jsi = [&](int c, int s) { return s > 0 ? (c + s >= int(IRCACHE_CASCADE_SIZE)) : (c < -s); };
        const bool was_just_scrolled_in = jsi(int(rcoord.cx), so[0]) || jsi(int(rcoord.cy), so[1]) || jsi(int(rcoord.cz), so[2]);
        const bool skip_allocation = query_rank >= IRCACHE_ENTRY_RANK_COUNT || (was_just_scrolled_in && query_rank > 0);
        const uint cell_idx = rcoord.cell_idx();
        const uint entry_flags = b.grid_meta[cell_idx * 2 + 1];
        just_allocated = (entry_flags & IRCACHE_ENTRY_META_JUST_ALLOCATED) != 0;
        if (!skip_allocation && (entry_flags & IRCACHE_ENTRY_META_OCCUPIED) == 0) {
            const uint prev = __atomic_fetch_or(&b.grid_meta[cell_idx * 2 + 1], IRCACHE_ENTRY_META_OCCUPIED | IRCACHE_ENTRY_META_JUST_ALLOCATED, __ATOMIC_RELAXED);
            if ((prev & IRCACHE_ENTRY_META_OCCUPIED) == 0) {
                just_allocated = true; allocated_by_us = true;
                const uint alloc_idx = __atomic_fetch_add(&b.meta[IRCACHE_META_ALLOC_COUNT], 1u, __ATOMIC_RELAXED);
                if (alloc_idx >= 1024 * 64) {
                    __atomic_fetch_add(&b.meta[IRCACHE_META_ALLOC_COUNT], uint(-1), __ATOMIC_RELAXED);
                    __atomic_fetch_and(&b.grid_meta[cell_idx * 2 + 1], ~(IRCACHE_ENTRY_META_OCCUPIED | IRCACHE_ENTRY_META_JUST_ALLOCATED), __ATOMIC_RELAXED);
                } else {
                    const uint entry_idx = b.pool[alloc_idx];
                    uint cur = b.meta[IRCACHE_META_ENTRY_COUNT];
                    while (cur < entry_idx + 1 && !__atomic_compare_exchange_n(&b.meta[IRCACHE_META_ENTRY_COUNT], &cur, entry_idx + 1, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                    b.life[entry_idx] = ircache_entry_life_for_rank(query_rank);
                    b.entry_cell[entry_idx] = cell_idx;
                    b.grid_meta[cell_idx * 2 + 0] = entry_idx;
                }
            }
        }
    }
    // ircache_lookup (:19-36)
    uint lookup_count = 0, lookup_entry = 0;
    const IrcacheCoord rc = ws_pos_to_ircache_coord(fc, pt_ws, normal_ws, jitter);
    {
        const uint cell_idx = rc.cell_idx();
        if (b.grid_meta[cell_idx * 2 + 1] & IRCACHE_ENTRY_META_OCCUPIED) { lookup_entry = b.grid_meta[cell_idx * 2 + 0]; lookup_count = 1; }
    }
    const float cell_diameter = ircache_grid_cell_diameter_in_cascade(rc.cascade);
    float3 offset_towards_query = query_from_ws - pt_ws;
    const float MAX_OFFSET = cell_diameter, MAX_OFFSET_AS_FRAC = 0.5f;
    offset_towards_query = offset_towards_query * (MAX_OFFSET / max(MAX_OFFSET / MAX_OFFSET_AS_FRAC, length(offset_towards_query)));
    Vertex new_entry; new_entry.position = pt_ws + offset_towards_query; new_entry.normal = normal_ws;
    if (allocated_by_us && lookup_count) b.reposition_proposal[lookup_entry] = pack_vertex(new_entry);
    if (just_allocated) return float3(0.0f);

    float3 irradiance_sum(0.0f);
    if (lookup_count) {
        const uint entry_idx = lookup_entry;
        float3 irradiance(0.0f);
        if (precise) {
            float weight_sum = 0;
            for (uint octa_idx = 0; octa_idx < IRCACHE_OCTA_DIMS2; ++octa_idx) {
                const float4 ra = b.aux[entry_idx * IRCACHE_AUX_STRIDE + octa_idx];
                const Reservoir1spp r = Reservoir1spp::from_raw(uint2(asuint(ra.x), asuint(ra.y)));
                SampleParams sp; sp.value = r.payload;
                const float3 dir = sp.direction();
                const float wt = dot(dir, normal_ws);
                if (wt > 0.0f) {
                    const float4 contrib = b.aux[entry_idx * IRCACHE_AUX_STRIDE + IRCACHE_OCTA_DIMS2 + octa_idx];
                    irradiance += contrib.xyz() * wt * contrib.w;
                    weight_sum += wt;
                }
            }
            irradiance = irradiance / max(1.0f, weight_sum);
        } else {
            irradiance.x += eval_sh_geometrics(b.irradiance[entry_idx * 3 + 0], normal_ws);
            irradiance.y += eval_sh_geometrics(b.irradiance[entry_idx * 3 + 1], normal_ws);
            irradiance.z += eval_sh_geometrics(b.irradiance[entry_idx * 3 + 2], normal_ws);
        }
        irradiance = max(float3(0.0f), irradiance);
        irradiance_sum += irradiance * 1.0f;
        const uint prev_life = b.life[entry_idx];
        if (prev_life < IRCACHE_ENTRY_LIFE_RECYCLE) {
            const uint new_life = ircache_entry_life_for_rank(query_rank);
            if (new_life < prev_life) { uint cur = b.life[entry_idx]; while (cur > new_life && !__atomic_compare_exchange_n(&b.life[entry_idx], &cur, new_life, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} }
            const uint prev_rank = ircache_entry_life_to_rank(prev_life);
            if (query_rank <= prev_rank) {
                const uint prev_vote_count = __atomic_fetch_add(&b.reposition_count[entry_idx], 1u, __ATOMIC_RELAXED);
                const float dart = uint_to_u01_float(hash1_mut(rng));
                const float prob = 1.0f / (float(prev_vote_count) + 1.0f);
                if (dart <= prob) b.reposition_proposal[entry_idx] = pack_vertex(new_entry);
            }
        }
    }
    return irradiance_sum;
}

}  // namespace kjo
