// Irradiance cache: constants, grid addressing, sample parameters and the lookup-with-allocation used by every pass that
// binds the cache (ircache/{ircache_constants,ircache_grid,ircache_sampler_common.inc,lookup}.hlsl).
//
// The cache is a racy structure BY DESIGN upstream (docs/gi-overview.md "Irradiance cache"; ircache.rs:68-76 binds everything
// write_no_sync): cells are claimed with an atomic OR, entries popped with an atomic ADD, keep-alive is an atomic MIN and the
// reposition vote is last-writer-wins.  The same atomics are used here; which thread wins is scheduling dependent, exactly as
// on the reference's GPU.
#pragma once
#include "kjb_trace.cuh"

namespace kjb {

#define KJB_IRCACHE_GRID_CELL_DIAMETER (0.16f * 0.125f)
constexpr uint32_t IRCACHE_CASCADE_SIZE = 32, IRCACHE_CASCADE_COUNT = 12;
constexpr uint32_t IRCACHE_META_TRACING_ALLOC_COUNT = 0, IRCACHE_META_ENTRY_COUNT = 2, IRCACHE_META_ALLOC_COUNT = 3;   // u32 slots of meta_buf
constexpr uint32_t IRCACHE_ENTRY_META_OCCUPIED = 1u, IRCACHE_ENTRY_META_JUST_ALLOCATED = 2u;
constexpr uint32_t IRCACHE_ENTRY_LIFE_RECYCLE = 0x8000000u, IRCACHE_ENTRY_LIFE_RECYCLED = 0x8000001u;
constexpr uint32_t IRCACHE_ENTRY_LIFE_PER_RANK = 4, IRCACHE_ENTRY_RANK_COUNT = 3;
constexpr uint32_t IRCACHE_OCTA_DIMS = 4, IRCACHE_OCTA_DIMS2 = 16, IRCACHE_IRRADIANCE_STRIDE = 3, IRCACHE_AUX_STRIDE = 64;
constexpr uint32_t IRCACHE_SAMPLES_PER_FRAME = 4, IRCACHE_VALIDATION_SAMPLES_PER_FRAME = 4, IRCACHE_RESTIR_M_CLAMP = 30;
KJB_DEV bool is_ircache_entry_life_valid(uint32_t life) { return life < IRCACHE_ENTRY_LIFE_PER_RANK * IRCACHE_ENTRY_RANK_COUNT; }
KJB_DEV uint32_t ircache_entry_life_to_rank(uint32_t life) { return life / IRCACHE_ENTRY_LIFE_PER_RANK; }
KJB_DEV uint32_t ircache_entry_life_for_rank(uint32_t rank) { return rank * IRCACHE_ENTRY_LIFE_PER_RANK; }

// ---- device atomics (the CPU test emulator maps them to the compiler builtins)
#if defined(__CUDA_ARCH__)
KJB_DEV uint32_t atom_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
KJB_DEV uint32_t atom_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
KJB_DEV uint32_t atom_and(uint32_t* p, uint32_t v) { return atomicAnd(p, v); }
KJB_DEV uint32_t atom_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
KJB_DEV uint32_t atom_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
#else
KJB_DEV uint32_t atom_add(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
KJB_DEV uint32_t atom_or(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
KJB_DEV uint32_t atom_and(uint32_t* p, uint32_t v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
KJB_DEV uint32_t atom_min(uint32_t* p, uint32_t v) { uint32_t c = *p; while (c > v && !__atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return c; }
KJB_DEV uint32_t atom_max(uint32_t* p, uint32_t v) { uint32_t c = *p; while (c < v && !__atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return c; }
#endif

// DEFINE_IRCACHE_BINDINGS (ircache/bindings.hlsl) + the aux buffer for IRCACHE_LOOKUP_PRECISE
struct IrcacheBufs {
    uint32_t *meta, *pool, *reposition_count, *grid_meta, *entry_cell, *life;
    float4 *reposition_proposal, *spatial, *irradiance, *aux;
    KJB_DEV bool bound() const { return meta != nullptr; }
};
inline IrcacheBufs ircache_bufs(const kjb_ircache_bindings& b) {
    IrcacheBufs r; r.meta = (uint32_t*)b.meta_buf.data; r.pool = (uint32_t*)b.pool_buf.data; r.reposition_count = (uint32_t*)b.reposition_proposal_count_buf.data;
    r.grid_meta = (uint32_t*)b.grid_meta_buf.data; r.entry_cell = (uint32_t*)b.entry_cell_buf.data; r.life = (uint32_t*)b.life_buf.data;
    r.reposition_proposal = (float4*)b.reposition_proposal_buf.data; r.spatial = (float4*)b.spatial_buf.data; r.irradiance = (float4*)b.irradiance_buf.data; r.aux = nullptr;   /* not in the binding block: the rtdgi / rtr lookups are never IRCACHE_LOOKUP_PRECISE */
    return r;
}

// ---- ircache_grid.hlsl
struct IrcacheCoord { uint32_t cx, cy, cz, cascade; };
KJB_DEV uint32_t ircache_cell_idx(const IrcacheCoord& c) { return c.cx + c.cy * 32u + c.cz * 1024u + c.cascade * 32768u; }
KJB_DEV uint32_t ws_local_pos_to_cascade_idx(float3 local_pos, uint32_t reserved_cells) {   // :34-39
    const float3 fcoord = local_pos / KJB_IRCACHE_GRID_CELL_DIAMETER;
    const float max_coord = kjb_max(kjb_abs(fcoord.x), kjb_max(kjb_abs(fcoord.y), kjb_abs(fcoord.z)));
    const float cascade_float = kjb_log2(max_coord / float(IRCACHE_CASCADE_SIZE / 2 - reserved_cells));
    return kjb_cvt_u32(kjb_clamp(kjb_ceil(kjb_max(0.0f, cascade_float)), 0.0f, float(IRCACHE_CASCADE_COUNT - 1)));
}
KJB_DEV float ircache_grid_cell_diameter_in_cascade(uint32_t cascade) { return KJB_IRCACHE_GRID_CELL_DIAMETER * float(1u << cascade); }
KJB_DEV IrcacheCoord ws_pos_to_ircache_coord(const kjb_frame_constants& fc, float3 pos, float3 normal, float3 jitter) {   // :41-75
    const float3 center = f3(fc.ircache_grid_center[0], fc.ircache_grid_center[1], fc.ircache_grid_center[2]);
    const uint32_t reserved_cells = 1;
    {
        const uint32_t cascade = ws_local_pos_to_cascade_idx(pos - center, reserved_cells);
        pos = pos + ircache_grid_cell_diameter_in_cascade(cascade) * jitter;
    }
    const uint32_t cascade = ws_local_pos_to_cascade_idx(pos - center, reserved_cells);
    const float cell_diameter = ircache_grid_cell_diameter_in_cascade(cascade);
    const int32_t* co = fc.ircache_cascades[cascade].origin;
    const float3 cell_offset = normal * cell_diameter * 0.5f;
    const float3 q = (pos + cell_offset) / cell_diameter;
    const int ix = kjb_cvt_i32(kjb_floor(q.x)) - co[0], iy = kjb_cvt_i32(kjb_floor(q.y)) - co[1], iz = kjb_cvt_i32(kjb_floor(q.z)) - co[2];
    IrcacheCoord r; r.cascade = cascade;
    r.cx = uint32_t(ix < 0 ? 0 : (ix > 31 ? 31 : ix)); r.cy = uint32_t(iy < 0 ? 0 : (iy > 31 ? 31 : iy)); r.cz = uint32_t(iz < 0 ? 0 : (iz > 31 ? 31 : iz));
    return r;
}

// ---- pack_unpack.hlsl:79-88 / ircache_sampler_common.inc.hlsl
KJB_DEV float3 octa_decode(float2 f) {
    f = f * 2.0f - 1.0f;
    float3 n = f3(f.x, f.y, 1.0f - kjb_abs(f.x) - kjb_abs(f.y));
    const float t = kjb_clamp(-n.z, 0.0f, 1.0f);
    n.x -= (kjb_step(0.0f, n.x) * 2 - 1) * t;
    n.y -= (kjb_step(0.0f, n.y) * 2 - 1) * t;
    return normalize(n);
}
struct SampleParams {
    uint32_t value;
    KJB_DEV static SampleParams from_spf_entry_sample_frame(uint32_t spf, uint32_t entry_idx, uint32_t sample_idx, uint32_t frame_idx) {
        const uint32_t PERIOD = IRCACHE_OCTA_DIMS2 / spf;
        uint32_t xy = sample_idx * PERIOD + (frame_idx % PERIOD);
        xy ^= (xy & 4u) >> 2u;   // checkerboard
        SampleParams r; r.value = xy + ((frame_idx << 16u) ^ entry_idx) * IRCACHE_OCTA_DIMS2; return r;
    }
    KJB_DEV static SampleParams from_raw(uint32_t raw) { SampleParams r; r.value = raw; return r; }
    KJB_DEV uint32_t octa_idx() const { return value % IRCACHE_OCTA_DIMS2; }
    KJB_DEV uint32_t rng() const { return hash1(value >> 4u); }
    KJB_DEV float3 direction() const {
        const uint32_t oi = octa_idx();
        const float2 urand = r2_sequence(rng() % 1024u);
        return octa_decode((f2(float(oi % IRCACHE_OCTA_DIMS), float(oi / IRCACHE_OCTA_DIMS)) + urand) / 4.0f);
    }
};

// ---- inc/mesh.hlsl:25-46
struct IrcacheVertex { float3 position, normal; };
KJB_DEV IrcacheVertex unpack_vertex(float4 p) {
    const uint32_t pck = kjb_f2u(p.w);
    IrcacheVertex v; v.position = xyz(p);
    v.normal = f3(float(pck & 2047u) * (2.0f / 2047.0f) - 1.0f, float((pck >> 11u) & 1023u) * (2.0f / 1023.0f) - 1.0f, float(pck >> 21u) * (2.0f / 2047.0f) - 1.0f);
    return v;
}
KJB_DEV float4 pack_vertex(const IrcacheVertex& v) { return f4(v.position, kjb_u2f(pack_normal_11_10_11(v.normal))); }

KJB_DEV float eval_sh_geometrics(float4 sh, float3 normal) {   // lookup.hlsl:197-212
    const float R0 = sh.x;
    const float3 R1 = 0.5f * f3(sh.y, sh.z, sh.w);
    const float lenR1 = length(R1);
    const float q = 0.5f * (1.0f + dot(R1 / lenR1, normal));
    const float p = 1.0f + 2.0f * lenR1 / R0;
    const float a = (1.0f - lenR1 / R0) / (1.0f + lenR1 / R0);
    return R0 * (a + (1.0f - a) * (p + 1.0f) * kjb_pow(q, p));
}

// IrcacheLookupParams::lookup (lookup.hlsl:76-311) with the maybe-allocating lookup it wraps (:19-74,:120-190).
// PRECISE = IRCACHE_LOOKUP_PRECISE: sum the per-direction reservoirs instead of evaluating the SH (the cache's own tracing passes).
template <bool PRECISE>
KJB_DEV float3 ircache_lookup(const Globals& g, const IrcacheBufs& b, float3 query_from_ws, float3 pt_ws, float3 normal_ws, uint32_t query_rank, uint32_t& rng, bool stochastic_interpolation = false) {
    if (!b.bound()) return f3(0.0f);
    const kjb_frame_constants& fc = g.fc;
    bool allocated_by_us = false, just_allocated = false;
    // lookup.hlsl:80-86: the three rng draws sit in the arguments of select(), i.e. they happen whether or not interpolation is on
    float3 jr; jr.x = rand01(rng); jr.y = rand01(rng); jr.z = rand01(rng);
    const IrcacheCoord rc = ws_pos_to_ircache_coord(fc, pt_ws, normal_ws, stochastic_interpolation ? jr - 0.5f : f3(0.0f));
    const uint32_t cell_idx = ircache_cell_idx(rc);
    {
        const int32_t* so = fc.ircache_cascades[rc.cascade].voxels_scrolled_this_frame;
        const int c3[3] = {int(rc.cx), int(rc.cy), int(rc.cz)};
        bool was_just_scrolled_in = false;
        for (int k = 0; k < 3; ++k) was_just_scrolled_in = was_just_scrolled_in || (so[k] > 0 ? (c3[k] + so[k] >= int(IRCACHE_CASCADE_SIZE)) : (c3[k] < -so[k]));
        const bool skip_allocation = query_rank >= IRCACHE_ENTRY_RANK_COUNT || (was_just_scrolled_in && query_rank > 0);
        const uint32_t entry_flags = b.grid_meta[cell_idx * 2 + 1];
        just_allocated = (entry_flags & IRCACHE_ENTRY_META_JUST_ALLOCATED) != 0;
        if (!skip_allocation && (entry_flags & IRCACHE_ENTRY_META_OCCUPIED) == 0) {
            const uint32_t prev = atom_or(&b.grid_meta[cell_idx * 2 + 1], IRCACHE_ENTRY_META_OCCUPIED | IRCACHE_ENTRY_META_JUST_ALLOCATED);
            if ((prev & IRCACHE_ENTRY_META_OCCUPIED) == 0) {   // we claimed the cell
                just_allocated = true; allocated_by_us = true;
                const uint32_t alloc_idx = atom_add(&b.meta[IRCACHE_META_ALLOC_COUNT], 1u);
                if (alloc_idx >= 1024u * 64u) {   // pool exhausted: undo
                    atom_add(&b.meta[IRCACHE_META_ALLOC_COUNT], uint32_t(-1));
                    atom_and(&b.grid_meta[cell_idx * 2 + 1], ~(IRCACHE_ENTRY_META_OCCUPIED | IRCACHE_ENTRY_META_JUST_ALLOCATED));
                } else {
                    const uint32_t entry_idx = b.pool[alloc_idx];
                    atom_max(&b.meta[IRCACHE_META_ENTRY_COUNT], entry_idx + 1);
                    b.life[entry_idx] = ircache_entry_life_for_rank(query_rank);   // clear dead state, mark used
                    b.entry_cell[entry_idx] = cell_idx;
                    b.grid_meta[cell_idx * 2 + 0] = entry_idx;
                }
            }
        }
    }
    uint32_t lookup_count = 0, lookup_entry = 0;
    if (b.grid_meta[cell_idx * 2 + 1] & IRCACHE_ENTRY_META_OCCUPIED) { lookup_entry = b.grid_meta[cell_idx * 2 + 0]; lookup_count = 1; }

    const float cell_diameter = ircache_grid_cell_diameter_in_cascade(rc.cascade);
    float3 offset_towards_query = query_from_ws - pt_ws;
    const float MAX_OFFSET = cell_diameter, MAX_OFFSET_AS_FRAC = 0.5f;
    offset_towards_query = offset_towards_query * (MAX_OFFSET / kjb_max(MAX_OFFSET / MAX_OFFSET_AS_FRAC, length(offset_towards_query)));
    IrcacheVertex new_entry; new_entry.position = pt_ws + offset_towards_query; new_entry.normal = normal_ws;
    if (allocated_by_us && lookup_count) b.reposition_proposal[lookup_entry] = pack_vertex(new_entry);
    if (just_allocated) return f3(0.0f);

    float3 irradiance_sum = f3(0.0f);
    if (lookup_count) {
        const uint32_t entry_idx = lookup_entry;
        float3 irradiance = f3(0.0f);
        if (PRECISE) {
            float weight_sum = 0;
            for (uint32_t octa_idx = 0; octa_idx < IRCACHE_OCTA_DIMS2; ++octa_idx) {
                const float4 ra = b.aux[entry_idx * IRCACHE_AUX_STRIDE + octa_idx];
                const Reservoir r = Reservoir::from_raw(u2(kjb_f2u(ra.x), kjb_f2u(ra.y)));
                const float3 dir = SampleParams::from_raw(r.payload).direction();
                const float wt = dot(dir, normal_ws);
                if (wt > 0.0f) {
                    const float4 contrib = b.aux[entry_idx * IRCACHE_AUX_STRIDE + IRCACHE_OCTA_DIMS2 + octa_idx];
                    irradiance += xyz(contrib) * wt * contrib.w;
                    weight_sum += wt;
                }
            }
            irradiance = irradiance / kjb_max(1.0f, weight_sum);
        } else {
            irradiance.x += eval_sh_geometrics(b.irradiance[entry_idx * 3 + 0], normal_ws);
            irradiance.y += eval_sh_geometrics(b.irradiance[entry_idx * 3 + 1], normal_ws);
            irradiance.z += eval_sh_geometrics(b.irradiance[entry_idx * 3 + 2], normal_ws);
        }
        irradiance = vmax(f3(0.0f), irradiance);
        irradiance_sum += irradiance * 1.0f;
        const uint32_t prev_life = b.life[entry_idx];
        if (prev_life < IRCACHE_ENTRY_LIFE_RECYCLE) {
            const uint32_t new_life = ircache_entry_life_for_rank(query_rank);
            if (new_life < prev_life) atom_min(&b.life[entry_idx], new_life);
            const uint32_t prev_rank = ircache_entry_life_to_rank(prev_life);
            if (query_rank <= prev_rank) {   // IRCACHE_USE_POSITION_VOTING + IRCACHE_USE_UNIFORM_VOTING
                const uint32_t prev_vote_count = atom_add(&b.reposition_count[entry_idx], 1u);
                const float dart = rand01(rng);
                const float prob = 1.0f / (float(prev_vote_count) + 1.0f);
                if (dart <= prob) b.reposition_proposal[entry_idx] = pack_vertex(new_entry);
            }
        }
    }
    return irradiance_sum;
}

}  // namespace kjb
