// Screen-space tile staging for the stencil / gather kernels: one TMA bulk-tensor copy (cp.async.bulk.tensor.2d, completion on an
// mbarrier) brings the (block + apron) footprint of an image into shared memory — the copy engine does the addressing, the bounds
// handling (texels outside the image arrive as zeros, which is exactly the ABI's "reads outside an image return 0") and the wait costs one
// elected thread a handful of instructions instead of every thread a guarded LDG per texel.  Warp-level exchange helpers (SHFL) live here
// too.  Under the test-only CPU launch emulator (KJB_EMU) the same calls are plain loops / a shared scratch array.
//
// Tensor maps are built on the host (kjb::tile_source, kjb_api.cu) with cuTensorMapEncodeTiled fetched through
// cudaGetDriverEntryPoint (no libcuda link).  TMA needs 16-byte aligned rows: images whose row pitch is not a multiple of 16 bytes
// (odd test extents) get `use_tma = 0` and the kernels stage the tile with cooperative guarded loads instead — same contents.
#pragma once
#include "kjb_device.cuh"
#if !defined(KJB_EMU)
#include <cuda.h>
#endif

#if defined(KJB_EMU)
#define KJB_DEVONLY inline
#else
#define KJB_DEVONLY __device__ __forceinline__
#endif

namespace kjb {

#if defined(KJB_EMU)
struct TileSource { int use_tma; };
#else
struct alignas(64) TileSource { CUtensorMap map; int use_tma; int pad[15]; };
#endif

// number of texels of a tile row as staged in shared memory: the box width rounded up so that a row is a whole number of 16-byte units
template <int TEXEL_BYTES> constexpr int tile_pitch(int w) { return ((w * TEXEL_BYTES + 15) / 16 * 16) / TEXEL_BYTES; }

#if !defined(KJB_EMU)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KJB_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KJB_MBAR_DONE;\n"
        "bra KJB_MBAR_WAIT;\n"
        "KJB_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
#endif

// Staging protocol (every thread of the block executes all three calls, uniform control flow):
//     tile_group_begin(bar, phase, use_tma, total_bytes, tid);          // arms the block's mbarrier with the bytes of ALL tiles of the group
//     tile_issue<T, TW, TH>(dst, src, img, x0, y0, bar, use_tma, tid, nthreads);   // once per tile: TW x TH texels whose top-left texel is (x0, y0)
//     tile_group_wait(bar, phase, use_tma);                             // on return every tile of the group is visible to the whole block
// `dst` has a row pitch of tile_pitch<sizeof(T)>(TW) texels and is 128-byte aligned; `phase` counts the groups this block has already
// pushed through `bar` (0 for the first); `use_tma` is one flag per launch (the host clears it unless every source qualifies).
template <typename T, int TW, int TH> constexpr uint32_t tile_bytes() { return uint32_t(tile_pitch<int(sizeof(T))>(TW) * TH * sizeof(T)); }

KJB_DEVONLY void tile_group_begin(uint64_t* bar, uint32_t phase, int use_tma, uint32_t total_bytes, int tid) {
#if !defined(KJB_EMU)
    if (use_tma) {
        if (phase == 0) { if (tid == 0) mbar_init(bar, 1); __syncthreads(); }
        if (tid == 0) mbar_expect_tx(bar, total_bytes);
    }
#endif
}
template <typename T, int TW, int TH>
KJB_DEVONLY void tile_issue(T* dst, const TileSource& src, const Img& img, int x0, int y0, uint64_t* bar, int use_tma, int tid, int nthreads) {
    constexpr int PITCH = tile_pitch<int(sizeof(T))>(TW);
#if !defined(KJB_EMU)
    if (use_tma) {
        // the map describes the image as rows of 32-bit words (4/8/16-byte texels) or of its 1/2-byte elements: scale the x coordinate
        if (tid == 0) tma_load_2d(dst, &src.map, x0 * (sizeof(T) >= 4 ? int(sizeof(T) / 4) : 1), y0, bar);
        return;
    }
#endif
    for (int i = tid; i < PITCH * TH; i += nthreads) {
        const int lx = i % PITCH, ly = i / PITCH;
        T v; memset(&v, 0, sizeof(T));
        if (inb(img, x0 + lx, y0 + ly)) v = ld_raw<T>(img, x0 + lx, y0 + ly);
        dst[i] = v;
    }
}
KJB_DEVONLY void tile_group_wait(uint64_t* bar, uint32_t phase, int use_tma) {
#if !defined(KJB_EMU)
    if (use_tma) { mbar_wait(bar, phase & 1u); return; }
#endif
    __syncthreads();
}

// ---- warp exchange: value of lane (lane ^ mask) of the same warp.  The blocks that use it are 32 threads wide, so a warp is one row of
// the block and lane == threadIdx.x.  Emulator: through a per-block scratch array with two barriers (all threads of the block take part).
KJB_DEVONLY float warp_xor(float v, int mask) {
#if !defined(KJB_EMU)
    return __shfl_xor_sync(0xffffffffu, v, mask);
#else
    static thread_local float scratch[1024];
    const int t = int(threadIdx.x + threadIdx.y * blockDim.x);
    scratch[t] = v; __syncthreads();
    const float r = scratch[(t & ~31) | ((t & 31) ^ mask)]; __syncthreads();
    return r;
#endif
}

}  // namespace kjb
