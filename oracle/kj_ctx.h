// ORACLE — TEST INFRASTRUCTURE ONLY (see kj_math.h header).
#pragma once
#include "kj_scene.h"
#include <string>
#include <functional>
#include <thread>

struct kjb_context {
    kjo::Scene scene;
    kjo::Globals g;
    std::string last_error;
    int num_threads = 0;
    // Cache-touching passes: serial in launch order by default (the deterministic schedule every parity test compares against).
    // kjb_set_debug_serial(ctx, 0) lets them run on all host threads — racy by design like the reference's GPU dispatch (the lookups use
    // the same atomics), used ONLY by bench.py's CPU-baseline / `--impl reference` legs so that the CPU arm really has every core.
    bool cache_passes_parallel = false;
    kjb_allgather_fn ag_fn = nullptr; void* ag_user = nullptr; uint32_t rank = 0, nranks = 1;
    uint32_t scissor_y0 = 0, scissor_y1 = 0;   // kjb_set_scissor: rows [y0, y1) of the pass's output grid (0,0 = all)
};

namespace kjo {

// Row-parallel loop over a WxH grid (oracle passes are embarrassingly parallel over pixels;
// every pass reads inputs and writes DIFFERENT output images, except the in-place validate pass
// which only touches its own pixel).
// rows of a pass's grid to compute under the context's scissor
inline void scissor_rows(const kjb_context* c, int h, int& y0, int& y1) {
    y0 = 0; y1 = h;
    if (c->scissor_y1 > c->scissor_y0) { y0 = int(c->scissor_y0) < h ? int(c->scissor_y0) : h; y1 = int(c->scissor_y1) < h ? int(c->scissor_y1) : h; }
}
inline void parallel_rows_range(int y0, int y1, const std::function<void(int)>& row_fn, int nthreads = 0);
inline void parallel_rows(int h, const std::function<void(int)>& row_fn, int nthreads = 0) {
    if (nthreads <= 0) nthreads = int(std::thread::hardware_concurrency());
    if (nthreads <= 1 || h < 8) { for (int y = 0; y < h; ++y) row_fn(y); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&]() { for (;;) { int y = next.fetch_add(1); if (y >= h) break; row_fn(y); } });
    for (auto& t : th) t.join();
}

inline void parallel_rows_range(int y0, int y1, const std::function<void(int)>& row_fn, int nthreads) {
    parallel_rows(y1 - y0, [&](int y) { row_fn(y + y0); }, nthreads);
}
// every oracle pass iterates its output rows through this: honours the tile scissor
inline void pass_rows(const kjb_context* ctx, int H, const std::function<void(int)>& row_fn, int nthreads = 0) {
    int y0, y1; scissor_rows(ctx, H, y0, y1);
    parallel_rows_range(y0, y1, row_fn, nthreads);
}

// Passes that touch the irradiance cache run on ONE thread in the order a serialised GPU launch would execute them: 8x16 pixel
// blocks, row-major over blocks, row-major inside a block (the product's launch shape for the two rtdgi ray passes).
inline void pass_pixels(const kjb_context* ctx, int W, int H, bool serial_tiles, const std::function<void(int, int)>& px_fn) {
    if (!serial_tiles || ctx->cache_passes_parallel) { pass_rows(ctx, H, [&](int y) { for (int x = 0; x < W; ++x) px_fn(x, y); }, ctx->num_threads); return; }
    int y0, y1; scissor_rows(ctx, H, y0, y1);
    for (int by = y0; by < y1; by += 16) for (int bx = 0; bx < W; bx += 8)
        for (int y = by; y < by + 16 && y < y1; ++y) for (int x = bx; x < bx + 8 && x < W; ++x) px_fn(x, y);
}

// 1-D cache passes (one item per entry sample): serial index order, or chunks of 64 items over the host threads in the parallel schedule
inline void pass_items(const kjb_context* ctx, uint32_t n, const std::function<void(uint32_t)>& fn) {
    if (!ctx->cache_passes_parallel) { for (uint32_t i = 0; i < n; ++i) fn(i); return; }
    parallel_rows(int((n + 63) / 64), [&](int c) { const uint32_t e = uint32_t(c) * 64 + 64 < n ? uint32_t(c) * 64 + 64 : n; for (uint32_t i = uint32_t(c) * 64; i < e; ++i) fn(i); }, ctx->num_threads);
}

inline float4 f4(const float* p) { return float4(p[0], p[1], p[2], p[3]); }

}  // namespace kjo
