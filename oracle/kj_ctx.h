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
};

namespace kjo {

// Row-parallel loop over a WxH grid (oracle passes are embarrassingly parallel over pixels;
// every pass reads inputs and writes DIFFERENT output images, except the in-place validate pass
// which only touches its own pixel).
inline void parallel_rows(int h, const std::function<void(int)>& row_fn, int nthreads = 0) {
    if (nthreads <= 0) nthreads = int(std::thread::hardware_concurrency());
    if (nthreads <= 1 || h < 8) { for (int y = 0; y < h; ++y) row_fn(y); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&]() { for (;;) { int y = next.fetch_add(1); if (y >= h) break; row_fn(y); } });
    for (auto& t : th) t.join();
}

inline float4 f4(const float* p) { return float4(p[0], p[1], p[2], p[3]); }

}  // namespace kjo
