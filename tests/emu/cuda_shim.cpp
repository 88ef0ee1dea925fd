// TEST INFRASTRUCTURE ONLY (see cuda_shim.h)
#include "cuda_shim.h"
thread_local kjb_emu::Idx threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
int kjb_emu::g_serial = 0;
int kjb_emu::num_workers() { static int n = [] { const char* e = getenv("KJB_EMU_THREADS"); int v = e ? atoi(e) : int(std::thread::hardware_concurrency()); return v < 1 ? 1 : v; }(); return n; }

#include <ucontext.h>
#include <memory>
namespace {
struct Fiber { ucontext_t ctx; std::unique_ptr<char[]> stack; bool done = true; };
struct BlockState {
    std::vector<Fiber> fibers; ucontext_t sched; unsigned cur = 0;
    void (*body)(void*) = nullptr; void* arg = nullptr;
};
thread_local BlockState* g_block = nullptr;
const size_t STACK = 256 * 1024;
void trampoline() {
    BlockState* b = g_block;
    b->body(b->arg);
    b->fibers[b->cur].done = true;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}
}
void kjb_emu::fiber_barrier() { BlockState* b = g_block; swapcontext(&b->fibers[b->cur].ctx, &b->sched); }
void kjb_emu::fiber_run_block(unsigned nthreads, dim3 block, void (*thread_body)(void*), void* arg) {
    static thread_local BlockState state;
    BlockState& b = state; g_block = &b; b.body = thread_body; b.arg = arg;
    if (b.fibers.size() < nthreads) { b.fibers.resize(nthreads); for (auto& f : b.fibers) if (!f.stack) f.stack.reset(new char[STACK]); }
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = b.fibers[t];
        getcontext(&f.ctx); f.ctx.uc_stack.ss_sp = f.stack.get(); f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0); f.done = false;
    }
    for (bool any = true; any;) {
        any = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            if (b.fibers[t].done) continue;
            any = true; b.cur = t;
            threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
            swapcontext(&b.sched, &b.fibers[t].ctx);
        }
    }
}
