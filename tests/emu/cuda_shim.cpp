// TEST INFRASTRUCTURE ONLY (see cuda_shim.h)
#include "cuda_shim.h"
thread_local kjb_emu::Idx threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
int kjb_emu::num_workers() { static int n = [] { const char* e = getenv("KJB_EMU_THREADS"); int v = e ? atoi(e) : int(std::thread::hardware_concurrency()); return v < 1 ? 1 : v; }(); return n; }
