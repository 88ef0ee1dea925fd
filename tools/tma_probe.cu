// Probe: which tensor-map / kernel-parameter shapes does UTMALDG accept on this GPU?  (development aid for kjb_tile.cuh)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/tma_probe tools/tma_probe.cu ; tools/_bin/tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
#include <cuda/barrier>
#include <cuda/ptx>
namespace cde = cuda::device::experimental;
using barrier_t = cuda::barrier<cuda::thread_scope_block>;
template <int BW, int BH>
__global__ void k_libcu(const __grid_constant__ CUtensorMap map, uint32_t* out, int x0, int y0) {
    __shared__ alignas(128) uint32_t tile[BW * BH];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier_t::arrival_token token;
    if (threadIdx.x == 0) { cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &map, x0, y0, bar); token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile)); }
    else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
template <int BW, int BH>
__global__ void k_global_desc(const CUtensorMap* map, uint32_t* out, int x0, int y0) {
    __shared__ __align__(128) uint32_t tile[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, BW * BH * 4); tma_load_2d(tile, map, x0, y0, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_mbar_only(uint32_t* out) {
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, 0);
    mbar_wait(&bar, 0);
    out[threadIdx.x] = 7u + threadIdx.x;
}
__global__ void k_bulk_1d(const uint32_t* src, uint32_t* out) {   // cp.async.bulk (no tensor map): 2 KB from global to shared
    __shared__ __align__(128) uint32_t tile[512];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 2048);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tile)), "l"(src), "r"(2048), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = tile[i];
}
struct alignas(64) Wrapped { CUtensorMap map; int flag; int pad[15]; };

template <int BW, int BH>
__global__ void k_direct(const __grid_constant__ CUtensorMap map, uint32_t* out, int x0, int y0) {
    __shared__ __align__(128) uint32_t tile[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, BW * BH * 4); tma_load_2d(tile, &map, x0, y0, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
template <int BW, int BH>
__global__ void k_wrapped(const __grid_constant__ Wrapped w, uint32_t* out, int x0, int y0) {
    __shared__ __align__(128) uint32_t tile[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, BW * BH * 4); tma_load_2d(tile, &w.map, x0, y0, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 320, H = 90;   // words per row, rows
    std::vector<uint32_t> h(W * H); for (int i = 0; i < W * H; ++i) h[i] = 0x10000u + i;
    uint32_t *d, *out; cudaMalloc(&d, W * H * 4); cudaMalloc(&out, 256 * 64 * 4); cudaMemcpy(d, h.data(), W * H * 4, cudaMemcpyHostToDevice); cudaMemset(out, 0xff, 256 * 64 * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int BW = (variant & (1 | 32 | 64)) ? 64 : 68, BH = (variant & (2 | 32 | 64)) ? 8 : 10;
    CUtensorMap m; memset(&m, 0, sizeof m);
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 4}; cuuint32_t box[2] = {cuuint32_t(BW), cuuint32_t(BH)}, es[2] = {1, 1};
    const int dtype = argc > 2 ? atoi(argv[2]) : int(CU_TENSOR_MAP_DATA_TYPE_UINT32), swz = argc > 3 ? atoi(argv[3]) : 0, rank = argc > 4 ? atoi(argv[4]) : 2;
    cuuint64_t dims3[3] = {W, H, 1}, strides3[2] = {W * 4, cuuint64_t(W) * H * 4}; cuuint32_t box3[3] = {cuuint32_t(BW), cuuint32_t(BH), 1}, es3[3] = {1, 1, 1};
    CUresult r = enc(&m, CUtensorMapDataType(dtype), rank, d, rank == 3 ? dims3 : dims, rank == 3 ? strides3 : strides, rank == 3 ? box3 : box, rank == 3 ? es3 : es, CU_TENSOR_MAP_INTERLEAVE_NONE, CUtensorMapSwizzle(swz),
                     (variant & 8) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  dtype %d swizzle %d rank %d\n", dtype, swz, rank);
    printf("variant %d: box %dx%d encode rc=%d query=%d\n", variant, BW, BH, int(r), int(q));
    const int x0 = (variant & 16) ? -2 : 30, y0 = (variant & 16) ? -1 : 7;
    if (variant & 128) { k_mbar_only<<<1, 128>>>(out); cudaError_t e2 = cudaDeviceSynchronize(); printf("  mbarrier only: %s\n", cudaGetErrorString(e2)); return 0; }
    if (variant & 256) {
        k_bulk_1d<<<1, 128>>>(d, out); cudaError_t e2 = cudaDeviceSynchronize(); printf("  cp.async.bulk 1-D: %s\n", cudaGetErrorString(e2));
        if (e2 == cudaSuccess) { std::vector<uint32_t> o(512); cudaMemcpy(o.data(), out, 2048, cudaMemcpyDeviceToHost); int bad = 0; for (int i = 0; i < 512; ++i) bad += o[i] != h[i]; printf("  mismatches: %d\n", bad); }
        return 0;
    }
    { const unsigned char* mb = (const unsigned char*)&m; printf("  tensor map bytes:"); for (int i = 0; i < 128; ++i) { if (i % 32 == 0) printf("\n   "); printf(" %02x", mb[i]); } printf("\n"); }
    if (variant & 32) {          // NVIDIA's libcu++ wrapper, descriptor as __grid_constant__ parameter
        k_libcu<64, 8><<<1, 128>>>(m, out, x0, y0);
    } else if (variant & 64) {   // descriptor in global memory
        CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        k_global_desc<64, 8><<<1, 128>>>(dm, out, x0, y0);
    } else if (variant & 4) {
        Wrapped w; memset(&w, 0, sizeof w); w.map = m; w.flag = 1;
        if (BW == 68 && BH == 10) k_wrapped<68, 10><<<1, 128>>>(w, out, x0, y0); else if (BW == 64 && BH == 8) k_wrapped<64, 8><<<1, 128>>>(w, out, x0, y0);
        else if (BW == 68) k_wrapped<68, 8><<<1, 128>>>(w, out, x0, y0); else k_wrapped<64, 10><<<1, 128>>>(w, out, x0, y0);
    } else {
        if (BW == 68 && BH == 10) k_direct<68, 10><<<1, 128>>>(m, out, x0, y0); else if (BW == 64 && BH == 8) k_direct<64, 8><<<1, 128>>>(m, out, x0, y0);
        else if (BW == 68) k_direct<68, 8><<<1, 128>>>(m, out, x0, y0); else k_direct<64, 10><<<1, 128>>>(m, out, x0, y0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("  sync: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint32_t> o(BW * BH); cudaMemcpy(o.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
            const int gx = x0 + x, gy = y0 + y; const uint32_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * W + gx] : 0u;
            if (o[y * BW + x] != want) ++bad;
        }
        printf("  mismatches: %d of %d\n", bad, BW * BH);
    }
    return 0;
}
