// What FMNMX does with signed zeros and NaNs on this GPU (decides whether kjb_min/kjb_max may lower to it; see include/kjb_numeric.h).
#include <cstdio>
#include <cstdint>
#include <cstring>
__global__ void k(const float* a, const float* b, float* mn, float* mx, float* sat, int n) {
    int i = threadIdx.x; if (i >= n) return;
    mn[i] = fminf(a[i], b[i]); mx[i] = fmaxf(a[i], b[i]); sat[i] = __saturatef(a[i]);
}
int main() {
    const uint32_t NANQ = 0x7fc00000u, NANS = 0xffc12345u;
    uint32_t av[] = {0x00000000u, 0x80000000u, 0x00000000u, 0x80000000u, NANQ, 0x3f800000u, NANS, 0x80000000u, NANQ, 0x00000001u, 0x80000001u};
    uint32_t bv[] = {0x80000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x3f800000u, NANQ, 0x80000000u, NANS, NANS, 0x80000001u, 0x00000001u};
    const int n = sizeof(av) / 4;
    float *a, *b, *mn, *mx, *st;
    cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&mn, n * 4); cudaMallocManaged(&mx, n * 4); cudaMallocManaged(&st, n * 4);
    memcpy(a, av, n * 4); memcpy(b, bv, n * 4);
    k<<<1, 32>>>(a, b, mn, mx, st, n);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    for (int i = 0; i < n; ++i) { uint32_t r0, r1, r2; memcpy(&r0, mn + i, 4); memcpy(&r1, mx + i, 4); memcpy(&r2, st + i, 4);
        printf("a=%08x b=%08x  fminf=%08x fmaxf=%08x  saturate(a)=%08x\n", av[i], bv[i], r0, r1, r2); }
    return 0;
}
