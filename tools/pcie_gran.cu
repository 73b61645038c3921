// pcie_gran.cu -- experiment: how many bytes cross PCIe when a kernel reads ONE byte per `stride` bytes of page-locked host
// memory in place, for the load flavours a gather could use.  Built by tools/pcie_gran.py (not part of libvfk.so).
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void gather(const uint8_t *__restrict__ p, int64_t n, int64_t stride, const uint32_t *__restrict__ perm, unsigned long long *out) {
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *a = p + (int64_t)perm[i] * stride + (perm[i] & 31);
        uint32_t v;
        if (MODE == 0) v = __ldg(a);                                                             // ld.global.nc
        else if (MODE == 1) v = *(const volatile uint8_t *)a;                                    // ld.volatile
        else if (MODE == 2) asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(a));
        else if (MODE == 3) asm volatile("ld.global.cs.u8 %0, [%1];" : "=r"(v) : "l"(a));
        else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(a));
        else asm volatile("ld.global.cv.u8 %0, [%1];" : "=r"(v) : "l"(a));
        acc += v;
    }
    if (acc == 0xFFFFFFFFFFFFull) *out = acc;
}
extern "C" int pcie_gather(int mode, const void *dev_alias, int64_t n, int64_t stride, const uint32_t *perm, void *out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = 148 * 8, block = 256;
    switch (mode) {
        case 0: gather<0><<<grid, block, 0, st>>>((const uint8_t *)dev_alias, n, stride, perm, (unsigned long long *)out); break;
        case 1: gather<1><<<grid, block, 0, st>>>((const uint8_t *)dev_alias, n, stride, perm, (unsigned long long *)out); break;
        case 2: gather<2><<<grid, block, 0, st>>>((const uint8_t *)dev_alias, n, stride, perm, (unsigned long long *)out); break;
        case 3: gather<3><<<grid, block, 0, st>>>((const uint8_t *)dev_alias, n, stride, perm, (unsigned long long *)out); break;
        case 4: gather<4><<<grid, block, 0, st>>>((const uint8_t *)dev_alias, n, stride, perm, (unsigned long long *)out); break;
        default: gather<5><<<grid, block, 0, st>>>((const uint8_t *)dev_alias, n, stride, perm, (unsigned long long *)out); break;
    }
    return (int)cudaGetLastError();
}
extern "C" int pcie_set_granularity(int bytes) { return (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes); }
