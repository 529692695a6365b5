// micro-benchmark: latency of bringing one 32 KiB tile from DRAM into shared memory on an otherwise idle /
// moderately loaded SM: (a) one cp.async.bulk, (b) 8 bulk copies of 4 KiB, (c) cp.async 16 B per thread,
// (d) LDG.128 + STS.128.   nvcc -arch=sm_100a -O3 load_lat.cu -o load_lat && ./load_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TILE = 32768, NT = 256;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(NT) k(const uint8_t *in, size_t n_tiles, int mode, int iters, unsigned long long *out, unsigned *sink) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long acc = 0;
    unsigned x = 0;
    for (int it = 0; it < iters; it++) {
        size_t t = ((size_t)blockIdx.x + (size_t)it * gridDim.x) % n_tiles;
        const uint8_t *src = in + t * TILE;
        __syncthreads();
        long long c0 = clock64();
        if (mode == 0 || mode == 1) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(TILE) : "memory");
                int piece = mode == 0 ? TILE : 4096;
                for (int o = 0; o < TILE; o += piece)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + o)), "l"(src + o), "r"(piece), "r"(smem_u32(&bar)) : "memory");
            }
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(it & 1) : "memory");
        } else if (mode == 2) {
            for (int k = 0; k < TILE / 16 / NT; k++) {
                int o = (k * NT + tid) * 16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + o)), "l"(src + o) : "memory");
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        } else {
            uint4 v[TILE / 16 / NT];
#pragma unroll
            for (int k = 0; k < TILE / 16 / NT; k++) v[k] = *reinterpret_cast<const uint4 *>(src + (k * NT + tid) * 16);
#pragma unroll
            for (int k = 0; k < TILE / 16 / NT; k++) *reinterpret_cast<uint4 *>(buf + (k * NT + tid) * 16) = v[k];
            __syncthreads();
        }
        long long c1 = clock64();
        acc += (unsigned long long)(c1 - c0);
        x += buf[(tid * 97 + it) % TILE];
    }
    if (tid == 0) atomicAdd(out, acc);
    if (x == 0xdeadbeef) *sink = x;
}
int main() {
    size_t bytes = 4ull << 30;
    uint8_t *d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    unsigned long long *out; cudaMalloc(&out, 8); unsigned *sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE);
    const char *names[4] = {"bulk 32K", "bulk 8x4K", "cp.async16", "ldg+sts"};
    for (int ctas = 1; ctas <= 4; ctas *= 2)
        for (int mode = 0; mode < 4; mode++) {
            int grid = 148 * ctas, iters = 64;
            cudaMemset(out, 0, 8);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            k<<<grid, NT, TILE>>>(d, bytes / TILE, mode, iters, out, sink);
            cudaEventRecord(b); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b);
            unsigned long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("ctas/SM %d  %-10s: avg %.0f cycles per 32 KiB load; aggregate %.1f GB/s  (%s)\n", ctas, names[mode], (double)h / grid / iters,
                   (double)grid * iters * TILE / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
