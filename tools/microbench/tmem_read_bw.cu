// Tensor-memory read bandwidth of one SM: W warps (1, 2 or 4 warpgroups' worth) read 32-lane x 32-column fp32 blocks with
// tcgen05.ld.32x32b.x32 back to back (DEPTH loads in flight before each wait) for ITERS rounds; bytes / clock64 cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_read_bw tools/microbench/tmem_read_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
        "%27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int DEPTH, int X16>
__global__ void __launch_bounds__(512, 1) bw_kernel(int warps, int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s + ((uint32_t)((warp % 4) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        for (int it = 0; it < iters; ++it) {
            if (X16) {
                float v[DEPTH][16];
#pragma unroll
                for (int k = 0; k < DEPTH; ++k) ld16(base + (uint32_t)(((it * DEPTH + k) * 16) % 512), v[k]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int k = 0; k < DEPTH; ++k) acc += v[k][0] + v[k][15];
            } else {
                float v[DEPTH][32];
#pragma unroll
                for (int k = 0; k < DEPTH; ++k) ld32(base + (uint32_t)(((it * DEPTH + k) * 32) % 512), v[k]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int k = 0; k < DEPTH; ++k) acc += v[k][0] + v[k][31];
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 12345.678f) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512));
}

template <int DEPTH, int X16>
void run(int warps, long long* d_cycles, float* d_sink) {
    const int iters = 2000;
    bw_kernel<DEPTH, X16><<<1, 512>>>(warps, iters, d_cycles, d_sink);
    bw_kernel<DEPTH, X16><<<1, 512>>>(warps, iters, d_cycles, d_sink);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * iters * DEPTH * 32 * (X16 ? 16 : 32) * 4;
    printf("x%-2d depth %d warps %2d: %9lld cycles, %7.1f B/clk/SM, %6.1f cycles per load per warp  (%s)\n", X16 ? 16 : 32, DEPTH, warps, c,
           bytes / (double)c, (double)c / (iters * DEPTH), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* d_cycles; float* d_sink;
    cudaMalloc(&d_cycles, 64); cudaMalloc(&d_sink, 64);
    for (int warps : {1, 4, 8, 16}) {
        run<1, 0>(warps, d_cycles, d_sink);
        run<2, 0>(warps, d_cycles, d_sink);
        run<1, 1>(warps, d_cycles, d_sink);
        run<4, 1>(warps, d_cycles, d_sink);
    }
    return 0;
}
