// Micro-benchmark (development aid, not product): throughput of tcgen05.st 32x32b.x32 (the A-operand write of conv_ts.cu)
// with 1, 2 and 4 warpgroups storing concurrently, with and without a tcgen05.wait::st per store.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sttm_bw sttm_bw.cu && ./sttm_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define STTM_X32(addr, v)                                                                                                           \
    asm volatile(                                                                                                                   \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"     \
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                                                                         \
        ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),  \
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),   \
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),   \
          "r"(v[30]), "r"(v[31])                                                                                                    \
        : "memory")

__global__ void __launch_bounds__(512, 1) k(int nwarps, int wait_each, int iters, int lds, long long *out) {
    __shared__ uint32_t slot;
    __shared__ __align__(16) uint32_t buf[8192];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = i;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 32 + i;
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            if (lds) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(buf) + (((threadIdx.x * 4 + i) & 511) << 6);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[4 * q]), "=r"(v[4 * q + 1]), "=r"(v[4 * q + 2]), "=r"(v[4 * q + 3]) : "r"(sa ^ ((q & 3) << 4)));
            }
            STTM_X32(addr + (i & 1) * 32, v);
            if (wait_each) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (lane == 0 && warp < nwarps) atomicMax((unsigned long long *)out + 1, (unsigned long long)(t1 - t0));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int iters = 2000;
    for (int lds = 0; lds < 2; ++lds)
        for (int wait_each = 0; wait_each < 2; ++wait_each)
            for (int nw : {4, 8, 16}) {
                cudaMemset(d, 0, 16);
                k<<<1, 512>>>(nw, wait_each, iters, lds, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                // one x32 store of a warpgroup = 128 lanes x 32 columns x 4 B = 16 KB
                printf("lds %d wait_each %d warps %2d: %s  %.1f cycles per x32 store per warp, %.1f B/clk into TMEM\n", lds, wait_each, nw,
                       cudaGetErrorString(e), (double)h[1] / iters, (double)nw * 32 * 32 * 4 * iters / (double)h[1]);
            }
    return 0;
}
