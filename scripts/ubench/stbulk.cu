// Micro-benchmark (development aid): shared-memory zero fill by st.bulk (UMEMSETS) vs st.shared.v4, and the ordering of
// st.bulk followed by cp.async to the same bytes issued by the same warp.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_bulk_zero(uint32_t a, uint64_t n) {
    asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(a), "l"(n) : "memory");
}

// mode 0: every thread st.shared.v4 zero (256 threads cover 16 KB in 4 stores each)
// mode 1: lane 0 of each of 8 warps: 4 x st.bulk 512 B
// mode 2: thread 0: one st.bulk 16 KB
// mode 3: lane 0 of each of 8 warps: 1 x st.bulk 2 KB
__global__ void fill_kernel(int mode, int iters, long long *cycles, int *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t tile = base + (it & 3) * 16384;
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(tile + (i * 256 + tid) * 16), "r"(0) : "memory");
        } else if (mode == 1) {
            if (lane == 0)
                for (int i = 0; i < 4; ++i) st_bulk_zero(tile + i * 4096 + warp * 512, 512);
        } else if (mode == 2) {
            if (tid == 0) st_bulk_zero(tile, 16384);
        } else {
            if (lane == 0) st_bulk_zero(tile + warp * 2048, 2048);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) cycles[mode] = t1 - t0;
    if (sink) sink[tid] = ((int *)smem)[tid];
}

// ordering: a warp zeroes 512 B with st.bulk, __syncwarp, then cp.async's a pattern over the same bytes; after the copies
// land the region must hold the pattern (never zeros).  Repeated many times with the region pre-dirtied.
__global__ void order_kernel(const uint4 *__restrict__ src, int iters, int *errors) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int bad = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t mine = base + warp * 512;
        // dirty
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(mine + lane * 16), "r"(0x7f7f7f7f) : "memory");
        __syncwarp();
        if (lane == 0) st_bulk_zero(mine, 512);
        __syncwarp();
        // only even lanes copy ("hits"); odd pieces must read back as zero, even pieces as the source pattern
        if (!(lane & 1))
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mine + lane * 16), "l"(src + (it * 32 + lane) % 4096) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(mine + lane * 16));
        uint4 want = (lane & 1) ? make_uint4(0, 0, 0, 0) : src[(it * 32 + lane) % 4096];
        bad += (v.x != want.x) | (v.y != want.y) | (v.z != want.z) | (v.w != want.w);
        __syncwarp();
    }
    if (bad) atomicAdd(errors, bad);
}

int main() {
    long long *cyc;
    int *sink, *err;
    uint4 *src;
    cudaMalloc(&cyc, 64);
    cudaMalloc(&sink, 4096);
    cudaMalloc(&err, 4);
    cudaMemset(err, 0, 4);
    cudaMalloc(&src, 4096 * 16);
    uint4 *h = (uint4 *)malloc(4096 * 16);
    for (int i = 0; i < 4096; ++i) h[i] = make_uint4(i * 4 + 1, i * 4 + 2, i * 4 + 3, i * 4 + 4);
    cudaMemcpy(src, h, 4096 * 16, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 2000;
    for (int mode = 0; mode < 4; ++mode) {
        fill_kernel<<<148, 256, 65536>>>(mode, iters, cyc, sink);
        cudaDeviceSynchronize();
    }
    long long hc[4];
    cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
    const char *names[4] = {"st.shared.v4 x4 / thread (256 thr)", "st.bulk 512 B x4 / warp (8 warps)", "st.bulk 16 KB x1 (1 thread)",
                            "st.bulk 2 KB x1 / warp (8 warps)"};
    for (int m = 0; m < 4; ++m) printf("%-40s %8.1f cycles per 16 KB (incl. __syncthreads)\n", names[m], (double)hc[m] / iters);
    order_kernel<<<148, 256, 65536>>>(src, 20000, err);
    cudaDeviceSynchronize();
    int he = 0;
    cudaMemcpy(&he, err, 4, cudaMemcpyDeviceToHost);
    printf("ordering st.bulk -> cp.async (same warp): %d mismatches, last error: %s\n", he, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
