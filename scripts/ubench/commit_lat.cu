// Micro-benchmark (development aid): latency of tcgen05.commit -> mbarrier phase completion, with and without MMAs in front.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
    uint64_t d = 0; d |= (uint64_t)((a >> 4) & 0x3fff); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
}
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long *out, int nmma, int ts) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *sm = raw + (base - smem_u32(raw));
    const uint32_t bar = base + 65536, slot = bar + 64;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 128) ((uint32_t *)sm)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8 * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)(sm + (slot - base));
    if (tid == 0 && ts < 2) {
        const uint64_t ad = desc_k_sw128(base), bd = desc_k_sw128(base + 32768);
        for (int rep = 0; rep < 4; ++rep) {
            long long t0 = clock64();
            for (int j = 0; j < nmma; ++j) {
                if (ts)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 256 + 8 * (j & 3)), "l"(bd + 2 * (j & 3)), "r"(idesc(128, N)), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * (j & 3)), "l"(bd + 2 * (j & 3)), "r"(idesc(128, N)), "r"(1u) : "memory");
            }
            long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * rep) : "memory");
            long long t2 = clock64();
            wait(bar + 8 * rep, 0);
            long long t3 = clock64();
            out[rep * 3] = t1 - t0; out[rep * 3 + 1] = t2 - t1; out[rep * 3 + 2] = t3 - t2;
        }
    }
    __syncthreads();
    if (warp == 1 && ts == 2) {
        // warp-uniform control flow, one elected lane, fully unrolled groups of 8 MMAs (A in TMEM)
        const uint64_t bd = desc_k_sw128(base + 32768);
        const uint32_t a0 = tmem + 256;
        for (int rep = 0; rep < 4; ++rep) {
            long long t0 = clock64();
            for (int it = 0; it < nmma / 8; ++it) {
                uint32_t pred;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                if (pred) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a0 + 8 * (j & 3) + 32 * (j >> 2)), "l"(bd + 2 * (j & 3)), "r"(idesc(128, N)), "r"(1u) : "memory");
                }
                __syncwarp();
            }
            long long t1 = clock64();
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * rep) : "memory");
            __syncwarp();
            long long t2 = clock64();
            wait(bar + 8 * rep, 0);
            long long t3 = clock64();
            if ((tid & 31) == 0) { out[rep * 3] = t1 - t0; out[rep * 3 + 1] = t2 - t1; out[rep * 3 + 2] = t3 - t2; }
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int N> void run(int nmma, int ts) {
    long long *d, h[12];
    cudaMalloc(&d, 96);
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    k<N><<<1, 128, 70 * 1024>>>(d, nmma, ts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("fail %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost);
    printf("N=%3d %s nmma=%2d: ", N, ts == 2 ? "TS-uniform" : (ts ? "TS" : "SS"), nmma);
    for (int r = 0; r < 4; ++r) printf("[issue %lld commit %lld wait %lld] ", h[r * 3], h[r * 3 + 1], h[r * 3 + 2]);
    printf("\n");
    cudaFree(d);
}
int main() {
    run<32>(8, 2); run<32>(32, 2); run<64>(32, 2); run<128>(32, 2); run<128>(64, 2);
    for (int ts = 0; ts < 2; ++ts) {
        run<32>(0, ts); run<32>(4, ts); run<32>(8, ts); run<32>(32, ts);
        run<64>(8, ts); run<128>(8, ts); run<128>(32, ts);
    }
    return 0;
}
