// Micro-test (development aid, not product): validates the operand conventions the conv kernels rely on.
//   case 0: kind::f16 (bf16), A from TMEM (tcgen05.st 32x32b, lane = row, 2 bf16 per 32-bit column), B K-major SW128 smem
//   case 1: kind::tf32, A and B K-major SW128 in shared memory (32 fp32 per 128-byte row, UMMA K = 8)
//   case 2: kind::tf32, A from TMEM (one fp32 per column), B K-major SW128 smem
//   case 3: kind::tf32, A and B MN-major SW128 in shared memory (wgrad form: tiles [k rows][32 fp32 of M/N])
// D[128 x N] = A[128 x K] * B[N x K]^T, K = 64 (bf16) or 32 (tf32) = one 128-byte swizzle row.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_mma ts_mma.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t a, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int fmt, int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int N>
__global__ void __launch_bounds__(128, 1) k(int mode, const void *A, const void *B, float *D) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *sm = raw + (base - smem_u32(raw));
    const uint32_t a_s = base, b_s = base + 32768, bar = base + 65536, slot = bar + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *(volatile uint32_t *)(sm + (slot - base));
    const uint32_t acol = 256;   // A operand columns (TS cases); accumulator at column 0
    const int row = tid;         // 128 threads = 128 rows
    // ---- B operand (all cases except 3: K-major [N][128 B]) ----
    if (mode != 3) {
        for (int e = tid; e < N * 8; e += 128) {
            int r = e >> 3, p = e & 7;
            uint4 v = *((const uint4 *)B + (size_t)r * 8 + p);
            *(uint4 *)(sm + 32768 + r * 128 + ((p ^ (r & 7)) << 4)) = v;
        }
    }
    if (mode == 0 || mode == 2) {
        // A row (128 bytes) -> 32 registers -> TMEM columns acol .. acol+31 of lane `row`
        uint32_t v[32];
        const uint4 *src = (const uint4 *)A + (size_t)row * 8;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            uint4 q = src[p];
            v[4 * p] = q.x; v[4 * p + 1] = q.y; v[4 * p + 2] = q.z; v[4 * p + 3] = q.w;
        }
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + acol;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
            ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
              "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
              "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
              "r"(v[30]), "r"(v[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (mode == 1) {
        for (int e = tid; e < 128 * 8; e += 128) {
            int r = e >> 3, p = e & 7;
            uint4 v = *((const uint4 *)A + (size_t)r * 8 + p);
            *(uint4 *)(sm + r * 128 + ((p ^ (r & 7)) << 4)) = v;
        }
    } else {
        // mode 3: MN-major tiles.  A given as [K=32 rows][M=128 fp32] row-major, B as [K=32][N fp32] row-major.
        // smem: per 32-wide (128-byte) M block: [32 k-rows][128 B] SW128, blocks `32*128` bytes apart.
        for (int e = tid; e < 32 * 32; e += 128) {       // A: 32 k-rows x 4 blocks x 8 pieces
            int kr = e >> 5, blk = (e >> 3) & 3, p = e & 7;
            uint4 v = *((const uint4 *)A + (size_t)kr * 32 + blk * 8 + p);
            *(uint4 *)(sm + blk * 4096 + kr * 128 + ((p ^ (kr & 7)) << 4)) = v;
        }
        for (int e = tid; e < 32 * (N / 4); e += 128) {
            int kr = e / (N / 4), rest = e % (N / 4), blk = rest >> 3, p = rest & 7;
            uint4 v = *((const uint4 *)B + (size_t)kr * (N / 4) + blk * 8 + p);
            *(uint4 *)(sm + 32768 + blk * 4096 + kr * 128 + ((p ^ (kr & 7)) << 4)) = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t bd = desc_k_sw128(b_s), ad = desc_k_sw128(a_s);
        if (mode == 0) {
            const uint32_t id = idesc(1, 128, N, 0, 0);
            for (int j = 0; j < 4; ++j) {     // K = 16 bf16 per MMA = 8 TMEM columns of A, 32 bytes of B
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + acol + 8 * j),
                             "l"(bd + 2 * j), "r"(id), "r"((uint32_t)(j > 0)) : "memory");
            }
        } else if (mode == 1) {
            const uint32_t id = idesc(2, 128, N, 0, 0);
            for (int j = 0; j < 4; ++j) {     // K = 8 tf32 per MMA = 32 bytes
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * j), "l"(bd + 2 * j),
                             "r"(id), "r"((uint32_t)(j > 0)) : "memory");
            }
        } else if (mode == 2) {
            const uint32_t id = idesc(2, 128, N, 0, 0);
            for (int j = 0; j < 4; ++j) {     // K = 8 tf32 per MMA = 8 TMEM columns of A
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + acol + 8 * j),
                             "l"(bd + 2 * j), "r"(id), "r"((uint32_t)(j > 0)) : "memory");
            }
        } else {
            const uint32_t id = idesc(2, 128, N, 1, 1);
            const uint64_t am = desc_mn_sw128(a_s, 4096), bm = desc_mn_sw128(b_s, 4096);
            for (int j = 0; j < 4; ++j) {     // K = 8 rows = one 1024-byte swizzle atom: start address field + 64
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(am + 64 * j), "l"(bm + 64 * j),
                             "r"(id), "r"((uint32_t)(j > 0)) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    {
        uint32_t done = 0, spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0) : "memory");
            if (!done && ++spins > (1u << 22)) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int n0 = 0; n0 < N; n0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + n0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + n0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float tf(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

template <int N> static int run(int mode) {
    const int K = mode == 0 ? 64 : 32;
    float *Ah = (float *)malloc(128 * K * 4), *Bh = (float *)malloc(N * K * 4), *Dh = (float *)malloc(128 * N * 4);
    srand(7 + mode);
    for (int i = 0; i < 128 * K; ++i) Ah[i] = (rand() % 2001 - 1000) / 1000.f;
    for (int i = 0; i < N * K; ++i) Bh[i] = (rand() % 2001 - 1000) / 1000.f;
    void *Ad, *Bd; float *Dd;
    CK(cudaMalloc(&Ad, 128 * K * 4)); CK(cudaMalloc(&Bd, N * K * 4)); CK(cudaMalloc(&Dd, 128 * N * 4));
    if (mode == 0) {
        __nv_bfloat16 *t = (__nv_bfloat16 *)malloc(128 * K * 2);
        for (int i = 0; i < 128 * K; ++i) { t[i] = __float2bfloat16(Ah[i]); Ah[i] = bf(Ah[i]); }
        CK(cudaMemcpy(Ad, t, 128 * K * 2, cudaMemcpyHostToDevice));
        for (int i = 0; i < N * K; ++i) { t[i] = __float2bfloat16(Bh[i]); Bh[i] = bf(Bh[i]); }
        CK(cudaMemcpy(Bd, t, N * K * 2, cudaMemcpyHostToDevice));
        free(t);
    } else if (mode == 3) {
        // device layouts: A^T [K][128], B^T [K][N]
        float *t = (float *)malloc(128 * K * 4);
        for (int m = 0; m < 128; ++m) for (int kk = 0; kk < K; ++kk) t[kk * 128 + m] = Ah[m * K + kk];
        CK(cudaMemcpy(Ad, t, 128 * K * 4, cudaMemcpyHostToDevice));
        for (int n = 0; n < N; ++n) for (int kk = 0; kk < K; ++kk) t[kk * N + n] = Bh[n * K + kk];
        CK(cudaMemcpy(Bd, t, N * K * 4, cudaMemcpyHostToDevice));
        free(t);
    } else {
        CK(cudaMemcpy(Ad, Ah, 128 * K * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(Bd, Bh, N * K * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024));
    k<N><<<1, 128, 68 * 1024>>>(mode, Ad, Bd, Dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d N %d: kernel failed: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(Dh, Dd, 128 * N * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_err_tf = 0, ref_max = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0, st = 0;
            for (int kk = 0; kk < K; ++kk) { s += (double)Ah[m * K + kk] * Bh[n * K + kk]; st += (double)tf(Ah[m * K + kk]) * tf(Bh[n * K + kk]); }
            max_err = fmax(max_err, fabs(s - Dh[m * N + n]));
            max_err_tf = fmax(max_err_tf, fabs(st - Dh[m * N + n]));
            ref_max = fmax(ref_max, fabs(s));
        }
    printf("mode %d N %d: max|err| vs exact-operand ref %.3e, vs tf32-truncated ref %.3e (max|ref| %.3f)  %s\n", mode, N, max_err, max_err_tf,
           ref_max, (mode == 0 ? max_err < 1e-4 : max_err < 2e-2) ? "OK" : "MISMATCH");
    cudaFree(Ad); cudaFree(Bd); cudaFree(Dd); free(Ah); free(Bh); free(Dh);
    return 0;
}

int main() {
    for (int mode = 0; mode < 4; ++mode) { run<64>(mode); run<128>(mode); }
    return 0;
}
