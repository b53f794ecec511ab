// Occupancy index kernels + C ABI (see include/toda_b200.h, "Occupancy index").
#include <stdarg.h>

#include "common.cuh"
#include "scan.cuh"

// ------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";

void toda_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *toda_last_error(void) { return g_err; }
long long g_toda_launches = 0;
extern "C" long long toda_launch_count(void) { return g_toda_launches; }
extern "C" int toda_version(void) { return 100; }

extern "C" int toda_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    TODA_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    TODA_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return TODA_OK;
}

// ------------------------------------------------------------------ kernels
__global__ void index_insert_kernel(GridIndex g, const int4 *__restrict__ coords, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(coords + i);  // (b,z,y,x)
        if ((unsigned)c.x >= (unsigned)g.batch || (unsigned)c.y >= (unsigned)g.D || (unsigned)c.z >= (unsigned)g.H ||
            (unsigned)c.w >= (unsigned)g.W)
            continue;
        index_mark(g, index_cell(g, c.x, c.y, c.z, c.w));
    }
}

struct ConvGeom {
    int k[3], s[3], p[3];
};

// every input marks the outputs it reaches: o*s = p_in + pad - k, 0 <= o < out
__global__ void index_insert_strided_kernel(GridIndex g, const int4 *__restrict__ in_coords, int n_in, ConvGeom cg) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(in_coords + i);
        int oz[3], oy[3], ox[3];
        int nz = 0, ny = 0, nx = 0;
        for (int k = 0; k < cg.k[0]; ++k) {
            int t = c.y + cg.p[0] - k;
            if (t >= 0 && t % cg.s[0] == 0 && t / cg.s[0] < g.D) oz[nz++] = t / cg.s[0];
        }
        for (int k = 0; k < cg.k[1]; ++k) {
            int t = c.z + cg.p[1] - k;
            if (t >= 0 && t % cg.s[1] == 0 && t / cg.s[1] < g.H) oy[ny++] = t / cg.s[1];
        }
        for (int k = 0; k < cg.k[2]; ++k) {
            int t = c.w + cg.p[2] - k;
            if (t >= 0 && t % cg.s[2] == 0 && t / cg.s[2] < g.W) ox[nx++] = t / cg.s[2];
        }
        for (int a = 0; a < nz; ++a)
            for (int b = 0; b < ny; ++b)
                for (int d = 0; d < nx; ++d) index_mark(g, index_cell(g, c.x, oz[a], oy[b], ox[d]));
    }
}

// one warp per L1 word: cnt1[w1] = number of marked cells under it
__global__ void index_count_kernel(GridIndex g) {
    int warps_per_block = blockDim.x >> 5;
    int lane = threadIdx.x & 31;
    for (int w1 = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w1 < g.words1; w1 += gridDim.x * warps_per_block) {
        unsigned long long m = g.bits1[w1];
        int c = 0;
        if (m) {
            long long base = (long long)w1 * 64;
            if ((m >> lane) & 1ull) c += __popcll(g.bits0[base + lane]);
            if ((m >> (lane + 32)) & 1ull) c += __popcll(g.bits0[base + lane + 32]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) g.pref1[w1] = c;
    }
}

// one warp per L1 word: per-word ranks, sorted coordinate list / cell list
__global__ void index_emit_kernel(GridIndex g, int4 *__restrict__ out_coords, int cap, uint32_t *__restrict__ out_cells) {
    int warps_per_block = blockDim.x >> 5;
    int lane = threadIdx.x & 31;
    long long hw = (long long)g.H * g.W;
    for (int w1 = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w1 < g.words1; w1 += gridDim.x * warps_per_block) {
        unsigned long long m = g.bits1[w1];
        if (!m) continue;
        long long base = (long long)w1 * 64;
        unsigned long long b0 = ((m >> lane) & 1ull) ? g.bits0[base + lane] : 0ull;
        unsigned long long b1 = ((m >> (lane + 32)) & 1ull) ? g.bits0[base + lane + 32] : 0ull;
        int p0 = __popcll(b0), p1 = __popcll(b1);
        int i0 = p0, i1 = p1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t0 = __shfl_up_sync(0xffffffffu, i0, o);
            int t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += t0; i1 += t1; }
        }
        int tot0 = __shfl_sync(0xffffffffu, i0, 31);
        int start = g.pref1[w1];
        int r0 = start + i0 - p0;
        int r1 = start + tot0 + i1 - p1;
        if (b0) g.rank0[base + lane] = r0;
        if (b1) g.rank0[base + lane + 32] = r1;
        if (out_coords || out_cells) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                unsigned long long bits = half ? b1 : b0;
                int row = half ? r1 : r0;
                long long cell0 = (base + lane + 32 * half) * 64;
                while (bits) {
                    int bit = __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    long long cell = cell0 + bit;
                    if (row < cap) {
                        if (out_cells) out_cells[row] = (uint32_t)cell;
                        if (out_coords) {
                            int b = (int)(cell / g.frame_stride);
                            long long r = cell - (long long)b * g.frame_stride;
                            int z = (int)(r / hw);
                            int rem = (int)(r - (long long)z * hw);
                            out_coords[row] = make_int4(b, z, rem / g.W, rem % g.W);
                        }
                    }
                    ++row;
                }
            }
        }
    }
}

// n_out[0] = total, n_out[1+b] = marked cells of frame b (frame boundaries are L1-word aligned)
__global__ void index_counts_kernel(GridIndex g, int32_t *n_out) {
    int words1_per_frame = (int)(g.frame_stride / 4096);
    for (int b = threadIdx.x; b <= g.batch; b += blockDim.x) {
        if (b == 0) n_out[0] = g.pref1[g.words1];
        if (b < g.batch) n_out[1 + b] = g.pref1[(b + 1) * words1_per_frame] - g.pref1[b * words1_per_frame];
    }
}

__global__ void index_rows_kernel(GridIndex g, const int4 *__restrict__ coords, int n, int *__restrict__ rows) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(coords + i);
        int r = -1;
        if ((unsigned)c.x < (unsigned)g.batch && (unsigned)c.y < (unsigned)g.D && (unsigned)c.z < (unsigned)g.H &&
            (unsigned)c.w < (unsigned)g.W)
            r = index_lookup(g, index_cell(g, c.x, c.y, c.z, c.w));
        rows[i] = r;
    }
}

__global__ void index_release_kernel(GridIndex g, const int4 *__restrict__ coords, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(coords + i);
        if ((unsigned)c.x >= (unsigned)g.batch || (unsigned)c.y >= (unsigned)g.D || (unsigned)c.z >= (unsigned)g.H ||
            (unsigned)c.w >= (unsigned)g.W)
            continue;
        long long w = index_cell(g, c.x, c.y, c.z, c.w) >> 6;
        g.bits0[w] = 0ull;
        g.bits1[w >> 6] = 0ull;
    }
}

// ------------------------------------------------------------------ shared build step
int index_build_ranks(const GridIndex &g, int32_t *out_coords, int cap, uint32_t *out_cells, int32_t *n_out,
                      cudaStream_t st) {
    const int block = 256;
    int grid = wave_grid((int64_t)g.words1 * 32, block);
    index_count_kernel<<<grid, block, 0, st>>>(g);
    TODA_LAUNCH_OK();
    int rc = scan_exclusive(LoadInt{g.pref1}, g.words1, g.pref1, g.tile_sums, nullptr, st);
    if (rc != TODA_OK) return rc;
    index_emit_kernel<<<grid, block, 0, st>>>(g, (int4 *)out_coords, cap, out_cells);
    TODA_LAUNCH_OK();
    if (n_out) {
        index_counts_kernel<<<1, 64, 0, st>>>(g, n_out);
        TODA_LAUNCH_OK();
    }
    return TODA_OK;
}

// ------------------------------------------------------------------ C ABI
static int check_dims(int batch, int D, int H, int W) {
    TODA_CHECK_ARG(batch > 0 && D > 0 && H > 0 && W > 0, "index: bad dims batch=%d D=%d H=%d W=%d", batch, D, H, W);
    long long cells = index_frame_stride(D, H, W) * batch;
    TODA_CHECK_ARG(cells < (1ll << 32) - 1, "index: %lld cells exceed the 32-bit cell id range", cells);
    return TODA_OK;
}

extern "C" size_t toda_index_bytes(int batch, int D, int H, int W) {
    if (batch <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return index_layout(nullptr, nullptr, batch, D, H, W);
}

extern "C" int toda_index_insert(void *index, int batch, int D, int H, int W, const int32_t *coords, int n,
                                 void *stream) {
    if (int rc = check_dims(batch, D, H, W)) return rc;
    TODA_CHECK_ARG(index && (coords || n == 0) && n >= 0, "index_insert: null pointer or negative n");
    if (n == 0) return TODA_OK;
    GridIndex g;
    index_layout(&g, index, batch, D, H, W);
    index_insert_kernel<<<wave_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(g, (const int4 *)coords, n);
    TODA_LAUNCH_OK();
    return TODA_OK;
}

extern "C" int toda_index_insert_strided(void *index_out, int batch, int oD, int oH, int oW, const int32_t *in_coords,
                                         int n_in, const int *k, const int *s, const int *p, void *stream) {
    if (int rc = check_dims(batch, oD, oH, oW)) return rc;
    TODA_CHECK_ARG(index_out && (in_coords || n_in == 0) && n_in >= 0 && k && s && p, "index_insert_strided: bad args");
    ConvGeom cg;
    for (int a = 0; a < 3; ++a) {
        TODA_CHECK_ARG(k[a] >= 1 && k[a] <= 3 && s[a] >= 1 && p[a] >= 0, "index_insert_strided: unsupported k/s/p");
        cg.k[a] = k[a]; cg.s[a] = s[a]; cg.p[a] = p[a];
    }
    if (n_in == 0) return TODA_OK;
    GridIndex g;
    index_layout(&g, index_out, batch, oD, oH, oW);
    index_insert_strided_kernel<<<wave_grid(n_in, 256), 256, 0, (cudaStream_t)stream>>>(g, (const int4 *)in_coords,
                                                                                        n_in, cg);
    TODA_LAUNCH_OK();
    return TODA_OK;
}

extern "C" int toda_index_build(void *index, int batch, int D, int H, int W, int32_t *out_coords, int cap,
                                int32_t *n_out, void *stream) {
    if (int rc = check_dims(batch, D, H, W)) return rc;
    TODA_CHECK_ARG(index && cap >= 0, "index_build: bad args");
    GridIndex g;
    index_layout(&g, index, batch, D, H, W);
    return index_build_ranks(g, out_coords, cap, nullptr, n_out, (cudaStream_t)stream);
}

extern "C" int toda_index_rows(const void *index, int batch, int D, int H, int W, const int32_t *coords, int n,
                               int32_t *rows, void *stream) {
    if (int rc = check_dims(batch, D, H, W)) return rc;
    TODA_CHECK_ARG(index && n >= 0 && (n == 0 || (coords && rows)), "index_rows: bad args");
    if (n == 0) return TODA_OK;
    GridIndex g;
    index_layout(&g, (void *)index, batch, D, H, W);
    index_rows_kernel<<<wave_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(g, (const int4 *)coords, n, rows);
    TODA_LAUNCH_OK();
    return TODA_OK;
}

extern "C" int toda_index_release(void *index, int batch, int D, int H, int W, const int32_t *coords, int n,
                                  void *stream) {
    if (int rc = check_dims(batch, D, H, W)) return rc;
    TODA_CHECK_ARG(index && n >= 0 && (n == 0 || coords), "index_release: bad args");
    if (n == 0) return TODA_OK;
    GridIndex g;
    index_layout(&g, index, batch, D, H, W);
    index_release_kernel<<<wave_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(g, (const int4 *)coords, n);
    TODA_LAUNCH_OK();
    return TODA_OK;
}
