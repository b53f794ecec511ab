// K3 / K4 rulebook builders: neighbour tables from the occupancy index (see include/toda_b200.h).
#include "common.cuh"

namespace {

struct Geom3 {
    int k[3], s[3], p[3];
};

// SubMConv3d: input at coord(o) + (k - K/2).  One thread per output row, all kernel offsets:
// the coordinate is read once (one 16-byte load) and the 27 probes are independent loads in flight.
// The three x-offsets of a (dz,dy) pair hit the same or adjacent 64-cell words, so they share sectors.
template <int KZ, int KY, int KX>
__global__ void __launch_bounds__(256) rulebook_subm_kernel(GridIndex g, const int4 *__restrict__ coords, int n,
                                                            int *__restrict__ nbr) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        int4 c = __ldg(coords + o);  // b,z,y,x
#pragma unroll
        for (int kz = 0; kz < KZ; ++kz) {
            int z = c.y + kz - KZ / 2;
#pragma unroll
            for (int ky = 0; ky < KY; ++ky) {
                int y = c.z + ky - KY / 2;
                bool row_ok = (unsigned)z < (unsigned)g.D && (unsigned)y < (unsigned)g.H;
                long long base = index_cell(g, c.x, z, y, 0);
#pragma unroll
                for (int kx = 0; kx < KX; ++kx) {
                    int x = c.w + kx - KX / 2;
                    int r = -1;
                    if (row_ok && (unsigned)x < (unsigned)g.W) r = index_lookup(g, base + x);
                    nbr[(size_t)((kz * KY + ky) * KX + kx) * n + o] = r;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) rulebook_subm_generic_kernel(GridIndex g, const int4 *__restrict__ coords, int n,
                                                                    int KZ, int KY, int KX, int *__restrict__ nbr) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        int4 c = __ldg(coords + o);
        for (int kz = 0; kz < KZ; ++kz)
            for (int ky = 0; ky < KY; ++ky)
                for (int kx = 0; kx < KX; ++kx) {
                    int z = c.y + kz - KZ / 2, y = c.z + ky - KY / 2, x = c.w + kx - KX / 2;
                    int r = -1;
                    if ((unsigned)z < (unsigned)g.D && (unsigned)y < (unsigned)g.H && (unsigned)x < (unsigned)g.W)
                        r = index_lookup(g, index_cell(g, c.x, z, y, x));
                    nbr[(size_t)((kz * KY + ky) * KX + kx) * n + o] = r;
                }
    }
}

// SparseConv3d forward table: output o gathers input at o*s - pad + k.
__global__ void __launch_bounds__(256) rulebook_sparse_fwd_kernel(GridIndex gin, const int4 *__restrict__ out_coords,
                                                                  int n_out, Geom3 cg, int *__restrict__ nbr) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += gridDim.x * blockDim.x) {
        int4 c = __ldg(out_coords + o);
        int k = 0;
        for (int kz = 0; kz < cg.k[0]; ++kz)
            for (int ky = 0; ky < cg.k[1]; ++ky)
                for (int kx = 0; kx < cg.k[2]; ++kx, ++k) {
                    int z = c.y * cg.s[0] - cg.p[0] + kz;
                    int y = c.z * cg.s[1] - cg.p[1] + ky;
                    int x = c.w * cg.s[2] - cg.p[2] + kx;
                    int r = -1;
                    if ((unsigned)z < (unsigned)gin.D && (unsigned)y < (unsigned)gin.H && (unsigned)x < (unsigned)gin.W)
                        r = index_lookup(gin, index_cell(gin, c.x, z, y, x));
                    nbr[(size_t)k * n_out + o] = r;
                }
    }
}

// SparseConv3d input-stationary table: input i feeds output (p_in + pad - k)/s under offset k.
__global__ void __launch_bounds__(256) rulebook_sparse_bwd_kernel(GridIndex gout, const int4 *__restrict__ in_coords,
                                                                  int n_in, Geom3 cg, int *__restrict__ nbr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
        int4 c = __ldg(in_coords + i);
        int k = 0;
        for (int kz = 0; kz < cg.k[0]; ++kz)
            for (int ky = 0; ky < cg.k[1]; ++ky)
                for (int kx = 0; kx < cg.k[2]; ++kx, ++k) {
                    int tz = c.y + cg.p[0] - kz, ty = c.z + cg.p[1] - ky, tx = c.w + cg.p[2] - kx;
                    int r = -1;
                    if (tz >= 0 && ty >= 0 && tx >= 0 && tz % cg.s[0] == 0 && ty % cg.s[1] == 0 && tx % cg.s[2] == 0) {
                        int z = tz / cg.s[0], y = ty / cg.s[1], x = tx / cg.s[2];
                        if (z < gout.D && y < gout.H && x < gout.W) r = index_lookup(gout, index_cell(gout, c.x, z, y, x));
                    }
                    nbr[(size_t)k * n_in + i] = r;
                }
    }
}

}  // namespace

extern "C" int toda_rulebook_subm(const void *index, int batch, int D, int H, int W, const int32_t *coords, int n,
                                  const int *k, int32_t *nbr, void *stream) {
    TODA_CHECK_ARG(index && k && n >= 0 && batch > 0 && D > 0 && H > 0 && W > 0, "rulebook_subm: bad args");
    TODA_CHECK_ARG(k[0] >= 1 && k[1] >= 1 && k[2] >= 1 && (k[0] & 1) && (k[1] & 1) && (k[2] & 1),
                   "rulebook_subm: kernel sizes must be odd");
    if (n == 0) return TODA_OK;
    TODA_CHECK_ARG(coords && nbr, "rulebook_subm: null pointer");
    GridIndex g;
    index_layout(&g, (void *)index, batch, D, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    int grid = wave_grid(n, 256);
    if (k[0] == 3 && k[1] == 3 && k[2] == 3)
        rulebook_subm_kernel<3, 3, 3><<<grid, 256, 0, st>>>(g, (const int4 *)coords, n, nbr);
    else
        rulebook_subm_generic_kernel<<<grid, 256, 0, st>>>(g, (const int4 *)coords, n, k[0], k[1], k[2], nbr);
    TODA_LAUNCH_OK();
    return TODA_OK;
}

extern "C" int toda_rulebook_sparse(const void *index_in, int iD, int iH, int iW, const void *index_out, int oD, int oH,
                                    int oW, int batch, const int32_t *in_coords, int n_in, const int32_t *out_coords,
                                    int n_out, const int *k, const int *s, const int *p, int32_t *nbr_fwd,
                                    int32_t *nbr_bwd, void *stream) {
    TODA_CHECK_ARG(index_in && index_out && k && s && p && n_in >= 0 && n_out >= 0 && batch > 0, "rulebook_sparse: bad args");
    Geom3 cg;
    for (int a = 0; a < 3; ++a) {
        TODA_CHECK_ARG(k[a] >= 1 && s[a] >= 1 && p[a] >= 0, "rulebook_sparse: bad k/s/p");
        cg.k[a] = k[a]; cg.s[a] = s[a]; cg.p[a] = p[a];
    }
    cudaStream_t st = (cudaStream_t)stream;
    GridIndex gin, gout;
    index_layout(&gin, (void *)index_in, batch, iD, iH, iW);
    index_layout(&gout, (void *)index_out, batch, oD, oH, oW);
    if (n_out > 0 && nbr_fwd) {
        TODA_CHECK_ARG(out_coords, "rulebook_sparse: null out_coords");
        rulebook_sparse_fwd_kernel<<<wave_grid(n_out, 256), 256, 0, st>>>(gin, (const int4 *)out_coords, n_out, cg, nbr_fwd);
        TODA_LAUNCH_OK();
    }
    if (n_in > 0 && nbr_bwd) {
        TODA_CHECK_ARG(in_coords, "rulebook_sparse: null in_coords");
        rulebook_sparse_bwd_kernel<<<wave_grid(n_in, 256), 256, 0, st>>>(gout, (const int4 *)in_coords, n_in, cg, nbr_bwd);
        TODA_LAUNCH_OK();
    }
    return TODA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Parity order of the input rows of a strided convolution (used by its dgrad, see toda_spconv_fwd's out_rows):
// class(row) = (((z+pz) % sz) * sy + (y+py) % sy) * sx + (x+px) % sx; rows sorted by class, canonical order kept inside a
// class (stable counting sort, deterministic).  Two launches over 1024-row tiles: per-tile class histograms, then
// class bases (sum over all tiles, tiny) + per-row ranks by warp ballots.
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kPoThreads = 256, kPoTile = 1024, kPoMaxClasses = 8;

struct ParityGeom { int s[3], p[3]; };

__device__ __forceinline__ int parity_class(const int32_t *__restrict__ coords, int row, const ParityGeom &g) {
    const int4 c = __ldg((const int4 *)coords + row);          // (b, z, y, x)
    return (((c.y + g.p[0]) % g.s[0]) * g.s[1] + (c.z + g.p[1]) % g.s[1]) * g.s[2] + (c.w + g.p[2]) % g.s[2];
}

__global__ void __launch_bounds__(kPoThreads) parity_count_kernel(const int32_t *__restrict__ coords, int n, ParityGeom g,
                                                                  int nclasses, int *__restrict__ tile_counts) {
    __shared__ int cnt[kPoMaxClasses];
    if (threadIdx.x < kPoMaxClasses) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int tile0 = blockIdx.x * kPoTile, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < kPoTile / kPoThreads; ++j) {
        const int row = tile0 + j * kPoThreads + threadIdx.x;
        const int c = row < n ? parity_class(coords, row, g) : -1;
        for (int q = 0; q < nclasses; ++q) {
            const uint32_t b = __ballot_sync(0xffffffffu, c == q);
            if (lane == 0 && b) atomicAdd(&cnt[q], __popc(b));      // shared-memory integer adds: order-independent
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nclasses) tile_counts[blockIdx.x * kPoMaxClasses + threadIdx.x] = cnt[threadIdx.x];
}

__global__ void __launch_bounds__(kPoThreads) parity_scatter_kernel(const int32_t *__restrict__ coords, int n, ParityGeom g,
                                                                    int nclasses, const int *__restrict__ tile_counts,
                                                                    int num_tiles, int32_t *__restrict__ order,
                                                                    int32_t *__restrict__ pos_of_row) {
    __shared__ int before[kPoMaxClasses], total[kPoMaxClasses], base[kPoMaxClasses];
    __shared__ int slice_cnt[kPoTile / 32][kPoMaxClasses];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per class: rows in earlier tiles, rows in all tiles (one warp per class, lanes stride over the tiles)
    if (warp < nclasses) {
        int b = 0, t = 0;
        for (int i = lane; i < num_tiles; i += 32) {
            const int v = tile_counts[i * kPoMaxClasses + warp];
            t += v;
            if (i < (int)blockIdx.x) b += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
        if (lane == 0) { before[warp] = b; total[warp] = t; }
    }
    const int tile0 = blockIdx.x * kPoTile;
    int cls[kPoTile / kPoThreads], rank[kPoTile / kPoThreads];
#pragma unroll
    for (int j = 0; j < kPoTile / kPoThreads; ++j) {
        const int row = tile0 + j * kPoThreads + threadIdx.x;
        const int c = row < n ? parity_class(coords, row, g) : -1;
        cls[j] = c;
        rank[j] = 0;
        const int slice = j * (kPoThreads / 32) + warp;          // slices are in row order
        for (int q = 0; q < nclasses; ++q) {
            const uint32_t b = __ballot_sync(0xffffffffu, c == q);
            if (c == q) rank[j] = __popc(b & ((1u << lane) - 1u));
            if (lane == 0) slice_cnt[slice][q] = __popc(b);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nclasses) {
        const int q = threadIdx.x;
        int start = 0;
        for (int c = 0; c < q; ++c) start += total[c];
        base[q] = start + before[q];
        int run = 0;                                              // exclusive prefix over the tile's 32 slices
        for (int s = 0; s < kPoTile / 32; ++s) { const int v = slice_cnt[s][q]; slice_cnt[s][q] = run; run += v; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPoTile / kPoThreads; ++j) {
        const int row = tile0 + j * kPoThreads + threadIdx.x;
        if (row >= n) continue;
        const int slice = j * (kPoThreads / 32) + warp;
        const int pos = base[cls[j]] + slice_cnt[slice][cls[j]] + rank[j];
        order[pos] = row;
        if (pos_of_row) pos_of_row[row] = pos;
    }
}

}  // namespace

extern "C" size_t toda_parity_order_workspace_bytes(int n) {
    return n < 0 ? 0 : (size_t)ceil_div(n > 0 ? n : 1, kPoTile) * kPoMaxClasses * sizeof(int) + 256;
}

extern "C" int toda_parity_order(const int32_t *coords, int n, const int *stride_host, const int *pad_host, int32_t *order,
                                 int32_t *pos_of_row, void *workspace, size_t workspace_bytes, void *stream) {
    TODA_CHECK_ARG(n >= 0 && stride_host && pad_host, "parity_order: bad arguments");
    ParityGeom g;
    int nclasses = 1;
    for (int a = 0; a < 3; ++a) {
        g.s[a] = stride_host[a];
        g.p[a] = pad_host[a];
        TODA_CHECK_ARG(g.s[a] >= 1 && g.p[a] >= 0, "parity_order: bad stride / padding");
        nclasses *= g.s[a];
    }
    TODA_CHECK_ARG(nclasses <= kPoMaxClasses, "parity_order: %d parity classes > %d", nclasses, kPoMaxClasses);
    if (n == 0) return TODA_OK;
    TODA_CHECK_ARG(coords && order && workspace, "parity_order: null pointer");
    if (workspace_bytes < toda_parity_order_workspace_bytes(n)) { toda_set_error("parity_order: workspace too small"); return TODA_ERR_WORKSPACE; }
    const int tiles = ceil_div(n, kPoTile);
    cudaStream_t st = (cudaStream_t)stream;
    parity_count_kernel<<<tiles, kPoThreads, 0, st>>>(coords, n, g, nclasses, (int *)workspace);
    TODA_LAUNCH_OK();
    parity_scatter_kernel<<<tiles, kPoThreads, 0, st>>>(coords, n, g, nclasses, (const int *)workspace, tiles, order, pos_of_row);
    TODA_LAUNCH_OK();
    return TODA_OK;
}
