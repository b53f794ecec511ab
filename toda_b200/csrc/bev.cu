// K9 HeightCompression: dense BEV scatter (and its backward gather).  See include/toda_b200.h.
// out is (B, C*D, H, W) fp32 with out[b, c*D+z, y, x] = features[n, c].
// Rows arrive in canonical (b,z,y,x) order, so consecutive rows are x-neighbours: with a warp laid along
// rows and looping over channels, the 32 stores of one channel land in (mostly) consecutive addresses.
#include "common.cuh"

namespace {

constexpr int kTile = 32;  // rows per tile; tile is transposed through shared memory

__global__ void __launch_bounds__(256) bev_scatter_kernel(const float *__restrict__ feat, const int4 *__restrict__ coords,
                                                          int n, int c, int D, int H, int W, float *__restrict__ out) {
    __shared__ float tile[kTile][129];
    __shared__ long long base[kTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
    const long long hw = (long long)H * W;
    int tiles = (n + kTile - 1) / kTile;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        int r0 = t * kTile;
        if (threadIdx.x < kTile) {
            int r = r0 + threadIdx.x;
            long long b = -1;
            if (r < n) {
                int4 q = __ldg(coords + r);  // b,z,y,x
                b = (((long long)q.x * c) * D + q.y) * hw + (long long)q.z * W + q.w;  // address of channel 0
            }
            base[threadIdx.x] = b;
        }
        for (int c0 = 0; c0 < c; c0 += 128) {
            int cw = min(128, c - c0);
            __syncthreads();
            // coalesced read: each warp reads rows warp, warp+8, ... ; lanes along channels
            for (int rr = warp; rr < kTile; rr += 8) {
                int r = r0 + rr;
                for (int cc = lane; cc < cw; cc += 32) tile[rr][cc] = (r < n) ? __ldg(feat + (size_t)r * c + c0 + cc) : 0.f;
            }
            __syncthreads();
            // transposed write: lanes along rows (x-neighbours), warps along channels
            long long b = base[lane];
            for (int cc = warp; cc < cw; cc += 8)
                if (b >= 0) out[b + (long long)(c0 + cc) * D * hw] = tile[lane][cc];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bev_gather_kernel(const float *__restrict__ dout, const int4 *__restrict__ coords,
                                                         int n, int c, int D, int H, int W, float *__restrict__ dfeat) {
    __shared__ float tile[kTile][129];
    __shared__ long long base[kTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long hw = (long long)H * W;
    int tiles = (n + kTile - 1) / kTile;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        int r0 = t * kTile;
        if (threadIdx.x < kTile) {
            int r = r0 + threadIdx.x;
            long long b = -1;
            if (r < n) {
                int4 q = __ldg(coords + r);
                b = (((long long)q.x * c) * D + q.y) * hw + (long long)q.z * W + q.w;
            }
            base[threadIdx.x] = b;
        }
        for (int c0 = 0; c0 < c; c0 += 128) {
            int cw = min(128, c - c0);
            __syncthreads();
            long long b = base[lane];
            for (int cc = warp; cc < cw; cc += 8) tile[lane][cc] = (b >= 0) ? __ldg(dout + b + (long long)(c0 + cc) * D * hw) : 0.f;
            __syncthreads();
            for (int rr = warp; rr < kTile; rr += 8) {
                int r = r0 + rr;
                if (r < n)
                    for (int cc = lane; cc < cw; cc += 32) dfeat[(size_t)r * c + c0 + cc] = tile[rr][cc];
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int toda_bev_scatter_fwd(const float *features, const int32_t *coords, int n, int c, int batch, int D, int H, int W,
                                    float *out, void *stream) {
    TODA_CHECK_ARG(n >= 0 && c > 0 && batch > 0 && D > 0 && H > 0 && W > 0 && out, "bev_scatter_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    TODA_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)batch * c * D * H * W * sizeof(float), st));
    if (n == 0) return TODA_OK;
    TODA_CHECK_ARG(features && coords, "bev_scatter_fwd: null pointer");
    int tiles = ceil_div(n, kTile);
    bev_scatter_kernel<<<wave_grid((int64_t)tiles * 256, 256), 256, 0, st>>>(features, (const int4 *)coords, n, c, D, H, W, out);
    TODA_LAUNCH_OK();
    return TODA_OK;
}

extern "C" int toda_bev_scatter_bwd(const float *d_out, const int32_t *coords, int n, int c, int batch, int D, int H, int W,
                                    float *d_features, void *stream) {
    TODA_CHECK_ARG(n >= 0 && c > 0 && batch > 0 && D > 0 && H > 0 && W > 0, "bev_scatter_bwd: bad args");
    if (n == 0) return TODA_OK;
    TODA_CHECK_ARG(d_out && coords && d_features, "bev_scatter_bwd: null pointer");
    int tiles = ceil_div(n, kTile);
    bev_gather_kernel<<<wave_grid((int64_t)tiles * 256, 256), 256, 0, (cudaStream_t)stream>>>(d_out, (const int4 *)coords, n, c,
                                                                                             D, H, W, d_features);
    TODA_LAUNCH_OK();
    return TODA_OK;
}
