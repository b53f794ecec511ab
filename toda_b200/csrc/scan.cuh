// Two-launch exclusive scan for the small integer arrays of the geometry stage
// (<= a few hundred thousand elements: L1 bitmap words, point-flag words).
//   launch 1: per-tile sums (tile = 1024 elements)
//   launch 2: every CTA reduces the sums of the tiles before it (tiles <= ~4k, so this is cheaper
//             than a third launch), scans its own tile and adds the offset.
// No inter-CTA waiting (the profiling guide forbids spin-waits between CTAs of different launches,
// and a decoupled look-back would need a zeroed status array per call).
#pragma once
#include "common.cuh"

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename Load>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(Load load, int n, int *tile_sums) {
    __shared__ int warp_sums[kScanThreads / 32];
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += load(base + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sums[w];
        tile_sums[blockIdx.x] = t;
    }
}

template <typename Load>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(Load load, int n, const int *tile_sums, int *out,
                                                                 int *total_out) {
    __shared__ int warp_sums[kScanThreads / 32];
    __shared__ int tile_offset;
    // offset of this tile = sum of the sums of all earlier tiles
    int part = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += kScanThreads) part += tile_sums[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sums[w];
        tile_offset = t;
    }
    __syncthreads();
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? load(base + i) : 0;
        s += v[i];
    }
    // exclusive scan of the per-thread sums across the CTA
    int incl = s;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // warp_sums reuse
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_sums[w];
    int run = tile_offset + woff + incl - s;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    // the thread that owns the last element publishes the grand total
    if (base <= n - 1 && n - 1 < base + kScanItems) {
        if (total_out) *total_out = run;
        out[n] = run;  // every scan output array has n+1 slots
    }
}

// out must have n+1 entries (out[n] = total).  tile_sums: ceil(n/1024) ints of scratch.
template <typename Load>
static inline int scan_exclusive(Load load, int n, int *out, int *tile_sums, int *total_out, cudaStream_t st) {
    if (n <= 0) {
        // nothing to scan: total = 0
        TODA_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int), st));
        if (total_out) TODA_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int), st));
        return TODA_OK;
    }
    int tiles = ceil_div(n, kScanTile);
    scan_tile_sums_kernel<<<tiles, kScanThreads, 0, st>>>(load, n, tile_sums);
    TODA_LAUNCH_OK();
    scan_apply_kernel<<<tiles, kScanThreads, 0, st>>>(load, n, tile_sums, out, total_out);
    TODA_LAUNCH_OK();
    return TODA_OK;
}

struct LoadInt {
    const int *p;
    __device__ __forceinline__ int operator()(int i) const { return p[i]; }
};
struct LoadPopc32 {
    const unsigned *p;
    __device__ __forceinline__ int operator()(int i) const { return __popc(p[i]); }
};
