// Shared epilogue helper of the tcgen05 convolution kernels: per-channel sums for the following BatchNorm.
#pragma once

// v[16] = this lane's row, 16 consecutive columns.  Returns, on every lane, the sum over the warp's 32 rows of column
// ((lane >> 1) & 15): a butterfly that halves the number of live columns at every step (8+4+2+1+1 = 16 shuffles
// instead of 16 x 5 for sixteen independent warp reductions).
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2], d;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float keep = (lane & 16) ? v[j + 8] : v[j];
        float send = (lane & 16) ? v[j] : v[j + 8];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float keep = (lane & 8) ? a[j + 4] : a[j];
        float send = (lane & 8) ? a[j] : a[j + 4];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float keep = (lane & 4) ? b[j + 2] : b[j];
        float send = (lane & 4) ? b[j] : b[j + 2];
        c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        float keep = (lane & 2) ? c[1] : c[0];
        float send = (lane & 2) ? c[0] : c[1];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}
