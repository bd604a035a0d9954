// Device helpers shared by the row-scan kernels (K2 exact scan, K4 IVF list scan).
#pragma once

#include "ts_common.cuh"

namespace ts {

template <int NCHUNK>
struct RowsPerTile {
    static constexpr int value = NCHUNK == 1 ? 16 : (NCHUNK == 2 ? 8 : (NCHUNK <= 4 ? 4 : 2));
};

// dot of one 16-byte chunk with the matching slice of q
template <int ELEM>
struct Chunk;
template <>
struct Chunk<2> {  // 8 bf16
    static constexpr int N = 8;
    __device__ static __forceinline__ float dot(const uint4& v, const float* q, float acc) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc = fmaf(__uint_as_float(w[i] << 16), q[2 * i], acc);
            acc = fmaf(__uint_as_float(w[i] & 0xFFFF0000u), q[2 * i + 1], acc);
        }
        return acc;
    }
};
template <>
struct Chunk<4> {  // 4 fp32
    static constexpr int N = 4;
    __device__ static __forceinline__ float dot(const uint4& v, const float* q, float acc) {
        acc = fmaf(__uint_as_float(v.x), q[0], acc);
        acc = fmaf(__uint_as_float(v.y), q[1], acc);
        acc = fmaf(__uint_as_float(v.z), q[2], acc);
        acc = fmaf(__uint_as_float(v.w), q[3], acc);
        return acc;
    }
};

// Transposing reduction: in: a[r] = this lane's partial sum for row r of the tile.
// out: a[0] = full sum for row `row_of_lane(lane)`, replicated over a group of 32/R lanes.
template <int R>
__device__ __forceinline__ void transpose_reduce(float (&a)[R], int lane) {
    int o = 16;
#pragma unroll
    for (int r = R; r > 1; r >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < r / 2; ++i) {
            float send = upper ? a[i] : a[i + r / 2];
            float keep = upper ? a[i + r / 2] : a[i];
            a[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, o);
        }
        o >>= 1;
    }
#pragma unroll
    for (; o > 0; o >>= 1) a[0] += __shfl_xor_sync(0xFFFFFFFFu, a[0], o);
}
template <int R>
__device__ __forceinline__ int row_of_lane(int lane) {
    int row = 0, o = 16;
#pragma unroll
    for (int r = R; r > 1; r >>= 1) {
        if (lane & o) row += r / 2;
        o >>= 1;
    }
    return row;
}


// 16 e4m3 values (one 16-byte chunk) against 16 fp32 query values
template <>
struct Chunk<1> {
    static constexpr int N = 16;
    __device__ static __forceinline__ float dot(const uint4& v, const float* q, float acc) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] & 0xFFFFu), __NV_E4M3);
            const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] >> 16), __NV_E4M3);
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
            acc = fmaf(a.x, q[4 * i + 0], acc);
            acc = fmaf(a.y, q[4 * i + 1], acc);
            acc = fmaf(b.x, q[4 * i + 2], acc);
            acc = fmaf(b.y, q[4 * i + 3], acc);
        }
        return acc;
    }
};

}  // namespace ts
