// K1 — fused L2-normalise + cast (+ zero-pad rows to a 16-byte multiple).
//
// Replaces the trailing F.normalize of `model.encode(..., normalize_embeddings=True)`
// (reference streamlit_app.py:173; ec2/generate_embeddings/embeddings.py:27,35) and the
// F.normalize pair inside sentence_transformers.util.cos_sim (test_app.py:76), applied ONCE
// when the corpus is written instead of on every query.
//
// Arithmetic (bit-defined, restated by oracle.normalize_f64):
//   ss   = sum_i (double)x_i^2            (fp64 accumulate, lane-strided + butterfly)
//   nrm  = max((float)sqrt(ss), 1e-12f)
//   y_i  = x_i / nrm                      (IEEE fp32 division)
//   out  = bf16_rn(y_i) | y_i
// HBM-bound: reads 4*D, writes 2*D (bf16) bytes per row. One warp per row, grid-stride.
#include "ts_common.cuh"

#include <algorithm>

namespace ts {

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__device__ __forceinline__ void store_from_float(T* p, float v);
template <>
__device__ __forceinline__ void store_from_float<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
}

// One warp per row, the row held in registers: lane l owns elements l, l+32, ... (NJ of them).
//   * one coalesced pass over the source (all NJ loads of a lane are independent and in flight together);
//   * ||x||^2 accumulated in fp64 in exactly the order K2's in-kernel query normalisation uses;
//   * the IEEE fp32 division x / nrm is evaluated as (float)((double)x * (1.0 / (double)nrm)): 3 instructions
//     instead of ~10, and bit-identical — the fp64 product is within 2^-52 of the exact quotient, and the
//     quotient of two 24-bit floats is never closer than 2^-49 (relative) to an fp32 rounding boundary;
//     results in the fp32-subnormal range (where that bound does not apply) take the division instruction;
//   * bf16 stores are 4-byte: lane pairs swap one value per two elements (even lanes store element pair
//     (l, l+1) of step j, odd lanes that of step j+1).
template <typename SRC, typename DST, int NJ>
__global__ void __launch_bounds__(256) normalize_cast_kernel(const SRC* __restrict__ src, int64_t n,
                                                             int dim, int dim_pad, int normalize,
                                                             DST* __restrict__ dst,
                                                             float* __restrict__ max_norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp0; row < n; row += nwarps) {
        const SRC* x = src + row * (int64_t)dim;
        DST* y = dst + row * (int64_t)dim_pad;
        float v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int i = lane + 32 * j;
            v[j] = (i < dim) ? load_as_float<SRC>(x + i) : 0.0f;
        }
        if (normalize) {
            double ss = 0.0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const double d = (double)v[j];
                ss = fma(d, d, ss);          // elements past dim are zeros: they leave ss unchanged
            }
            ss = warp_sum(ss);
            const float nrm = fmaxf((float)sqrt(ss), 1e-12f);
            const double rd = 1.0 / (double)nrm;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float q = (float)((double)v[j] * rd);
                if (fabsf(q) < 1.17549435e-38f) q = __fdiv_rn(v[j], nrm);
                v[j] = q;
            }
        }
        float qs = 0.f;  // squared norm of the row AS STORED (bounds |<dq, row>| in K3's certificate)
        if constexpr (sizeof(DST) == 4) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int i = lane + 32 * j;
                if (i < dim_pad) store_from_float<DST>(y + i, v[j]);
                qs = fmaf(v[j], v[j], qs);
            }
        } else {
            const bool odd = (lane & 1) != 0;
#pragma unroll
            for (int j = 0; j < NJ; j += 2) {
                const float recv = __shfl_xor_sync(0xFFFFFFFFu, odd ? v[j] : v[j + 1], 1);
                const float a0 = odd ? recv : v[j];
                const float a1 = odd ? v[j + 1] : recv;
                const int i = odd ? (lane - 1 + 32 * (j + 1)) : (lane + 32 * j);
                const __nv_bfloat162 b = __floats2bfloat162_rn(a0, a1);
                if (i < dim_pad) *reinterpret_cast<__nv_bfloat162*>(y + i) = b;
                const float2 sb = __bfloat1622float2(b);
                qs = fmaf(sb.x, sb.x, fmaf(sb.y, sb.y, qs));
            }
        }
        if (max_norm2 != nullptr) {
            qs = warp_sum(qs);
            // non-negative floats order like their bit patterns; NaN/Inf rows poison the bound on purpose
            if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(max_norm2), __float_as_uint(qs));
        }
    }
}

template <typename SRC>
__global__ void __launch_bounds__(256) dequant_rows_kernel(const SRC* __restrict__ src, int64_t n,
                                                           int dim, int dim_pad,
                                                           float* __restrict__ dst) {
    int64_t total = n * (int64_t)dim;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / dim;
        int c = (int)(t - r * dim);
        dst[t] = load_as_float<SRC>(src + r * (int64_t)dim_pad + c);
    }
}

template <typename SRC, int NJ>
static int launch_nc_nj(const void* src, int64_t n, int dim, int dim_pad, int normalize, void* dst,
                        int dst_dtype, cudaStream_t s, float* max_norm2) {
    int64_t blocks64 = (n + 7) / 8;
    int blocks = (int)(blocks64 > 148 * 32 ? 148 * 32 : blocks64);
    if (dst_dtype == TS_BF16) {
        normalize_cast_kernel<SRC, __nv_bfloat16, NJ><<<blocks, 256, 0, s>>>(
            (const SRC*)src, n, dim, dim_pad, normalize, (__nv_bfloat16*)dst, max_norm2);
    } else if (dst_dtype == TS_F32) {
        normalize_cast_kernel<SRC, float, NJ><<<blocks, 256, 0, s>>>((const SRC*)src, n, dim, dim_pad,
                                                                     normalize, (float*)dst, max_norm2);
    } else {
        set_error("normalize_cast: unsupported destination dtype %d", dst_dtype);
        return TS_ERR_UNSUPPORTED;
    }
    TS_LAUNCH_CHECK();
    return TS_OK;
}

template <typename SRC>
static int launch_nc(const void* src, int64_t n, int dim, int dim_pad, int normalize, void* dst,
                     int dst_dtype, cudaStream_t s, float* max_norm2) {
    if (n == 0) return TS_OK;
    if (dim_pad <= 256) return launch_nc_nj<SRC, 8>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
    if (dim_pad <= 512) return launch_nc_nj<SRC, 16>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
    if (dim_pad <= 768) return launch_nc_nj<SRC, 24>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
    if (dim_pad <= 1024) return launch_nc_nj<SRC, 32>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
    return launch_nc_nj<SRC, 64>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
}

int launch_normalize_cast(const void* src, int src_dtype, int64_t n, int dim, int dim_pad,
                          int normalize, void* dst, int dst_dtype, cudaStream_t s, float* max_norm2) {
    switch (src_dtype) {
        case TS_F32: return launch_nc<float>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
        case TS_BF16:
            return launch_nc<__nv_bfloat16>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
        case TS_F16: return launch_nc<__half>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2);
        default:
            set_error("normalize_cast: unsupported source dtype %d", src_dtype);
            return TS_ERR_BAD_ARG;
    }
}

int launch_prepare_queries(const void* q, int q_dtype, int nq, int dim, int dim_pad, int normalize,
                           float* out_f32, cudaStream_t s) {
    return launch_normalize_cast(q, q_dtype, nq, dim, dim_pad, normalize, out_f32, TS_F32, s);
}

// max over rows of ||row||^2 (as stored), atomically folded into *out — the bound K3's certificate uses.
template <typename T>
__global__ void __launch_bounds__(256) max_norm2_kernel(const T* __restrict__ rows, int64_t n, int dim_pad,
                                                        float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    for (int64_t r = w0; r < n; r += (int64_t)gridDim.x * 8) {
        float s = 0.f;
        for (int i = lane; i < dim_pad; i += 32) {
            const float v = load_as_float<T>(rows + r * (int64_t)dim_pad + i);
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(s));
    }
}

int launch_max_norm2(const void* rows, int dtype, int64_t n, int dim_pad, float* out, cudaStream_t s) {
    if (n == 0) return TS_OK;
    const int blocks = (int)std::min<int64_t>((n + 7) / 8, 148 * 16);
    if (dtype == TS_BF16)
        max_norm2_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)rows, n, dim_pad, out);
    else if (dtype == TS_F32)
        max_norm2_kernel<float><<<blocks, 256, 0, s>>>((const float*)rows, n, dim_pad, out);
    else {
        set_error("max_norm2: unsupported dtype %d", dtype);
        return TS_ERR_UNSUPPORTED;
    }
    TS_LAUNCH_CHECK();
    return TS_OK;
}

int launch_dequant_rows(const void* src, int src_dtype, int64_t n, int dim, int dim_pad, float* dst,
                        cudaStream_t s) {
    if (n == 0) return TS_OK;
    int64_t total = n * (int64_t)dim;
    int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (src_dtype == TS_BF16)
        dequant_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)src, n, dim,
                                                                 dim_pad, dst);
    else if (src_dtype == TS_F32)
        dequant_rows_kernel<float><<<blocks, 256, 0, s>>>((const float*)src, n, dim, dim_pad, dst);
    else {
        set_error("dequant_rows: unsupported dtype %d", src_dtype);
        return TS_ERR_UNSUPPORTED;
    }
    TS_LAUNCH_CHECK();
    return TS_OK;
}

}  // namespace ts
