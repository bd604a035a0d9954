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

// One warp per row, the row held in registers: lane l owns the element PAIRS (2l, 2l+1) + 64j, j < NP.
//   * one coalesced pass over the source (8-byte loads, all of a lane's loads independent and in flight);
//   * ||x||^2 accumulated in fp64, lane-local in ascending element order, then the xor butterfly — exactly the
//     order K2's in-kernel query normalisation uses, so both produce the same bits;
//   * the IEEE fp32 division x / nrm is evaluated as (float)((double)x * (1.0 / (double)nrm)): 3 instructions
//     instead of ~10, and bit-identical — the fp64 product is within 2^-52 of the exact quotient, and the
//     quotient of two 24-bit floats is never closer than 2^-49 (relative) to an fp32 rounding boundary. The
//     bound does not hold for results in the fp32-subnormal range: a lane whose smallest |result| is below
//     FLT_MIN (exact zeros included) redoes its elements with the division instruction;
//   * stores are one bf16x2 (or float2) per pair, no shuffles;
//   * max ||stored row||^2 for K3's certificate is an upper bound: sum q^2 in fp32, times (1 + 2^-8)^2 for the
//     bf16 rounding and a little slack for the fp32 summation.
// a pair that lies inside the row, as one vector load (row and pair aligned) / two scalar loads
template <typename SRC>
__device__ __forceinline__ float2 ld_pair_vec(const SRC* p);
template <>
__device__ __forceinline__ float2 ld_pair_vec<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <>
__device__ __forceinline__ float2 ld_pair_vec<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <>
__device__ __forceinline__ float2 ld_pair_vec<__half>(const __half* p) {
    return __half22float2(*reinterpret_cast<const __half2*>(p));
}
// Branch-free inside the unrolled loop (predicated loads), so all of a lane's loads issue back to back.
template <typename SRC, int NP>
__device__ __forceinline__ void load_row_pairs(const SRC* x, int lane, int dim, bool vec_ok, float2 (&v)[NP]) {
    if (vec_ok) {   // warp-uniform: even dim => every pair is aligned and entirely inside or outside the row
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int i = 2 * lane + 64 * j;
            v[j] = make_float2(0.f, 0.f);
            if (i < dim) v[j] = ld_pair_vec<SRC>(x + i);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int i = 2 * lane + 64 * j;
            v[j].x = (i < dim) ? load_as_float<SRC>(x + i) : 0.f;
            v[j].y = (i + 1 < dim) ? load_as_float<SRC>(x + i + 1) : 0.f;
        }
    }
}

template <typename SRC, typename DST, int NP>
__global__ void __launch_bounds__(256, NP <= 16 ? 2 : 1) normalize_cast_kernel(const SRC* __restrict__ src, int64_t n,
                                                             int dim, int dim_pad, int normalize,
                                                             DST* __restrict__ dst,
                                                             float* __restrict__ max_norm2,
                                                             const int64_t* __restrict__ dst_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool vec_ok = (dim & 1) == 0;   // even rows keep every pair aligned to its vector load
    float warp_max = 0.f;                 // running max over this warp's rows: ONE atomic per warp at the end
    for (int64_t row = warp0; row < n; row += nwarps) {
        const SRC* x = src + row * (int64_t)dim;
        // upsert: source row `row` lands in stored row dst_rows[row] (negative = superseded by a later source row
        // with the same id, skipped); append / build: rows land contiguously
        int64_t out_row = row;
        if (dst_rows != nullptr) {
            out_row = dst_rows[row];
            if (out_row < 0) continue;
        }
        DST* y = dst + out_row * (int64_t)dim_pad;
        float2 v[NP];
        load_row_pairs<SRC, NP>(x, lane, dim, vec_ok, v);
        if (normalize) {
            double ss = 0.0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {   // elements past dim are zeros: they leave ss unchanged
                const double a = (double)v[j].x, b = (double)v[j].y;
                ss = fma(a, a, ss);
                ss = fma(b, b, ss);
            }
            ss = warp_sum(ss);
            const float nrm = fmaxf((float)sqrt(ss), 1e-12f);
            const double rd = 1.0 / (double)nrm;
            float amin = 3.0e38f;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                v[j].x = (float)((double)v[j].x * rd);
                v[j].y = (float)((double)v[j].y * rd);
                amin = fminf(amin, fminf(fabsf(v[j].x), fabsf(v[j].y)));
            }
            if (amin < 1.17549435e-38f) {   // subnormal (or zero) results in this lane: reload, divide exactly
                load_row_pairs<SRC, NP>(x, lane, dim, vec_ok, v);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    v[j].x = __fdiv_rn(v[j].x, nrm);
                    v[j].y = __fdiv_rn(v[j].y, nrm);
                }
            }
        }
        float qs = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int i = 2 * lane + 64 * j;
            qs = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, qs));
            if (i < dim_pad) {   // dim_pad is a multiple of 8 and i is even: the pair is inside the row
                if constexpr (sizeof(DST) == 4) *reinterpret_cast<float2*>(y + i) = v[j];
                else *reinterpret_cast<__nv_bfloat162*>(y + i) = __floats2bfloat162_rn(v[j].x, v[j].y);
            }
        }
        if (max_norm2 != nullptr) {
            qs = warp_sum(qs) * (sizeof(DST) == 4 ? 1.0001f : 1.0081f);
            // non-negative floats order like their bit patterns; NaN/Inf rows poison the bound on purpose
            warp_max = __uint_as_float(max(__float_as_uint(warp_max), __float_as_uint(qs)));
        }
    }
    // (a per-row atomic on one address serialises in L2: 10^6 of them cost more than streaming the rows)
    if (max_norm2 != nullptr && lane == 0 && __float_as_uint(warp_max) != 0u)
        atomicMax(reinterpret_cast<unsigned int*>(max_norm2), __float_as_uint(warp_max));
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

template <typename SRC, int NP>
static int launch_nc_nj(const void* src, int64_t n, int dim, int dim_pad, int normalize, void* dst,
                        int dst_dtype, cudaStream_t s, float* max_norm2, const int64_t* dst_rows) {
    int64_t blocks64 = (n + 7) / 8;
    int blocks = (int)(blocks64 > 148 * 32 ? 148 * 32 : blocks64);
    if (dst_dtype == TS_BF16) {
        normalize_cast_kernel<SRC, __nv_bfloat16, NP><<<blocks, 256, 0, s>>>(
            (const SRC*)src, n, dim, dim_pad, normalize, (__nv_bfloat16*)dst, max_norm2, dst_rows);
    } else if (dst_dtype == TS_F32) {
        normalize_cast_kernel<SRC, float, NP><<<blocks, 256, 0, s>>>((const SRC*)src, n, dim, dim_pad,
                                                                     normalize, (float*)dst, max_norm2, dst_rows);
    } else {
        set_error("normalize_cast: unsupported destination dtype %d", dst_dtype);
        return TS_ERR_UNSUPPORTED;
    }
    TS_LAUNCH_CHECK();
    return TS_OK;
}

template <typename SRC>
static int launch_nc(const void* src, int64_t n, int dim, int dim_pad, int normalize, void* dst,
                     int dst_dtype, cudaStream_t s, float* max_norm2, const int64_t* dst_rows) {
    if (n == 0) return TS_OK;
    if (dim_pad <= 256) return launch_nc_nj<SRC, 4>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
    if (dim_pad <= 512) return launch_nc_nj<SRC, 8>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
    if (dim_pad <= 768) return launch_nc_nj<SRC, 12>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
    if (dim_pad <= 1024) return launch_nc_nj<SRC, 16>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
    return launch_nc_nj<SRC, 32>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
}

int launch_normalize_cast(const void* src, int src_dtype, int64_t n, int dim, int dim_pad,
                          int normalize, void* dst, int dst_dtype, cudaStream_t s, float* max_norm2,
                          const int64_t* dst_rows) {
    switch (src_dtype) {
        case TS_F32: return launch_nc<float>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
        case TS_BF16:
            return launch_nc<__nv_bfloat16>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
        case TS_F16: return launch_nc<__half>(src, n, dim, dim_pad, normalize, dst, dst_dtype, s, max_norm2, dst_rows);
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
    float warp_max = 0.f;   // one atomic per warp, not per row (same-address atomics serialise in L2)
    for (int64_t r = w0; r < n; r += (int64_t)gridDim.x * 8) {
        float s = 0.f;
        for (int i = lane; i < dim_pad; i += 32) {
            const float v = load_as_float<T>(rows + r * (int64_t)dim_pad + i);
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        warp_max = __uint_as_float(max(__float_as_uint(warp_max), __float_as_uint(s)));
    }
    if (lane == 0 && __float_as_uint(warp_max) != 0u)
        atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(warp_max));
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
