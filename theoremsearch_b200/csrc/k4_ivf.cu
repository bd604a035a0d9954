// K4 — IVF-Flat (the pgvector `ivfflat` equivalent BASELINE.json config 5 asks for).
//
// The reference never creates the index (rds_schema.sql has no CREATE INDEX, so production runs the
// exact seq-scan of streamlit_app.py:275-282); pgvector's ivfflat semantics are what is restated:
// k-means centroids over a sample, every row filed under its nearest centroid, a query probes the
// `nprobe` nearest lists and ranks only their rows.
//
// Kernels
//   K4a assign_argmax_kernel   rows x centroids bf16 GEMM on tcgen05/TMEM with the arg-max over
//                              centroids fused into the epilogue (thread <-> row, columns <-> centroids,
//                              so the running maximum is thread-local). Tensor-bound: 2*N*nlist*D FLOPs.
//                              Used by k-means (sample rows) and by the build (all rows).
//   update_centroids_kernel    spherical k-means update: one CTA per centroid sums its rows in ascending
//                              row order (deterministic, no atomics), normalises, re-seeds empty lists.
//   gather_quantize_kernel     build: rows permuted into list order, stored e4m3 with one fp32 scale per
//                              row (scale = max|x|/448) or as bf16.
//   K4b list_scan_kernel       search: the probed lists of one query form one virtual row sequence that is
//                              split evenly over all warps of `parts` CTAs; per-warp 1-D TMA pipelines,
//                              the same dot/transposing-reduce/WarpTopK body as K2, last CTA merges.
//                              HBM-bound: (rows in probed lists) * row_bytes per query.
//   K4c ivf_rescore_kernel     the rescore_k survivors are re-scored against the full-precision corpus
//                              rows with the fp32 query in K2's summation order (so returned scores are
//                              bit-identical to the exact path's), sorted, top-k emitted.
// The coarse step (top-nprobe centroids) is the exact-search path itself (K2 / K3) run over the
// centroid table.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "merge_device.cuh"
#include "scan_device.cuh"
#include "umma_device.cuh"

namespace ts {

namespace k4 {
using namespace umma;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr size_t SMEM_TILES = (size_t)STAGES * (A_BYTES + B_BYTES);
constexpr size_t SMEM_BYTES = SMEM_TILES + 16 * sizeof(uint64_t) + 1024;
}  // namespace k4

struct AssignParams {
    int64_t n_rows;
    int nlist;
    int num_k_blocks, num_m_blocks, num_n_blocks;
    int bn;              // centroids per n-block (multiple of 32, <= 256)
    uint32_t* assign;    // [n_rows] nearest centroid (ties -> lower centroid)
    float* best;         // [n_rows] its score, or nullptr
};

// ---------------------------------------------------------------------------------- K4a
__global__ void __launch_bounds__(k4::THREADS, 1)
assign_argmax_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const AssignParams p) {
    using namespace k4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_TILES);
    uint64_t* full = bars;              // [STAGES]
    uint64_t* empty = bars + 4;         // [STAGES]
    uint64_t* tmem_full = bars + 8;     // [ACC_STAGES]
    uint64_t* tmem_empty = bars + 10;   // [ACC_STAGES]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    // A CTA owns whole m-blocks (128 rows) and walks every centroid block for them, so the running
    // arg-max never leaves the epilogue thread's registers.
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol_a = l2_policy_evict_first();
            const uint64_t pol_b = l2_policy_evict_last();   // the centroid table is re-read by every m-block
            int stage = 0;
            uint32_t phase = 0;
            for (int mb = blockIdx.x; mb < p.num_m_blocks; mb += gridDim.x) {
                for (int nb = 0; nb < p.num_n_blocks; ++nb) {
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        mbar_wait_wd(&empty[stage], phase ^ 1u);
                        mbar_expect_tx(&full[stage], A_BYTES + (uint32_t)p.bn * (BK * 2));
                        tma_load_2d(smem_a + (size_t)stage * A_BYTES, &tmap_a, kb * BK, mb * BM, &full[stage], pol_a);
                        tma_load_2d(smem_b + (size_t)stage * B_BYTES, &tmap_b, kb * BK, nb * p.bn, &full[stage], pol_b);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int mb = blockIdx.x; mb < p.num_m_blocks; mb += gridDim.x) {
                for (int nb = 0; nb < p.num_n_blocks; ++nb, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                    mbar_wait_wd(&tmem_empty[acc], acc_phase ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        mbar_wait_wd(&full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = umma_smem_desc(smem_a + (size_t)stage * A_BYTES);
                        const uint64_t db = umma_smem_desc(smem_b + (size_t)stage * B_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        tcgen05_commit(&empty[stage]);
                        if (kb == p.num_k_blocks - 1) tcgen05_commit(&tmem_full[acc]);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else {
        const int lane_base = 32 * (warp & 3);
        int it = 0;
        for (int mb = blockIdx.x; mb < p.num_m_blocks; mb += gridDim.x) {
            float best = -INFINITY;
            uint32_t best_idx = 0;
            for (int nb = 0; nb < p.num_n_blocks; ++nb, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                const int c0 = nb * p.bn;
                const int valid = (p.nlist - c0 < p.bn) ? (p.nlist - c0) : p.bn;   // real centroids in this block
                mbar_wait_wd(&tmem_full[acc], acc_phase);
                tcgen05_fence_after();
                const uint32_t taddr0 = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int c = 0; c * 32 < valid; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    if (c * 32 + 32 <= valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float f = __uint_as_float(v[j]);
                            if (f > best) {   // strict: the lower centroid keeps a tie
                                best = f;
                                best_idx = (uint32_t)(c0 + c * 32 + j);
                            }
                        }
                    } else {   // ragged tail: columns >= valid are zero-filled padding, not centroids
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float f = __uint_as_float(v[j]);
                            if (c * 32 + j < valid && f > best) {
                                best = f;
                                best_idx = (uint32_t)(c0 + c * 32 + j);
                            }
                        }
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(&tmem_empty[acc]);
            }
            const int64_t row = (int64_t)mb * BM + lane_base + lane;
            if (row < p.n_rows) {
                p.assign[row] = best_idx;
                if (p.best != nullptr) p.best[row] = best;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// rows: bf16, `row_stride` bytes apart (a strided view of the corpus is a valid sample); centroids: bf16
// [nlist, dim_pad] contiguous.
static int launch_assign(int device, const void* rows, int64_t n_rows, size_t row_stride, const __nv_bfloat16* cent,
                         int nlist, int dim_pad, uint32_t* assign, float* best, cudaStream_t s) {
    using namespace k4;
    if (n_rows == 0) return TS_OK;
    TS_REQUIRE(n_rows < (int64_t)1 << 31, TS_ERR_UNSUPPORTED, "ivf assign: more than 2^31-1 rows");
    CUtensorMap tmap_a, tmap_b;
    int rc = make_tmap_bf16_rows(&tmap_a, rows, (uint64_t)n_rows, (uint64_t)dim_pad, row_stride, BM);
    if (rc) return rc;
    AssignParams p;
    p.n_rows = n_rows;
    p.nlist = nlist;
    p.num_k_blocks = (dim_pad + BK - 1) / BK;
    p.num_m_blocks = (int)((n_rows + BM - 1) / BM);
    p.num_n_blocks = (nlist + BN - 1) / BN;
    p.bn = ((nlist + p.num_n_blocks - 1) / p.num_n_blocks + 31) / 32 * 32;
    p.assign = assign;
    p.best = best;
    rc = make_tmap_bf16_rows(&tmap_b, cent, (uint64_t)nlist, (uint64_t)dim_pad, (uint64_t)dim_pad * 2, (uint32_t)p.bn);
    if (rc) return rc;
    static bool attr_set[64] = {false};   // function attributes are per device
    if (!attr_set[device & 63]) {
        TS_CHECK_CUDA(cudaFuncSetAttribute(assign_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)SMEM_BYTES));
        attr_set[device & 63] = true;
    }
    const int grid = std::min(p.num_m_blocks, sm_count(device));
    assign_argmax_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tmap_a, tmap_b, p);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

// ---------------------------------------------------------------------------------- k-means pieces
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// centroid c starts as sample row c*step + hash % step: nlist distinct rows spread over the sample
__global__ void __launch_bounds__(256) init_centroids_kernel(const uint8_t* __restrict__ rows, size_t row_stride,
                                                             int64_t n_rows, int nlist, int dim_pad, uint64_t seed,
                                                             float* __restrict__ cent, __nv_bfloat16* __restrict__ cent16) {
    const int c = blockIdx.x;
    const int64_t step = n_rows / nlist;
    const int64_t r = (int64_t)c * step + (int64_t)(splitmix64(seed ^ (uint64_t)c) % (uint64_t)step);
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(rows + (size_t)r * row_stride);
    for (int i = threadIdx.x; i < dim_pad; i += blockDim.x) {
        cent[(size_t)c * dim_pad + i] = __bfloat162float(src[i]);
        cent16[(size_t)c * dim_pad + i] = src[i];
    }
}

__global__ void iota_u32_kernel(uint32_t* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint32_t)i;
}

// offsets[c] = first position in the ascending `sorted_assign` whose list id is >= c  (c in [0, nlist])
__global__ void list_offsets_kernel(const uint32_t* __restrict__ sorted_assign, int64_t n, int nlist,
                                    int64_t* __restrict__ offsets) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nlist) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_assign[mid] < (uint32_t)c) lo = mid + 1; else hi = mid;
    }
    offsets[c] = lo;
}

// One CTA per centroid: sum of its member rows (ascending row order), normalised. Empty or
// degenerate lists are re-seeded from a pseudo-random sample row.
__global__ void __launch_bounds__(256) update_centroids_kernel(const uint8_t* __restrict__ rows, size_t row_stride,
                                                               int64_t n_rows, const uint32_t* __restrict__ members,
                                                               const int64_t* __restrict__ offsets, int dim_pad,
                                                               uint64_t seed, float* __restrict__ cent,
                                                               __nv_bfloat16* __restrict__ cent16,
                                                               const uint32_t* __restrict__ assign, int nlist,
                                                               int split_small) {
    constexpr int MAXC = TS_MAX_DIM / 256;   // columns per thread
    __shared__ float red[8];
    __shared__ float s_norm;
    __shared__ long long s_alt;
    const int c = blockIdx.x;
    const int64_t b = offsets[c], e = offsets[c + 1];
    float sum[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) sum[i] = 0.f;
#pragma unroll 4
    for (int64_t m = b; m < e; ++m) {
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(rows + (size_t)members[m] * row_stride);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int col = threadIdx.x + 256 * i;
            if (col < dim_pad) sum[i] += __bfloat162float(src[col]);
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) ss = fmaf(sum[i], sum[i], ss);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        s_norm = sqrtf(t);
    }
    __syncthreads();
    const float nrm = s_norm;
    // Re-seed: empty / degenerate lists always; with split_small also lists below a quarter of the mean size —
    // random-row initialisation leaves some true clusters without a centroid and others with several, a local
    // optimum Lloyd iterations cannot leave. The new centroid is a member of a list at least twice the mean
    // size (a few hashed tries), i.e. the big merged lists get split; the last iterations run without it.
    const double mean = (double)n_rows / (double)nlist;
    const bool reseed = !(nrm > 1e-20f) || (split_small && (double)(e - b) < 0.25 * mean);
    if (threadIdx.x == 0) {
        long long pick = -1;
        if (reseed) {
            for (int t = 0; t < 16; ++t) {
                const uint64_t r = splitmix64(seed ^ (0xABCDull << 32) ^ ((uint64_t)t << 48) ^ (uint64_t)c) % (uint64_t)n_rows;
                const uint32_t l = assign[r];
                pick = (long long)r;
                if ((double)(offsets[l + 1] - offsets[l]) > 2.0 * mean) break;
            }
        }
        s_alt = pick;
    }
    __syncthreads();
    const __nv_bfloat16* alt = nullptr;
    if (reseed) alt = reinterpret_cast<const __nv_bfloat16*>(rows + (size_t)s_alt * row_stride);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int col = threadIdx.x + 256 * i;
        if (col < dim_pad) {
            const float v = reseed ? __bfloat162float(alt[col]) : __fdiv_rn(sum[i], nrm);
            cent[(size_t)c * dim_pad + col] = v;
            cent16[(size_t)c * dim_pad + col] = __float2bfloat16_rn(v);
        }
    }
}

// max ||centroid_bf16||^2 (the exactness certificate of the batched coarse search needs the bound)
__global__ void __launch_bounds__(256) max_norm2_bf16_kernel(const __nv_bfloat16* __restrict__ rows, int n, int dim_pad,
                                                             float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    float s = 0.f;
    for (int i = lane; i < dim_pad; i += 32) {
        const float v = __bfloat162float(rows[(size_t)r * dim_pad + i]);
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(s));
}

// ---------------------------------------------------------------------------------- build: permute + quantise
// One warp per list position: fetch the corpus row filed there, store it as e4m3 with a per-row scale
// (scale = max|x| / 448, 1 for an all-zero row; q = e4m3_rn(x / scale), saturating) or as bf16.
template <int LIST_ELEM>
__global__ void __launch_bounds__(256) gather_quantize_kernel(const uint8_t* __restrict__ corpus, uint32_t src_row_bytes,
                                                              const uint32_t* __restrict__ list_rows, int64_t n,
                                                              int dim_pad, uint32_t dst_row_bytes,
                                                              uint8_t* __restrict__ dst, float* __restrict__ scales) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    for (int64_t pos = w0; pos < n; pos += (int64_t)gridDim.x * 8) {
        const uint8_t* src = corpus + (size_t)list_rows[pos] * src_row_bytes;
        uint8_t* out = dst + (size_t)pos * dst_row_bytes;
        if constexpr (LIST_ELEM == 2) {
            for (uint32_t off = lane * 16u; off < dst_row_bytes; off += 512u)
                *reinterpret_cast<uint4*>(out + off) = __ldg(reinterpret_cast<const uint4*>(src + off));
        } else {
            float amax = 0.f;
            for (uint32_t off = lane * 16u; off < src_row_bytes; off += 512u) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + off));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    amax = fmaxf(amax, fabsf(__uint_as_float(w[i] << 16)));
                    amax = fmaxf(amax, fabsf(__uint_as_float(w[i] & 0xFFFF0000u)));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
            const float scale = amax > 0.f ? __fdiv_rn(amax, 448.0f) : 1.0f;
            if (lane == 0) scales[pos] = scale;
            // 16 output bytes per lane per step = 16 elements = 32 source bytes
            for (uint32_t ob = lane * 16u; ob < dst_row_bytes; ob += 512u) {
                uint32_t packed[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t sb = ob * 2u + h * 16u;
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (sb < src_row_bytes) v = __ldg(reinterpret_cast<const uint4*>(src + sb));
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i += 2) {
                        const float a0 = __fdiv_rn(__uint_as_float(w[i] << 16), scale);
                        const float a1 = __fdiv_rn(__uint_as_float(w[i] & 0xFFFF0000u), scale);
                        const float a2 = __fdiv_rn(__uint_as_float(w[i + 1] << 16), scale);
                        const float a3 = __fdiv_rn(__uint_as_float(w[i + 1] & 0xFFFF0000u), scale);
                        const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a0, a1), __NV_SATFINITE, __NV_E4M3);
                        const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(a2, a3), __NV_SATFINITE, __NV_E4M3);
                        packed[h * 2 + i / 2] = lo | (hi << 16);
                    }
                }
                *reinterpret_cast<uint4*>(out + ob) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
        }
    }
}

// dequantise list positions [first, first+n) to fp32 [n, dim] (tests / diagnostics)
__global__ void dequant_list_kernel(const uint8_t* __restrict__ data, uint32_t row_bytes, int list_dtype,
                                    const float* __restrict__ scales, int64_t first, int64_t n, int dim,
                                    float* __restrict__ out) {
    const int64_t total = n * (int64_t)dim;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / dim;
        const int c = (int)(t - r * dim);
        const uint8_t* row = data + (size_t)(first + r) * row_bytes;
        float v;
        if (list_dtype == TS_BF16) {
            v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(row)[c]);
        } else {
            const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)row[c], __NV_E4M3);
            v = __half2float(*reinterpret_cast<const __half*>(&h)) * scales[first + r];
        }
        out[t] = v;
    }
}

__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int64_t)in[i];
}
__global__ void list_sizes_kernel(const int64_t* __restrict__ offsets, int nlist, int64_t* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nlist) out[c] = offsets[c + 1] - offsets[c];
}

__global__ void add_list_sizes_kernel(const int64_t* __restrict__ offsets, int nlist, int64_t* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nlist) out[c] += offsets[c + 1] - offsets[c];
}

// ---------------------------------------------------------------------------------- K4c rescore
// One CTA per query: candidates (list positions) -> corpus rows -> exact fp32-query scores in K2's
// summation order -> bitonic sort -> top-k.
// The body is a device function run by ALL threads of a CTA for query q: the stand-alone kernel below wraps it, and
// the list scan's last CTA calls it directly for small batches (one launch fewer on the single-query latency path).
// Needs blockDim.x >= P (the candidates padded to a power of two >= 64) and P * 8 bytes of shared memory at rs_buf.
template <int ELEM>
__device__ __forceinline__ void ivf_rescore_body(const uint64_t* __restrict__ cand, int kc, int k, int q,
                                                 const uint32_t* __restrict__ list_rows,
                                                 const float* __restrict__ q32,
                                                 const uint8_t* __restrict__ corpus, uint32_t row_bytes,
                                                 int dim_pad, const int64_t* __restrict__ id_map,
                                                 uint64_t* __restrict__ out_keys, float* __restrict__ out_scores,
                                                 int64_t* __restrict__ out_ids, uint64_t* rs_buf) {
    constexpr int CN = Chunk<ELEM>::N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    int P = 64;
    while (P < kc) P <<= 1;
    const float* qv = q32 + (size_t)q * dim_pad;
    // warp w owns candidates w, w + nwarps, ...; lane i looks up the i-th of them (key -> list position ->
    // corpus row), so all of a warp's dependent lookups are in flight together; rows are then scored four at
    // a time (their loads are independent and overlap), each by the whole warp in K2's summation order.
    {
        const int jm = warp + lane * nwarps;
        const uint64_t mykey = (jm < kc) ? cand[(size_t)q * kc + jm] : 0ull;
        const uint32_t myrow = mykey ? list_rows[key_row(mykey)] : 0u;
        const unsigned live = __ballot_sync(0xFFFFFFFFu, mykey != 0ull);
        uint64_t outkey = 0ull;
        for (int c0 = 0; c0 < 32 && warp + c0 * nwarps < P; c0 += 4) {
            if (((live >> c0) & 0xFu) == 0u) continue;
            uint32_t row[4];
            const uint8_t* r[4];
            float acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                row[u] = __shfl_sync(0xFFFFFFFFu, myrow, (c0 + u) & 31);   // dead slots read row 0: harmless
                r[u] = corpus + (size_t)row[u] * row_bytes;
                acc[u] = 0.f;
            }
#pragma unroll 2
            for (uint32_t off = (uint32_t)lane * 16u; off < row_bytes; off += 512u) {   // K2's chunk order
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(r[u] + off));
                float ql[CN];
#pragma unroll
                for (int i = 0; i < CN; i += 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(qv + off / ELEM + i));
                    ql[i] = f.x;
                    ql[i + 1] = f.y;
                    ql[i + 2] = f.z;
                    ql[i + 3] = f.w;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = Chunk<ELEM>::dot(v[u], ql, acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xFFFFFFFFu, acc[u], o);   // == K2's reduce tree
                if (lane == c0 + u && ((live >> (c0 + u)) & 1u)) outkey = pack_key(acc[u], row[u]);
            }
        }
        if (jm < P) rs_buf[jm] = outkey;
    }
    __syncthreads();
    // rank by counting: P <= 256 keys, one per thread, unique (distinct rows) or 0 — a key's rank is the number
    // of larger keys; 128 x 128 compares cost a fraction of a 28-level bitonic network's barriers
    {
        const int i = threadIdx.x;
        const uint64_t key = (i < P) ? rs_buf[i] : 0ull;
        const int nnz = __syncthreads_count(key != 0ull);
        if (key != 0ull) {
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < P; ++j) rank += (rs_buf[j] > key) ? 1 : 0;
            if (rank < k) {
                const size_t o = (size_t)q * k + rank;
                if (out_keys) out_keys[o] = key;
                if (out_scores) out_scores[o] = key_score(key);
                if (out_ids) {
                    const uint32_t row = key_row(key);
                    out_ids[o] = id_map ? id_map[row] : (int64_t)row;
                }
            }
        }
        for (int r = nnz + i; r < k; r += blockDim.x) {   // fewer eligible rows than k: padding
            const size_t o = (size_t)q * k + r;
            if (out_keys) out_keys[o] = 0ull;
            if (out_scores) out_scores[o] = -INFINITY;
            if (out_ids) out_ids[o] = -1;
        }
    }
}

template <int ELEM>
__global__ void __launch_bounds__(1024) ivf_rescore_kernel(const uint64_t* __restrict__ cand, int kc, int k,
                                                           const uint32_t* __restrict__ list_rows,
                                                           const float* __restrict__ q32,
                                                           const uint8_t* __restrict__ corpus, uint32_t row_bytes,
                                                           int dim_pad, const int64_t* __restrict__ id_map,
                                                           uint64_t* __restrict__ out_keys, float* __restrict__ out_scores,
                                                           int64_t* __restrict__ out_ids) {
    extern __shared__ uint64_t rs_buf_dyn[];
    ivf_rescore_body<ELEM>(cand, kc, k, (int)blockIdx.x, list_rows, q32, corpus, row_bytes, dim_pad, id_map, out_keys,
                           out_scores, out_ids, rs_buf_dyn);
}

// ---------------------------------------------------------------------------------- K4b list scan
struct ListScanParams {
    const uint8_t* list_data;     // [n, row_bytes] rows in list order
    uint32_t row_bytes;           // multiple of 16
    int dim_pad;                  // valid fp32 query elements
    const float* scales;          // [n] per-row scale (fp8 lists) or nullptr
    const int64_t* list_offsets;  // [nlist + 1]
    const uint64_t* probes;       // [nq, nprobe] packed (score, list) keys from the coarse step; 0 = none
    int nprobe;
    const float* queries;         // [nq, dim_pad] fp32 normalised
    const uint32_t* mask;         // allow bitmask over CORPUS rows (SQL WHERE before ORDER BY/LIMIT) or nullptr
    const uint32_t* list_rows;    // [n] corpus row of every list position (only read when mask != nullptr or has_dead)
    int has_dead;                 // some positions are tombstones (list_rows[pos] == TS_DEAD_ROW): skip them
    int k;                        // candidates kept per query (rescore_k)
    uint64_t* part_keys;          // [nq][gridDim.x][k]
    uint32_t* tickets;            // [nq] zero on entry, left zero
    uint64_t* out_keys;           // [nq][k] merged candidates, key row = list POSITION
    const uint32_t* run_flag;     // non-null: run only if *run_flag != 0 (fallback of the list-major path K4d)
    int stages;
    unsigned long long* timeline; // diagnostics: [gridDim.x][8] globaltimer stamps of query 0, or nullptr
    // fused exact re-score (K4c inside the last CTA of the query; small batches): k_out > 0 enables it
    int k_out;                    // final k (0 = stop at the candidates, a separate K4c launch follows)
    const uint8_t* corpus;        // stored bf16 rows
    uint32_t corpus_row_bytes;
    const int64_t* id_map;
    uint64_t* fin_keys;           // [nq][k_out] any of the three may be nullptr
    float* fin_scores;
    int64_t* fin_ids;
};

__device__ unsigned long long g_ivf_timeline[148 * 12];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define IVF_STAMP(i)                                                                         \
    do {                                                                                     \
        if (p.timeline != nullptr && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 148) \
            p.timeline[blockIdx.x * 12 + (i)] = globaltimer_ns();                            \
    } while (0)

template <int ELEM, int NCHUNK, int KPL, int R>
__global__ void __launch_bounds__(512, 1) list_scan_kernel(const ListScanParams p) {
    constexpr int CN = Chunk<ELEM>::N;
    constexpr int GROUP = 32 / R;
    extern __shared__ __align__(128) uint8_t smem[];
    if (p.run_flag != nullptr && *p.run_flag == 0u) return;   // the list-major path did the work

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const int stages = p.stages;
    const uint32_t tile_bytes = R * p.row_bytes;
    const int qi = blockIdx.y;
    const int nprobe = p.nprobe;

    uint8_t* my_slots = smem + (size_t)warp * stages * tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)W * stages * tile_bytes);
    uint64_t* my_bars = bars + warp * stages;
    // probed-list table: first position, length, tiles before this list
    int64_t* s_start = reinterpret_cast<int64_t*>(bars + W * stages);
    int* s_len = reinterpret_cast<int*>(s_start + nprobe);
    int* s_tpref = s_len + nprobe;   // [nprobe + 1]

    IVF_STAMP(0);
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&my_bars[s], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        int carry = 0;
        for (int j0 = 0; j0 < nprobe; j0 += 32) {
            const int j = j0 + lane;
            int len = 0;
            int64_t start = 0;
            if (j < nprobe) {
                const uint64_t key = p.probes[(size_t)qi * nprobe + j];
                if (key != 0ull) {
                    const uint32_t l = key_row(key);
                    start = p.list_offsets[l];
                    len = (int)(p.list_offsets[l + 1] - start);
                }
                s_start[j] = start;
                s_len[j] = len;
            }
            int t = (len + R - 1) / R;
            int inc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += u;
            }
            if (j < nprobe) s_tpref[j] = carry + inc - t;
            carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) s_tpref[nprobe] = carry;
    }
    __syncthreads();
    IVF_STAMP(1);

    const int64_t T = s_tpref[nprobe];
    const int64_t g = (int64_t)blockIdx.x * W + warp;
    const int64_t G = (int64_t)gridDim.x * W;
    const int t0 = (int)(T * g / G), t1 = (int)(T * (g + 1) / G);
    const uint64_t policy = l2_policy_evict_first();
    const int k = p.k;
    const int my_row = row_of_lane<R>(lane);
    const bool leader = (lane & (GROUP - 1)) == 0;

    // cursor: (list slot j, tile tt inside it); both cursors start at tile t0
    int jc = 0;
    if (t0 < t1) {
        int lo = 0, hi = nprobe - 1;   // largest j with s_tpref[j] <= t0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_tpref[mid] <= t0) lo = mid; else hi = mid - 1;
        }
        jc = lo;
        while (s_tpref[jc + 1] <= t0) ++jc;   // skip empty lists that share the prefix value
    }
    int tc = t0 - s_tpref[jc];
    int ji = jc, ti = tc;
    auto advance = [&](int& j, int& tt) {
        ++tt;
        while (j < nprobe && tt >= s_tpref[j + 1] - s_tpref[j]) {
            ++j;
            tt = 0;
        }
    };
    auto issue = [&](int j, int tt, int s) {
        const int64_t pos0 = s_start[j] + (int64_t)tt * R;
        const int left = s_len[j] - tt * R;
        const uint32_t bytes = (uint32_t)(left < R ? left : R) * p.row_bytes;
        mbar_expect_tx(&my_bars[s], bytes);
        tma_load_1d_hint(my_slots + (size_t)s * tile_bytes, p.list_data + (size_t)pos0 * p.row_bytes, bytes, &my_bars[s],
                         policy);
    };
    {
        int it = t0;
        for (int s = 0; s < stages && it < t1; ++s, ++it) {
            if (lane == 0) issue(ji, ti, s);
            advance(ji, ti);
        }
    }

    // query slice of this lane: fp32 for bf16 lists, packed half2 for e4m3 lists (see dot16_e4m3_h2)
    constexpr int QF = (ELEM == 1) ? 1 : NCHUNK * CN;
    constexpr int QH = (ELEM == 1) ? NCHUNK * CN / 2 : 1;
    float q[QF];
    __half2 qh[QH];
    {
        const float* qv = p.queries + (size_t)qi * p.dim_pad;
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            const int e0 = (j * 32 + lane) * CN;
            // dim_pad is a multiple of 8 and rows of the prepared query array are 32-byte aligned: float4 loads
            if constexpr (ELEM == 1) {
#pragma unroll
                for (int i = 0; i < CN; i += 4) {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e0 + i < p.dim_pad) f = __ldg(reinterpret_cast<const float4*>(qv + e0 + i));
                    qh[(j * CN + i) / 2] = __floats2half2_rn(f.x, f.y);
                    qh[(j * CN + i) / 2 + 1] = __floats2half2_rn(f.z, f.w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < CN; i += 4) {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e0 + i < p.dim_pad) f = __ldg(reinterpret_cast<const float4*>(qv + e0 + i));
                    q[j * CN + i] = f.x;
                    q[j * CN + i + 1] = f.y;
                    q[j * CN + i + 2] = f.z;
                    q[j * CN + i + 3] = f.w;
                }
            }
        }
    }

    WarpSelect<KPL> sel;
    sel.init();
    int s = 0;
    uint32_t parity = 0;
    IVF_STAMP(2);
    for (int t = t0; t < t1; ++t) {
        const int left = s_len[jc] - tc * R;
        const int64_t pos = s_start[jc] + (int64_t)tc * R + my_row;
        bool mine = leader && my_row < left;
        if ((p.mask != nullptr || p.has_dead) && mine) {   // position -> corpus row -> tombstone? -> allow bit
            const uint32_t row = __ldg(p.list_rows + pos);
            mine = row != TS_DEAD_ROW;
            if (mine && p.mask != nullptr) mine = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;
        }
        float scale = 1.0f;
        if (ELEM == 1 && mine) scale = __ldg(p.scales + pos);

        mbar_wait(&my_bars[s], parity);
        const uint8_t* slot = my_slots + (size_t)s * tile_bytes;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            const uint32_t off = (uint32_t)(j * 32 + lane) * 16u;
            if (off < p.row_bytes) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint4 v = *reinterpret_cast<const uint4*>(slot + (size_t)r * p.row_bytes + off);
                    if constexpr (ELEM == 1) acc[r] = dot16_e4m3_h2(v, &qh[j * CN / 2], acc[r]);
                    else acc[r] = Chunk<ELEM>::dot(v, &q[j * CN], acc[r]);
                }
            }
        }
        __syncwarp();
        if (t + stages < t1) {
            if (lane == 0) issue(ji, ti, s);
            advance(ji, ti);
        }
        advance(jc, tc);

        transpose_reduce<R>(acc, lane);
        sel.offer(mine ? pack_key(acc[0] * scale, (uint32_t)pos) : 0ull, k, lane);
        if (++s == stages) {
            s = 0;
            parity ^= 1u;
        }
    }
    IVF_STAMP(3);

    // ---- CTA selection. Warps do NOT sort their own candidates: every warp appends its (sorted) best list
    // and its unsorted pending keys to one dense smem array, and the whole CTA sorts that array once with a
    // shared-memory bitonic network (all threads on one sort instead of every warp sorting a padded list and
    // one warp folding them serially: 2-3 us instead of 40+ when a warp has seen only tens of rows).
    __syncthreads();                                          // every slot has been consumed
    uint64_t* cbuf = reinterpret_cast<uint64_t*>(smem);       // aliases the drained TMA slots
    __shared__ int s_n, s_is_last;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    {
        const uint32_t lt_mask = (1u << lane) - 1u;
        int nb = 0;                                           // non-empty entries of the sorted list (a prefix)
#pragma unroll
        for (int j = 0; j < KPL; ++j) nb += __popc(__ballot_sync(0xFFFFFFFFu, sel.best.key[j] != 0ull));
        const int n_w = nb + sel.count;
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_n, n_w);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
        for (int j = 0; j < KPL; ++j)
            if (j * 32 + lane < nb) cbuf[base + j * 32 + lane] = sel.best.key[j];
        int off = base + nb;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            const unsigned msk = __ballot_sync(0xFFFFFFFFu, sel.pend[j] != 0ull);
            if (sel.pend[j] != 0ull) cbuf[off + __popc(msk & lt_mask)] = sel.pend[j];
            off += __popc(msk);
        }
    }
    __syncthreads();
    IVF_STAMP(4);
    {
        const int n = s_n;
        int P = 64;
        while (P < n) P <<= 1;
        for (int i = n + threadIdx.x; i < P; i += blockDim.x) cbuf[i] = 0ull;
        __syncthreads();
        cta_bitonic_sort_desc(cbuf, P);
        uint64_t* out = p.part_keys + ((size_t)qi * gridDim.x + blockIdx.x) * k;
        for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = (i < P) ? cbuf[i] : 0ull;
    }
    // publish: the barrier orders every thread's stores before thread 0's gpu-scope fence (fences are
    // cumulative), so ONE fence suffices — 512 threads fencing at once cost several microseconds
    __syncthreads();
    IVF_STAMP(5);
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(p.tickets + qi, 1u);
        s_is_last = (t == gridDim.x - 1);
        if (s_is_last) {
            p.tickets[qi] = 0u;
            __threadfence();   // acquire side: the other CTAs' lists are read (from L2) after this
        }
    }
    __syncthreads();
    if (!s_is_last) return;

    // ---- final merge by the last CTA of the query: gridDim.x sorted lists of k keys. Prune first with a lower
    // bound L on the k-th largest key overall: for any j, the j-th largest of the lists' entries at position
    // ceil(k/j)-1 is such a bound (j lists hold at least ceil(k/j) keys >= it). j = k is "the k-th largest
    // head" (good when the best keys are spread over the lists), j = 1 is "the largest k-th entry" (good when
    // they sit in a few lists, e.g. the CTAs that scanned the query's own cluster); powers of four in between.
    // Only keys >= L can be in the result — typically a few hundred of the lists*k keys.
    IVF_STAMP(6);
    const int nl = gridDim.x;
    const uint64_t* mine = p.part_keys + (size_t)qi * nl * k;
    constexpr int CAP = 4096;                                 // keys; fits the smallest smem carve-out (32 KB)
    constexpr int MAXJ = 12;
    __shared__ int s_cnt, s_nj;
    __shared__ int s_j[MAXJ], s_pos[MAXJ];
    __shared__ unsigned long long s_low;
    if (threadIdx.x == 0) {
        int nj = 0;
        for (int j = 1; j < nl && j < k && nj < MAXJ - 1; j <<= 2) {   // 1, 4, 16, 64: each choice costs lists^2 compares
            s_j[nj] = j;
            s_pos[nj++] = (k + j - 1) / j - 1;
        }
        const int jl = k < nl ? k : nl;                       // the widest choice: k heads, or every list
        s_j[nj] = jl;
        s_pos[nj++] = (k + jl - 1) / jl - 1;
        s_nj = nj;
        s_cnt = 0;
        s_low = 0ull;                                         // stays 0 when no bound applies: every key survives
    }
    __syncthreads();
    {
        const int nj = s_nj;
        uint64_t* hv = cbuf;                                  // [nj][nl]
        for (int i = threadIdx.x; i < nj * nl; i += blockDim.x)
            hv[i] = __ldcg(mine + (size_t)(i % nl) * k + s_pos[i / nl]);
        __syncthreads();
        for (int i = threadIdx.x; i < nj * nl; i += blockDim.x) {
            const uint64_t key = hv[i];
            if (key == 0ull) continue;
            const uint64_t* row = hv + (size_t)(i / nl) * nl;
            int rank = 0;
#pragma unroll 4
            for (int l = 0; l < nl; ++l) rank += (row[l] > key) ? 1 : 0;
            if (rank == s_j[i / nl] - 1) atomicMax(&s_low, (unsigned long long)key);
        }
        __syncthreads();
    }
    IVF_STAMP(8);
    const uint64_t low = s_low;
    uint64_t* sel_buf = cbuf;                                 // the head sort is finished with cbuf (barrier above)
    for (int i0 = threadIdx.x; i0 < nl * k; i0 += 32 * blockDim.x) {   // up to 32 L2 loads in flight per thread
        uint64_t v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const int i = i0 + u * blockDim.x;
            v[u] = (i < nl * k) ? __ldcg(mine + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (v[u] != 0ull && v[u] >= low) {
                const int pos = atomicAdd(&s_cnt, 1);
                if (pos < CAP) sel_buf[pos] = v[u];
            }
        }
    }
    __syncthreads();
    IVF_STAMP(9);
    const int cnt = s_cnt;
    uint64_t* dst = p.out_keys + (size_t)qi * k;
    if (cnt <= 384) {
        // few survivors (the usual case): rank by counting; the rank is the output position
        for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = 0ull;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const uint64_t key = sel_buf[i];
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < cnt; ++j) rank += (sel_buf[j] > key) ? 1 : 0;
            if (rank < k) dst[rank] = key;
        }
        IVF_STAMP(10);
    } else if (cnt <= CAP) {
        int P = 64;
        while (P < cnt) P <<= 1;
        for (int i = cnt + threadIdx.x; i < P; i += blockDim.x) sel_buf[i] = 0ull;
        __syncthreads();
        cta_bitonic_sort_desc(sel_buf, P);
        for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = (i < P) ? sel_buf[i] : 0ull;
    } else {
        // more survivors than the buffer holds (lists with long runs of near-equal keys): the register
        // merge is slower but needs no bound
        __syncthreads();
        MergeParams mp;
        mp.keys = p.part_keys;
        mp.nlists = nl;
        mp.nq = gridDim.y;
        mp.k = k;
        mp.stride_list = k;
        mp.stride_query = (int64_t)nl * k;
        mp.list_base = nullptr;
        mp.id_map = nullptr;
        mp.out_keys = p.out_keys;
        mp.out_scores = nullptr;
        mp.out_ids = nullptr;
        mp.out_stride = k;
        mp.qlist = nullptr;
        mp.qcount = nullptr;
        mp.done_flag = nullptr;
        mp.done_value = 0;
        merge_lists<KPL>(mp, qi, qi, cbuf, W);
    }
    IVF_STAMP(7);
    if (p.k_out > 0) {
        // exact re-score of the k candidates by this CTA (they were written to dst by this CTA: visible after the barrier)
        __syncthreads();
        ivf_rescore_body<2>(p.out_keys, k, p.k_out, qi, p.list_rows, p.queries, p.corpus, p.corpus_row_bytes, p.dim_pad,
                            p.id_map, p.fin_keys, p.fin_scores, p.fin_ids, cbuf);
    }
}

// A query spread over several CTAs whose warps would each see only a few hundred rows is latency-bound
// (short per-warp instruction streams matter more than bytes in flight); long scans — many rows per warp,
// e.g. the one-list fp8 shadow of the whole corpus — stream best with K2's configuration.
static bool latency_mode(const ts_index* ix, int nprobe, int parts) {
    if (parts <= 1) return false;
    const double rows = (double)nprobe * (double)ix->size / (double)std::max(ix->nlist, 1);
    return rows / ((double)parts * 16.0) < 512.0;
}

struct ListScanConfig {
    int warps, stages;
    size_t smem;
};

template <int ELEM, int NCHUNK, int KPL, int R>
static int launch_list_scan_r(const ts_index* ix, ListScanParams p, int nq, int parts, cudaStream_t s) {
    const Tunables& t = tunables();
    const size_t tile_bytes = (size_t)R * p.row_bytes;
    int stages = t.scan_stages < 2 ? 2 : t.scan_stages;
    // latency mode (a query is spread over several CTAs): 16 warps per SM halve the rows — and so the serial
    // instruction stream — per warp; throughput mode (one CTA per query, many queries): 8 warps stream at HBM rate
    int warps = t.ivf_warps > 0 ? t.ivf_warps : (latency_mode(ix, p.nprobe, parts) ? 16 : 8);
    warps = warps > 16 ? 16 : warps;
    const size_t table = (size_t)p.nprobe * 16 + 16;   // s_start, s_len, s_tpref
    const size_t budget = (size_t)(220 * 1024) - 1024 - table;
    while (warps > 1 && (size_t)warps * stages * tile_bytes + 8 * warps * stages > budget) --warps;
    size_t smem = (size_t)warps * stages * tile_bytes + 8 * (size_t)warps * stages + table;
    // the epilogue reuses the slot area: every warp's sorted list + pending keys, and the 4096-key final buffer
    const size_t select_bytes = std::max<size_t>((size_t)warps * 2 * KPL * 32 * 8, 4096 * 8);
    if (smem < select_bytes) smem = select_bytes;
    p.stages = stages;
    if (p.k_out > 0) {
        int P = 64;
        while (P < p.k) P <<= 1;
        TS_REQUIRE(warps * 32 >= P, TS_ERR_UNSUPPORTED, "ivf list scan: %d threads cannot re-score %d candidates in place "
                   "(set tunable ivf.fuse_rescore = 0)", warps * 32, p.k);
    }
    p.timeline = nullptr;
    if (t.ivf_timeline) TS_CHECK_CUDA(cudaGetSymbolAddress((void**)&p.timeline, g_ivf_timeline));
    auto kern = list_scan_kernel<ELEM, NCHUNK, KPL, R>;
    TS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(parts, nq), warps * 32, smem, s>>>(p);
    TS_LAUNCH_CHECK();
    (void)ix;
    return TS_OK;
}

template <int ELEM, int NCHUNK>
static int launch_list_scan_k(const ts_index* ix, const ListScanParams& p, int nq, int parts, cudaStream_t s) {
    constexpr int R = (ELEM == 1) ? (NCHUNK == 1 ? 16 : (NCHUNK == 2 ? 8 : 4)) : RowsPerTile<NCHUNK>::value;
    if constexpr (ELEM == 1 && NCHUNK == 2) {   // the flagship shape (e4m3, 512 < D <= 1024): 4-row tiles too
        int rows = tunables().ivf_tile_rows;
        if (rows == 0) rows = latency_mode(ix, p.nprobe, parts) ? 4 : 8;
        if (rows == 4 && p.k <= 128)
            return p.k <= 32 ? launch_list_scan_r<ELEM, NCHUNK, 1, 4>(ix, p, nq, parts, s)
                             : launch_list_scan_r<ELEM, NCHUNK, 4, 4>(ix, p, nq, parts, s);
    }
    if (p.k <= 32) return launch_list_scan_r<ELEM, NCHUNK, 1, R>(ix, p, nq, parts, s);
    if (p.k <= 128) return launch_list_scan_r<ELEM, NCHUNK, 4, R>(ix, p, nq, parts, s);
    if (p.k <= 256) return launch_list_scan_r<ELEM, NCHUNK, 8, R>(ix, p, nq, parts, s);
    set_error("ivf list scan: max(k, rescore_k) = %d exceeds %d", p.k, TS_IVF_MAX_CANDIDATES);
    return TS_ERR_UNSUPPORTED;
}

static int launch_list_scan(const ts_index* ix, const ListScanParams& p, int nq, int parts, cudaStream_t s) {
    const int nchunk = (int)((p.row_bytes + 511) / 512);
    if (ix->list_dtype == TS_FP8_E4M3) {
        if (nchunk <= 1) return launch_list_scan_k<1, 1>(ix, p, nq, parts, s);
        if (nchunk <= 2) return launch_list_scan_k<1, 2>(ix, p, nq, parts, s);
        if (nchunk <= 4) return launch_list_scan_k<1, 4>(ix, p, nq, parts, s);
    } else {
        switch (nchunk) {
            case 1: return launch_list_scan_k<2, 1>(ix, p, nq, parts, s);
            case 2: return launch_list_scan_k<2, 2>(ix, p, nq, parts, s);
            case 3: return launch_list_scan_k<2, 3>(ix, p, nq, parts, s);
            case 4: return launch_list_scan_k<2, 4>(ix, p, nq, parts, s);
            default:
                if (nchunk <= 8) return launch_list_scan_k<2, 8>(ix, p, nq, parts, s);
        }
    }
    set_error("ivf list scan: no kernel for list dtype %d dim %d", ix->list_dtype, ix->dim);
    return TS_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------- host: state
static void ivf_free_lists(ts_index* ix) {
    cudaFree(ix->list_offsets);
    cudaFree(ix->list_rows);
    cudaFree(ix->list_data);
    cudaFree(ix->list_scales);
    ix->list_offsets = nullptr;
    ix->list_rows = nullptr;
    ix->list_data = nullptr;
    ix->list_scales = nullptr;
    side_table_destroy(&ix->pos_store, &ix->pos_of_row);
    cudaFree(ix->ovf_set);
    ix->ovf_set = nullptr;
    ix->list_cap = ix->built_n = ix->ovf_n = ix->ivf_dead = ix->ovf_cap = ix->ivf_moved = 0;
    ix->ivf_built = false;
}
static void ivf_free_centroids(ts_index* ix) {
    cudaFree(ix->centroids);
    cudaFree(ix->centroids_bf16);
    cudaFree(ix->centroid_max_norm2);
    ix->centroids = nullptr;
    ix->centroids_bf16 = nullptr;
    ix->centroid_max_norm2 = nullptr;
    ix->nlist = 0;
}
static int ivf_alloc_centroids(ts_index* ix, int nlist) {
    ivf_free_lists(ix);
    ivf_free_centroids(ix);
    TS_CHECK_CUDA(cudaMalloc(&ix->centroids, (size_t)nlist * ix->dim_pad * sizeof(float)));
    TS_CHECK_CUDA(cudaMalloc(&ix->centroids_bf16, (size_t)nlist * ix->dim_pad * sizeof(__nv_bfloat16)));
    TS_CHECK_CUDA(cudaMalloc(&ix->centroid_max_norm2, sizeof(float)));
    ix->nlist = nlist;
    return TS_OK;
}
static int ivf_finish_centroids(ts_index* ix, cudaStream_t s) {
    TS_CHECK_CUDA(cudaMemsetAsync(ix->centroid_max_norm2, 0, sizeof(float), s));
    max_norm2_bf16_kernel<<<(ix->nlist + 7) / 8, 256, 0, s>>>(ix->centroids_bf16, ix->nlist, ix->dim_pad,
                                                              ix->centroid_max_norm2);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

// temp device buffers of the train/build calls (build-time only; searches never allocate)
struct TempBufs {
    std::vector<void*> ptrs;
    ~TempBufs() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    int get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) {
            set_error("ivf: cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return TS_ERR_OOM;
        }
        ptrs.push_back(p);
        *out = (T*)p;
        return TS_OK;
    }
};

// assign -> (members sorted by (list, row), offsets). Stable LSD radix sort on the list id keeps rows ascending.
static int sort_by_list(const uint32_t* assign, int64_t n, int nlist, uint32_t* sorted_assign, uint32_t* rows_in,
                        uint32_t* members, int64_t* offsets, TempBufs& tmp, cudaStream_t s) {
    TS_REQUIRE(n < (int64_t)1 << 31, TS_ERR_UNSUPPORTED, "ivf: more than 2^31-1 rows in one sort");
    iota_u32_kernel<<<1024, 256, 0, s>>>(rows_in, n);
    TS_LAUNCH_CHECK();
    int end_bit = 1;
    while ((1 << end_bit) < nlist) ++end_bit;
    size_t tbytes = 0;
    TS_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, assign, sorted_assign, rows_in, members, (int)n, 0,
                                                  end_bit, s));
    uint8_t* t = nullptr;
    int rc = tmp.get(&t, tbytes);
    if (rc) return rc;
    TS_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(t, tbytes, assign, sorted_assign, rows_in, members, (int)n, 0, end_bit,
                                                  s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    list_offsets_kernel<<<(nlist + 1 + 255) / 256, 256, 0, s>>>(sorted_assign, n, nlist, offsets);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

// ---------------------------------------------------------------------------------- incremental upkeep
__global__ void invert_rows_kernel(const uint32_t* __restrict__ list_rows, int64_t n, int64_t first_pos,
                                   uint32_t* __restrict__ pos_of_row) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t row = list_rows[first_pos + i];
        if (row != TS_DEAD_ROW) pos_of_row[row] = (uint32_t)(first_pos + i);
    }
}
// Replaced rows: the main-list position holding the old content dies; the row joins the overflow set (rows already
// in the overflow just stay there — their content is re-gathered below).
__global__ void tombstone_kernel(const uint32_t* __restrict__ replaced, int64_t n, uint32_t* __restrict__ pos_of_row,
                                 int64_t built_n, uint32_t* __restrict__ list_rows, uint32_t* __restrict__ ovf_set,
                                 unsigned int* __restrict__ moved) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t row = replaced[i];
        const uint32_t pos = pos_of_row[row];
        if ((int64_t)pos < built_n && list_rows[pos] == row) {
            list_rows[pos] = TS_DEAD_ROW;
            ovf_set[atomicAdd(moved, 1u)] = row;
        }
    }
}
__global__ void iota_from_kernel(uint32_t* out, int64_t n, uint32_t first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = first + (uint32_t)i;
}
__global__ void gather_rows_kernel(const uint8_t* __restrict__ corpus, uint32_t row_bytes, const uint32_t* __restrict__ rows,
                                   int64_t n, uint8_t* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (int64_t)gridDim.x * 8) {
        const uint8_t* src = corpus + (size_t)rows[i] * row_bytes;
        for (uint32_t off = lane * 16u; off < row_bytes; off += 512u)
            *reinterpret_cast<uint4*>(dst + (size_t)i * row_bytes + off) = __ldg(reinterpret_cast<const uint4*>(src + off));
    }
}
// members[j] indexes the (ascending) overflow set: overflow position j holds corpus row ovf_set[members[j]]
__global__ void overflow_rows_kernel(const uint32_t* __restrict__ ovf_set, const uint32_t* __restrict__ members, int64_t m,
                                     int64_t built_n, uint32_t* __restrict__ list_rows, uint32_t* __restrict__ pos_of_row) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t row = ovf_set[members[j]];
        list_rows[built_n + j] = row;
        pos_of_row[row] = (uint32_t)(built_n + j);
    }
}
__global__ void overflow_offsets_kernel(const int64_t* __restrict__ tmp_offsets, int nlist, int64_t built_n,
                                        int64_t* __restrict__ list_offsets) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l <= nlist) list_offsets[nlist + l] = built_n + tmp_offsets[l];
}
// probe (score, list l) -> probes (score, l) and (score, nlist + l): the main list and its overflow list
__global__ void expand_probes_kernel(const uint64_t* __restrict__ in, int64_t n, int nlist, uint64_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = in[i];
        out[2 * i] = key;
        out[2 * i + 1] = key ? pack_key(key_score(key), key_row(key) + (uint32_t)nlist) : 0ull;
    }
}

template <typename T>
static int grow_buffer(T** buf, size_t old_count, size_t new_count) {
    T* fresh = nullptr;
    TS_CHECK_CUDA(cudaMalloc(&fresh, std::max<size_t>(new_count, 1) * sizeof(T)));
    if (*buf != nullptr && old_count > 0) {
        cudaError_t e = cudaMemcpy(fresh, *buf, old_count * sizeof(T), cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) {
            cudaFree(fresh);
            TS_CHECK_CUDA(e);
        }
    }
    cudaFree(*buf);
    *buf = fresh;
    return TS_OK;
}

// ovf_set[0, m) holds the corpus rows that live in the overflow lists (any order): sort them, file them under the
// existing centroids, quantise them behind the main lists and publish the overflow offsets.
static int ivf_rebuild_overflow(ts_index* ix, int64_t m, TempBufs& tmp, cudaStream_t s) {
    int rc;
    if (m == 0) {
        if (ix->ovf_n > 0) {     // the overflow emptied (deletes): every overflow list is [built_n, built_n)
            int64_t* zeros = nullptr;
            if ((rc = tmp.get(&zeros, (size_t)ix->nlist + 1))) return rc;
            TS_CHECK_CUDA(cudaMemsetAsync(zeros, 0, ((size_t)ix->nlist + 1) * sizeof(int64_t), s));
            overflow_offsets_kernel<<<(ix->nlist + 1 + 255) / 256, 256, 0, s>>>(zeros, ix->nlist, ix->built_n, ix->list_offsets);
            TS_LAUNCH_CHECK();
            TS_CHECK_CUDA(cudaStreamSynchronize(s));
            ix->ovf_n = 0;
        }
        return TS_OK;
    }
    TS_REQUIRE(m < (int64_t)1 << 31, TS_ERR_UNSUPPORTED, "ivf: more than 2^31-1 overflow rows");
    // 2. the overflow set in ascending row order (so rows ascend inside every overflow list, like the main lists)
    uint32_t* sorted_set = nullptr;
    if ((rc = tmp.get(&sorted_set, (size_t)m))) return rc;
    {
        size_t tbytes = 0;
        TS_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tbytes, ix->ovf_set, sorted_set, (int)m, 0, 32, s));
        uint8_t* t = nullptr;
        if ((rc = tmp.get(&t, tbytes))) return rc;
        TS_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(t, tbytes, ix->ovf_set, sorted_set, (int)m, 0, 32, s));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        TS_CHECK_CUDA(cudaMemcpyAsync(ix->ovf_set, sorted_set, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    // 3. file them under the existing centroids: gather -> K4a -> (list, row) order
    uint8_t* rows = nullptr;
    uint32_t *assign = nullptr, *sorted_assign = nullptr, *iota = nullptr, *members = nullptr;
    int64_t* offsets = nullptr;
    if ((rc = tmp.get(&rows, (size_t)m * ix->row_bytes())) || (rc = tmp.get(&assign, (size_t)m)) ||
        (rc = tmp.get(&sorted_assign, (size_t)m)) || (rc = tmp.get(&iota, (size_t)m)) || (rc = tmp.get(&members, (size_t)m)) ||
        (rc = tmp.get(&offsets, (size_t)ix->nlist + 1)))
        return rc;
    const int gblocks = (int)std::min<int64_t>((m + 7) / 8, 148 * 16);
    gather_rows_kernel<<<gblocks, 256, 0, s>>>((const uint8_t*)ix->data, (uint32_t)ix->row_bytes(), ix->ovf_set, m, rows);
    TS_LAUNCH_CHECK();
    rc = launch_assign(ix->device, rows, m, ix->row_bytes(), ix->centroids_bf16, ix->nlist, ix->dim_pad, assign, nullptr, s);
    if (rc) return rc;
    rc = sort_by_list(assign, m, ix->nlist, sorted_assign, iota, members, offsets, tmp, s);
    if (rc) return rc;
    overflow_rows_kernel<<<(int)std::min<int64_t>((m + 255) / 256, 1024), 256, 0, s>>>(ix->ovf_set, members, m, ix->built_n,
                                                                                      ix->list_rows, ix->pos_of_row);
    TS_LAUNCH_CHECK();
    if (ix->list_dtype == TS_BF16)
        gather_quantize_kernel<2><<<gblocks, 256, 0, s>>>((const uint8_t*)ix->data, (uint32_t)ix->row_bytes(),
                                                          ix->list_rows + ix->built_n, m, ix->dim_pad, ix->list_row_bytes,
                                                          (uint8_t*)ix->list_data + (size_t)ix->built_n * ix->list_row_bytes,
                                                          nullptr);
    else
        gather_quantize_kernel<1><<<gblocks, 256, 0, s>>>((const uint8_t*)ix->data, (uint32_t)ix->row_bytes(),
                                                          ix->list_rows + ix->built_n, m, ix->dim_pad, ix->list_row_bytes,
                                                          (uint8_t*)ix->list_data + (size_t)ix->built_n * ix->list_row_bytes,
                                                          ix->list_scales + ix->built_n);
    TS_LAUNCH_CHECK();
    overflow_offsets_kernel<<<(ix->nlist + 1 + 255) / 256, 256, 0, s>>>(offsets, ix->nlist, ix->built_n, ix->list_offsets);
    TS_LAUNCH_CHECK();
    TS_CHECK_CUDA(cudaStreamSynchronize(s));   // temporaries are freed on return
    ix->ovf_n = m;
    return TS_OK;
}

void ivf_free_all(ts_index* ix) {
    ivf_free_lists(ix);
    ivf_free_centroids(ix);
}

int ivf_apply_mutation(ts_index* ix, const uint32_t* replaced_rows, int64_t n_replaced, int64_t app_first, int64_t app_n,
                       cudaStream_t s) {
    if (!ix->ivf_built || (n_replaced == 0 && app_n == 0)) return TS_OK;
    // Past a tenth of the corpus in tombstones + overflow the lists are re-packed from scratch (one K4a pass over
    // all rows: ~1 s per 40M rows); below that only the overflow lists are rebuilt, O(overflow).
    const int64_t churn = ix->ivf_dead + ix->ovf_n + n_replaced + app_n;
    if (churn > std::max<int64_t>(4096, ix->size / 10)) return ts_ivf_build(ix, ix->list_dtype, s);
    TS_CHECK_CUDA(cudaStreamSynchronize(s));
    const int64_t m_max = ix->ovf_n + n_replaced + app_n;
    if (m_max > ix->ovf_cap) {
        const int64_t cap = std::max<int64_t>(m_max, 2 * ix->ovf_cap);
        int rc = grow_buffer(&ix->ovf_set, (size_t)ix->ovf_n, (size_t)cap);
        if (rc) return rc;
        ix->ovf_cap = cap;
    }
    if (ix->built_n + m_max > ix->list_cap) {
        const int64_t cap = std::max<int64_t>(ix->built_n + m_max, ix->list_cap + ix->list_cap / 8);
        const size_t used = (size_t)(ix->built_n + ix->ovf_n);
        int rc = grow_buffer(&ix->list_rows, used, (size_t)cap);
        if (!rc) rc = grow_buffer((uint8_t**)&ix->list_data, used * ix->list_row_bytes, (size_t)cap * ix->list_row_bytes);
        if (!rc && ix->list_scales) rc = grow_buffer(&ix->list_scales, used, (size_t)cap);
        if (rc) return rc;
        ix->list_cap = cap;
    }
    TempBufs tmp;
    int rc;
    // 1. tombstones; replaced main-list rows join the overflow set
    unsigned int moved = 0;
    if (n_replaced > 0) {
        unsigned int* d_moved = nullptr;
        if ((rc = tmp.get(&d_moved, 1))) return rc;
        TS_CHECK_CUDA(cudaMemsetAsync(d_moved, 0, sizeof(unsigned int), s));
        tombstone_kernel<<<(int)std::min<int64_t>((n_replaced + 255) / 256, 1024), 256, 0, s>>>(
            replaced_rows, n_replaced, ix->pos_of_row, ix->built_n, ix->list_rows, ix->ovf_set + ix->ovf_n, d_moved);
        TS_LAUNCH_CHECK();
        TS_CHECK_CUDA(cudaMemcpyAsync(&moved, d_moved, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
        TS_CHECK_CUDA(cudaStreamSynchronize(s));
        ix->ivf_dead += moved;
    }
    int64_t m = ix->ovf_n + moved;
    if (app_n > 0) {
        iota_from_kernel<<<(int)std::min<int64_t>((app_n + 255) / 256, 1024), 256, 0, s>>>(ix->ovf_set + m, app_n,
                                                                                          (uint32_t)app_first);
        TS_LAUNCH_CHECK();
        m += app_n;
    }
    return ivf_rebuild_overflow(ix, m, tmp, s);
}

// ---- delete by id (ts_index_delete) -------------------------------------------------------------------------
// The row store is compacted: deleted rows vanish and the last rows move into the freed slots. remap[old row] is the
// row's new position (itself for most rows) or TS_DEAD_ROW for a deleted row.
__global__ void remap_main_lists_kernel(uint32_t* __restrict__ list_rows, int64_t built_n, const uint32_t* __restrict__ remap,
                                        uint32_t* __restrict__ pos_of_row, unsigned int* __restrict__ died) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < built_n; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t row = list_rows[p];
        if (row == TS_DEAD_ROW) continue;
        const uint32_t now = remap[row];
        if (now == row) continue;
        list_rows[p] = now;                       // a deleted row's entry becomes a tombstone
        if (now == TS_DEAD_ROW) atomicAdd(died, 1u);
        else pos_of_row[now] = (uint32_t)p;
    }
}
__global__ void remap_overflow_set_kernel(const uint32_t* __restrict__ in, int64_t m, const uint32_t* __restrict__ remap,
                                          uint32_t* __restrict__ out, unsigned int* __restrict__ kept) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t now = remap[in[i]];
        if (now != TS_DEAD_ROW) out[atomicAdd(kept, 1u)] = now;
    }
}

// ix->size and the rows are already the compacted ones. Main-list entries of deleted rows are tombstoned and those
// of moved rows renamed in place; the overflow set is filtered / renamed and its lists rebuilt (O(overflow)).
int ivf_apply_delete(ts_index* ix, const uint32_t* remap, int64_t n_deleted, cudaStream_t s) {
    if (!ix->ivf_built || n_deleted == 0) return TS_OK;
    if (ix->size == 0) {          // nothing left to list: the centroids stay, the lists go (ts_ivf_build after the next add)
        TS_CHECK_CUDA(cudaStreamSynchronize(s));
        ivf_free_lists(ix);
        return TS_OK;
    }
    const int64_t churn = ix->ivf_dead + ix->ovf_n + n_deleted;
    if (churn > std::max<int64_t>(4096, ix->size / 10)) return ts_ivf_build(ix, ix->list_dtype, s);
    TempBufs tmp;
    int rc;
    unsigned int* counters = nullptr;     // [0] main-list entries that died, [1] overflow rows kept
    uint32_t* kept_set = nullptr;
    if ((rc = tmp.get(&counters, 2)) || (rc = tmp.get(&kept_set, (size_t)ix->ovf_n))) return rc;
    TS_CHECK_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), s));
    remap_main_lists_kernel<<<(int)std::min<int64_t>((ix->built_n + 255) / 256, 148 * 8), 256, 0, s>>>(
        ix->list_rows, ix->built_n, remap, ix->pos_of_row, counters);
    TS_LAUNCH_CHECK();
    if (ix->ovf_n > 0) {
        remap_overflow_set_kernel<<<(int)std::min<int64_t>((ix->ovf_n + 255) / 256, 1024), 256, 0, s>>>(
            ix->ovf_set, ix->ovf_n, remap, kept_set, counters + 1);
        TS_LAUNCH_CHECK();
    }
    unsigned int h[2] = {0, 0};
    TS_CHECK_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, s));
    TS_CHECK_CUDA(cudaStreamSynchronize(s));
    ix->ivf_dead += h[0];
    ix->ivf_moved += n_deleted;
    if (ix->ovf_n == 0) return TS_OK;
    TS_CHECK_CUDA(cudaMemcpyAsync(ix->ovf_set, kept_set, (size_t)h[1] * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    return ivf_rebuild_overflow(ix, (int64_t)h[1], tmp, s);
}

}  // namespace ts

using namespace ts;

extern "C" {

// ------------------------------------------------------------------------------------ train
int ts_ivf_train(ts_index* ix, const float* sample, int64_t n_sample, int nlist, int iters, uint64_t seed,
                 void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_train: index is NULL");
    TS_REQUIRE(nlist >= 1 && nlist <= (1 << 20), TS_ERR_BAD_ARG, "ivf_train: nlist=%d out of range [1, 2^20]", nlist);
    TS_REQUIRE(iters >= 0 && iters <= 1000, TS_ERR_BAD_ARG, "ivf_train: iters=%d", iters);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_train: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    TempBufs tmp;
    const uint8_t* rows = nullptr;
    size_t row_stride = 0;
    int64_t n = 0;
    if (sample != nullptr) {
        TS_REQUIRE(n_sample >= nlist, TS_ERR_BAD_ARG, "ivf_train: %lld sample rows < nlist %d", (long long)n_sample, nlist);
        __nv_bfloat16* buf = nullptr;
        int rc = tmp.get(&buf, (size_t)n_sample * ix->dim_pad);
        if (rc) return rc;
        rc = launch_normalize_cast(sample, TS_F32, n_sample, ix->dim, ix->dim_pad, 1, buf, TS_BF16, s);
        if (rc) return rc;
        rows = (const uint8_t*)buf;
        row_stride = (size_t)ix->dim_pad * 2;
        n = n_sample;
    } else {
        // sample = every (size / n_sample)-th stored row, read in place through a strided tensor map
        TS_REQUIRE(ix->dtype == TS_BF16, TS_ERR_UNSUPPORTED, "ivf_train: training on stored rows needs a bf16 index");
        TS_REQUIRE(ix->size >= nlist, TS_ERR_BAD_ARG, "ivf_train: %lld stored rows < nlist %d", (long long)ix->size, nlist);
        if (n_sample <= 0 || n_sample > ix->size) n_sample = ix->size;
        if (n_sample < nlist) n_sample = nlist;
        const int64_t stride = ix->size / n_sample;
        rows = (const uint8_t*)ix->data;
        row_stride = (size_t)stride * ix->row_bytes();
        n = n_sample;
    }
    int rc = ivf_alloc_centroids(ix, nlist);
    if (rc) return rc;
    init_centroids_kernel<<<nlist, 256, 0, s>>>(rows, row_stride, n, nlist, ix->dim_pad, seed, ix->centroids,
                                                ix->centroids_bf16);
    TS_LAUNCH_CHECK();
    uint32_t *assign = nullptr, *sorted_assign = nullptr, *iota = nullptr, *members = nullptr;
    int64_t* offsets = nullptr;
    if ((rc = tmp.get(&assign, (size_t)n)) || (rc = tmp.get(&sorted_assign, (size_t)n)) || (rc = tmp.get(&iota, (size_t)n)) ||
        (rc = tmp.get(&members, (size_t)n)) || (rc = tmp.get(&offsets, (size_t)nlist + 1)))
        return rc;
    for (int it = 0; it < iters; ++it) {
        rc = launch_assign(ix->device, rows, n, row_stride, ix->centroids_bf16, nlist, ix->dim_pad, assign, nullptr, s);
        if (rc) return rc;
        rc = sort_by_list(assign, n, nlist, sorted_assign, iota, members, offsets, tmp, s);
        if (rc) return rc;
        update_centroids_kernel<<<nlist, 256, 0, s>>>(rows, row_stride, n, members, offsets, ix->dim_pad,
                                                      splitmix64(seed + 1 + (uint64_t)it), ix->centroids,
                                                      ix->centroids_bf16, assign, nlist, it + 2 < iters ? 1 : 0);
        TS_LAUNCH_CHECK();
    }
    rc = ivf_finish_centroids(ix, s);
    if (rc) return rc;
    TS_CHECK_CUDA(cudaStreamSynchronize(s));   // temporaries are freed on return
    return TS_OK;
}

int ts_ivf_set_centroids(ts_index* ix, const float* centroids, int nlist, void* stream) {
    TS_REQUIRE(ix != nullptr && centroids != nullptr, TS_ERR_BAD_ARG, "ivf_set_centroids: NULL argument");
    TS_REQUIRE(nlist >= 1 && nlist <= (1 << 20), TS_ERR_BAD_ARG, "ivf_set_centroids: nlist=%d", nlist);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_set_centroids: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ivf_alloc_centroids(ix, nlist);
    if (rc) return rc;
    rc = launch_normalize_cast(centroids, TS_F32, nlist, ix->dim, ix->dim_pad, 0, ix->centroids, TS_F32, s);
    if (rc) return rc;
    rc = launch_normalize_cast(centroids, TS_F32, nlist, ix->dim, ix->dim_pad, 0, ix->centroids_bf16, TS_BF16, s);
    if (rc) return rc;
    return ivf_finish_centroids(ix, s);
}

int ts_ivf_get_centroids(const ts_index* ix, float* out, void* stream) {
    TS_REQUIRE(ix != nullptr && out != nullptr, TS_ERR_BAD_ARG, "ivf_get_centroids: NULL argument");
    TS_REQUIRE(ix->nlist > 0, TS_ERR_STATE, "ivf_get_centroids: no centroids (call ts_ivf_train first)");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_get_centroids: cannot select CUDA device %d", ix->device);
    return launch_dequant_rows(ix->centroids, TS_F32, ix->nlist, ix->dim, ix->dim_pad, out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ build
int ts_ivf_build(ts_index* ix, int list_dtype, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_build: index is NULL");
    TS_REQUIRE(ix->nlist > 0, TS_ERR_STATE, "ivf_build: no centroids (call ts_ivf_train or ts_ivf_set_centroids first)");
    TS_REQUIRE(list_dtype == TS_BF16 || list_dtype == TS_FP8_E4M3, TS_ERR_BAD_ARG,
               "ivf_build: list dtype must be TS_BF16 or TS_FP8_E4M3 (got %d)", list_dtype);
    TS_REQUIRE(ix->dtype == TS_BF16, TS_ERR_UNSUPPORTED, "ivf_build: the corpus must be stored as bf16");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_build: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    ivf_free_lists(ix);
    const int64_t n = ix->size;
    const uint32_t lrow = list_dtype == TS_BF16 ? (uint32_t)ix->row_bytes() : (uint32_t)((ix->dim + 15) / 16 * 16);
    TempBufs tmp;
    uint32_t *assign = nullptr, *sorted_assign = nullptr, *iota = nullptr;
    int rc;
    if ((rc = tmp.get(&assign, (size_t)n)) || (rc = tmp.get(&sorted_assign, (size_t)n)) || (rc = tmp.get(&iota, (size_t)n)))
        return rc;
    TS_CHECK_CUDA(cudaMalloc(&ix->list_offsets, (2 * (size_t)ix->nlist + 1) * sizeof(int64_t)));
    if ((rc = side_table_create(&ix->pos_store, &ix->pos_of_row, ix->device, ix->capacity))) return rc;
    TS_CHECK_CUDA(cudaMalloc(&ix->list_rows, std::max<size_t>((size_t)n, 1) * sizeof(uint32_t)));
    TS_CHECK_CUDA(cudaMalloc(&ix->list_data, std::max<size_t>((size_t)n, 1) * lrow));
    if (list_dtype == TS_FP8_E4M3)
        TS_CHECK_CUDA(cudaMalloc(&ix->list_scales, std::max<size_t>((size_t)n, 1) * sizeof(float)));
    ix->list_dtype = list_dtype;
    ix->list_row_bytes = lrow;
    rc = launch_assign(ix->device, ix->data, n, ix->row_bytes(), ix->centroids_bf16, ix->nlist, ix->dim_pad, assign,
                       nullptr, s);
    if (rc) return rc;
    rc = sort_by_list(assign, n, ix->nlist, sorted_assign, iota, ix->list_rows, ix->list_offsets, tmp, s);
    if (rc) return rc;
    if (n > 0) {
        const int blocks = (int)std::min<int64_t>((n + 7) / 8, 148 * 16);
        if (list_dtype == TS_BF16)
            gather_quantize_kernel<2><<<blocks, 256, 0, s>>>((const uint8_t*)ix->data, (uint32_t)ix->row_bytes(),
                                                             ix->list_rows, n, ix->dim_pad, lrow, (uint8_t*)ix->list_data,
                                                             nullptr);
        else
            gather_quantize_kernel<1><<<blocks, 256, 0, s>>>((const uint8_t*)ix->data, (uint32_t)ix->row_bytes(),
                                                             ix->list_rows, n, ix->dim_pad, lrow, (uint8_t*)ix->list_data,
                                                             ix->list_scales);
        TS_LAUNCH_CHECK();
        invert_rows_kernel<<<1024, 256, 0, s>>>(ix->list_rows, n, 0, ix->pos_of_row);
        TS_LAUNCH_CHECK();
    }
    TS_CHECK_CUDA(cudaStreamSynchronize(s));
    ix->list_cap = std::max<int64_t>(n, 1);
    ix->built_n = n;
    ix->ovf_n = 0;
    ix->ivf_dead = 0;
    ix->ivf_built = true;
    return TS_OK;
}

int ts_ivf_nlist(const ts_index* ix) { return ix ? ix->nlist : -1; }

int ts_ivf_repack(ts_index* ix, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_repack: index is NULL");
    TS_REQUIRE(ix->ivf_built, TS_ERR_STATE, "ivf_repack: lists are not built (call ts_ivf_build)");
    if (ix->ovf_n == 0 && ix->ivf_dead == 0) return TS_OK;
    return ts_ivf_build(ix, ix->list_dtype, stream);
}

int ts_ivf_pending(const ts_index* ix, int64_t* overflow_rows, int64_t* dead_positions) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_pending: index is NULL");
    if (overflow_rows) *overflow_rows = ix->ivf_built ? ix->ovf_n : 0;
    if (dead_positions) *dead_positions = ix->ivf_built ? ix->ivf_dead : 0;
    return TS_OK;
}

int ts_debug_ivf_timeline(uint64_t* out_host, int n_ctas) {
    TS_REQUIRE(out_host != nullptr && n_ctas >= 1 && n_ctas <= 148, TS_ERR_BAD_ARG, "debug_ivf_timeline: bad argument");
    TS_CHECK_CUDA(cudaDeviceSynchronize());
    TS_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_ivf_timeline, (size_t)n_ctas * 12 * sizeof(uint64_t)));
    return TS_OK;
}

int ts_ivf_list_sizes(const ts_index* ix, int64_t* out, void* stream) {
    TS_REQUIRE(ix != nullptr && out != nullptr, TS_ERR_BAD_ARG, "ivf_list_sizes: NULL argument");
    TS_REQUIRE(ix->ivf_built, TS_ERR_STATE, "ivf_list_sizes: lists are not built (call ts_ivf_build)");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_list_sizes: cannot select CUDA device %d", ix->device);
    list_sizes_kernel<<<(ix->nlist + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ix->list_offsets, ix->nlist, out);
    TS_LAUNCH_CHECK();
    if (ix->ovf_n > 0) {   // + the overflow lists (tombstoned positions still count until the re-pack)
        add_list_sizes_kernel<<<(ix->nlist + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ix->list_offsets + ix->nlist,
                                                                                         ix->nlist, out);
        TS_LAUNCH_CHECK();
    }
    return TS_OK;
}

int ts_ivf_get_lists(const ts_index* ix, int64_t* offsets_out, int64_t* rows_out, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_get_lists: index is NULL");
    TS_REQUIRE(ix->ivf_built, TS_ERR_STATE, "ivf_get_lists: lists are not built (call ts_ivf_build)");
    TS_REQUIRE(ix->ovf_n == 0 && ix->ivf_dead == 0 && ix->ivf_moved == 0, TS_ERR_STATE,
               "ivf_get_lists: rows were added, replaced or deleted since the build (%lld in overflow lists, %lld "
               "tombstones); call ts_ivf_repack first", (long long)ix->ovf_n, (long long)ix->ivf_dead);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_get_lists: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (offsets_out)
        TS_CHECK_CUDA(cudaMemcpyAsync(offsets_out, ix->list_offsets, ((size_t)ix->nlist + 1) * sizeof(int64_t),
                                      cudaMemcpyDeviceToDevice, s));
    if (rows_out && ix->size > 0) {
        widen_u32_kernel<<<1024, 256, 0, s>>>(ix->list_rows, ix->size, rows_out);
        TS_LAUNCH_CHECK();
    }
    return TS_OK;
}

int ts_ivf_get_list_data(const ts_index* ix, int64_t first, int64_t n, float* out, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_get_list_data: index is NULL");
    TS_REQUIRE(ix->ivf_built, TS_ERR_STATE, "ivf_get_list_data: lists are not built (call ts_ivf_build)");
    TS_REQUIRE(first >= 0 && n >= 0 && first + n <= ix->built_n + ix->ovf_n, TS_ERR_BAD_ARG,
               "ivf_get_list_data: [%lld, %lld) outside [0, %lld)", (long long)first, (long long)(first + n),
               (long long)(ix->built_n + ix->ovf_n));
    if (n == 0) return TS_OK;
    TS_REQUIRE(out != nullptr, TS_ERR_BAD_ARG, "ivf_get_list_data: out is NULL");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_get_list_data: cannot select CUDA device %d", ix->device);
    const int64_t total = n * (int64_t)ix->dim;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    dequant_list_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)ix->list_data, ix->list_row_bytes,
                                                                  ix->list_dtype, ix->list_scales, first, n, ix->dim, out);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

// ------------------------------------------------------------------------------------ search
}  // extern "C"

namespace ts {

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

static ts_index centroid_view(const ts_index* ix) {
    ts_index v;
    v.device = ix->device;
    v.dim = ix->dim;
    v.dim_pad = ix->dim_pad;
    v.dtype = TS_BF16;
    v.capacity = ix->nlist;
    v.size = ix->nlist;
    v.data = ix->centroids_bf16;
    v.max_norm2 = ix->centroid_max_norm2;
    return v;
}

static int ivf_parts(const ts_index* ix, int nq) {
    const int sms = sm_count(ix->device);
    if (tunables().ivf_parts > 0) return std::min(sms, tunables().ivf_parts);
    return std::max(1, std::min(sms, (2 * sms + nq - 1) / std::max(nq, 1)));
}

// K4d hands whole lists to CTAs: it needs many lists (one list = one CTA per query group, so an index with a
// handful of huge lists — e.g. the one-list fp8 shadow — would run on a handful of SMs) and a batch.
static bool ivf_use_grouped(const ts_index* ix, int nq, int kc, int nprobe) {
    const int m = tunables().ivf_group_min_nq;
    if (m <= 0 || nq < m || !ivf_grouped_supported(ix, kc)) return false;
    // overflow lists are ordinary (virtual) lists for the list-major scan; tombstones are skipped by the tcgen05 back
    // end's epilogue only — the older back ends fall back to the per-query scan until the re-pack
    if (ix->ivf_dead > 0 && tunables().ivf_group_mma < 3) return false;
    const int64_t pairs = (int64_t)nq * std::min(nprobe, ix->nlist);
    const int min_lists = tunables().ivf_group_min_lists > 0 ? tunables().ivf_group_min_lists : 4 * sm_count(ix->device);
    return ix->nlist >= min_lists && pairs >= min_lists;
}

struct IvfWs {
    void* grouped;
    float* q32;
    uint64_t* probes;
    uint64_t* probes2;
    void* coarse;
    size_t coarse_bytes;
    uint64_t* part_keys;
    uint32_t* tickets;
    uint64_t* cand;
    size_t bytes;
};
static IvfWs carve_ivf(const ts_index* ix, int nq, int kc, int nprobe, void* base) {
    IvfWs w;
    const ts_index view = centroid_view(ix);
    size_t off = 0;
    auto take = [&](size_t b) {
        char* p = (char*)base + off;
        off += al256(b);
        return (void*)p;
    };
    w.q32 = (float*)take((size_t)nq * ix->dim_pad * 4);
    w.probes = (uint64_t*)take((size_t)nq * nprobe * 8);
    w.probes2 = (uint64_t*)take((size_t)nq * nprobe * 2 * 8);   // (main list, overflow list) per probe, when rows were added
    w.coarse_bytes = ts_workspace_bytes(&view, nq, nprobe);
    w.coarse = take(w.coarse_bytes);
    w.part_keys = (uint64_t*)take((size_t)nq * ivf_parts(ix, nq) * kc * 8);
    w.tickets = (uint32_t*)take((size_t)nq * 4);
    w.cand = (uint64_t*)take((size_t)nq * kc * 8);
    w.grouped = ivf_use_grouped(ix, nq, kc, nprobe) ? take(ivf_grouped_workspace_bytes(ix, nq, nprobe)) : nullptr;
    w.bytes = off;
    return w;
}

static int ivf_search_impl(ts_index* ix, const void* queries, int q_dtype, int nq, int k, int nprobe, int rescore_k,
                           int normalize, const uint32_t* allow_mask, uint64_t* out_keys, float* out_scores, int64_t* out_ids, void* workspace,
                           size_t workspace_bytes, cudaStream_t s) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ivf_search: index is NULL");
    TS_REQUIRE(ix->ivf_built, TS_ERR_STATE, "ivf_search: lists are not built (call ts_ivf_train + ts_ivf_build)");
    TS_REQUIRE(nq >= 0, TS_ERR_BAD_ARG, "ivf_search: nq=%d", nq);
    TS_REQUIRE(k >= 1 && k <= TS_IVF_MAX_CANDIDATES, TS_ERR_BAD_ARG, "ivf_search: k=%d out of range [1, %d]", k,
               TS_IVF_MAX_CANDIDATES);
    TS_REQUIRE(nprobe >= 1, TS_ERR_BAD_ARG, "ivf_search: nprobe=%d", nprobe);
    TS_REQUIRE(rescore_k <= TS_IVF_MAX_CANDIDATES, TS_ERR_BAD_ARG, "ivf_search: rescore_k=%d exceeds %d", rescore_k,
               TS_IVF_MAX_CANDIDATES);
    TS_REQUIRE(q_dtype == TS_F32 || q_dtype == TS_BF16 || q_dtype == TS_F16, TS_ERR_BAD_ARG, "ivf_search: query dtype %d",
               q_dtype);
    if (nq == 0) return TS_OK;
    TS_REQUIRE(queries != nullptr && workspace != nullptr, TS_ERR_BAD_ARG, "ivf_search: NULL buffer");
    nprobe = std::min(std::min(nprobe, ix->nlist), TS_MAX_K);
    const int kc = std::max(k, rescore_k);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_search: cannot select CUDA device %d", ix->device);
    IvfWs w = carve_ivf(ix, nq, kc, nprobe, workspace);
    TS_REQUIRE(workspace_bytes >= w.bytes, TS_ERR_CAPACITY, "ivf_search: workspace %zu < %zu bytes", workspace_bytes,
               w.bytes);
    // 1. coarse: exact top-nprobe over the centroid table (K2 for a few queries, K3 for a batch)
    // (the coarse scan normalises the raw queries itself and also writes the prepared fp32 copy the list scan and
    // the re-score read — no separate preparation launch; a batch (K3) prepares them with K1's kernel)
    ts_index view = centroid_view(ix);
    // both ticket arrays are zeroed up front so that the coarse scan and the list scan are adjacent launches
    TS_CHECK_CUDA(cudaMemsetAsync(w.tickets, 0, (size_t)nq * sizeof(uint32_t), s));
    int rc = search_impl(&view, queries, q_dtype, nq, nprobe, normalize, nullptr, w.probes, nullptr, nullptr, w.coarse,
                         w.coarse_bytes, s, nullptr, nullptr, nullptr, 0, w.q32);
    if (rc) return rc;
    // 2. scan the probed lists, keep kc candidates per query
    ListScanParams p;
    p.list_data = (const uint8_t*)ix->list_data;
    p.row_bytes = ix->list_row_bytes;
    p.dim_pad = ix->dim_pad;
    p.scales = ix->list_scales;
    p.list_offsets = ix->list_offsets;
    p.probes = w.probes;
    p.nprobe = nprobe;
    if (ix->ovf_n > 0) {   // rows added since the build: every probed list brings its overflow list along
        const int64_t np = (int64_t)nq * nprobe;
        expand_probes_kernel<<<(int)std::min<int64_t>((np + 255) / 256, 1024), 256, 0, s>>>(w.probes, np, ix->nlist,
                                                                                         w.probes2);
        TS_LAUNCH_CHECK();
        p.probes = w.probes2;
        p.nprobe = 2 * nprobe;
    }
    p.queries = w.q32;
    p.mask = allow_mask;
    p.list_rows = ix->list_rows;
    p.has_dead = ix->ivf_dead > 0 ? 1 : 0;
    p.k = kc;
    p.part_keys = w.part_keys;
    p.tickets = w.tickets;
    p.out_keys = w.cand;
    p.run_flag = nullptr;
    p.stages = 0;
    p.k_out = 0;
    p.corpus = (const uint8_t*)ix->data;
    p.corpus_row_bytes = (uint32_t)ix->row_bytes();
    p.id_map = ix->has_ids ? ix->ids : nullptr;
    p.fin_keys = out_keys;
    p.fin_scores = out_scores;
    p.fin_ids = out_ids;
    // small batches (the latency path): the list scan's last CTA re-scores its candidates itself — one launch fewer
    const bool fuse_rescore = w.grouped == nullptr && nq < 64 && tunables().ivf_fuse_rescore != 0;
    if (fuse_rescore) p.k_out = k;
    if (w.grouped != nullptr) {
        // large batch: list-major scan (each probed list read once per 4 queries); K4b below runs only if the
        // score buffer turned out too small for this batch (decided on the device)
        rc = launch_ivf_grouped(ix, p.probes, w.q32, nq, nprobe, ix->ovf_n > 0 ? 2 : 1, kc, allow_mask, w.grouped, w.cand,
                                &p.run_flag, s);
        if (rc) return rc;
    }
    rc = launch_list_scan(ix, p, nq, ivf_parts(ix, nq), s);
    if (rc) return rc;
    if (fuse_rescore) return TS_OK;
    // 3. exact re-score of the survivors against the stored corpus rows
    int P = 64;
    while (P < kc) P <<= 1;
    // a lone query spreads its candidates over 32 warps (latency), a batch keeps CTAs small (throughput)
    // (P <= 1024 candidates, <= 32 per warp: at least P/32 warps)
    const int rs_threads = nq < 64 ? 1024 : std::max(256, P);
    ivf_rescore_kernel<2><<<nq, rs_threads, (size_t)P * 8, s>>>(w.cand, kc, k, ix->list_rows, w.q32, (const uint8_t*)ix->data,
                                                         (uint32_t)ix->row_bytes(), ix->dim_pad,
                                                         ix->has_ids ? ix->ids : nullptr, out_keys, out_scores, out_ids);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

}  // namespace ts

extern "C" {

size_t ts_ivf_workspace_bytes(const ts_index* ix, int nq, int k, int nprobe, int rescore_k) {
    if (!ix || ix->nlist <= 0 || nq < 0 || k < 1 || nprobe < 1) return 0;
    nprobe = std::min(std::min(nprobe, ix->nlist), TS_MAX_K);
    return carve_ivf(ix, std::max(nq, 1), std::max(k, rescore_k), nprobe, nullptr).bytes;
}

int ts_ivf_search(ts_index* ix, const void* queries, int q_dtype, int nq, int k, int nprobe, int rescore_k,
                  int normalize_queries, const uint32_t* allow_mask, float* out_scores, int64_t* out_ids,
                  void* workspace, size_t workspace_bytes, void* stream) {
    TS_REQUIRE(nq == 0 || (out_scores != nullptr && out_ids != nullptr), TS_ERR_BAD_ARG,
               "ivf_search: output pointers are NULL");
    return ivf_search_impl(ix, queries, q_dtype, nq, k, nprobe, rescore_k, normalize_queries, allow_mask, nullptr, out_scores,
                           out_ids, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ts_ivf_search_keys(ts_index* ix, const void* queries, int q_dtype, int nq, int k, int nprobe, int rescore_k,
                       int normalize_queries, const uint32_t* allow_mask, uint64_t* out_keys, void* workspace,
                       size_t workspace_bytes, void* stream) {
    TS_REQUIRE(nq == 0 || out_keys != nullptr, TS_ERR_BAD_ARG, "ivf_search_keys: out_keys is NULL");
    return ivf_search_impl(ix, queries, q_dtype, nq, k, nprobe, rescore_k, normalize_queries, allow_mask, out_keys, nullptr, nullptr,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
