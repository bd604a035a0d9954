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


// Same 16 e4m3 values against 16 query values held as 8 half2: packed HFMA2 into two fp16 accumulators
// (8 products each, |e4m3| <= 448 and |q| <= 1 so no overflow), widened to fp32 once per chunk.
// 1.8 instructions per element instead of 3; the fp16 rounding it adds (~1e-5 on a unit-vector score)
// is two orders below the e4m3 quantisation noise of the list rows, and candidates are re-scored exactly.
__device__ __forceinline__ float dot16_e4m3_h2(const uint4& v, const __half2* q, float acc) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    __half2 s0 = __float2half2_rn(0.f), s1 = __float2half2_rn(0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] & 0xFFFFu), __NV_E4M3);
        const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] >> 16), __NV_E4M3);
        s0 = __hfma2(*reinterpret_cast<const __half2*>(&lo), q[2 * i], s0);
        s1 = __hfma2(*reinterpret_cast<const __half2*>(&hi), q[2 * i + 1], s1);
    }
    const float2 f = __half22float2(__hadd2(s0, s1));
    return acc + (f.x + f.y);
}

// ---------------------------------------------------------------------------------- buffered warp select
// WarpTopK plus a register buffer of pending candidates (FAISS-WarpSelect style): a key that beats the
// current threshold is APPENDED (a handful of instructions); when 32*KPL keys are pending they are sorted
// with a register bitonic network and folded into the list with one bitonic merge. Element-wise insertion
// costs ~70 dependent instructions per key at KPL = 4, which dominates short scans where most rows still
// enter the list (IVF list scans: tens of rows per warp, k' = 100).
template <int KPL>
struct WarpSelect {
    WarpTopK<KPL> best;
    uint64_t pend[KPL];
    int count;
    uint64_t thr;   // k-th key of `best` as of the last flush

    __device__ __forceinline__ void init() {
        best.clear();
#pragma unroll
        for (int j = 0; j < KPL; ++j) pend[j] = 0ull;
        count = 0;
        thr = 0ull;
    }

    // descending bitonic sort of the 32*KPL pending keys (position p = j*32 + lane)
    __device__ __forceinline__ void sort_pending(int lane) {
        constexpr int LOGK = 5 + ilog2_c(KPL);
        // unit-step loops over log2(size) / log2(stride): fully unrolled, all register indices static
#pragma unroll
        for (int lsize = 1; lsize <= LOGK; ++lsize) {
            const int size = 1 << lsize;
#pragma unroll
            for (int lstride = lsize - 1; lstride >= 0; --lstride) {
                const int stride = 1 << lstride;
                if (lstride >= 5) {
                    const int sj = stride >> 5;
#pragma unroll
                    for (int j = 0; j < KPL; ++j) {
                        if ((j & sj) == 0) {
                            const bool desc = ((j * 32) & size) == 0;
                            const uint64_t x = pend[j], y = pend[j | sj];
                            const bool sw = desc ? (x < y) : (x > y);
                            pend[j] = sw ? y : x;
                            pend[j | sj] = sw ? x : y;
                        }
                    }
                } else {
                    const bool lower = (lane & stride) != 0;
#pragma unroll
                    for (int j = 0; j < KPL; ++j) {
                        const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, pend[j], stride);
                        const bool desc = (((j * 32) | lane) & size) == 0;
                        const bool keep_max = desc != lower;
                        const bool take = keep_max ? (o > pend[j]) : (o < pend[j]);
                        pend[j] = take ? o : pend[j];
                    }
                }
            }
        }
    }

    __device__ __forceinline__ void flush(int k, int lane) {
        if (count == 0) return;   // warp-uniform
        sort_pending(lane);
        best.merge_desc(pend, lane);
#pragma unroll
        for (int j = 0; j < KPL; ++j) pend[j] = 0ull;
        count = 0;
        thr = best.kth(k);
    }

    // every lane passes its candidate key (0 = none); warp-uniform control flow. New keys always land in
    // register 0 (lane = count mod 32); a full register row is rotated up — static indices only.
    __device__ __forceinline__ void offer(uint64_t key, int k, int lane) {
        unsigned m = __ballot_sync(0xFFFFFFFFu, key > thr);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t x = __shfl_sync(0xFFFFFFFFu, key, src);
            if (x <= thr) continue;   // the threshold may have risen in a flush since the ballot
            if (lane == (count & 31)) pend[0] = x;
            ++count;
            if ((count & 31) == 0) {
                if (count == 32 * KPL) {
                    flush(k, lane);
                } else {
#pragma unroll
                    for (int j = KPL - 1; j >= 1; --j) pend[j] = pend[j - 1];
                    pend[0] = 0ull;
                }
            }
        }
    }
};

// One warp selects the best k keys of a run of `len` fp32 scores (positions pos0 .. pos0 + len - 1; -inf = masked).
// Fast path: a lane loads four scores at once and compares them with the running threshold as FLOATS; only when
// some lane of the warp has a candidate are keys packed and offered to the WarpSelect — after the first few hundred
// rows that happens for ~k ln(n) / n of the steps. `run` must be 16-byte aligned; len is padded to 4 in memory.
__device__ __forceinline__ float pick4(const float4& v, int i) {   // register selects, never local memory
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}
template <int KPL>
__device__ __forceinline__ void warp_select_run(WarpSelect<KPL>& sel, const float* __restrict__ run, int len, uint32_t pos0,
                                                int k, int lane) {
    float thr_f = sel.thr ? key_score(sel.thr) : -INFINITY;
    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    // software-pipelined: the next 512 scores are requested before the current ones are examined (a warp walks its run
    // alone, so without this every step waits out a full DRAM round trip: 4096 queries x 78k scores took 0.45 ms)
    float4 nxt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int r = u * 128 + 4 * lane;
        nxt[u] = (r < len) ? __ldg(reinterpret_cast<const float4*>(run + r)) : ninf4;
    }
    for (int r0 = 0; r0 < len; r0 += 512) {            // 4 x (32 lanes x float4) per iteration, loads issued together
        float4 v[4];
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = nxt[u];
            const int rn = r0 + 512 + u * 128 + 4 * lane;
            nxt[u] = (rn < len) ? __ldg(reinterpret_cast<const float4*>(run + rn)) : ninf4;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * 128 + 4 * lane;
            const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) any |= (r + i < len) && (f[i] >= thr_f) && (f[i] > -INFINITY);
        }
        if (!__any_sync(0xFFFFFFFFu, any)) continue;
        // slow path, ONE call site of offer(): its flush (a 128..256-key register sort + merge, ~2000 instructions)
        // is inlined wherever offer() is called, and sixteen copies of it thrashed the instruction cache
        // (dense select 1.55 ms instead of 0.1 ms)
#pragma unroll 1
        for (int t = 0; t < 16; ++t) {
            const int u = t >> 2, i = t & 3;
            const float4 vu = u == 0 ? v[0] : (u == 1 ? v[1] : (u == 2 ? v[2] : v[3]));
            const float f = pick4(vu, i);
            const int r = r0 + u * 128 + 4 * lane + i;
            const bool live = (r < len) && (f >= thr_f) && (f > -INFINITY);
            if (!__any_sync(0xFFFFFFFFu, live)) continue;
            sel.offer(live ? pack_key(f, pos0 + (uint32_t)r) : 0ull, k, lane);
            thr_f = sel.thr ? key_score(sel.thr) : -INFINITY;
        }
    }
}

}  // namespace ts
