// tcgen05 / TMEM / 2-D TMA device wrappers and the host-side tensor-map helper, shared by the
// GEMM-shaped kernels (K3 batched search, K4a IVF assignment).
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>

#include "ts_common.cuh"

namespace ts {

namespace umma {
constexpr int BM = 128;          // rows of the streamed operand per tile (UMMA M)
constexpr int BN = 256;          // rows of the resident operand per tile (UMMA N), upper bound
constexpr int BK = 64;           // bf16 per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
}  // namespace umma

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    // watchdog: a pipeline bug must trap, not hang the GPU box
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// fp32 operands, read by the tensor core as TF32 (the low 13 mantissa bits are ignored): K = 8 per instruction
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile whose rows are 128 bytes apart: 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_smem_desc(const void* smem) {
    uint64_t d = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu);
    d |= (uint64_t)(1024u >> 4) << 32;  // stride byte offset
    d |= 1ull << 46;                    // descriptor version (sm_100)
    d |= 2ull << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------- host: tensor maps
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    return fn;
}

// rows of `cols` elements (bf16, or fp32 when f32 = true), `row_bytes` apart; one box = box_rows rows x one
// 128-byte swizzle span (64 bf16 / 32 fp32); columns past `cols` read as zero
inline int make_tmap_rows(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_bytes,
                          uint32_t box_rows, bool f32) {
    auto enc = get_encode_fn();
    TS_REQUIRE(enc != nullptr, TS_ERR_CUDA, "batched: cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? umma::BK / 2 : umma::BK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TS_REQUIRE(r == CUDA_SUCCESS, TS_ERR_CUDA, "batched: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return TS_OK;
}
inline int make_tmap_bf16_rows(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_bytes,
                               uint32_t box_rows) {
    return make_tmap_rows(map, base, rows, cols, row_bytes, box_rows, false);
}


}  // namespace ts
