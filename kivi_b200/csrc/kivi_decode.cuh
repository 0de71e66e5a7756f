// kivi_decode.cuh -- KIVI cache layout in HBM + mbarrier / bulk-copy (TMA) PTX helpers (sm_100a).
//
// Cache layout (one layer; U = B * Hkv units, D = 128 channels, cell = 32 consecutive elements):
//
//   K store  [U][k_cap_blocks][4 quarters][QB bytes]    block = 128 tokens, per-channel quantised
//            quarter qt holds channels d in [32*qt, 32*qt+32):
//              codes [32 rows][4 cells][cbk bytes]        cbk = 4*k_bits (cell of 32 TOKENS of channel d)
//              meta  [32 rows][128/g][half2(scale, zero)]
//            -> a (block, quarter) is ONE contiguous, 16-B aligned run of QB bytes = one bulk copy.
//   V store  codes [U][v_cap][4 cells][cbv bytes]         cbv = 4*v_bits (cell of 32 CHANNELS of token t)
//            meta  [U][v_cap][128/g][half2(scale, zero)]  -> a run of tokens = two bulk copies.
//   K residual [U][R][128] fp16           (tokens tk .. tk+r-1, newest last)
//   V residual [U][v_res_cap][128] fp16   ring buffer, head = state.vhead, L valid tokens
//   state    int32[8] on the device, shared by all layers of a model (every layer sees the same
//            lengths): {tk, r, tv, L, vhead, kv_len, 0, 0}
//
// Policy restated from models/llama_kivi.py:343-356 (K: the fp16 window is quantised per channel in
// groups of g tokens as soon as it holds R tokens) and :386-399 (V: the window holds the newest R
// tokens; each step the oldest one is quantised per token in groups of g channels).
#pragma once
#include "kivi_common.cuh"

namespace kivi {

constexpr int kD = 128;            // head_dim of every model the reference ships (Llama / Mistral)
constexpr int kBlockTokens = 128;  // K store block
constexpr int kCell = 32;

struct CacheDesc {
    int B, Hkv, H, k_bits, v_bits, g, R;
    int k_cap_blocks, v_cap, v_res_cap;
    uint8_t* k_store;
    uint8_t* v_codes;
    uint8_t* v_meta;
    __half* k_res;
    __half* v_res;
    int* state;
};

enum { ST_TK = 0, ST_R = 1, ST_TV = 2, ST_L = 3, ST_VHEAD = 4, ST_KVLEN = 5 };

__host__ __device__ inline int k_cell_bytes(int bits) { return 4 * bits; }
constexpr int kQRows = 32;         // channels per K quarter
__host__ __device__ inline int k_q_code_bytes(int bits) { return kQRows * 4 * k_cell_bytes(bits); }
__host__ __device__ inline int k_q_meta_bytes(int g) { return kQRows * (kBlockTokens / g) * 4; }
__host__ __device__ inline int k_q_bytes(int bits, int g) { return k_q_code_bytes(bits) + k_q_meta_bytes(g); }
// byte offset of the (block, channel d) row inside a unit's K store, and of its meta row
__host__ __device__ inline int64_t k_row_off(int blk, int d, int bits, int g) {
    return ((int64_t)blk * 4 + d / kQRows) * k_q_bytes(bits, g) + (d % kQRows) * (4 * k_cell_bytes(bits));
}
__host__ __device__ inline int64_t k_meta_off(int blk, int d, int bits, int g) {
    return ((int64_t)blk * 4 + d / kQRows) * k_q_bytes(bits, g) + k_q_code_bytes(bits) + (d % kQRows) * ((kBlockTokens / g) * 4);
}
__host__ __device__ inline int64_t k_unit_bytes(int cap_blocks, int bits, int g) { return (int64_t)cap_blocks * 4 * k_q_bytes(bits, g); }
__host__ __device__ inline int v_tok_code_bytes(int bits) { return 4 * 4 * bits; }           // 128 channels
__host__ __device__ inline int v_tok_meta_bytes(int g) { return (kD / g) * 4; }

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP), completion on an mbarrier.
// dst/src 16-B aligned, bytes % 16 == 0.  The packed cache is read exactly once per step: evict-first.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace kivi
