// kivi_decode.cuh -- KIVI cache layout in HBM (tensor-core friendly blocks) + mbarrier / bulk-copy (TMA)
// PTX helpers (sm_100a).
//
// Every packed store is a sequence of 128 x 128 BLOCKS "inner x outer":
//     K store: inner = channel d (the reduction index of q.K^T), outer = token   -> one block = 128 tokens
//     V store: inner = token t  (the reduction index of p.V),   outer = channel  -> one block = 128 tokens
// Quantisation groups run along the OUTER dim (g tokens per channel for K, g channels per token for V --
// exactly the reference's per-channel K / per-token V scheme), so a block is the same object for both.
//
//   block = [ codes: 8 chunks x kChunkBytes ][ meta: 8 chunks x (128/g) groups x 4 x 16 bytes ]
//   chunk c = inner indices 16c .. 16c+15.  Its codes are stored as the A-operand fragments of
//   mma.sync.m16n8k16 (row = outer, col = inner), so that a lane's 128-bit shared-memory load yields its four
//   A registers for F = 16/bits consecutive MMAs at once:
//       word(lane = 4*g8 + t, r), r = 0..3:   inner pair 16c + 2t + 8*(r >> 1) + {0, 1},  row = g8 + 8*(r & 1)
//       low  16 bits: field j (bits [bits*j, bits*j + bits)) = code[inner even][outer 16*j + row]
//       high 16 bits: field j                                = code[inner odd ][outer 16*j + row]
//   (a "slab" = the 16*F outer rows covered by one word set: 128 rows at 2 bits, 64 rows at 4 bits.)
//   A code becomes an fp16 MMA operand with ONE LOP3 per PAIR of codes (see Lay<>::shr below).
//   meta entry (c, G, t) = 8 halfs { z[i0], z[i0+1], s[i0], s[i0+1], z[i0+8], z[i0+9], s[i0+8], s[i0+9] },
//   i0 = 16c + 2t, of outer group G: ONE 128-bit load gives a lane the two half2 of scales it multiplies with
//   its half2 of x (q or p) to build its B fragment, and its four registers ARE the A operand of the zero-term
//   MMA (rows 0..7 = zeros of group G; rows 8..15 = the scales, whose products are ignored).
//
//   K window [U][R][128] fp16, V window [U][R+1][128] fp16 ring (head = state.vhead).  Inside a 256-byte row the eight-
//   channel (16-byte) units are XOR-swizzled with the row's slot index: unit' = unit ^ (slot & 7)  (win_off below), so that
//   the MMA fragment loads of 8 consecutive rows at one channel offset hit 8 different bank groups of shared memory.
//   state    int32[8] on the device, shared by the layers of a model: {tk, r, tv, L, vhead, kv_len}
//
// Policy restated from models/llama_kivi.py:343-356 (K: the fp16 window is quantised per channel in
// groups of g tokens as soon as it holds R tokens) and :386-399 (V: the window holds the newest R
// tokens; each step the oldest one is quantised per token in groups of g channels).
#pragma once
#include "kivi_common.cuh"

namespace kivi {

constexpr int kD = 128;            // head_dim of every model the reference ships (Llama / Mistral)
constexpr int kBlockTokens = 128;  // tokens per block (K: outer rows, V: inner rows)

struct CacheDesc {
    int B, Hkv, H, k_bits, v_bits, g, R;
    int k_cap_blocks, v_cap_blocks, v_res_cap;
    uint8_t* k_store;
    uint8_t* v_store;
    __half* k_res;
    __half* v_res;
    int* state;
};

enum { ST_TK = 0, ST_R = 1, ST_TV = 2, ST_L = 3, ST_VHEAD = 4, ST_KVLEN = 5 };

template <int BITS>
struct Lay {
    static constexpr int F = 16 / BITS;                  // fields per 16-bit half = MMAs per slab
    static constexpr int kSlabRows = 16 * F;             // outer rows per slab (128 / 64)
    static constexpr int kSlabs = 128 / kSlabRows;       // slabs per block (1 / 2)
    static constexpr int kChunkBytes = 512 * kSlabs;     // 128 words per slab
    static constexpr int kCodeBytes = 8 * kChunkBytes;   // 4096 / 8192
    // Unpack: field j of a 16-bit half is brought to bit offset P(j) in [4, 10) by an optional shift of the whole
    // word, isolated with one AND, and consumed AS IS: the fp16 denormal  code * 2^(P - 24).  mma.sync handles
    // denormal inputs exactly when their set bits sit at offset >= 4 (measured: tools/probes/mma_unpack_variants.cu,
    // error identical to normal inputs; fields at offsets 0..3 lose up to 4 bits), so no magic-number subtraction
    // is needed: ONE LOP3 per pair of codes.
    //   2-bit: fields 0,1 <- (w << 4);  fields 2,3,4 in place;  fields 5,6,7 <- (w >> 6)
    //   4-bit: field 0 <- (w << 4);  field 1 in place;  field 2 <- (w >> 4);  field 3 <- (w >> 8)
    __host__ __device__ static constexpr int shr(int j) {            // > 0: right shift, < 0: left shift
        return BITS == 2 ? (j < 2 ? -4 : (j < 5 ? 0 : 6)) : (j == 0 ? -4 : (j == 1 ? 0 : (j == 2 ? 4 : 8)));
    }
    __host__ __device__ static constexpr int bitpos(int j) { return BITS * j - shr(j); }
};

__host__ __device__ inline int lay_code_bytes(int bits) { return bits == 2 ? 4096 : 8192; }
__host__ __device__ inline int lay_meta_bytes(int g) { return 8 * (128 / g) * 4 * 16; }
__host__ __device__ inline int lay_block_bytes(int bits, int g) { return lay_code_bytes(bits) + lay_meta_bytes(g); }

// byte offset (inside a block) of the word holding element (inner i, outer o), and its bit position
__host__ __device__ inline int lay_word_off(int bits, int i, int o) {
    const int F = 16 / bits, slab_rows = 16 * F, slabs = 128 / slab_rows;
    const int c = i >> 4, ii = i & 15;
    const int p = ((ii & 7) >> 1) + 4 * (ii >> 3);
    const int sl = o / slab_rows, row = o % 16;
    const int lane = (row & 7) * 4 + (p & 3), r = (p >> 2) * 2 + (row >> 3);
    return ((c * slabs + sl) * 128 + lane * 4 + r) * 4;
}
__host__ __device__ inline int lay_bit_pos(int bits, int i, int o) {
    const int F = 16 / bits, slab_rows = 16 * F;
    return 16 * (i & 1) + bits * ((o % slab_rows) >> 4);
}
// element offset (halfs, inside a unit's window) of (slot, channel): 16-byte units swizzled with the slot index
__host__ __device__ inline int win_off(int slot, int ch) { return slot * kD + ((((ch >> 3) ^ (slot & 7)) << 3) | (ch & 7)); }
// the same for a 16-byte unit index (0..15) of the row
__host__ __device__ inline int win_unit(int slot, int unit) { return slot * 16 + (unit ^ (slot & 7)); }

// byte offsets (inside a block) of the fp16 zero / scale of (inner i, outer group G)
__host__ __device__ inline int lay_zero_off(int bits, int g, int i, int G) {
    const int c = i >> 4, ii = i & 15;
    const int t = (ii & 7) >> 1;
    return lay_code_bytes(bits) + (((c * (128 / g)) + G) * 4 + t) * 16 + (ii >> 3) * 8 + (ii & 1) * 2;
}
__host__ __device__ inline int lay_scale_off(int bits, int g, int i, int G) { return lay_zero_off(bits, g, i, G) + 4; }

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
// L2 prefetch of a contiguous piece of global memory (a hint: no completion, no destination); src 16-B aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// the same copy without a cache hint (data other warps read too)
__device__ __forceinline__ void bulk_g2s_plain(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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

// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col)      (SASS: HMMA.16816.F32)
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// D = A * B (no accumulator input: the zero registers are free, and no instruction is spent clearing D beforehand)
__device__ __forceinline__ void mma_16816_init(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// four 8x8 b16 matrices from shared memory, each delivered TRANSPOSED: lane (g8, t) receives, of matrix i, the elements
// (memory row 2t, column g8) and (memory row 2t+1, column g8); lanes 8i .. 8i+7 supply the row addresses of matrix i
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(row_ptr)));
}

}  // namespace kivi
