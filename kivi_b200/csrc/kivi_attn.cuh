// kivi_attn.cuh -- fused KIVI decode attention over the blocked cache (sm_100a), one launch per layer.
//
// Replaces the ~30 launches of the reference's decode branch (models/llama_kivi.py:314-399): q.Kq^T with
// in-register dequantisation, the fp16 K window, scale, mask, fp32 softmax, p.Vq, the fp16 V window, the
// fp16 add, and the per-unit cache data movement (window append, K flush, V token pack).  Rounding points
// of the reference are reproduced (fp16 logits -> fp16 scale -> fp32 softmax -> fp16 probs -> fp16 partial
// outputs -> fp16 add).
//
// Execution model
//   * persistent grid, CTA c handles units c, c+grid, ...; unit = (b, kv-head, chunk of G query heads):
//     packed bytes are read once per KV head for all G heads (GQA);
//   * 8 warps per CTA, each with S PRIVATE shared-memory stages.  A warp streams its own work items
//     HBM -> shared memory with 1-D bulk copies (cp.async.bulk = the TMA engine, SASS UBLKCP, L2
//     evict-first) completing on the stage's mbarrier; right after consuming a stage its elected lane issues
//     the copy of the item S positions ahead (fence.proxy.async orders its reads before the async write).
//     The issue cursor runs ahead across the K -> softmax -> V phases and across units, so HBM never idles
//     behind a barrier.  Items (dealt round-robin to the warps): one 128-token packed K block, <= 24 tokens of
//     the fp16 K window, one 128-token packed V block, <= 24 tokens of the fp16 V ring.
//
// Arithmetic of a packed block (128 inner x 128 outer, kivi_decode.cuh):
//     sum_i x_i * (s_i,G * c_i,o + z_i,G) = sum_i (x_i * s_i,G) * c_i,o  +  sum_i x_i * z_i,G
//   SIMT needs one LOP3 + one FFMA per code and is bound by the 16-lane ALU pipe (measured: 2.5 TB/s
//   equivalent, profiles/r01_decode_attention_ncu_summary.txt).  Here the tensor cores are an UNPACK
//   AMORTISER: the only per-code work left is isolating the field, ONE LOP3 per PAIR of codes:
//     A (16 outer x 16 inner, fp16)  codes as fp16 denormals code * 2^(P-24), P >= 4 (exact in mma.sync: Lay<>::shr)
//     B (16 inner x 8 cols,  fp16)  column (group, head, part): x_i * s_i,G split EXACTLY with two half2
//                                   instructions: hi = x*s (rounded), lo = fma(x, s, -hi) (the residual of
//                                   an fp16 product is an fp16); G query heads share the MMA (GQA is free)
//     C (16 outer x 8 cols,  fp32)  row o, columns (G(o), h, hi | lo) are the wanted sums; the other
//                                   columns are cross terms and are ignored.  Products exact, fp32 accumulate.
//   The zero term is one more MMA per 16 inner indices with exact fp16 operands (rows = z_G, cols = x_h).
#pragma once
#include <cstdlib>
#include "kivi_decode.cuh"

namespace kivi {

int make_desc(const kivi_cache_t* k, CacheDesc* d);

constexpr int kCW = 8;                 // warps per CTA
constexpr int kThreads = kCW * 32;
constexpr int kResTile = 24;           // tokens per fp16-window item (24 * 256 B = 6 KB)
constexpr int kResBytes = kResTile * kD * 2;
constexpr float kRcpSqrtD = 1.0f / 11.313708f;   // ATen: x * (1.0f / float(math.sqrt(128)))  (llama_kivi.py:339)
// probabilities are kept x 2^6 in the logits row during the V phase: exact, and it keeps the fp16 residual
// fma(p, s, -hi) of small probabilities out of the denormal range
constexpr float kProbScale = 64.f, kProbScaleInv = 1.f / 64.f;
constexpr int kScratchBytes = 256;     // per-CTA scratch of commit_unit (V token codes)

struct AttnParams {
    CacheDesc c;
    const __half* q; const __half* k_new; const __half* v_new; const __half* mask;
    __half* out; __half* dbg_logits; __half* dbg_probs;
    long long dbg_stride;
    __half* ws; long long ld;          // optional global fp16 workspace [B*H][ld] for the logits rows (long contexts)
    int t_cap, stage_bytes, spw /*stages per warp*/, hchunks, n_units, use_ws;
};

struct Sched {                          // per-step constants, identical for every unit
    int tk, r, tv, L, vhead, T, seg1;
    int n_kb, n_kr, n_vb, vr1, n_vr;
};

__device__ __forceinline__ Sched make_sched(const CacheDesc& c) {
    Sched s;
    s.tk = c.state[ST_TK]; s.r = c.state[ST_R]; s.tv = c.state[ST_TV]; s.L = c.state[ST_L]; s.vhead = c.state[ST_VHEAD];
    s.T = s.tk + s.r + 1;
    s.n_kb = cdiv(s.tk, kBlockTokens);
    s.n_kr = cdiv(s.r, kResTile);
    s.n_vb = cdiv(s.tv, kBlockTokens);
    s.seg1 = min(s.L, c.v_res_cap - s.vhead);
    s.vr1 = cdiv(s.seg1, kResTile);
    s.n_vr = s.vr1 + cdiv(s.L - s.seg1, kResTile);
    return s;
}

// items are dealt round-robin: index i of a list that starts at round-robin position `base` goes to
// warp (base + i) % kCW
__device__ __forceinline__ int rr_first(int base, int w) { return (w - base % kCW + kCW) % kCW; }
__device__ __forceinline__ int rr_count(int n, int first) { return n > first ? (n - first - 1) / kCW + 1 : 0; }

// The item stream of one warp.  Per unit, in order:
//   K phase:  KB(block = w + 8a) a < nkb;   KR(i = kr0 + 8b) b < nkr
//   V phase:  VB(block = w + 8a) a < nvb;   VR(i = vr0 + 8b) b < nvr
struct WarpPlan {
    int nkb, nkr, kr0, nvb, nvr, vr0, per_unit;
    __device__ __forceinline__ WarpPlan(const Sched& s, int w) {
        nkb = rr_count(s.n_kb, w);
        kr0 = rr_first(s.n_kb, w); nkr = rr_count(s.n_kr, kr0);
        nvb = rr_count(s.n_vb, w);
        vr0 = rr_first(s.n_vb, w); nvr = rr_count(s.n_vr, vr0);
        per_unit = nkb + nkr + nvb + nvr;
    }
};

struct Pipe {                           // a warp's private stages + its issue cursor
    uint8_t* base; uint64_t* full; int spw, stage_bytes;
    int iss_unit, iss_j, iss_n;         // next item to issue: (unit, index within the unit's list), count issued
    __device__ __forceinline__ uint8_t* stage(int m) const { return base + (size_t)(m % spw) * stage_bytes; }
    __device__ __forceinline__ void wait_full(int m) const { mbar_wait(&full[m % spw], (uint32_t)((m / spw) & 1)); }
};

// Issue the warp's next item (executed by the whole warp, copies issued by lane 0).
template <int KB, int VB>
__device__ __forceinline__ void issue_next(Pipe& pp, const AttnParams& p, const Sched& s, const WarpPlan& wp,
                                           int warp, int lane, uint64_t pol)
{
    if (pp.iss_unit >= p.n_units || wp.per_unit == 0) return;
    const CacheDesc& c = p.c;
    const int u = pp.iss_unit / p.hchunks;
    uint8_t* dst = pp.stage(pp.iss_n);
    uint64_t* bar = &pp.full[pp.iss_n % pp.spw];
    int j = pp.iss_j;
    if (lane == 0) {
        // order this warp's earlier generic-proxy reads of the stage before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (j < wp.nkb) {
            const uint32_t bb = (uint32_t)lay_block_bytes(KB, c.g);
            mbar_expect_tx(bar, bb);
            bulk_g2s(dst, c.k_store + ((int64_t)u * c.k_cap_blocks + (warp + kCW * j)) * bb, bb, bar, pol);
        } else if ((j -= wp.nkb) < wp.nkr) {
            const int t0 = (wp.kr0 + kCW * j) * kResTile, nt = min(kResTile, s.r - t0);
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.k_res + ((int64_t)u * c.R + t0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        } else if ((j -= wp.nkr) < wp.nvb) {
            const uint32_t bb = (uint32_t)lay_block_bytes(VB, c.g);
            mbar_expect_tx(bar, bb);
            bulk_g2s(dst, c.v_store + ((int64_t)u * c.v_cap_blocks + (warp + kCW * j)) * bb, bb, bar, pol);
        } else {
            j -= wp.nvb;
            const int i = wp.vr0 + kCW * j;
            int slot0, nt;
            if (i < s.vr1) { const int t0 = i * kResTile; slot0 = s.vhead + t0; nt = min(kResTile, s.seg1 - t0); }
            else { const int t0 = (i - s.vr1) * kResTile; slot0 = t0; nt = min(kResTile, s.L - s.seg1 - t0); }
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.v_res + ((int64_t)u * c.v_res_cap + slot0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        }
    }
    ++pp.iss_n;
    if (++pp.iss_j == wp.per_unit) { pp.iss_j = 0; pp.iss_unit += gridDim.x; }
}

// logits (fp16 kernel output) -> fp16 scaled, the value that enters the softmax
__device__ __forceinline__ __half scale_logit(float acc) {
    return __float2half_rn(__half2float(__float2half_rn(acc)) * kRcpSqrtD);
}

__device__ __forceinline__ uint32_t h2_as_u32(const __half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(const uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }

// One B-fragment register: column part 0 -> hi = fp16(x*s); part 1 -> lo = x*s - hi (exact).  Branch-free:
// nh = hi * (part ? -1 : 0);  b = fma(x, s, nh).
__device__ __forceinline__ uint32_t b_prep(uint32_t x2, uint32_t s2, __half2 msel) {
    const __half2 x = u32_as_h2(x2), s = u32_as_h2(s2);
    const __half2 nh = __hmul2(__hmul2(x, s), msel);
    return h2_as_u32(__hfma2(x, s, nh));
}

// exact power of two 2^(24 - P) that undoes the denormal scaling of the fields of MMA mm
template <int BITS>
__device__ __forceinline__ float inv_pos_scale(int mm) {
    return __uint_as_float((uint32_t)(127 + 24 - Lay<BITS>::bitpos(mm % Lay<BITS>::F)) << 23);
}

// column bookkeeping of the B fragments (G query heads, NG outer groups per block)
template <int G, int GS>
struct Cols {
    static constexpr int NG = 128 / GS;               // outer groups per block (g >= 32 -> NG <= 4)
    static constexpr int GPF = 4 / G;                  // groups per B fragment (8 columns = GPF x G heads x hi/lo)
    static constexpr int NF = (NG + GPF - 1) / GPF;    // B fragments per 16 inner indices
};

// ------------------------------------------------------------------------------------------------
// One packed block (128 inner x 128 outer) on the tensor cores.
//   st      : the block in shared memory (codes, then meta)
//   getx    : (chunk c, head h, &xa, &xb) -> the lane's x values (half2) of inner indices 16c+2t+{0,1} and
//             16c+2t+{8,9} for head h
//   acc[mm] : accumulators of MMA mm (outer rows 16mm .. 16mm+15): lane (g8, t) holds rows g8 / g8+8 of
//             columns 2t, 2t+1 = (group-in-fragment t / G, head t % G, hi | lo)
//   zc      : zero-term accumulator: row g8 = group min(g8 >> 1, NG-1), columns 2t, 2t+1 = head t % G
// ------------------------------------------------------------------------------------------------
template <int BITS, int G, int GS, class XF>
__device__ __forceinline__ void mma_block(const uint8_t* st, XF&& getx, float (&acc)[8][4], float (&zc)[4], int lane)
{
    using L = Lay<BITS>;
    using CL = Cols<G, GS>;
    constexpr int NG = CL::NG, GPF = CL::GPF, NF = CL::NF;
    const int g8 = lane >> 2, t = lane & 3;
    const int hb = (g8 % (2 * G)) >> 1;                 // head of this lane's B column
    const int gi = g8 / (2 * G);                        // group-in-fragment of this lane's B column
    const int gz = min(g8 >> 1, NG - 1);                // group of this lane's A rows in the zero-term MMA
    const __half2 msel = (g8 & 1) ? __float2half2_rn(-1.f) : __float2half2_rn(0.f);
    const uint8_t* meta = st + L::kCodeBytes + t * 16;
    constexpr uint32_t kField = ((1u << BITS) - 1u) * 0x00010001u;
    #pragma unroll 2
    for (int c = 0; c < 8; ++c) {
        uint32_t xa, xb;
        getx(c, hb, xa, xb);
        // {z(2t,2t+1), s(2t,2t+1), z(2t+8,2t+9), s(2t+8,2t+9)} of group gz: as is, the A operand of the zero-term MMA
        const uint4 mz = *reinterpret_cast<const uint4*>(meta + (c * NG + gz) * 64);
        mma_16816(zc, mz.x, mz.y, mz.z, mz.w, xa, xb);
        uint32_t b0[NF], b1[NF];
        if (G == 1 && NF == 1) {                        // the B column's group is gz
            b0[0] = b_prep(xa, mz.y, msel); b1[0] = b_prep(xb, mz.w, msel);
        } else {
            #pragma unroll
            for (int f = 0; f < NF; ++f) {
                const int grp = min(f * GPF + gi, NG - 1);
                const uint4 ms = *reinterpret_cast<const uint4*>(meta + (c * NG + grp) * 64);
                b0[f] = b_prep(xa, ms.y, msel); b1[f] = b_prep(xb, ms.w, msel);
            }
        }
        #pragma unroll
        for (int sl = 0; sl < L::kSlabs; ++sl) {
            const uint4 w4 = *reinterpret_cast<const uint4*>(st + (c * L::kSlabs + sl) * 512 + lane * 16);
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
            uint32_t wl4[4], wr4[4], wr6[4], wr8[4];    // the shifted copies a bit width needs (the others fold away)
            #pragma unroll
            for (int r = 0; r < 4; ++r) { wl4[r] = w[r] << 4; wr4[r] = w[r] >> 4; wr6[r] = w[r] >> 6; wr8[r] = w[r] >> 8; }
            #pragma unroll
            for (int j = 0; j < L::F; ++j) {
                uint32_t a[4];
                #pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int sh = L::shr(j);
                    const uint32_t src = sh == -4 ? wl4[r] : sh == 0 ? w[r] : sh == 4 ? wr4[r] : sh == 6 ? wr6[r] : wr8[r];
                    a[r] = src & (kField << L::bitpos(j));
                }
                const int mm = sl * L::F + j;
                const int f = ((16 * mm) / GS) / GPF;
                mma_16816(acc[mm], a[0], a[1], a[2], a[3], b0[f], b1[f]);
            }
        }
    }
}

// The lane's zero term for every outer group: Z[head t % G][grp] lives in the lanes with g8 = 2 * grp.
template <int G, int GS>
__device__ __forceinline__ void gather_z(const float (&zc)[4], int lane, float (&zsel)[Cols<G, GS>::NG]) {
    #pragma unroll
    for (int grp = 0; grp < Cols<G, GS>::NG; ++grp)
        zsel[grp] = __shfl_sync(0xffffffffu, zc[0], 8 * grp + (lane & 3));
}

// Hand every (outer row, value) this lane owns to `emit`: value = ((hi + lo) * 2^(24-P) + Z) * post for its head
// t % G.  Lane (g8, t) owns rows g8 / g8+8 of the MMAs whose group-in-fragment is t / G.
template <int BITS, int G, int GS, class EF>
__device__ __forceinline__ void finalize(const float (&acc)[8][4], const float (&zsel)[Cols<G, GS>::NG], int lane,
                                         float post, EF&& emit)
{
    constexpr int GPF = Cols<G, GS>::GPF;
    const int g8 = lane >> 2, t = lane & 3;
    if (G == 1 && GS == 32) {
        // lane t owns MMAs 2t and 2t+1 (group t): pick them with selects instead of 8 predicated copies of the tail
        float lo[4], hi[4];
        #pragma unroll
        for (int e = 0; e < 4; ++e) {
            lo[e] = t == 0 ? acc[0][e] : t == 1 ? acc[2][e] : t == 2 ? acc[4][e] : acc[6][e];
            hi[e] = t == 0 ? acc[1][e] : t == 1 ? acc[3][e] : t == 2 ? acc[5][e] : acc[7][e];
        }
        const float sl = (t == 0 ? inv_pos_scale<BITS>(0) : t == 1 ? inv_pos_scale<BITS>(2) : t == 2 ? inv_pos_scale<BITS>(4) : inv_pos_scale<BITS>(6)) * post;
        const float sh = (t == 0 ? inv_pos_scale<BITS>(1) : t == 1 ? inv_pos_scale<BITS>(3) : t == 2 ? inv_pos_scale<BITS>(5) : inv_pos_scale<BITS>(7)) * post;
        const float zt = (t == 0 ? zsel[0] : t == 1 ? zsel[1] : t == 2 ? zsel[2] : zsel[3]) * post;
        const int o = 32 * t + g8;
        emit(o, fmaf(lo[0] + lo[1], sl, zt));
        emit(o + 8, fmaf(lo[2] + lo[3], sl, zt));
        emit(o + 16, fmaf(hi[0] + hi[1], sh, zt));
        emit(o + 24, fmaf(hi[2] + hi[3], sh, zt));
    } else {
        const int gi_l = t / G;
        #pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
            const int grp = (16 * mm) / GS;
            if (grp % GPF == gi_l) {
                const float sc = inv_pos_scale<BITS>(mm) * post, zt = zsel[grp] * post;
                emit(16 * mm + g8, fmaf(acc[mm][0] + acc[mm][1], sc, zt));
                emit(16 * mm + g8 + 8, fmaf(acc[mm][2] + acc[mm][3], sc, zt));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cache data movement of one unit (models/llama_kivi.py:343-356, :386-399); cold path, kept out of line.
// Executed by ALL threads of the CTA (it contains CTA barriers on the K-flush path, taken uniformly).
//   scratch: kScratchBytes private to this call; flush_scratch: >= 4 KB, free after a CTA barrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float q_code(float x, float mnf, float scf, float maxq) {
    const __half t1 = __float2half_rn(x - mnf);
    const __half t2 = __float2half_rn(__fdiv_rn(__half2float(t1), scf));
    return fminf(fmaxf(__half2float(t2), 0.f), maxq);       // quant/new_pack.py:238-241 (rint follows)
}

template <int KB, int VB>
__device__ __noinline__ void commit_unit(const AttnParams& p, const Sched& s, int u, int tid, uint8_t* scratch,
                                         uint8_t* flush_scratch)
{
    const CacheDesc& c = p.c;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = c.g;
    // ---- V: v_new joins the ring; if the window would exceed R, its oldest token is quantised per token
    if (tid < kD / 8)
        reinterpret_cast<uint4*>(c.v_res + ((int64_t)u * c.v_res_cap + (s.vhead + s.L) % c.v_res_cap) * kD)[tid] =
            __ldg(reinterpret_cast<const uint4*>(p.v_new + (int64_t)u * kD) + tid);
    if (s.L + 1 > c.R && warp == 1) {
        constexpr int F = 16 / VB, kSlabRows = 16 * F, kSlabs = 128 / kSlabRows;
        const float maxq = (float)((1 << VB) - 1);
        const __half* src = c.v_res + ((int64_t)u * c.v_res_cap + s.vhead) * kD;
        const int bb = lay_block_bytes(VB, g);
        uint8_t* blk = c.v_store + ((int64_t)u * c.v_cap_blocks + s.tv / kBlockTokens) * bb;
        const int inner = s.tv % kBlockTokens;
        const uint2 raw = *reinterpret_cast<const uint2*>(src + lane * 4);              // 4 channels per lane
        const __half2* rh = reinterpret_cast<const __half2*>(&raw);
        const float2 x01 = __half22float2(rh[0]), x23 = __half22float2(rh[1]);
        const float x[4] = {x01.x, x01.y, x23.x, x23.y};
        float mnf = fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), mxf = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
        const int lpg = g / 4;                                                          // lanes per group
        for (int o = 1; o < lpg; o <<= 1) {
            mnf = fminf(mnf, __shfl_xor_sync(0xffffffffu, mnf, o));
            mxf = fmaxf(mxf, __shfl_xor_sync(0xffffffffu, mxf, o));
        }
        const __half d16 = __float2half_rn(mxf - mnf);
        const __half sc = __float2half_rn(__fdiv_rn(__half2float(d16), maxq));
        const float scf = __half2float(sc);
        uint32_t four = 0;
        #pragma unroll
        for (int e = 0; e < 4; ++e) four |= (uint32_t)__float2int_rn(q_code(x[e], mnf, scf, maxq)) << (8 * e);
        reinterpret_cast<uint32_t*>(scratch)[lane] = four;                              // codes[channel] as bytes
        if (lane % lpg == 0) {
            *reinterpret_cast<__half*>(blk + lay_scale_off(VB, g, inner, lane / lpg)) = sc;
            *reinterpret_cast<__half*>(blk + lay_zero_off(VB, g, inner, lane / lpg)) = __float2half_rn(mnf);
        }
        __syncwarp();
        if (lane < kSlabs * 16) {                                                       // one 16-bit half-word per lane
            const int sl = lane >> 4, row = lane & 15;
            uint32_t hw = 0;
            #pragma unroll
            for (int j = 0; j < F; ++j) hw |= (uint32_t)scratch[sl * kSlabRows + 16 * j + row] << (VB * j);
            *reinterpret_cast<uint16_t*>(blk + lay_word_off(VB, inner, sl * kSlabRows + row) + 2 * (inner & 1)) = (uint16_t)hw;
        }
        __syncwarp();
    }
    // ---- K: k_new joins the window, or completes it -> quantise the R tokens per channel
    if (s.r + 1 < c.R) {
        if (tid >= 16 && tid < 16 + kD / 8)
            reinterpret_cast<uint4*>(c.k_res + ((int64_t)u * c.R + s.r) * kD)[tid - 16] =
                __ldg(reinterpret_cast<const uint4*>(p.k_new + (int64_t)u * kD) + (tid - 16));
    } else {
        constexpr int F = 16 / KB, kSlabRows = 16 * F, kSlabs = 128 / kSlabRows;
        const float maxq = (float)((1 << KB) - 1);
        const int bb = lay_block_bytes(KB, g);
        uint8_t* ub = c.k_store + (int64_t)u * c.k_cap_blocks * bb;
        const __half* win = c.k_res + (int64_t)u * c.R * kD;
        const __half* knew = p.k_new + (int64_t)u * kD;
        __half2* stats = reinterpret_cast<__half2*>(flush_scratch);                     // [R/g][128] (scale, mn)
        auto tokval = [&](int t, int d) -> float {
            return __half2float(t < c.R - 1 ? win[(int64_t)t * kD + d] : knew[d]);
        };
        __syncthreads();                                                                // flush_scratch is free now
        for (int w = tid; w < kD * (c.R / g); w += kThreads) {
            const int d = w % kD, grp = w / kD;
            float mnf = tokval(grp * g, d), mxf = mnf;
            for (int i = 1; i < g; ++i) { const float x = tokval(grp * g + i, d); mnf = fminf(mnf, x); mxf = fmaxf(mxf, x); }
            const __half d16 = __float2half_rn(mxf - mnf);
            const __half sc = __float2half_rn(__fdiv_rn(__half2float(d16), maxq));
            const __half mn = __float2half_rn(mnf);
            stats[w] = __halves2half2(sc, mn);
            const int tok = s.tk + grp * g;
            uint8_t* blk = ub + (int64_t)(tok / kBlockTokens) * bb;
            *reinterpret_cast<__half*>(blk + lay_scale_off(KB, g, d, (tok % kBlockTokens) / g)) = sc;
            *reinterpret_cast<__half*>(blk + lay_zero_off(KB, g, d, (tok % kBlockTokens) / g)) = mn;
        }
        __syncthreads();
        const int nblk = max(1, c.R / kBlockTokens);                                    // R in {32, 64, 128, 256}
        for (int bi = 0; bi < nblk; ++bi) {
            const int tb = s.tk + bi * kBlockTokens;                                    // first flushed token of this block
            const int o0 = tb % kBlockTokens, cnt = min(c.R, kBlockTokens);
            uint8_t* blk = ub + (int64_t)(tb / kBlockTokens) * bb;
            for (int id = tid; id < kD * kSlabs * 16; id += kThreads) {                 // one 16-bit half-word per item
                const int d = id % kD, row = (id / kD) % 16, sl = id / (kD * 16);
                uint16_t* hp = reinterpret_cast<uint16_t*>(blk + lay_word_off(KB, d, sl * kSlabRows + row) + 2 * (d & 1));
                uint32_t hw = *hp;
                #pragma unroll 1
                for (int j = 0; j < F; ++j) {
                    const int o = sl * kSlabRows + 16 * j + row;
                    if (o < o0 || o >= o0 + cnt) continue;
                    const int tl = (tb - s.tk) + (o - o0);                              // token index within the window
                    const float2 sm = __half22float2(stats[(tl / g) * kD + d]);
                    const uint32_t code = (uint32_t)__float2int_rn(q_code(tokval(tl, d), sm.y, sm.x, maxq));
                    hw = (hw & ~(((1u << KB) - 1u) << (KB * j))) | (code << (KB * j));
                }
                *hp = (uint16_t)hw;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int KB, int VB, int G, int GS, bool WS>
__global__ void __launch_bounds__(kThreads, 2)
attention_kernel(const AttnParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const CacheDesc& c = p.c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int n_stages = kCW * p.spw;
    using CL = Cols<G, GS>;
    constexpr int NG = CL::NG, GPF = CL::GPF;

    // carve shared memory
    uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * p.stage_bytes);
    uint8_t* ptr = smem + (((size_t)n_stages * (p.stage_bytes + 8) + 127) & ~(size_t)127);
    uint2* q2 = reinterpret_cast<uint2*>(ptr); ptr += G * 32 * 8;    // [G][8 chunks][4 t] {q(16c+2t,+1), q(16c+2t+8,+9)} half2 pairs
    float* qlin = reinterpret_cast<float*>(ptr); ptr += G * kD * 4;  // [G][128] q in fp32, channel order
    float* stats = reinterpret_cast<float*>(ptr); ptr += 16 * 4;     // softmax block-reduce scratch
    float* pnew = reinterpret_cast<float*>(ptr); ptr += 16 * 4;      // probability of the new token, per head
    uint8_t* scratch = ptr; ptr += kScratchBytes;
    float* red = reinterpret_cast<float*>(ptr);                      // [kCW][G][2][128] partial outputs
    // logits row(s): [G][t_cap] scaled logits, then probabilities x 2^6 -- in shared memory (aliasing `red`, which
    // is only written after the last read of the probabilities) or in the caller's global workspace
    const long long lg_stride = WS ? p.ld : (long long)p.t_cap;

    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) mbar_init(&full_all[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const Sched s = make_sched(c);
    const WarpPlan wp(s, warp);
    const uint64_t pol = policy_evict_first();
    Pipe pp;
    pp.base = smem + (size_t)warp * p.spw * p.stage_bytes;
    pp.full = full_all + warp * p.spw;
    pp.spw = p.spw; pp.stage_bytes = p.stage_bytes;
    pp.iss_unit = blockIdx.x; pp.iss_j = 0; pp.iss_n = 0;
    for (int i = 0; i < p.spw; ++i) issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);

    const int ratio = c.H / c.Hkv;
    const int h_l = t4 % G;                                         // the head this lane finalises
    int m = 0;                                                      // items consumed so far by this warp
    #pragma unroll 1
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int b = u / c.Hkv;
        const int uq0 = u * ratio + hc * G;                         // first query head row of this chunk
        __half* lg = reinterpret_cast<__half*>(red);
        if (WS) lg = p.ws + (int64_t)uq0 * p.ld;

        // -- stage q: fp32 in channel order (window items) and the half2 pairs of the B fragments
        for (int i = tid; i < G * kD; i += kThreads) qlin[i] = __half2float(p.q[(int64_t)uq0 * kD + i]);
        for (int i = tid; i < G * 32; i += kThreads) {
            const int h = i >> 5, cc = (i >> 2) & 7, tt = i & 3;
            const uint32_t* qh = reinterpret_cast<const uint32_t*>(p.q + (int64_t)(uq0 + h) * kD + 16 * cc + 2 * tt);
            q2[i] = make_uint2(__ldg(qh), __ldg(qh + 4));
        }
        __syncthreads();

        // ================= K phase =================
        #pragma unroll 1
        for (int a = 0; a < wp.nkb; ++a) {
            const int blk = warp + kCW * a;
            float acc[8][4];
            float zc[4] = {0.f, 0.f, 0.f, 0.f};
            #pragma unroll
            for (int mm = 0; mm < 8; ++mm)
                #pragma unroll
                for (int e = 0; e < 4; ++e) acc[mm][e] = 0.f;
            pp.wait_full(m);
            mma_block<KB, G, GS>(pp.stage(m), [&](int cc, int h, uint32_t& xa, uint32_t& xb) {
                const uint2 v = q2[(h * 8 + cc) * 4 + t4];
                xa = v.x; xb = v.y;
            }, acc, zc, lane);
            __syncwarp();
            issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
            ++m;
            float zsel[NG];
            gather_z<G, GS>(zc, lane, zsel);
            __half* row = lg + (int64_t)h_l * lg_stride + blk * kBlockTokens;
            const int nvalid = s.tk - blk * kBlockTokens;             // < 128 only in the last block when R < 128
            finalize<KB, G, GS>(acc, zsel, lane, 1.f, [&](int o, float v) {
                if (o < nvalid) row[o] = scale_logit(v);
            });
        }
        // fp16 K window
        {
            const int part = lane & 7, tok = lane >> 3;
            #pragma unroll 1
            for (int bq = 0; bq < wp.nkr; ++bq) {
                const int i = wp.kr0 + kCW * bq;
                const int t0 = i * kResTile, nt = min(kResTile, s.r - t0);
                pp.wait_full(m);
                const uint8_t* st = pp.stage(m);
                #pragma unroll 1
                for (int ts = 0; ts < nt; ts += 4) {
                    const int t = ts + tok;
                    float sum[G];
                    #pragma unroll
                    for (int h = 0; h < G; ++h) sum[h] = 0.f;
                    if (t < nt) {
                        const uint4 a4 = *reinterpret_cast<const uint4*>(st + t * 256 + part * 16);
                        const uint4 b4 = *reinterpret_cast<const uint4*>(st + t * 256 + 128 + part * 16);
                        const __half2* ah = reinterpret_cast<const __half2*>(&a4);
                        const __half2* bh = reinterpret_cast<const __half2*>(&b4);
                        #pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
                            #pragma unroll
                            for (int h = 0; h < G; ++h) {
                                const float2 qa = *reinterpret_cast<const float2*>(qlin + h * kD + part * 8 + 2 * e);
                                const float2 qb = *reinterpret_cast<const float2*>(qlin + h * kD + 64 + part * 8 + 2 * e);
                                sum[h] = fmaf(qa.x, fa.x, sum[h]); sum[h] = fmaf(qa.y, fa.y, sum[h]);
                                sum[h] = fmaf(qb.x, fb.x, sum[h]); sum[h] = fmaf(qb.y, fb.y, sum[h]);
                            }
                        }
                    }
                    #pragma unroll
                    for (int h = 0; h < G; ++h) {
                        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
                        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
                        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 4);
                        if (part == 0 && t < nt)
                            lg[(int64_t)h * lg_stride + s.tk + t0 + t] = scale_logit(sum[h]);
                    }
                }
                __syncwarp();
                issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
                ++m;
            }
            // the new token (k_new, not yet in the window): one warp, plain loads
            if ((s.n_kb + s.n_kr) % kCW == warp) {
                const uint2 kv = __ldg(reinterpret_cast<const uint2*>(p.k_new + (int64_t)u * kD) + lane);
                const __half2* kh = reinterpret_cast<const __half2*>(&kv);
                const float2 k01 = __half22float2(kh[0]), k23 = __half22float2(kh[1]);
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float4 qv = *reinterpret_cast<const float4*>(qlin + h * kD + lane * 4);
                    float sum = qv.x * k01.x;
                    sum = fmaf(qv.y, k01.y, sum); sum = fmaf(qv.z, k23.x, sum); sum = fmaf(qv.w, k23.y, sum);
                    sum = warp_sum(sum);
                    if (lane == 0) lg[(int64_t)h * lg_stride + s.T - 1] = scale_logit(sum);
                    if (lane >= 1 && lane < 8) lg[(int64_t)h * lg_stride + s.T - 1 + lane] = __float2half_rn(-65504.f);   // pad: exp -> 0
                }
            }
        }
        __syncthreads();

        // ================= softmax (fp32), one block-wide reduction per head =================
        #pragma unroll 1
        for (int h = 0; h < G; ++h) {
            __half* row = lg + (int64_t)h * lg_stride;
            const bool slow = p.mask || p.dbg_logits || p.dbg_probs;
            const int nvec = (s.T + 7) >> 3;                        // the row is padded with -65504 up to 8 * nvec
            float ml = -INFINITY, sl = 0.f;
            if (slow) {
                for (int t = tid; t < s.T; t += kThreads) {
                    __half v = row[t];
                    if (p.mask) {
                        v = __hadd_rn(v, p.mask[(int64_t)b * s.T + t]);                   // llama_kivi.py:369
                        if (__half2float(v) < -65504.f) v = __float2half_rn(-65504.f);    // :370-372 (max with finfo.min)
                        row[t] = v;
                    }
                    if (p.dbg_logits) p.dbg_logits[(int64_t)(uq0 + h) * p.dbg_stride + t] = v;
                    ml = fmaxf(ml, __half2float(v));
                }
                for (int t = tid; t < s.T; t += kThreads) sl += __expf(__half2float(row[t]) - ml);
            } else {
                for (int i = tid; i < nvec; i += kThreads) {
                    const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * i);
                    const __half2* hh = reinterpret_cast<const __half2*>(&u);
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hh[e]); ml = fmaxf(ml, fmaxf(f.x, f.y)); }
                }
                for (int i = tid; i < nvec; i += kThreads) {
                    const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * i);
                    const __half2* hh = reinterpret_cast<const __half2*>(&u);
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hh[e]); sl += __expf(f.x - ml) + __expf(f.y - ml); }
                }
            }
            // (max, sum) pairs: warp, then block
            #pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const float mo = __shfl_xor_sync(0xffffffffu, ml, o), so = __shfl_xor_sync(0xffffffffu, sl, o);
                const float mn = fmaxf(ml, mo);
                sl = (ml == -INFINITY ? 0.f : sl * __expf(ml - mn)) + (mo == -INFINITY ? 0.f : so * __expf(mo - mn));
                ml = mn;
            }
            if (lane == 0) { stats[warp] = ml; stats[8 + warp] = sl; }
            __syncthreads();
            float M = stats[0];
            #pragma unroll
            for (int w = 1; w < kCW; ++w) M = fmaxf(M, stats[w]);
            float S = 0.f;
            #pragma unroll
            for (int w = 0; w < kCW; ++w) S += stats[w] == -INFINITY ? 0.f : stats[8 + w] * __expf(stats[w] - M);
            if (slow) {
                for (int t = tid; t < s.T; t += kThreads) {
                    const __half pr = __float2half_rn(__fdiv_rn(__expf(__half2float(row[t]) - M), S));   // :375
                    row[t] = __float2half_rn(__half2float(pr) * kProbScale);                            // exact
                    if (p.dbg_probs) p.dbg_probs[(int64_t)(uq0 + h) * p.dbg_stride + t] = pr;
                    if (t == s.T - 1) pnew[h] = __half2float(pr);
                }
            } else {
                const float rS = __frcp_rn(S);
                const __half2 k64 = __float2half2_rn(kProbScale);
                for (int i = tid; i < nvec; i += kThreads) {
                    uint4 u = *reinterpret_cast<const uint4*>(row + 8 * i);
                    __half2* hh = reinterpret_cast<__half2*>(&u);
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(hh[e]);
                        const float e0 = __expf(f.x - M), e1 = __expf(f.y - M);
                        // e / S, correctly rounded: one Newton step on the quotient (same result as __fdiv_rn for normal
                        // operands, without its range checks); then fp16 (:375)
                        float q0 = e0 * rS, q1 = e1 * rS;
                        q0 = fmaf(fmaf(-q0, S, e0), rS, q0);
                        q1 = fmaf(fmaf(-q1, S, e1), rS, q1);
                        const __half2 pr = __floats2half2_rn(q0, q1);
                        if (8 * i + 2 * e == s.T - 1) pnew[h] = __low2float(pr);
                        if (8 * i + 2 * e + 1 == s.T - 1) pnew[h] = __high2float(pr);
                        hh[e] = __hmul2(pr, k64);                   // exact
                    }
                    *reinterpret_cast<uint4*>(row + 8 * i) = u;
                }
            }
            if (G > 1) __syncthreads();                             // stats are reused by the next head
        }
        __syncthreads();

        // ================= V phase =================
        float acc[8][4];                                            // packed part, over all of this warp's blocks
        float zc[4] = {0.f, 0.f, 0.f, 0.f};
        float orr[G][4];                                            // fp16 window part: lane = 4 channels
        #pragma unroll
        for (int mm = 0; mm < 8; ++mm)
            #pragma unroll
            for (int e = 0; e < 4; ++e) acc[mm][e] = 0.f;
        #pragma unroll
        for (int h = 0; h < G; ++h)
            #pragma unroll
            for (int e = 0; e < 4; ++e) orr[h][e] = 0.f;
        #pragma unroll 1
        for (int a = 0; a < wp.nvb; ++a) {
            const int blk = warp + kCW * a;
            const int t0 = blk * kBlockTokens, nt = s.tv - t0;      // nt >= 128 except in the last block
            pp.wait_full(m);
            const __half* prow = lg + t0 + 2 * t4;
            mma_block<VB, G, GS>(pp.stage(m), [&](int cc, int h, uint32_t& xa, uint32_t& xb) {
                const __half* pr = prow + (int64_t)h * lg_stride + 16 * cc;
                xa = *reinterpret_cast<const uint32_t*>(pr);
                xb = *reinterpret_cast<const uint32_t*>(pr + 8);
                if (nt < kBlockTokens) {                            // tokens beyond the packed length belong to the window
                    const int i0 = 16 * cc + 2 * t4;
                    if (i0 >= nt) xa = 0u; else if (i0 + 1 >= nt) xa &= 0xFFFFu;
                    if (i0 + 8 >= nt) xb = 0u; else if (i0 + 9 >= nt) xb &= 0xFFFFu;
                }
            }, acc, zc, lane);
            __syncwarp();
            issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
            ++m;
        }
        #pragma unroll 1
        for (int bq = 0; bq < wp.nvr; ++bq) {
            const int i = wp.vr0 + kCW * bq;
            int l0, nt;                                             // logical index of the item's first token
            if (i < s.vr1) { l0 = i * kResTile; nt = min(kResTile, s.seg1 - l0); }
            else { const int tt0 = (i - s.vr1) * kResTile; l0 = s.seg1 + tt0; nt = min(kResTile, s.L - s.seg1 - tt0); }
            pp.wait_full(m);
            const uint8_t* st = pp.stage(m);
            #pragma unroll 2
            for (int t = 0; t < nt; ++t) {
                const uint2 vv = *reinterpret_cast<const uint2*>(st + t * 256 + lane * 8);
                const __half2* vh = reinterpret_cast<const __half2*>(&vv);
                const float2 v01 = __half22float2(vh[0]), v23 = __half22float2(vh[1]);
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float pr = __half2float(lg[(int64_t)h * lg_stride + s.tv + l0 + t]);   // x 2^6, undone below
                    orr[h][0] = fmaf(pr, v01.x, orr[h][0]); orr[h][1] = fmaf(pr, v01.y, orr[h][1]);
                    orr[h][2] = fmaf(pr, v23.x, orr[h][2]); orr[h][3] = fmaf(pr, v23.y, orr[h][3]);
                }
            }
            __syncwarp();
            issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
            ++m;
        }
        float zsel[NG];
        gather_z<G, GS>(zc, lane, zsel);
        if (!WS) __syncthreads();                                   // everyone is done reading the probabilities (red aliases them)
        {
            float* rq = red + ((size_t)(warp * G + h_l) * 2 + 0) * kD;
            finalize<VB, G, GS>(acc, zsel, lane, kProbScaleInv, [&](int o, float v) { rq[o] = v; });
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                float* rr = red + ((size_t)(warp * G + h) * 2 + 1) * kD;
                *reinterpret_cast<float4*>(rr + lane * 4) =
                    make_float4(orr[h][0] * kProbScaleInv, orr[h][1] * kProbScaleInv, orr[h][2] * kProbScaleInv, orr[h][3] * kProbScaleInv);
            }
        }
        __syncthreads();
        for (int i = tid; i < G * kD; i += kThreads) {
            const int h = i / kD, d = i % kD;
            float q_sum = 0.f, r_sum = 0.f;
            #pragma unroll
            for (int w = 0; w < kCW; ++w) {
                q_sum += red[((size_t)(w * G + h) * 2 + 0) * kD + d];
                r_sum += red[((size_t)(w * G + h) * 2 + 1) * kD + d];
            }
            r_sum = fmaf(pnew[h], __half2float(p.v_new[(int64_t)u * kD + d]), r_sum);
            __half o = __float2half_rn(r_sum);                                          // llama_kivi.py:380 / :384
            if (s.tv > 0) o = __hadd_rn(__float2half_rn(q_sum), o);                     // :382-384
            p.out[(int64_t)(uq0 + h) * kD + d] = o;
        }
        if (hc == 0) commit_unit<KB, VB>(p, s, u, tid, scratch, reinterpret_cast<uint8_t*>(red));
        __syncthreads();                                            // q2 / qlin / lg / red are reused by the next unit
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0, g_max_smem = 0;

template <int KB, int VB, int G, int GS>
static int launch_attention(AttnParams& p, int max_kv_len, cudaStream_t st)
{
    const CacheDesc& c = p.c;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    int stage = max(lay_block_bytes(KB, c.g), lay_block_bytes(VB, c.g));
    stage = max(stage, kResBytes);
    p.stage_bytes = (stage + 127) / 128 * 128;
    const int red_bytes = kCW * G * 2 * kD * 4;
    const int base = 512 /*barriers, alignment*/ + G * 32 * 8 + G * kD * 4 + 128 + kScratchBytes;
    // logits rows in shared memory when they fit next to >= 2 stages per warp, else in the caller's workspace
    const int t_need = (max_kv_len + 8 + 63) / 64 * 64;
    p.t_cap = max(red_bytes / (2 * G), t_need);
    int fixed = base + G * p.t_cap * 2;
    p.use_ws = 0;
    const bool force_ws = p.ws && getenv("KIVI_FORCE_WORKSPACE");       // tests: exercise the workspace path at small sizes
    if (force_ws || (g_max_smem - fixed) / (kCW * p.stage_bytes) < 2) {
        if (!p.ws) return KIVI_ERR_CAPACITY;
        if (p.ld < max_kv_len + 8) return KIVI_ERR_CAPACITY;
        p.use_ws = 1;
        p.t_cap = red_bytes / (2 * G);
        fixed = base + red_bytes;
    }
    int ctas = 2;
    p.spw = min(4, (g_max_smem / ctas - 1024 - fixed) / (kCW * p.stage_bytes));
    if (p.spw < 2) {
        ctas = 1;
        p.spw = min(4, (g_max_smem - fixed) / (kCW * p.stage_bytes));
    }
    if (p.spw < 1) return KIVI_ERR_CAPACITY;
    const size_t smem = (size_t)kCW * p.spw * p.stage_bytes + fixed;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<KB, VB, G, GS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attention_kernel<KB, VB, G, GS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = min(p.n_units, g_num_sms * ctas);
    if (p.use_ws) attention_kernel<KB, VB, G, GS, true><<<grid, kThreads, smem, st>>>(p);
    else          attention_kernel<KB, VB, G, GS, false><<<grid, kThreads, smem, st>>>(p);
    return post_launch();
}

template <int KB, int VB>
static int dispatch_attention(AttnParams& p, int G, int max_kv_len, cudaStream_t st)
{
    #define KIVI_GS(GS_)                                                                  \
        if (p.c.g == GS_) {                                                               \
            if (G == 4) return launch_attention<KB, VB, 4, GS_>(p, max_kv_len, st);       \
            if (G == 2) return launch_attention<KB, VB, 2, GS_>(p, max_kv_len, st);       \
            return launch_attention<KB, VB, 1, GS_>(p, max_kv_len, st);                   \
        }
    KIVI_GS(32)
    KIVI_GS(64)
    KIVI_GS(128)
    #undef KIVI_GS
    return KIVI_ERR_GROUP;
}

}  // namespace kivi
