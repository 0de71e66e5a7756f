// kivi_decode.cu -- fused KIVI decode attention over the blocked cache (sm_100a).
//
// One launch per layer per step replaces the ~30 launches of the reference's decode branch
// (models/llama_kivi.py:314-399): q.Kq^T with in-register dequantisation, the fp16 K window, scale,
// mask, fp32 softmax, p.Vq, the fp16 V window, the fp16 add, and the per-unit cache data movement
// (window append, K flush, V token pack).  Rounding points of the reference are reproduced
// (fp16 logits -> fp16 scale -> fp32 softmax -> fp16 probs -> fp16 partial outputs -> fp16 add).
//
// Execution model
//   * persistent grid, CTA c handles units c, c+grid, ...; unit = (b, kv-head, chunk of G query
//     heads) -- packed bytes are read once per KV head for all G heads (GQA);
//   * 8 warps per CTA, each with S PRIVATE shared-memory stages.  A warp streams its own work items
//     HBM -> shared memory with 1-D bulk copies (cp.async.bulk, the TMA engine; SASS UBLKCP) that
//     complete on the stage's mbarrier: right after it has consumed the item in a stage, its elected
//     lane issues the copy of the item S positions ahead into the same stage.  The issue cursor runs
//     ahead across the K -> softmax -> V phases and across units, so HBM never idles behind a
//     barrier; a warp walks the rounds of its own stages in order, so the mbarrier parity can never
//     alias a round it has not reached.  ~6 KB work items, dealt round-robin to the warps:
//       KQ  quarter of a 512-token K tile (4 blocks x 32 channels): lane = (row parity, block, cell),
//           one 64/128-bit LDS of codes + one 32-bit LDS of (scale, zero) per row, then per element
//           1 LOP3 (denormal unpack, kivi_common.cuh) + 1 FFMA per query head, fp32 accumulate;
//       KR  <= 24 tokens of the fp16 K window;          VQ  128 tokens of packed V;
//       VR  <= 24 tokens of the fp16 V ring.
//   * G == 1 (MHA) kernels fit 2 CTAs per SM: while one CTA sits in a phase barrier or the softmax,
//     the other keeps the FMA/ALU pipes and the TMA queue busy.
#pragma once
#include "kivi_decode.cuh"

namespace kivi {

int make_desc(const kivi_cache_t* k, CacheDesc* d);

constexpr int kCW = 8;                 // warps per CTA
constexpr int kThreads = kCW * 32;
constexpr int kVTile = 128;            // tokens per VQ item
constexpr int kResTile = 24;           // tokens per KR / VR item (24 * 256 B = 6 KB)
constexpr int kResBytes = kResTile * kD * 2;
constexpr float kRcpSqrtD = 1.0f / 11.313708f;   // ATen: x * (1.0f / float(math.sqrt(128)))  (llama_kivi.py:339)

struct DecodeParams {
    CacheDesc c;
    const __half* q; const __half* k_new; const __half* v_new; const __half* mask;
    __half* out; __half* dbg_logits; __half* dbg_probs;
    long long dbg_stride;
    int t_cap, stage_bytes, spw /*stages per warp*/, kb_stride, hchunks, n_units;
};

struct Sched {                          // per-step constants, identical for every unit
    int tk, r, tv, L, vhead, T, seg1;
    int n_ktiles, n_kr, n_vq, vr1, n_vr;
};

__device__ __forceinline__ Sched make_sched(const CacheDesc& c) {
    Sched s;
    s.tk = c.state[ST_TK]; s.r = c.state[ST_R]; s.tv = c.state[ST_TV]; s.L = c.state[ST_L]; s.vhead = c.state[ST_VHEAD];
    s.T = s.tk + s.r + 1;
    s.n_ktiles = cdiv(cdiv(s.tk, kBlockTokens), 4);
    s.n_kr = cdiv(s.r, kResTile);
    s.n_vq = cdiv(s.tv, kVTile);
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
//   K phase:  KQ(tile = w + 8a, quarter) a < ntile, quarter < 4;   KR(i = kr0 + 8b) b < nkr
//   V phase:  VQ(i = w + 8a) a < nvq;                               VR(i = vr0 + 8b) b < nvr
struct WarpPlan {
    int ntile, nkr, kr0, nvq, nvr, vr0, per_unit;
    __device__ __forceinline__ WarpPlan(const Sched& s, int w) {
        ntile = rr_count(s.n_ktiles, w);
        kr0 = rr_first(s.n_ktiles, w); nkr = rr_count(s.n_kr, kr0);
        nvq = rr_count(s.n_vq, w);
        vr0 = rr_first(s.n_vq, w); nvr = rr_count(s.n_vr, vr0);
        per_unit = 4 * ntile + nkr + nvq + nvr;
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
__device__ __forceinline__ void issue_next(Pipe& pp, const DecodeParams& p, const Sched& s, const WarpPlan& wp,
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
        if (j < 4 * wp.ntile) {
            const int QB = k_q_bytes(KB, c.g);
            const int tile = warp + kCW * (j >> 2), qt = j & 3;
            const int b0 = tile * 4, nb = min(4, cdiv(s.tk, kBlockTokens) - b0);
            mbar_expect_tx(bar, (uint32_t)(nb * QB));
            const uint8_t* src = c.k_store + (int64_t)u * k_unit_bytes(c.k_cap_blocks, KB, c.g) + ((int64_t)b0 * 4 + qt) * QB;
            for (int jb = 0; jb < nb; ++jb) bulk_g2s(dst + jb * p.kb_stride, src + (int64_t)jb * 4 * QB, (uint32_t)QB, bar, pol);
        } else if ((j -= 4 * wp.ntile) < wp.nkr) {
            const int t0 = (wp.kr0 + kCW * j) * kResTile, nt = min(kResTile, s.r - t0);
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.k_res + ((int64_t)u * c.R + t0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        } else if ((j -= wp.nkr) < wp.nvq) {
            const int vcb = v_tok_code_bytes(VB), vmb = v_tok_meta_bytes(c.g);
            const int t0 = (warp + kCW * j) * kVTile, nt = min(kVTile, s.tv - t0);
            const uint32_t cb = (uint32_t)(nt * vcb), mb = (uint32_t)((nt * vmb + 15) & ~15);
            mbar_expect_tx(bar, cb + mb);
            bulk_g2s(dst, c.v_codes + ((int64_t)u * c.v_cap + t0) * vcb, cb, bar, pol);
            bulk_g2s(dst + kVTile * vcb, c.v_meta + ((int64_t)u * c.v_cap + t0) * vmb, mb, bar, pol);
        } else {
            j -= wp.nvq;
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

// ------------------------------------------------------------------------------------------------
// compute pieces
// ------------------------------------------------------------------------------------------------
template <int BITS> struct Cell;
template <> struct Cell<2> { using vec_t = uint2; static constexpr int kWords = 2; };
template <> struct Cell<4> { using vec_t = uint4; static constexpr int kWords = 4; };

template <int BITS>
__device__ __forceinline__ void fma_cell32(float (&acc)[32], const typename Cell<BITS>::vec_t& cw, float a2) {
    constexpr int FPI = 32 / BITS;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&cw);
    #pragma unroll
    for (int j = 0; j < Cell<BITS>::kWords; ++j) {
        float (&sub)[FPI] = *reinterpret_cast<float (*)[FPI]>(&acc[j * FPI]);
        fma_word<BITS>(sub, w[j], a2);
    }
}
template <int BITS>
__device__ __forceinline__ float rescale32(int e) { return field_rescale<BITS>(e % (32 / BITS)); }

// logits (fp16 kernel output) -> fp16 scaled, the value that enters the softmax
__device__ __forceinline__ __half scale_logit(float acc) {
    return __float2half_rn(__half2float(__float2half_rn(acc)) * kRcpSqrtD);
}

// One quarter (32 channels) of a 4-block K tile.  lane = rp*16 + j*4 + tg; local rows rp + 2i, i < 16.
// padded stride of one K block-quarter in a stage: (stride mod 128) == bytes of one 4-cell row, so that
// the 4 blocks read by a half-warp land on disjoint banks
template <int KB, int GS>
struct KStage {
    static constexpr int QB = kQRows * 4 * 4 * KB + kQRows * (kBlockTokens / GS) * 4;
    static constexpr int kRow = 4 * 4 * KB;
    static constexpr int kStride = QB + ((kRow - QB % 128) % 128 + 128) % 128;
};

template <int KB, int G, int GS>
__device__ __forceinline__ void kq_quarter(const uint8_t* st, int qt, const float* qsp,
                                           float (&acc)[G][32], float (&zs)[G], int lane)
{
    using vec_t = typename Cell<KB>::vec_t;
    constexpr int cbk = 4 * KB;
    constexpr int g = GS;
    constexpr int kb_stride = KStage<KB, GS>::kStride;
    const int rp = lane >> 4, j = (lane >> 2) & 3, tg = lane & 3;
    constexpr int mpb = kBlockTokens / g;
    const uint8_t* blk = st + j * kb_stride;
    const uint8_t* cp = blk + rp * (4 * cbk) + tg * cbk;
    const uint8_t* mp = blk + kQRows * 4 * cbk + rp * (mpb * 4) + ((tg * kCell) / g) * 4;
    const float* qp = qsp + (qt * 2 + rp) * 16;                     // [h][qt][rp][i]
    #pragma unroll 1
    for (int i4 = 0; i4 < 4; ++i4) {
        float4 qv[G];
        #pragma unroll
        for (int h = 0; h < G; ++h) qv[h] = *reinterpret_cast<const float4*>(qp + h * kD + i4 * 4);
        vec_t cw[4];
        float2 sz[4];
        #pragma unroll
        for (int ii = 0; ii < 4; ++ii) {                            // all loads of the 4 rows first
            const int row2 = (i4 * 4 + ii) * 2;                     // local row = rp + row2
            cw[ii] = *reinterpret_cast<const vec_t*>(cp + row2 * (4 * cbk));
            sz[ii] = __half22float2(*reinterpret_cast<const __half2*>(mp + row2 * (mpb * 4)));
        }
        #pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const float x2 = ii == 0 ? qv[h].x : ii == 1 ? qv[h].y : ii == 2 ? qv[h].z : qv[h].w;
                zs[h] = fmaf(x2, sz[ii].y, zs[h]);
                fma_cell32<KB>(acc[h], cw[ii], x2 * sz[ii].x);
            }
        }
    }
}

// cache data movement of one unit (models/llama_kivi.py:343-356, :386-399); cold path, kept out of line
// executed by a team of `tsize` threads (multiple of 32); `tid` = index within the team
template <int KB, int VB>
__device__ __noinline__ void commit_unit(const DecodeParams& p, const Sched& s, int u, int tid, int tsize)
{
    const CacheDesc& c = p.c;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = c.g;
    // V: v_new joins the ring; if the window would exceed R, its oldest token is quantised
    if (tid < kD / 8)
        reinterpret_cast<uint4*>(c.v_res + ((int64_t)u * c.v_res_cap + (s.vhead + s.L) % c.v_res_cap) * kD)[tid] =
            __ldg(reinterpret_cast<const uint4*>(p.v_new + (int64_t)u * kD) + tid);
    if (s.L + 1 > c.R && warp == (tsize > 32 ? 1 : 0)) {
        constexpr int FPI = 32 / VB, WPT = kD / FPI;
        const float maxq = (float)((1 << VB) - 1);
        const __half* src = c.v_res + ((int64_t)u * c.v_res_cap + s.vhead) * kD;
        const bool act = lane < WPT;
        float x[FPI];
        #pragma unroll
        for (int e = 0; e < FPI; ++e) x[e] = act ? __half2float(src[lane * FPI + e]) : 0.f;
        float mnf = x[0], mxf = x[0];
        #pragma unroll
        for (int e = 1; e < FPI; ++e) { mnf = fminf(mnf, x[e]); mxf = fmaxf(mxf, x[e]); }
        const int lpg = g / FPI;
        for (int o = 1; o < lpg; o <<= 1) {
            mnf = fminf(mnf, __shfl_xor_sync(0xffffffffu, mnf, o));
            mxf = fmaxf(mxf, __shfl_xor_sync(0xffffffffu, mxf, o));
        }
        if (act) {
            const __half d16 = __float2half_rn(mxf - mnf);
            const __half sc = __float2half_rn(__fdiv_rn(__half2float(d16), maxq));
            const float scf = __half2float(sc);
            uint32_t word = 0;
            #pragma unroll
            for (int e = 0; e < FPI; ++e) {
                const __half t1 = __float2half_rn(x[e] - mnf);
                const __half t2 = __float2half_rn(__fdiv_rn(__half2float(t1), scf));
                const float f = fminf(fmaxf(__half2float(t2), 0.f), maxq);
                word |= (uint32_t)__float2int_rn(f) << (VB * e);
            }
            const int64_t tokidx = (int64_t)u * c.v_cap + s.tv;
            reinterpret_cast<uint32_t*>(c.v_codes)[tokidx * WPT + lane] = word;
            if (lane % lpg == 0)
                reinterpret_cast<__half2*>(c.v_meta)[tokidx * (kD / g) + lane / lpg] =
                    __halves2half2(sc, __float2half_rn(mnf));
        }
    }
    // K: k_new joins the window, or completes it -> quantise the R tokens per channel
    if (s.r + 1 < c.R) {
        if (tid >= 16 && tid < 16 + kD / 8)
            reinterpret_cast<uint4*>(c.k_res + ((int64_t)u * c.R + s.r) * kD)[tid - 16] =
                __ldg(reinterpret_cast<const uint4*>(p.k_new + (int64_t)u * kD) + (tid - 16));
    } else {
        constexpr int FPI = 32 / KB;
        constexpr int cbk = 4 * KB;
        const float maxq = (float)((1 << KB) - 1);
        uint8_t* ubase = c.k_store + (int64_t)u * k_unit_bytes(c.k_cap_blocks, KB, g);
        const __half* win = c.k_res + (int64_t)u * c.R * kD;
        const __half* knew = p.k_new + (int64_t)u * kD;
        for (int w = tid; w < kD * (c.R / g); w += tsize) {
            const int d = w % kD, grp = w / kD;
            auto tokval = [&](int t) -> float {
                return __half2float(t < c.R - 1 ? win[(int64_t)t * kD + d] : knew[d]);
            };
            float mnf = tokval(grp * g), mxf = mnf;
            for (int i = 1; i < g; ++i) { const float x = tokval(grp * g + i); mnf = fminf(mnf, x); mxf = fmaxf(mxf, x); }
            const __half d16 = __float2half_rn(mxf - mnf);
            const __half sc = __float2half_rn(__fdiv_rn(__half2float(d16), maxq));
            const float scf = __half2float(sc);
            for (int wi = 0; wi < g / FPI; ++wi) {
                uint32_t word = 0;
                #pragma unroll 1
                for (int e = 0; e < FPI; ++e) {
                    const float x = tokval(grp * g + wi * FPI + e);
                    const __half t1 = __float2half_rn(x - mnf);
                    const __half t2 = __float2half_rn(__fdiv_rn(__half2float(t1), scf));
                    const float f = fminf(fmaxf(__half2float(t2), 0.f), maxq);
                    word |= (uint32_t)__float2int_rn(f) << (KB * e);
                }
                const int tok = s.tk + grp * g + wi * FPI;          // absolute token of the word's first element
                const int blk = tok / kBlockTokens, bt = tok % kBlockTokens;
                *reinterpret_cast<uint32_t*>(ubase + k_row_off(blk, d, KB, g) + (bt / kCell) * cbk + ((bt % kCell) / FPI) * 4) = word;
                if (wi == 0)
                    *reinterpret_cast<__half2*>(ubase + k_meta_off(blk, d, KB, g) + (bt / g) * 4) =
                        __halves2half2(sc, __float2half_rn(mnf));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int KB, int VB, int G, int GS>
__global__ void __launch_bounds__(kThreads, G == 1 ? 2 : 1)
decode_attention_kernel(const DecodeParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const CacheDesc& c = p.c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_stages = kCW * p.spw;

    // carve shared memory (every pointer is derived from `smem` so that LDS/STS are generated)
    uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * p.stage_bytes);
    uint8_t* ptr = smem + (((size_t)n_stages * (p.stage_bytes + 8) + 127) & ~(size_t)127);
    float* qsp = reinterpret_cast<float*>(ptr); ptr += G * kD * 4;   // [G][4][2][16] q*2^90: channel d = qt*32 + rp + 2i
    float* qlin = reinterpret_cast<float*>(ptr); ptr += G * kD * 4;  // [G][128] q*2^90 in channel order
    float* stats = reinterpret_cast<float*>(ptr); ptr += 16 * 4;     // softmax block-reduce scratch
    float* pnew = reinterpret_cast<float*>(ptr); ptr += 16 * 4;      // probability of the new token, per head
    __half* lg = reinterpret_cast<__half*>(ptr);                     // [G][t_cap] scaled logits, then probabilities
    float* red = reinterpret_cast<float*>(ptr);                      // aliases lg: [kCW][G][2][128]

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

    constexpr int g = GS;
    const int ratio = c.H / c.Hkv;
    int m = 0;                                                      // items consumed so far by this warp
    #pragma unroll 1
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int b = u / c.Hkv;
        const int uq0 = u * ratio + hc * G;                         // first query head row of this chunk

        // -- stage q (x 2^90): channel order and the (quarter, parity, i) permutation used by kq_quarter
        for (int i = tid; i < G * kD; i += kThreads) {
            const int h = i / kD, d = i % kD;
            const float v = __half2float(p.q[(int64_t)(uq0 + h) * kD + d]) * kPreScale;
            qlin[i] = v;
            const int qt = d / kQRows, lr = d % kQRows;
            qsp[h * kD + (qt * 2 + (lr & 1)) * 16 + (lr >> 1)] = v;
        }
        __syncthreads();

        // ================= K phase =================
        #pragma unroll 1
        for (int a = 0; a < wp.ntile; ++a) {
            const int tile = warp + kCW * a;
            float acc[G][32];
            float zs[G];
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                zs[h] = 0.f;
                #pragma unroll
                for (int e = 0; e < 32; ++e) acc[h][e] = 0.f;
            }
            #pragma unroll 1
            for (int qt = 0; qt < 4; ++qt) {
                pp.wait_full(m);
                kq_quarter<KB, G, GS>(pp.stage(m), qt, qsp, acc, zs, lane);
                __syncwarp();
                issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
                ++m;
            }
            // combine the two row parities, finalise 16 tokens per lane
            const int rp = lane >> 4, j = (lane >> 2) & 3, tg = lane & 3;
            const int tok0 = (tile * 4 + j) * kBlockTokens + tg * kCell + rp * 16;
            const bool valid = tok0 < s.tk;
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                zs[h] += __shfl_xor_sync(0xffffffffu, zs[h], 16);
                const float zt = zs[h] * kPreScaleInv;
                __align__(16) __half o[16];
                #pragma unroll
                for (int e = 0; e < 16; ++e) {
                    // lane rp keeps elements [16*rp, 16*rp+16): send the other half, receive ours
                    const float mine = rp ? acc[h][16 + e] : acc[h][e];
                    const float send = rp ? acc[h][e] : acc[h][16 + e];
                    const float tot = mine + __shfl_xor_sync(0xffffffffu, send, 16);
                    o[e] = scale_logit(fmaf(tot, rescale32<KB>(e), zt));   // rescale32(16+e) == rescale32(e)
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(lg + (size_t)h * p.t_cap + tok0);
                    dst[0] = *reinterpret_cast<const uint4*>(&o[0]);
                    dst[1] = *reinterpret_cast<const uint4*>(&o[8]);
                }
            }
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
                            lg[(size_t)h * p.t_cap + s.tk + t0 + t] = scale_logit(sum[h] * kPreScaleInv);
                    }
                }
                __syncwarp();
                issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
                ++m;
            }
            // the new token (k_new, not yet in the window): one warp, plain loads
            if ((s.n_ktiles + s.n_kr) % kCW == warp) {
                const uint2 kv = __ldg(reinterpret_cast<const uint2*>(p.k_new + (int64_t)u * kD) + lane);
                const __half2* kh = reinterpret_cast<const __half2*>(&kv);
                const float2 k01 = __half22float2(kh[0]), k23 = __half22float2(kh[1]);
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float4 qv = *reinterpret_cast<const float4*>(qlin + h * kD + lane * 4);
                    float sum = qv.x * k01.x;
                    sum = fmaf(qv.y, k01.y, sum); sum = fmaf(qv.z, k23.x, sum); sum = fmaf(qv.w, k23.y, sum);
                    sum = warp_sum(sum);
                    if (lane == 0) lg[(size_t)h * p.t_cap + s.T - 1] = scale_logit(sum * kPreScaleInv);
                }
            }
        }
        __syncthreads();

        // ================= softmax (fp32), one block-wide reduction per head =================
        #pragma unroll 1
        for (int h = 0; h < G; ++h) {
            __half* row = lg + (size_t)h * p.t_cap;
            float ml = -INFINITY;
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
            float sl = 0.f;
            for (int t = tid; t < s.T; t += kThreads) sl += __expf(__half2float(row[t]) - ml);
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
            for (int t = tid; t < s.T; t += kThreads) {
                const __half pr = __float2half_rn(__fdiv_rn(__expf(__half2float(row[t]) - M), S));   // :375
                row[t] = pr;
                if (p.dbg_probs) p.dbg_probs[(int64_t)(uq0 + h) * p.dbg_stride + t] = pr;
                if (t == s.T - 1) pnew[h] = __half2float(pr);
            }
            if (G > 1) __syncthreads();                             // stats are reused by the next head
        }
        __syncthreads();

        // ================= V phase =================
        float oq[G][32];                                            // packed part: lane = (token row tr, cell)
        float ozs[G];
        float orr[G][4];                                            // fp16 window part: lane = 4 channels
        #pragma unroll
        for (int h = 0; h < G; ++h) {
            ozs[h] = 0.f;
            #pragma unroll
            for (int e = 0; e < 32; ++e) oq[h][e] = 0.f;
            #pragma unroll
            for (int e = 0; e < 4; ++e) orr[h][e] = 0.f;
        }
        {
            using vec_t = typename Cell<VB>::vec_t;
            constexpr int cbv = 4 * VB, vcb = 4 * cbv;
            constexpr int vmb = (kD / g) * 4;
            const int tr = lane >> 2, cell = lane & 3;
            #pragma unroll 1
            for (int a = 0; a < wp.nvq; ++a) {
                const int i = warp + kCW * a;
                const int t0 = i * kVTile, nt = min(kVTile, s.tv - t0);
                pp.wait_full(m);
                const uint8_t* st = pp.stage(m);
                const uint8_t* cp = st + tr * vcb + cell * cbv;
                const uint8_t* mp = st + kVTile * vcb + tr * vmb + ((cell * kCell) / g) * 4;
                const __half* prow = lg + t0 + tr;
                auto step = [&]() {
                    const vec_t cw = *reinterpret_cast<const vec_t*>(cp);
                    const float2 sz = __half22float2(*reinterpret_cast<const __half2*>(mp));
                    #pragma unroll
                    for (int h = 0; h < G; ++h) {
                        const float x2 = __half2float(prow[(size_t)h * p.t_cap]) * kPreScale;
                        ozs[h] = fmaf(x2, sz.y, ozs[h]);
                        fma_cell32<VB>(oq[h], cw, x2 * sz.x);
                    }
                };
                const int full = nt >> 3;                           // steps in which all 8 token rows are valid
                #pragma unroll 2
                for (int si = 0; si < full; ++si) {
                    step();
                    cp += 8 * vcb; mp += 8 * vmb; prow += 8;
                }
                if (tr < (nt & 7)) step();                          // ragged tail of the last tile
                __syncwarp();
                issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
                ++m;
            }
            #pragma unroll 1
            for (int bq = 0; bq < wp.nvr; ++bq) {
                const int i = wp.vr0 + kCW * bq;
                int l0, nt;                                         // logical index of the item's first token
                if (i < s.vr1) { l0 = i * kResTile; nt = min(kResTile, s.seg1 - l0); }
                else { const int t0 = (i - s.vr1) * kResTile; l0 = s.seg1 + t0; nt = min(kResTile, s.L - s.seg1 - t0); }
                pp.wait_full(m);
                const uint8_t* st = pp.stage(m);
                #pragma unroll 2
                for (int t = 0; t < nt; ++t) {
                    const uint2 vv = *reinterpret_cast<const uint2*>(st + t * 256 + lane * 8);
                    const __half2* vh = reinterpret_cast<const __half2*>(&vv);
                    const float2 v01 = __half22float2(vh[0]), v23 = __half22float2(vh[1]);
                    #pragma unroll
                    for (int h = 0; h < G; ++h) {
                        const float pr = __half2float(lg[(size_t)h * p.t_cap + s.tv + l0 + t]);
                        orr[h][0] = fmaf(pr, v01.x, orr[h][0]); orr[h][1] = fmaf(pr, v01.y, orr[h][1]);
                        orr[h][2] = fmaf(pr, v23.x, orr[h][2]); orr[h][3] = fmaf(pr, v23.y, orr[h][3]);
                    }
                }
                __syncwarp();
                issue_next<KB, VB>(pp, p, s, wp, warp, lane, pol);
                ++m;
            }
        }
        // reduce the packed part over the 8 token rows held by different lanes (xor 4, 8, 16)
        #pragma unroll
        for (int h = 0; h < G; ++h) {
            #pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                ozs[h] += __shfl_xor_sync(0xffffffffu, ozs[h], o);
                #pragma unroll
                for (int e = 0; e < 32; ++e) oq[h][e] += __shfl_xor_sync(0xffffffffu, oq[h][e], o);
            }
        }
        __syncthreads();                                            // everyone is done reading the probabilities
        {
            const int cell = lane & 3;
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                float* rq = red + ((size_t)(warp * G + h) * 2 + 0) * kD;
                float* rr = red + ((size_t)(warp * G + h) * 2 + 1) * kD;
                if (lane < 4) {
                    const float zt = ozs[h] * kPreScaleInv;
                    #pragma unroll
                    for (int e = 0; e < 32; ++e) rq[cell * 32 + e] = fmaf(oq[h][e], rescale32<VB>(e), zt);
                }
                *reinterpret_cast<float4*>(rr + lane * 4) = make_float4(orr[h][0], orr[h][1], orr[h][2], orr[h][3]);
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
        if (hc == 0) commit_unit<KB, VB>(p, s, u, tid, kThreads);
        __syncthreads();                                            // qsp / lg / red are reused by the next unit
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0, g_max_smem = 0;

template <int KB, int VB, int G, int GS>
static int launch_decode(DecodeParams& p, int max_kv_len, cudaStream_t st)
{
    const CacheDesc& c = p.c;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    p.kb_stride = KStage<KB, GS>::kStride;
    int stage = 4 * p.kb_stride;
    stage = max(stage, kVTile * (v_tok_code_bytes(VB) + v_tok_meta_bytes(c.g)) + 16);
    stage = max(stage, kResBytes);
    p.stage_bytes = (stage + 127) / 128 * 128;
    p.t_cap = max(4096, (max_kv_len + 63) / 64 * 64);
    const int fixed = 512 /*barriers, alignment*/ + 2 * G * kD * 4 + 128 + G * p.t_cap * 2;
    // G == 1 kernels are compiled for 2 CTAs per SM (<= 128 registers)
    int ctas = (G == 1) ? 2 : 1;
    p.spw = min(4, (g_max_smem / ctas - 1024 - fixed) / (kCW * p.stage_bytes));
    if (ctas == 2 && p.spw < 2) {
        ctas = 1;
        p.spw = min(4, (g_max_smem - fixed) / (kCW * p.stage_bytes));
    }
    if (p.spw < 1) return KIVI_ERR_CAPACITY;
    const size_t smem = (size_t)kCW * p.spw * p.stage_bytes + fixed;
    auto kern = decode_attention_kernel<KB, VB, G, GS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = min(p.n_units, g_num_sms * ctas);
    kern<<<grid, kThreads, smem, st>>>(p);
    return post_launch();
}


template <int KB, int VB>
static int dispatch_decode(DecodeParams& p, int G, int max_kv_len, cudaStream_t st)
{
    #define KIVI_GS(GS_)                                                                  \
        if (p.c.g == GS_) {                                                               \
            if (G == 4) return launch_decode<KB, VB, 4, GS_>(p, max_kv_len, st);          \
            if (G == 2) return launch_decode<KB, VB, 2, GS_>(p, max_kv_len, st);          \
            return launch_decode<KB, VB, 1, GS_>(p, max_kv_len, st);                      \
        }
    KIVI_GS(32)
    KIVI_GS(64)
    KIVI_GS(128)
    #undef KIVI_GS
    return KIVI_ERR_GROUP;
}

}  // namespace kivi
