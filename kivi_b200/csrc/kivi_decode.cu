// kivi_decode.cu -- fused KIVI decode attention over the blocked cache (sm_100a).
//
// One launch per layer per step replaces the ~30 launches of the reference's decode branch
// (models/llama_kivi.py:314-399): q.Kq^T with in-register dequantisation, the fp16 K window, scale,
// mask, fp32 softmax, p.Vq, the fp16 V window, the fp16 add, and the per-unit cache data movement
// (window append, K flush, V token pack).  Rounding points of the reference are reproduced
// (fp16 logits -> fp16 scale -> fp32 softmax -> fp16 probs -> fp16 partial outputs -> fp16 add).
//
// Execution model
//   * persistent grid: one CTA per SM, CTA c handles units c, c+grid, ...; unit = (b, kv-head, chunk
//     of G query heads) -- packed bytes are read once per KV head for all G heads (GQA);
//   * warp 8 = producer: one elected lane streams the unit's packed tiles HBM -> shared memory with
//     1-D bulk copies (cp.async.bulk, the TMA engine; SASS UBLKCP), completion on per-stage "full"
//     mbarriers, back-pressure on per-stage "empty" mbarriers.  It runs ahead across the
//     K -> softmax -> V phases and across units, so HBM never idles behind a barrier;
//   * warps 0..7 = consumers, each with S PRIVATE stages (a warp walks the rounds of its own stages
//     in order, so the mbarrier parity can never alias a round it has not reached).  The producer
//     issues phase-major, then round-robin over the warps (no head-of-line blocking, and nothing it
//     blocks on ever depends on an item issued later).  ~6 KB work items:
//       KQ  quarter of a 512-token K tile (4 blocks x 32 channels): lane = (row parity, block, cell),
//           one 64/128-bit LDS of codes + one 32-bit LDS of (scale, zero) per row, then per element
//           1 LOP3 (denormal unpack, kivi_common.cuh) + 1 FFMA per query head, fp32 accumulate;
//       KR  <= 24 tokens of the fp16 K window;          VQ  128 tokens of packed V;
//       VR  <= 24 tokens of the fp16 V ring.
//   * G == 1 (MHA) kernels fit 2 CTAs per SM: while one CTA sits in a phase barrier or the softmax,
//     the other keeps the FMA/ALU pipes and the TMA queue busy.
#include "kivi_decode.cuh"

namespace kivi {

int make_desc(const kivi_cache_t* k, CacheDesc* d);

constexpr int kCW = 8;                 // consumer warps
constexpr int kCT = kCW * 32;          // consumer threads
constexpr int kThreads = kCT + 32;
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
    int tk, r, tv, L, vhead, T;
    int n_ktiles, n_kr, n_vq, vr1, vr2, n_vr, items_k, items_v;
};

__device__ __forceinline__ Sched make_sched(const CacheDesc& c) {
    Sched s;
    s.tk = c.state[ST_TK]; s.r = c.state[ST_R]; s.tv = c.state[ST_TV]; s.L = c.state[ST_L]; s.vhead = c.state[ST_VHEAD];
    s.T = s.tk + s.r + 1;
    s.n_ktiles = cdiv(cdiv(s.tk, kBlockTokens), 4);
    s.n_kr = cdiv(s.r, kResTile);
    s.n_vq = cdiv(s.tv, kVTile);
    const int seg1 = min(s.L, c.v_res_cap - s.vhead);
    s.vr1 = cdiv(seg1, kResTile);
    s.vr2 = cdiv(s.L - seg1, kResTile);
    s.n_vr = s.vr1 + s.vr2;
    s.items_k = 4 * s.n_ktiles + s.n_kr;
    s.items_v = s.n_vq + s.n_vr;
    return s;
}

// number of indices i in [0, n) with i % kCW == w, and the first such index shifted by `base`
__device__ __forceinline__ int count_rr(int n, int w) { return n > w ? (n - w - 1) / kCW + 1 : 0; }
__device__ __forceinline__ int first_rr(int base, int w) { return (w - base % kCW + kCW) % kCW; }

// stage (w, m): the m-th item ever consumed by warp w lives in that warp's stage m % spw
struct Ring {
    uint8_t* base; uint64_t* full; uint64_t* empty; int spw, stage_bytes;
    __device__ __forceinline__ int idx(int w, int m) const { return w * spw + m % spw; }
    __device__ __forceinline__ uint8_t* stage(int w, int m) const { return base + (size_t)idx(w, m) * stage_bytes; }
    __device__ __forceinline__ void wait_full(int w, int m) const { mbar_wait(&full[idx(w, m)], (uint32_t)((m / spw) & 1)); }
    __device__ __forceinline__ void release(int w, int m) const { mbar_arrive(&empty[idx(w, m)]); }
};

// ------------------------------------------------------------------------------------------------
// producer
// ------------------------------------------------------------------------------------------------
template <int KB, int VB>
__device__ void producer_loop(const DecodeParams& p, const Sched& s, const Ring& ring)
{
    const CacheDesc& c = p.c;
    const int QB = k_q_bytes(KB, c.g);
    const int vcb = v_tok_code_bytes(VB), vmb = v_tok_meta_bytes(c.g);
    const uint64_t pol = policy_evict_first();
    const int nblk = cdiv(s.tk, kBlockTokens);
    const int64_t kunit = k_unit_bytes(c.k_cap_blocks, KB, c.g);
    int mw[kCW];
    #pragma unroll
    for (int w = 0; w < kCW; ++w) mw[w] = 0;
    // acquire warp w's next stage: wait until its previous occupant was released, arm the barrier
    auto acquire = [&](int w, uint32_t bytes, uint64_t*& bar) -> uint8_t* {
        const int m = mw[w]++;
        const int st = ring.idx(w, m);
        if (m >= ring.spw) mbar_wait(&ring.empty[st], (uint32_t)(((m / ring.spw) - 1) & 1));
        bar = &ring.full[st];
        mbar_expect_tx(bar, bytes);
        return ring.base + (size_t)st * ring.stage_bytes;
    };
    const int seg1 = min(s.L, c.v_res_cap - s.vhead);
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int u = unit / p.hchunks;                              // (b, kv head)
        uint64_t* bar;
        // ---- K phase: round j, warp w -> the j-th K-phase item of warp w
        {
            int rounds = 0;
            #pragma unroll
            for (int w = 0; w < kCW; ++w)
                rounds = max(rounds, 4 * count_rr(s.n_ktiles, w) + count_rr(s.n_kr - first_rr(s.n_ktiles, w) + w, w));
            for (int j = 0; j < rounds; ++j) {
                #pragma unroll
                for (int w = 0; w < kCW; ++w) {
                    const int nkq = 4 * count_rr(s.n_ktiles, w);
                    if (j < nkq) {
                        const int tile = w + kCW * (j >> 2), qt = j & 3;
                        const int b0 = tile * 4, nb = min(4, nblk - b0);
                        uint8_t* dst = acquire(w, (uint32_t)(nb * QB), bar);
                        for (int jb = 0; jb < nb; ++jb)
                            bulk_g2s(dst + jb * p.kb_stride,
                                     c.k_store + (int64_t)u * kunit + ((int64_t)(b0 + jb) * 4 + qt) * QB,
                                     (uint32_t)QB, bar, pol);
                    } else {
                        const int i = first_rr(s.n_ktiles, w) + kCW * (j - nkq);
                        if (i < s.n_kr) {
                            const int t0 = i * kResTile, nt = min(kResTile, s.r - t0);
                            uint8_t* dst = acquire(w, (uint32_t)(nt * kD * 2), bar);
                            bulk_g2s(dst, c.k_res + ((int64_t)u * c.R + t0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
                        }
                    }
                }
            }
        }
        // ---- V phase
        {
            int rounds = 0;
            #pragma unroll
            for (int w = 0; w < kCW; ++w)
                rounds = max(rounds, count_rr(s.n_vq, w) + count_rr(s.n_vr - first_rr(s.n_vq, w) + w, w));
            for (int j = 0; j < rounds; ++j) {
                #pragma unroll
                for (int w = 0; w < kCW; ++w) {
                    const int nq = count_rr(s.n_vq, w);
                    if (j < nq) {
                        const int i = w + kCW * j;
                        const int t0 = i * kVTile, nt = min(kVTile, s.tv - t0);
                        const uint32_t cb = (uint32_t)(nt * vcb), mb = (uint32_t)((nt * vmb + 15) & ~15);
                        uint8_t* dst = acquire(w, cb + mb, bar);
                        bulk_g2s(dst, c.v_codes + ((int64_t)u * c.v_cap + t0) * vcb, cb, bar, pol);
                        bulk_g2s(dst + kVTile * vcb, c.v_meta + ((int64_t)u * c.v_cap + t0) * vmb, mb, bar, pol);
                    } else {
                        const int i = first_rr(s.n_vq, w) + kCW * (j - nq);
                        if (i < s.n_vr) {
                            int slot0, nt;
                            if (i < s.vr1) { const int t0 = i * kResTile; slot0 = s.vhead + t0; nt = min(kResTile, seg1 - t0); }
                            else { const int t0 = (i - s.vr1) * kResTile; slot0 = t0; nt = min(kResTile, s.L - seg1 - t0); }
                            uint8_t* dst = acquire(w, (uint32_t)(nt * kD * 2), bar);
                            bulk_g2s(dst, c.v_res + ((int64_t)u * c.v_res_cap + slot0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// consumer pieces
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
template <int KB, int G>
__device__ __forceinline__ void kq_quarter(const uint8_t* st, int kb_stride, int g, int qt, const float* qsp,
                                           float (&acc)[G][32], float (&zs)[G], int lane)
{
    using vec_t = typename Cell<KB>::vec_t;
    constexpr int cbk = 4 * KB;
    const int rp = lane >> 4, j = (lane >> 2) & 3, tg = lane & 3;
    const int mpb = kBlockTokens / g;
    const uint8_t* blk = st + j * kb_stride;
    const uint8_t* cp = blk + rp * (4 * cbk) + tg * cbk;
    const uint8_t* mp = blk + kQRows * 4 * cbk + rp * (mpb * 4) + ((tg * kCell) / g) * 4;
    const float* qp = qsp + (qt * 2 + rp) * 16;                     // [h][qt][rp][i]
    #pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
        float4 qv[G];
        #pragma unroll
        for (int h = 0; h < G; ++h) qv[h] = *reinterpret_cast<const float4*>(qp + h * kD + i4 * 4);
        #pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int row2 = (i4 * 4 + ii) * 2;                     // local row = rp + row2
            const vec_t cw = *reinterpret_cast<const vec_t*>(cp + row2 * (4 * cbk));
            const float2 sz = __half22float2(*reinterpret_cast<const __half2*>(mp + row2 * (mpb * 4)));
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const float x2 = ii == 0 ? qv[h].x : ii == 1 ? qv[h].y : ii == 2 ? qv[h].z : qv[h].w;
                zs[h] = fmaf(x2, sz.y, zs[h]);
                fma_cell32<KB>(acc[h], cw, x2 * sz.x);
            }
        }
    }
}

template <int G>
struct SmemView {
    float* qsp;        // [G][4][2][16] q * 2^90, permuted (h, quarter, row parity, i): channel d = qt*32 + rp + 2i
    float* qlin;       // [G][128] q * 2^90 in channel order
    __half* lg;        // [G][t_cap] scaled logits, then probabilities
    float* red;        // aliases lg: [kCW][G][2][128]
    float* stats;      // [16] block-reduce scratch
    float* pnew;       // [G] probability of the new token
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int KB, int VB, int G>
__global__ void __launch_bounds__(kThreads, G == 1 ? 2 : 1)
decode_attention_kernel(const DecodeParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const CacheDesc& c = p.c;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n_stages = kCW * p.spw;
    Ring ring;
    ring.base = smem; ring.spw = p.spw; ring.stage_bytes = p.stage_bytes;
    uint8_t* ptr = smem + (size_t)n_stages * p.stage_bytes;     // stage_bytes % 128 == 0
    ring.full = reinterpret_cast<uint64_t*>(ptr); ptr += n_stages * 8;
    ring.empty = reinterpret_cast<uint64_t*>(ptr);
    ptr = smem + (((size_t)n_stages * (p.stage_bytes + 16) + 127) & ~(size_t)127);   // keeps the shared address space
    SmemView<G> sv;
    sv.qsp = reinterpret_cast<float*>(ptr); ptr += G * kD * 4;
    sv.qlin = reinterpret_cast<float*>(ptr); ptr += G * kD * 4;
    sv.stats = reinterpret_cast<float*>(ptr); ptr += 16 * 4;
    sv.pnew = reinterpret_cast<float*>(ptr); ptr += 16 * 4;
    sv.lg = reinterpret_cast<__half*>(ptr);
    sv.red = reinterpret_cast<float*>(ptr);

    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) { mbar_init(&ring.full[i], 1); mbar_init(&ring.empty[i], 1); }
        mbar_fence_init();
    }
    __syncthreads();

    const Sched s = make_sched(c);

    if (warp == kCW) {                                              // ===== producer warp
        if (lane == 0) producer_loop<KB, VB>(p, s, ring);
        return;
    }

    // ===== consumers
    const int tid = threadIdx.x;                                    // 0..255
    const int g = c.g;
    const int ratio = c.H / c.Hkv;
    int m = 0;                                                      // items consumed so far by this warp
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int b = u / c.Hkv;
        const int uq0 = u * ratio + hc * G;                         // first query head row of this chunk

        // -- stage q (x 2^90): channel order and the (half, parity, i) permutation used by kq_half
        for (int i = tid; i < G * kD; i += kCT) {
            const int h = i / kD, d = i % kD;
            const float v = __half2float(p.q[(int64_t)(uq0 + h) * kD + d]) * kPreScale;
            sv.qlin[i] = v;
            const int qt = d / kQRows, lr = d % kQRows;
            sv.qsp[h * kD + (qt * 2 + (lr & 1)) * 16 + (lr >> 1)] = v;
        }
        named_bar_sync(1, kCT);

        // ================= K phase =================
        for (int tile = warp; tile < s.n_ktiles; tile += kCW) {
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
                ring.wait_full(warp, m);
                kq_quarter<KB, G>(ring.stage(warp, m), p.kb_stride, g, qt, sv.qsp, acc, zs, lane);
                __syncwarp();
                if (lane == 0) ring.release(warp, m);
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
                    const float rs = rp ? rescale32<KB>(16 + e) : rescale32<KB>(e);
                    o[e] = scale_logit(fmaf(tot, rs, zt));
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(sv.lg + (size_t)h * p.t_cap + tok0);
                    dst[0] = *reinterpret_cast<const uint4*>(&o[0]);
                    dst[1] = *reinterpret_cast<const uint4*>(&o[8]);
                }
            }
        }
        // fp16 K window: items after the KQ items, round-robin continues
        {
            float qf[G][16];                                        // this lane's 16 channels of q
            const int part = lane & 7, tok = lane >> 3;
            bool loaded = false;
            for (int i = 0; i < s.n_kr; ++i) {
                if ((s.n_ktiles + i) % kCW != warp) continue;
                if (!loaded) {
                    #pragma unroll
                    for (int h = 0; h < G; ++h)
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            qf[h][e] = sv.qlin[h * kD + part * 8 + e];
                            qf[h][8 + e] = sv.qlin[h * kD + 64 + part * 8 + e];
                        }
                    loaded = true;
                }
                const int t0 = i * kResTile, nt = min(kResTile, s.r - t0);
                ring.wait_full(warp, m);
                const uint8_t* st = ring.stage(warp, m);
                for (int ts = 0; ts < nt; ts += 4) {
                    const int t = ts + tok;
                    float sum[G];
                    #pragma unroll
                    for (int h = 0; h < G; ++h) sum[h] = 0.f;
                    if (t < nt) {
                        const uint4 a = *reinterpret_cast<const uint4*>(st + t * 256 + part * 16);
                        const uint4 bq = *reinterpret_cast<const uint4*>(st + t * 256 + 128 + part * 16);
                        const __half2* ah = reinterpret_cast<const __half2*>(&a);
                        const __half2* bh = reinterpret_cast<const __half2*>(&bq);
                        #pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
                            #pragma unroll
                            for (int h = 0; h < G; ++h) {
                                sum[h] = fmaf(qf[h][2 * e], fa.x, sum[h]);
                                sum[h] = fmaf(qf[h][2 * e + 1], fa.y, sum[h]);
                                sum[h] = fmaf(qf[h][8 + 2 * e], fb.x, sum[h]);
                                sum[h] = fmaf(qf[h][8 + 2 * e + 1], fb.y, sum[h]);
                            }
                        }
                    }
                    #pragma unroll
                    for (int h = 0; h < G; ++h) {
                        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
                        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
                        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 4);
                        if (part == 0 && t < nt)
                            sv.lg[(size_t)h * p.t_cap + s.tk + t0 + t] = scale_logit(sum[h] * kPreScaleInv);
                    }
                }
                __syncwarp();
                if (lane == 0) ring.release(warp, m);
                ++m;
            }
            // the new token (k_new, not yet in the window): one warp, plain loads
            if ((s.n_ktiles + s.n_kr) % kCW == warp) {
                const uint2 kv = __ldg(reinterpret_cast<const uint2*>(p.k_new + (int64_t)u * kD) + lane);
                const __half2* kh = reinterpret_cast<const __half2*>(&kv);
                const float2 k01 = __half22float2(kh[0]), k23 = __half22float2(kh[1]);
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float4 qv = *reinterpret_cast<const float4*>(sv.qlin + h * kD + lane * 4);
                    float sum = qv.x * k01.x;
                    sum = fmaf(qv.y, k01.y, sum); sum = fmaf(qv.z, k23.x, sum); sum = fmaf(qv.w, k23.y, sum);
                    sum = warp_sum(sum);
                    if (lane == 0) sv.lg[(size_t)h * p.t_cap + s.T - 1] = scale_logit(sum * kPreScaleInv);
                }
            }
        }
        named_bar_sync(1, kCT);

        // ================= softmax (fp32) =================
        #pragma unroll 1
        for (int h = 0; h < G; ++h) {
            __half* row = sv.lg + (size_t)h * p.t_cap;
            float m = -INFINITY;
            for (int t = tid; t < s.T; t += kCT) {
                __half v = row[t];
                if (p.mask) {
                    v = __hadd_rn(v, p.mask[(int64_t)b * s.T + t]);                   // llama_kivi.py:369
                    if (__half2float(v) < -65504.f) v = __float2half_rn(-65504.f);    // :370-372 (max with finfo.min)
                    row[t] = v;
                }
                if (p.dbg_logits) p.dbg_logits[(int64_t)(uq0 + h) * p.dbg_stride + t] = v;
                m = fmaxf(m, __half2float(v));
            }
            #pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) sv.stats[warp] = m;
            named_bar_sync(1, kCT);
            m = sv.stats[0];
            #pragma unroll
            for (int w = 1; w < kCW; ++w) m = fmaxf(m, sv.stats[w]);
            float sum = 0.f;
            for (int t = tid; t < s.T; t += kCT) sum += expf(__half2float(row[t]) - m);
            sum = warp_sum(sum);
            if (lane == 0) sv.stats[8 + warp] = sum;
            named_bar_sync(1, kCT);
            sum = 0.f;
            #pragma unroll
            for (int w = 0; w < kCW; ++w) sum += sv.stats[8 + w];
            for (int t = tid; t < s.T; t += kCT) {
                const __half pr = __float2half_rn(__fdiv_rn(expf(__half2float(row[t]) - m), sum));   // :375
                row[t] = pr;
                if (p.dbg_probs) p.dbg_probs[(int64_t)(uq0 + h) * p.dbg_stride + t] = pr;
                if (t == s.T - 1) sv.pnew[h] = __half2float(pr);
            }
        }
        named_bar_sync(1, kCT);

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
            const int vmb = v_tok_meta_bytes(g);
            const int tr = lane >> 2, cell = lane & 3;
            for (int i = warp; i < s.n_vq; i += kCW) {
                const int t0 = i * kVTile, nt = min(kVTile, s.tv - t0);
                ring.wait_full(warp, m);
                const uint8_t* st = ring.stage(warp, m);
                const uint8_t* cp = st + tr * vcb + cell * cbv;
                const uint8_t* mp = st + kVTile * vcb + tr * vmb + ((cell * kCell) / g) * 4;
                #pragma unroll 4
                for (int ts = 0; ts < kVTile; ts += 8) {
                    if (ts + tr < nt) {
                        const vec_t cw = *reinterpret_cast<const vec_t*>(cp + ts * vcb);
                        const float2 sz = __half22float2(*reinterpret_cast<const __half2*>(mp + ts * vmb));
                        #pragma unroll
                        for (int h = 0; h < G; ++h) {
                            const float x2 = __half2float(sv.lg[(size_t)h * p.t_cap + t0 + ts + tr]) * kPreScale;
                            ozs[h] = fmaf(x2, sz.y, ozs[h]);
                            fma_cell32<VB>(oq[h], cw, x2 * sz.x);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ring.release(warp, m);
                ++m;
            }
            for (int i = 0; i < s.n_vr; ++i) {
                if ((s.n_vq + i) % kCW != warp) continue;
                const int seg1 = min(s.L, c.v_res_cap - s.vhead);
                int l0, nt;                                         // logical index of the item's first token
                if (i < s.vr1) { l0 = i * kResTile; nt = min(kResTile, seg1 - l0); }
                else { const int t0 = (i - s.vr1) * kResTile; l0 = seg1 + t0; nt = min(kResTile, s.L - seg1 - t0); }
                ring.wait_full(warp, m);
                const uint8_t* st = ring.stage(warp, m);
                #pragma unroll 4
                for (int t = 0; t < nt; ++t) {
                    const uint2 vv = *reinterpret_cast<const uint2*>(st + t * 256 + lane * 8);
                    const __half2* vh = reinterpret_cast<const __half2*>(&vv);
                    const float2 v01 = __half22float2(vh[0]), v23 = __half22float2(vh[1]);
                    #pragma unroll
                    for (int h = 0; h < G; ++h) {
                        const float pr = __half2float(sv.lg[(size_t)h * p.t_cap + s.tv + l0 + t]);
                        orr[h][0] = fmaf(pr, v01.x, orr[h][0]); orr[h][1] = fmaf(pr, v01.y, orr[h][1]);
                        orr[h][2] = fmaf(pr, v23.x, orr[h][2]); orr[h][3] = fmaf(pr, v23.y, orr[h][3]);
                    }
                }
                __syncwarp();
                if (lane == 0) ring.release(warp, m);
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
        named_bar_sync(1, kCT);                                     // everyone is done reading the probabilities
        {
            const int cell = lane & 3;
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                float* rq = sv.red + ((size_t)(warp * G + h) * 2 + 0) * kD;
                float* rr = sv.red + ((size_t)(warp * G + h) * 2 + 1) * kD;
                if (lane < 4) {
                    const float zt = ozs[h] * kPreScaleInv;
                    #pragma unroll
                    for (int e = 0; e < 32; ++e) rq[cell * 32 + e] = fmaf(oq[h][e], rescale32<VB>(e), zt);
                }
                *reinterpret_cast<float4*>(rr + lane * 4) = make_float4(orr[h][0], orr[h][1], orr[h][2], orr[h][3]);
            }
        }
        named_bar_sync(1, kCT);
        for (int i = tid; i < G * kD; i += kCT) {
            const int h = i / kD, d = i % kD;
            float q_sum = 0.f, r_sum = 0.f;
            #pragma unroll
            for (int w = 0; w < kCW; ++w) {
                q_sum += sv.red[((size_t)(w * G + h) * 2 + 0) * kD + d];
                r_sum += sv.red[((size_t)(w * G + h) * 2 + 1) * kD + d];
            }
            r_sum = fmaf(sv.pnew[h], __half2float(p.v_new[(int64_t)u * kD + d]), r_sum);
            __half o = __float2half_rn(r_sum);                                          // llama_kivi.py:380 / :384
            if (s.tv > 0) o = __hadd_rn(__float2half_rn(q_sum), o);                     // :382-384
            p.out[(int64_t)(uq0 + h) * kD + d] = o;
        }

        // ================= cache data movement for this unit (llama_kivi.py:343-356, :386-399) =========
        if (hc == 0) {
            // V: v_new joins the ring; if the window would exceed R, its oldest token is quantised
            if (tid < kD / 8)
                reinterpret_cast<uint4*>(c.v_res + ((int64_t)u * c.v_res_cap + (s.vhead + s.L) % c.v_res_cap) * kD)[tid] =
                    __ldg(reinterpret_cast<const uint4*>(p.v_new + (int64_t)u * kD) + tid);
            if (s.L + 1 > c.R && warp == 1) {
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
                if (tid >= 32 && tid < 32 + kD / 8)
                    reinterpret_cast<uint4*>(c.k_res + ((int64_t)u * c.R + s.r) * kD)[tid - 32] =
                        __ldg(reinterpret_cast<const uint4*>(p.k_new + (int64_t)u * kD) + (tid - 32));
            } else {
                constexpr int FPI = 32 / KB;
                constexpr int cbk = 4 * KB;
                const float maxq = (float)((1 << KB) - 1);
                uint8_t* ubase = c.k_store + (int64_t)u * k_unit_bytes(c.k_cap_blocks, KB, g);
                const __half* win = c.k_res + (int64_t)u * c.R * kD;
                const __half* knew = p.k_new + (int64_t)u * kD;
                for (int w = tid; w < kD * (c.R / g); w += kCT) {
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
                        #pragma unroll
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
        named_bar_sync(1, kCT);                                     // qs / lg / red are reused by the next unit
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0, g_max_smem = 0;

template <int KB, int VB, int G>
static int launch_decode(DecodeParams& p, int max_kv_len, cudaStream_t st)
{
    const CacheDesc& c = p.c;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    // padded stride of one K block-quarter in a stage: (stride mod 128) == bytes of one 4-cell row, so
    // that the 4 blocks read by a half-warp land on disjoint banks
    const int QB = k_q_bytes(KB, c.g);
    const int row_bytes = 4 * 4 * KB;
    p.kb_stride = QB + ((row_bytes - QB % 128) % 128 + 128) % 128;
    int stage = 4 * p.kb_stride;
    stage = max(stage, kVTile * (v_tok_code_bytes(VB) + v_tok_meta_bytes(c.g)) + 16);
    stage = max(stage, kResBytes);
    p.stage_bytes = (stage + 127) / 128 * 128;
    p.t_cap = max(4096, (max_kv_len + 63) / 64 * 64);
    const int fixed = 512 /*barriers, alignment*/ + 2 * G * kD * 4 + 128 + G * p.t_cap * 2;
    // G == 1 kernels are compiled for 2 CTAs per SM (<= 112 registers)
    int ctas = (G == 1) ? 2 : 1;
    p.spw = min(4, (g_max_smem / ctas - 1024 - fixed) / (kCW * p.stage_bytes));
    if (ctas == 2 && p.spw < 2) {
        ctas = 1;
        p.spw = min(4, (g_max_smem - fixed) / (kCW * p.stage_bytes));
    }
    if (ctas == 2) p.spw = min(p.spw, 2);
    if (p.spw < 1) return KIVI_ERR_CAPACITY;
    const size_t smem = (size_t)kCW * p.spw * p.stage_bytes + fixed;
    auto kern = decode_attention_kernel<KB, VB, G>;
    static bool attr_set = false;
    static size_t attr_smem = 0;
    if (!attr_set || smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true; attr_smem = g_max_smem;
    }
    const int grid = min(p.n_units, g_num_sms * ctas);
    kern<<<grid, kThreads, smem, st>>>(p);
    return post_launch();
}

}  // namespace kivi

using namespace kivi;

extern "C" int kivi_decode_attention_f16(const kivi_cache_t* cache, const void* q, const void* k_new, const void* v_new,
                                         const void* mask, void* out, void* dbg_logits, void* dbg_probs,
                                         int64_t dbg_stride, int max_kv_len, void* stream)
{
    DecodeParams p;
    int rc = make_desc(cache, &p.c);
    if (rc) return rc;
    if (!q || !k_new || !v_new || !out) return KIVI_ERR_NULL;
    if (max_kv_len <= 0) return KIVI_ERR_SHAPE;
    if (max_kv_len > p.c.k_cap_blocks * kBlockTokens) return KIVI_ERR_CAPACITY;
    p.q = (const __half*)q; p.k_new = (const __half*)k_new; p.v_new = (const __half*)v_new; p.mask = (const __half*)mask;
    p.out = (__half*)out; p.dbg_logits = (__half*)dbg_logits; p.dbg_probs = (__half*)dbg_probs; p.dbg_stride = dbg_stride;
    const int ratio = p.c.H / p.c.Hkv;
    const int G = ratio % 4 == 0 ? 4 : (ratio % 2 == 0 ? 2 : 1);
    p.hchunks = ratio / G;
    p.n_units = p.c.B * p.c.Hkv * p.hchunks;
    cudaStream_t st = (cudaStream_t)stream;
    #define KIVI_DISPATCH(KB_, VB_)                                               \
        if (p.c.k_bits == KB_ && p.c.v_bits == VB_) {                             \
            if (G == 4) return launch_decode<KB_, VB_, 4>(p, max_kv_len, st);     \
            if (G == 2) return launch_decode<KB_, VB_, 2>(p, max_kv_len, st);     \
            return launch_decode<KB_, VB_, 1>(p, max_kv_len, st);                 \
        }
    KIVI_DISPATCH(2, 2)
    KIVI_DISPATCH(4, 4)
    KIVI_DISPATCH(2, 4)
    KIVI_DISPATCH(4, 2)
    #undef KIVI_DISPATCH
    return KIVI_ERR_BITS;
}
