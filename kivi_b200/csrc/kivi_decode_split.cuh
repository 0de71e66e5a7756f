// kivi_decode_split.cuh -- barrier-free three-kernel form of the KIVI decode attention (sm_100a).
//
//   qk_split_kernel      every WARP is an independent worker: it streams whole K tiles (4 quarter items
//                        through its private TMA stages), finalises 512 logits and writes them, scaled,
//                        as fp16 to a global workspace row; fp16-window items and the new token likewise.
//                        Items are dealt round-robin over ALL warps of the grid: no barrier, no tail.
//   softmax_rows_kernel  one CTA per (b, head) row: mask, fp32 softmax, fp16 probabilities in place.
//   sv_split_kernel      a TEAM of 1..8 warps per unit streams the packed V tiles (+ their probability
//                        slices, fetched by the same bulk copy group) and the fp16 window, combines in
//                        shared memory behind a team-sized named barrier, writes the output and performs
//                        the unit's cache update.
//
// Same arithmetic and rounding points as the fused kernel (kivi_decode_impl.cuh); the intermediate fp16
// logits/probabilities make one round trip through L2/HBM (2 x B*H*T*2 bytes, +7 % traffic at cfg 2) in
// exchange for removing every phase barrier, and there is no shared-memory bound on the context length
// (the fused kernel keeps a [G][T] fp16 row per unit in shared memory).
#pragma once
#include "kivi_decode_impl.cuh"

namespace kivi {

struct SplitParams {
    DecodeParams d;
    __half* ws;              // [B*H][ld] fp16 workspace: scaled logits, then probabilities
    long long ld;
    int team;                // warps per unit in sv_split_kernel (1, 2, 4 or 8)
};

// ------------------------------------------------------------------------------------------------
// q . K^T
// ------------------------------------------------------------------------------------------------
struct QkCursor {                       // a warp's issue cursor over its items (stride = warps in the grid)
    int x, sub;
};

template <int KB>
__device__ __forceinline__ void qk_issue_next(Pipe& pp, QkCursor& cur, const SplitParams& sp, const Sched& s,
                                              int ipu, int total, int nw, int lane, uint64_t pol)
{
    const DecodeParams& p = sp.d;
    const CacheDesc& c = p.c;
    // skip items that need no load (the new token)
    while (cur.x < total && cur.x % ipu == ipu - 1) cur.x += nw;
    if (cur.x >= total) return;
    const int unit = cur.x / ipu, j = cur.x % ipu;
    const int u = unit / p.hchunks;
    uint8_t* dst = pp.stage(pp.iss_n);
    uint64_t* bar = &pp.full[pp.iss_n % pp.spw];
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (j < s.n_ktiles) {
            const int QB = k_q_bytes(KB, c.g);
            const int b0 = j * 4, nb = min(4, cdiv(s.tk, kBlockTokens) - b0);
            mbar_expect_tx(bar, (uint32_t)(nb * QB));
            const uint8_t* src = c.k_store + (int64_t)u * k_unit_bytes(c.k_cap_blocks, KB, c.g) + ((int64_t)b0 * 4 + cur.sub) * QB;
            for (int jb = 0; jb < nb; ++jb) bulk_g2s(dst + jb * p.kb_stride, src + (int64_t)jb * 4 * QB, (uint32_t)QB, bar, pol);
        } else {
            const int t0 = (j - s.n_ktiles) * kResTile, nt = min(kResTile, s.r - t0);
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.k_res + ((int64_t)u * c.R + t0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        }
    }
    ++pp.iss_n;
    if (j < s.n_ktiles) { if (++cur.sub == 4) { cur.sub = 0; cur.x += nw; } }
    else cur.x += nw;
}

template <int KB, int G, int GS>
__global__ void __launch_bounds__(kThreads, G == 1 ? 2 : 1)
qk_split_kernel(const SplitParams sp)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const DecodeParams& p = sp.d;
    const CacheDesc& c = p.c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_stages = kCW * p.spw;
    uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * p.stage_bytes);
    uint8_t* ptr = smem + (((size_t)n_stages * (p.stage_bytes + 8) + 127) & ~(size_t)127);
    float* qsp = reinterpret_cast<float*>(ptr) + warp * (2 * G * kD);        // per-warp q buffers
    float* qlin = qsp + G * kD;

    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) mbar_init(&full_all[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const Sched s = make_sched(c);
    const uint64_t pol = policy_evict_first();
    const int ipu = s.n_ktiles + s.n_kr + 1;                                 // items per unit (last = new token)
    const int total = p.n_units * ipu;
    const int nw = gridDim.x * kCW, gw = blockIdx.x * kCW + warp;
    Pipe pp;
    pp.base = smem + (size_t)warp * p.spw * p.stage_bytes;
    pp.full = full_all + warp * p.spw;
    pp.spw = p.spw; pp.stage_bytes = p.stage_bytes; pp.iss_n = 0;
    QkCursor cur{gw, 0};
    for (int i = 0; i < p.spw; ++i) qk_issue_next<KB>(pp, cur, sp, s, ipu, total, nw, lane, pol);

    const int ratio = c.H / c.Hkv;
    int m = 0;
    int q_unit = -1;
    #pragma unroll 1
    for (int x = gw; x < total; x += nw) {
        const int unit = x / ipu, j = x % ipu;
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int uq0 = u * ratio + hc * G;
        if (unit != q_unit) {                                                // this warp's copy of q (x 2^90)
            __syncwarp();
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const uint2 qv = __ldg(reinterpret_cast<const uint2*>(p.q + (int64_t)(uq0 + h) * kD) + lane);
                const __half2* qh = reinterpret_cast<const __half2*>(&qv);
                const float2 a = __half22float2(qh[0]), b2 = __half22float2(qh[1]);
                const float v4[4] = {a.x * kPreScale, a.y * kPreScale, b2.x * kPreScale, b2.y * kPreScale};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int d = lane * 4 + e;
                    qlin[h * kD + d] = v4[e];
                    const int qt = d / kQRows, lr = d % kQRows;
                    qsp[h * kD + (qt * 2 + (lr & 1)) * 16 + (lr >> 1)] = v4[e];
                }
            }
            __syncwarp();
            q_unit = unit;
        }
        if (j < s.n_ktiles) {                                                // ---- packed K tile
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
                qk_issue_next<KB>(pp, cur, sp, s, ipu, total, nw, lane, pol);
                ++m;
            }
            const int rp = lane >> 4, jb = (lane >> 2) & 3, tg = lane & 3;
            const int tok0 = (j * 4 + jb) * kBlockTokens + tg * kCell + rp * 16;
            const bool valid = tok0 < s.tk;
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                zs[h] += __shfl_xor_sync(0xffffffffu, zs[h], 16);
                const float zt = zs[h] * kPreScaleInv;
                __align__(16) __half o[16];
                #pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float mine = rp ? acc[h][16 + e] : acc[h][e];
                    const float send = rp ? acc[h][e] : acc[h][16 + e];
                    const float tot = mine + __shfl_xor_sync(0xffffffffu, send, 16);
                    o[e] = scale_logit(fmaf(tot, rescale32<KB>(e), zt));
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(sp.ws + (int64_t)(uq0 + h) * sp.ld + tok0);
                    dst[0] = *reinterpret_cast<const uint4*>(&o[0]);
                    dst[1] = *reinterpret_cast<const uint4*>(&o[8]);
                }
            }
        } else if (j < ipu - 1) {                                            // ---- fp16 K window item
            const int part = lane & 7, tok = lane >> 3;
            const int t0 = (j - s.n_ktiles) * kResTile, nt = min(kResTile, s.r - t0);
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
                        sp.ws[(int64_t)(uq0 + h) * sp.ld + s.tk + t0 + t] = scale_logit(sum[h] * kPreScaleInv);
                }
            }
            __syncwarp();
            qk_issue_next<KB>(pp, cur, sp, s, ipu, total, nw, lane, pol);
            ++m;
        } else {                                                             // ---- the new token
            const uint2 kv = __ldg(reinterpret_cast<const uint2*>(p.k_new + (int64_t)u * kD) + lane);
            const __half2* kh = reinterpret_cast<const __half2*>(&kv);
            const float2 k01 = __half22float2(kh[0]), k23 = __half22float2(kh[1]);
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const float4 qv = *reinterpret_cast<const float4*>(qlin + h * kD + lane * 4);
                float sum = qv.x * k01.x;
                sum = fmaf(qv.y, k01.y, sum); sum = fmaf(qv.z, k23.x, sum); sum = fmaf(qv.w, k23.y, sum);
                sum = warp_sum(sum);
                if (lane == 0) sp.ws[(int64_t)(uq0 + h) * sp.ld + s.T - 1] = scale_logit(sum * kPreScaleInv);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// softmax over the workspace rows (one CTA per (b, head))
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
softmax_rows_kernel(__half* __restrict__ ws, long long ld, const int* __restrict__ state, const __half* __restrict__ mask,
                    int H, __half* __restrict__ dbg_logits, __half* __restrict__ dbg_probs, long long dbg_stride)
{
    // the row is held in registers between the passes: 8 halfs (one 128-bit access) per thread per 2048 tokens
    constexpr int kMaxIter = 20;                                   // rows of up to 40960 tokens are held in registers
    __shared__ float stats[16];
    const int T = state[ST_TK] + state[ST_R] + 1;
    const int rowi = blockIdx.x, b = rowi / H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __half* row = ws + (int64_t)rowi * ld;                          // 16-B aligned (ld % 8 == 0)
    const int nvec = (T + 7) / 8;
    uint4 v[kMaxIter];
    float ml = -INFINITY;
    #pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const int i = tid + it * 256;
        if (i < nvec) {
            uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
            __half* h = reinterpret_cast<__half*>(&u);
            #pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int t = i * 8 + e;
                if (t < T) {
                    if (mask) {
                        h[e] = __hadd_rn(h[e], mask[(int64_t)b * T + t]);                   // llama_kivi.py:369
                        if (__half2float(h[e]) < -65504.f) h[e] = __float2half_rn(-65504.f);  // :370-372
                    }
                    if (dbg_logits) dbg_logits[(int64_t)rowi * dbg_stride + t] = h[e];
                    ml = fmaxf(ml, __half2float(h[e]));
                } else {
                    h[e] = __float2half_rn(-65504.f);               // padding of the last vector: exp -> 0
                }
            }
            v[it] = u;
        }
    }
    float sl = 0.f;
    #pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const int i = tid + it * 256;
        if (i < nvec) {
            const __half* h = reinterpret_cast<const __half*>(&v[it]);
            #pragma unroll
            for (int e = 0; e < 8; ++e) if (i * 8 + e < T) sl += __expf(__half2float(h[e]) - ml);
        }
    }
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
    for (int w = 1; w < 8; ++w) M = fmaxf(M, stats[w]);
    float S = 0.f;
    #pragma unroll
    for (int w = 0; w < 8; ++w) S += stats[w] == -INFINITY ? 0.f : stats[8 + w] * __expf(stats[w] - M);
    #pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const int i = tid + it * 256;
        if (i < nvec) {
            uint4 u = v[it];
            __half* h = reinterpret_cast<__half*>(&u);
            #pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int t = i * 8 + e;
                h[e] = __float2half_rn(__fdiv_rn(__expf(__half2float(h[e]) - M), S));       // :375
                if (dbg_probs && t < T) dbg_probs[(int64_t)rowi * dbg_stride + t] = h[e];
            }
            *reinterpret_cast<uint4*>(row + i * 8) = u;             // the <= 7 halfs past T stay inside the row (ld >= T + 8)
        }
    }
}

// ------------------------------------------------------------------------------------------------
// p . V
// ------------------------------------------------------------------------------------------------
struct SvPlan {                         // items of warp `wt` of a team of `team` warps, per unit
    int nvq, nvr, vr0, per_unit;
    __device__ __forceinline__ SvPlan(const Sched& s, int wt, int team) {
        nvq = s.n_vq > wt ? (s.n_vq - wt - 1) / team + 1 : 0;
        vr0 = (wt - s.n_vq % team + team) % team;
        nvr = s.n_vr > vr0 ? (s.n_vr - vr0 - 1) / team + 1 : 0;
        per_unit = nvq + nvr;
    }
};

template <int VB, int G>
__device__ __forceinline__ void sv_issue_next(Pipe& pp, const SplitParams& sp, const Sched& s, const SvPlan& pl,
                                              int wt, int nslots, int ratio, int lane, uint64_t pol)
{
    const DecodeParams& p = sp.d;
    const CacheDesc& c = p.c;
    if (pp.iss_unit >= p.n_units || pl.per_unit == 0) return;
    const int u = pp.iss_unit / p.hchunks, hc = pp.iss_unit % p.hchunks;
    uint8_t* dst = pp.stage(pp.iss_n);
    uint64_t* bar = &pp.full[pp.iss_n % pp.spw];
    int j = pp.iss_j;
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (j < pl.nvq) {
            const int vcb = v_tok_code_bytes(VB), vmb = v_tok_meta_bytes(c.g);
            const int t0 = (wt + sp.team * j) * kVTile, nt = min(kVTile, s.tv - t0);
            const uint32_t cb = (uint32_t)(nt * vcb), mb = (uint32_t)((nt * vmb + 15) & ~15), pb = (uint32_t)((nt * 2 + 15) & ~15);
            mbar_expect_tx(bar, cb + mb + G * pb);
            bulk_g2s(dst, c.v_codes + ((int64_t)u * c.v_cap + t0) * vcb, cb, bar, pol);
            bulk_g2s(dst + kVTile * vcb, c.v_meta + ((int64_t)u * c.v_cap + t0) * vmb, mb, bar, pol);
            const int uq0 = u * ratio + hc * G;
            for (int h = 0; h < G; ++h)
                bulk_g2s(dst + kVTile * (vcb + vmb) + h * kVTile * 2, sp.ws + (int64_t)(uq0 + h) * sp.ld + t0, pb, bar, pol);
        } else {
            j -= pl.nvq;
            const int i = pl.vr0 + sp.team * j;
            int slot0, nt;
            if (i < s.vr1) { const int t0 = i * kResTile; slot0 = s.vhead + t0; nt = min(kResTile, s.seg1 - t0); }
            else { const int t0 = (i - s.vr1) * kResTile; slot0 = t0; nt = min(kResTile, s.L - s.seg1 - t0); }
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.v_res + ((int64_t)u * c.v_res_cap + slot0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        }
    }
    ++pp.iss_n;
    if (++pp.iss_j == pl.per_unit) { pp.iss_j = 0; pp.iss_unit += nslots; }
}

template <int KB, int VB, int G, int GS>
__global__ void __launch_bounds__(kThreads, G == 1 ? 2 : 1)
sv_split_kernel(const SplitParams sp)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const DecodeParams& p = sp.d;
    const CacheDesc& c = p.c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_stages = kCW * p.spw;
    uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * p.stage_bytes);
    uint8_t* ptr = smem + (((size_t)n_stages * (p.stage_bytes + 8) + 127) & ~(size_t)127);
    float* red_all = reinterpret_cast<float*>(ptr);                          // [kCW][G][2][128]

    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) mbar_init(&full_all[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const Sched s = make_sched(c);
    const uint64_t pol = policy_evict_first();
    const int team = sp.team, tpc = kCW / team;
    const int ti = warp / team, wt = warp % team;                            // team index in the CTA, warp in the team
    const int nslots = gridDim.x * tpc, slot = blockIdx.x * tpc + ti;
    const int ttid = wt * 32 + lane, tsize = team * 32;
    const SvPlan pl(s, wt, team);
    const int ratio = c.H / c.Hkv;
    Pipe pp;
    pp.base = smem + (size_t)warp * p.spw * p.stage_bytes;
    pp.full = full_all + warp * p.spw;
    pp.spw = p.spw; pp.stage_bytes = p.stage_bytes;
    pp.iss_unit = slot; pp.iss_j = 0; pp.iss_n = 0;
    for (int i = 0; i < p.spw; ++i) sv_issue_next<VB, G>(pp, sp, s, pl, wt, nslots, ratio, lane, pol);
    float* red = red_all + (size_t)ti * team * G * 2 * kD;                   // this team's [team][G][2][128]
    auto team_sync = [&]() {
        if (team == 1) __syncwarp();
        else named_bar_sync(1 + ti, tsize);
    };

    constexpr int g = GS;
    int m = 0;
    #pragma unroll 1
    for (int unit = slot; unit < p.n_units; unit += nslots) {
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int uq0 = u * ratio + hc * G;
        float oq[G][32];
        float ozs[G];
        float orr[G][4];
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
            for (int a = 0; a < pl.nvq; ++a) {
                const int t0 = (wt + team * a) * kVTile, nt = min(kVTile, s.tv - t0);
                pp.wait_full(m);
                const uint8_t* st = pp.stage(m);
                const uint8_t* cp = st + tr * vcb + cell * cbv;
                const uint8_t* mp = st + kVTile * vcb + tr * vmb + ((cell * kCell) / g) * 4;
                const __half* prow = reinterpret_cast<const __half*>(st + kVTile * (vcb + vmb)) + tr;
                auto step = [&]() {
                    const vec_t cw = *reinterpret_cast<const vec_t*>(cp);
                    const float2 sz = __half22float2(*reinterpret_cast<const __half2*>(mp));
                    #pragma unroll
                    for (int h = 0; h < G; ++h) {
                        const float x2 = __half2float(prow[h * kVTile]) * kPreScale;
                        ozs[h] = fmaf(x2, sz.y, ozs[h]);
                        fma_cell32<VB>(oq[h], cw, x2 * sz.x);
                    }
                };
                const int full = nt >> 3;
                #pragma unroll 2
                for (int si = 0; si < full; ++si) {
                    step();
                    cp += 8 * vcb; mp += 8 * vmb; prow += 8;
                }
                if (tr < (nt & 7)) step();
                __syncwarp();
                sv_issue_next<VB, G>(pp, sp, s, pl, wt, nslots, ratio, lane, pol);
                ++m;
            }
            #pragma unroll 1
            for (int bq = 0; bq < pl.nvr; ++bq) {
                const int i = pl.vr0 + team * bq;
                int l0, nt;
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
                        const float pr = __half2float(__ldg(sp.ws + (int64_t)(uq0 + h) * sp.ld + s.tv + l0 + t));
                        orr[h][0] = fmaf(pr, v01.x, orr[h][0]); orr[h][1] = fmaf(pr, v01.y, orr[h][1]);
                        orr[h][2] = fmaf(pr, v23.x, orr[h][2]); orr[h][3] = fmaf(pr, v23.y, orr[h][3]);
                    }
                }
                __syncwarp();
                sv_issue_next<VB, G>(pp, sp, s, pl, wt, nslots, ratio, lane, pol);
                ++m;
            }
        }
        #pragma unroll
        for (int h = 0; h < G; ++h) {
            #pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                ozs[h] += __shfl_xor_sync(0xffffffffu, ozs[h], o);
                #pragma unroll
                for (int e = 0; e < 32; ++e) oq[h][e] += __shfl_xor_sync(0xffffffffu, oq[h][e], o);
            }
        }
        {
            const int cell = lane & 3;
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                float* rq = red + ((size_t)(wt * G + h) * 2 + 0) * kD;
                float* rr = red + ((size_t)(wt * G + h) * 2 + 1) * kD;
                if (lane < 4) {
                    const float zt = ozs[h] * kPreScaleInv;
                    #pragma unroll
                    for (int e = 0; e < 32; ++e) rq[cell * 32 + e] = fmaf(oq[h][e], rescale32<VB>(e), zt);
                }
                *reinterpret_cast<float4*>(rr + lane * 4) = make_float4(orr[h][0], orr[h][1], orr[h][2], orr[h][3]);
            }
        }
        team_sync();
        for (int i = ttid; i < G * kD; i += tsize) {
            const int h = i / kD, d = i % kD;
            float q_sum = 0.f, r_sum = 0.f;
            for (int w = 0; w < team; ++w) {
                q_sum += red[((size_t)(w * G + h) * 2 + 0) * kD + d];
                r_sum += red[((size_t)(w * G + h) * 2 + 1) * kD + d];
            }
            const float pn = __half2float(__ldg(sp.ws + (int64_t)(uq0 + h) * sp.ld + s.T - 1));
            r_sum = fmaf(pn, __half2float(p.v_new[(int64_t)u * kD + d]), r_sum);
            __half o = __float2half_rn(r_sum);
            if (s.tv > 0) o = __hadd_rn(__float2half_rn(q_sum), o);
            p.out[(int64_t)(uq0 + h) * kD + d] = o;
        }
        if (hc == 0) commit_unit<KB, VB>(p, s, u, ttid, tsize);
        team_sync();                                                         // red is reused by the team's next unit
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
template <int KB, int VB, int G, int GS>
static int launch_decode_split(SplitParams& sp, cudaStream_t st)
{
    DecodeParams& p = sp.d;
    const CacheDesc& c = p.c;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    p.kb_stride = KStage<KB, GS>::kStride;
    int stage = 4 * p.kb_stride;
    stage = max(stage, kVTile * (v_tok_code_bytes(VB) + v_tok_meta_bytes(c.g)) + G * kVTile * 2 + 16);
    stage = max(stage, kResBytes);
    p.stage_bytes = (stage + 127) / 128 * 128;
    const int fixed = 512 + kCW * G * 2 * kD * 4;                            // barriers + per-warp q buffers / team reduce
    int ctas = (G == 1) ? 2 : 1;
    p.spw = min(4, (g_max_smem / ctas - 1024 - fixed) / (kCW * p.stage_bytes));
    if (ctas == 2 && p.spw < 2) { ctas = 1; p.spw = min(4, (g_max_smem - fixed) / (kCW * p.stage_bytes)); }
    if (p.spw < 1) return KIVI_ERR_CAPACITY;
    const size_t smem = (size_t)kCW * p.spw * p.stage_bytes + fixed;
    auto kqk = qk_split_kernel<KB, G, GS>;
    auto ksv = sv_split_kernel<KB, VB, G, GS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kqk, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ksv, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = g_num_sms * ctas;
    kqk<<<grid, kThreads, smem, st>>>(sp);
    int rc = post_launch(); if (rc) return rc;
    softmax_rows_kernel<<<c.B * c.H, 256, 0, st>>>(sp.ws, sp.ld, c.state, p.mask, c.H, p.dbg_logits, p.dbg_probs, p.dbg_stride);
    rc = post_launch(); if (rc) return rc;
    // team size: the smallest power of two that gives every warp slot of the grid a unit
    int team = 1;
    while (team < kCW && (long long)p.n_units * team < (long long)grid * kCW * 3 / 4) team *= 2;
    sp.team = team;
    ksv<<<grid, kThreads, smem, st>>>(sp);
    return post_launch();
}

template <int KB, int VB>
static int dispatch_decode_split(SplitParams& sp, int G, cudaStream_t st)
{
    #define KIVI_GS(GS_)                                                                  \
        if (sp.d.c.g == GS_) {                                                            \
            if (G == 4) return launch_decode_split<KB, VB, 4, GS_>(sp, st);               \
            if (G == 2) return launch_decode_split<KB, VB, 2, GS_>(sp, st);               \
            return launch_decode_split<KB, VB, 1, GS_>(sp, st);                           \
        }
    KIVI_GS(32)
    KIVI_GS(64)
    KIVI_GS(128)
    #undef KIVI_GS
    return KIVI_ERR_GROUP;
}

}  // namespace kivi
