// kivi_attn.cuh -- KIVI decode attention over the blocked cache (sm_100a): three barrier-free kernels.
//
// One call per layer per step replaces the ~30 launches of the reference's decode branch
// (models/llama_kivi.py:314-399): q.Kq^T with in-register dequantisation, the fp16 K window, scale, mask,
// fp32 softmax, p.Vq, the fp16 V window, the fp16 add, and the per-unit cache data movement (window
// append, K flush, V token pack).  Rounding points of the reference are reproduced (fp16 logits -> fp16
// scale -> fp32 softmax -> fp16 probs -> fp16 partial outputs -> fp16 add).
//
//   qk_kernel        every WARP is an independent worker.  Work items (dealt round-robin over all warps of
//                    the persistent grid, so there is no barrier and no tail): one 128-token K block, <= 24
//                    tokens of the fp16 K window, or the new token.  Scaled fp16 logits go to a global
//                    workspace row.
//   softmax_kernel   one CTA per (b, head) row: mask, fp32 softmax, fp16 probabilities in place.
//   sv_kernel        a TEAM of 1..8 warps per unit streams the 128-token V blocks (with their probability
//                    slices) and the fp16 V window, combines behind a team-sized named barrier, writes
//                    the output and performs the unit's cache update.
//
// Data movement: each warp owns S private shared-memory stages and streams its own items HBM -> smem with
// 1-D bulk copies (cp.async.bulk = the TMA engine, SASS UBLKCP, L2 evict-first) completing on the stage's
// mbarrier; right after consuming a stage its elected lane issues the copy of the item S positions ahead
// (fence.proxy.async orders its reads before the async write).  A warp walks the rounds of its own stages
// in order, so the mbarrier parity can never alias a round it has not reached.
//
// Arithmetic of a packed block (128 inner x 128 outer, kivi_decode.cuh): sum_i x_i*(s_i,G * c_i,o + z_i,G)
//   = sum_i (x_i*s_i,G) * c_i,o + sum_i x_i*z_i,G.  The first term runs on the tensor cores as an UNPACK
//   AMORTISER: SIMT needs one LOP3 + one FFMA per code (2 issue slots per element; the 16-lane ALU pipe
//   is the bound -- measured 2.9 TB/s-equivalent, profiles/), HMMA needs only the 2-bit -> fp16 expansion,
//   one LOP3 + one HADD2 per PAIR of codes:
//     A (16 outer x 16 inner, fp16)  codes, exact: (0x6400 | field) - 1024 = code * (2^bits)^pos
//     B (16 inner x 8 cols,  fp16)  column (G, part): x_i*s_i,G split EXACTLY into hi = fp16(a) and
//                                   lo = fp16((a - hi) * 2^11)  (a = x*s has 22 significant bits), zero for
//                                   the rows / groups that do not exist
//     C (16 outer x 8 cols,  fp32)  row o, columns (G(o), hi/lo) are the wanted sums; other columns are
//                                   cross terms and are ignored.  Products are exact, accumulation is fp32.
//   (fp16 denormal codes would save the HADD2 but HMMA handles denormal inputs with ~2^-18 relative
//   error -- tools/probes/mma_denorm.cu -- so they are not used.)  |x*s| must stay below 65504.
//   With G query heads per KV head the unpacked A fragments are reused by G MMAs (GQA costs ~nothing).
#pragma once
#include "kivi_decode.cuh"

namespace kivi {

int make_desc(const kivi_cache_t* k, CacheDesc* d);

constexpr int kCW = 8;                 // warps per CTA
constexpr int kThreads = kCW * 32;
constexpr int kResTile = 24;           // tokens per fp16-window item (24 * 256 B = 6 KB)
constexpr int kResBytes = kResTile * kD * 2;
constexpr float kRcpSqrtD = 1.0f / 11.313708f;   // ATen: x * (1.0f / float(math.sqrt(128)))  (llama_kivi.py:339)
constexpr float kLoScale = 2048.f, kLoScaleInv = 1.f / 2048.f;

struct AttnParams {
    CacheDesc c;
    const __half* q; const __half* k_new; const __half* v_new; const __half* mask;
    __half* out; __half* dbg_logits; __half* dbg_probs;
    long long dbg_stride;
    __half* ws; long long ld;          // fp16 workspace [B*H][ld]: scaled logits, then probabilities
    int stage_bytes, spw, hchunks, n_units, team;
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

struct Pipe {                           // a warp's private stages
    uint8_t* base; uint64_t* full; int spw, stage_bytes, iss_n;
    __device__ __forceinline__ uint8_t* stage(int m) const { return base + (size_t)(m % spw) * stage_bytes; }
    __device__ __forceinline__ void wait_full(int m) const { mbar_wait(&full[m % spw], (uint32_t)((m / spw) & 1)); }
};

// logits (fp16 kernel output) -> fp16 scaled, the value that enters the softmax
__device__ __forceinline__ __half scale_logit(float acc) {
    return __float2half_rn(__half2float(__float2half_rn(acc)) * kRcpSqrtD);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// One chunk (16 inner indices) of a packed block on the tensor cores.
//   codes : this chunk's fragment words in shared memory (kChunkBytes)
//   meta4 : this lane's four (scale, zero) entries of the chunk for group `gcol` (one 128-bit load)
//   x[4]  : the lane's inner-vector values x_i for i = 2t, 2t+1, 2t+8, 2t+9 of the chunk (0 for invalid rows)
//   C[h][m] : accumulators of MMA m (outer rows 16m .. 16m+15), Z[h]: sum x*zero of this lane's group
// ------------------------------------------------------------------------------------------------
template <int BITS, int G>
__device__ __forceinline__ void mma_chunk(const uint8_t* codes, const uint4 meta4, const float (&x)[G][4],
                                          bool col_valid, int part, int lane, float (&C)[G][8][4], float (&Z)[G])
{
    using L = Lay<BITS>;
    // ---- B fragment: column (group, part) of x*s, split into hi / lo fp16
    const __half2* mh = reinterpret_cast<const __half2*>(&meta4);
    float sf[4], zf[4];
    #pragma unroll
    for (int r = 0; r < 4; ++r) { const float2 f = __half22float2(mh[r]); sf[r] = f.x; zf[r] = f.y; }
    uint32_t b0[G], b1[G];
    // branch-free column select: part 0 -> hi = fp16(a); part 1 -> lo = fp16((a - hi) * 2^11)
    const float k1 = part ? 1.f : 0.f, k2 = part ? kLoScale : 1.f;
    #pragma unroll
    for (int h = 0; h < G; ++h) {
        float a[4];
        #pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float xv = col_valid ? x[h][r] : 0.f;
            a[r] = xv * sf[r];
            Z[h] = fmaf(xv, zf[r], Z[h]);
        }
        const float2 f01 = __half22float2(__floats2half2_rn(a[0], a[1]));
        const float2 f23 = __half22float2(__floats2half2_rn(a[2], a[3]));
        b0[h] = pack_h2(fmaf(-f01.x, k1, a[0]) * k2, fmaf(-f01.y, k1, a[1]) * k2);
        b1[h] = pack_h2(fmaf(-f23.x, k1, a[2]) * k2, fmaf(-f23.y, k1, a[3]) * k2);
    }
    // ---- A fragments: unpack (LOP3 + HADD2 per pair of codes) and multiply
    uint32_t magic;                                              // fp16 1024.0 in both halves, kept in a register
    asm volatile("mov.b32 %0, 0x64006400;" : "=r"(magic));
    constexpr uint32_t kField = ((1u << BITS) - 1u) * 0x00010001u;
    const __half2 k1024 = __float2half2_rn(1024.f);
    #pragma unroll
    for (int sl = 0; sl < L::kSlabs; ++sl) {
        const uint4 w4 = *reinterpret_cast<const uint4*>(codes + sl * 512 + lane * 16);
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        uint32_t ws[4];
        #pragma unroll
        for (int r = 0; r < 4; ++r) ws[r] = w[r] >> L::kShift;
        #pragma unroll
        for (int j = 0; j < L::F; ++j) {
            uint32_t a[4];
            #pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t src = j < L::kInPlace ? w[r] : ws[r];
                uint32_t m;                                      // (src & mask) | magic in ONE LOP3 (magic in a register)
                asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(m) : "r"(src), "r"(kField << (BITS * L::pos(j))), "r"(magic));
                const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&m), k1024);
                a[r] = *reinterpret_cast<const uint32_t*>(&v);
            }
            #pragma unroll
            for (int h = 0; h < G; ++h) mma_16816(C[h][sl * L::F + j], a[0], a[1], a[2], a[3], b0[h], b1[h]);
        }
    }
}

// exact power of two (2^bits)^-pos(j) that undoes the in-place field position of MMA m
template <int BITS>
__device__ __forceinline__ float inv_pos_scale(int m) {
    return __uint_as_float((uint32_t)(127 - BITS * Lay<BITS>::pos(m % Lay<BITS>::F)) << 23);
}

// ------------------------------------------------------------------------------------------------
// cache data movement of one unit (models/llama_kivi.py:343-356, :386-399); cold path, out of line.
// Executed by a team of `tsize` threads (multiple of 32); `tid` = index within the team.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void set_field(uint8_t* blk, int bits, int inner, int outer, uint32_t code) {
    uint32_t* w = reinterpret_cast<uint32_t*>(blk + lay_word_off(bits, inner, outer));
    const int pos = lay_bit_pos(bits, inner, outer);
    atomicAnd(w, ~(((1u << bits) - 1u) << pos));
    atomicOr(w, code << pos);
}

template <int KB, int VB>
__device__ __noinline__ void commit_unit(const AttnParams& p, const Sched& s, int u, int tid, int tsize)
{
    const CacheDesc& c = p.c;
    const int g = c.g;
    // ---- V: v_new joins the ring; if the window would exceed R, its oldest token is quantised per token
    if (tid < kD / 8)
        reinterpret_cast<uint4*>(c.v_res + ((int64_t)u * c.v_res_cap + (s.vhead + s.L) % c.v_res_cap) * kD)[tid] =
            __ldg(reinterpret_cast<const uint4*>(p.v_new + (int64_t)u * kD) + tid);
    if (s.L + 1 > c.R) {
        const float maxq = (float)((1 << VB) - 1);
        const __half* src = c.v_res + ((int64_t)u * c.v_res_cap + s.vhead) * kD;
        uint8_t* blk = c.v_store + ((int64_t)u * c.v_cap_blocks + s.tv / kBlockTokens) * lay_block_bytes(VB, g);
        const int inner = s.tv % kBlockTokens;
        for (int ch = tid; ch < kD; ch += tsize) {                           // one thread per channel (group stats recomputed)
            const int G = ch / g;
            float mnf = __half2float(src[G * g]), mxf = mnf;
            for (int i = 1; i < g; ++i) { const float x = __half2float(src[G * g + i]); mnf = fminf(mnf, x); mxf = fmaxf(mxf, x); }
            const __half d16 = __float2half_rn(mxf - mnf);
            const __half sc = __float2half_rn(__fdiv_rn(__half2float(d16), maxq));
            const float scf = __half2float(sc);
            const __half t1 = __float2half_rn(__half2float(src[ch]) - mnf);
            const __half t2 = __float2half_rn(__fdiv_rn(__half2float(t1), scf));
            const float f = fminf(fmaxf(__half2float(t2), 0.f), maxq);
            set_field(blk, VB, inner, ch, (uint32_t)__float2int_rn(f));
            if (ch % g == 0)
                *reinterpret_cast<__half2*>(blk + lay_meta_off(VB, g, inner, G)) = __halves2half2(sc, __float2half_rn(mnf));
        }
    }
    // ---- K: k_new joins the window, or completes it -> quantise the R tokens per channel
    if (s.r + 1 < c.R) {
        if (tid >= 16 && tid < 16 + kD / 8)
            reinterpret_cast<uint4*>(c.k_res + ((int64_t)u * c.R + s.r) * kD)[tid - 16] =
                __ldg(reinterpret_cast<const uint4*>(p.k_new + (int64_t)u * kD) + (tid - 16));
    } else {
        const float maxq = (float)((1 << KB) - 1);
        const int bb = lay_block_bytes(KB, g);
        uint8_t* ub = c.k_store + (int64_t)u * c.k_cap_blocks * bb;
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
            const int tok0 = s.tk + grp * g;                                 // absolute token of the group's first element
            uint8_t* blk = ub + (int64_t)(tok0 / kBlockTokens) * bb;
            for (int i = 0; i < g; ++i) {
                const __half t1 = __float2half_rn(tokval(grp * g + i) - mnf);
                const __half t2 = __float2half_rn(__fdiv_rn(__half2float(t1), scf));
                const float f = fminf(fmaxf(__half2float(t2), 0.f), maxq);
                set_field(blk, KB, d, tok0 % kBlockTokens + i, (uint32_t)__float2int_rn(f));
            }
            *reinterpret_cast<__half2*>(blk + lay_meta_off(KB, g, d, (tok0 % kBlockTokens) / g)) =
                __halves2half2(sc, __float2half_rn(mnf));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// q . K^T
// ------------------------------------------------------------------------------------------------
template <int KB>
__device__ __forceinline__ void qk_issue_next(Pipe& pp, int& cur, const AttnParams& p, const Sched& s,
                                              int ipu, int total, int nw, int lane, uint64_t pol)
{
    const CacheDesc& c = p.c;
    while (cur < total && cur % ipu == ipu - 1) cur += nw;       // the new token needs no load
    if (cur >= total) return;
    const int unit = cur / ipu, j = cur % ipu;
    const int u = unit / p.hchunks;
    uint8_t* dst = pp.stage(pp.iss_n);
    uint64_t* bar = &pp.full[pp.iss_n % pp.spw];
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (j < s.n_kb) {
            const uint32_t bb = (uint32_t)lay_block_bytes(KB, c.g);
            mbar_expect_tx(bar, bb);
            bulk_g2s(dst, c.k_store + ((int64_t)u * c.k_cap_blocks + j) * bb, bb, bar, pol);
        } else {
            const int t0 = (j - s.n_kb) * kResTile, nt = min(kResTile, s.r - t0);
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.k_res + ((int64_t)u * c.R + t0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        }
    }
    ++pp.iss_n;
    cur += nw;
}

template <int KB, int G, int GS>
__global__ void __launch_bounds__(kThreads, G <= 2 ? 2 : 1)
qk_kernel(const AttnParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const CacheDesc& c = p.c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_stages = kCW * p.spw;
    uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * p.stage_bytes);
    uint8_t* ptr = smem + (((size_t)n_stages * (p.stage_bytes + 8) + 127) & ~(size_t)127);
    float* qf = reinterpret_cast<float*>(ptr) + warp * (G * kD);                // this warp's q, channel order

    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) mbar_init(&full_all[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const Sched s = make_sched(c);
    const uint64_t pol = policy_evict_first();
    const int ipu = s.n_kb + s.n_kr + 1;                                     // items per unit (last = new token)
    const int total = p.n_units * ipu;
    const int nw = gridDim.x * kCW, gw = blockIdx.x * kCW + warp;
    Pipe pp;
    pp.base = smem + (size_t)warp * p.spw * p.stage_bytes;
    pp.full = full_all + warp * p.spw;
    pp.spw = p.spw; pp.stage_bytes = p.stage_bytes; pp.iss_n = 0;
    int cur = gw;
    for (int i = 0; i < p.spw; ++i) qk_issue_next<KB>(pp, cur, p, s, ipu, total, nw, lane, pol);

    constexpr int NG = 128 / GS;                                             // token groups per block
    const int ratio = c.H / c.Hkv;
    const int g8 = lane >> 2, t = lane & 3;
    const int gcol = g8 >> 1, part = g8 & 1;
    const bool col_valid = gcol < NG;
    int m = 0, q_unit = -1;
    #pragma unroll 1
    for (int x = gw; x < total; x += nw) {
        const int unit = x / ipu, j = x % ipu;
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int uq0 = u * ratio + hc * G;
        if (unit != q_unit) {                                                // this warp's copy of q
            __syncwarp();
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const uint2 qv = __ldg(reinterpret_cast<const uint2*>(p.q + (int64_t)(uq0 + h) * kD) + lane);
                const __half2* qh = reinterpret_cast<const __half2*>(&qv);
                const float2 a = __half22float2(qh[0]), b2 = __half22float2(qh[1]);
                *reinterpret_cast<float4*>(qf + h * kD + lane * 4) = make_float4(a.x, a.y, b2.x, b2.y);
            }
            __syncwarp();
            q_unit = unit;
        }
        if (j < s.n_kb) {                                                    // ---- packed K block (tensor cores)
            float C[G][8][4];
            float Z[G];
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                Z[h] = 0.f;
                #pragma unroll
                for (int mm = 0; mm < 8; ++mm)
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) C[h][mm][e] = 0.f;
            }
            pp.wait_full(m);
            const uint8_t* st = pp.stage(m);
            const uint8_t* meta = st + Lay<KB>::kCodeBytes;
            #pragma unroll 2
            for (int ch = 0; ch < 8; ++ch) {
                float xq[G][4];
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float2 lo = *reinterpret_cast<const float2*>(qf + h * kD + ch * 16 + 2 * t);
                    const float2 hi = *reinterpret_cast<const float2*>(qf + h * kD + ch * 16 + 2 * t + 8);
                    xq[h][0] = lo.x; xq[h][1] = lo.y; xq[h][2] = hi.x; xq[h][3] = hi.y;
                }
                const uint4 m4 = *reinterpret_cast<const uint4*>(meta + (((ch * NG) + (col_valid ? gcol : 0)) * 4 + t) * 16);
                mma_chunk<KB, G>(st + ch * Lay<KB>::kChunkBytes, m4, xq, col_valid, part, lane, C, Z);
            }
            __syncwarp();
            qk_issue_next<KB>(pp, cur, p, s, ipu, total, nw, lane, pol);
            ++m;
            // Z of group gcol: sum over the four t-lanes (each covers 4 of the chunk's 16 channels)
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                Z[h] += __shfl_xor_sync(0xffffffffu, Z[h], 1);
                Z[h] += __shfl_xor_sync(0xffffffffu, Z[h], 2);
            }
            // lane (g8, t) holds, for MMA mm, rows g8 / g8+8 of columns (group t, hi | lo): useful when the
            // token's group equals t
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const float zt = __shfl_sync(0xffffffffu, Z[h], (2 * t) * 4);   // Z of group t lives in lanes g8 = 2t, 2t+1
                __half* row = p.ws + (int64_t)(uq0 + h) * p.ld + j * kBlockTokens;
                #pragma unroll
                for (int mm = 0; mm < 8; ++mm) {
                    if ((16 * mm) / GS == t) {
                        const float sc = inv_pos_scale<KB>(mm);
                        const int tok = 16 * mm + g8;
                        const float v0 = fmaf(fmaf(C[h][mm][1], kLoScaleInv, C[h][mm][0]), sc, zt);
                        const float v1 = fmaf(fmaf(C[h][mm][3], kLoScaleInv, C[h][mm][2]), sc, zt);
                        if (j * kBlockTokens + tok < s.tk) row[tok] = scale_logit(v0);
                        if (j * kBlockTokens + tok + 8 < s.tk) row[tok + 8] = scale_logit(v1);
                    }
                }
            }
        } else if (j < ipu - 1) {                                            // ---- fp16 K window item
            const int part8 = lane & 7, tok = lane >> 3;
            const int t0 = (j - s.n_kb) * kResTile, nt = min(kResTile, s.r - t0);
            pp.wait_full(m);
            const uint8_t* st = pp.stage(m);
            #pragma unroll 1
            for (int ts = 0; ts < nt; ts += 4) {
                const int tt = ts + tok;
                float sum[G];
                #pragma unroll
                for (int h = 0; h < G; ++h) sum[h] = 0.f;
                if (tt < nt) {
                    const uint4 a4 = *reinterpret_cast<const uint4*>(st + tt * 256 + part8 * 16);
                    const uint4 b4 = *reinterpret_cast<const uint4*>(st + tt * 256 + 128 + part8 * 16);
                    const __half2* ah = reinterpret_cast<const __half2*>(&a4);
                    const __half2* bh = reinterpret_cast<const __half2*>(&b4);
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
                        #pragma unroll
                        for (int h = 0; h < G; ++h) {
                            const float2 qa = *reinterpret_cast<const float2*>(qf + h * kD + part8 * 8 + 2 * e);
                            const float2 qb = *reinterpret_cast<const float2*>(qf + h * kD + 64 + part8 * 8 + 2 * e);
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
                    if (part8 == 0 && tt < nt)
                        p.ws[(int64_t)(uq0 + h) * p.ld + s.tk + t0 + tt] = scale_logit(sum[h]);
                }
            }
            __syncwarp();
            qk_issue_next<KB>(pp, cur, p, s, ipu, total, nw, lane, pol);
            ++m;
        } else {                                                             // ---- the new token
            const uint2 kv = __ldg(reinterpret_cast<const uint2*>(p.k_new + (int64_t)u * kD) + lane);
            const __half2* kh = reinterpret_cast<const __half2*>(&kv);
            const float2 k01 = __half22float2(kh[0]), k23 = __half22float2(kh[1]);
            #pragma unroll
            for (int h = 0; h < G; ++h) {
                const float4 qv = *reinterpret_cast<const float4*>(qf + h * kD + lane * 4);
                float sum = qv.x * k01.x;
                sum = fmaf(qv.y, k01.y, sum); sum = fmaf(qv.z, k23.x, sum); sum = fmaf(qv.w, k23.y, sum);
                sum = warp_sum(sum);
                if (lane == 0) p.ws[(int64_t)(uq0 + h) * p.ld + s.T - 1] = scale_logit(sum);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// softmax over the workspace rows (one CTA per (b, head))
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
softmax_kernel(__half* __restrict__ ws, long long ld, const int* __restrict__ state, const __half* __restrict__ mask,
               int H, __half* __restrict__ dbg_logits, __half* __restrict__ dbg_probs, long long dbg_stride)
{
    __shared__ float stats[16];
    const int T = state[ST_TK] + state[ST_R] + 1;
    const int rowi = blockIdx.x, b = rowi / H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __half* row = ws + (int64_t)rowi * ld;
    float ml = -INFINITY;
    for (int t = tid; t < T; t += 256) {
        __half v = row[t];
        if (mask) {
            v = __hadd_rn(v, mask[(int64_t)b * T + t]);                       // llama_kivi.py:369
            if (__half2float(v) < -65504.f) v = __float2half_rn(-65504.f);    // :370-372 (max with finfo.min)
            row[t] = v;
        }
        if (dbg_logits) dbg_logits[(int64_t)rowi * dbg_stride + t] = v;
        ml = fmaxf(ml, __half2float(v));
    }
    float sl = 0.f;
    for (int t = tid; t < T; t += 256) sl += __expf(__half2float(row[t]) - ml);
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
    for (int t = tid; t < T; t += 256) {
        const __half pr = __float2half_rn(__fdiv_rn(__expf(__half2float(row[t]) - M), S));   // :375
        row[t] = pr;
        if (dbg_probs) dbg_probs[(int64_t)rowi * dbg_stride + t] = pr;
    }
}

// ------------------------------------------------------------------------------------------------
// p . V
// ------------------------------------------------------------------------------------------------
struct SvPlan {                         // items of warp `wt` of a team of `team` warps, per unit
    int nvb, nvr, vr0, per_unit;
    __device__ __forceinline__ SvPlan(const Sched& s, int wt, int team) {
        nvb = s.n_vb > wt ? (s.n_vb - wt - 1) / team + 1 : 0;
        vr0 = (wt - s.n_vb % team + team) % team;
        nvr = s.n_vr > vr0 ? (s.n_vr - vr0 - 1) / team + 1 : 0;
        per_unit = nvb + nvr;
    }
};

struct SvCursor { int unit, j; };

template <int VB, int G>
__device__ __forceinline__ void sv_issue_next(Pipe& pp, SvCursor& cur, const AttnParams& p, const Sched& s,
                                              const SvPlan& pl, int wt, int nslots, int ratio, int lane, uint64_t pol)
{
    const CacheDesc& c = p.c;
    if (cur.unit >= p.n_units || pl.per_unit == 0) return;
    const int u = cur.unit / p.hchunks, hc = cur.unit % p.hchunks;
    uint8_t* dst = pp.stage(pp.iss_n);
    uint64_t* bar = &pp.full[pp.iss_n % pp.spw];
    int j = cur.j;
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (j < pl.nvb) {
            const uint32_t bb = (uint32_t)lay_block_bytes(VB, c.g);
            const int blk = wt + p.team * j, t0 = blk * kBlockTokens, nt = min(kBlockTokens, s.tv - t0);
            const uint32_t pb = (uint32_t)((nt * 2 + 15) & ~15);
            mbar_expect_tx(bar, bb + G * pb);
            bulk_g2s(dst, c.v_store + ((int64_t)u * c.v_cap_blocks + blk) * bb, bb, bar, pol);
            const int uq0 = u * ratio + hc * G;
            for (int h = 0; h < G; ++h)
                bulk_g2s(dst + bb + h * kBlockTokens * 2, p.ws + (int64_t)(uq0 + h) * p.ld + t0, pb, bar, pol);
        } else {
            j -= pl.nvb;
            const int i = pl.vr0 + p.team * j;
            int slot0, nt;
            if (i < s.vr1) { const int t0 = i * kResTile; slot0 = s.vhead + t0; nt = min(kResTile, s.seg1 - t0); }
            else { const int t0 = (i - s.vr1) * kResTile; slot0 = t0; nt = min(kResTile, s.L - s.seg1 - t0); }
            mbar_expect_tx(bar, (uint32_t)(nt * kD * 2));
            bulk_g2s(dst, c.v_res + ((int64_t)u * c.v_res_cap + slot0) * kD, (uint32_t)(nt * kD * 2), bar, pol);
        }
    }
    ++pp.iss_n;
    if (++cur.j == pl.per_unit) { cur.j = 0; cur.unit += nslots; }
}

template <int KB, int VB, int G, int GS>
__global__ void __launch_bounds__(kThreads, G <= 2 ? 2 : 1)
sv_kernel(const AttnParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
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
    const int team = p.team, tpc = kCW / team;
    const int ti = warp / team, wt = warp % team;                            // team index in the CTA, warp in the team
    const int nslots = gridDim.x * tpc, slot = blockIdx.x * tpc + ti;
    const int ttid = wt * 32 + lane, tsize = team * 32;
    const SvPlan pl(s, wt, team);
    const int ratio = c.H / c.Hkv;
    Pipe pp;
    pp.base = smem + (size_t)warp * p.spw * p.stage_bytes;
    pp.full = full_all + warp * p.spw;
    pp.spw = p.spw; pp.stage_bytes = p.stage_bytes; pp.iss_n = 0;
    SvCursor cur{slot, 0};
    for (int i = 0; i < p.spw; ++i) sv_issue_next<VB, G>(pp, cur, p, s, pl, wt, nslots, ratio, lane, pol);
    float* red = red_all + (size_t)ti * team * G * 2 * kD;                   // this team's [team][G][2][128]
    auto team_sync = [&]() {
        if (team == 1) __syncwarp();
        else named_bar_sync(1 + ti, tsize);
    };

    constexpr int NG = 128 / GS;                                             // channel groups
    const int g8 = lane >> 2, t = lane & 3;
    const int gcol = g8 >> 1, part = g8 & 1;
    const bool col_valid = gcol < NG;
    constexpr int kVBlock = Lay<VB>::kCodeBytes + 8 * NG * 16 * 4;           // block bytes for (VB, GS)
    int m = 0;
    #pragma unroll 1
    for (int unit = slot; unit < p.n_units; unit += nslots) {
        const int u = unit / p.hchunks, hc = unit % p.hchunks;
        const int uq0 = u * ratio + hc * G;
        float C[G][8][4];
        float Z[G];
        float orr[G][4];                                                     // fp16 window part: lane = 4 channels
        #pragma unroll
        for (int h = 0; h < G; ++h) {
            Z[h] = 0.f;
            #pragma unroll
            for (int mm = 0; mm < 8; ++mm)
                #pragma unroll
                for (int e = 0; e < 4; ++e) C[h][mm][e] = 0.f;
            #pragma unroll
            for (int e = 0; e < 4; ++e) orr[h][e] = 0.f;
        }
        #pragma unroll 1
        for (int a = 0; a < pl.nvb; ++a) {                                   // ---- packed V blocks (tensor cores)
            const int blk = wt + team * a, t0 = blk * kBlockTokens, nt = min(kBlockTokens, s.tv - t0);
            pp.wait_full(m);
            const uint8_t* st = pp.stage(m);
            const uint8_t* meta = st + Lay<VB>::kCodeBytes;
            const __half* prob = reinterpret_cast<const __half*>(st + kVBlock);
            #pragma unroll 2
            for (int ch = 0; ch < 8; ++ch) {
                float xp[G][4];
                const int i0 = ch * 16 + 2 * t;                              // tokens i0, i0+1, i0+8, i0+9 of the block
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(prob + h * kBlockTokens + i0));
                    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(prob + h * kBlockTokens + i0 + 8));
                    xp[h][0] = i0 < nt ? lo.x : 0.f;     xp[h][1] = i0 + 1 < nt ? lo.y : 0.f;
                    xp[h][2] = i0 + 8 < nt ? hi.x : 0.f; xp[h][3] = i0 + 9 < nt ? hi.y : 0.f;
                }
                uint4 m4 = *reinterpret_cast<const uint4*>(meta + (((ch * NG) + (col_valid ? gcol : 0)) * 4 + t) * 16);
                // rows beyond the packed length hold stale meta: neutralise (0 * NaN would poison the sums)
                if (i0 >= nt) { m4.x = 0u; } if (i0 + 1 >= nt) { m4.y = 0u; }
                if (i0 + 8 >= nt) { m4.z = 0u; } if (i0 + 9 >= nt) { m4.w = 0u; }
                mma_chunk<VB, G>(st + ch * Lay<VB>::kChunkBytes, m4, xp, col_valid, part, lane, C, Z);
            }
            __syncwarp();
            sv_issue_next<VB, G>(pp, cur, p, s, pl, wt, nslots, ratio, lane, pol);
            ++m;
        }
        #pragma unroll 1
        for (int bq = 0; bq < pl.nvr; ++bq) {                                // ---- fp16 V window items
            const int i = pl.vr0 + team * bq;
            int l0, nt;
            if (i < s.vr1) { l0 = i * kResTile; nt = min(kResTile, s.seg1 - l0); }
            else { const int t0 = (i - s.vr1) * kResTile; l0 = s.seg1 + t0; nt = min(kResTile, s.L - s.seg1 - t0); }
            pp.wait_full(m);
            const uint8_t* st = pp.stage(m);
            #pragma unroll 2
            for (int tt = 0; tt < nt; ++tt) {
                const uint2 vv = *reinterpret_cast<const uint2*>(st + tt * 256 + lane * 8);
                const __half2* vh = reinterpret_cast<const __half2*>(&vv);
                const float2 v01 = __half22float2(vh[0]), v23 = __half22float2(vh[1]);
                #pragma unroll
                for (int h = 0; h < G; ++h) {
                    const float pr = __half2float(__ldg(p.ws + (int64_t)(uq0 + h) * p.ld + s.tv + l0 + tt));
                    orr[h][0] = fmaf(pr, v01.x, orr[h][0]); orr[h][1] = fmaf(pr, v01.y, orr[h][1]);
                    orr[h][2] = fmaf(pr, v23.x, orr[h][2]); orr[h][3] = fmaf(pr, v23.y, orr[h][3]);
                }
            }
            __syncwarp();
            sv_issue_next<VB, G>(pp, cur, p, s, pl, wt, nslots, ratio, lane, pol);
            ++m;
        }
        // this warp's partial output: packed part from the accumulators (channel 16mm + g8 (+8), group t)
        #pragma unroll
        for (int h = 0; h < G; ++h) {
            Z[h] += __shfl_xor_sync(0xffffffffu, Z[h], 1);
            Z[h] += __shfl_xor_sync(0xffffffffu, Z[h], 2);
            const float zt = __shfl_sync(0xffffffffu, Z[h], (2 * t) * 4);
            float* rq = red + ((size_t)(wt * G + h) * 2 + 0) * kD;
            float* rr = red + ((size_t)(wt * G + h) * 2 + 1) * kD;
            #pragma unroll
            for (int mm = 0; mm < 8; ++mm) {
                if ((16 * mm) / GS == t) {
                    const float sc = inv_pos_scale<VB>(mm);
                    rq[16 * mm + g8] = fmaf(fmaf(C[h][mm][1], kLoScaleInv, C[h][mm][0]), sc, zt);
                    rq[16 * mm + g8 + 8] = fmaf(fmaf(C[h][mm][3], kLoScaleInv, C[h][mm][2]), sc, zt);
                }
            }
            *reinterpret_cast<float4*>(rr + lane * 4) = make_float4(orr[h][0], orr[h][1], orr[h][2], orr[h][3]);
        }
        team_sync();
        for (int i = ttid; i < G * kD; i += tsize) {
            const int h = i / kD, d = i % kD;
            float q_sum = 0.f, r_sum = 0.f;
            for (int w = 0; w < team; ++w) {
                q_sum += red[((size_t)(w * G + h) * 2 + 0) * kD + d];
                r_sum += red[((size_t)(w * G + h) * 2 + 1) * kD + d];
            }
            const float pn = __half2float(__ldg(p.ws + (int64_t)(uq0 + h) * p.ld + s.T - 1));
            r_sum = fmaf(pn, __half2float(p.v_new[(int64_t)u * kD + d]), r_sum);
            __half o = __float2half_rn(r_sum);                                          // llama_kivi.py:380 / :384
            if (s.tv > 0) o = __hadd_rn(__float2half_rn(q_sum), o);                     // :382-384
            p.out[(int64_t)(uq0 + h) * kD + d] = o;
        }
        if (hc == 0) commit_unit<KB, VB>(p, s, u, ttid, tsize);
        team_sync();                                                         // red is reused by the team's next unit
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0, g_max_smem = 0;

template <int KB, int VB, int G, int GS>
static int launch_attention(AttnParams& p, cudaStream_t st)
{
    const CacheDesc& c = p.c;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    int stage = max(lay_block_bytes(KB, c.g), lay_block_bytes(VB, c.g) + G * kBlockTokens * 2);
    stage = max(stage, kResBytes);
    p.stage_bytes = (stage + 127) / 128 * 128;
    const int fixed = 512 + kCW * G * 2 * kD * 4;                            // barriers + per-warp q / team reduce
    const int ctas = G <= 2 ? 2 : 1;
    p.spw = min(4, (g_max_smem / ctas - 1024 - fixed) / (kCW * p.stage_bytes));
    if (p.spw < 1) return KIVI_ERR_CAPACITY;
    const size_t smem = (size_t)kCW * p.spw * p.stage_bytes + fixed;
    auto kqk = qk_kernel<KB, G, GS>;
    auto ksv = sv_kernel<KB, VB, G, GS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kqk, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ksv, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int grid = g_num_sms * ctas;
    kqk<<<grid, kThreads, smem, st>>>(p);
    int rc = post_launch(); if (rc) return rc;
    softmax_kernel<<<c.B * c.H, 256, 0, st>>>(p.ws, p.ld, c.state, p.mask, c.H, p.dbg_logits, p.dbg_probs, p.dbg_stride);
    rc = post_launch(); if (rc) return rc;
    // team size: the smallest power of two that gives (almost) every warp slot of the grid a unit
    int team = 1;
    while (team < kCW && (long long)p.n_units * team < (long long)grid * kCW * 3 / 4) team *= 2;
    p.team = team;
    ksv<<<grid, kThreads, smem, st>>>(p);
    return post_launch();
}

template <int KB, int VB>
static int dispatch_attention(AttnParams& p, int G, cudaStream_t st)
{
    #define KIVI_GS(GS_)                                                                  \
        if (p.c.g == GS_) {                                                               \
            if (G == 4) return launch_attention<KB, VB, 4, GS_>(p, st);                   \
            if (G == 2) return launch_attention<KB, VB, 2, GS_>(p, st);                   \
            return launch_attention<KB, VB, 1, GS_>(p, st);                               \
        }
    KIVI_GS(32)
    KIVI_GS(64)
    KIVI_GS(128)
    #undef KIVI_GS
    return KIVI_ERR_GROUP;
}

}  // namespace kivi
