// kivi_cache.cu -- pre-allocated, blocked KIVI cache: sizing, prefill pack, state advance, export to
// the reference's 9-tuple layout (sm_100a).  Layout: kivi_decode.cuh.
//
// Replaces the cache handling of LlamaFlashAttention_KIVI.forward (models/llama_kivi.py):
//   prefill split + pack  :425-452   -> kivi_cache_prefill_f16 (fused transpose+quantise of K, no
//                                       .transpose(2,3).contiguous() temp; bulk per-token pack of V)
//   torch.cat growth      :350-352, :393-395, :391 -> nothing: blocks / ring slots are written in place
//   9-tuple               :454-455   -> kivi_cache_export_f16 (tests / interop only)
#include "kivi_decode.cuh"

namespace kivi {

// same arithmetic as kivi_pack.cu (quant/new_pack.py:238-241), kept local to stay header-only free
__device__ __forceinline__ uint32_t q_one(float x, float mnf, float scf, float maxq) {
    const __half t1 = __float2half_rn(x - mnf);
    const __half t2 = __float2half_rn(__fdiv_rn(__half2float(t1), scf));
    float f = __half2float(t2);
    f = fminf(fmaxf(f, 0.f), maxq);
    return (uint32_t)__float2int_rn(f);
}
__device__ __forceinline__ __half scale_of(float mnf, float mxf, float maxq) {
    const __half d = __float2half_rn(mxf - mnf);
    return __float2half_rn(__fdiv_rn(__half2float(d), maxq));
}

// ------------------------------------------------------------------------------------------------
// K prefill: one CTA per (unit, 128-token block).  The [128 tokens][128 ch] fp16 tile is staged in
// shared memory (coalesced 256-B rows), then thread (d, group) quantises g tokens of channel d.
// ------------------------------------------------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(256)
k_prefill_kernel(CacheDesc c, const __half* __restrict__ k, int n, int nq)
{
    constexpr int FPI = 32 / BITS;
    __shared__ __half tile[kBlockTokens][kD + 8];                // +8 halfs: conflict-free column reads
    const int u = blockIdx.y, blk = blockIdx.x;
    const int t0 = blk * kBlockTokens;
    const int nt = min(kBlockTokens, nq - t0);                   // multiple of g
    const __half* src = k + ((int64_t)u * n + t0) * kD;
    for (int i = threadIdx.x; i < nt * (kD / 8); i += blockDim.x) {
        const int t = i / (kD / 8), p = i % (kD / 8);
        *reinterpret_cast<uint4*>(&tile[t][p * 8]) = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)t * kD) + p);
    }
    __syncthreads();
    const int g = c.g;
    const int cbk = k_cell_bytes(BITS);
    const float maxq = (float)((1 << BITS) - 1);
    uint8_t* ubase = c.k_store + (int64_t)u * k_unit_bytes(c.k_cap_blocks, BITS, g);
    for (int w = threadIdx.x; w < kD * (nt / g); w += blockDim.x) {
        const int d = w % kD, grp = w / kD;
        float mnf = __half2float(tile[grp * g][d]), mxf = mnf;
        for (int i = 1; i < g; ++i) {
            const float x = __half2float(tile[grp * g + i][d]);
            mnf = fminf(mnf, x); mxf = fmaxf(mxf, x);
        }
        const __half sc = scale_of(mnf, mxf, maxq);
        const float scf = __half2float(sc);
        uint8_t* rowp = ubase + k_row_off(blk, d, BITS, g);
        for (int wi = 0; wi < g / FPI; ++wi) {
            uint32_t word = 0;
            #pragma unroll
            for (int j = 0; j < FPI; ++j)
                word |= q_one(__half2float(tile[grp * g + wi * FPI + j][d]), mnf, scf, maxq) << (BITS * j);
            const int tok = grp * g + wi * FPI;                  // token in block of the word's first element
            *reinterpret_cast<uint32_t*>(rowp + (tok / kCell) * cbk + ((tok % kCell) / FPI) * 4) = word;
        }
        *reinterpret_cast<__half2*>(ubase + k_meta_off(blk, d, BITS, g) + grp * 4) = __halves2half2(sc, __float2half_rn(mnf));
    }
}

// V prefill: thread = one word of one token (like kivi_pack.cu), writing the interleaved meta.
template <int BITS>
__global__ void __launch_bounds__(256)
v_prefill_kernel(CacheDesc c, const __half* __restrict__ v, int n, int nqv, int lpg_log2)
{
    constexpr int FPI = 32 / BITS;
    constexpr int WPT = kD / FPI;                                // words per token
    const int64_t wid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)c.B * c.Hkv * nqv * WPT;
    const bool active = wid < total;
    const float maxq = (float)((1 << BITS) - 1);
    int64_t tokflat = active ? wid / WPT : 0;
    const int wi = (int)(wid % WPT);
    const int u = (int)(tokflat / nqv), t = (int)(tokflat % nqv);
    float x[FPI];
    const __half* src = v + ((int64_t)u * n + t) * kD + wi * FPI;
    #pragma unroll
    for (int j = 0; j < FPI; ++j) x[j] = active ? __half2float(__ldg(src + j)) : 0.f;
    float mnf = x[0], mxf = x[0];
    #pragma unroll
    for (int j = 1; j < FPI; ++j) { mnf = fminf(mnf, x[j]); mxf = fmaxf(mxf, x[j]); }
    for (int o = 1; o < (1 << lpg_log2); o <<= 1) {
        mnf = fminf(mnf, __shfl_xor_sync(0xffffffffu, mnf, o));
        mxf = fmaxf(mxf, __shfl_xor_sync(0xffffffffu, mxf, o));
    }
    if (!active) return;
    const __half sc = scale_of(mnf, mxf, maxq);
    const float scf = __half2float(sc);
    uint32_t word = 0;
    #pragma unroll
    for (int j = 0; j < FPI; ++j) word |= q_one(x[j], mnf, scf, maxq) << (BITS * j);
    const int64_t tokidx = (int64_t)u * c.v_cap + t;
    reinterpret_cast<uint32_t*>(c.v_codes)[tokidx * WPT + wi] = word;
    if ((wi & ((1 << lpg_log2) - 1)) == 0)
        reinterpret_cast<__half2*>(c.v_meta)[tokidx * (kD / c.g) + (wi >> lpg_log2)] =
            __halves2half2(sc, __float2half_rn(mnf));
}

// residual windows of the prompt + state
__global__ void __launch_bounds__(256)
residual_prefill_kernel(CacheDesc c, const __half* __restrict__ k, const __half* __restrict__ v, int n,
                        int nqk, int nqv)
{
    const int u = blockIdx.x;
    const int r = n - nqk, L = n - nqv;
    for (int i = threadIdx.x; i < r * (kD / 8); i += blockDim.x)
        reinterpret_cast<uint4*>(c.k_res + (int64_t)u * c.R * kD)[i] =
            __ldg(reinterpret_cast<const uint4*>(k + ((int64_t)u * n + nqk) * kD) + i);
    for (int i = threadIdx.x; i < L * (kD / 8); i += blockDim.x)
        reinterpret_cast<uint4*>(c.v_res + (int64_t)u * c.v_res_cap * kD)[i] =
            __ldg(reinterpret_cast<const uint4*>(v + ((int64_t)u * n + nqv) * kD) + i);
    if (u == 0 && threadIdx.x == 0) {
        c.state[ST_TK] = nqk; c.state[ST_R] = r; c.state[ST_TV] = nqv; c.state[ST_L] = L;
        c.state[ST_VHEAD] = 0; c.state[ST_KVLEN] = n; c.state[6] = 0; c.state[7] = 0;
    }
}

// One decode step's bookkeeping (models/llama_kivi.py:343-356, :386-399): the data movement is done by
// the decode kernel per unit; the lengths advance here, once per step for all layers.
__global__ void advance_kernel(int* state, int R, int v_res_cap)
{
    int tk = state[ST_TK], r = state[ST_R], tv = state[ST_TV], L = state[ST_L], vh = state[ST_VHEAD];
    r += 1;
    if (r == R) { tk += R; r = 0; }
    L += 1;
    if (L > R) { tv += 1; vh = (vh + 1) % v_res_cap; L = R; }
    state[ST_TK] = tk; state[ST_R] = r; state[ST_TV] = tv; state[ST_L] = L; state[ST_VHEAD] = vh;
    state[ST_KVLEN] += 1;
}

// ------------------------------------------------------------------------------------------------
// export to the reference layouts (the 9-tuple of models/llama_kivi.py:454-455)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
export_k_kernel(CacheDesc c, int tk, uint32_t* __restrict__ code, __half* __restrict__ scale, __half* __restrict__ mn)
{
    // code [U][128][tk/fpi] words, scale/mn [U][128][tk/g]
    const int bits = c.k_bits, fpi = 32 / bits, g = c.g;
    const int cbk = k_cell_bytes(bits);
    const int u = blockIdx.y;
    const uint8_t* ubase = c.k_store + (int64_t)u * k_unit_bytes(c.k_cap_blocks, bits, g);
    const int wpr = tk / fpi, gpr = tk / g;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kD * wpr; i += gridDim.x * blockDim.x) {
        const int d = i / wpr, w = i % wpr;
        const int tok = w * fpi, blk = tok / kBlockTokens, bt = tok % kBlockTokens;
        code[((int64_t)u * kD + d) * wpr + w] =
            *reinterpret_cast<const uint32_t*>(ubase + k_row_off(blk, d, bits, g) + (bt / kCell) * cbk + ((bt % kCell) / fpi) * 4);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kD * gpr; i += gridDim.x * blockDim.x) {
        const int d = i / gpr, gi = i % gpr;
        const int tok = gi * g, blk = tok / kBlockTokens, bt = tok % kBlockTokens;
        const __half2 m = *reinterpret_cast<const __half2*>(ubase + k_meta_off(blk, d, bits, g) + (bt / g) * 4);
        scale[((int64_t)u * kD + d) * gpr + gi] = __low2half(m);
        mn[((int64_t)u * kD + d) * gpr + gi] = __high2half(m);
    }
}

__global__ void __launch_bounds__(256)
export_v_kernel(CacheDesc c, int tv, int L, int vhead, int r,
                uint32_t* __restrict__ code, __half* __restrict__ scale, __half* __restrict__ mn,
                __half* __restrict__ k_full, __half* __restrict__ v_full)
{
    const int fpi = 32 / c.v_bits, wpt = kD / fpi, gpt = kD / c.g;
    const int u = blockIdx.y;
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = tid; i < tv * wpt; i += stride)
        code[(int64_t)u * tv * wpt + i] = reinterpret_cast<const uint32_t*>(c.v_codes)[(int64_t)u * c.v_cap * wpt + i];
    for (int i = tid; i < tv * gpt; i += stride) {
        const __half2 m = reinterpret_cast<const __half2*>(c.v_meta)[(int64_t)u * c.v_cap * gpt + i];
        scale[(int64_t)u * tv * gpt + i] = __low2half(m);
        mn[(int64_t)u * tv * gpt + i] = __high2half(m);
    }
    for (int i = tid; i < r * kD; i += stride)
        k_full[(int64_t)u * r * kD + i] = c.k_res[(int64_t)u * c.R * kD + i];
    for (int i = tid; i < L * kD; i += stride) {
        const int t = i / kD, d = i % kD;
        v_full[(int64_t)u * L * kD + i] = c.v_res[((int64_t)u * c.v_res_cap + (vhead + t) % c.v_res_cap) * kD + d];
    }
}

static int validate(const kivi_cache_t* k, CacheDesc* d)
{
    if (!k) return KIVI_ERR_NULL;
    if (!(k->k_bits == 2 || k->k_bits == 4) || !(k->v_bits == 2 || k->v_bits == 4)) return KIVI_ERR_BITS;
    if (k->head_dim != kD) return KIVI_ERR_SHAPE;
    if (!(k->group_size == 32 || k->group_size == 64 || k->group_size == 128)) return KIVI_ERR_GROUP;
    if (k->residual_length <= 0 || k->residual_length % k->group_size != 0) return KIVI_ERR_SHAPE;   // llama_kivi.py:344
    if (k->residual_length > 256) return KIVI_ERR_UNSUPPORTED;
    if (k->batch <= 0 || k->num_kv_heads <= 0 || k->num_heads % k->num_kv_heads != 0) return KIVI_ERR_GQA;
    if (k->k_cap_blocks <= 0 || k->v_cap <= 0 || k->v_cap % 256 != 0) return KIVI_ERR_SHAPE;
    if (k->v_res_cap < k->residual_length + 1) return KIVI_ERR_SHAPE;
    if (!k->k_store || !k->v_codes || !k->v_meta || !k->k_res || !k->v_res || !k->state) return KIVI_ERR_NULL;
    d->B = k->batch; d->Hkv = k->num_kv_heads; d->H = k->num_heads; d->k_bits = k->k_bits; d->v_bits = k->v_bits;
    d->g = k->group_size; d->R = k->residual_length; d->k_cap_blocks = k->k_cap_blocks; d->v_cap = k->v_cap;
    d->v_res_cap = k->v_res_cap;
    d->k_store = (uint8_t*)k->k_store; d->v_codes = (uint8_t*)k->v_codes; d->v_meta = (uint8_t*)k->v_meta;
    d->k_res = (__half*)k->k_res; d->v_res = (__half*)k->v_res; d->state = (int*)k->state;
    return KIVI_OK;
}

int make_desc(const kivi_cache_t* k, CacheDesc* d) { return validate(k, d); }

}  // namespace kivi

using namespace kivi;

extern "C" int kivi_cache_sizes(int batch, int num_kv_heads, int k_bits, int v_bits, int group_size,
                                int residual_length, int max_tokens, int64_t* out /* [8] */)
{
    if (!out) return KIVI_ERR_NULL;
    if (!(k_bits == 2 || k_bits == 4) || !(v_bits == 2 || v_bits == 4)) return KIVI_ERR_BITS;
    if (!(group_size == 32 || group_size == 64 || group_size == 128)) return KIVI_ERR_GROUP;
    if (residual_length <= 0 || residual_length % group_size != 0 || batch <= 0 || num_kv_heads <= 0 || max_tokens <= 0)
        return KIVI_ERR_SHAPE;
    const int64_t U = (int64_t)batch * num_kv_heads;
    const int64_t capb = cdiv(max_tokens, kBlockTokens) + 1;
    const int64_t vcap = (int64_t)cdiv(max_tokens, 256) * 256 + 256;
    const int64_t vres = residual_length + 1;
    out[0] = capb;                                                // k_cap_blocks
    out[1] = vcap;                                                // v_cap
    out[2] = vres;                                                // v_res_cap
    out[3] = U * k_unit_bytes((int)capb, k_bits, group_size);    // bytes: k_store
    out[4] = U * vcap * v_tok_code_bytes(v_bits);                 // bytes: v_codes
    out[5] = U * vcap * v_tok_meta_bytes(group_size) + 64;        // bytes: v_meta (+ tail slack for 16-B rounding)
    out[6] = U * residual_length * kD * 2;                        // bytes: k_res
    out[7] = U * vres * kD * 2;                                   // bytes: v_res
    return KIVI_OK;
}

extern "C" int kivi_cache_prefill_f16(const kivi_cache_t* cache, const void* k, const void* v, int n, void* stream)
{
    CacheDesc c;
    int rc = make_desc(cache, &c);
    if (rc) return rc;
    if (n < 0) return KIVI_ERR_SHAPE;
    if (n > 0 && (!k || !v)) return KIVI_ERR_NULL;
    const int R = c.R;
    // models/llama_kivi.py:425-434 (K) and :442-449 (V)
    const int nqk = (n % R != 0) ? (n < R ? 0 : n - n % R) : n;
    const int nqv = (n <= R) ? 0 : n - R;
    if (cdiv(nqk, kBlockTokens) > c.k_cap_blocks || nqv > c.v_cap) return KIVI_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    const int U = c.B * c.Hkv;
    if (U > 65535) return KIVI_ERR_SHAPE;
    if (nqk > 0) {
        dim3 grid(cdiv(nqk, kBlockTokens), U);
        if (c.k_bits == 2) k_prefill_kernel<2><<<grid, 256, 0, st>>>(c, (const __half*)k, n, nqk);
        else               k_prefill_kernel<4><<<grid, 256, 0, st>>>(c, (const __half*)k, n, nqk);
        rc = post_launch(); if (rc) return rc;
    }
    if (nqv > 0) {
        const int fpi = 32 / c.v_bits;
        int lg = 0; while ((fpi << lg) < c.g) ++lg;
        const int64_t total = (int64_t)U * nqv * (kD / fpi);
        const int64_t blocks = cdiv64(total, 256);
        if (blocks > 0x7fffffffLL) return KIVI_ERR_SHAPE;
        if (c.v_bits == 2) v_prefill_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(c, (const __half*)v, n, nqv, lg);
        else               v_prefill_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(c, (const __half*)v, n, nqv, lg);
        rc = post_launch(); if (rc) return rc;
    }
    residual_prefill_kernel<<<U, 256, 0, st>>>(c, (const __half*)k, (const __half*)v, n, nqk, nqv);
    return post_launch();
}

extern "C" int kivi_cache_advance(const kivi_cache_t* cache, void* stream)
{
    CacheDesc c;
    int rc = make_desc(cache, &c);
    if (rc) return rc;
    advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(c.state, c.R, c.v_res_cap);
    return post_launch();
}

extern "C" int kivi_cache_export_f16(const kivi_cache_t* cache, int tk, int r, int tv, int L, int vhead,
                                     void* k_code, void* k_scale, void* k_mn, void* k_full,
                                     void* v_code, void* v_scale, void* v_mn, void* v_full, void* stream)
{
    CacheDesc c;
    int rc = make_desc(cache, &c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int U = c.B * c.Hkv;
    if (tk > 0) {
        if (!k_code || !k_scale || !k_mn) return KIVI_ERR_NULL;
        export_k_kernel<<<dim3(8, U), 256, 0, st>>>(c, tk, (uint32_t*)k_code, (__half*)k_scale, (__half*)k_mn);
        rc = post_launch(); if (rc) return rc;
    }
    export_v_kernel<<<dim3(8, U), 256, 0, st>>>(c, tv, L, vhead, r, (uint32_t*)v_code, (__half*)v_scale, (__half*)v_mn,
                                               (__half*)k_full, (__half*)v_full);
    return post_launch();
}
