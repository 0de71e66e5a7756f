// kivi_decode.cu -- C-ABI entry of the fused decode attention; the kernel lives in kivi_decode_impl.cuh and is
// instantiated per (k_bits, v_bits) pair in kivi_decode_k{2,4}v{2,4}.cu (compiled in parallel).
#include <cstdlib>
#include "kivi_decode_split.cuh"

namespace kivi {
int decode_k2v2(DecodeParams& p, int G, int max_kv_len, cudaStream_t st);
int decode_k4v4(DecodeParams& p, int G, int max_kv_len, cudaStream_t st);
int decode_k2v4(DecodeParams& p, int G, int max_kv_len, cudaStream_t st);
int decode_k4v2(DecodeParams& p, int G, int max_kv_len, cudaStream_t st);
int decode_split_k2v2(SplitParams& sp, int G, cudaStream_t st);
int decode_split_k4v4(SplitParams& sp, int G, cudaStream_t st);
int decode_split_k2v4(SplitParams& sp, int G, cudaStream_t st);
int decode_split_k4v2(SplitParams& sp, int G, cudaStream_t st);
}

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
    int G = ratio % 4 == 0 ? 4 : (ratio % 2 == 0 ? 2 : 1);
    if (const char* e = getenv("KIVI_GQA_G")) { const int g = atoi(e); if ((g == 1 || g == 2 || g == 4) && ratio % g == 0) G = g; }
    p.hchunks = ratio / G;
    p.n_units = p.c.B * p.c.Hkv * p.hchunks;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.c.k_bits == 2 && p.c.v_bits == 2) return decode_k2v2(p, G, max_kv_len, st);
    if (p.c.k_bits == 4 && p.c.v_bits == 4) return decode_k4v4(p, G, max_kv_len, st);
    if (p.c.k_bits == 2 && p.c.v_bits == 4) return decode_k2v4(p, G, max_kv_len, st);
    if (p.c.k_bits == 4 && p.c.v_bits == 2) return decode_k4v2(p, G, max_kv_len, st);
    return KIVI_ERR_BITS;
}

extern "C" int kivi_decode_attention_split_f16(const kivi_cache_t* cache, const void* q, const void* k_new, const void* v_new,
                                               const void* mask, void* out, void* workspace, int64_t ld,
                                               void* dbg_logits, void* dbg_probs, int64_t dbg_stride, void* stream)
{
    SplitParams sp;
    DecodeParams& p = sp.d;
    int rc = make_desc(cache, &p.c);
    if (rc) return rc;
    if (!q || !k_new || !v_new || !out || !workspace) return KIVI_ERR_NULL;
    if (ld <= 0 || ld % 8 != 0) return KIVI_ERR_ALIGN;
    if (ld > 40960) return KIVI_ERR_CAPACITY;                 // softmax_rows_kernel keeps a row in registers
    if (reinterpret_cast<uintptr_t>(workspace) % 16 != 0) return KIVI_ERR_ALIGN;
    p.q = (const __half*)q; p.k_new = (const __half*)k_new; p.v_new = (const __half*)v_new; p.mask = (const __half*)mask;
    p.out = (__half*)out; p.dbg_logits = (__half*)dbg_logits; p.dbg_probs = (__half*)dbg_probs; p.dbg_stride = dbg_stride;
    p.t_cap = 0;
    sp.ws = (__half*)workspace; sp.ld = ld; sp.team = 1;
    const int ratio = p.c.H / p.c.Hkv;
    const int G = ratio % 4 == 0 ? 4 : (ratio % 2 == 0 ? 2 : 1);
    p.hchunks = ratio / G;
    p.n_units = p.c.B * p.c.Hkv * p.hchunks;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.c.k_bits == 2 && p.c.v_bits == 2) return decode_split_k2v2(sp, G, st);
    if (p.c.k_bits == 4 && p.c.v_bits == 4) return decode_split_k4v4(sp, G, st);
    if (p.c.k_bits == 2 && p.c.v_bits == 4) return decode_split_k2v4(sp, G, st);
    if (p.c.k_bits == 4 && p.c.v_bits == 2) return decode_split_k4v2(sp, G, st);
    return KIVI_ERR_BITS;
}
