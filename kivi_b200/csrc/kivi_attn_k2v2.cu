// instantiation of the fused decode attention kernel for k_bits = 2, v_bits = 2 (all G, all group sizes)
#include "kivi_attn.cuh"
namespace kivi {
int attention_k2v2(AttnParams& p, int G, int max_kv_len, cudaStream_t st) { return dispatch_attention<2, 2>(p, G, max_kv_len, st); }
}
