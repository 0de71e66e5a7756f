// instantiation of the fused decode attention kernel for k_bits = 4, v_bits = 4 (all G, all group sizes)
#include "kivi_attn.cuh"
namespace kivi {
int attention_k4v4(AttnParams& p, int G, int max_kv_len, cudaStream_t st) { return dispatch_attention<4, 4>(p, G, max_kv_len, st); }
}
