// instantiation of the decode attention kernels for k_bits = 2, v_bits = 4 (all G, all group sizes)
#include "kivi_attn.cuh"
namespace kivi {
int attention_k2v4(AttnParams& p, int G, cudaStream_t st) { return dispatch_attention<2, 4>(p, G, st); }
}
