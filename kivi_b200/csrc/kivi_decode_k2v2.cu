// instantiation of the fused decode kernel for k_bits = 2, v_bits = 2 (all G, all group sizes)
#include "kivi_decode_impl.cuh"
namespace kivi {
int decode_k2v2(DecodeParams& p, int G, int max_kv_len, cudaStream_t st) { return dispatch_decode<2, 2>(p, G, max_kv_len, st); }
}
