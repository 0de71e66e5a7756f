// instantiation of the fused decode kernel for k_bits = 4, v_bits = 4 (all G, all group sizes)
#include "kivi_decode_impl.cuh"
namespace kivi {
int decode_k4v4(DecodeParams& p, int G, int max_kv_len, cudaStream_t st) { return dispatch_decode<4, 4>(p, G, max_kv_len, st); }
}
