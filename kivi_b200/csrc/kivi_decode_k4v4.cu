// instantiation of the fused and split decode kernels for k_bits = 4, v_bits = 4 (all G, all group sizes)
#include "kivi_decode_split.cuh"
namespace kivi {
int decode_k4v4(DecodeParams& p, int G, int max_kv_len, cudaStream_t st) { return dispatch_decode<4, 4>(p, G, max_kv_len, st); }
int decode_split_k4v4(SplitParams& sp, int G, cudaStream_t st) { return dispatch_decode_split<4, 4>(sp, G, st); }
}
