// instantiation of the decode attention kernels for k_bits = 4, v_bits = 2 (all G, all group sizes)
#include "kivi_attn.cuh"
namespace kivi {
int attention_k4v2(AttnParams& p, int G, bool overlap_prologue, cudaStream_t st) { return dispatch_attention<4, 2>(p, G, overlap_prologue, st); }
int64_t workspace_k4v2(const CacheDesc& c, int n_units, int G, int max_kv_len, void* base, Workspace* w) {
    return carve_workspace(c, n_units, G, max_kv_len, base, w);
}
}

#if KIVI_TIMELINE
namespace kivi { int timeline_k4v2(unsigned long long* host_out) { return timeline_fetch(host_out); } }   // tuning builds (tools/timeline.py)
#endif
