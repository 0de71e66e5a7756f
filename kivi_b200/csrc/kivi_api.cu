// kivi_api.cu -- library identification, error strings, launch counter.
#include "kivi_common.cuh"

namespace kivi {
unsigned long long g_launch_count = 0;
}

extern "C" int kivi_version(void) { return 100; }   // 0.1.0

extern "C" uint64_t kivi_launch_count(void) { return kivi::g_launch_count; }

extern "C" const char* kivi_error_string(int code)
{
    switch (code) {
        case KIVI_OK: return "ok";
        case KIVI_ERR_BITS: return "unsupported bit width";
        case KIVI_ERR_SHAPE: return "shape / divisibility requirement violated";
        case KIVI_ERR_GQA: return "nh must be a positive multiple of nh_kv";
        case KIVI_ERR_GROUP: return "unsupported group_size";
        case KIVI_ERR_ALIGN: return "pointer or stride alignment requirement violated";
        case KIVI_ERR_NULL: return "required pointer is NULL";
        case KIVI_ERR_LAYOUT: return "unknown layout id";
        case KIVI_ERR_CAPACITY: return "cache capacity exceeded";
        case KIVI_ERR_UNSUPPORTED: return "not implemented by this build";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown kivi error";
}
