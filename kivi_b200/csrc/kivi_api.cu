// kivi_api.cu -- library identification, error strings, launch counter.
#include "kivi_common.cuh"

#include <cstdlib>
#include <mutex>

namespace kivi {
std::atomic<unsigned long long> g_launch_count{0};

int device_info(DeviceInfo* out)
{
    static DeviceInfo table[kMaxDevices];
    static std::atomic<bool> known[kMaxDevices];
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= kMaxDevices) return KIVI_ERR_UNSUPPORTED;
    if (!known[dev].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(mu);
        if (!known[dev].load(std::memory_order_relaxed)) {
            DeviceInfo d{dev, 0, 0};
            e = cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            if (e != cudaSuccess) return (int)e;
            table[dev] = d;
            known[dev].store(true, std::memory_order_release);
        }
    }
    *out = table[dev];
    return KIVI_OK;
}

const Tuning& tuning()
{
    static const Tuning t = [] {
        auto geti = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
        Tuning v;
        v.gqa_g = geti("KIVI_GQA_G");
        v.ctas_per_sm = geti("KIVI_CTAS_PER_SM");
        v.stages_per_warp = geti("KIVI_STAGES_PER_WARP");
        v.no_pdl = getenv("KIVI_NO_PDL") != nullptr;
        v.no_mma_gemv = getenv("KIVI_NO_MMA_GEMV") != nullptr;       // A/B: SIMT kernels only (tools/microbench.py)
        return v;
    }();
    return t;
}
}

extern "C" int kivi_version(void) { return 100; }   // 0.1.0

extern "C" uint64_t kivi_launch_count(void) { return kivi::g_launch_count.load(); }

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
