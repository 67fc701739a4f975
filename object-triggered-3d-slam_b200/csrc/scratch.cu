// scratch.cu -- per-device cache of device allocations for the operators' temporaries and results
// (see common.cuh: DevBuf).  Best-fit reuse of released blocks; cudaMalloc only on a miss.
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace otslam {

thread_local uint64_t g_error_count = 0;

namespace {
struct Block {
    void* p;
    size_t bytes;
};
struct DeviceCache {
    std::vector<Block> free_list;
    size_t cached_bytes = 0;
};
std::mutex g_mu;
std::unordered_map<int, DeviceCache> g_cache;
std::unordered_map<void*, std::pair<int, size_t>> g_live;     // ptr -> (device, bytes)
constexpr size_t kMaxCachedBytes = 8ull << 30;                // per device; 180 GB of HBM make this cheap

void trim_locked(int dev, size_t keep_bytes) {
    DeviceCache& c = g_cache[dev];
    if (c.cached_bytes <= keep_bytes) return;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    std::sort(c.free_list.begin(), c.free_list.end(), [](const Block& a, const Block& b) { return a.bytes < b.bytes; });
    while (!c.free_list.empty() && c.cached_bytes > keep_bytes) {        // largest first
        cudaFree(c.free_list.back().p);
        c.cached_bytes -= c.free_list.back().bytes;
        c.free_list.pop_back();
    }
    if (cur != dev) cudaSetDevice(cur);
}
}  // namespace

cudaError_t scratch_alloc(void** p, size_t bytes) {
    *p = nullptr;
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        DeviceCache& c = g_cache[dev];
        int best = -1;
        const size_t limit = bytes * 2 + (1u << 20);
        for (int i = 0; i < (int)c.free_list.size(); ++i) {
            const size_t b = c.free_list[i].bytes;
            if (b >= bytes && b <= limit && (best < 0 || b < c.free_list[best].bytes)) best = i;
        }
        if (best >= 0) {
            const Block blk = c.free_list[best];
            c.free_list[best] = c.free_list.back();
            c.free_list.pop_back();
            c.cached_bytes -= blk.bytes;
            g_live[blk.p] = {dev, blk.bytes};
            *p = blk.p;
            return cudaSuccess;
        }
    }
    e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {                    // give the cache back to the driver and retry once
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(g_mu);
            trim_locked(dev, 0);
        }
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_mu);
    g_live[*p] = {dev, bytes};
    return cudaSuccess;
}

void scratch_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_live.find(p);
    if (it == g_live.end()) {                                // not ours (defensive): plain free
        cudaFree(p);
        return;
    }
    const int dev = it->second.first;
    const size_t bytes = it->second.second;
    g_live.erase(it);
    DeviceCache& c = g_cache[dev];
    c.free_list.push_back({p, bytes});
    c.cached_bytes += bytes;
    if (c.cached_bytes > kMaxCachedBytes) trim_locked(dev, kMaxCachedBytes / 2);
}

}  // namespace otslam

extern "C" int otslam_host_alloc(uint64_t bytes, void** out) {
    if (!out) return otslam::set_error(OTSLAM_ERR_INVALID, "null output");
    *out = nullptr;
    OT_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return OTSLAM_OK;
}

extern "C" int otslam_host_free(void* p) {
    if (p) OT_CUDA(cudaFreeHost(p));
    return OTSLAM_OK;
}

extern "C" int otslam_trim_scratch(void) {
    std::lock_guard<std::mutex> lk(otslam::g_mu);
    for (auto& kv : otslam::g_cache) otslam::trim_locked(kv.first, 0);
    return OTSLAM_OK;
}
