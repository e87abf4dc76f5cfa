// cp_mem.cpp -- caching device allocator of libcudapath.so.
//
// A render job is: create a context, upload, build the BVH (about twenty-five scratch arrays, up to several GB each for a
// 32 M-reference scene), render (twenty queue arrays, 14 GB at the default wave size), destroy.  Going to the driver for each
// of these costs 0.2-0.6 s per job: cudaMalloc / cudaFree map and unmap physical memory and synchronise the device, and the
// stream-ordered pool (cudaMallocAsync) that was used before remaps physical pages whenever it cannot find a contiguous range
// for a multi-GB request, which made the build of an identical scene take anywhere between 15 and 650 ms.
//
// Freed blocks are therefore kept in a per-device, size-keyed free list and handed out again to the next request of (nearly)
// the same size -- a job that repeats sees only exact hits.  Nothing is returned to the driver until cudapath_trim_memory(),
// a failed cudaMalloc (the cache is emptied, the request retried) or process exit.
//
// Contract of dev_free(): no work that touches the block may still be pending on the device (the callers synchronise the
// stream they used first), because the next owner may use it on another stream.
#include "cp_host.h"
#include <map>
#include <mutex>
#include <unordered_map>

namespace cp {

namespace {
struct DevCache {
    std::mutex m;
    std::multimap<size_t, void *> freeBlocks;          // size -> block
    std::unordered_map<void *, size_t> live;           // block -> size (handed out)
    size_t cachedBytes = 0;
};
DevCache &cache_for_current_device() {
    static std::mutex gm;
    static std::map<int, DevCache *> caches;
    int dev = 0; cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(gm);
    DevCache *&c = caches[dev];
    if (!c) c = new DevCache();
    return *c;
}
size_t round_size(size_t b) {
    if (b == 0) b = 1;
    const size_t q = b < (1u << 20) ? 512 : (size_t) 2 << 20;       // 2 MB: the granularity the driver maps large allocations with
    return (b + q - 1) / q * q;
}
void trim_locked(DevCache &c) {
    for (auto &kv : c.freeBlocks) cudaFree(kv.second);
    c.freeBlocks.clear(); c.cachedBytes = 0;
}
}

cudaError_t dev_alloc(void **p, size_t bytes) {
    DevCache &c = cache_for_current_device();
    const size_t need = round_size(bytes);
    std::lock_guard<std::mutex> g(c.m);
    auto it = c.freeBlocks.lower_bound(need);
    if (it != c.freeBlocks.end() && it->first <= need + need / 8 + 4096) {          // best fit, at most 12.5 % larger
        *p = it->second; c.live[*p] = it->first; c.cachedBytes -= it->first; c.freeBlocks.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess && !c.freeBlocks.empty()) {                                   // out of memory: give the cached blocks back, retry
        cudaGetLastError();
        trim_locked(c);
        e = cudaMalloc(p, need);
    }
    if (e == cudaSuccess) c.live[*p] = need;
    return e;
}

void dev_free(const void *p) {
    if (!p) return;
    DevCache &c = cache_for_current_device();
    std::lock_guard<std::mutex> g(c.m);
    auto it = c.live.find(const_cast<void *>(p));
    if (it == c.live.end()) { cudaFree(const_cast<void *>(p)); return; }               // not ours (should not happen)
    c.freeBlocks.emplace(it->second, it->first); c.cachedBytes += it->second;
    c.live.erase(it);
}

size_t dev_cached_bytes() {
    DevCache &c = cache_for_current_device();
    std::lock_guard<std::mutex> g(c.m);
    return c.cachedBytes;
}

void dev_trim() {
    DevCache &c = cache_for_current_device();
    std::lock_guard<std::mutex> g(c.m);
    cudaDeviceSynchronize();
    trim_locked(c);
}

} // namespace cp
