// Runtime: error slot, launch counter, memory helpers of the C ABI (include/rip_b200.h).
#include "rip_rt.h"

#include <map>
#include <mutex>

namespace rip {
static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void configure_smem_once(const void* fn, size_t smem, bool prefer_smem_carveout) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;  // largest size configured so far
    int dev = 0;
    RIP_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find({dev, fn});
    if (it != done.end() && it->second >= smem) return;
    RIP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (prefer_smem_carveout) RIP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    done[{dev, fn}] = smem;
}
}  // namespace rip

using namespace rip;

extern "C" {

const char* rip_last_error(void) { return g_err; }
int rip_abi_version(void) { return RIP_ABI_VERSION; }
long long rip_launch_count(void) { return g_launches.load(); }

long rip_struct_size(int which) {
    switch (which) {
        case 0: return (long)sizeof(rip_ramp_slice);
        case 1: return (long)sizeof(rip_ramp_plan);
        case 2: return (long)sizeof(rip_caldir_desc);
        case 3: return (long)sizeof(rip_l1l2_params);
        case 4: return (long)sizeof(rip_l2_out);
        case 5: return (long)sizeof(rip_fwd_params);
    }
    return -1;
}

int rip_device_count(int* count) {
    RIP_API_BEGIN
    RIP_REQUIRE(count != nullptr, "rip_device_count: null pointer");
    RIP_CUDA(cudaGetDeviceCount(count));
    RIP_API_END
}

int rip_device_sync(int device) {
    RIP_API_BEGIN
    use_device(device);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

int rip_host_alloc(void** p, size_t bytes) {
    RIP_API_BEGIN
    RIP_REQUIRE(p != nullptr, "rip_host_alloc: null pointer");
    RIP_CUDA(cudaMallocHost(p, bytes));
    RIP_API_END
}

int rip_host_free(void* p) {
    RIP_API_BEGIN
    if (p) RIP_CUDA(cudaFreeHost(p));
    RIP_API_END
}

int rip_dev_alloc(int device, void** p, size_t bytes) {
    RIP_API_BEGIN
    RIP_REQUIRE(p != nullptr, "rip_dev_alloc: null pointer");
    use_device(device);
    RIP_CUDA(cudaMalloc(p, bytes));
    RIP_API_END
}

int rip_dev_free(int device, void* p) {
    RIP_API_BEGIN
    use_device(device);
    if (p) RIP_CUDA(cudaFree(p));
    RIP_API_END
}

int rip_copy_h2d(int device, void* dst, const void* src, size_t bytes, void* stream) {
    RIP_API_BEGIN
    use_device(device);
    RIP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    RIP_API_END
}

int rip_copy_d2h(int device, void* dst, const void* src, size_t bytes, void* stream) {
    RIP_API_BEGIN
    use_device(device);
    RIP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RIP_API_END
}

int rip_stream_sync(int device, void* stream) {
    RIP_API_BEGIN
    use_device(device);
    RIP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    RIP_API_END
}

}  // extern "C"
