// Internal runtime helpers shared by the translation units of librip_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rip_b200.h"

namespace rip {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define RIP_CUDA(expr)                                                                                 \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char b_[512];                                                                              \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw rip::Error(b_);                                                                      \
        }                                                                                              \
    } while (0)

#define RIP_REQUIRE(cond, ...)                     \
    do {                                           \
        if (!(cond)) {                             \
            char b_[512];                          \
            snprintf(b_, sizeof b_, __VA_ARGS__);  \
            throw rip::Error(b_);                  \
        }                                          \
    } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches
#define RIP_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
        ++rip::g_launches;                                                 \
        RIP_CUDA(cudaGetLastError());                                      \
    } while (0)

// C-ABI boundary: no exception crosses it
#define RIP_API_BEGIN try {
#define RIP_API_END                                  \
    return 0;                                        \
    }                                                \
    catch (const std::exception& e) {                \
        rip::set_error("%s", e.what());              \
        return 1;                                    \
    }                                                \
    catch (...) {                                    \
        rip::set_error("unknown error");             \
        return 2;                                    \
    }

inline void use_device(int device) { RIP_CUDA(cudaSetDevice(device)); }

// RAII device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) return;
        RIP_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void upload(const T* host, size_t count, cudaStream_t st = 0) {
        if (n < count) alloc(count);
        RIP_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void download(T* host, size_t count, cudaStream_t st = 0) const {
        RIP_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    void zero(cudaStream_t st = 0) {
        if (p) RIP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    }
};

inline size_t dtype_size(int dt) {
    switch (dt) {
        case RIP_F32: return 4;
        case RIP_F64: return 8;
        case RIP_I32: return 4;
        case RIP_U16: return 2;
    }
    throw Error("bad dtype tag");
}

// upload an untyped host plane
struct DevRaw {
    void* p = nullptr;
    size_t bytes = 0;
    DevRaw() {}
    DevRaw(const DevRaw&) = delete;
    DevRaw& operator=(const DevRaw&) = delete;
    ~DevRaw() { if (p) cudaFree(p); }
    void alloc(size_t b) {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        if (!b) return;
        RIP_CUDA(cudaMalloc(&p, b));
        bytes = b;
    }
    void upload(const void* host, size_t b, cudaStream_t st = 0) {
        if (bytes < b) alloc(b);
        RIP_CUDA(cudaMemcpyAsync(p, host, b, cudaMemcpyHostToDevice, st));
    }
};

// device plan cache (ramp plan in __constant__ memory + exact weights in global memory); defined in rip_fit.cu
const double* plan_to_device(int device, const rip_ramp_plan* plan, const double* w_exact, cudaStream_t st);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (device, function): configure each pair once (rip_rt.cu)
void configure_smem_once(const void* fn, size_t smem, bool prefer_smem_carveout = false);

}  // namespace rip
