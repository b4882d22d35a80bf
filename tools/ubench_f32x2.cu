// Micro-benchmark (development tool, not part of the library): issue rate of scalar vs packed (f32x2) FP32
// instructions on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = x[i].x * a; x[i].y = x[i].y * a; }            // 2 FMUL
            if (MODE == 1) { x[i] = __fmul2_rn(x[i], A); }                           // 1 FMUL2
            if (MODE == 2) { x[i].x = x[i].x + b; x[i].y = x[i].y + b; }            // 2 FADD
            if (MODE == 3) { x[i] = __fadd2_rn(x[i], B); }                           // 1 FADD2
            if (MODE == 4) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); }
            if (MODE == 5) { x[i] = __ffma2_rn(x[i], A, B); }
            if (MODE == 6) { x[i].x = x[i].x * a; x[i].y = x[i].y + b; }            // FMUL + FADD mix
            if (MODE == 7) { x[i] = __fmul2_rn(x[i], A); x[i] = __fadd2_rn(x[i], B); }  // 2 packed = 4 flops
            if (MODE == 8) { x[i].x = x[i].x * a; x[i].y = x[i].y * a; x[i].x = x[i].x + b; x[i].y = x[i].y + b; }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int lane_ops_per_iter) {
    const int blocks = 148 * 4, threads = 256, iters = 4096;
    float* d;
    cudaMalloc(&d, blocks * threads * 4);
    k<MODE><<<blocks, threads>>>(d, 16, 1.0000001f, 1e-9f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * threads * iters * lane_ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-ops/clk/SM @1.965GHz)\n", name, ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}

int main() {
    run<0>("FMUL x2 (scalar)", 16);
    run<1>("FMUL2 (packed)", 16);
    run<2>("FADD x2 (scalar)", 16);
    run<3>("FADD2 (packed)", 16);
    run<4>("FFMA x2 (scalar)", 16);
    run<5>("FFMA2 (packed)", 16);
    run<6>("FMUL+FADD mix (scalar)", 16);
    run<8>("2FMUL+2FADD (scalar)", 32);
    run<7>("FMUL2+FADD2 (packed)", 32);
    return 0;
}
