// cal_fused v3 / v2t (see rip_v2_core.cuh, "v3"): the role-split and TMA-fed variants of the fused kernel, kept behind
// the development selector (RIP_FUSED_VARIANT / params.threads < 0) for A/B measurements.  Own translation unit: the
// sixteen instantiations compile in parallel with rip_v2.cu.
#include "rip_handle.h"
#include "rip_launch.h"
#include "rip_v2_core.cuh"

namespace rip {

__constant__ RampPlanDev c_plan_v3;
__constant__ v2::FastTab c_fast_v3;

namespace v2 {

// v3: role-split CTA (rip_v2_core.cuh, "v3"): 256 threads, X = warps 0-3, Y = warps 4-7, TMA-fed raw ring
template <int G, int P, bool BX, int MINB, int XR, int YR>
__global__ void __launch_bounds__(2 * TW, MINB) cal_fused_v3_kernel(const Args A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    v3_body<G, P, BX, XR, YR>(A, c_plan_v3, c_fast_v3, smem_raw);
}

template <int G, int P, int MINB>
__global__ void __launch_bounds__(TW, MINB) cal_fused_v2t_kernel(const Args A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    v2t_body<G, P>(A, c_plan_v3, c_fast_v3, smem_raw);
}

static void configure_once(const void* fn, size_t smem) { configure_smem_once(fn, smem, true); }

template <int G, int P, bool BX>
static void launch_v3(const Args& A, cudaStream_t st) {
    constexpr int MINB = (G <= 8) ? 3 : 2;
    const size_t smem = v3_smem_bytes<G>(BX);
    // launch allocation 80 (G <= 8: 3 CTAs/SM) or 128 (2 CTAs/SM) registers per thread, re-split between the roles
    constexpr int XR = BX ? ((G <= 8) ? 72 : 112) : ((G <= 8) ? 64 : 96);
    constexpr int YR = BX ? ((G <= 8) ? 88 : 144) : ((G <= 8) ? 96 : 160);
    auto kern = cal_fused_v3_kernel<G, P, BX, MINB, XR, YR>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, 2 * TW, smem, st, A);
}

template <int G, int P>
static void launch_v2t(const Args& A, cudaStream_t st) {
    const size_t smem = Smem<G>::bytes();
    auto kern = cal_fused_v2t_kernel<G, P, (G <= 8) ? 4 : 2>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, TW, smem, st, A);
}

}  // namespace v2

// variant: 0 = v2 (one role, 128 threads), 1 = v3 (X: a0 a1 | Y: b c), 2 = v3 (X: a0 a1 b | Y: c)
void launch_cal_fused_v3(const v2::Args& A, int G, int P, int variant, cudaStream_t st) {
    P = v2_pad_P(P);
    const bool bx = variant == 2;
#define RIP_V3(GG, PP) \
    if (G == GG && P == PP) { if (variant == 3) v2::launch_v2t<GG, PP>(A, st); else if (bx) v2::launch_v3<GG, PP, true>(A, st); else v2::launch_v3<GG, PP, false>(A, st); return; }
    RIP_V3(8, 11) RIP_V3(8, 4) RIP_V3(16, 11) RIP_V3(16, 4)
#undef RIP_V3
    throw Error("cal_fused v3: unsupported (G, P)");
}

void v3_plan_to_device(const rip_ramp_plan* plan, const void* fast_tab, cudaStream_t st) {
    RIP_CUDA(cudaMemcpyToSymbolAsync(c_plan_v3, plan, sizeof(rip_ramp_plan), 0, cudaMemcpyHostToDevice, st));
    RIP_CUDA(cudaMemcpyToSymbolAsync(c_fast_v3, fast_tab, sizeof(v2::FastTab), 0, cudaMemcpyHostToDevice, st));
}

}  // namespace rip
