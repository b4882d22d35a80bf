// Ramp fitting (stage entry points) and the fused L1->L2 kernel.  One translation unit because both
// read the ramp plan from the same __constant__ symbol (uniform operands fold into the FP instructions).
#include <mutex>

#include "rip_launch.h"

namespace rip {

__constant__ RampPlanDev c_plan;

// ---------------------------------------------------------------------------------------------------------
// plan cache: one plan resident per device; exact weights in global memory
// ---------------------------------------------------------------------------------------------------------
namespace {
struct PlanSlot {
    bool valid = false;
    rip_ramp_plan host;
    std::vector<double> w;
    double* d_w = nullptr;
    size_t d_w_cap = 0;
};
std::mutex g_plan_mu;
PlanSlot g_plan[64];
}  // namespace

const double* plan_to_device(int device, const rip_ramp_plan* plan, const double* w_exact, cudaStream_t st) {
    RIP_REQUIRE(device >= 0 && device < 64, "bad device ordinal %d", device);
    RIP_REQUIRE(plan != nullptr && w_exact != nullptr, "ramp plan / exact weights missing");
    RIP_REQUIRE(plan->G >= 3 && plan->G <= RIP_GMAX, "ramp plan: G=%d outside 3..%d", plan->G, RIP_GMAX);
    RIP_REQUIRE(plan->nvar >= 1 && plan->nvar <= RIP_MAXVAR, "ramp plan: nvar=%d outside 1..%d", plan->nvar, RIP_MAXVAR);
    const int nsl = plan->var_slice_off[plan->nvar];
    RIP_REQUIRE(nsl >= 0 && nsl <= RIP_MAXSLICE, "ramp plan: %d slices exceed %d", nsl, RIP_MAXSLICE);
    std::lock_guard<std::mutex> lk(g_plan_mu);
    PlanSlot& s = g_plan[device];
    const size_t nw = (size_t)nsl * RIP_GMAX;
    const bool same = s.valid && memcmp(&s.host, plan, sizeof(rip_ramp_plan)) == 0 && s.w.size() == nw &&
                      (nw == 0 || memcmp(s.w.data(), w_exact, nw * sizeof(double)) == 0);
    if (!same) {
        // a different plan may still be in use by kernels in flight on other streams
        RIP_CUDA(cudaDeviceSynchronize());
        s.host = *plan;
        s.w.assign(w_exact, w_exact + nw);
        if (s.d_w_cap < nw + 1) {
            if (s.d_w) cudaFree(s.d_w);
            RIP_CUDA(cudaMalloc((void**)&s.d_w, (nw + 1) * sizeof(double)));
            s.d_w_cap = nw + 1;
        }
        if (nw) RIP_CUDA(cudaMemcpyAsync(s.d_w, s.w.data(), nw * sizeof(double), cudaMemcpyHostToDevice, st));
        RIP_CUDA(cudaMemcpyToSymbolAsync(c_plan, &s.host, sizeof(rip_ramp_plan), 0, cudaMemcpyHostToDevice, st));
        v2_plan_to_device(&s.host, st);
        RIP_CUDA(cudaStreamSynchronize(st));
        s.valid = true;
    }
    return s.d_w;
}

// ---------------------------------------------------------------------------------------------------------
// stage kernels: jump_detect / ramp_fit on [G,ny,nx] cubes
// ---------------------------------------------------------------------------------------------------------
template <typename TG>
__global__ void jump_detect_kernel(const float* __restrict__ data, uint8_t* __restrict__ rdq, int ny, int nx, int nb,
                                   int variant, const TG* __restrict__ gain, const float* __restrict__ read,
                                   const double* __restrict__ w, float* __restrict__ slope, float* __restrict__ err_read,
                                   float* __restrict__ err_poisson, float* __restrict__ smap) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    const long npl = (long)ny * nx, p = (long)y * nx + x;
    const int ngrp = c_plan.var_ngrp[variant];
    float d[RIP_GMAX];
#pragma unroll
    for (int g = 0; g < RIP_GMAX; ++g) d[g] = (g < c_plan.G) ? data[(long)g * npl + p] : 0.0f;
    const bool active = (y >= nb && y < ny - nb && x >= nb && x < nx - nb);
    FitResult r = jump_detect_pixel<RIP_GMAX, TG, false>(d, variant, gain[p], read[p], active, c_plan, w,
                                                         smap ? smap + p : nullptr, npl);
    slope[p] = r.slope;
    err_read[p] = r.err_read;
    err_poisson[p] = r.err_poisson;
    for (int g = 0; g < ngrp; ++g)
        if ((r.jump_mask >> g) & 1u) rdq[(long)g * npl + p] |= (uint8_t)DQ_JUMP_DET;
}

template <typename TG, bool FAST>
__global__ void ramp_fit_kernel(const float* __restrict__ data, uint8_t* __restrict__ rdq, uint32_t* __restrict__ pdq,
                                int ny, int nx, int nb, const TG* __restrict__ gain, const float* __restrict__ read,
                                const double* __restrict__ w, float* __restrict__ slope, float* __restrict__ err_read,
                                float* __restrict__ err_poisson) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    const long npl = (long)ny * nx, p = (long)y * nx + x;
    const int G = c_plan.G;
    float d[RIP_GMAX];
    GroupFlags gf = {0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int g = 0; g < RIP_GMAX; ++g) {
        d[g] = (g < G) ? data[(long)g * npl + p] : 0.0f;
        if (g < G) {
            const uint32_t q = rdq[(long)g * npl + p];
            if (q & DQ_DO_NOT_USE) gf.dnu |= 1u << g;
            if (q & DQ_SATURATED) gf.sat |= 1u << g;
            if (q & DQ_JUMP_DET) gf.jump |= 1u << g;
            if (q & DQ_AD_FLOOR) gf.adf |= 1u << g;
            if (!(q & DQ_SATURATED)) gf.other_unsat |= q & ~(DQ_DO_NOT_USE | DQ_SATURATED | DQ_JUMP_DET | DQ_AD_FLOOR);
        }
    }
    const uint32_t jump_in = gf.jump;
    const bool active = (y >= nb && y < ny - nb && x >= nb && x < nx - nb);
    uint32_t pd = pdq[p];
    FitResult r = ramp_fit_pixel<RIP_GMAX, TG, FAST>(d, gf, pd, gain[p], read[p], active, c_plan, w);
    slope[p] = r.slope;
    err_read[p] = r.err_read;
    err_poisson[p] = r.err_poisson;
    pdq[p] = pd;
    const uint32_t added = gf.jump & ~jump_in;
    for (int g = 0; g < G; ++g)
        if ((added >> g) & 1u) rdq[(long)g * npl + p] |= (uint8_t)DQ_JUMP_DET;
}

// ---------------------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------------------
template <int GMAX, int PMAX, typename TG, typename TK>
__global__ void __launch_bounds__(256) cal_fused_kernel(const CalArgs A) {
    typedef typename Promote<float, TG>::type TIM;
    typedef typename Promote<TIM, TK>::type TI;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CalSmem<GMAX, TIM, TI> sm;
    sm.carve(smem_raw, blockDim.x);
    const int tw = blockDim.x - 6;
    const int c0 = blockIdx.x * tw;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = min(r0 + A.band_rows, A.n);
    for (int s = r0 - 3; s <= r1 + 5; ++s) {
        cal_step<GMAX, PMAX, TG, TK>(A, c_plan, sm, threadIdx.x, blockDim.x, c0, r0, r1, s);
        __syncthreads();
    }
}

template <int GMAX, int PMAX, typename TG, typename TK>
static void launch_cal_t(const CalArgs& A, int threads, cudaStream_t st) {
    typedef typename Promote<float, TG>::type TIM;
    typedef typename Promote<TIM, TK>::type TI;
    const size_t smem = CalSmem<GMAX, TIM, TI>::bytes(threads);
    auto kern = cal_fused_kernel<GMAX, PMAX, TG, TK>;
    if (smem > 48 * 1024) configure_smem_once((const void*)kern, smem);
    const int tw = threads - 6;
    dim3 grid((A.n + tw - 1) / tw, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, threads, smem, st, A);
}

template <typename TG, typename TK>
static void launch_cal_gp(const CalArgs& A, int threads, cudaStream_t st) {
    if (A.G <= 8) {
        if (A.P <= 4) launch_cal_t<8, 4, TG, TK>(A, threads, st);
        else if (A.P <= 11) launch_cal_t<8, 11, TG, TK>(A, threads, st);
        else launch_cal_t<8, RIP_PMAX, TG, TK>(A, threads, st);
    } else {
        if (A.P <= 4) launch_cal_t<16, 4, TG, TK>(A, threads, st);
        else if (A.P <= 11) launch_cal_t<16, 11, TG, TK>(A, threads, st);
        else launch_cal_t<16, RIP_PMAX, TG, TK>(A, threads, st);
    }
}

size_t cal_fused_smem_bytes(int G, int g_dtype, int k_dtype, int threads) {
    const int gm = (G <= 8) ? 8 : 16;
    const size_t tim = (g_dtype == RIP_F64) ? 8 : 4, ti = (g_dtype == RIP_F64 || k_dtype == RIP_F64) ? 8 : 4;
    return (size_t)threads * (tim * D_DEPTH * gm + ti * O_DEPTH * gm + 2 * R_DEPTH * gm + 4 * S_DEPTH + 4 * F_DEPTH + F_DEPTH) + 64;
}

void launch_cal_fused(const CalArgs& A, int g_dtype, int k_dtype, int threads, cudaStream_t st) {
    RIP_REQUIRE(A.G >= 3 && A.G <= RIP_GMAX, "fused L1->L2: G=%d outside 3..%d", A.G, RIP_GMAX);
    RIP_REQUIRE(A.P >= 1 && A.P <= RIP_PMAX, "fused L1->L2: P=%d outside 1..%d", A.P, RIP_PMAX);
    RIP_REQUIRE(threads >= 32 && threads <= 256 && threads % 32 == 0, "fused L1->L2: threads=%d must be a multiple of 32 in 32..256", threads);
    RIP_REQUIRE(cal_fused_smem_bytes(A.G, g_dtype, k_dtype, threads) <= 227 * 1024, "fused L1->L2: tile does not fit shared memory; use fewer threads");
    const bool gd = g_dtype == RIP_F64, kd = k_dtype == RIP_F64;
    if (!gd && !kd) launch_cal_gp<float, float>(A, threads, st);
    else if (gd && !kd) launch_cal_gp<double, float>(A, threads, st);
    else if (!gd && kd) launch_cal_gp<float, double>(A, threads, st);
    else launch_cal_gp<double, double>(A, threads, st);
}

}  // namespace rip

using namespace rip;

extern "C" int rip_jump_detect(int device, const float* data, uint8_t* rdq, int ny, int nx, int nb, const rip_ramp_plan* plan,
                               const double* w_exact, int variant, const void* gain, int g_dtype, const float* read,
                               float* slope, float* err_read, float* err_poisson, float* smap) {
    RIP_API_BEGIN
    use_device(device);
    const double* dw = plan_to_device(device, plan, w_exact, 0);
    RIP_REQUIRE(variant >= 0 && variant < plan->nvar, "rip_jump_detect: variant %d outside plan (nvar=%d)", variant, plan->nvar);
    const int G = plan->G;
    const long npl = (long)ny * nx;
    const int nsl = plan->var_slice_off[variant + 1] - plan->var_slice_off[variant];
    DevBuf<float> d, rd, s(npl), er(npl), ep(npl), sm;
    DevBuf<uint8_t> q;
    DevRaw g;
    d.upload(data, (size_t)G * npl);
    rd.upload(read, npl);
    q.upload(rdq, (size_t)G * npl);
    g.upload(gain, npl * dtype_size(g_dtype));
    if (smap && nsl > 0) sm.alloc((size_t)nsl * npl);
    dim3 grid((nx + 127) / 128, ny);
    if (g_dtype == RIP_F64)
        RIP_LAUNCH(jump_detect_kernel<double>, grid, 128, 0, 0, d.p, q.p, ny, nx, nb, variant, (const double*)g.p, rd.p, dw, s.p, er.p, ep.p, sm.p);
    else
        RIP_LAUNCH(jump_detect_kernel<float>, grid, 128, 0, 0, d.p, q.p, ny, nx, nb, variant, (const float*)g.p, rd.p, dw, s.p, er.p, ep.p, sm.p);
    s.download(slope, npl);
    er.download(err_read, npl);
    ep.download(err_poisson, npl);
    q.download(rdq, (size_t)G * npl);
    if (smap && nsl > 0) sm.download(smap, (size_t)nsl * npl);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_ramp_fit(int device, const float* data, uint8_t* rdq, uint32_t* pdq, int ny, int nx, int nb,
                            const rip_ramp_plan* plan, const double* w_exact, const void* gain, int g_dtype,
                            const float* read, int fast, float* slope, float* err_read, float* err_poisson) {
    RIP_API_BEGIN
    use_device(device);
    const double* dw = plan_to_device(device, plan, w_exact, 0);
    const int G = plan->G;
    const long npl = (long)ny * nx;
    DevBuf<float> d, rd, s(npl), er(npl), ep(npl);
    DevBuf<uint8_t> q;
    DevBuf<uint32_t> pd;
    DevRaw g;
    d.upload(data, (size_t)G * npl);
    rd.upload(read, npl);
    q.upload(rdq, (size_t)G * npl);
    pd.upload(pdq, npl);
    g.upload(gain, npl * dtype_size(g_dtype));
    dim3 grid((nx + 127) / 128, ny);
#define RF(T, F) RIP_LAUNCH((ramp_fit_kernel<T, F>), grid, 128, 0, 0, d.p, q.p, pd.p, ny, nx, nb, (const T*)g.p, rd.p, dw, s.p, er.p, ep.p)
    if (g_dtype == RIP_F64) { if (fast) RF(double, true); else RF(double, false); }
    else { if (fast) RF(float, true); else RF(float, false); }
#undef RF
    s.download(slope, npl);
    er.download(err_read, npl);
    ep.download(err_poisson, npl);
    q.download(rdq, (size_t)G * npl);
    pd.download(pdq, npl);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}
