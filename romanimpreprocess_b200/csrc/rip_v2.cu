// cal_fused v2 (see rip_v2_core.cuh): record packing, kernel, launcher.
#include "rip_handle.h"
#include "rip_launch.h"
#include "rip_v2_core.cuh"

#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

namespace rip {

__constant__ RampPlanDev c_plan_v2;
__constant__ v2::FastTab c_fast_v2;

namespace v2 {

__global__ void pack_rec1_kernel(const PackSrc S, int ntile, int nq, f4* __restrict__ out) {
    const int c = threadIdx.x, tile = blockIdx.x, row = (int)blockIdx.y - PADR;  // out points at row 0
    const int x = tile * TS + c;
    f4* o = out + ((long)row * ntile + tile) * ((long)nq * TW) + c;
    for (int q = 0; q < nq; ++q) {
        f4 v;
        v.x = rec1_word(S, row, x, 4 * q);
        v.y = rec1_word(S, row, x, 4 * q + 1);
        v.z = rec1_word(S, row, x, 4 * q + 2);
        v.w = rec1_word(S, row, x, 4 * q + 3);
        o[q * TW] = v;
    }
}

__global__ void pack_recK_kernel(const PackSrc S, int ntile, f4* __restrict__ out) {
    const int c = threadIdx.x, tile = blockIdx.x, row = (int)blockIdx.y - PADR;
    const int x = tile * TS + c;
    f4* o = out + ((long)row * ntile + tile) * ((long)KQ * TW) + c;
    for (int q = 0; q < KQ; ++q) {
        f4 v;
        v.x = recK_word(S, row, x, 4 * q);
        v.y = recK_word(S, row, x, 4 * q + 1);
        v.z = recK_word(S, row, x, 4 * q + 2);
        v.w = recK_word(S, row, x, 4 * q + 3);
        o[q * TW] = v;
    }
}

__global__ void pack_recK64_kernel(const PackSrc S, int ntile, f4* __restrict__ out) {
    const int c = threadIdx.x, tile = blockIdx.x, row = (int)blockIdx.y - PADR;
    const int x = tile * TS + c;
    f4* o = out + ((long)row * ntile + tile) * ((long)KQ64 * TW) + c;
    for (int q = 0; q < KQ64; ++q) {
        f4 v;
        v.x = recK64_word(S, row, x, 4 * q);
        v.y = recK64_word(S, row, x, 4 * q + 1);
        v.z = recK64_word(S, row, x, 4 * q + 2);
        v.w = recK64_word(S, row, x, 4 * q + 3);
        o[q * TW] = v;
    }
}

// One CTA per (column tile, row band).  (A persistent-CTA variant pulling work items from an atomic counter was
// measured on B200: no gain -- 1120 items on 592 resident slots already overlap well -- and its outer loop cost
// registers, i.e. spills at the 128-register budget; dropped.)
template <int G, int P, int MINB>
__global__ void __launch_bounds__(TW, MINB) cal_fused_v2_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<G> sm;
    sm.carve(smem_raw);
    Regs<G, P> R;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = min(r0 + A.band_rows, A.n);
    prologue<G, P>(A, sm, R, tid, tile, r0, r1);
    __syncthreads();
    unsigned o5s = first_o5<G>(r0);
    for (int s = r0 - 3; s <= r1 + 5; ++s) {
        step<G, P>(A, c_plan_v2, c_fast_v2, sm, R, tid, tile, r0, r1, s, o5s);
        o5s = next_o5<G>(o5s);
        __syncthreads();
    }
}

// v6 (rip_v2_core.cuh, "v6"): five resident CTAs per SM -- 96 registers, depth-4 record ring, stage c one row behind
// stage b with a barrier in between
template <int G, int P, int SCHED = 0>
__global__ void __launch_bounds__(TW, 5) cal_fused_v6_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem6<G> sm;
    sm.carve(smem_raw);
    Regs<G, P> R;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = min(r0 + A.band_rows, A.n);
    prologue6<G, P>(A, sm, R, tid, tile, r0, r1);
    __syncthreads();
    for (int s = r0 - 3; s <= r1 + 4; ++s) {
        step6a<G, P, SCHED>(A, sm, R, tid, tile, r0, r1, s);
        __syncthreads();
        step6b<G, P, SCHED>(A, c_plan_v2, c_fast_v2, sm, R, tid, tile, r0, r1, s);
        __syncthreads();
    }
}

// v6 for a float64 ipc4d: the same schedule with the IPC stages in float64 (O1 ring of doubles: 3 x 8320 B): 56.0 KB of rings
// -> FOUR resident CTAs per SM (4 x (57344 + 1024) B = the SM's 228 KB exactly) where the v2 organisation allows three
template <int G, int P>
__global__ void __launch_bounds__(TW, 4) cal_fused_v6k64_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem6<G, true> sm;
    sm.carve(smem_raw);
    Regs<G, P> R;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = min(r0 + A.band_rows, A.n);
    prologue6<G, P, true>(A, sm, R, tid, tile, r0, r1);
    __syncthreads();
    for (int s = r0 - 3; s <= r1 + 4; ++s) {
        step6a<G, P, 2, true>(A, sm, R, tid, tile, r0, r1, s);
        __syncthreads();
        step6b<G, P, 2, true>(A, c_plan_v2, c_fast_v2, sm, R, tid, tile, r0, r1, s);
        __syncthreads();
    }
}

// v6 with split-phase barriers (rip_v2_core.cuh, SCHED 3): the two CTA barriers of a v6 step become mbarrier
// arrive / wait pairs with independent work in between (stage a0 behind the mid-step arrival, the ramp fit and epilogue of
// stage c behind the end-of-step arrival), so a warp rarely blocks on its slowest sibling.
//     wait-end(prev)  a1  [Lc]  b  arrive-mid  a0  wait-mid  c: stencil, arrive-end, ramp fit ...  [L1 Lb]
template <int G, int P>
__global__ void __launch_bounds__(TW, 5) cal_fused_v6s_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = Smem6<G>;
    SM sm;
    sm.carve(smem_raw);
    uint64_t* mbar = (uint64_t*)(smem_raw + (size_t)SM::DEPTH * SM::ROW5 + (size_t)3 * SM::ROWO);  // the 64 spare bytes
    Regs<G, P> R;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = min(r0 + A.band_rows, A.n);
    if (tid == 0) {
        tma::mbar_init(mbar, TW);
        tma::mbar_init(mbar + 1, TW);
        tma::fence_mbar_init();
    }
    prologue6<G, P>(A, sm, R, tid, tile, r0, r1);
    __syncthreads();
    unsigned it = 0;
    for (int s = r0 - 3; s <= r1 + 4; ++s, ++it) {
        StepCtx C;
        make_ctx6<G>(C, A, tid, tile, r0, r1, s);
        const unsigned (&o5)[5] = C.o5;
        if (it) sp_wait(mbar + 1, (it - 1u) & 1u);  // everything the previous step published (and released) is visible
        row_async<G, P>(A, sm, R, s + 1, 1, RIP_OS(1), tile, tid, r0 - 3, r1 + 3);
        prefetch_records<G, P, KQ>(A, s, tile, tid, r0, r1);
        prefetch_raw<G>(A, s + 2, tile, tid, r0 - 3, r1 + 3);
        stage_a1<G, P>(A, sm, R, C);
        load_c<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
        stage_b<G, P>(A, sm, R, C);
        sp_arrive(mbar);       // O1 of row s-4 and D of row s-2 are written
        stage_a0<G, P>(A, sm, C);
        sp_wait(mbar, it & 1u);
        stage_c<G, P>(A, c_plan_v2, c_fast_v2, sm, R, C, NoHook(), ArriveHook{mbar + 1});
        load_a1<G, P>(A, R, fold_row(s - 1, r0 - 2, r1 + 2), tile, tid);
        load_b6<G, P>(A, R, fold_row(s - 3, r0 - 1, r1 + 1), tile, tid);
        R.orow += (unsigned)A.n;
    }
}

// float64 ipc4d (K64): same march, IPC stages in float64 (3 CTAs/SM: the O1 ring holds doubles)
template <int G, int P, int MINB>
__global__ void __launch_bounds__(TW, MINB) cal_fused_v2k64_kernel(const Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<G, true> sm;
    sm.carve(smem_raw);
    Regs<G, P> R;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = min(r0 + A.band_rows, A.n);
    prologue<G, P, true>(A, sm, R, tid, tile, r0, r1);
    __syncthreads();
    unsigned o5s = first_o5<G>(r0);
    for (int s = r0 - 3; s <= r1 + 5; ++s) {
        step<G, P, true>(A, c_plan_v2, c_fast_v2, sm, R, tid, tile, r0, r1, s, o5s);
        o5s = next_o5<G>(o5s);
        __syncthreads();
    }
}

static void configure_once(const void* fn, size_t smem) { configure_smem_once(fn, smem, true); }

template <int G, int P, int SCHED = 0>
static void launch_v6(const Args& A, cudaStream_t st) {
    const size_t smem = Smem6<G>::bytes();
    auto kern = cal_fused_v6_kernel<G, P, SCHED>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, TW, smem, st, A);
}

template <int G, int P>
static void launch_v6k64(const Args& A, cudaStream_t st) {
    const size_t smem = Smem6<G, true>::bytes();
    auto kern = cal_fused_v6k64_kernel<G, P>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, TW, smem, st, A);
}

template <int G, int P>
static void launch_v6s(const Args& A, cudaStream_t st) {
    const size_t smem = Smem6<G>::bytes();
    auto kern = cal_fused_v6s_kernel<G, P>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, TW, smem, st, A);
}

template <int G, int P>
static void launch_k64(const Args& A, cudaStream_t st) {
    const size_t smem = Smem<G, true>::bytes();
    auto kern = cal_fused_v2k64_kernel<G, P, 3>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, TW, smem, st, A);
}

template <int G, int P, int MINB>
static void launch_tb(const Args& A, cudaStream_t st) {
    const size_t smem = Smem<G>::bytes();
    auto kern = cal_fused_v2_kernel<G, P, MINB>;
    configure_once((const void*)kern, smem);
    dim3 grid(A.ntile, (A.n + A.band_rows - 1) / A.band_rows);
    RIP_LAUNCH(kern, grid, TW, smem, st, A);
}

// resident CTAs per SM the kernel is compiled for: 4 (128 registers/thread) for G <= 8, 3 for G = 16 (larger rings);
template <int G, int P>
static void launch_t(const Args& A, cudaStream_t st) {
    launch_tb<G, P, (G <= 8) ? 4 : 3>(A, st);
}

}  // namespace v2

// Band height: about 64 rows (measured best on B200: taller bands lose to the wave tail, shorter ones to the 9 halo
// steps per band), adjusted so that the CTA count ends just under a whole number of waves of resident CTAs
// (4096^2, 148 SMs x 4: 67 bands of 62 rows = 2345 CTAs = 3.96 waves instead of 64 x 35 = 3.78; profiles/r02).
int v2_default_band_rows(int device, int n, int G, int ctas_per_sm) {
    static thread_local int sms_dev = -1, sms = 0;
    if (sms_dev != device) {
        RIP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        sms_dev = device;
    }
    const int slots = sms * (ctas_per_sm > 0 ? ctas_per_sm : ((G <= 8) ? 4 : 3)), ntile = v2::ntiles(n);
    const int nb0 = (n + 63) / 64;
    const int waves = (ntile * nb0 + slots - 1) / slots;
    int nb = waves * slots / ntile;
    if (nb < nb0) nb = nb0;
    if (nb > 2 * nb0) nb = 2 * nb0;
    const int rows = (n + nb - 1) / nb;
    return rows < 16 ? 16 : rows;
}

bool v2_supported(int G, int P, bool k64) {
    P = v2_pad_P(P);
    if (k64) return G == 8 && (P == 4 || P == 11);  // (G = 16 with float64 taps would leave one CTA per SM: generic kernel)
    return (G == 8 && P == 4) || (G == 8 && P == 11) || (G == 16 && P == 11) || (G == 16 && P == 4);
}

bool v6_supported(int G, int P) { return G == 8 && (v2_pad_P(P) == 4 || v2_pad_P(P) == 11); }
void launch_cal_fused_v6(const v2::Args& A, int G, int P, cudaStream_t st) {
    P = v2_pad_P(P);
    if (G == 8 && P == 4) v2::launch_v6<8, 4, 2>(A, st);
    else if (G == 8 && P == 11) {
        // measured (profiles/r02/ab_v6_sched.log): record of stage c issued before stage b (SCHED 2) 1.3245 ms, after it
        // 1.328 ms, stage a0 in the first half 1.342 ms
        static const int split = [] { const char* e = getenv("RIP_V6_SPLIT"); return e ? atoi(e) : 0; }();  // (development A/B)
        if (split) v2::launch_v6s<8, 11>(A, st);
        else v2::launch_v6<8, 11, 2>(A, st);
    }
    else throw Error("cal_fused v6: unsupported (G, P)");
}

void launch_cal_fused_v6k64(const v2::Args& A, int G, int P, cudaStream_t st) {
    P = v2_pad_P(P);
    if (G == 8 && P == 4) v2::launch_v6k64<8, 4>(A, st);
    else if (G == 8 && P == 11) v2::launch_v6k64<8, 11>(A, st);
    else throw Error("cal_fused v6 (float64 ipc4d): unsupported (G, P)");
}

void launch_cal_fused_v2k64(const v2::Args& A, int G, int P, cudaStream_t st) {
    P = v2_pad_P(P);
    if (G == 8 && P == 4) v2::launch_k64<8, 4>(A, st);
    else if (G == 8 && P == 11) v2::launch_k64<8, 11>(A, st);
    else throw Error("cal_fused v2 (float64 ipc4d): unsupported (G, P)");
}

void launch_cal_fused_v2(const v2::Args& A, int G, int P, cudaStream_t st) {
    P = v2_pad_P(P);
    if (G == 16 && P == 4) v2::launch_t<16, 4>(A, st);
    else if (G == 8 && P == 4) v2::launch_t<8, 4>(A, st);
    else if (G == 8 && P == 11) {
        static const int minb = [] { const char* e = getenv("RIP_FUSED_MINB"); return e ? atoi(e) : 4; }();  // (development: register budget A/B)
        if (minb == 5) v2::launch_tb<8, 11, 5>(A, st);
        else v2::launch_t<8, 11>(A, st);
    }
    else if (G == 16 && P == 11) v2::launch_t<16, 11>(A, st);
    else throw Error("cal_fused v2: unsupported (G, P)");
}

void v2_plan_to_device(const rip_ramp_plan* plan, cudaStream_t st) {
    RIP_CUDA(cudaMemcpyToSymbolAsync(c_plan_v2, plan, sizeof(rip_ramp_plan), 0, cudaMemcpyHostToDevice, st));
    // (the copy is taken from pageable memory at enqueue time, so the temporary may die here)
    const v2::FastTab ft = v2::make_fast_tab(*plan);
    RIP_CUDA(cudaMemcpyToSymbolAsync(c_fast_v2, &ft, sizeof ft, 0, cudaMemcpyHostToDevice, st));
    v3_plan_to_device(plan, &ft, st);  // (the role-split variants live in their own translation unit: rip_v3.cu)
}

// record of detector row 0 inside the padded allocations
v2::f4* v2_rec1_row0(rip_caldir* h, int G) {
    return (v2::f4*)h->v2_rec1.p + (size_t)v2::PADR * v2::ntiles(h->n) * v2::nq1(G, v2_pad_P(h->P)) * v2::TW;
}
v2::f4* v2_recK_row0(rip_caldir* h) {
    return (v2::f4*)h->v2_recK.p + (size_t)v2::PADR * v2::ntiles(h->n) * v2::kq_of(h->d.ipc_dtype == RIP_F64) * v2::TW;
}

// (re)build the packed records of a handle for G groups
void v2_pack(rip_caldir* h, int G, cudaStream_t st) {
    if (h->v2_G == G) return;
    const bool k64 = h->d.ipc_dtype == RIP_F64;
    const int kq = v2::kq_of(k64);
    const int n = h->n, ntile = v2::ntiles(n), nq = v2::nq1(G, v2_pad_P(h->P));
    const int nrow = n + 2 * v2::PADR;  // zero rows on both sides: the kernel's loaders never clamp
    h->v2_rec1.alloc((size_t)nrow * ntile * nq * v2::TW * 4);
    h->v2_recK.alloc((size_t)nrow * ntile * kq * v2::TW * 4);
    v2::PackSrc S;
    S.n = n; S.nb = h->nb; S.G = G; S.P = h->P;
    S.dark = h->dark_cube.p;
    S.bias = h->has_bias ? h->biascorr.p + (size_t)(h->d.n_bias - G) * h->na * h->na : nullptr;
    S.coefs = h->coefs.p; S.Smin = h->Smin.p; S.Smax = h->Smax.p; S.Sref = h->Sref.p;
    S.gain = (const float*)h->gain.p; S.aux = h->aux.p; S.ipc = (const float*)h->ipc.p;
    S.read = h->read.p; S.dslope = h->dslope_ipc.p; S.flat = h->flat_ipc.p; S.sdq = h->sdq.p;
    dim3 grid(ntile, nrow);
    RIP_LAUNCH(v2::pack_rec1_kernel, grid, v2::TW, 0, st, S, ntile, nq, v2_rec1_row0(h, G));
    if (k64) RIP_LAUNCH(v2::pack_recK64_kernel, grid, v2::TW, 0, st, S, ntile, v2_recK_row0(h));
    else RIP_LAUNCH(v2::pack_recK_kernel, grid, v2::TW, 0, st, S, ntile, v2_recK_row0(h));
    h->v2_G = G;
}

}  // namespace rip
