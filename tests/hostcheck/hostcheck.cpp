// TEST INFRASTRUCTURE ONLY -- never linked into librip_b200.so and never imported by romanimpreprocess_b200/.
//
// The per-pixel sources of the fused L1->L2 kernel (csrc/rip_math.cuh, csrc/rip_cal_core.cuh) are written
// __host__ __device__.  This file compiles the *same* source with g++ and walks it through the CTA / march-step
// schedule of cal_fused_kernel (csrc/rip_fit.cu) sequentially, so that tile logic, ring-buffer depths, halo
// handling and DQ propagation can be checked against the oracle in the GPU-less build container before GPU time
// is spent.  The GPU parity tests (tests/test_gpu_*.py, -m gpu) remain the gate for the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rip_cal_core.cuh"

using namespace rip;

template <int GMAX, int PMAX, typename TG, typename TK>
static void run_t(const CalArgs& A, const rip_ramp_plan& pl, int threads) {
    typedef typename Promote<float, TG>::type TIM;
    typedef typename Promote<TIM, TK>::type TI;
    const size_t smem = CalSmem<GMAX, TIM, TI>::bytes(threads);
    std::vector<unsigned char> buf(smem + 64);
    const int tw = threads - 6;
    const int gx = (A.n + tw - 1) / tw, gy = (A.n + A.band_rows - 1) / A.band_rows;
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            // poison the "shared memory" so that reads of never-written ring slots show up as garbage
            memset(buf.data(), 0xA5, buf.size());
            CalSmem<GMAX, TIM, TI> sm;
            unsigned char* base = buf.data();
            base += (16 - ((size_t)base & 15)) & 15;
            sm.carve(base, threads);
            const int c0 = bx * tw, r0 = by * A.band_rows;
            const int r1 = (r0 + A.band_rows < A.n) ? r0 + A.band_rows : A.n;
            for (int s = r0 - 3; s <= r1 + 5; ++s)
                for (int tid = 0; tid < threads; ++tid) cal_step<GMAX, PMAX, TG, TK>(A, pl, sm, tid, threads, c0, r0, r1, s);
        }
}

template <typename TG, typename TK>
static void run_gp(const CalArgs& A, const rip_ramp_plan& pl, int threads) {
    if (A.G <= 8) {
        if (A.P <= 4) run_t<8, 4, TG, TK>(A, pl, threads);
        else if (A.P <= 11) run_t<8, 11, TG, TK>(A, pl, threads);
        else run_t<8, RIP_PMAX, TG, TK>(A, pl, threads);
    } else {
        if (A.P <= 4) run_t<16, 4, TG, TK>(A, pl, threads);
        else if (A.P <= 11) run_t<16, 11, TG, TK>(A, pl, threads);
        else run_t<16, RIP_PMAX, TG, TK>(A, pl, threads);
    }
}

extern "C" int hostcheck_cal_fused(const CalArgs* A, const rip_ramp_plan* plan, int g_dtype, int k_dtype, int threads) {
    const bool gd = g_dtype == RIP_F64, kd = k_dtype == RIP_F64;
    if (!gd && !kd) run_gp<float, float>(*A, *plan, threads);
    else if (gd && !kd) run_gp<double, float>(*A, *plan, threads);
    else if (!gd && kd) run_gp<float, double>(*A, *plan, threads);
    else run_gp<double, double>(*A, *plan, threads);
    return 0;
}

extern "C" int hostcheck_sizeof_calargs(void) { return (int)sizeof(CalArgs); }
