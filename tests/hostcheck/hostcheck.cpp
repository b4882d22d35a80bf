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
#include "rip_v2_core.cuh"

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


// ---- v2 (rip_v2_core.cuh): host pack with the shared rec*_word functions + the kernel's march schedule ---------
template <int G, int P, bool K64 = false>
static void run_v2_t(rip::v2::Args A, const rip::v2::PackSrc& S, const rip_ramp_plan& pl) {
    using namespace rip::v2;
    constexpr int KQ = K64 ? rip::v2::KQ64 : rip::v2::KQ;
    const int n = A.n, ntile = ntiles(n), nq = nq1(G, P);
    // records with PADR zero rows on both sides, as v2_pack lays them out (rip_v2.cu)
    const int nrow = n + 2 * PADR;
    std::vector<f4> rec1((size_t)nrow * ntile * nq * TW), recK((size_t)nrow * ntile * KQ * TW);
    for (int prow = 0; prow < nrow; ++prow)
        for (int tile = 0; tile < ntile; ++tile)
            for (int c = 0; c < TW; ++c) {
                const int x = tile * TS + c, row = prow - PADR;
                for (int q = 0; q < nq; ++q)
                    rec1[((size_t)(prow * ntile + tile) * nq + q) * TW + c] =
                        f4{rec1_word(S, row, x, 4 * q), rec1_word(S, row, x, 4 * q + 1), rec1_word(S, row, x, 4 * q + 2), rec1_word(S, row, x, 4 * q + 3)};
                for (int q = 0; q < KQ; ++q)
                    recK[((size_t)(prow * ntile + tile) * KQ + q) * TW + c] =
                        K64 ? f4{recK64_word(S, row, x, 4 * q), recK64_word(S, row, x, 4 * q + 1), recK64_word(S, row, x, 4 * q + 2), recK64_word(S, row, x, 4 * q + 3)}
                            : f4{recK_word(S, row, x, 4 * q), recK_word(S, row, x, 4 * q + 1), recK_word(S, row, x, 4 * q + 2), recK_word(S, row, x, 4 * q + 3)};
            }
    A.ntile = ntile;
    A.rec1 = rec1.data() + (size_t)PADR * ntile * nq * TW;
    A.recK = recK.data() + (size_t)PADR * ntile * KQ * TW;
    // the tabulated channel lines K0 (k0_chan_kernel) hands to the kernel: m * row + c in f64, unfused
    std::vector<double> line;
    if (A.do_refpix) {
        line.resize((size_t)G * 32 * n);
        for (int g = 0; g < G; ++g)
            for (int ch = 0; ch < 32; ++ch)
                for (int row = 0; row < n; ++row) {
                    const double prod = A.chan_m[g * 32 + ch] * (double)row;
                    line[((size_t)g * 32 + ch) * n + row] = prod + A.chan_c[g * 32 + ch];
                }
        A.chan_line = line.data();
    }
    const FastTab ft = make_fast_tab(pl);
    const size_t smem = Smem<G, K64>::bytes();
    std::vector<unsigned char> buf(smem + 64);
    std::vector<Regs<G, P>> regs(TW);
    const int gy = (n + A.band_rows - 1) / A.band_rows;
    for (int by = 0; by < gy; ++by)
        for (int tile = 0; tile < ntile; ++tile) {
            memset(buf.data(), 0xA5, buf.size());
            memset((void*)regs.data(), 0xA5, sizeof(Regs<G, P>) * TW);
            unsigned char* base = buf.data();
            base += (16 - ((size_t)base & 15)) & 15;
            Smem<G, K64> sm;
            sm.carve(base);
            const int r0 = by * A.band_rows, r1 = (r0 + A.band_rows < n) ? r0 + A.band_rows : n;
            for (int tid = 0; tid < TW; ++tid) prologue<G, P, K64>(A, sm, regs[tid], tid, tile, r0, r1);
            unsigned o5s = first_o5<G>(r0);
            for (int s = r0 - 3; s <= r1 + 5; ++s) {
                for (int tid = 0; tid < TW; ++tid) step<G, P, K64>(A, pl, ft, sm, regs[tid], tid, tile, r0, r1, s, o5s);
                o5s = next_o5<G>(o5s);
            }
        }
}

// ---- v6: depth-4 record ring, two barriers per step (all threads walk the first half of a step, then the second) --------
template <int G, int P, bool K64 = false>
static void run_v6_t(rip::v2::Args A, const rip::v2::PackSrc& S, const rip_ramp_plan& pl) {
    using namespace rip::v2;
    constexpr int KQ = K64 ? rip::v2::KQ64 : rip::v2::KQ;
    const int n = A.n, ntile = ntiles(n), nq = nq1(G, P);
    const int nrow = n + 2 * PADR;
    std::vector<f4> rec1((size_t)nrow * ntile * nq * TW), recK((size_t)nrow * ntile * KQ * TW);
    for (int prow = 0; prow < nrow; ++prow)
        for (int tile = 0; tile < ntile; ++tile)
            for (int c = 0; c < TW; ++c) {
                const int x = tile * TS + c, row = prow - PADR;
                for (int q = 0; q < nq; ++q)
                    rec1[((size_t)(prow * ntile + tile) * nq + q) * TW + c] =
                        f4{rec1_word(S, row, x, 4 * q), rec1_word(S, row, x, 4 * q + 1), rec1_word(S, row, x, 4 * q + 2), rec1_word(S, row, x, 4 * q + 3)};
                for (int q = 0; q < KQ; ++q)
                    recK[((size_t)(prow * ntile + tile) * KQ + q) * TW + c] =
                        K64 ? f4{recK64_word(S, row, x, 4 * q), recK64_word(S, row, x, 4 * q + 1), recK64_word(S, row, x, 4 * q + 2), recK64_word(S, row, x, 4 * q + 3)}
                            : f4{recK_word(S, row, x, 4 * q), recK_word(S, row, x, 4 * q + 1), recK_word(S, row, x, 4 * q + 2), recK_word(S, row, x, 4 * q + 3)};
            }
    A.ntile = ntile;
    A.rec1 = rec1.data() + (size_t)PADR * ntile * nq * TW;
    A.recK = recK.data() + (size_t)PADR * ntile * KQ * TW;
    std::vector<double> line;
    if (A.do_refpix && A.chan_m && A.chan_c) {
        line.resize((size_t)G * 32 * n);
        for (int g = 0; g < G; ++g)
            for (int ch = 0; ch < 32; ++ch)
                for (int j = 0; j < n; ++j) line[((size_t)g * 32 + ch) * n + j] = A.chan_m[g * 32 + ch] * (double)j + A.chan_c[g * 32 + ch];
        A.chan_line = line.data();
    }
    const FastTab ft = make_fast_tab(pl);
    const size_t smem = Smem6<G, K64>::bytes();
    std::vector<unsigned char> buf(smem + 64);
    std::vector<Regs<G, P>> regs(TW);
    const int gy = (n + A.band_rows - 1) / A.band_rows;
    for (int by = 0; by < gy; ++by)
        for (int tile = 0; tile < ntile; ++tile) {
            memset(buf.data(), 0xA5, buf.size());
            memset((void*)regs.data(), 0xA5, sizeof(Regs<G, P>) * TW);
            unsigned char* base = buf.data();
            base += (16 - ((size_t)base & 15)) & 15;
            Smem6<G, K64> sm;
            sm.carve(base);
            const int r0 = by * A.band_rows, r1 = (r0 + A.band_rows < n) ? r0 + A.band_rows : n;
            for (int tid = 0; tid < TW; ++tid) prologue6<G, P, K64>(A, sm, regs[tid], tid, tile, r0, r1);
            for (int s = r0 - 3; s <= r1 + 4; ++s) {
                for (int tid = 0; tid < TW; ++tid) step6a<G, P, 2, K64>(A, sm, regs[tid], tid, tile, r0, r1, s);
                for (int tid = 0; tid < TW; ++tid) step6b<G, P, 2, K64>(A, pl, ft, sm, regs[tid], tid, tile, r0, r1, s);
            }
        }
}

extern "C" int hostcheck_cal_fused_v6(const rip::v2::Args* A, const rip::v2::PackSrc* S, const rip_ramp_plan* plan) {
    const int G = S->G, P = S->P <= 4 ? 4 : (S->P <= 11 ? 11 : S->P);
    if (G == 8 && P == 4) run_v6_t<8, 4>(*A, *S, *plan);
    else if (G == 8 && P == 11) run_v6_t<8, 11>(*A, *S, *plan);
    else return 1;
    return 0;
}

extern "C" int hostcheck_cal_fused_v6k64(const rip::v2::Args* A, const rip::v2::PackSrc* S, const rip_ramp_plan* plan) {
    const int G = S->G, P = S->P <= 4 ? 4 : (S->P <= 11 ? 11 : S->P);
    if (G == 8 && P == 4) run_v6_t<8, 4, true>(*A, *S, *plan);
    else if (G == 8 && P == 11) run_v6_t<8, 11, true>(*A, *S, *plan);
    else return 1;
    return 0;
}

// S->ipc points at float64 taps (the v2 kernel for float64 ipc4d: G = 8 only)
extern "C" int hostcheck_cal_fused_v2k64(const rip::v2::Args* A, const rip::v2::PackSrc* S, const rip_ramp_plan* plan) {
    const int G = S->G, P = S->P <= 4 ? 4 : (S->P <= 11 ? 11 : S->P);  // (v2_pad_P: records padded with zero coefficients)
    if (G == 8 && P == 4) run_v2_t<8, 4, true>(*A, *S, *plan);
    else if (G == 8 && P == 11) run_v2_t<8, 11, true>(*A, *S, *plan);
    else return 1;
    return 0;
}

extern "C" int hostcheck_cal_fused_v2(const rip::v2::Args* A, const rip::v2::PackSrc* S, const rip_ramp_plan* plan) {
    const int G = S->G, P = S->P <= 4 ? 4 : (S->P <= 11 ? 11 : S->P);
    if (G == 8 && P == 4) run_v2_t<8, 4>(*A, *S, *plan);
    else if (G == 8 && P == 11) run_v2_t<8, 11>(*A, *S, *plan);
    else if (G == 16 && P == 11) run_v2_t<16, 11>(*A, *S, *plan);
    else if (G == 16 && P == 4) run_v2_t<16, 4>(*A, *S, *plan);
    else return 1;
    return 0;
}
extern "C" int hostcheck_sizeof_v2args(void) { return (int)sizeof(rip::v2::Args); }
extern "C" int hostcheck_sizeof_packsrc(void) { return (int)sizeof(rip::v2::PackSrc); }

// Shared-reciprocal division (rip::v2::SharedDiv) against true IEEE division, with the hardware reciprocal emulated
// as a 1-ulp-perturbed 1/d.  Returns the number of mismatches over `count` pseudo-random (x, d) pairs.
extern "C" long hostcheck_shared_div(long count, unsigned seed, float dlo, float dhi, float xmax) {
    unsigned long long st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(st >> 11) / 9007199254740992.0; };
    long bad = 0;
    for (long i = 0; i < count; ++i) {
        const float d = (float)(dlo * pow((double)dhi / dlo, rnd())) * (rnd() < 0.5 ? 1.f : -1.f);
        const float x = (float)((2.0 * rnd() - 1.0) * xmax * pow(2.0, -20.0 * rnd()));
        rip::v2::SharedDiv sd;
        sd.d = d;
        sd.ok = true;
        float r0 = 1.0f / d;
        const int pert = (int)(rnd() * 3.0) - 1;  // -1, 0, +1 ulp: MUFU.RCP is not correctly rounded
        r0 = nextafterf(r0, pert > 0 ? INFINITY : (pert < 0 ? -INFINITY : r0));
        const float e = fmaf(-d, r0, 1.0f);
        sd.r = fmaf(r0, e, r0);
        const rip::v2::f2 q = sd.div2(rip::v2::f2{x, -x});
        if (q.x != x / d || q.y != (-x) / d) ++bad;
    }
    return bad;
}

// Certified fast inverse (rip::invlin_fast_z) against the plain 24-step search, on pseudo-random pixels with
// realistic and with hostile coefficient sets, for ramps of increasing signal (root carried from read to read) and
// for adversarial signals placed exactly on / one ulp off the float32 values the search compares against.
// Returns the number of mismatching z.  stats[0] = calls, [1] = exact evaluations, [2] = uncertified pixels.
template <int P>
static long invlin_check_t(long npix, unsigned seed, double hostile, double* stats) {
    unsigned long long st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(st >> 11) / 9007199254740992.0; };
    auto slow_z = [&](double Slin, const float (&c)[P]) {
        double z = 0.0, step = 1.0;
        bool ex;
        for (int j = 1; j < 25; ++j) {
            step *= 0.5;
            const float phi = legendre_eval<double, P, false>(z, c, P, ex);
            z = z + (((double)phi < Slin) ? step : -step);
        }
        return z;
    };
    long bad = 0;
    for (long px = 0; px < npix; ++px) {
        float c[P];
        double cd[P];
        const double c1 = 20000.0 + 12000.0 * rnd();
        c[1] = (float)c1;
        c[0] = (float)(c1 * (0.9 + 0.2 * rnd()) + 3000.0 * rnd());
        const double c2 = (rnd() < 0.5 ? -1.0 : 1.0) * (20.0 + 180.0 * rnd());
        if (P > 2) c[2] = (float)c2;
        for (int L = 3; L < P; ++L) c[L] = (float)((rnd() - 0.5) * 0.04 * c2 * (1.0 + hostile * 30.0 * rnd()));
        for (int L = 0; L < P; ++L) cd[L] = (double)c[L];
        float A, m;
        invlin_certify<P>(c, P, A, m);
        if (!(m > 0.0f)) { stats[2] += 1; continue; }
        bool ex;
        const double lo = legendre_eval<double, P, false>(-1.0, c, P, ex), hi = legendre_eval<double, P, false>(1.0, c, P, ex);
        double r = 0.0;
        auto check = [&](double Slin) {
            int ne = 0;
            const double zf = invlin_fast_z<P>(Slin, cd, A, m, r, &ne);
            const double zs = slow_z(Slin, c);
            stats[0] += 1; stats[1] += ne;
            if (!(zf == zs)) ++bad;
        };
        // a ramp of reads (faint or bright pixel)
        const double rate = (rnd() < 0.7 ? 40.0 : 1800.0) * rnd();
        double sig = lo + 200.0 + 3000.0 * rnd();
        for (int k = 0; k < 12; ++k) { check(sig); sig += rate * (1.0 + rnd()); }
        // adversarial: exactly the float32 value the search sees at a point of its grid, and its neighbours
        for (int k = 0; k < 10; ++k) {
            const int bits = 8 + (int)(rnd() * 17.0);  // grid of step 2^-bits
            const double zt = clamp_pm1((floor(rnd() * ldexp(2.0, bits)) - ldexp(1.0, bits)) * ldexp(1.0, -bits));
            const float pf = legendre_eval<double, P, false>(zt, c, P, ex);
            check((double)pf);
            check((double)nextafterf(pf, INFINITY));
            check((double)nextafterf(pf, -INFINITY));
            check(nextafter((double)pf, INFINITY));
            check(nextafter((double)pf, -INFINITY));
            check((double)pf + (rnd() - 0.5) * 0.02);
        }
        // edges of the range and beyond, non-finite signals
        check(lo); check(hi); check(lo - 1e-3); check(hi + 1e-3); check(lo - 5000.0); check(hi + 5000.0);
        check(nextafter(hi, INFINITY)); check(nextafter(lo, -INFINITY));
        check(0.0); check(NAN); check(INFINITY); check(-INFINITY);
        check(0.5 * (lo + hi));
    }
    return bad;
}

extern "C" long hostcheck_invlin_fast(long npix, unsigned seed, int P, double hostile, double* stats) {
    stats[0] = stats[1] = stats[2] = 0.0;
    if (P == 4) return invlin_check_t<4>(npix, seed, hostile, stats);
    if (P == 11) return invlin_check_t<11>(npix, seed, hostile, stats);
    if (P == 16) return invlin_check_t<16>(npix, seed, hostile, stats);
    return -1;
}
