// Per-pixel arithmetic of the L1->L2 path, written once for device and host.
//
// Everything here is __host__ __device__ so that tests/hostcheck can run the *same* source on the CPU (this
// container has no GPU); the product never calls the host instantiation.  All translation units are compiled
// with -fmad=false (nvcc) / -ffp-contract=off (g++): NumPy never fuses multiply-add and the parity contract
//  is "same roundings as the reference", so every * and + below is a separate IEEE operation.
#pragma once
#include <math.h>
#include <stdint.h>

#include "rip_b200.h"

#ifdef __CUDACC__
#define RIP_HD __host__ __device__ __forceinline__
#define RIP_HD_COLD inline __host__ __device__ __noinline__   // rarely executed paths: keep them out of the hot instruction stream
#else
#define RIP_HD inline
#define RIP_HD_COLD inline
#endif

namespace rip {

// DQ bits (roman_datamodels.dqflags.pixel)
constexpr uint32_t DQ_DO_NOT_USE = 1u;
constexpr uint32_t DQ_SATURATED = 2u;
constexpr uint32_t DQ_JUMP_DET = 4u;
constexpr uint32_t DQ_AD_FLOOR = 64u;
constexpr uint32_t DQ_GW_AFFECTED_DATA = 16u;
constexpr uint32_t DQ_NO_FLAT_FIELD = 1u << 18;
constexpr uint32_t DQ_NO_GAIN_VALUE = 1u << 19;
constexpr uint32_t DQ_NO_LIN_CORR = 1u << 20;
constexpr uint32_t DQ_NO_SAT_CHECK = 1u << 21;
constexpr uint32_t DQ_REFERENCE_PIXEL = 1u << 31;


RIP_HD float rip_sqrt(float x) { return sqrtf(x); }
RIP_HD double rip_sqrt(double x) { return sqrt(x); }

// ---- NumPy-flavoured min/max/clip: NaN propagates (np.clip / np.maximum semantics) ----
template <typename T> RIP_HD T np_max(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
template <typename T> RIP_HD T np_min(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
template <typename T> RIP_HD T np_clip(T x, T lo, T hi) { return np_min(np_max(x, lo), hi); }

// ---------------------------------------------------------------------------------------------------------
// Legendre evaluation  (reference utils/ipc_linearity.py:215-231)
//   phi = c0 + sum_{L>=1} c_L * (|z|>1 ? sign(z)^L (1 + L(L+1)/2 (|z|-1)) : P_L(z)),
//   P_{L+1} = ((2L+1)/(L+1) z) P_L - (L/(L+1)) P_{L-1}, rational constants rounded to T first.
// ---------------------------------------------------------------------------------------------------------
//   NB (observed in the reference, pinned by the golden vectors): the accumulator `phi` is created as a copy of
//   coefs[0], i.e. it has the dtype of the coefficient planes (float32) even when z is float64; each
//   `phi += c_L * term` is evaluated in the promoted type TZ and rounded back to float32.
template <typename TZ, int PMAX, bool LINEXTRAP>
RIP_HD float legendre_eval(TZ z, const float (&c)[PMAX], int P, bool& ex) {
    TZ az = z < 0 ? -z : z;  // fabs; NaN stays NaN
    ex = az > (TZ)1;
    float phi = c[0];
    TZ prev = (TZ)1;
    TZ cur = z;
#pragma unroll
    for (int L = 1; L < PMAX; ++L) {
        if (L < P) {
            TZ term = cur;
            if (LINEXTRAP && ex) {
                TZ s = (z > 0) ? (TZ)1 : ((L & 1) ? (TZ)-1 : (TZ)1);  // sign(z)^L, z != 0 here
                TZ e = (TZ)(L * (L + 1) / 2.0);
                term = s * ((TZ)1 + e * (az - (TZ)1));
            }
            phi = (float)((TZ)phi + (TZ)c[L] * term);
            TZ a = (TZ)((2 * L + 1) / (double)(L + 1));
            TZ b = (TZ)(L / (double)(L + 1));
            TZ nxt = (a * z) * cur - b * prev;
            prev = cur;
            cur = nxt;
        }
    }
    return phi;
}

// z = -1 + 2 (S - Smin) / (Smax - Smin)   (ipc_linearity.py:330)
template <typename T> RIP_HD T lin_z(T S, T Smin, T Smax) { return (T)-1 + ((T)2 * (S - Smin)) / (Smax - Smin); }

// One pixel through multilin (ipc_linearity.py:329-342): G groups in, G linearised groups out, dq updated.
//   attempt_mask bit g = "try to flag group g" (the reference passes ~rdq & SATURATED, gen_cal_image.py:585)
template <int GMAX, int PMAX>
RIP_HD void multilin_pixel(const float (&S)[GMAX], int G, const float (&c)[PMAX], int P, float Smin, float Smax,
                           float Sref, uint32_t& dq, uint32_t attempt_mask, bool do_not_flag_first,
                           float (&phi)[GMAX]) {
#pragma unroll
    for (int j = 0; j < GMAX; ++j) {
        if (j < G) {
            float z = lin_z<float>(S[j], Smin, Smax);
            const bool first = (j == 0) && do_not_flag_first;
            if (first) z = np_clip<float>(z, -1.0f, 1.0f);
            bool ex;
            float p = legendre_eval<float, PMAX, true>(z, c, P, ex);
            phi[j] = ((dq & (DQ_NO_LIN_CORR | DQ_REFERENCE_PIXEL)) == 0) ? p : (S[j] - Sref);
            if (!first && ex && ((attempt_mask >> j) & 1u)) dq |= DQ_NO_LIN_CORR;
        }
    }
}

// 24-step bisection inverse (ipc_linearity.py:381-390) with z in type TZ (the reference runs it in float64
// inside IL.apply because counts are int32: SURVEY App. A10).  (Smax-Smin)/2 is an f32 expression in the
// reference (both planes are f32) and is promoted afterwards.
template <typename TZ, int PMAX>
RIP_HD TZ invlin_pixel(TZ Slin, const float (&c)[PMAX], int P, float Smin, float Smax, bool& ex) {
    TZ z = (TZ)0;
    TZ step = (TZ)1;
    ex = false;
    for (int j = 1; j < 25; ++j) {
        step = step * (TZ)0.5;  // exact powers of two
        float phi = legendre_eval<TZ, PMAX, false>(z, c, P, ex);
        z = z + (((TZ)phi < Slin) ? step : -step);
    }
    float half = (Smax - Smin) / 2.0f;
    return (TZ)Smin + (TZ)half * ((TZ)1 + z);
}

// ---------------------------------------------------------------------------------------------------------
// The same 24-step inverse, float64, WITHOUT evaluating the polynomial at every step (forward model hot loop).
//
// The reference's comparison at step j is  phi_fp(z_j) < Slin  with phi_fp the float32-accumulated evaluation above.
// Let phi be the exact polynomial and r a point where phi(r) - Slin = rho is known.  If phi' >= m > 0 on [-1,1]
// (certified once per pixel by invlin_certify) then for s = z - r:   s > 0: phi(z) - Slin >= m s + rho,
// s < 0: phi(z) - Slin <= m s + rho.  phi_fp differs from phi by at most delta (sum of the float32 roundings of the
// accumulator: each <= 2^-24 |partial sum|), so the comparison is DECIDED without evaluation whenever
// |m s + rho| > delta on the right side; only the last few steps (z_j within the rounding noise of the root) run the
// exact evaluation.  r comes from a few float64 Newton steps started at the previous read's root; Newton need not
// converge for correctness (rho carries whatever residual is left).  delta: global bound (P-1) 2^-24 sum|c_L| far
// from r, the measured partial sums at r (plus their Lipschitz drift) within 2^-16 of r.  The result is the
// reference's z bit for bit (tests/test_hostcheck.py: exhaustive random + adversarial comparison on the host;
// tests/test_gpu_forward.py: exact DN parity with the oracle).
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define RIP_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define RIP_FMA(a, b, c) fma((a), (b), (c))
#endif

// sum |c_L| and a certified lower bound of phi'(z) on [-1,1] (0 = not certified -> plain bisection)
template <int PMAX>
RIP_HD void invlin_certify(const float (&c)[PMAX], int P, float& A_out, float& m_out) {
    double A = 0.0, B2 = 0.0;
    for (int L = 0; L < P; ++L) {
        const double ac = fabs((double)c[L]);
        A += ac;
        B2 += ac * ((double)(L - 1) * L * (L + 1) * (L + 2) / 8.0);  // max |P_L''| on [-1,1] = P_L''(1)
    }
    constexpr int K = 32;
    const double h = 2.0 / K;
    double lo = 1.0e300;
    for (int i = 0; i <= K; ++i) {
        const double z = -1.0 + h * i;
        double p0 = 1.0, p1 = z, d0 = 0.0, d1 = 1.0, dphi = 0.0;
        for (int L = 1; L < P; ++L) {
            dphi += (double)c[L] * d1;
            const double pn = ((2 * L + 1) * z * p1 - L * p0) / (L + 1);
            const double dn = d0 + (2 * L + 1) * p1;
            p0 = p1; p1 = pn; d0 = d1; d1 = dn;
        }
        lo = dphi < lo ? dphi : lo;
    }
    lo = lo - 0.5 * h * B2 * (1.0 + 1e-9) - 1e-9 * A * P * P;
    const bool ok = (A == A) && (A < 1.0e30) && (lo > 1.0e-6 * A) && (lo == lo);
    float af = (float)A;
    if ((double)af < A) af = nextafterf(af, INFINITY);
    float mf = ok ? (float)lo : 0.0f;
    if (ok && (double)mf > lo) mf = nextafterf(mf, -INFINITY);
    A_out = af;
    m_out = (ok && mf > 0.0f) ? mf : 0.0f;
}

// exact evaluation (== legendre_eval<double, PMAX, false>) with the coefficients already promoted to float64
template <int P>
RIP_HD float legendre_eval_cd(double z, const double (&cd)[P]) {
    float phi = (float)cd[0];
    double prev = 1.0, cur = z;
#pragma unroll
    for (int L = 1; L < P; ++L) {
        phi = (float)((double)phi + cd[L] * cur);
        const double a = (2 * L + 1) / (double)(L + 1), b = L / (double)(L + 1);
        const double nxt = (a * z) * cur - b * prev;
        prev = cur;
        cur = nxt;
    }
    return phi;
}

// phi, phi' and sum_L |partial sum_L| (L >= 1) of the exact polynomial, float64 with fused multiply-adds
template <int P>
RIP_HD void legendre_newton_eval(double z, const double (&cd)[P], double& phi, double& dphi, double& sabs) {
    double p0 = 1.0, p1 = z, d0 = 0.0, d1 = 1.0;
    phi = cd[0];
    dphi = 0.0;
    sabs = 0.0;
#pragma unroll
    for (int L = 1; L < P; ++L) {
        phi = RIP_FMA(cd[L], p1, phi);
        dphi = RIP_FMA(cd[L], d1, dphi);
        sabs += fabs(phi);
        const double a = (2 * L + 1) / (double)(L + 1), b = L / (double)(L + 1);
        const double pn = RIP_FMA(a * z, p1, -(b * p0));
        const double dn = RIP_FMA((double)(2 * L + 1), p1, d0);
        p0 = p1; p1 = pn; d0 = d1; d1 = dn;
    }
}

RIP_HD double clamp_pm1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

// Returns the z of the reference's 24-step search.  r: in = starting guess of the root, out = last Newton iterate
// (feed it to the next read of the same pixel).  n_exact (optional) counts exact evaluations.
template <int P>
RIP_HD double invlin_fast_z(double Slin, const double (&cd)[P], float A, float m, double& r, int* n_exact) {
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double slack = 1.0 + 1.0 / 1024.0;
    const double eps = (double)A * 9.094947017729282e-13;  // 2^-40 sum|c|: float64 evaluation errors, generously
    const double dglob = (double)(P - 1) * u * (double)A * slack + eps;
    const double D1 = (double)A * (double)(P * (P - 1) / 2) * (double)(P - 1);  // drift of sum_L |S_L| per unit z
    double rr = clamp_pm1(r), phi, dphi, sabs, rho;
    if (!(rr == rr)) rr = 0.0;
    legendre_newton_eval<P>(rr, cd, phi, dphi, sabs);
    rho = phi - Slin;
    for (int it = 0; it < 8; ++it) {
        if (fabs(rho) <= 0.5 * u * sabs) break;
        const double den = dphi > (double)m ? dphi : (double)m;
        const double rn = clamp_pm1(rr - rho / den);
        if (!(rn != rr)) break;  // converged to the grid of doubles, pinned at +-1, or NaN
        rr = rn;
        legendre_newton_eval<P>(rr, cd, phi, dphi, sabs);
        rho = phi - Slin;
    }
    r = rr;
    // |s| > mu_far  =>  |m s| > dglob + |rho|  =>  m s + rho has the sign of s and clears the global bound: decided.
    // (true for all lanes of a warp during the first ~17 steps: one subtraction and one comparison per step)
    const double mu_far = (dglob + fabs(rho)) / (double)m * (1.0 + 1.0 / 1024.0);
    // The search runs on the grid z = Z 2^-24 (24 halvings of a unit step starting at 0), so the far test is done in
    // integers: Z < Zlo  =>  z < r - mu_far, Z > Zhi  =>  z > r + mu_far, with one grid unit of margin for the roundings of
    // the scaling (the far test is only a SUFFICIENT condition: a step it misses is decided by the tests below, with the
    // same outcome).  Two integer comparisons and an integer add per far step instead of five float64 operations.
    const double SC = 16777216.0, lim = 1073741824.0;
    double dlo = floor(rr * SC - mu_far * SC) - 1.0, dhi = ceil(rr * SC + mu_far * SC) + 1.0;
    if (!(dlo > -lim)) dlo = -lim;  // (also catches NaN: then nothing is "far")
    if (!(dhi < lim)) dhi = lim;
    if (!(mu_far == mu_far) || !(rr == rr)) { dlo = -lim; dhi = lim; }
    const int Zlo = (int)dlo, Zhi = (int)dhi;
    int Z = 0, istep = 1 << 24, j0 = 1;
    // Closed form for the leading run of far steps.  While every visited Z is far, each decision goes towards r, i.e. the
    // search walks the midpoints of the dyadic intervals that contain r: after k steps Z + 2^24 is an odd multiple of
    // 2^(24-k).  Let [t, u] = [Zlo, Zhi] + 2^24 be the not-far window and 2^b the granularity of the COARSEST dyadic point
    // inside it (highest bit in which t-1 and u differ).  No midpoint of granularity > 2^b lies in the window, so the
    // first 23-b visited points are far and the (23-b)-th is the odd multiple of 2^(b+1) of the interval of width
    // 2^(b+2) that contains the window: the loop starts there (about 6 iterations remain instead of 24).
    {
        const int t = Zlo + (1 << 24), uu = Zhi + (1 << 24);
        if (t >= 1 && uu < (1 << 25) && uu >= t) {
            const unsigned x = (unsigned)(t - 1) ^ (unsigned)uu;
#if defined(__CUDA_ARCH__)
            const int b = 31 - __clz((int)x);
#else
            const int b = 31 - __builtin_clz(x);
#endif
            const int nskip = 23 - b;
            if (nskip >= 1) {
                Z = (int)((((unsigned)uu >> (b + 2)) << (b + 2)) + (1u << (b + 1))) - (1 << 24);
                istep = 1 << (b + 1);
                j0 = nskip + 1;
            }
        }
    }
    for (int j = j0; j < 25; ++j) {
        istep >>= 1;
        bool lt;
        if (Z < Zlo) {
            lt = true;
        } else if (Z > Zhi) {
            lt = false;
        } else {
            const double z = (double)Z * (1.0 / 16777216.0);
            const double s = z - rr, as = fabs(s);
            double dl = dglob;
            if (as <= 1.52587890625e-05) {  // 2^-16: partial sums are those at r up to their drift
                const double dloc = u * (sabs + as * D1) * slack + eps;
                dl = dloc < dglob ? dloc : dglob;
            }
            const double g = RIP_FMA((double)m, s, rho);
            if (s > 0.0 && g > dl) lt = false;
            else if (s < 0.0 && g < -dl) lt = true;
            else if (s == 0.0 && fabs(rho) > dl) lt = rho < 0.0;
            else {
                lt = (double)legendre_eval_cd<P>(z, cd) < Slin;
                if (n_exact) ++*n_exact;
            }
        }
        Z += lt ? istep : -istep;
    }
    return (double)Z * (1.0 / 16777216.0);
}

// ---------------------------------------------------------------------------------------------------------
// Ramp-fit plan (built on the host with NumPy so that every scalar has the reference's rounding; see
// romanimpreprocess_b200/utils/fitting.py:build_plan).  Variant 0 is the full ramp, variant v>=1 is the ramp
// truncated at iend = G - v (reference utils/fitting.py:165-169,326).
// ---------------------------------------------------------------------------------------------------------
using RampSlice = rip_ramp_slice;
using RampPlanDev = rip_ramp_plan;

// Exact variance of one slice's slope difference in the reference's op order (fitting.py:233-241; App. A7).
//   TD = float when the gain plane is f32 (dvardt f32, inner term rounded to f32), double when it is f64.
//   Takes only scalars (no register-array indexing) so the cold path never forces the ramp into local memory.
template <typename TD>
RIP_HD_COLD float smap_exact(float delta, int ngrp, TD dvardt, float sig2read, const RampPlanDev& pl, const double* w) {
    double var = 0.0;
    for (int a = 0; a < ngrp; ++a) {
        TD inner = dvardt * (TD)pl.tau[a] + (TD)(sig2read / pl.nreads[a]);
        var = var + (w[a] * w[a]) * (double)inner;
        for (int b = 0; b < a; ++b) var = var + (((double)2 * w[a]) * w[b] * (double)dvardt) * (double)pl.tbar[b];
    }
    return delta / (float)sqrt(var);
}

// sthresh (f64) for a slope (fitting.py:215-217).  logf: CUDA/glibc logf vs NumPy's SIMD logf may differ in the
// last ulp (SURVEY 7 "Bit-exact DQ vs float thresholds"); evaluated via double log and rounded to f32.
RIP_HD_COLD double jump_threshold(float slope, const RampPlanDev& pl) {
    float x = np_clip<float>(slope, pl.IthreshA_f, pl.IthreshB_f);
    float t = (float)log((double)(x / pl.IthreshA_f));
    double xx = (double)t / pl.logIratio;
    return pl.SthreshA + (pl.SthreshB - pl.SthreshA) * xx;
}

struct FitResult {
    float slope, err_read, err_poisson;
    uint32_t jump_mask;  // bit i = JUMP_DET on group i
};

// jump_detect for one pixel and one plan variant (fitting.py:157-255).
//   FAST=false: every slice in the reference's exact op order (used by the stage entry point; writes smap).
//   FAST=true : factorised f32 variance; slices whose significance is within `band` of the threshold are
//               re-evaluated exactly, so the flags are those of the exact path.
// The (i, di) loops are the reference's (fitting.py:225-229) and fully unrolled so d[] is indexed statically.
template <int GMAX, typename TD, bool FAST>
RIP_HD FitResult jump_detect_pixel(const float (&d)[GMAX], int v, TD gain, float read, bool active,
                                   const RampPlanDev& pl, const double* w_all, float* smap_out, long smap_stride) {
    FitResult r;
    const int ngrp = pl.var_ngrp[v];
    const int start = pl.start;
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < GMAX; ++t)
        if (t < ngrp) acc = acc + pl.var_K[v][t] * (d[t] - d[1]);
    r.slope = acc;
    TD gc = np_clip<TD>(gain, (TD)1e-4, (TD)1e4);
    TD dvardt = np_max<TD>((TD)r.slope / gc, (TD)0);
    r.err_poisson = (float)rip_sqrt(np_max<TD>((TD)pl.var_coef[v] * dvardt, (TD)0));
    r.err_read = read * pl.var_rfac[v];
    r.jump_mask = 0u;
    if (FAST && !active) return r;
    const float sig2read = read * read;
    int s = pl.var_slice_off[v];
    const int s0 = s;
    double thr_exact = 0.0;
    bool have_thr = false;
    float hi = 0.0f, lo = 0.0f;
    const float dv = (float)dvardt;
    if (FAST) {
        // approximate threshold: logf is within ~1e-6 relative of the reference's; the band absorbs it
        float x = np_clip<float>(r.slope, pl.IthreshA_f, pl.IthreshB_f);
        float thr = (float)pl.SthreshA + (float)(pl.SthreshB - pl.SthreshA) * (logf(x / pl.IthreshA_f) / (float)pl.logIratio);
        hi = thr * (1.0f + pl.band);
        lo = thr * (1.0f - pl.band);
        if (hi < lo) { float tt = hi; hi = lo; lo = tt; }
    } else {
        thr_exact = jump_threshold(r.slope, pl);
        have_thr = true;
    }
#pragma unroll
    for (int i = 0; i < GMAX - 1; ++i) {
        if (i >= start && i < ngrp - 1) {
            const int dimax = (i == ngrp - 2 || ngrp - 1 - start == 2) ? 1 : 2;
#pragma unroll
            for (int di = 1; di <= 2; ++di) {
                if (di <= dimax) {
                    const RampSlice& sl = pl.slices[s];
                    const float diff = d[(i + di < GMAX) ? (i + di) : (GMAX - 1)] - d[i];
                    if (!FAST) {
                        float sm = smap_exact<TD>(diff / sl.dt - r.slope, ngrp, dvardt, sig2read, pl, w_all + (long)s * RIP_GMAX);
                        if (smap_out) smap_out[(long)(s - s0) * smap_stride] = sm;
                        if (active && ((double)sm > thr_exact)) r.jump_mask |= 1u << i;
                    } else {
                        float var = dv * sl.A + sig2read * sl.B;
                        float sm = (diff * sl.inv_dt - r.slope) / sqrtf(var);
                        if (sm > hi) {
                            r.jump_mask |= 1u << i;
                        } else if (!(sm < lo)) {  // borderline, NaN or degenerate variance -> exact evaluation
                            if (!have_thr) { thr_exact = jump_threshold(r.slope, pl); have_thr = true; }
                            float sme = smap_exact<TD>(diff / sl.dt - r.slope, ngrp, dvardt, sig2read, pl, w_all + (long)s * RIP_GMAX);
                            if ((double)sme > thr_exact) r.jump_mask |= 1u << i;
                        }
                    }
                    ++s;
                }
            }
        }
    }
    return r;
}

// Per-pixel group flags as bit masks over groups (bit g <-> group g).
struct GroupFlags {
    uint32_t dnu, sat, jump, adf;
    uint32_t other_unsat;  // OR over unsaturated groups of any other rdq bits (stage entry point only)
};

// ramp_fit for one pixel (fitting.py:310-355).  Updates gf.jump and pdq.
template <int GMAX, typename TD, bool FAST>
RIP_HD FitResult ramp_fit_pixel(const float (&d)[GMAX], GroupFlags& gf, uint32_t& pdq, TD gain, float read,
                                bool active, const RampPlanDev& pl, const double* w_all) {
    const int G = pl.G;
    FitResult r = jump_detect_pixel<GMAX, TD, FAST>(d, 0, gain, read, active, pl, w_all, nullptr, 0);
    const bool unsat = ((gf.sat >> (G - 1)) & 1u) == 0u;
    if (unsat) gf.jump |= r.jump_mask;
    for (int iend = G - 1; iend > 2 + pl.start; --iend) {
        const bool layer = ((gf.sat >> iend) & 1u) && !((gf.sat >> (iend - 1)) & 1u);
        if (layer) {
            FitResult t = jump_detect_pixel<GMAX, TD, FAST>(d, G - iend, gain, read, active, pl, w_all, nullptr, 0);
            r.slope = t.slope;
            r.err_read = t.err_read;
            r.err_poisson = t.err_poisson;
            gf.jump |= t.jump_mask;
        }
    }
    const uint32_t allg = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t unsat_g = ~gf.sat & allg;
    uint32_t pdq2 = gf.other_unsat & ~DQ_DO_NOT_USE;
    if (gf.jump & unsat_g) pdq2 |= DQ_JUMP_DET;
    if (gf.adf & unsat_g) pdq2 |= DQ_AD_FLOOR;
    if ((gf.dnu & allg) == allg) pdq2 |= DQ_DO_NOT_USE;
    if ((gf.sat >> (1 + pl.start)) & 1u) pdq2 |= DQ_DO_NOT_USE;
    if (gf.sat & allg) pdq2 |= DQ_SATURATED;
    if ((pdq & DQ_REFERENCE_PIXEL) == 0u) pdq |= pdq2;
    return r;
}

// do_ramp_fit packaging + dark + error split + flat (gen_cal_image.py:458-475,223-229,607-629).
//   flat_area = f32(flat_ipc / AreaFactor) is computed by the caller (f64 or f32 area plane).
RIP_HD void l2_epilogue(FitResult& r, bool active, float dark_slope_ipc, float flat_area) {
    float err = (float)sqrt((double)r.err_read * (double)r.err_read + (double)r.err_poisson * (double)r.err_poisson);
    float varp = r.err_poisson * r.err_poisson;
    float slope = r.slope;
    if (!active) { slope = 0.0f; varp = 0.0f; err = 0.0f; }
    if (active) slope = slope - dark_slope_ipc;
    float ep = sqrtf(varp);
    float er = sqrtf(np_max<float>(err * err - ep * ep, 0.0f));
    r.slope = slope / flat_area;
    r.err_read = er / flat_area;
    r.err_poisson = ep / flat_area;
}

}  // namespace rip
