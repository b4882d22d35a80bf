// Forward model (scene electrons -> L1 resultants): IL.apply (reference utils/ipc_linearity.py:461-513) and
// make_l1_fullcal (reference from_sim/sim_to_isim.py:163-262) with romanisim's apportioning / read noise
// restated (parity unpinned, validated statistically).
//
// ALU-bound (24 bisection steps x Legendre order x n_reads in float64), not HBM-bound.
#include <memory>
#include <vector>

#include "rip_launch.h"
#include "rip_rng.cuh"

namespace rip {

// Stirling tail log(k!) - [(k+1/2)log(k+1) - (k+1) + (1/2)log(2pi)]  (Hormann 1993)
__constant__ double c_stirling[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509,
                                      0.0166446911898211, 0.0138761288230707, 0.0118967099458917, 0.0104112652619720,
                                      0.00925546218271273, 0.00833056343336287};
__device__ __forceinline__ double stirling_tail(double k) {
    if (k <= 9.0) return c_stirling[(int)k];
    const double kp1sq = (k + 1.0) * (k + 1.0);
    return (1.0 / 12.0 - (1.0 / 360.0 - 1.0 / 1260.0 / kp1sq) / kp1sq) / (k + 1.0);
}

// Per-read constants of the apportioning (pixel independent): the binomial probability of read k given the reads
// before it, folded to q = min(p, 1-p), and what the float32 sampler derives from q alone.
struct ReadTab {
    double p;
    float q, omq, l1mq, s;  // q, 1-q, log1p(-q), q/(1-q)
    int flip, draw;         // p > 1/2; this read draws at all (t > t_prev)
};

__device__ long binomial_draw(Philox& rng, long n, double p);

// Binomial(n, p_k) for n < 2^22 with the per-read constants (same algorithm as binomial_draw's float32 branch)
__device__ __forceinline__ int binomial_draw_tab(Philox& rng, int n, const ReadTab& T) {
    if (n <= 0 || T.p <= 0.0) return 0;
    if (T.p >= 1.0) return n;
    const float nf = (float)n, qf = T.q, omq = T.omq;
    const float npq = nf * qf;
    int k;
    // (inversion up to n q = 64 -- NumPy switches at 30: for small means BTRS's squeeze accepts about half of the candidates
    //  and every other one pays four float64 logarithms, while an inversion step is 8 float32 instructions; measured on
    //  B200: threshold 10 -> 54 ms, 30 -> 9.9 ms per 4096^2 x 35 reads on a 300-electron scene)
    const float lf0 = nf * T.l1mq;  // log of P(X = 0) = (1-q)^n
    // the float32 recurrence needs P(X = 0) well inside the normal range (__expf flushes denormals to zero, and a zero
    // start would never terminate): e^-60 = 9e-27
    if (npq < 64.0f && lf0 > -60.0f) {
        const float f0 = __expf(lf0), s = T.s;
        const float bound = fminf(nf, npq + 10.0f * sqrtf(npq * omq + 1.0f));
        float x = 0.0f, f = f0, u = rng.uniform();
        int guard = 0;
        while (u > f) {
            x += 1.0f;
            if (x > bound) {
                if (++guard > 64) { x = rintf(npq); break; }  // (unreachable for f0 >= e^-60; a bound on the loop all the same)
                x = 0.0f; f = f0; u = rng.uniform();
            } else { u -= f; f = __fdividef((nf - x + 1.0f) * s * f, x); }
        }
        k = (int)x;
    } else {
        const float spq = sqrtf(npq * omq);
        const float b = 1.15f + 2.53f * spq, a = -0.0873f + 0.0248f * b + 0.01f * qf, c = npq + 0.5f;
        const float vr = 0.92f - __fdividef(4.2f, b);
        for (;;) {
            const float u = rng.uniform() - 0.5f;
            const float v = rng.uniform();
            const float us = 0.5f - fabsf(u);
            const float kk = floorf((__fdividef(2.0f * a, us) + b) * u + c);
            if (kk < 0.0f || kk > nf) continue;
            if (us >= 0.07f && v <= vr) { k = (int)kk; break; }
            const double nd = (double)n, kd = (double)kk, usd = (double)us, q = (double)qf;
            const double r = q / (1.0 - q), alpha = (2.83 + 5.1 / (double)b) * (double)spq;
            const double m = floor((nd + 1.0) * q);
            const double lv = log((double)v * alpha / ((double)a / (usd * usd) + (double)b));
            const double ub = (m + 0.5) * log((m + 1.0) / (r * (nd - m + 1.0))) +
                              (nd + 1.0) * log((nd - m + 1.0) / (nd - kd + 1.0)) +
                              (kd + 0.5) * log(r * (nd - kd + 1.0) / (kd + 1.0)) + stirling_tail(m) +
                              stirling_tail(nd - m) - stirling_tail(kd) - stirling_tail(nd - kd);
            if (lv <= ub) { k = (int)kk; break; }
        }
    }
    return T.flip ? n - k : k;
}

// Binomial(n, p) sampler: sequential inversion of the CDF (BINV) for n*min(p,1-p) < 10, Hormann's BTRS transformed
// rejection otherwise.  For n < 2^22 (every realistic well) the set-up, the inversion recurrence and the BTRS squeeze
// run in float32 (integers up to n are exact; the relative error 1e-7 of the probabilities is far below anything a
// realisation ensemble can resolve); the rarely reached exact acceptance test of BTRS and larger n use float64.
__device__ long binomial_draw(Philox& rng, long n, double p) {
    if (n <= 0 || p <= 0.0) return 0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    const double q = flip ? 1.0 - p : p;
    long k;
    if (n < (1L << 22)) {
        const float nf = (float)n, qf = (float)q, omq = 1.0f - qf;
        const float npq = nf * qf;
        if (npq < 10.0f) {
            const float f0 = expf(nf * log1pf(-qf)), s = qf / omq;
            const float bound = fminf(nf, npq + 10.0f * sqrtf(npq * omq + 1.0f));
            float x = 0.0f, f = f0, u = rng.uniform();
            while (u > f) {
                x += 1.0f;
                if (x > bound) { x = 0.0f; f = f0; u = rng.uniform(); }
                else { u -= f; f = ((nf - x + 1.0f) * s * f) / x; }
            }
            k = (long)x;
        } else {
            const float spq = sqrtf(npq * omq);
            const float b = 1.15f + 2.53f * spq, a = -0.0873f + 0.0248f * b + 0.01f * qf, c = npq + 0.5f;
            const float vr = 0.92f - 4.2f / b;
            for (;;) {
                const float u = rng.uniform() - 0.5f;
                const float v = rng.uniform();
                const float us = 0.5f - fabsf(u);
                const float kk = floorf((2.0f * a / us + b) * u + c);
                if (kk < 0.0f || kk > nf) continue;
                if (us >= 0.07f && v <= vr) { k = (long)kk; break; }
                const double nd = (double)n, kd = (double)kk, usd = (double)us;
                const double r = q / (1.0 - q), alpha = (2.83 + 5.1 / (double)b) * (double)spq;
                const double m = floor((nd + 1.0) * q);
                const double lv = log((double)v * alpha / ((double)a / (usd * usd) + (double)b));
                const double ub = (m + 0.5) * log((m + 1.0) / (r * (nd - m + 1.0))) +
                                  (nd + 1.0) * log((nd - m + 1.0) / (nd - kd + 1.0)) +
                                  (kd + 0.5) * log(r * (nd - kd + 1.0) / (kd + 1.0)) + stirling_tail(m) +
                                  stirling_tail(nd - m) - stirling_tail(kd) - stirling_tail(nd - kd);
                if (lv <= ub) { k = (long)kk; break; }
            }
        }
        return flip ? n - k : k;
    }
    if ((double)n * q < 10.0) {
        // sequential inversion via geometric waiting times
        const double lq = log1p(-q);
        long x = 0, sum = 0;
        for (;;) {
            const double u = rng.uniform53();
            sum += (long)floor(log(u) / lq) + 1;
            if (sum > n) break;
            ++x;
        }
        k = x;
    } else {
        const double nd = (double)n;
        const double spq = sqrt(nd * q * (1.0 - q));
        const double b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * q, c = nd * q + 0.5;
        const double vr = 0.92 - 4.2 / b, r = q / (1.0 - q), alpha = (2.83 + 5.1 / b) * spq;
        const double m = floor((nd + 1.0) * q);
        for (;;) {
            const double u = rng.uniform53() - 0.5;
            double v = rng.uniform53();
            const double us = 0.5 - fabs(u);
            const double kk = floor((2.0 * a / us + b) * u + c);
            if (kk < 0.0 || kk > nd) continue;
            if (us >= 0.07 && v <= vr) { k = (long)kk; break; }
            v = log(v * alpha / (a / (us * us) + b));
            const double ub = (m + 0.5) * log((m + 1.0) / (r * (nd - m + 1.0))) +
                              (nd + 1.0) * log((nd - m + 1.0) / (nd - kk + 1.0)) +
                              (kk + 0.5) * log(r * (nd - kk + 1.0) / (kk + 1.0)) + stirling_tail(m) +
                              stirling_tail(nd - m) - stirling_tail(kk) - stirling_tail(nd - kk);
            if (v <= ub) { k = (long)kk; break; }
        }
    }
    return flip ? n - k : k;
}

// ---------------------------------------------------------------------------------------------------------
// IL.apply chain on a window [ny,nx] of pitched calibration planes
// ---------------------------------------------------------------------------------------------------------
template <typename TC, typename TS, typename TO>
__global__ void il_sum_kernel(const TC* __restrict__ counts, const TS* __restrict__ start_e, double start_scalar, long npix,
                              TO* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    out[p] = start_e ? (TO)counts[p] + (TO)start_e[p] : (TO)counts[p] + (TO)start_scalar;
}

// conv (type TI) = 9-tap source-indexed IPC of im; optional; then / gain -> Slin (TL); bisection -> S (TL)
template <typename TIM, typename TK, typename TG, typename TL, int PMAX>
__global__ void il_chain_kernel(const TIM* __restrict__ im, const TK* __restrict__ K, const TG* __restrict__ gain,
                                long gain_off, int gain_pitch, int ny, int nx, const float* __restrict__ coefs,
                                long lin_plane, long lin_off, int lin_pitch, int P, const float* __restrict__ Smin,
                                const float* __restrict__ Smax, const float* __restrict__ Sref, int gain_in, int electrons_out,
                                double* __restrict__ out) {
    typedef typename Promote<TIM, TK>::type TI;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    const long pl = (long)ny * nx;
    TI conv;
    if (K) {
        const int DY[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1};
        const int DX[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
        conv = (TI)im[(long)y * nx + x] * (TI)K[4 * pl + (long)y * nx + x];
#pragma unroll
        for (int q = 1; q < 9; ++q) {
            const int ys = y - DY[q], xs = x - DX[q];
            if (ys >= 0 && ys < ny && xs >= 0 && xs < nx)
                conv = conv + (TI)im[(long)ys * nx + xs] * (TI)K[(long)((1 + DY[q]) * 3 + (1 + DX[q])) * pl + (long)ys * nx + xs];
        }
    } else {
        conv = (TI)im[(long)y * nx + x];
    }
    const long lp = lin_off + (long)y * lin_pitch + x;
    TL slin = gain_in ? (TL)conv / (TL)gain[gain_off + (long)y * gain_pitch + x] : (TL)conv;
    float c[PMAX];
#pragma unroll
    for (int L = 0; L < PMAX; ++L) c[L] = (L < P) ? coefs[(long)L * lin_plane + lp] : 0.0f;
    bool ex;
    TL S = invlin_pixel<TL, PMAX>(slin, c, P, Smin[lp], Smax[lp], ex);
    double r = (double)S;
    if (electrons_out) {
        // g_out * (S - Sref): S has type TL, Sref f32, gain TG  (ipc_linearity.py:513)
        TL diff = S - (TL)Sref[lp];
        r = (double)((typename Promote<TL, TG>::type)gain[gain_off + (long)y * gain_pitch + x] * (typename Promote<TL, TG>::type)diff);
    }
    out[(long)y * nx + x] = r;
}

struct IlWindow {
    int ny, nx;
    const void* K; int k_dtype;
    const void* gain; int g_dtype; long gain_off; int gain_pitch;
    const float *coefs, *Smin, *Smax, *Sref; long lin_plane, lin_off; int lin_pitch; int P;
};

template <typename TIM, typename TK, typename TG, typename TL>
static void il_chain_launch(const void* im, const IlWindow& w, bool use_gain, int electrons_out, double* out, cudaStream_t st) {
    dim3 grid((w.nx + 127) / 128, w.ny), block(128);
#define ILC(PM)                                                                                                        \
    RIP_LAUNCH((il_chain_kernel<TIM, TK, TG, TL, PM>), grid, block, 0, st, (const TIM*)im, (const TK*)w.K,              \
               (const TG*)((use_gain || electrons_out) ? w.gain : nullptr), w.gain_off, w.gain_pitch, w.ny, w.nx, w.coefs, \
               w.lin_plane, w.lin_off, w.lin_pitch, w.P, w.Smin, w.Smax, w.Sref, use_gain ? 1 : 0, electrons_out, out)
    if (w.P <= 4) ILC(4);
    else if (w.P <= 11) ILC(11);
    else ILC(RIP_PMAX);
#undef ILC
}

// Runs IL.apply on device data.  d_counts: device window [ny,nx] of c_dtype; d_start: device f32/f64 plane or null.
// Returns the NumPy result dtype tag (values are written as f64).
static int il_apply_device(const void* d_counts, int c_dtype, const void* d_start, int s_dtype, double start_scalar,
                           const IlWindow& w, int electrons, int electrons_out, double* d_out, void* d_tmp,
                           cudaStream_t st) {
    const long npix = (long)w.ny * w.nx;
    // dtype of counts + start_e: ints (>=32 bit) + f32 -> f64; f32 + (f32 | python scalar) -> f32
    const bool cd = (c_dtype == RIP_F64 || c_dtype == RIP_I32);
    const bool sd = d_start && s_dtype == RIP_F64;
    const bool sumd = cd || sd;
    const unsigned nb = (unsigned)((npix + 255) / 256);
#define SUM(TC, TS, TO) RIP_LAUNCH((il_sum_kernel<TC, TS, TO>), nb, 256, 0, st, (const TC*)d_counts, (const TS*)d_start, start_scalar, npix, (TO*)d_tmp)
    if (c_dtype == RIP_I32) { if (sd) SUM(int32_t, double, double); else SUM(int32_t, float, double); }
    else if (c_dtype == RIP_F64) { if (sd) SUM(double, double, double); else SUM(double, float, double); }
    else { if (sd) SUM(float, double, double); else SUM(float, float, float); }
#undef SUM
    const bool kd = w.K && w.k_dtype == RIP_F64;
    const bool gd = w.g_dtype == RIP_F64;
    const bool convd = sumd || kd;
    const bool use_gain = electrons != 0;
    const bool slind = convd || (use_gain && gd);
    // TL = dtype of counts_conv / g_in
    if (!sumd) {
        if (!kd) { if (slind) il_chain_launch<float, float, double, double>(d_tmp, w, use_gain, electrons_out, d_out, st);
                   else il_chain_launch<float, float, float, float>(d_tmp, w, use_gain, electrons_out, d_out, st); }
        else { if (gd) il_chain_launch<float, double, double, double>(d_tmp, w, use_gain, electrons_out, d_out, st);
               else il_chain_launch<float, double, float, double>(d_tmp, w, use_gain, electrons_out, d_out, st); }
    } else {
        if (!kd) { if (gd) il_chain_launch<double, float, double, double>(d_tmp, w, use_gain, electrons_out, d_out, st);
                   else il_chain_launch<double, float, float, double>(d_tmp, w, use_gain, electrons_out, d_out, st); }
        else { if (gd) il_chain_launch<double, double, double, double>(d_tmp, w, use_gain, electrons_out, d_out, st);
               else il_chain_launch<double, double, float, double>(d_tmp, w, use_gain, electrons_out, d_out, st); }
    }
    bool outd = slind;
    if (electrons_out && gd) outd = true;
    return outd ? RIP_F64 : RIP_F32;
}

// ---------------------------------------------------------------------------------------------------------
// K3: forward ramp, two kernels.
//   fwd_apportion_kernel  one thread per active pixel: reset-noise electrons (start_e) and the cumulative electrons
//                         at every read (binomial apportioning of the exposure's total) -> HBM (4 B per pixel and read;
//                         140 MB per read plane set at 4096^2 -- the divergent, log-heavy sampler stays out of the
//                         arithmetic kernel and no halo pixel is ever drawn twice)
//   fwd_ramp_kernel       one CTA = 32 x 8 active pixels, all threads owners; per read: (32+2) x (8+2) halo tile of
//                         electrons in shared memory, IPC 3x3, /gain, certified fast inverse (rip_math.cuh
//                         invlin_fast_z: the z of the reference's 24-step float64 search with ~2-3 Newton
//                         evaluations + the last few exact ones instead of 24), group mean, read noise, biascorr
// ---------------------------------------------------------------------------------------------------------
struct FwdArgs {
    int n, nb, na, G, P, n_reads;
    int reads_per_group[RIP_GMAX];
    int read_index[64];
    double read_time;
    uint64_t seed;
    int add_read_noise, add_reset_noise, add_biascorr, quantize;
    double biascorr_t0;
    int has_bias;
    ReadTab rtab[64];           // per-read apportioning constants (by value: kernel parameters live in constant memory)
    float extrap[64];           // (t_k - t_{k-1}) / (t_{k-1} - t_{k-2}): linear extrapolation of the root from read to read
    const int32_t* counts;      // [na,na] total electrons of the exposure (already Poisson)
    int32_t* cum;               // [n_reads,na,na] cumulative electrons per read (written by the apportioning kernel
                                //  unless supplied by the caller: tests)
    int cum_given;
    float* start;               // [na,na] electrons in the well at the reset
    const float* linA; const float* linm;  // [n,n] certificate planes of the fast inverse
    const float* coefs; const float* Smin; const float* Smax;  // full-frame planes
    const void* gain; const void* ipc; const float* read; const float* resetnoise; const float* dark_slope;
    const float* bias;          // [G,na,na] or null (offset applied)
    float* out;                 // [G,na,na]
};

constexpr int FTX = 32, FTY = 4;

__global__ void invlin_certify_kernel(const float* __restrict__ coefs, int P, long npl, float* __restrict__ linA,
                                      float* __restrict__ linm) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npl) return;
    float c[RIP_PMAX];
#pragma unroll
    for (int L = 0; L < RIP_PMAX; ++L) c[L] = (L < P) ? coefs[(long)L * npl + p] : 0.0f;
    float A, m;
    invlin_certify<RIP_PMAX>(c, P, A, m);
    linA[p] = A;
    linm[p] = m;
}

template <typename TG>
__global__ void __launch_bounds__(128) fwd_apportion_kernel(const FwdArgs A) {
    const int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    const int na = A.na, n = A.n, nb = A.nb;
    if (xa >= na) return;
    const long pa = (long)ya * na + xa, pf = (long)(ya + nb) * n + (xa + nb), npa = (long)na * na;
    Philox rng;
    rng.init(A.seed, (uint64_t)pa, 1u);
    // reset noise in electrons (sim_to_isim.py:195-215): N(0,1)*resetnoise*gain - t0*dark_slope/gain, float32
    const TG g = ((const TG*)A.gain)[pf];
    float start_e = 0.0f;
    if (A.add_reset_noise) {
        float rn = rng.normal();
        rn = rn * A.resetnoise[pf];
        rn = (float)((typename Promote<float, TG>::type)rn * (typename Promote<float, TG>::type)g);
        start_e = rn;
    }
    if (A.has_bias) {
        typedef typename Promote<float, TG>::type TP;
        // tbias * dark_slope / gain : python float * f32 array -> f32, / gain -> TP
        const float td = (float)A.biascorr_t0 * A.dark_slope[pf];
        start_e = (float)((TP)start_e - (TP)td / (TP)g);
    }
    A.start[pa] = start_e;
    if (A.cum_given) return;
    int remaining = 0;
    if (A.counts) {
        const int c = A.counts[pa];
        remaining = c < 0 ? 0 : (c > 2000000000 ? 2000000000 : c);
    }
    int cum = 0;
    for (int k = 0; k < A.n_reads; ++k) {
        const ReadTab& T = A.rtab[k];
        if (remaining > 0 && T.draw) {
            const int d = remaining < (1 << 22) ? binomial_draw_tab(rng, remaining, T) : (int)binomial_draw(rng, (long)remaining, T.p);
            cum += d;
            remaining -= d;
        }
        A.cum[(long)k * npa + pa] = cum;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Cosmic rays: romanisim.cr.simulate_crs (romanisim 0.x, called per read by l1.apportion_counts_to_resultants when
// crparam is not None -- the reference passes crparam={} = the defaults, from_sim/sim_to_isim.py:233-242), restated
// from the published algorithm (third-party source absent: parity unpinned, validated statistically):
//   N ~ Poisson(flux * area * dt) events per read; position uniform on the array, direction phi uniform in [0, 2 pi),
//   projected path length [um] from a power law x^-4.33 on [10, 2000], energy loss dE/dx [eV/um] from a Moyal
//   distribution (location 120, scale 50) on [10, 10000], both by inverse-transform sampling of the pdf tabulated on
//   10000 grid points (cumulative sum, linear interpolation of the inverse); the end point is clipped to
//   [-0.5, N + 0.5] per axis; every pixel the segment crosses receives Poisson(dE/dx * pixel_size / conversion_factor
//   * sqrt((depth / pixel_size)^2 + l2d^2)) electrons, l2d = path length inside the pixel [pixels].
// The electrons stay in the well: they are added to the cumulative counts of the read and of every later read.
// One CTA per read; one thread per event (a few hundred events per read at the default flux).
// ---------------------------------------------------------------------------------------------------------
constexpr int CR_GRID = 10000;
struct CrArgs {
    int na, n_reads;
    uint64_t seed;
    double lam[64];             // expected events per read: flux * area * (t_k - t_{k-1})
    int group_of_read[64];
    double counts_per_pix;      // pixel_size / conversion_factor  (times dE/dx)
    double depth_ratio;         // pixel_depth / pixel_size
    double inv_pixel_size;      // 1 / pixel_size [1/um]
    const float* len_cdf;       // [CR_GRID] cumulative distributions of the two samplers
    const float* dedx_cdf;
    int32_t* cum;               // [n_reads, na, na]
    uint32_t* groups;           // [na, na] bit g = a cosmic ray deposited electrons during group g
};

// inverse-transform sample: x(u) with cdf tabulated on linspace(lo, hi, CR_GRID)
__device__ inline double cr_sample(const float* __restrict__ cdf, double lo, double hi, float u) {
    int a = 0, b = CR_GRID - 1;  // cdf[a] <= u <= cdf[b]
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (cdf[m] <= u) a = m; else b = m;
    }
    const float ca = cdf[a], cb = cdf[b];
    const double f = cb > ca ? (double)(u - ca) / (double)(cb - ca) : 0.0;
    return lo + (hi - lo) * ((double)a + f) / (double)(CR_GRID - 1);
}

__global__ void __launch_bounds__(128) fwd_cr_kernel(const CrArgs A) {
    const int k = blockIdx.x;
    __shared__ int n_ev;
    if (threadIdx.x == 0) {
        Philox r;
        r.init(A.seed, (uint64_t)k, 200u);
        long nn = poisson_draw(r, A.lam[k]);
        n_ev = nn > (1 << 20) ? (1 << 20) : (int)nn;
    }
    __syncthreads();
    const int na = A.na;
    const long npa = (long)na * na;
    const uint32_t gbit = 1u << A.group_of_read[k];
    for (int ev = threadIdx.x; ev < n_ev; ev += blockDim.x) {
        Philox r;
        r.init(A.seed, ((uint64_t)k << 32) | (uint64_t)ev, 201u);
        const double i0 = (double)r.uniform() * na, j0 = (double)r.uniform() * na;
        const double phi = 6.283185307179586 * (double)r.uniform();
        const double len = cr_sample(A.len_cdf, 10.0, 2000.0, r.uniform()) * A.inv_pixel_size;
        const double dedx = cr_sample(A.dedx_cdf, 10.0, 10000.0, r.uniform());
        double sn, cs;
        sincos(phi, &sn, &cs);
        const double i1 = fmin(fmax(i0 + len * cs, -0.5), na + 0.5), j1 = fmin(fmax(j0 + len * sn, -0.5), na + 0.5);
        const double di = i1 - i0, dj = j1 - j0;
        const double L = sqrt(di * di + dj * dj);
        const double cpp = dedx * A.counts_per_pix;
        // walk the pixel-border crossings (borders at half-integers, pixel centres at integers) in order of t in [0, 1]
        const double inv_di = di != 0.0 ? 1.0 / di : 0.0, inv_dj = dj != 0.0 ? 1.0 / dj : 0.0;
        double bi = di > 0.0 ? floor(i0 + 0.5) + 0.5 : ceil(i0 - 0.5) - 0.5;   // next border along i
        double bj = dj > 0.0 ? floor(j0 + 0.5) + 0.5 : ceil(j0 - 0.5) - 0.5;
        const double si = di > 0.0 ? 1.0 : -1.0, sj = dj > 0.0 ? 1.0 : -1.0;
        double t = 0.0;
        for (int guard = 0; guard < 1024 && t < 1.0; ++guard) {
            double ti = di != 0.0 ? (bi - i0) * inv_di : 2.0, tj = dj != 0.0 ? (bj - j0) * inv_dj : 2.0;
            if (ti <= t) { bi += si; continue; }   // (a start exactly on a border)
            if (tj <= t) { bj += sj; continue; }
            double tn = fmin(fmin(ti, tj), 1.0);
            const double tm = 0.5 * (t + tn);
            const long ii = llrint(i0 + tm * di), jj = llrint(j0 + tm * dj);   // pixel of the segment's midpoint
            const double l2 = (tn - t) * L;
            if (ii >= 0 && ii < na && jj >= 0 && jj < na && l2 > 0.0) {
                const double l3 = sqrt(A.depth_ratio * A.depth_ratio + l2 * l2);
                const long d = poisson_draw(r, cpp * l3);
                if (d > 0) {
                    const long pa = ii * na + jj;
                    const int dd = d > 1000000000L ? 1000000000 : (int)d;
                    for (int kk = k; kk < A.n_reads; ++kk) atomicAdd(&A.cum[(long)kk * npa + pa], dd);
                    atomicOr(&A.groups[pa], gbit);
                }
            }
            if (tn == ti) bi += si;
            if (tn == tj) bj += sj;
            t = tn;
        }
    }
}

// the two cumulative tables of romanisim.cr.create_sampler: cumsum(pdf) - pdf[0], normalised to its maximum
static void cr_tables(std::vector<float>& len_cdf, std::vector<float>& dedx_cdf) {
    len_cdf.resize(CR_GRID);
    dedx_cdf.resize(CR_GRID);
    std::vector<double> y(CR_GRID);
    auto build = [&](std::vector<float>& out) {
        double acc = 0.0;
        std::vector<double> c(CR_GRID);
        for (int i = 0; i < CR_GRID; ++i) { acc += y[i]; c[i] = acc - y[0]; }
        const double mx = c[CR_GRID - 1];
        for (int i = 0; i < CR_GRID; ++i) out[i] = (float)(c[i] / mx);
    };
    for (int i = 0; i < CR_GRID; ++i) {
        const double x = 10.0 + (2000.0 - 10.0) * i / (CR_GRID - 1);
        y[i] = pow(x, -4.33);
    }
    build(len_cdf);
    for (int i = 0; i < CR_GRID; ++i) {
        const double x = 10.0 + (10000.0 - 10.0) * i / (CR_GRID - 1);
        const double xs = (x - 120.0) / 50.0;
        y[i] = exp(-(xs + exp(-xs)) / 2.0);
    }
    build(dedx_cdf);
}

template <int P, typename TG, typename TK>
__global__ void __launch_bounds__(FTX * FTY, 6) fwd_ramp_kernel(const FwdArgs A) {
    __shared__ double se[2][FTY + 2][FTX + 2];  // electrons in the well at this read (counts so far + start_e), double-buffered
    __shared__ float ss[FTY + 2][FTX + 2];   // start_e of the tile + halo
    const int tid = threadIdx.x, tx = tid % FTX, ty = tid / FTX;
    const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY;
    const int xa = x0 + tx, ya = y0 + ty;
    const int na = A.na, n = A.n, nb = A.nb;
    const bool owner = xa < na && ya < na;
    const long pa = (long)ya * na + xa;
    const long pf = (long)(ya + nb) * n + (xa + nb);
    const long npa = (long)na * na, npl = (long)n * n;
    constexpr int HALO = (FTX + 2) * (FTY + 2);
    // the (at most two) halo-tile entries this thread stages per read
    long hsrc[2];
    int hpos[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * FTX * FTY;
        hpos[q] = -1;
        hsrc[q] = -1;
        if (e < HALO) {
            const int hy = e / (FTX + 2), hx = e - hy * (FTX + 2);
            const int gx = x0 + hx - 1, gy = y0 + hy - 1;
            hpos[q] = e;
            if (gx >= 0 && gx < na && gy >= 0 && gy < na) hsrc[q] = (long)gy * na + gx;
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
        if (hpos[q] >= 0) (&ss[0][0])[hpos[q]] = hsrc[q] >= 0 ? A.start[hsrc[q]] : 0.0f;

    double cd[P];
    float smin = 0.f, smax = 1.f, linA = 0.f, linm = 0.f;
    TG g = (TG)1;
    TK kt[9];
    bool kok[9];
    const TK* K = (const TK*)A.ipc;
    {
        const int DY[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1};
        const int DX[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int ys = ya - DY[q], xs = xa - DX[q];
            kok[q] = owner && K && ys >= 0 && ys < na && xs >= 0 && xs < na;
            kt[q] = kok[q] ? K[(long)((1 + DY[q]) * 3 + (1 + DX[q])) * npa + (long)ys * na + xs] : (TK)0;
        }
    }
    if (owner) {
#pragma unroll
        for (int L = 0; L < P; ++L) cd[L] = (L < A.P) ? (double)A.coefs[(long)L * npl + pf] : 0.0;
        smin = A.Smin[pf];
        smax = A.Smax[pf];
        linA = A.linA[pf];
        linm = A.linm[pf];
        g = ((const TG*)A.gain)[pf];
    } else {
#pragma unroll
        for (int L = 0; L < P; ++L) cd[L] = 0.0;
    }
    const float half = (smax - smin) / 2.0f;  // f32 expression in the reference (ipc_linearity.py:390)
    double root = 0.0, root_pp = 0.0;
    bool have_root = false;
    int k = 0;
    // cumulative electrons of the NEXT read for this thread's halo entries: loaded one read ahead, so that the global
    // loads fly during the inversion of the current read (ncu before: 1.1 long-scoreboard stall cycles per issue)
    int cnext[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; ++q)
        if (hsrc[q] >= 0) cnext[q] = A.cum[hsrc[q]];
    for (int grp = 0; grp < A.G; ++grp) {
        double acc = 0.0;
        for (int r = 0; r < A.reads_per_group[grp]; ++r, ++k) {
            // one barrier per read: read k stages into buffer k & 1, which was last read during read k-2, i.e. before
            // every thread arrived at the barrier of read k-1
            double (*seb)[FTX + 2] = se[k & 1];
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (hpos[q] >= 0) (&seb[0][0])[hpos[q]] = hsrc[q] >= 0 ? (double)cnext[q] + (double)(&ss[0][0])[hpos[q]] : 0.0;
            __syncthreads();
            if (k + 1 < A.n_reads) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (hsrc[q] >= 0) cnext[q] = A.cum[(long)(k + 1) * npa + hsrc[q]];
            }
            if (owner) {
                const int sy = ty + 1, sx = tx + 1;
                double (*se)[FTX + 2] = seb;  // (the names below are those of the single-buffer version)
                double conv;
                if (K) {
                    conv = se[sy][sx] * (double)kt[0];
                    if (kok[1]) conv = conv + se[sy - 1][sx] * (double)kt[1];
                    if (kok[2]) conv = conv + se[sy + 1][sx] * (double)kt[2];
                    if (kok[3]) conv = conv + se[sy][sx - 1] * (double)kt[3];
                    if (kok[4]) conv = conv + se[sy][sx + 1] * (double)kt[4];
                    if (kok[5]) conv = conv + se[sy - 1][sx - 1] * (double)kt[5];
                    if (kok[6]) conv = conv + se[sy - 1][sx + 1] * (double)kt[6];
                    if (kok[7]) conv = conv + se[sy + 1][sx - 1] * (double)kt[7];
                    if (kok[8]) conv = conv + se[sy + 1][sx + 1] * (double)kt[8];
                } else {
                    conv = se[sy][sx];
                }
                const double slin = conv / (double)g;
                double z;
                if (linm > 0.0f) {
                    if (!have_root) {
                        root = cd[1] != 0.0 ? (slin - cd[0]) / cd[1] : 0.0;
                        root_pp = root;
                        have_root = true;
                    }
                    // starting guess: the root moves almost linearly in time (constant flux, mild non-linearity), so the
                    // extrapolation from the last two reads lands within the Poisson noise of the increment and one
                    // Newton step suffices even for bright pixels
                    const double guess = root + (root - root_pp) * (double)A.extrap[k];
                    root_pp = root;
                    root = guess;
                    z = invlin_fast_z<P>(slin, cd, linA, linm, root, nullptr);
                } else {  // no monotonicity certificate for this pixel: the plain search
                    z = 0.0;
                    double step = 1.0;
                    for (int j = 1; j < 25; ++j) {
                        step = step * 0.5;
                        const float phi = legendre_eval_cd<P>(z, cd);
                        z = z + (((double)phi < slin) ? step : -step);
                    }
                }
                const double S = (double)smin + (double)half * (1.0 + z);
                acc = acc + S;
            }
        }
        if (owner) {
            // romanisim: resultant = mean of the reads in the group, stored float32
            float res = (float)(acc / (double)A.reads_per_group[grp]);
            if (A.add_read_noise) {
                Philox rn;
                rn.init(A.seed, (uint64_t)pa, 16u + (uint32_t)grp);
                res = res + rn.normal() * (A.read[pf] / sqrtf((float)A.reads_per_group[grp]));
            }
            if (A.add_biascorr && A.bias) res = res + A.bias[(long)grp * npa + pa];
            if (A.quantize) res = rintf(res);
            A.out[(long)grp * npa + pa] = res;
        }
    }
}

}  // namespace rip

using namespace rip;

extern "C" int rip_il_apply_planes(int device, const void* counts, int c_dtype, int ny, int nx, const void* start_e,
                                   int s_dtype, double start_scalar, const void* kernel, int k_dtype, const void* gain,
                                   int g_dtype, const float* coefs, int P, const float* Smin, const float* Smax,
                                   const float* Sref, int electrons, int electrons_out, double* out, int* out_dtype) {
    RIP_API_BEGIN
    RIP_REQUIRE(counts && coefs && Smin && Smax && out, "rip_il_apply_planes: null argument");
    RIP_REQUIRE(P >= 1 && P <= RIP_PMAX, "rip_il_apply_planes: P=%d outside 1..%d", P, RIP_PMAX);
    RIP_REQUIRE(c_dtype == RIP_F32 || c_dtype == RIP_F64 || c_dtype == RIP_I32, "rip_il_apply_planes: counts dtype");
    RIP_REQUIRE(!(electrons || electrons_out) || gain, "rip_il_apply_planes: gain needed for electron units");
    RIP_REQUIRE(!electrons_out || Sref, "rip_il_apply_planes: Sref needed for electrons_out");
    use_device(device);
    const long npix = (long)ny * nx;
    DevRaw dc, ds, dk, dg, tmp;
    DevBuf<float> dco, dmin, dmax, dref;
    DevBuf<double> dout(npix);
    dc.upload(counts, npix * dtype_size(c_dtype));
    if (start_e) ds.upload(start_e, npix * dtype_size(s_dtype));
    if (kernel) dk.upload(kernel, 9 * npix * dtype_size(k_dtype));
    if (gain) dg.upload(gain, npix * dtype_size(g_dtype));
    dco.upload(coefs, (size_t)P * npix);
    dmin.upload(Smin, npix);
    dmax.upload(Smax, npix);
    if (Sref) dref.upload(Sref, npix);
    tmp.alloc(npix * 8);
    IlWindow w;
    w.ny = ny; w.nx = nx; w.K = kernel ? dk.p : nullptr; w.k_dtype = k_dtype;
    w.gain = gain ? dg.p : nullptr; w.g_dtype = gain ? g_dtype : RIP_F32; w.gain_off = 0; w.gain_pitch = nx;
    w.coefs = dco.p; w.Smin = dmin.p; w.Smax = dmax.p; w.Sref = dref.p; w.lin_plane = npix; w.lin_off = 0; w.lin_pitch = nx; w.P = P;
    const int tag = il_apply_device(dc.p, c_dtype, start_e ? ds.p : nullptr, s_dtype, start_scalar, w, electrons, electrons_out, dout.p, tmp.p, 0);
    dout.download(out, npix);
    if (out_dtype) *out_dtype = tag;
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

// ---- handle-based entry points ----------------------------------------------------------------------------
#include "rip_handle.h"

extern "C" int rip_il_apply(rip_caldir* h, const void* counts, int c_dtype, const float* start_e, int electrons,
                            int electrons_out, double* out) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && counts && out, "rip_il_apply: null argument");
    RIP_REQUIRE(c_dtype == RIP_F32 || c_dtype == RIP_F64 || c_dtype == RIP_I32, "rip_il_apply: counts dtype");
    use_device(h->device);
    cudaStream_t st = h->stream;
    const int na = h->na, n = h->n, nb = h->nb;
    const long npa = (long)na * na;
    DevRaw dc, ds, tmp;
    DevBuf<double> dout(npa);
    dc.upload(counts, npa * dtype_size(c_dtype), st);
    if (start_e) ds.upload(start_e, npa * 4, st);
    tmp.alloc(npa * 8);
    IlWindow w;
    w.ny = na; w.nx = na; w.K = h->has_ipc ? h->ipc.p : nullptr; w.k_dtype = h->d.ipc_dtype;
    w.gain = h->gain.p; w.g_dtype = h->d.gain_dtype; w.gain_off = (long)nb * n + nb; w.gain_pitch = n;
    w.coefs = h->coefs.p; w.Smin = h->Smin.p; w.Smax = h->Smax.p; w.Sref = h->Sref.p;
    w.lin_plane = (long)n * n; w.lin_off = (long)nb * n + nb; w.lin_pitch = n; w.P = h->P;
    il_apply_device(dc.p, c_dtype, start_e ? ds.p : nullptr, RIP_F32, 0.0, w, electrons, electrons_out, dout.p, tmp.p, st);
    dout.download(out, npa, st);
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}

static void make_l1_impl(rip_caldir* h, const int32_t* d_counts, const int32_t* d_cum, const rip_fwd_params* prm,
                         float* d_out, cudaStream_t st) {
    RIP_REQUIRE(prm->G >= 1 && prm->G <= RIP_GMAX, "rip_make_l1: G=%d outside 1..%d", prm->G, RIP_GMAX);
    RIP_REQUIRE(prm->n_reads >= 1 && prm->n_reads <= 64, "rip_make_l1: n_reads=%d outside 1..64", prm->n_reads);
    int tot = 0;
    for (int g = 0; g < prm->G; ++g) { RIP_REQUIRE(prm->reads_per_group[g] >= 1, "rip_make_l1: empty group"); tot += prm->reads_per_group[g]; }
    RIP_REQUIRE(tot == prm->n_reads, "rip_make_l1: reads_per_group sums to %d, n_reads=%d", tot, prm->n_reads);
    RIP_REQUIRE(!prm->add_reset_noise || h->resetnoise.p, "rip_make_l1: read file has no resetnoise plane");
    RIP_REQUIRE(!(prm->add_biascorr && h->has_bias) || h->d.n_bias >= prm->G, "rip_make_l1: biascorr has fewer groups than the read pattern");
    FwdArgs A;
    memset(&A, 0, sizeof A);
    A.n = h->n; A.nb = h->nb; A.na = h->na; A.G = prm->G; A.P = h->P; A.n_reads = prm->n_reads;
    for (int g = 0; g < RIP_GMAX; ++g) A.reads_per_group[g] = prm->reads_per_group[g];
    for (int k = 0; k < 64; ++k) A.read_index[k] = prm->read_index[k];
    A.read_time = prm->read_time; A.seed = prm->seed;
    A.add_read_noise = prm->add_read_noise; A.add_reset_noise = prm->add_reset_noise;
    A.add_biascorr = prm->add_biascorr; A.quantize = prm->quantize;
    A.biascorr_t0 = h->d.biascorr_t0; A.has_bias = h->has_bias ? 1 : 0;
    const size_t npa = (size_t)h->na * h->na, npl = (size_t)h->n * h->n;
    // per-CALDIR certificate of the fast inverse (once) and the per-call workspace
    if (!h->lin_A.p) {
        h->lin_A.alloc(npl);
        h->lin_m.alloc(npl);
        RIP_LAUNCH(invlin_certify_kernel, (unsigned)((npl + 127) / 128), 128, 0, st, (const float*)h->coefs.p, h->P, (long)npl,
                   h->lin_A.p, h->lin_m.p);
    }
    if (h->f_start.n < npa) h->f_start.alloc(npa);
    if (!d_cum && h->f_cum.n < npa * prm->n_reads) h->f_cum.alloc(npa * prm->n_reads);
    {   // romanisim starts the clock at the reset: read k draws Binomial(remaining, (t_k - t_prev)/(t_last - t_prev))
        const double t_last = prm->read_time * (double)prm->read_index[prm->n_reads - 1];
        double t_prev = 0.0;
        for (int k = 0; k < prm->n_reads; ++k) {
            const double t = prm->read_time * (double)prm->read_index[k];
            ReadTab& T = A.rtab[k];
            T.draw = t > t_prev ? 1 : 0;
            double p = (t_last > t_prev) ? (t - t_prev) / (t_last - t_prev) : 1.0;
            p = p >= 1.0 ? 1.0 : p;
            T.p = p;
            T.flip = p > 0.5 ? 1 : 0;
            const double q = T.flip ? 1.0 - p : p;
            T.q = (float)q; T.omq = 1.0f - T.q; T.l1mq = log1pf(-T.q); T.s = T.q / T.omq;
            t_prev = t;
            A.extrap[k] = 0.0f;
            if (k >= 2) {
                const double t1 = prm->read_time * (double)prm->read_index[k - 1], t2 = prm->read_time * (double)prm->read_index[k - 2];
                if (t1 > t2 && t > t1) A.extrap[k] = (float)((t - t1) / (t1 - t2));
            }
        }
    }
    A.counts = d_counts;
    if (d_cum && prm->cr_enable) {  // cosmic rays add to the cumulative counts: work on a copy of the caller's cube
        if (h->f_cum.n < npa * prm->n_reads) h->f_cum.alloc(npa * prm->n_reads);
        RIP_CUDA(cudaMemcpyAsync(h->f_cum.p, d_cum, npa * prm->n_reads * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        d_cum = h->f_cum.p;
    }
    A.cum = d_cum ? const_cast<int32_t*>(d_cum) : h->f_cum.p;
    A.cum_given = d_cum ? 1 : 0;
    A.start = h->f_start.p;
    A.linA = h->lin_A.p; A.linm = h->lin_m.p;
    A.coefs = h->coefs.p; A.Smin = h->Smin.p; A.Smax = h->Smax.p;
    A.gain = h->gain.p; A.ipc = h->has_ipc ? h->ipc.p : nullptr; A.read = h->read.p; A.resetnoise = h->resetnoise.p;
    A.dark_slope = h->dark_slope.p;
    // sim_to_isim.py:256-258 adds the whole biascorr cube (same number of groups as the pattern)
    A.bias = h->has_bias ? h->biascorr.p + (size_t)(h->d.n_bias - prm->G) * h->na * h->na : nullptr;
    A.out = d_out;
    const bool gd = h->d.gain_dtype == RIP_F64, kd = h->has_ipc && h->d.ipc_dtype == RIP_F64;
    dim3 ga((h->na + 127) / 128, h->na);
    if (gd) RIP_LAUNCH(fwd_apportion_kernel<double>, ga, 128, 0, st, A);
    else RIP_LAUNCH(fwd_apportion_kernel<float>, ga, 128, 0, st, A);
    if (prm->cr_enable) {  // cosmic rays on top of the apportioned (or supplied) cumulative counts
        RIP_REQUIRE(prm->cr_flux >= 0.0 && prm->cr_area >= 0.0 && prm->cr_conversion_factor > 0.0 && prm->cr_pixel_size > 0.0 &&
                    prm->cr_pixel_depth >= 0.0, "rip_make_l1: bad cosmic-ray parameters");
        if (!h->cr_len_cdf.p) {
            std::vector<float> a, b;
            cr_tables(a, b);
            h->cr_len_cdf.upload(a.data(), a.size(), st);
            h->cr_dedx_cdf.upload(b.data(), b.size(), st);
            RIP_CUDA(cudaStreamSynchronize(st));  // (the host vectors die here)
        }
        if (h->f_crg.n < npa) h->f_crg.alloc(npa);
        RIP_CUDA(cudaMemsetAsync(h->f_crg.p, 0, npa * sizeof(uint32_t), st));
        CrArgs Cr;
        memset(&Cr, 0, sizeof Cr);
        Cr.na = h->na; Cr.n_reads = prm->n_reads; Cr.seed = prm->seed;
        double t_prev = 0.0;
        int k = 0;
        for (int g = 0; g < prm->G; ++g)
            for (int r = 0; r < prm->reads_per_group[g]; ++r, ++k) {
                const double t = prm->read_time * (double)prm->read_index[k];
                Cr.lam[k] = t > t_prev ? prm->cr_flux * prm->cr_area * (t - t_prev) : 0.0;
                Cr.group_of_read[k] = g;
                t_prev = t;
            }
        Cr.counts_per_pix = prm->cr_pixel_size / prm->cr_conversion_factor;
        Cr.depth_ratio = prm->cr_pixel_depth / prm->cr_pixel_size;
        Cr.inv_pixel_size = 1.0 / prm->cr_pixel_size;
        Cr.len_cdf = h->cr_len_cdf.p; Cr.dedx_cdf = h->cr_dedx_cdf.p;
        Cr.cum = A.cum; Cr.groups = h->f_crg.p;
        RIP_LAUNCH(fwd_cr_kernel, prm->n_reads, 128, 0, st, Cr);
        h->f_crg_valid = true;
    } else {
        h->f_crg_valid = false;
    }
    dim3 grid((h->na + FTX - 1) / FTX, (h->na + FTY - 1) / FTY);
    const int threads = FTX * FTY;
#define FW(PM)                                                                                   \
    do {                                                                                         \
        if (!gd && !kd) RIP_LAUNCH((fwd_ramp_kernel<PM, float, float>), grid, threads, 0, st, A);  \
        else if (gd && !kd) RIP_LAUNCH((fwd_ramp_kernel<PM, double, float>), grid, threads, 0, st, A); \
        else if (!gd && kd) RIP_LAUNCH((fwd_ramp_kernel<PM, float, double>), grid, threads, 0, st, A); \
        else RIP_LAUNCH((fwd_ramp_kernel<PM, double, double>), grid, threads, 0, st, A);           \
    } while (0)
    if (h->P <= 4) FW(4);
    else if (h->P <= 11) FW(11);
    else FW(RIP_PMAX);
#undef FW
}

extern "C" int rip_make_l1_dev(rip_caldir* h, const int32_t* d_counts, const rip_fwd_params* prm, float* d_resultants,
                               void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_counts && prm && d_resultants, "rip_make_l1_dev: null argument");
    use_device(h->device);
    make_l1_impl(h, d_counts, nullptr, prm, d_resultants, (cudaStream_t)stream);
    RIP_API_END
}

extern "C" int rip_make_l1_host(rip_caldir* h, const int32_t* counts, const int32_t* cum_counts,
                                const rip_fwd_params* prm, float* resultants) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && (counts || cum_counts) && prm && resultants, "rip_make_l1_host: null argument");
    use_device(h->device);
    cudaStream_t st = h->stream;
    const size_t npa = (size_t)h->na * h->na;
    DevBuf<int32_t> dc, dcum;
    if (counts) dc.upload(counts, npa, st);
    if (cum_counts) dcum.upload(cum_counts, npa * prm->n_reads, st);
    DevBuf<float> dout(npa * prm->G);
    make_l1_impl(h, counts ? dc.p : nullptr, cum_counts ? dcum.p : nullptr, prm, dout.p, st);
    dout.download(resultants, npa * prm->G, st);
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}

// bit g of plane [na,na] = a cosmic ray deposited electrons during group g of the LAST rip_make_l1_* call with
// cr_enable (the per-group JUMP_DET flag romanisim's apportioning returns in its dq cube)
extern "C" int rip_fwd_cr_groups_host(rip_caldir* h, uint32_t* groups) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && groups, "rip_fwd_cr_groups_host: null argument");
    RIP_REQUIRE(h->f_crg_valid, "rip_fwd_cr_groups_host: the last forward ramp of this handle ran without cosmic rays");
    use_device(h->device);
    h->f_crg.download(groups, (size_t)h->na * h->na, h->stream);
    RIP_CUDA(cudaStreamSynchronize(h->stream));
    RIP_API_END
}

extern "C" int rip_fwd_cr_groups_dev(rip_caldir* h, const uint32_t** d_groups) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_groups, "rip_fwd_cr_groups_dev: null argument");
    RIP_REQUIRE(h->f_crg_valid, "rip_fwd_cr_groups_dev: the last forward ramp of this handle ran without cosmic rays");
    *d_groups = h->f_crg.p;
    RIP_API_END
}

// cumulative electrons per read [n_reads,na,na] of the LAST forward ramp of the handle (binomial apportioning + cosmic
// rays): for the statistical tests of the samplers
extern "C" int rip_fwd_cum_counts_host(rip_caldir* h, int n_reads, int32_t* cum) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && cum && n_reads >= 1, "rip_fwd_cum_counts_host: bad argument");
    const size_t npa = (size_t)h->na * h->na;
    RIP_REQUIRE(h->f_cum.n >= npa * (size_t)n_reads, "rip_fwd_cum_counts_host: no forward ramp with %d reads has run on this handle", n_reads);
    use_device(h->device);
    h->f_cum.download(cum, npa * (size_t)n_reads, h->stream);
    RIP_CUDA(cudaStreamSynchronize(h->stream));
    RIP_API_END
}
