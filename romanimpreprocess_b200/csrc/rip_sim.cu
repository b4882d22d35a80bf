// Forward path, the rows either side of make_l1_fullcal (a15, a20) and the many-realisations
// bookkeeping (BASELINE configs[4]; reference validation_tests/many_realizations.py:58-106):
//
//   sim_calprep / sim_counts   Image2D.simulate (from_sim/sim_to_isim.py:615-648): calibration planes of the scene
//                              (IPC-deconvolved dark rate and flat, clips) and the Poisson draw of the scene electrons
//   noise_1f_frames            noise_1f_frame (sim_to_isim.py:265-303): 1/f noise blocks by a length-(n/4)^2 complex FFT
//                              (four-step, two shared-memory passes) of an on-the-fly Gaussian spectrum
//   fill_refdata_1f            fill_in_refdata_and_1f (sim_to_isim.py:306-402): reference pixels, banding, amp33
//   mask_build / moments_*     CombinedMask.build (utils/maskhandling.py:82-117) fused with the moment sums
//   stack_median               np.median(stack, axis=0) for the realisation stacks
//
// Random numbers are counter-based Philox (rip_rng.cuh), not GalSim's Boost-MT: everything stochastic is validated
// statistically against the oracle (tests/test_gpu_sim.py); every deterministic step is checked exactly.
#include <math.h>
#include <algorithm>
#include <stdint.h>

#include "rip_handle.h"
#include "rip_launch.h"
#include "rip_math.cuh"
#include "rip_rng.cuh"

namespace rip {

// this_dark (e/s) input of ipc_rev: dark_slope * gain on the full frame (sim_to_isim.py:624)
template <typename TG>
__global__ void dark_e_kernel(const float* __restrict__ dark_slope, const TG* __restrict__ gain, long npix, float* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    out[p] = (float)((typename Promote<float, TG>::type)dark_slope[p] * (typename Promote<float, TG>::type)gain[p]);
}

// clips of sim_to_isim.py:631-633 on the active window of full-frame planes -> dense [na,na] outputs
__global__ void sim_clip_kernel(const float* __restrict__ dark_full, const float* __restrict__ flat_full, int n, int nb,
                                float* __restrict__ this_dark, float* __restrict__ this_flat) {
    const int na = n - 2 * nb;
    const int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= na) return;
    const long q = (long)(ya + nb) * n + (xa + nb), p = (long)ya * na + xa;
    const float fl = np_clip<float>(flat_full[q], 0.0f, 1.99999952316284179688f);  // 2 - 2**-21
    const float lo = -0.1f * fl;  // python float * f32 array -> f32
    this_flat[p] = fl;
    this_dark[p] = np_max<float>(dark_full[q], lo);
}

// counts (+)= Poisson(clip(C t g / g_ideal * image * flat / area, 0)) [+ Poisson(dark * t_dark)]
template <typename TG, typename TA>
__global__ void sim_counts_kernel(const float* __restrict__ image, const float* __restrict__ this_flat,
                                  const float* __restrict__ this_dark, const TG* __restrict__ gain, const TA* __restrict__ area,
                                  int n, int nb, double ct, double g_ideal, double t_dark, uint64_t seed, int accumulate,
                                  int32_t* __restrict__ counts, double* __restrict__ rate_out) {
    const int na = n - 2 * nb;
    const int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= na) return;
    const long q = (long)(ya + nb) * n + (xa + nb), p = (long)ya * na + xa;
    const double fa = area ? (double)this_flat[p] / (double)area[p] : (double)this_flat[p];
    double lam = ct * (double)gain[q] / g_ideal * (double)image[p] * fa;
    lam = lam > 0.0 ? lam : 0.0;  // np.clip(., 0, None); NaN -> no electrons
    if (rate_out) rate_out[p] = lam;
    Philox rng;
    rng.init(seed, (uint64_t)p, 2u);
    long c = poisson_draw(rng, lam);
    if (t_dark > 0.0) {
        Philox r2;
        r2.init(seed, (uint64_t)p, 3u);
        c += poisson_draw(r2, (double)this_dark[p] * t_dark);
    }
    if (accumulate) c += counts[p];
    counts[p] = (int32_t)(c > 2147483647L ? 2147483647L : c);
}

// ---------------------------------------------------------------------------------------------------------
// 1/f frames.  m = 2 * nside * (nside/32) = L^2 with L = nside/4.  Four-step FFT:
//   X[k1 + L k2] = sum_{n2} W_L^{n2 k2} [ W_m^{n2 k1} sum_{n1} x[L n1 + n2] W_L^{n1 k1} ]
// pass 1: for 8 values of n2 per CTA, the L-point transforms over n1 and the twiddle -> A[k1][n2]
// pass 2: for 8 values of k1 per CTA, the L-point transforms over n2 -> real part of X[k] for k < m/2, / sqrt 2
// ---------------------------------------------------------------------------------------------------------
constexpr int FCOL = 8;

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place decimation-in-time transforms of FCOL columns of length L (input already bit-reversed).  Two radix-2 stages
// are done per pass on four elements held in registers (half the shared-memory traffic and barriers of stage-by-stage
// radix 2); an odd number of stages starts with one plain radix-2 pass.  tw[t] = exp(-2 pi i t / L).
__device__ void cta_fft(float2* s, const float2* tw, int L) {
    int len = 2;
    int logL = 0;
    while ((1 << logL) < L) ++logL;
    if (logL & 1) {  // stage len = 2: twiddle 1
        for (int b = threadIdx.x; b < FCOL * (L >> 1); b += blockDim.x) {
            const int col = b / (L >> 1), i = b - col * (L >> 1);
            float2* a = s + (long)col * L + 2 * i;
            const float2 u = a[0], v = a[1];
            a[0] = make_float2(u.x + v.x, u.y + v.y);
            a[1] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
        len = 4;
    }
    for (; len <= (L >> 1); len <<= 2) {  // stages len and 2 len together
        const int h = len >> 1, t1 = L / len, t2 = L / (2 * len);
        for (int b = threadIdx.x; b < FCOL * (L >> 2); b += blockDim.x) {
            const int col = b / (L >> 2), i = b - col * (L >> 2);
            const int grp = i / h, k = i - grp * h;
            float2* a = s + (long)col * L + grp * (2 * len) + k;
            const float2 w1 = tw[k * t1];
            float2 a0 = a[0], a1 = cmul(a[h], w1), a2 = a[2 * h], a3 = cmul(a[3 * h], w1);
            const float2 b0 = make_float2(a0.x + a1.x, a0.y + a1.y), b1 = make_float2(a0.x - a1.x, a0.y - a1.y);
            const float2 b2 = make_float2(a2.x + a3.x, a2.y + a3.y), b3 = make_float2(a2.x - a3.x, a2.y - a3.y);
            const float2 c2 = cmul(b2, tw[k * t2]), c3 = cmul(b3, tw[(k + h) * t2]);
            a[0] = make_float2(b0.x + c2.x, b0.y + c2.y);
            a[2 * h] = make_float2(b0.x - c2.x, b0.y - c2.y);
            a[h] = make_float2(b1.x + c3.x, b1.y + c3.y);
            a[3 * h] = make_float2(b1.x - c3.x, b1.y - c3.y);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void fill_twiddles(float2* tw, int L) {
    for (int t = threadIdx.x; t < (L >> 1); t += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * (float)t / (float)L, &sn, &cs);
        tw[t] = make_float2(cs, sn);
    }
}

// amplitude of spectrum element j (sim_to_isim.py:287-292): |k m|^-1/2 with k wrapped to [-1/2, 1/2), 0 at j = 0
__device__ __forceinline__ float amp_1f(unsigned j, unsigned m) {
    if (j == 0) return 0.0f;
    const unsigned d = j < m / 2 ? j : m - j;
    return rsqrtf((float)d);
}

__global__ void __launch_bounds__(256) fft1f_pass1_kernel(int L, int logL, uint64_t seed, unsigned frame0,
                                                          const double* __restrict__ draws, float2* __restrict__ A) {
    extern __shared__ float2 sm1[];
    float2* s = sm1;
    float2* tw = sm1 + (long)FCOL * L;
    const unsigned m = (unsigned)L * (unsigned)L;
    const int frame = blockIdx.y, n2_0 = blockIdx.x * FCOL;
    fill_twiddles(tw, L);
    for (int idx = threadIdx.x; idx < FCOL * L; idx += blockDim.x) {
        const int c = idx % FCOL, n1 = idx / FCOL;
        const unsigned j = (unsigned)L * (unsigned)n1 + (unsigned)(n2_0 + c);
        float re, im;
        if (draws) {
            re = (float)draws[(long)frame * 2 * m + j];
            im = (float)draws[(long)frame * 2 * m + m + j];
        } else {
            Philox rng;
            rng.init(seed, (uint64_t)j, 1024u + frame0 + (unsigned)frame);
            rng.normal2(re, im);
        }
        const float a = amp_1f(j, m);
        s[(long)c * L + bitrev((unsigned)n1, logL)] = make_float2(re * a, im * a);
    }
    __syncthreads();
    cta_fft(s, tw, L);
    for (int idx = threadIdx.x; idx < FCOL * L; idx += blockDim.x) {
        const int c = idx % FCOL, k1 = idx / FCOL;
        const unsigned n2 = (unsigned)(n2_0 + c);
        float sn, cs;
        sincospif(-2.0f * (float)(n2 * (unsigned)k1) / (float)m, &sn, &cs);  // n2 k1 < 2^20: exact
        A[((long)frame * L + k1) * L + n2] = cmul(s[(long)c * L + k1], make_float2(cs, sn));
    }
}

__global__ void __launch_bounds__(256) fft1f_pass2_kernel(int L, int logL, const float2* __restrict__ A,
                                                          float* __restrict__ frames, double* __restrict__ sums) {
    extern __shared__ float2 sm1[];
    float2* s = sm1;
    float2* tw = sm1 + (long)FCOL * L;
    __shared__ double red[256];
    const long m = (long)L * L;
    const int frame = blockIdx.y, k1_0 = blockIdx.x * FCOL;
    fill_twiddles(tw, L);
    for (int idx = threadIdx.x; idx < FCOL * L; idx += blockDim.x) {
        const int r = idx / L, n2 = idx - r * L;
        s[(long)r * L + bitrev((unsigned)n2, logL)] = A[((long)frame * L + (k1_0 + r)) * L + n2];
    }
    __syncthreads();
    cta_fft(s, tw, L);
    double acc = 0.0;
    for (int idx = threadIdx.x; idx < FCOL * (L >> 1); idx += blockDim.x) {
        const int r = idx % FCOL, k2 = idx / FCOL;
        const float v = s[(long)r * L + k2].x * 0.70710678118654752440f;
        frames[(long)frame * (m / 2) + (long)(k1_0 + r) + (long)L * k2] = v;
        acc += (double)v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(&sums[frame], red[0]);
}

__global__ void frame_demean_kernel(float* __restrict__ frames, const double* __restrict__ sums, long half) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.y;
    if (p >= half) return;
    frames[(long)frame * half + p] = (float)((double)frames[(long)frame * half + p] - sums[frame] / (double)half);
}

static int ilog2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return ((1 << l) == v) ? l : -1;
}

// nframes blocks [nside, nside/32] f32 into d_frames; d_work >= nframes * (nside/4)^2 float2; d_sums >= nframes doubles
static void noise_1f_frames_impl(int nside, int nframes, uint64_t seed, unsigned frame0, const double* d_draws,
                                 float* d_frames, float2* d_work, double* d_sums, cudaStream_t st) {
    const int L = nside / 4, logL = ilog2_exact(L);
    RIP_REQUIRE(logL >= 3 && L <= 1024 && nside == 4 * L, "noise_1f_frames: frame side %d must be a power of two in 32..4096", nside);
    const size_t smem = ((size_t)FCOL * L + L / 2) * sizeof(float2);
    configure_smem_once((const void*)fft1f_pass1_kernel, (FCOL * 1024 + 512) * 8);
    configure_smem_once((const void*)fft1f_pass2_kernel, (FCOL * 1024 + 512) * 8);
    const long half = (long)L * L / 2;
    RIP_CUDA(cudaMemsetAsync(d_sums, 0, (size_t)nframes * sizeof(double), st));
    dim3 grid(L / FCOL, nframes);
    RIP_LAUNCH(fft1f_pass1_kernel, grid, 256, smem, st, L, logL, seed, frame0, d_draws, d_work);
    RIP_LAUNCH(fft1f_pass2_kernel, grid, 256, smem, st, L, logL, (const float2*)d_work, d_frames, d_sums);
    dim3 g2((unsigned)((half + 255) / 256), nframes);
    RIP_LAUNCH(frame_demean_kernel, g2, 256, 0, st, d_frames, (const double*)d_sums, half);
}

// ---------------------------------------------------------------------------------------------------------
// fill_in_refdata_and_1f: one thread per (group, pixel)
// ---------------------------------------------------------------------------------------------------------
struct FillArgs {
    int n, nb, G, cw, banding;
    uint64_t seed;
    float inv_rn[RIP_GMAX];  // sqrt(len(tij[j])) as f32 (python float ** 0.5, weak scalar -> f32)
    float u_pink, c_pink;
    const float* read; const float* resetnoise; const float* dark;  // dark: last G groups of the dark cube
    const float* frames;  // [33][n, cw] of the current group: 0 = common, 1 + ch = channel ch
    uint16_t* im;
};

__global__ void fill_refdata_kernel(const FillArgs A, int g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int n = A.n, nb = A.nb;
    if (x >= n) return;
    const long p = (long)y * n + x, npl = (long)n * n;
    uint16_t* dst = A.im + (long)g * npl + p;
    float v;
    if (x >= nb && x < n - nb && y >= nb && y < n - nb) {
        v = (float)*dst;  // active pixels keep the simulated signal (sim_to_isim.py:356-358)
    } else {
        Philox r1, r2;
        r1.init(A.seed, (uint64_t)p, 32u + (unsigned)g);
        r2.init(A.seed, (uint64_t)p, 8u);
        float a = r1.normal() * A.read[p];
        a = a / A.inv_rn[g];
        const float b = r2.normal() * A.resetnoise[p];
        v = (a + b) + A.dark[(long)g * npl + p];
    }
    if (A.banding) {
        const int cw = A.cw, ch = x / cw, xin = x - ch * cw;
        const int sc = (ch & 1) ? (cw - 1 - xin) : xin;  // odd channels are read out mirrored (:384-385)
        const long fp = (long)y * cw + sc, fsz = (long)n * cw;
        const float common = A.frames[fp] * A.c_pink;
        const float pink = A.frames[(long)(1 + ch) * fsz + fp] * A.u_pink + common;
        v = v + pink / A.inv_rn[g];
    }
    v = rintf(v);
    v = v < 0.0f ? 0.0f : (v > 65535.0f ? 65535.0f : v);
    *dst = (uint16_t)v;
}

// amp33[g] = u16(med + (N std + (RU_PINK frame33 + M_PINK common)) / rn)   (sim_to_isim.py:392-399)
__global__ void fill_amp33_kernel(int n, int cw, int g, uint64_t seed, float rn, float c_pink, float ru_pink, float m_pink,
                                  const float* __restrict__ med, const float* __restrict__ stdv, const float* __restrict__ common_f,
                                  const float* __restrict__ frame33, int banding, uint16_t* __restrict__ amp33) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x, fsz = (long)n * cw;
    if (p >= fsz) return;
    Philox r;
    r.init(seed, (uint64_t)p, 48u + (unsigned)g);
    const long q = (p / cw) * 128 + (p % cw);  // statistics planes are [n,128]
    const float white = r.normal() * stdv[q];
    const float common = common_f[p] * c_pink;
    const float pink = ru_pink * frame33[p] + m_pink * common;
    const float v = med[q] + (white + pink) / rn;
    amp33[(long)g * fsz + p] = (uint16_t)(int)v;  // ndarray.astype(uint16): truncation
    (void)banding;
}

// ---------------------------------------------------------------------------------------------------------
// CombinedMask.build (maskhandling.py:82-117) on a window, fused with the moment sums of many_realizations.py:74-77
// ---------------------------------------------------------------------------------------------------------
struct GrowSets { uint32_t m1, m5, m9, m25; };

__device__ __forceinline__ bool grown_mask(const uint32_t* __restrict__ dq, long pitch, int ny, int nx, int y, int x, const GrowSets S) {
    const uint32_t any = S.m1 | S.m5 | S.m9 | S.m25;
    uint32_t hit = dq[(long)y * pitch + x] & any;
    // scipy.signal.convolve(mode="same") pads with zeros: neighbours outside the window do not contribute
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= ny) continue;
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= nx || (dx == 0 && dy == 0)) continue;
            const int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
            uint32_t sel = S.m25;
            if (ady <= 1 && adx <= 1) sel |= S.m9;
            if (ady + adx == 1) sel |= S.m5;
            hit |= dq[(long)yy * pitch + xx] & sel;
        }
    }
    return hit != 0u;
}

__global__ void mask_build_kernel(const uint32_t* __restrict__ dq, long pitch, int ny, int nx, const GrowSets S, uint8_t* __restrict__ mask) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    mask[(long)y * nx + x] = grown_mask(dq, pitch, ny, nx, y, x, S) ? 1 : 0;
}

// moments[0] += w; moments[1] += w ? data : 0; moments[2] += w ? data**2 : 0      (float32 accumulators)
__global__ void moments_accumulate_kernel(const float* __restrict__ data, const uint32_t* __restrict__ dq, long pitch, int ny,
                                          int nx, const GrowSets S, float* __restrict__ mom) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    const bool w = !grown_mask(dq, pitch, ny, nx, y, x, S);
    const long p = (long)y * nx + x, np_ = (long)ny * nx;
    const float d = data[(long)y * pitch + x];
    mom[p] = mom[p] + (w ? 1.0f : 0.0f);
    mom[np_ + p] = mom[np_ + p] + (w ? d : 0.0f);
    mom[2 * np_ + p] = mom[2 * np_ + p] + (w ? d * d : 0.0f);
}

// many_realizations.py:80-83: mean, std, sentinel -1000 where nothing was accumulated
__global__ void moments_finalize_kernel(float* __restrict__ mom, long npix) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float m0 = mom[p];
    const float den = m0 + 1e-25f;
    float m1 = mom[npix + p] / den, m2 = mom[2 * npix + p] / den;
    m2 = sqrtf(np_max<float>(m2 - m1 * m1, 0.0f));  // np.clip(., 0, None): NaN propagates
    const bool ok = m0 > 0.1f;
    mom[npix + p] = ok ? m1 : -1000.0f;
    mom[2 * npix + p] = ok ? m2 : -1000.0f;
}

// np.median over the leading axis of a [R, npix] stack (R <= RP, RP a power of two): bitonic network in registers
template <int RP>
__global__ void stack_median_kernel(const float* __restrict__ stack, int R, long npix, float* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float v[RP];
    bool nan = false;
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        v[i] = (i < R) ? stack[(long)i * npix + p] : INFINITY;
        nan = nan || (v[i] != v[i]);
    }
#pragma unroll
    for (int k = 2; k <= RP; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < RP; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = ((i & k) == 0);
                    const float a = v[i], b = v[l];
                    const bool sw = up ? (a > b) : (a < b);
                    v[i] = sw ? b : a;
                    v[l] = sw ? a : b;
                }
            }
        }
    }
    // v ascending; padded +inf entries sit at the top.  Dynamic index on a register array would spill: select by scan.
    const int lo = (R - 1) >> 1, hi = R >> 1;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        a = (i == lo) ? v[i] : a;
        b = (i == hi) ? v[i] : b;
    }
    float med = (lo == hi) ? a : (a + b) / 2.0f;  // np.mean of the two middle values in float32
    out[p] = nan ? NAN : med;
}


// romanisim.l1.make_asdf restated: resultants f32 [G,na,na] -> active window of the u16 cube [G,n,n]
__global__ void l1_embed_kernel(const float* __restrict__ res, int G, int n, int nb, uint16_t* __restrict__ im) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, g = blockIdx.z;
    if (x >= n) return;
    const int na = n - 2 * nb;
    uint16_t v = 0;
    if (x >= nb && x < n - nb && y >= nb && y < n - nb) {
        float f = res[((long)g * na + (y - nb)) * na + (x - nb)];
        f = f < 0.0f ? 0.0f : (f > 65535.0f ? 65535.0f : f);  // NaN -> 0
        v = (uint16_t)(int)f;
    }
    im[((long)g * n + y) * n + x] = v;
}

// one realisation's planes of validation_tests/many_realizations.py:69-73 (full frames, zero border)
__global__ void realization_record_kernel(const uint16_t* __restrict__ im, int G, int n, int nb, const float* __restrict__ slope,
                                          const float* __restrict__ er, const float* __restrict__ ep, float* __restrict__ diffs,
                                          float* __restrict__ images, float* __restrict__ err) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n) return;
    const long p = (long)y * n + x, npl = (long)n * n;
    diffs[p] = (float)im[(long)(G - 1) * npl + p] - (float)im[npl + p];
    const bool act = x >= nb && x < n - nb && y >= nb && y < n - nb;
    images[p] = act ? slope[p] : 0.0f;
    err[p] = act ? sqrtf(er[p] * er[p] + ep[p] * ep[p]) : 0.0f;
}

static GrowSets grow_sets(const uint8_t* grow32) {
    GrowSets S{0u, 0u, 0u, 0u};
    for (int b = 0; b < 32; ++b) {
        const uint32_t bit = 1u << b;
        switch (grow32[b]) {
            case 0: break;
            case 1: S.m1 |= bit; break;
            case 5: S.m5 |= bit; break;
            case 9: S.m9 |= bit; break;
            case 25: S.m25 |= bit; break;
            default: RIP_REQUIRE(false, "mask grow code %d for bit %d (expected 0, 1, 5, 9 or 25)", (int)grow32[b], b);
        }
    }
    return S;
}

static void sim_calprep_impl(rip_caldir* h, float* d_this_dark, float* d_this_flat, cudaStream_t st) {
    const int n = h->n, nb = h->nb, na = h->na;
    const long npl = (long)n * n;
    DevBuf<float> dk(npl), fl(npl);
    DevRaw tmp;
    tmp.alloc((size_t)na * na * 8);
    const unsigned nblk = (unsigned)((npl + 255) / 256);
    if (h->d.gain_dtype == RIP_F64) RIP_LAUNCH(dark_e_kernel<double>, nblk, 256, 0, st, h->dark_slope.p, (const double*)h->gain.p, npl, dk.p);
    else RIP_LAUNCH(dark_e_kernel<float>, nblk, 256, 0, st, h->dark_slope.p, (const float*)h->gain.p, npl, dk.p);
    RIP_CUDA(cudaMemcpyAsync(fl.p, h->flat.p, npl * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (h->has_ipc) {
        launch_ipc_rev_dn(dk.p, n, n, nb, h->ipc.p, h->d.ipc_dtype, nullptr, RIP_F32, false, 0.f, tmp.p, st);
        launch_ipc_rev_dn(fl.p, n, n, nb, h->ipc.p, h->d.ipc_dtype, h->gain.p, h->d.gain_dtype, false, 0.f, tmp.p, st);
    }
    dim3 grid((na + 127) / 128, na);
    RIP_LAUNCH(sim_clip_kernel, grid, 128, 0, st, (const float*)dk.p, (const float*)fl.p, n, nb, d_this_dark, d_this_flat);
    RIP_CUDA(cudaStreamSynchronize(st));
}

static void ensure_sim_planes(rip_caldir* h) {
    if (h->sim_dark.p) return;
    const long npa = (long)h->na * h->na;
    h->sim_dark.alloc(npa);
    h->sim_flat.alloc(npa);
    sim_calprep_impl(h, h->sim_dark.p, h->sim_flat.p, h->stream);
}

static void sim_counts_impl(rip_caldir* h, const float* d_image, const void* d_area, int area_dtype, double t_exp,
                            double cnorm, double g_ideal, double t_dark, uint64_t seed, int32_t* d_counts, int accumulate,
                            double* d_rate, cudaStream_t st) {
    ensure_sim_planes(h);
    const int n = h->n, nb = h->nb, na = h->na;
    dim3 grid((na + 127) / 128, na);
    const double ct = cnorm * t_exp;
    const bool gd = h->d.gain_dtype == RIP_F64, ad = d_area && area_dtype == RIP_F64;
#define SC(TG, TA)                                                                                                     \
    RIP_LAUNCH((sim_counts_kernel<TG, TA>), grid, 128, 0, st, d_image, (const float*)h->sim_flat.p, (const float*)h->sim_dark.p, \
               (const TG*)h->gain.p, (const TA*)d_area, n, nb, ct, g_ideal, t_dark, seed, accumulate, d_counts, d_rate)
    if (gd) { if (ad) SC(double, double); else SC(double, float); }
    else { if (ad) SC(float, double); else SC(float, float); }
#undef SC
}

static void fill_refdata_impl(rip_caldir* h, uint16_t* d_im, uint16_t* d_amp33, int G, const int32_t* reads_per_group,
                              uint64_t seed, int banding, cudaStream_t st) {
    const int n = h->n, nb = h->nb, cw = n / 32;
    RIP_REQUIRE(G >= 1 && G <= RIP_GMAX, "rip_fill_refdata_1f: G=%d outside 1..%d", G, RIP_GMAX);
    RIP_REQUIRE(h->resetnoise.p, "rip_fill_refdata_1f: read file has no resetnoise plane");
    RIP_REQUIRE(h->d.n_dark >= G && h->dark_cube.p, "rip_fill_refdata_1f: dark cube has %d groups, need %d", h->d.n_dark, G);
    const bool want33 = d_amp33 && h->has_amp33 && h->amp_std.p;
    RIP_REQUIRE(!want33 || cw <= 128, "rip_fill_refdata_1f: amp33 statistics are [n,128]; frame side %d has %d-column channels", n, cw);
    const long npl = (long)n * n, fsz = (long)n * cw;
    const int nfr = 34;
    if (banding) {
        const int L = n / 4;
        if (h->f_frames.n < (size_t)nfr * fsz) h->f_frames.alloc((size_t)nfr * fsz);
        if (h->f_work.n < (size_t)nfr * L * L * 2) h->f_work.alloc((size_t)nfr * L * L * 2);
        if (h->f_sums.n < (size_t)nfr) h->f_sums.alloc(nfr);
    }
    FillArgs A;
    memset(&A, 0, sizeof A);
    A.n = n; A.nb = nb; A.G = G; A.cw = cw; A.banding = banding; A.seed = seed;
    for (int g = 0; g < G; ++g) A.inv_rn[g] = (float)sqrt((double)reads_per_group[g]);
    A.u_pink = (float)h->d.u_pink; A.c_pink = (float)h->d.c_pink;
    A.read = h->read.p; A.resetnoise = h->resetnoise.p;
    A.dark = h->dark_cube.p + (size_t)(h->d.n_dark - G) * npl;
    A.frames = h->f_frames.p; A.im = d_im;
    dim3 grid((n + 127) / 128, n);
    for (int g = 0; g < G; ++g) {
        if (banding)
            noise_1f_frames_impl(n, want33 ? 34 : 33, seed, (unsigned)(g * 64), nullptr, h->f_frames.p, (float2*)h->f_work.p,
                                 h->f_sums.p, st);
        RIP_LAUNCH(fill_refdata_kernel, grid, 128, 0, st, A, g);
        if (want33 && banding)
            RIP_LAUNCH(fill_amp33_kernel, (unsigned)((fsz + 255) / 256), 256, 0, st, n, cw, g, seed, A.inv_rn[g], A.c_pink,
                       (float)h->d.ru_pink, (float)h->d.m_pink, (const float*)h->amp_med.p, (const float*)h->amp_std.p,
                       (const float*)h->f_frames.p, (const float*)(h->f_frames.p + 33 * fsz), banding, d_amp33);
    }
}

template <int RP>
static void stack_median_t(const float* d_stack, int R, long npix, float* d_out, cudaStream_t st) {
    RIP_LAUNCH(stack_median_kernel<RP>, (unsigned)((npix + 127) / 128), 128, 0, st, d_stack, R, npix, d_out);
}

}  // namespace rip

using namespace rip;

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" int rip_sim_calprep(rip_caldir* h, float* this_dark, float* this_flat) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && this_dark && this_flat, "rip_sim_calprep: null argument");
    use_device(h->device);
    ensure_sim_planes(h);
    const long npa = (long)h->na * h->na;
    h->sim_dark.download(this_dark, npa, h->stream);
    h->sim_flat.download(this_flat, npa, h->stream);
    RIP_CUDA(cudaStreamSynchronize(h->stream));
    RIP_API_END
}

extern "C" int rip_sim_counts_dev(rip_caldir* h, const float* d_image, const void* d_area, int area_dtype, double t_exp,
                                  double cnorm, double g_ideal, double t_dark, uint64_t seed, int32_t* d_counts,
                                  int accumulate, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_image && d_counts, "rip_sim_counts_dev: null argument");
    use_device(h->device);
    sim_counts_impl(h, d_image, d_area, area_dtype, t_exp, cnorm, g_ideal, t_dark, seed, d_counts, accumulate, nullptr,
                    (cudaStream_t)stream);
    RIP_API_END
}

extern "C" int rip_sim_counts_host(rip_caldir* h, const float* image, const void* area, int area_dtype, double t_exp,
                                   double cnorm, double g_ideal, double t_dark, uint64_t seed, int32_t* counts,
                                   int accumulate, double* rate_out) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && image && counts, "rip_sim_counts_host: null argument");
    use_device(h->device);
    cudaStream_t st = h->stream;
    const long npa = (long)h->na * h->na;
    DevBuf<float> di;
    DevRaw da;
    DevBuf<int32_t> dc(npa);
    DevBuf<double> dr;
    di.upload(image, npa, st);
    if (area) da.upload(area, npa * dtype_size(area_dtype), st);
    if (accumulate) dc.upload(counts, npa, st);
    if (rate_out) dr.alloc(npa);
    sim_counts_impl(h, di.p, area ? da.p : nullptr, area_dtype, t_exp, cnorm, g_ideal, t_dark, seed, dc.p, accumulate,
                    rate_out ? dr.p : nullptr, st);
    dc.download(counts, npa, st);
    if (rate_out) dr.download(rate_out, npa, st);
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}

extern "C" int rip_noise_1f_frames_host(int device, int nside, int nframes, uint64_t seed, const double* draws, float* out) {
    RIP_API_BEGIN
    RIP_REQUIRE(out && nframes >= 1, "rip_noise_1f_frames_host: null argument");
    use_device(device);
    const int L = nside / 4;
    const long m = (long)L * L, half = m / 2;
    DevBuf<double> dd, sums(nframes);
    DevBuf<float> fr((size_t)nframes * half), work((size_t)nframes * m * 2);
    if (draws) dd.upload(draws, (size_t)nframes * 2 * m, 0);
    noise_1f_frames_impl(nside, nframes, seed, 0u, draws ? dd.p : nullptr, fr.p, (float2*)work.p, sums.p, 0);
    fr.download(out, (size_t)nframes * half, 0);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_fill_refdata_1f_dev(rip_caldir* h, uint16_t* d_im, uint16_t* d_amp33, int G, const int32_t* reads_per_group,
                                       uint64_t seed, int fill_in_banding, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_im && reads_per_group, "rip_fill_refdata_1f_dev: null argument");
    use_device(h->device);
    fill_refdata_impl(h, d_im, d_amp33, G, reads_per_group, seed, fill_in_banding, (cudaStream_t)stream);
    RIP_API_END
}

extern "C" int rip_fill_refdata_1f_host(rip_caldir* h, uint16_t* im, uint16_t* amp33, int G, const int32_t* reads_per_group,
                                        uint64_t seed, int fill_in_banding) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && im && reads_per_group, "rip_fill_refdata_1f_host: null argument");
    use_device(h->device);
    cudaStream_t st = h->stream;
    const long npl = (long)h->n * h->n, a33 = (long)h->n * (h->n / 32);
    DevBuf<uint16_t> di, da;
    di.upload(im, (size_t)G * npl, st);
    if (amp33) { da.alloc((size_t)G * a33); da.zero(st); }
    fill_refdata_impl(h, di.p, amp33 ? da.p : nullptr, G, reads_per_group, seed, fill_in_banding, st);
    di.download(im, (size_t)G * npl, st);
    if (amp33 && h->has_amp33 && h->amp_std.p && fill_in_banding) da.download(amp33, (size_t)G * a33, st);
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}

extern "C" int rip_mask_build_host(int device, const uint32_t* dq, int ny, int nx, const uint8_t* grow32, uint8_t* mask) {
    RIP_API_BEGIN
    RIP_REQUIRE(dq && grow32 && mask && ny > 0 && nx > 0, "rip_mask_build_host: null argument");
    use_device(device);
    const GrowSets S = grow_sets(grow32);
    const long npix = (long)ny * nx;
    DevBuf<uint32_t> d;
    DevBuf<uint8_t> m(npix);
    d.upload(dq, npix, 0);
    dim3 grid((nx + 127) / 128, ny);
    RIP_LAUNCH(mask_build_kernel, grid, 128, 0, 0, (const uint32_t*)d.p, (long)nx, ny, nx, S, m.p);
    m.download(mask, npix, 0);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_moments_accumulate_dev(int device, const float* d_slope, const uint32_t* d_pdq, int n, int nb,
                                          const uint8_t* grow32, float* d_moments, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_slope && d_pdq && grow32 && d_moments, "rip_moments_accumulate_dev: null argument");
    use_device(device);
    const GrowSets S = grow_sets(grow32);
    const int na = n - 2 * nb;
    const long off = (long)nb * n + nb;
    dim3 grid((na + 127) / 128, na);
    RIP_LAUNCH(moments_accumulate_kernel, grid, 128, 0, (cudaStream_t)stream, d_slope + off, d_pdq + off, (long)n, na, na, S, d_moments);
    RIP_API_END
}

extern "C" int rip_moments_finalize_dev(int device, float* d_moments, long npix, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_moments && npix > 0, "rip_moments_finalize_dev: null argument");
    use_device(device);
    RIP_LAUNCH(moments_finalize_kernel, (unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream, d_moments, npix);
    RIP_API_END
}

extern "C" int rip_stack_median_dev(int device, const float* d_stack, int R, long npix, float* d_out, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_stack && d_out && npix > 0, "rip_stack_median_dev: null argument");
    RIP_REQUIRE(R >= 1 && R <= 128, "rip_stack_median_dev: R=%d outside 1..128", R);
    use_device(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (R <= 8) stack_median_t<8>(d_stack, R, npix, d_out, st);
    else if (R <= 16) stack_median_t<16>(d_stack, R, npix, d_out, st);
    else if (R <= 32) stack_median_t<32>(d_stack, R, npix, d_out, st);
    else if (R <= 64) stack_median_t<64>(d_stack, R, npix, d_out, st);
    else stack_median_t<128>(d_stack, R, npix, d_out, st);
    RIP_API_END
}

extern "C" int rip_l1_embed_dev(int device, const float* d_resultants, int G, int n, int nb, uint16_t* d_im, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_resultants && d_im && G >= 1 && n > 2 * nb, "rip_l1_embed_dev: bad argument");
    use_device(device);
    dim3 grid((n + 127) / 128, n, G);
    RIP_LAUNCH(l1_embed_kernel, grid, 128, 0, (cudaStream_t)stream, d_resultants, G, n, nb, d_im);
    RIP_API_END
}

extern "C" int rip_realization_record_dev(int device, const uint16_t* d_im, int G, int n, int nb, const float* d_slope,
                                          const float* d_err_read, const float* d_err_poisson, float* d_diffs,
                                          float* d_images, float* d_err, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_im && d_slope && d_err_read && d_err_poisson && d_diffs && d_images && d_err && G >= 2,
                "rip_realization_record_dev: bad argument");
    use_device(device);
    dim3 grid((n + 127) / 128, n);
    RIP_LAUNCH(realization_record_kernel, grid, 128, 0, (cudaStream_t)stream, d_im, G, n, nb, d_slope, d_err_read,
               d_err_poisson, d_diffs, d_images, d_err);
    RIP_API_END
}

// =========================================================================================================
// Sky model (utils/sky.py:98-190 medfit): medians of N x N regions by radix select, polynomial evaluation
// =========================================================================================================
namespace rip {

__device__ __forceinline__ uint32_t f2key_sky(float v) {  // order-preserving key; NaNs are filtered before
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f_sky(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// np.nanmedian of one ky x kx region per CTA: four 8-bit radix passes over order-preserving keys, tracking the two
// middle ranks at once; float32 mean of the two middle values for even counts (np.median), NaN when nothing is finite.
__global__ void __launch_bounds__(1024) block_nanmedian_kernel(const float* __restrict__ arr, long pitch, int py, int px, int ky,
                                                               int kx, int N, float* __restrict__ meds) {
    __shared__ uint32_t hist[2][256];
    __shared__ uint32_t prefix[2], rank[2], mask_s, cnt_s;
    const int ry = blockIdx.y, rx = blockIdx.x;
    const float* base = arr + (long)(py + ry * ky) * pitch + (px + rx * kx);
    const long total = (long)ky * kx;
    if (threadIdx.x == 0) { prefix[0] = prefix[1] = 0u; rank[0] = rank[1] = 0u; mask_s = 0u; cnt_s = 0u; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 512; i += blockDim.x) (&hist[0][0])[i] = 0u;
        __syncthreads();
        const uint32_t m = mask_s, p0 = prefix[0], p1 = prefix[1];
        for (long e = threadIdx.x; e < total; e += blockDim.x) {
            const float v = base[(e / kx) * pitch + (e % kx)];
            if (v != v) continue;
            const uint32_t k = f2key_sky(v), b = (k >> shift) & 255u;
            if ((k & m) == p0) atomicAdd(&hist[0][b], 1u);
            if (pass > 0 && (k & m) == p1) atomicAdd(&hist[1][b], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (pass == 0) {
                uint32_t c = 0;
                for (int b = 0; b < 256; ++b) c += hist[0][b];
                cnt_s = c;
                rank[0] = c ? (c - 1) / 2 : 0u;
                rank[1] = c / 2;
                for (int b = 0; b < 256; ++b) hist[1][b] = hist[0][b];
            }
            for (int s = 0; s < 2; ++s) {
                uint32_t r = rank[s], acc = 0;
                int b = 0;
                for (; b < 255; ++b) {
                    if (acc + hist[s][b] > r) break;
                    acc += hist[s][b];
                }
                rank[s] = r - acc;
                prefix[s] |= (uint32_t)b << shift;
            }
            mask_s |= 255u << shift;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float out = NAN;
        if (cnt_s) {
            const float a = key2f_sky(prefix[0]), b = key2f_sky(prefix[1]);
            out = (cnt_s & 1u) ? a : (a + b) / 2.0f;
        }
        meds[ry * N + rx] = out;
    }
}

// arrmed = sum_k x[k] * outer(LPY[j], LPX[i]) in float64, terms in the reference's order, cast to float32; optionally
// subtracted from `arr` in place (float32 subtraction of the float32 model, as `slope -= skymodel` does)
__global__ void medfit_eval_kernel(int ny, int nx, int order, const double* __restrict__ coef, const double* __restrict__ LPX,
                                   const double* __restrict__ LPY, float* __restrict__ model, float* __restrict__ arr, long pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    double acc = 0.0;
    int k = 0;
    for (int i = 0; i <= order; ++i)
        for (int j = 0; j <= order - i; ++j) {
            acc = acc + coef[k] * (LPY[(long)j * ny + y] * LPX[(long)i * nx + x]);
            ++k;
        }
    const float mval = (float)acc;
    if (model) model[(long)y * nx + x] = mval;
    if (arr) arr[(long)y * pitch + x] = arr[(long)y * pitch + x] - mval;
}

}  // namespace rip

extern "C" int rip_block_nanmedian_dev(int device, const float* d_arr, long pitch, int ny, int nx, int N, float* d_meds,
                                       void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_arr && d_meds && N >= 1 && ny >= N && nx >= N, "rip_block_nanmedian_dev: bad argument");
    use_device(device);
    const int kx = nx / N, ky = ny / N, px = (nx % N) / 2, py = (ny % N) / 2;
    RIP_LAUNCH(block_nanmedian_kernel, dim3(N, N), 1024, 0, (cudaStream_t)stream, d_arr, pitch, py, px, ky, kx, N, d_meds);
    RIP_API_END
}

extern "C" int rip_medfit_eval_dev(int device, int ny, int nx, int order, const double* coef, const double* LPX,
                                   const double* LPY, float* d_model, float* d_arr, long pitch, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(coef && LPX && LPY && (d_model || d_arr) && order >= 0 && order <= 8, "rip_medfit_eval_dev: bad argument");
    use_device(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nc = (order + 1) * (order + 2) / 2;
    DevBuf<double> dc, dx, dy;
    dc.upload(coef, nc, st);
    dx.upload(LPX, (size_t)(order + 1) * nx, st);
    dy.upload(LPY, (size_t)(order + 1) * ny, st);
    dim3 grid((nx + 127) / 128, ny);
    RIP_LAUNCH(medfit_eval_kernel, grid, 128, 0, st, ny, nx, order, (const double*)dc.p, (const double*)dx.p,
               (const double*)dy.p, d_model, d_arr, pitch);
    RIP_CUDA(cudaStreamSynchronize(st));  // the coefficient tables are freed on return
    RIP_API_END
}

extern "C" int rip_medfit_host(int device, const float* arr, int ny, int nx, int N, float* meds) {
    RIP_API_BEGIN
    RIP_REQUIRE(arr && meds, "rip_medfit_host: null argument");
    use_device(device);
    DevBuf<float> d, m((size_t)N * N);
    d.upload(arr, (size_t)ny * nx, 0);
    const int kx = nx / N, ky = ny / N, px = (nx % N) / 2, py = (ny % N) / 2;
    RIP_REQUIRE(N >= 1 && kx >= 1 && ky >= 1, "rip_medfit_host: %d regions do not fit a %d x %d array", N, ny, nx);
    RIP_LAUNCH(block_nanmedian_kernel, dim3(N, N), 1024, 0, 0, (const float*)d.p, (long)nx, py, px, ky, kx, N, m.p);
    m.download(meds, (size_t)N * N, 0);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

// =========================================================================================================
// Noise layers (L1_to_L2/gen_noise_image.py:60-331 make_noise_cube), the device-side pieces of directive "R"
// =========================================================================================================
namespace rip {

// white read noise on the active pixels of a u16 cube (gen_noise_image.py:121-135):
//   im = N(0,1);  im *= read / sqrt(N_k)  (float32 array times a float64 quotient, stored float32);
//   resultants = float32(data) + im;  data = round(clip(resultants, 0, 65535))
__global__ void add_read_noise_kernel(uint16_t* __restrict__ data, int G, int n, int nb, const float* __restrict__ read,
                                      const double* __restrict__ sqrt_n, uint64_t seed) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x + nb, y = blockIdx.y + nb, k = blockIdx.z;
    if (x >= n - nb) return;
    const long p = (long)y * n + x;
    Philox r;
    r.init(seed, (uint64_t)p, 96u + (unsigned)k);
    float im = r.normal();
    im = (float)((double)im * ((double)read[p] / sqrt_n[k]));
    float res = (float)data[(long)k * n * n + p] + im;
    res = res < 0.0f ? 0.0f : (res > 65535.0f ? 65535.0f : res);
    data[(long)k * n * n + p] = (uint16_t)rintf(res);
}

// dark cube (last G groups) cast to the cube's integer type: gen_noise_image.py:101-109 (astype truncates)
__global__ void dark_to_u16_kernel(const float* __restrict__ dark, long count, uint16_t* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    const float v = dark[p];
    out[p] = (uint16_t)(int)(v < 0.0f ? 0.0f : (v > 65535.0f ? 65535.0f : v));
}

// diff = a - b on the active window of two full-frame float32 planes -> dense [na,na]
__global__ void active_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int nb, float* __restrict__ out) {
    const int na = n - 2 * nb;
    const int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= na) return;
    const long q = (long)(ya + nb) * n + (xa + nb);
    out[(long)ya * na + xa] = a[q] - b[q];
}

}  // namespace rip

extern "C" int rip_add_read_noise_dev(rip_caldir* h, uint16_t* d_data, int G, const int32_t* reads_per_group, uint64_t seed,
                                      void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_data && reads_per_group && G >= 1 && G <= RIP_GMAX, "rip_add_read_noise_dev: bad argument");
    use_device(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    double sq[RIP_GMAX];
    for (int g = 0; g < G; ++g) sq[g] = sqrt((double)reads_per_group[g]);
    if (h->f_sums.n < 64) h->f_sums.alloc(64);
    RIP_CUDA(cudaMemcpyAsync(h->f_sums.p + 40, sq, G * sizeof(double), cudaMemcpyHostToDevice, st));
    RIP_CUDA(cudaStreamSynchronize(st));  // (sq lives on this stack frame)
    const int na = h->na;
    dim3 grid((na + 127) / 128, na, G);
    RIP_LAUNCH(add_read_noise_kernel, grid, 128, 0, st, d_data, G, h->n, h->nb, (const float*)h->read.p,
               (const double*)(h->f_sums.p + 40), seed);
    RIP_API_END
}

extern "C" int rip_dark_as_l1_dev(rip_caldir* h, int G, uint16_t* d_data, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_data, "rip_dark_as_l1_dev: null argument");
    RIP_REQUIRE(h->d.n_dark - G == 0 || h->d.n_dark - G == 1, "Dark date cube has the wrong shape.");  // (the reference's message)
    use_device(h->device);
    const long npl = (long)h->n * h->n, count = (long)G * npl;
    RIP_LAUNCH(dark_to_u16_kernel, (unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream,
               (const float*)(h->dark_cube.p + (size_t)(h->d.n_dark - G) * npl), count, d_data);
    RIP_API_END
}

extern "C" int rip_active_diff_dev(int device, const float* d_a, const float* d_b, int n, int nb, float* d_out, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_a && d_b && d_out && n > 2 * nb, "rip_active_diff_dev: bad argument");
    use_device(device);
    const int na = n - 2 * nb;
    dim3 grid((na + 127) / 128, na);
    RIP_LAUNCH(active_diff_kernel, grid, 128, 0, (cudaStream_t)stream, d_a, d_b, n, nb, d_out);
    RIP_API_END
}

// =========================================================================================================
// Order statistics of a flat float32 array (np.percentile's inputs for the z clip of the noise layers,
// gen_noise_image.py:165-171) and the clip itself
// =========================================================================================================
namespace rip {

// K ranks at once (0-based, ascending, NaNs excluded): four 8-bit radix passes; each pass = one grid-wide histogram
// kernel (per-CTA shared-memory histograms of the elements matching each rank's prefix, flushed with atomics) and one
// single-CTA scan that narrows every rank's prefix.  state: [K] prefix, [K] remaining rank (u64 as 2 x u32), count.
constexpr int OS_KMAX = 16;
struct OsState {
    uint32_t prefix[OS_KMAX];
    unsigned long long rank[OS_KMAX];
    unsigned long long count;
    uint32_t mask;
};

__global__ void __launch_bounds__(256) order_hist_kernel(const float* __restrict__ arr, long count, int K, int pass,
                                                         const OsState* __restrict__ st, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[OS_KMAX][256];
    for (int i = threadIdx.x; i < K * 256; i += blockDim.x) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const uint32_t m = st->mask;
    // identical prefixes share a histogram row (the first of them): fewer atomics when ranks are neighbours
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
        const float v = arr[e];
        if (v != v) continue;
        const uint32_t k = f2key_sky(v), bin = (k >> shift) & 255u, kp = k & m;
        for (int r = 0; r < K; ++r) {
            const uint32_t pr = st->prefix[r];
            if (kp != pr) continue;
            bool dup = false;
            for (int q = 0; q < r; ++q) dup = dup || (st->prefix[q] == pr);
            if (!dup) atomicAdd(&sh[r][bin], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * 256; i += blockDim.x) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

__global__ void order_scan_kernel(int K, int pass, OsState* __restrict__ st, uint32_t* __restrict__ hist, float* __restrict__ out) {
    const int r = threadIdx.x;
    __shared__ unsigned long long cnt;
    const int shift = 24 - 8 * pass;
    if (pass == 0 && r == 0) {
        unsigned long long c = 0;
        for (int b = 0; b < 256; ++b) c += hist[b];
        cnt = c;
        st->count = c;
    }
    __syncthreads();
    unsigned long long new_rank = 0;
    uint32_t new_prefix = 0;
    if (r < K) {
        int row = r;  // the histogram row of this rank: the first rank with the same prefix
        for (int q = r - 1; q >= 0; --q)
            if (st->prefix[q] == st->prefix[r]) row = q;
        unsigned long long rk = st->rank[r];
        if (pass == 0 && rk >= cnt) rk = cnt ? cnt - 1 : 0;
        unsigned long long acc = 0;
        int b = 0;
        for (; b < 255; ++b) {
            const uint32_t h = hist[row * 256 + b];
            if (acc + h > rk) break;
            acc += h;
        }
        new_rank = rk - acc;
        new_prefix = st->prefix[r] | ((uint32_t)b << shift);
    }
    __syncthreads();  // every thread has read the prefixes / histograms of this pass
    if (r < K) {
        st->rank[r] = new_rank;
        st->prefix[r] = new_prefix;
        if (pass == 3) out[r] = st->count ? key2f_sky(new_prefix) : NAN;
    }
    for (int i = threadIdx.x; i < K * 256; i += blockDim.x) hist[i] = 0u;
    if (r == 0) st->mask |= 255u << shift;
}

__global__ void clip_kernel(float* __restrict__ a, long count, float lo, float hi) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    a[p] = np_clip<float>(a[p], lo, hi);
}

}  // namespace rip

extern "C" int rip_order_stats_dev(int device, const float* d_arr, long count, int K, const long* ranks, float* out,
                                   long* n_valid, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_arr && ranks && out && K >= 1 && K <= OS_KMAX && count >= 1, "rip_order_stats_dev: bad argument");
    use_device(device);
    cudaStream_t st = (cudaStream_t)stream;
    OsState hs;
    memset(&hs, 0, sizeof hs);
    for (int r = 0; r < K; ++r) hs.rank[r] = (unsigned long long)(ranks[r] < 0 ? 0 : ranks[r]);
    DevRaw dst;
    DevBuf<uint32_t> hist((size_t)OS_KMAX * 256);
    DevBuf<float> dout(K);
    dst.upload(&hs, sizeof hs, st);
    hist.zero(st);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const unsigned nblk = (unsigned)std::min<long>((count + 255) / 256, (long)sms * 8);
    for (int pass = 0; pass < 4; ++pass) {
        RIP_LAUNCH(order_hist_kernel, nblk, 256, 0, st, d_arr, count, K, pass, (const OsState*)dst.p, hist.p);
        RIP_LAUNCH(order_scan_kernel, 1, 32, 0, st, K, pass, (OsState*)dst.p, hist.p, dout.p);
    }
    dout.download(out, K, st);
    RIP_CUDA(cudaMemcpyAsync(&hs, dst.p, sizeof hs, cudaMemcpyDeviceToHost, st));
    RIP_CUDA(cudaStreamSynchronize(st));
    if (n_valid) *n_valid = (long)hs.count;
    RIP_API_END
}

extern "C" int rip_clip_dev(int device, float* d_arr, long count, float lo, float hi, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(d_arr && count >= 1, "rip_clip_dev: bad argument");
    use_device(device);
    RIP_LAUNCH(clip_kernel, (unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream, d_arr, count, lo, hi);
    RIP_API_END
}

// =========================================================================================================
// Noise directive "P" with flag r: re-sampled Poisson noise propagated through the ramp-fit weights
// (L1_to_L2/gen_noise_image.py:258-321)
// =========================================================================================================
namespace rip {

struct PoisArgs {
    int n, nb, G, n_samp;
    float frame_time;
    int group_of_read[64];            // group whose read list contains sample i, or -1
    float n_in_group[RIP_GMAX];       // len(read_pattern[j])
    float w[RIP_GMAX][RIP_GMAX];      // w[es][j]: weight of resultant j for pixels whose ramp ends at es (0 where undefined)
    uint8_t w_defined[RIP_GMAX];
    uint64_t seed;
};

template <typename TG>
__global__ void poisson_resample_kernel(const PoisArgs A, const float* __restrict__ skylevel, const TG* __restrict__ gain,
                                        const int8_t* __restrict__ endslice, float* __restrict__ diff) {
    const int na = A.n - 2 * A.nb;
    const int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= na) return;
    const long p = (long)ya * na + xa, q = (long)(ya + A.nb) * A.n + (xa + A.nb);
    typedef typename Promote<float, TG>::type TP;
    const TG g = np_clip<TG>(gain[q], (TG)1e-4, (TG)1e4);
    // e_per_slice = skylevel * gain * frame_time, clipped at 0 (:283, 289)
    TP e = (TP)skylevel[p] * (TP)g * (TP)A.frame_time;
    e = np_max<TP>(e, (TP)0);
    const int es_raw = endslice ? (int)endslice[p] : -1;
    const int es = es_raw > 0 ? es_raw : A.G - 1;
    Philox rng;
    rng.init(A.seed, (uint64_t)p, 128u);
    float cur = 0.0f;
    float delta[RIP_GMAX];
#pragma unroll
    for (int j = 0; j < RIP_GMAX; ++j) delta[j] = 0.0f;
    const double ed = (double)e;
    // sky-level expectations are a few electrons per sample: float32 inversion with exp(-e) hoisted out of
    // the sample loop (the float64 sampler called per draw took 25 ms per layer at 4096^2, the multiplication method 17 ms); PTRS above 10 electrons
    const bool small = ed < 10.0;
    const double inv_g = 1.0 / (double)g;
    const float lamf = (float)ed;
    const float enlam = small ? __expf(-lamf) : 0.0f;
    for (int i = 0; i < A.n_samp; ++i) {
        double s;
        if (small) {
            // inversion by sequential search with ONE uniform per draw (the multiplication method spent k+1 Philox
            // outputs per draw: 12 of 32 lanes active and the ALU pipe 51 % busy in profiles/r01/poisson_resample_*);
            // the search stops in the far tail once the float32 CDF no longer grows (term < 2^-25, beyond the mode)
            int k = 0;
            const float u = rng.uniform();
            float term = enlam, cdf = enlam;
            while (u > cdf) {
                ++k;
                term *= __fdividef(lamf, (float)k);
                const float nxt = cdf + term;
                if (nxt == cdf) break;
                cdf = nxt;
            }
            s = (double)k;
        } else {
            s = (double)poisson_draw(rng, ed);  // NaN expectation -> 0 draws; the difference below is NaN as in NumPy
        }
        s = s - ed;
        s = s * inv_g;  // (the reference divides; one reciprocal per pixel instead of 35 float64 divisions: the quantity is a
                        //  random draw, validated statistically)
        cur = (float)((double)cur + s);
        const int j = A.group_of_read[i];  // warp-uniform
        if (j >= 0) {
            const float qv = cur / A.n_in_group[j];
#pragma unroll
            for (int t = 0; t < RIP_GMAX; ++t)
                if (t == j) delta[t] = delta[t] + qv;
        }
    }
    float d = diff[p];
    if (es < RIP_GMAX && A.w_defined[es]) {
#pragma unroll
        for (int j = 0; j < RIP_GMAX; ++j)
            if (j < A.G) d = d + A.w[es][j] * delta[j];
    }
    diff[p] = d;
}

}  // namespace rip

namespace rip {

// ---------------------------------------------------------------------------------------------------------
// Noise directive "O" (reference L1_to_L2/gen_noise_image.py:173-227): pseudo-Poisson draws from the Pearson family with
// the 2nd-4th moments of the ramp-fitted Poisson noise, diff += draw / gain.  Per active pixel: I = max(gain *
// data_withsky, 0.01), the (tilde nu_21, nu_31, nu_41) of the pixel's ramp end, beta_1 = nu31^2 / (nu21^3 I),
// beta_2 = (3 nu21^2 I + nu41) / (nu21^2 I); outside the admissible region the draw is 0
// (GalPoisson/draw_with_tilnus.py:42-60); below the Type III line (beta_2 < 1.5 beta_1 + 3) the Pearson Type I = a
// shifted and scaled Beta(a, b) with (a, b) from the closed form of :160-198 -- the only type the production read
// patterns reach (their nu_41 is negative); between the Type III and Type V lines the Type VI = beta prime of :670-721.
// Types III and V (exact equalities), IV (Devroye's sampler) are counted in *unsupported and draw 0: the host raises.
// Beta(a, b) = X / (X + Y) with Gamma variates by Marsaglia & Tsang (2000).
// ---------------------------------------------------------------------------------------------------------
__device__ inline double gamma_draw(Philox& rng, double a) {
    double boost = 1.0;
    if (a < 1.0) {  // Gamma(a) = Gamma(a + 1) U^(1/a)
        boost = pow(rng.uniform53(), 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 64; ++it) {
        float xa, xb;
        rng.normal2(xa, xb);
        const double x = (double)xa;
        const double t = 1.0 + c * x;
        if (t <= 0.0) continue;
        const double v = t * t * t;
        const double u = rng.uniform53();
        if (u < 1.0 - 0.0331 * (x * x) * (x * x) || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;  // (never reached in practice: acceptance > 95 % per trial)
}

// Beta(a, b): Joehnk's algorithm when both shapes are <= 1 (with its log-space branch for draws that underflow -- near
// the admissibility boundary beta_2 -> beta_1 + 1 the shapes tend to 0 and the distribution to two points), two Gamma
// variates otherwise: the selection NumPy's generator makes.
__device__ inline double beta_draw(Philox& rng, double a, double b) {
    if (a <= 1.0 && b <= 1.0) {
        for (int it = 0; it < 4096; ++it) {
            const double U = rng.uniform53(), V = rng.uniform53();
            const double X = pow(U, 1.0 / a), Y = pow(V, 1.0 / b);
            const double XpY = X + Y;
            if (XpY <= 1.0) {
                if (XpY > 0.0) return X / XpY;
                double lx = log(U) / a, ly = log(V) / b;
                const double lm = lx > ly ? lx : ly;
                lx -= lm;
                ly -= lm;
                return exp(lx - log(exp(lx) + exp(ly)));
            }
        }
        return a / (a + b);
    }
    const double X = gamma_draw(rng, a), Y = gamma_draw(rng, b);
    return X / (X + Y);
}

// Re ln Gamma(x + i y), x > 0: recurrence up to Re z >= 16, then Stirling's series (truncation < 2e-12 there).
// (GalPoisson/draw_with_tilnus.py:296-306 takes it from scipy.special.loggamma.)
__host__ __device__ inline double re_lgamma_cplx(double x, double y) {
    double acc = 0.0;
    while (x < 16.0) {
        acc += log(x * x + y * y);  // 2 Re ln z
        x += 1.0;
    }
    const double r2 = x * x + y * y, y2 = y * y, x2 = x * x;
    const double re1 = x / r2;                                                        // Re z^-1
    const double re3 = x * (x2 - 3.0 * y2) / (r2 * r2 * r2);                          // Re z^-3
    const double re5 = x * (x2 * x2 - 10.0 * x2 * y2 + 5.0 * y2 * y2) / (r2 * r2 * r2 * r2 * r2);  // Re z^-5
    return (x - 0.5) * (0.5 * log(r2)) - y * atan2(y, x) - x + 0.91893853320467274178 + re1 / 12.0 - re3 / 360.0 +
           re5 / 1260.0 - 0.5 * acc;
}

// log of the normalisation of the Pearson IV density in the angle t = atan((x - lam) / a):
// g(t) = k cos(t)^(2m-2) exp(-nu t), k = 2^(2m-2) |Gamma(m + i nu/2)|^2 / (pi Gamma(2m-1))   (Heinrich 2004, eq. 6-8)
__host__ __device__ inline double pearson4_logk(double m, double nu) {
    return (2.0 * m - 2.0) * 0.69314718055994530942 + 2.0 * re_lgamma_cplx(m, 0.5 * nu) - 1.14472988584940017414 - lgamma(2.0 * m - 1.0);
}

// One Pearson IV deviate, f(x) ~ (1 + xi^2)^-m exp(-nu atan xi), xi = (x - lam) / a, m > 1: Devroye's rejection method
// for log-concave densities applied to the angle (Heinrich 2004, section 7; the sampler of the reference's
// pt4_rvs_devroye, GalPoisson/draw_with_tilnus.py:444-483).  The hat needs rc = 1 / g(mode); the reference's rc carries a
// factor `a` on top (its log k is the x-space constant): its hat still dominates for a >= 1 (same distribution, low
// acceptance -- the reason it switches to a second sampler) and does not for a < 1.  Here rc is the exact one:
// acceptance >= 1/4 for every (m, nu), one sampler.
__device__ inline double pearson4_draw(Philox& rng, double m, double nu, double a, double lam) {
    const double b = 2.0 * m - 2.0;
    const double M = atan2(-nu, b);
    const double r_const = b * log(b / hypot(b, nu)) - nu * M;
    const double rc = exp(-r_const - pearson4_logk(m, nu));
    for (int it = 0; it < 512; ++it) {
        double x = 4.0 * rng.uniform53(), z = 0.0;
        bool right = false;
        if (x > 2.0) { x -= 2.0; right = true; }
        if (x > 1.0) { z = log(x - 1.0); x = 1.0 - z; }
        x = right ? M + rc * x : M - rc * x;
        if (!(fabs(x) < 1.57079632679489661923)) continue;
        if (z + log(rng.uniform53()) > b * log(cos(x)) - nu * x - r_const) continue;
        return a * tan(x) + lam;
    }
    return a * tan(M) + lam;  // (not reached: 512 rejections in a row have probability < 1e-60)
}

struct PearsonArgs {
    int n, nb, G, start;
    double nu21[RIP_GMAX], nu31[RIP_GMAX], nu41[RIP_GMAX];
    unsigned char defined[RIP_GMAX];
    uint64_t seed;
};

template <typename TG>
__global__ void pearson_noise_kernel(const PearsonArgs A, const float* __restrict__ withsky, const TG* __restrict__ gain,
                                     const int8_t* __restrict__ endslice, float* __restrict__ diff, int* __restrict__ unsupported) {
    const int na = A.n - 2 * A.nb;
    const int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= na) return;
    const long p = (long)ya * na + xa, q = (long)(ya + A.nb) * A.n + (xa + A.nb);
    const int es0 = endslice ? (int)endslice[p] : 0;
    const int es = es0 > 0 ? es0 : A.G - 1;  // np.where(endslice > 0, endslice, ngrp - 1)
    if (es < A.start + 1 || es >= A.G || !A.defined[es]) return;
    typedef typename Promote<float, TG>::type TP;
    const TP g = np_clip<TP>((TP)gain[q], (TP)1e-4, (TP)1e4);
    const double gI = (double)(g * (TP)withsky[p]);
    const double I = gI < 0.01 ? 0.01 : gI;  // np.clip(., 0.01, None): NaN stays NaN -> every test below is false
    const double n21 = A.nu21[es], n31 = A.nu31[es], n41 = A.nu41[es];
    const double b1 = n31 * n31 / (n21 * n21 * n21 * I);
    const double b2 = (3.0 * n21 * n21 * I + n41) / (n21 * n21 * I);
    const bool base = (b2 > 0.0) && (b1 >= 0.0) && (b2 > b1 + 1.0) && (b2 > 0.75 * b1);
    if (!base) return;
    const double rhs1 = 1.5 * b1 + 3.0;
    if (!(b2 < rhs1)) {  // Types III (==), VI, V (==), IV
        const double rhs2 = (48.0 + 39.0 * b1 + 6.0 * pow(4.0 + b1, 1.5)) / (32.0 - b1);
        if (b2 > rhs1 && b2 < rhs2) {
            // Type VI = shifted / scaled beta prime (GalPoisson/draw_with_tilnus.py:670-721): y = X / Y with
            // X ~ Gamma(alpha), Y ~ Gamma(beta); draw = sign (scale y - shift)
            const double r = 6.0 * (b2 - b1 - 1.0) / (3.0 * b1 - 2.0 * b2 + 6.0);
            const double eps = r * r / (4.0 + (b1 / 4.0) * (r + 2.0) * (r + 2.0) / (r + 1.0));
            const double dd = sqrt(r * r - 4.0 * eps);
            const double q1 = (2.0 - r + dd) / 2.0, q2 = (r - 2.0 + dd) / 2.0;
            const double al = q2 + 1.0, be = q1 - q2 - 1.0;
            const double var1 = al * (al + be - 1.0) / ((be - 2.0) * (be - 1.0) * (be - 1.0));
            const double sc = sqrt(n21 * I / var1), sh = sc * (al / (be - 1.0));
            if (!(al > 0.0) || !(be > 2.0) || !(sc == sc)) { atomicAdd(unsupported, 1); return; }
            Philox rng;
            rng.init(A.seed, (uint64_t)p, 160u);
            const double X = gamma_draw(rng, al), Y = gamma_draw(rng, be);
            const double draw = (n31 >= 0.0 ? 1.0 : -1.0) * (sc * (X / Y) - sh);
            diff[p] = (float)((TP)diff[p] + (TP)(float)draw / g);
            return;
        }
        const double sgn = (n31 >= 0.0) ? 1.0 : -1.0;
        double draw;
        if (b2 == rhs1) {
            // Type III = shifted / scaled Gamma (:256-281; its sign is that of nu31 with 0 counted as negative)
            const double sc = fabs(n31) / (2.0 * n21), shape = 4.0 * n21 * n21 * n21 * I / (n31 * n31);
            if (!(shape > 0.0) || !(sc == sc)) { atomicAdd(unsupported, 1); return; }
            Philox rng;
            rng.init(A.seed, (uint64_t)p, 160u);
            draw = ((n31 > 0.0) ? 1.0 : -1.0) * (sc * gamma_draw(rng, shape) - shape * sc);
        } else if (b2 == rhs2) {
            // Type V = shifted inverse Gamma (:601-668): y = beta / Gamma(alpha)
            const double st = sqrt(4.0 + b1);
            const double pp = 4.0 * (1.0 + 2.0 / b1 + st / b1), pm = 4.0 * (1.0 + 2.0 / b1 - st / b1);
            const double pq = pp > 4.0 ? pp : pm;
            const double be = sqrt(n21 * I) * (pq - 2.0) * sqrt(pq - 3.0), al = pq - 1.0;
            if (!(al > 1.0) || !(be > 0.0)) { atomicAdd(unsupported, 1); return; }
            Philox rng;
            rng.init(A.seed, (uint64_t)p, 160u);
            draw = sgn * (be / gamma_draw(rng, al) - be / (al - 1.0));
        } else if (b2 > rhs2 && b1 < 32.0) {
            // Type IV (:535-598)
            const double r = 6.0 * (b2 - b1 - 1.0) / (2.0 * b2 - 3.0 * b1 - 6.0);
            const double inner = 16.0 * (r - 1.0) - b1 * (r - 2.0) * (r - 2.0);
            if (!(r > 1.0) || !(inner > 0.0)) { atomicAdd(unsupported, 1); return; }  // the reference raises ValueError here
            const double nu = -sgn * (r * (r - 2.0) * sqrt(b1) / sqrt(inner));  // sign(mu_3) = -sign(nu)
            const double a4 = sqrt(n21 * I * inner) / 4.0, m4 = r / 2.0 + 1.0;
            Philox rng;
            rng.init(A.seed, (uint64_t)p, 160u);
            draw = pearson4_draw(rng, m4, nu, a4, a4 * nu / (2.0 * (m4 - 1.0)));
        } else {
            return;  // (beta_1 >= 32 above the Type V line: no type in the reference either, the pixel stays 0)
        }
        diff[p] = (float)((TP)diff[p] + (TP)(float)draw / g);
        return;
    }
    // Type I: u = a + b, v = (a - b)^2 / (a b)
    const double u = 3.0 * (b1 - b2 + 1.0) / ((b2 - 3.0) - 1.5 * b1);
    const double v = b1 * (u + 2.0) * (u + 2.0) / (4.0 * (u + 1.0));
    const double sq = sqrt(v / (v + 4.0));
    const double ap = 0.5 * u * (1.0 + sq), bp = 0.5 * u * (1.0 - sq);
    const bool cond = (n31 < 0.0) ? (ap > bp) : (ap < bp);  // sign of the skew picks the branch
    const double a = cond ? ap : bp, b = cond ? bp : ap;
    const double mean = a / (a + b), var = a * b / ((a + b) * (a + b) * (a + b + 1.0));
    const double scale = sqrt(n21 * I / var);
    if (!(a > 0.0) || !(b > 0.0) || !(scale == scale)) return;
    Philox rng;
    rng.init(A.seed, (uint64_t)p, 160u);
    const double y = beta_draw(rng, a, b);
    diff[p] = (float)((TP)diff[p] + (TP)(float)(scale * (y - mean)) / g);
}

}  // namespace rip

// tilnu: [G][3] = (nu21, nu31, nu41) in e/s units of the ramp ending at group i (rows with defined[i] == 0 are skipped);
// d_unsupported: device int counter of pixels whose Pearson parameters are invalid (Type IV with r <= 1 or a non-positive
// discriminant, where the reference raises ValueError; Type VI / III / V with non-positive shapes); the caller zeroes it
extern "C" int rip_pearson_noise_dev(rip_caldir* h, const float* d_withsky, const int8_t* d_endslice, int G, int start,
                                     const double* tilnu, const uint8_t* defined, uint64_t seed, float* d_diff,
                                     int32_t* d_unsupported, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_withsky && tilnu && defined && d_diff && d_unsupported, "rip_pearson_noise_dev: null argument");
    RIP_REQUIRE(G >= 1 && G <= RIP_GMAX && start >= 0, "rip_pearson_noise_dev: G=%d / start=%d out of range", G, start);
    use_device(h->device);
    PearsonArgs A;
    memset(&A, 0, sizeof A);
    A.n = h->n; A.nb = h->nb; A.G = G; A.start = start; A.seed = seed;
    for (int i = 0; i < G; ++i) {
        A.nu21[i] = tilnu[3 * i]; A.nu31[i] = tilnu[3 * i + 1]; A.nu41[i] = tilnu[3 * i + 2];
        A.defined[i] = defined[i];
    }
    dim3 grid((h->na + 127) / 128, h->na);
    if (h->d.gain_dtype == RIP_F64)
        RIP_LAUNCH(pearson_noise_kernel<double>, grid, 128, 0, (cudaStream_t)stream, A, d_withsky, (const double*)h->gain.p, d_endslice, d_diff, (int*)d_unsupported);
    else
        RIP_LAUNCH(pearson_noise_kernel<float>, grid, 128, 0, (cudaStream_t)stream, A, d_withsky, (const float*)h->gain.p, d_endslice, d_diff, (int*)d_unsupported);
    RIP_API_END
}

extern "C" double rip_pearson4_logk_host(double m, double nu) { return rip::pearson4_logk(m, nu); }

extern "C" int rip_poisson_resample_dev(rip_caldir* h, const float* d_skylevel, const int8_t* d_endslice, int G, int n_samp,
                                        const int32_t* group_of_read, const float* weights /*[G][G] row es*/,
                                        const uint8_t* w_defined, double frame_time, uint64_t seed, float* d_diff, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_skylevel && group_of_read && weights && w_defined && d_diff, "rip_poisson_resample_dev: null argument");
    RIP_REQUIRE(G >= 1 && G <= RIP_GMAX && n_samp >= 1 && n_samp <= 64, "rip_poisson_resample_dev: G=%d / n_samp=%d out of range", G, n_samp);
    use_device(h->device);
    PoisArgs A;
    memset(&A, 0, sizeof A);
    A.n = h->n; A.nb = h->nb; A.G = G; A.n_samp = n_samp; A.frame_time = (float)frame_time; A.seed = seed;
    for (int i = 0; i < 64; ++i) A.group_of_read[i] = i < n_samp ? group_of_read[i] : -1;
    for (int j = 0; j < G; ++j) {
        int c = 0;
        for (int i = 0; i < n_samp; ++i) c += group_of_read[i] == j;
        A.n_in_group[j] = (float)(c > 0 ? c : 1);
        A.w_defined[j] = w_defined[j];
        for (int t = 0; t < G; ++t) A.w[j][t] = weights[j * G + t];
    }
    dim3 grid((h->na + 127) / 128, h->na);
    if (h->d.gain_dtype == RIP_F64)
        RIP_LAUNCH(poisson_resample_kernel<double>, grid, 128, 0, (cudaStream_t)stream, A, d_skylevel, (const double*)h->gain.p, d_endslice, d_diff);
    else
        RIP_LAUNCH(poisson_resample_kernel<float>, grid, 128, 0, (cudaStream_t)stream, A, d_skylevel, (const float*)h->gain.p, d_endslice, d_diff);
    RIP_API_END
}
