// Stage kernels: one per reference function of the hot path (host-pointer C ABI, include/rip_b200.h).
// These are the parity surface for the reference's public functions; the throughput path is the fused kernel in
// rip_fit.cu.  Compiled with -fmad=false: every multiply and add is a separate IEEE operation.
#include <cfloat>

#include "rip_launch.h"

namespace rip {

// =========================================================================================================
// Legendre linearity
// =========================================================================================================
template <typename TZ, int PMAX>
__global__ void lin_eval_kernel(const TZ* __restrict__ z, const float* __restrict__ coefs, int P, long npix,
                                int linextrap, float* __restrict__ phi, uint8_t* __restrict__ exflag) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float c[PMAX];
#pragma unroll
    for (int L = 0; L < PMAX; ++L) c[L] = (L < P) ? coefs[(long)L * npix + p] : 0.0f;
    bool ex;
    float r = linextrap ? legendre_eval<TZ, PMAX, true>(z[p], c, P, ex) : legendre_eval<TZ, PMAX, false>(z[p], c, P, ex);
    phi[p] = r;
    if (exflag) exflag[p] = ex ? 1 : 0;
}

template <int GMAX, int PMAX>
__global__ void multilin_kernel(const float* __restrict__ S, int G, long npix, const float* __restrict__ coefs, int P,
                                const float* __restrict__ Smin, const float* __restrict__ Smax,
                                const float* __restrict__ Sref, const uint32_t* __restrict__ lin_dq,
                                const uint8_t* __restrict__ attempt, int dnff, int single_frame,
                                float* __restrict__ phi, uint32_t* __restrict__ dq_out) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float c[PMAX];
#pragma unroll
    for (int L = 0; L < PMAX; ++L) c[L] = (L < P) ? coefs[(long)L * npix + p] : 0.0f;
    uint32_t dq = lin_dq[p];
    if (single_frame) {  // linearity(): no override, flag every extrapolated pixel (ipc_linearity.py:270-272)
        bool ex;
        float z = lin_z<float>(S[p], Smin[p], Smax[p]);
        phi[p] = legendre_eval<float, PMAX, true>(z, c, P, ex);
        if (ex) dq |= DQ_NO_LIN_CORR;
        dq_out[p] = dq;
        return;
    }
    float s[GMAX], out[GMAX];
    uint32_t am = 0u;
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
        s[g] = (g < G) ? S[(long)g * npix + p] : 0.0f;
        if (g < G && (!attempt || attempt[(long)g * npix + p])) am |= 1u << g;
    }
    multilin_pixel<GMAX, PMAX>(s, G, c, P, Smin[p], Smax[p], Sref[p], dq, am, dnff != 0, out);
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
        if (g < G) phi[(long)g * npix + p] = out[g];
    dq_out[p] = dq;
}

template <typename TZ, int PMAX>
__global__ void invlin_kernel(const TZ* __restrict__ Slin, long npix, const float* __restrict__ coefs, int P,
                              const float* __restrict__ Smin, const float* __restrict__ Smax, TZ* __restrict__ Sout,
                              uint8_t* __restrict__ exflag) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float c[PMAX];
#pragma unroll
    for (int L = 0; L < PMAX; ++L) c[L] = (L < P) ? coefs[(long)L * npix + p] : 0.0f;
    bool ex;
    Sout[p] = invlin_pixel<TZ, PMAX>(Slin[p], c, P, Smin[p], Smax[p], ex);
    if (exflag) exflag[p] = ex ? 1 : 0;
}

// =========================================================================================================
// IPC
// =========================================================================================================
// fwd at (y,x): sum over taps of img(y-dy,x-dx) * K[1+dy][1+dx][y-dy][x-dx], reference accumulation order.
template <typename TI, typename TK, typename F>
__device__ __forceinline__ TI ipc_fwd_at(F img, const TK* __restrict__ K, int ny, int nx, int y, int x) {
    const long pl = (long)ny * nx;
    const int DY[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1};
    const int DX[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
    TI acc = (TI)img(y, x) * (TI)K[4 * pl + (long)y * nx + x];
#pragma unroll
    for (int q = 1; q < 9; ++q) {
        const int ys = y - DY[q], xs = x - DX[q];
        if (ys >= 0 && ys < ny && xs >= 0 && xs < nx)
            acc = acc + (TI)img(ys, xs) * (TI)K[(long)((1 + DY[q]) * 3 + (1 + DX[q])) * pl + (long)ys * nx + xs];
    }
    return acc;
}

// im2 = gain * image (type TIM) or a converting copy
template <typename TIMG, typename TG, typename TIM>
__global__ void mulgain_kernel(const TIMG* __restrict__ img, const TG* __restrict__ gain, long npix, TIM* __restrict__ out) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    out[p] = gain ? (TIM)gain[p] * (TIM)img[p] : (TIM)img[p];
}

// out = ipc_fwd(im) [/ gain]
template <typename TIM, typename TK, typename TG>
__global__ void ipc_fwd_kernel(const TIM* __restrict__ im, const TK* __restrict__ K, const TG* __restrict__ gain, int ny,
                               int nx, typename Promote<TIM, TK>::type* __restrict__ out) {
    typedef typename Promote<TIM, TK>::type TI;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    TI v = ipc_fwd_at<TI, TK>([&](int yy, int xx) { return im[(long)yy * nx + xx]; }, K, ny, nx, y, x);
    if (gain) v = v / (TI)gain[(long)y * nx + x];
    out[(long)y * nx + x] = v;
}

// one ipc_rev iteration: out = (prev + im2) - ipc_fwd(prev)        (ipc_linearity.py:139)
template <typename TP, typename TIM, typename TK>
__global__ void ipc_rev_iter_kernel(const TP* __restrict__ prev, const TIM* __restrict__ im2, const TK* __restrict__ K,
                                    int ny, int nx, typename Promote<TP, TK>::type* __restrict__ out) {
    typedef typename Promote<TP, TK>::type TI;
    typedef typename Promote<TP, TIM>::type TS;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx) return;
    const long p = (long)y * nx + x;
    TI f = ipc_fwd_at<TI, TK>([&](int yy, int xx) { return prev[(long)yy * nx + xx]; }, K, ny, nx, y, x);
    TS sum = (TS)prev[p] + (TS)im2[p];
    out[p] = (TI)sum - f;
}

template <typename T, typename TG>
__global__ void divgain_kernel(T* __restrict__ a, const TG* __restrict__ gain, long npix) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    a[p] = a[p] / (T)gain[p];
}

// ---- DN-space deconvolution of the active region of a full-frame f32 plane (correct_cube / get_flat) ----
template <typename TG, typename TK>
__global__ void ipc_dn_pass1_kernel(const float* __restrict__ plane, int ny, int nx, int nb, const TK* __restrict__ K,
                                    const TG* __restrict__ gain, int clip_gain, float g_lo,
                                    typename Promote<typename Promote<float, TG>::type, TK>::type* __restrict__ o1) {
    typedef typename Promote<float, TG>::type TIM;
    typedef typename Promote<TIM, TK>::type TI;
    const int nya = ny - 2 * nb, nxa = nx - 2 * nb;
    int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= nxa) return;
    auto D = [&](int yy, int xx) -> TIM {
        const long q = (long)(yy + nb) * nx + (xx + nb);
        if (!gain) return (TIM)plane[q];
        TG g = gain[q];
        if (clip_gain) g = np_max<TG>(g, (TG)g_lo);
        // correct_cube: data*g (utils/ipc_linearity.py:186); get_flat: gain*image (:136) -- commutative
        return (TIM)plane[q] * (TIM)g;
    };
    TI f = ipc_fwd_at<TI, TK>(D, K, nya, nxa, ya, xa);
    TIM d0 = D(ya, xa);
    o1[(long)ya * nxa + xa] = (TI)(TIM)(d0 + d0) - f;
}

template <typename TG, typename TK>
__global__ void ipc_dn_pass2_kernel(float* __restrict__ plane, int ny, int nx, int nb, const TK* __restrict__ K,
                                    const TG* __restrict__ gain, int clip_gain, float g_lo,
                                    const typename Promote<typename Promote<float, TG>::type, TK>::type* __restrict__ o1) {
    typedef typename Promote<float, TG>::type TIM;
    typedef typename Promote<TIM, TK>::type TI;
    const int nya = ny - 2 * nb, nxa = nx - 2 * nb;
    int xa = blockIdx.x * blockDim.x + threadIdx.x, ya = blockIdx.y;
    if (xa >= nxa) return;
    const long q = (long)(ya + nb) * nx + (xa + nb);
    TI f = ipc_fwd_at<TI, TK>([&](int yy, int xx) { return o1[(long)yy * nxa + xx]; }, K, nya, nxa, ya, xa);
    TIM d0;
    TI res;
    if (gain) {
        TG g = gain[q];
        if (clip_gain) g = np_max<TG>(g, (TG)g_lo);
        d0 = (TIM)plane[q] * (TIM)g;
        res = ((o1[(long)ya * nxa + xa] + (TI)d0) - f) / (TI)g;
    } else {
        d0 = (TIM)plane[q];
        res = (o1[(long)ya * nxa + xa] + (TI)d0) - f;
    }
    plane[q] = (float)res;
}

template <typename TG, typename TK>
static void ipc_rev_dn_t(float* plane, int ny, int nx, int nb, const void* K, const void* gain, bool clip_gain,
                         float g_lo, void* tmp, cudaStream_t st) {
    typedef typename Promote<typename Promote<float, TG>::type, TK>::type TI;
    const int nya = ny - 2 * nb, nxa = nx - 2 * nb;
    dim3 grid((nxa + 127) / 128, nya), block(128);
    RIP_LAUNCH((ipc_dn_pass1_kernel<TG, TK>), grid, block, 0, st, plane, ny, nx, nb, (const TK*)K, (const TG*)gain,
               clip_gain ? 1 : 0, g_lo, (TI*)tmp);
    RIP_LAUNCH((ipc_dn_pass2_kernel<TG, TK>), grid, block, 0, st, plane, ny, nx, nb, (const TK*)K, (const TG*)gain,
               clip_gain ? 1 : 0, g_lo, (const TI*)tmp);
}

void launch_ipc_rev_dn(float* plane, int ny, int nx, int nb, const void* K, int k_dtype, const void* gain_full,
                       int g_dtype, bool clip_gain, float g_lo, void* tmp, cudaStream_t st) {
    const bool gd = gain_full && g_dtype == RIP_F64, kd = (k_dtype == RIP_F64);
    if (!gd && !kd) ipc_rev_dn_t<float, float>(plane, ny, nx, nb, K, gain_full, clip_gain, g_lo, tmp, st);
    else if (gd && !kd) ipc_rev_dn_t<double, float>(plane, ny, nx, nb, K, gain_full, clip_gain, g_lo, tmp, st);
    else if (!gd && kd) ipc_rev_dn_t<float, double>(plane, ny, nx, nb, K, gain_full, clip_gain, g_lo, tmp, st);
    else ipc_rev_dn_t<double, double>(plane, ny, nx, nb, K, gain_full, clip_gain, g_lo, tmp, st);
}

// =========================================================================================================
// Medians (reference-pixel statistics) -- block-level bitonic sort in shared memory
// =========================================================================================================
__device__ __forceinline__ void block_bitonic_sort(float* v, int npow2) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    float a = v[i], b = v[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { v[i] = b; v[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// np.median of `count` values already sorted ascending in v (f32 mean of the two middle values when even)
__device__ __forceinline__ float median_sorted(const float* v, int count) {
    if (count & 1) return v[count / 2];
    return (v[count / 2 - 1] + v[count / 2]) / 2.0f;
}

// per-row median of up to two column segments of a [n, ncols] f32 image
__global__ void row_median_kernel(const float* __restrict__ image, int ncols, int s0, int l0, int s1, int l1,
                                  int npow2, float* __restrict__ out) {
    extern __shared__ float sv[];
    const int row = blockIdx.x;
    const float* r = image + (long)row * ncols;
    const int count = l0 + l1;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        float v = FLT_MAX;
        if (i < l0) v = r[s0 + i];
        else if (i < count) v = r[s1 + (i - l0)];
        else v = INFINITY;
        sv[i] = v;
    }
    __syncthreads();
    block_bitonic_sort(sv, npow2);
    if (threadIdx.x == 0) out[row] = median_sorted(sv, count);
}

// ctr = median(ref_med[0..n)) -> ctr_out[0]
__global__ void vec_median_kernel(const float* __restrict__ v, int count, int npow2, float* __restrict__ ctr_out) {
    extern __shared__ float sv[];
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) sv[i] = (i < count) ? v[i] : INFINITY;
    __syncthreads();
    block_bitonic_sort(sv, npow2);
    if (threadIdx.x == 0) ctr_out[0] = median_sorted(sv, count);
}

// image[i,:] = f32(f64(image[i,:]) - m*(ref_med[i]-ctr))     (reference_subtraction.py:122-123; App. A2)
__global__ void refsub_row_apply_kernel(float* __restrict__ image, int n, int ncols, double m_med,
                                        const float* __restrict__ ref_med, const float* __restrict__ ctr) {
    const int row = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= ncols) return;
    const float dm = ref_med[row] - ctr[0];
    const double corr = m_med * (double)dm;
    const long p = (long)row * ncols + x;
    image[p] = (float)((double)image[p] - corr);
}

// medians of the bottom (rows 0:4) and top (rows n-4:n) reference pixels of each 128-column channel
__global__ void chan_median_kernel(const float* __restrict__ image, int n, int ncols, float* __restrict__ med /*[nchan][2]*/) {
    __shared__ float sv[512];
    const int ch = blockIdx.x, side = blockIdx.y;
    const int rbase = side ? (n - 4) : 0;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sv[i] = image[(long)(rbase + (i >> 7)) * ncols + ch * 128 + (i & 127)];
    __syncthreads();
    block_bitonic_sort(sv, 512);
    if (threadIdx.x == 0) med[ch * 2 + side] = median_sorted(sv, 512);
}

// line through (1.5, bottom), (n-2.5, top); image[j, ch] -= m*j + c in f64   (reference_subtraction.py:57-68)
// NB the reference solves the 2x2 system with np.linalg.lstsq (LAPACK gelsd); the closed form used here agrees to
// ~1e-16 relative, far below the f32 rounding of the stored pixel (DESIGN.md "known deviations").
__global__ void refsub_chan_apply_kernel(float* __restrict__ image, int n, int ncols, int nchan, const float* __restrict__ med) {
    const int row = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nchan * 128) return;
    const int ch = x >> 7;
    const double b = (double)med[ch * 2], t = (double)med[ch * 2 + 1];
    const double x0 = 1.5, x1 = (double)n - 2.5;
    const double m = (t - b) / (x1 - x0);
    const double c = b - m * x0;
    const double line = m * (double)row + c;
    const long p = (long)row * ncols + x;
    image[p] = (float)((double)image[p] - line);
}

// =========================================================================================================
// Flat (utils/flatutils.py:44-69): pad, flag, clip.  The IPC part is launch_ipc_rev_dn with the clipped gain.
// =========================================================================================================
template <typename TG>
__global__ void flat_prepare_kernel(const float* __restrict__ flat, int n, int nb, const TG* __restrict__ gain,
                                    uint32_t* __restrict__ pdq, int ipc_deconvolve, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n) return;
    const long p = (long)y * n + x;
    const bool active = (y >= nb && y < n - nb && x >= nb && x < n - nb);
    float f = active ? flat[p] : 1.0f;
    uint32_t fl = 0u;
    if (f < 0.1f || f > 10.0f) fl |= DQ_NO_FLAT_FIELD;
    f = np_clip<float>(f, 0.1f, 10.0f);
    if (ipc_deconvolve && active && gain && gain[p] <= (TG)0.1) fl |= DQ_NO_GAIN_VALUE;
    if (pdq && fl) pdq[p] |= fl;
    out[p] = f;
}

void launch_flat_prepare(const float* flat, int n, int nb, const void* gain, int g_dtype, uint32_t* pdq,
                         int ipc_deconvolve, float* out, cudaStream_t st) {
    dim3 grid((n + 255) / 256, n);
    if (gain && g_dtype == RIP_F64)
        RIP_LAUNCH((flat_prepare_kernel<double>), grid, 256, 0, st, flat, n, nb, (const double*)gain, pdq, ipc_deconvolve, out);
    else
        RIP_LAUNCH((flat_prepare_kernel<float>), grid, 256, 0, st, flat, n, nb, (const float*)gain, pdq, ipc_deconvolve, out);
}

// =========================================================================================================
// Saturation flagging (restatement)
// =========================================================================================================
__global__ void sat_bits_kernel(const uint16_t* __restrict__ raw, int G, int n, const float* __restrict__ thr,
                                const uint32_t* __restrict__ sat_dq, int skip, uint32_t* __restrict__ bits,
                                uint32_t* __restrict__ pdq) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long npl = (long)n * n;
    if (p >= npl) return;
    float t = thr[p];
    const bool nocheck_flag = (sat_dq[p] & DQ_NO_SAT_CHECK) != 0u;
    if (nocheck_flag || t != t) t = INFINITY;
    if (nocheck_flag && pdq) pdq[p] |= DQ_NO_SAT_CHECK;
    uint32_t b = 0u;
    bool cum = false;
    for (int g = skip; g < G; ++g) {
        const float v = (float)raw[(long)g * npl + p];
        cum = cum || (v >= t);
        if (cum) b |= 1u << g;
        if (v <= 0.0f) b |= 1u << (16 + g);
    }
    bits[p] = b;
}

__global__ void sat_grow_kernel(const uint32_t* __restrict__ bits, int G, int n, int backup, int skip,
                                uint8_t* __restrict__ rdq) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n) return;
    const long npl = (long)n * n;
    uint32_t grown = 0u;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < n && xx >= 0 && xx < n) grown |= bits[(long)yy * n + xx];
        }
    grown &= 0xffffu;
    uint32_t satm = grown;
    for (int b = 1; b <= backup; ++b) satm |= grown >> b;
    satm &= ((1u << G) - 1u) & ~((1u << skip) - 1u);
    const uint32_t adf = bits[(long)y * n + x] >> 16;
    for (int g = 0; g < G; ++g) {
        uint8_t v = 0;
        if ((satm >> g) & 1u) v |= (uint8_t)DQ_SATURATED;
        if ((adf >> g) & 1u) v |= (uint8_t)(DQ_AD_FLOOR | DQ_DO_NOT_USE);
        if (v) rdq[(long)g * npl + (long)y * n + x] |= v;
    }
}

}  // namespace rip

// =========================================================================================================
// C ABI
// =========================================================================================================
using namespace rip;

static inline unsigned nblk(long n, int b) { return (unsigned)((n + b - 1) / b); }

template <int PMAX>
static void lin_eval_dispatch(const void* dz, int z_dtype, const float* dc, int P, long npix, int linextrap, float* dphi,
                              uint8_t* dex) {
    if (z_dtype == RIP_F64)
        RIP_LAUNCH((lin_eval_kernel<double, PMAX>), nblk(npix, 256), 256, 0, 0, (const double*)dz, dc, P, npix, linextrap, dphi, dex);
    else
        RIP_LAUNCH((lin_eval_kernel<float, PMAX>), nblk(npix, 256), 256, 0, 0, (const float*)dz, dc, P, npix, linextrap, dphi, dex);
}

extern "C" int rip_lin_eval(int device, const void* z, int z_dtype, const float* coefs, int P, long npix, int linextrap,
                            float* phi, uint8_t* exflag) {
    RIP_API_BEGIN
    RIP_REQUIRE(P >= 1 && P <= RIP_PMAX, "rip_lin_eval: P=%d outside 1..%d", P, RIP_PMAX);
    RIP_REQUIRE(z_dtype == RIP_F32 || z_dtype == RIP_F64, "rip_lin_eval: z dtype must be f32 or f64");
    if (npix == 0) return 0;
    use_device(device);
    DevRaw dz;
    dz.upload(z, npix * dtype_size(z_dtype));
    DevBuf<float> dc, dphi(npix);
    dc.upload(coefs, (size_t)P * npix);
    DevBuf<uint8_t> dex(npix);
    if (P <= 4) lin_eval_dispatch<4>(dz.p, z_dtype, dc.p, P, npix, linextrap, dphi.p, dex.p);
    else if (P <= 11) lin_eval_dispatch<11>(dz.p, z_dtype, dc.p, P, npix, linextrap, dphi.p, dex.p);
    else lin_eval_dispatch<RIP_PMAX>(dz.p, z_dtype, dc.p, P, npix, linextrap, dphi.p, dex.p);
    dphi.download(phi, npix);
    if (exflag) dex.download(exflag, npix);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_multilin(int device, const float* S, int G, long npix, const float* coefs, int P, const float* Smin,
                            const float* Smax, const float* Sref, const uint32_t* lin_dq, const uint8_t* attempt,
                            int do_not_flag_first, int single_frame, float* phi, uint32_t* dq_out) {
    RIP_API_BEGIN
    RIP_REQUIRE(P >= 1 && P <= RIP_PMAX, "rip_multilin: P=%d outside 1..%d", P, RIP_PMAX);
    RIP_REQUIRE(G >= 1 && G <= RIP_GMAX, "rip_multilin: G=%d outside 1..%d", G, RIP_GMAX);
    RIP_REQUIRE(!single_frame || G == 1, "rip_multilin: single_frame needs G=1");
    if (npix == 0) return 0;
    use_device(device);
    DevBuf<float> dS, dc, dmin, dmax, dref, dphi((size_t)G * npix);
    DevBuf<uint32_t> ddq, dout(npix);
    DevBuf<uint8_t> datt;
    dS.upload(S, (size_t)G * npix);
    dc.upload(coefs, (size_t)P * npix);
    dmin.upload(Smin, npix);
    dmax.upload(Smax, npix);
    dref.upload(Sref, npix);
    ddq.upload(lin_dq, npix);
    if (attempt) datt.upload(attempt, (size_t)G * npix);
#define ML(PM)                                                                                                      \
    RIP_LAUNCH((multilin_kernel<RIP_GMAX, PM>), nblk(npix, 128), 128, 0, 0, dS.p, G, npix, dc.p, P, dmin.p, dmax.p, \
               dref.p, ddq.p, datt.p, do_not_flag_first, single_frame, dphi.p, dout.p)
    if (P <= 4) ML(4);
    else if (P <= 11) ML(11);
    else ML(RIP_PMAX);
#undef ML
    dphi.download(phi, (size_t)G * npix);
    dout.download(dq_out, npix);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_invlinearity(int device, const void* Slin, int dtype, long npix, const float* coefs, int P,
                                const float* Smin, const float* Smax, void* S_out, uint8_t* exflag) {
    RIP_API_BEGIN
    RIP_REQUIRE(P >= 1 && P <= RIP_PMAX, "rip_invlinearity: P=%d outside 1..%d", P, RIP_PMAX);
    RIP_REQUIRE(dtype == RIP_F32 || dtype == RIP_F64, "rip_invlinearity: dtype must be f32 or f64");
    if (npix == 0) return 0;
    use_device(device);
    const size_t es = dtype_size(dtype);
    DevRaw din, dout;
    din.upload(Slin, npix * es);
    dout.alloc(npix * es);
    DevBuf<float> dc, dmin, dmax;
    dc.upload(coefs, (size_t)P * npix);
    dmin.upload(Smin, npix);
    dmax.upload(Smax, npix);
    DevBuf<uint8_t> dex(npix);
#define IL(T, PM) RIP_LAUNCH((invlin_kernel<T, PM>), nblk(npix, 128), 128, 0, 0, (const T*)din.p, npix, dc.p, P, dmin.p, dmax.p, (T*)dout.p, dex.p)
    if (dtype == RIP_F64) { if (P <= 4) IL(double, 4); else if (P <= 11) IL(double, 11); else IL(double, RIP_PMAX); }
    else { if (P <= 4) IL(float, 4); else if (P <= 11) IL(float, 11); else IL(float, RIP_PMAX); }
#undef IL
    RIP_CUDA(cudaMemcpy(S_out, dout.p, npix * es, cudaMemcpyDeviceToHost));
    if (exflag) dex.download(exflag, npix);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

// ---- generic ipc_fwd / ipc_rev with NumPy's dtype promotion --------------------------------------------
namespace {
struct IpcCtx {
    int ny, nx;
    long npix;
    DevRaw img, K, gain, im2, a, b;
    bool imgd, kd, gd, has_gain;
    bool timd, tid_;  // TIM / TI are double?
};

template <typename TIMG, typename TG, typename TIM>
void run_mulgain(IpcCtx& c) {
    RIP_LAUNCH((mulgain_kernel<TIMG, TG, TIM>), nblk(c.npix, 256), 256, 0, 0, (const TIMG*)c.img.p,
               (const TG*)(c.has_gain ? c.gain.p : nullptr), c.npix, (TIM*)c.im2.p);
}

void make_im2(IpcCtx& c) {
    c.im2.alloc(c.npix * (c.timd ? 8 : 4));
    if (!c.timd) run_mulgain<float, float, float>(c);
    else if (c.imgd && (!c.has_gain || c.gd)) run_mulgain<double, double, double>(c);
    else if (c.imgd) run_mulgain<double, float, double>(c);
    else run_mulgain<float, double, double>(c);  // image f32, gain f64
}

void setup(IpcCtx& c, const void* image, int img_dtype, int ny, int nx, const void* kernel, int k_dtype, const void* gain,
           int g_dtype) {
    RIP_REQUIRE(img_dtype == RIP_F32 || img_dtype == RIP_F64, "ipc: image dtype must be f32 or f64");
    RIP_REQUIRE(k_dtype == RIP_F32 || k_dtype == RIP_F64, "ipc: kernel dtype must be f32 or f64");
    RIP_REQUIRE(!gain || g_dtype == RIP_F32 || g_dtype == RIP_F64, "ipc: gain dtype must be f32 or f64");
    c.ny = ny; c.nx = nx; c.npix = (long)ny * nx;
    c.imgd = img_dtype == RIP_F64; c.kd = k_dtype == RIP_F64; c.has_gain = gain != nullptr; c.gd = c.has_gain && g_dtype == RIP_F64;
    c.timd = c.imgd || c.gd;
    c.tid_ = c.timd || c.kd;
    c.img.upload(image, c.npix * (c.imgd ? 8 : 4));
    c.K.upload(kernel, 9 * c.npix * (c.kd ? 8 : 4));
    if (c.has_gain) c.gain.upload(gain, c.npix * (c.gd ? 8 : 4));
    make_im2(c);
}

template <typename TP, typename TIM, typename TK>
void run_iter(IpcCtx& c, const void* prev, void* out) {
    dim3 grid((c.nx + 127) / 128, c.ny), block(128);
    RIP_LAUNCH((ipc_rev_iter_kernel<TP, TIM, TK>), grid, block, 0, 0, (const TP*)prev, (const TIM*)c.im2.p, (const TK*)c.K.p,
               c.ny, c.nx, (typename Promote<TP, TK>::type*)out);
}
}  // namespace

extern "C" int rip_ipc_fwd(int device, const void* image, int img_dtype, int ny, int nx, const void* kernel, int k_dtype,
                           const void* gain, int g_dtype, void* out, int* out_dtype) {
    RIP_API_BEGIN
    use_device(device);
    IpcCtx c;
    setup(c, image, img_dtype, ny, nx, kernel, k_dtype, gain, g_dtype);
    c.a.alloc(c.npix * (c.tid_ ? 8 : 4));
    dim3 grid((nx + 127) / 128, ny), block(128);
    const void* g = c.has_gain ? c.gain.p : nullptr;
    // out /= gain happens in the output dtype (in-place NumPy division)
#define FW(TIM, TK, TG) RIP_LAUNCH((ipc_fwd_kernel<TIM, TK, TG>), grid, block, 0, 0, (const TIM*)c.im2.p, (const TK*)c.K.p, (const TG*)g, ny, nx, (Promote<TIM, TK>::type*)c.a.p)
    if (!c.timd && !c.kd) FW(float, float, float);
    else if (!c.timd && c.kd) FW(float, double, float);
    else if (c.timd && !c.kd) { if (c.gd || !c.has_gain) FW(double, float, double); else FW(double, float, float); }
    else { if (c.gd || !c.has_gain) FW(double, double, double); else FW(double, double, float); }
#undef FW
    RIP_CUDA(cudaMemcpy(out, c.a.p, c.npix * (c.tid_ ? 8 : 4), cudaMemcpyDeviceToHost));
    if (out_dtype) *out_dtype = c.tid_ ? RIP_F64 : RIP_F32;
    RIP_API_END
}

extern "C" int rip_ipc_rev(int device, const void* image, int img_dtype, int ny, int nx, const void* kernel, int k_dtype,
                           int order, const void* gain, int g_dtype, void* out, int* out_dtype) {
    RIP_API_BEGIN
    RIP_REQUIRE(order >= 0 && order <= 16, "rip_ipc_rev: order %d outside 0..16", order);
    use_device(device);
    IpcCtx c;
    setup(c, image, img_dtype, ny, nx, kernel, k_dtype, gain, g_dtype);
    const bool outd = (order == 0) ? c.timd : c.tid_;
    c.a.alloc(c.npix * 8);
    c.b.alloc(c.npix * 8);
    const void* prev = c.im2.p;
    void* cur = c.a.p;
    for (int it = 0; it < order; ++it) {
        const bool prevd = (it == 0) ? c.timd : c.tid_;
        if (!prevd && !c.kd) run_iter<float, float, float>(c, prev, cur);  // all f32
        else if (!prevd && c.kd) run_iter<float, float, double>(c, prev, cur);  // first iteration, f64 kernel
        else if (prevd && !c.timd) { if (c.kd) run_iter<double, float, double>(c, prev, cur); else run_iter<double, float, float>(c, prev, cur); }
        else { if (c.kd) run_iter<double, double, double>(c, prev, cur); else run_iter<double, double, float>(c, prev, cur); }
        prev = cur;
        cur = (cur == c.a.p) ? c.b.p : c.a.p;
    }
    if (c.has_gain) {
        void* res = const_cast<void*>(prev);
        if (outd) { if (c.gd) RIP_LAUNCH((divgain_kernel<double, double>), nblk(c.npix, 256), 256, 0, 0, (double*)res, (const double*)c.gain.p, c.npix);
                    else RIP_LAUNCH((divgain_kernel<double, float>), nblk(c.npix, 256), 256, 0, 0, (double*)res, (const float*)c.gain.p, c.npix); }
        else RIP_LAUNCH((divgain_kernel<float, float>), nblk(c.npix, 256), 256, 0, 0, (float*)res, (const float*)c.gain.p, c.npix);
    }
    RIP_CUDA(cudaMemcpy(out, prev, c.npix * (outd ? 8 : 4), cudaMemcpyDeviceToHost));
    if (out_dtype) *out_dtype = outd ? RIP_F64 : RIP_F32;
    RIP_API_END
}

extern "C" int rip_correct_cube(int device, float* data, int G, int ny, int nx, const void* kernel, int k_dtype, int nya,
                                int nxa, const void* gain_full, int g_dtype) {
    RIP_API_BEGIN
    RIP_REQUIRE(k_dtype == RIP_F32 || k_dtype == RIP_F64, "rip_correct_cube: kernel dtype must be f32 or f64");
    const int nb = (8192 + (nx - nxa) / 2) % 16;  // utils/ipc_linearity.py:177
    RIP_REQUIRE(ny - 2 * nb == nya && nx - 2 * nb == nxa, "rip_correct_cube: kernel shape (%d,%d) does not match data (%d,%d) minus border %d", nya, nxa, ny, nx, nb);
    use_device(device);
    const long npl = (long)ny * nx, npa = (long)nya * nxa;
    DevBuf<float> d;
    d.upload(data, (size_t)G * npl);
    DevRaw K, g, tmp;
    K.upload(kernel, 9 * npa * dtype_size(k_dtype));
    if (gain_full) g.upload(gain_full, npl * dtype_size(g_dtype));
    tmp.alloc(npa * 8);
    for (int i = 0; i < G; ++i)
        launch_ipc_rev_dn(d.p + (size_t)i * npl, ny, nx, nb, K.p, k_dtype, gain_full ? g.p : nullptr, g_dtype, false, 0.f, tmp.p, 0);
    d.download(data, (size_t)G * npl);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

// ---- reference-pixel stage functions ----------------------------------------------------------------------
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

extern "C" int rip_row_medians(int device, const float* image, int n, int ncols, int use_ref_channel, float* ref_med,
                               float* sci_med) {
    RIP_API_BEGIN
    RIP_REQUIRE(n >= 16 && ncols >= n, "rip_row_medians: bad shape");
    RIP_REQUIRE(!use_ref_channel || ncols >= n + 128, "rip_row_medians: use_ref_channel needs ncols >= n+128");
    use_device(device);
    DevBuf<float> img, out(n);
    img.upload(image, (size_t)n * ncols);
    if (use_ref_channel) RIP_LAUNCH(row_median_kernel, n, 128, 128 * 4, 0, img.p, ncols, n, 128, 0, 0, 128, out.p);
    else RIP_LAUNCH(row_median_kernel, n, 32, 8 * 4, 0, img.p, ncols, 0, 4, n - 4, 4, 8, out.p);
    out.download(ref_med, n);
    if (sci_med) {
        DevBuf<float> o2(n);
        const int cnt = n - 8, np2 = next_pow2(cnt);
        RIP_REQUIRE(np2 * 4 <= 48 * 1024, "rip_row_medians: row too long");
        RIP_LAUNCH(row_median_kernel, n, 512, np2 * 4, 0, img.p, ncols, 4, cnt, 0, 0, np2, o2.p);
        o2.download(sci_med, n);
        RIP_CUDA(cudaDeviceSynchronize());
    }
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_refsub_row_apply(int device, float* image, int n, int ncols, double m_med, const float* ref_med) {
    RIP_API_BEGIN
    use_device(device);
    DevBuf<float> img, rm, ctr(1);
    img.upload(image, (size_t)n * ncols);
    rm.upload(ref_med, n);
    const int np2 = next_pow2(n);
    RIP_REQUIRE(np2 * 4 <= 48 * 1024, "rip_refsub_row_apply: n too large");
    RIP_LAUNCH(vec_median_kernel, 1, 1024, np2 * 4, 0, rm.p, n, np2, ctr.p);
    dim3 grid((ncols + 255) / 256, n);
    RIP_LAUNCH(refsub_row_apply_kernel, grid, 256, 0, 0, img.p, n, ncols, m_med, rm.p, ctr.p);
    img.download(image, (size_t)n * ncols);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_refsub_channel(int device, float* image, int n, int ncols, int nchan) {
    RIP_API_BEGIN
    RIP_REQUIRE(nchan >= 1 && nchan * 128 <= ncols, "rip_refsub_channel: %d channels do not fit %d columns", nchan, ncols);
    use_device(device);
    DevBuf<float> img, med((size_t)nchan * 2);
    img.upload(image, (size_t)n * ncols);
    RIP_LAUNCH(chan_median_kernel, dim3(nchan, 2), 256, 0, 0, img.p, n, ncols, med.p);
    dim3 grid((nchan * 128 + 255) / 256, n);
    RIP_LAUNCH(refsub_chan_apply_kernel, grid, 256, 0, 0, img.p, n, ncols, nchan, med.p);
    img.download(image, (size_t)n * ncols);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_get_flat(int device, const float* flat, int n, int nb, const void* gain, int g_dtype, const void* kernel,
                            int k_dtype, uint32_t* pdq, int ipc_deconvolve, float* out) {
    RIP_API_BEGIN
    RIP_REQUIRE(!ipc_deconvolve || (gain && kernel), "rip_get_flat: ipc_deconvolve needs gain and ipc4d");
    use_device(device);
    const long npl = (long)n * n, npa = (long)(n - 2 * nb) * (n - 2 * nb);
    DevBuf<float> f, o(npl);
    f.upload(flat, npl);
    DevBuf<uint32_t> dq;
    if (pdq) dq.upload(pdq, npl);
    DevRaw g, K, tmp;
    if (gain) g.upload(gain, npl * dtype_size(g_dtype));
    launch_flat_prepare(f.p, n, nb, gain ? g.p : nullptr, g_dtype, pdq ? dq.p : nullptr, ipc_deconvolve, o.p, 0);
    if (ipc_deconvolve) {
        K.upload(kernel, 9 * npa * dtype_size(k_dtype));
        tmp.alloc(npa * 8);
        // flatutils.py:69 clips the gain only when pdq is given
        launch_ipc_rev_dn(o.p, n, n, nb, K.p, k_dtype, g.p, g_dtype, pdq != nullptr, 0.1f, tmp.p, 0);
    }
    o.download(out, npl);
    if (pdq) dq.download(pdq, npl);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}

extern "C" int rip_flag_saturation(int device, const uint16_t* raw, int G, int n, const float* sat_thresh,
                                   const uint32_t* sat_dq, int backup, int skip_firstn, uint8_t* rdq, uint32_t* pdq) {
    RIP_API_BEGIN
    RIP_REQUIRE(G >= 1 && G <= RIP_GMAX, "rip_flag_saturation: G=%d outside 1..%d", G, RIP_GMAX);
    RIP_REQUIRE(backup >= 0 && skip_firstn >= 0, "rip_flag_saturation: negative backup/skip");
    use_device(device);
    const long npl = (long)n * n;
    DevBuf<uint16_t> r;
    r.upload(raw, (size_t)G * npl);
    DevBuf<float> t;
    t.upload(sat_thresh, npl);
    DevBuf<uint32_t> sd, bits(npl), pd;
    sd.upload(sat_dq, npl);
    pd.upload(pdq, npl);
    DevBuf<uint8_t> q;
    q.upload(rdq, (size_t)G * npl);
    RIP_LAUNCH(sat_bits_kernel, nblk(npl, 256), 256, 0, 0, r.p, G, n, t.p, sd.p, skip_firstn, bits.p, pd.p);
    RIP_LAUNCH(sat_grow_kernel, dim3((n + 255) / 256, n), 256, 0, 0, bits.p, G, n, backup, skip_firstn, q.p);
    q.download(rdq, (size_t)G * npl);
    pd.download(pdq, npl);
    RIP_CUDA(cudaDeviceSynchronize());
    RIP_API_END
}
