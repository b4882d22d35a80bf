// CALDIR handle (device-resident calibration planes of one SCA), static products, reference-pixel
// statistics and the fused L1->L2 entry points (include/rip_b200.h).
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>

#include "rip_handle.h"
#include "rip_launch.h"
#include "rip_v2_core.cuh"

namespace rip {

// =========================================================================================================
// K2: static, exposure-independent products
// =========================================================================================================
__global__ void static_planes_kernel(int n, int nb, const float* __restrict__ sat_thr, const uint32_t* __restrict__ sat_dq,
                                     const uint32_t* __restrict__ lin_dq, const uint32_t* __restrict__ mask_dq,
                                     const uint32_t* __restrict__ dark_dq, float* __restrict__ thr_eff,
                                     uint8_t* __restrict__ aux, uint32_t* __restrict__ sdq) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long)n * n) return;
    const int y = (int)(p / n), x = (int)(p % n);
    const bool active = (y >= nb && y < n - nb && x >= nb && x < n - nb);
    float t = sat_thr[p];
    const uint32_t sd = sat_dq[p];
    if ((sd & DQ_NO_SAT_CHECK) || t != t) t = INFINITY;
    thr_eff[p] = t;
    const uint32_t ld = lin_dq[p];
    uint32_t md = mask_dq ? mask_dq[p] : 0u;
    if (mask_dq) {
        // do_dqinit(..., expand_gw_flagging=1) (gen_cal_image.py:118): GW_AFFECTED_DATA grown by one pixel before it
        // enters pixeldq.  Restated from upstream (binary dilation, one iteration, 4-connected structure).
        uint32_t nbq = 0u;
        if (y > 0) nbq |= mask_dq[p - n];
        if (y < n - 1) nbq |= mask_dq[p + n];
        if (x > 0) nbq |= mask_dq[p - 1];
        if (x < n - 1) nbq |= mask_dq[p + 1];
        md |= nbq & DQ_GW_AFFECTED_DATA;
    }
    uint8_t a = 0;
    if (ld & (DQ_NO_LIN_CORR | DQ_REFERENCE_PIXEL)) a |= 1;
    if ((ld | md) & DQ_REFERENCE_PIXEL) a |= 2;
    aux[p] = a;
    // pdq sources that do not depend on the exposure: mask (do_dqinit), NO_SAT_CHECK (flag_saturation), lin dq
    // (ipc_linearity.py:328), dark dq on the active region (subtract_dark_current); flat flags are OR'ed later
    sdq[p] |= md | (sd & DQ_NO_SAT_CHECK) | ld | ((active && dark_dq) ? dark_dq[p] : 0u);
}

// =========================================================================================================
// K0: reference-pixel statistics (gen_cal_image.py:531-555 + utils/reference_subtraction.py)
// =========================================================================================================
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}


// ---- radix select of the two middle order statistics of e = f32(amp33) - med over the whole reference-output frame:
// 3 histogram passes over order-preserving keys (11 + 11 + 10 bits).  The block that finishes a pass last (ticket
// counter) also locates the buckets of the two ranks, extends their prefixes and clears the histogram, so a pass
// is ONE launch; pass 0 is fused into the per-row sorting kernel below.

// executed by all 256 threads of one block; hist = this group's [2][2048] counters
__device__ void k0_scan_block(uint32_t* __restrict__ hist, SelState* __restrict__ st, int pass, uint32_t M, uint32_t* sh /*[256]*/) {
    const int tid = threadIdx.x;
    const int nbits = (pass == 2) ? 10 : 11, nbin = 1 << nbits;
    for (int r = 0; r < 2; ++r) {
        const uint32_t* h = hist + ((pass == 0) ? 0 : r * 2048);  // pass 0: both ranks share the (empty) prefix
        uint32_t loc[8], sum = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int bin = tid * 8 + k;
            loc[k] = (bin < nbin) ? __ldcg(h + bin) : 0u;
            sum += loc[k];
        }
        sh[tid] = sum;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {  // Hillis-Steele inclusive scan
            const uint32_t v = (tid >= off) ? sh[tid - off] : 0u;
            __syncthreads();
            sh[tid] += v;
            __syncthreads();
        }
        uint32_t before = sh[tid] - sum;
        const uint32_t rank = (pass == 0) ? (M / 2 - 1 + r) : st->rank[r];  // M even: np.median averages the two middle values
        const uint32_t prefix = (pass == 0) ? 0u : st->prefix[r];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (rank >= before && rank < before + loc[k]) {
                st->prefix[r] = (prefix << nbits) | (uint32_t)(tid * 8 + k);
                st->rank[r] = rank - before;
            }
            before += loc[k];
        }
        __syncthreads();
    }
    for (int i = tid; i < 4096; i += 256) hist[i] = 0u;
}

// flush a block's shared histogram to the group's global one; the last block of the group runs the scan
__device__ void k0_flush_and_scan(uint32_t* sh_hist /*[2][2048]*/, int nhist, uint32_t* __restrict__ ghist, SelState* __restrict__ st,
                                  uint32_t* __restrict__ ticket, int nblocks, int pass, uint32_t M) {
    __shared__ uint32_t scan_sh[256];
    __shared__ int is_last;
    __syncthreads();
    for (int i = threadIdx.x; i < nhist * 2048; i += blockDim.x) {
        const uint32_t v = sh_hist[i];
        if (v) atomicAdd(&ghist[i], v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == (uint32_t)nblocks - 1u);
    __syncthreads();
    if (is_last) {
        __threadfence();
        k0_scan_block(ghist, st, pass, M, scan_sh);
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// passes 1 and 2: elements must match the already-known high bits (prefix) of each of the two order statistics
__global__ void __launch_bounds__(256) k0_hist_kernel(const uint16_t* __restrict__ amp33, const float* __restrict__ med, long M,
                                                      int pass, SelState* __restrict__ st, uint32_t* __restrict__ hist /*[G][2][2048]*/,
                                                      uint32_t* __restrict__ ticket /*[G]*/) {
    __shared__ uint32_t sh[2][2048];
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const SelState s = st[g];
    const int shift_hi = (pass == 1) ? 21 : 10;
    const int shift = (pass == 1) ? 10 : 0;
    const uint32_t mask = (pass == 2) ? 1023u : 2047u;
    const uint16_t* a = amp33 + (long)g * M;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long)gridDim.x * blockDim.x) {
        const float e = (float)a[i] - med[i];
        const uint32_t k = f2key(e);
        const uint32_t hi = k >> shift_hi;
        const uint32_t b = (k >> shift) & mask;
        if (hi == s.prefix[0]) atomicAdd(&sh[0][b], 1u);
        if (hi == s.prefix[1]) atomicAdd(&sh[1][b], 1u);
    }
    k0_flush_and_scan(&sh[0][0], 2, hist + (long)g * 4096, st + g, ticket + g, gridDim.x, pass, (uint32_t)M);
}

// One warp per row of the reference output: bitonic sort of its 128 values e = f32(amp33) - med (4 per lane, element
// index i = 32 k + lane) -> the two middle order statistics (ranks 63, 64); plus pass 0 of the global radix select.
constexpr int K0_ROWS = 32;
__global__ void __launch_bounds__(256) k0_rows_kernel(const uint16_t* __restrict__ amp33, const float* __restrict__ med, int n,
                                                      float* __restrict__ rowA, float* __restrict__ rowB, SelState* __restrict__ st,
                                                      uint32_t* __restrict__ hist, uint32_t* __restrict__ ticket) {
    __shared__ uint32_t sh[2048];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    // K0_ROWS rows per CTA, 8 at a time (one per warp): 4x fewer histogram flushes (2048 global atomics each) than one
    // pass of 8 rows per CTA
    for (int it = 0; it < K0_ROWS / 8; ++it) {
    const int row = blockIdx.x * K0_ROWS + it * 8 + w;
    if (row < n) {
        const uint16_t* a = amp33 + ((long)g * n + row) * 128;
        const float* m = med + (long)row * 128;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = (float)a[lane + 32 * k] - m[lane + 32 * k];
            atomicAdd(&sh[f2key(v[k]) >> 21], 1u);
        }
#pragma unroll
        for (int k2 = 2; k2 <= 128; k2 <<= 1) {
#pragma unroll
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                if (j >= 32) {  // partner in the same lane: registers k and k ^ (j / 32)
                    const int dk = j >> 5;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if ((k & dk) == 0) {
                            const bool up = (((k * 32) & k2) == 0);  // (i & k2) with i = 32 k + lane, k2 >= 64 here
                            const float lo = fminf(v[k], v[k | dk]), hi = fmaxf(v[k], v[k | dk]);
                            v[k] = up ? lo : hi;
                            v[k | dk] = up ? hi : lo;
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float o = __shfl_xor_sync(0xffffffffu, v[k], j);
                        const bool lower = (lane & j) == 0;
                        const bool up = (((k * 32 + lane) & k2) == 0);
                        v[k] = (lower == up) ? fminf(v[k], o) : fmaxf(v[k], o);
                    }
                }
            }
        }
        if (lane == 31) rowA[(long)g * n + row] = v[1];  // rank 63 = 32*1 + 31
        if (lane == 0) rowB[(long)g * n + row] = v[2];   // rank 64 = 32*2 + 0
    }
    }
    k0_flush_and_scan(sh, 1, hist + (long)g * 4096, st + g, ticket + g, gridDim.x, 0, (uint32_t)n * 128u);
}

__device__ __forceinline__ void bitonic_sort_block(float* v, int npow2) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = v[i], b = v[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { v[i] = b; v[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// np.median of v[0..count) (shared memory) without sorting: four 8-bit radix passes over order-preserving keys,
// both middle ranks tracked at once (warp 0 / warp 1 scan their 256-bin histogram with a shuffle prefix sum).
// Every thread of the CTA calls it (blockDim >= 64); NaN anywhere gives NaN, as NumPy does.  ws: shared uint32[520].
__device__ float block_median_radix(const float* v, int count, uint32_t* ws) {
    uint32_t* hist = ws;          // [2][256]
    uint32_t* st = ws + 512;      // prefix[2], rank[2], mask, nan
    if (threadIdx.x == 0) {
        st[0] = st[1] = 0u;
        st[2] = (uint32_t)((count - 1) / 2);
        st[3] = (uint32_t)(count / 2);
        st[4] = 0u;
        st[5] = 0u;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const uint32_t m = st[4], p0 = st[0], p1 = st[1];
        for (int i = threadIdx.x; i < count; i += blockDim.x) {
            const float x = v[i];
            if (x != x) { st[5] = 1u; continue; }
            const uint32_t k = f2key(x), b = (k >> shift) & 255u;
            if ((k & m) == p0) atomicAdd(&hist[b], 1u);
            if ((k & m) == p1) atomicAdd(&hist[256 + b], 1u);
        }
        __syncthreads();
        if (warp < 2) {
            const uint32_t* h = hist + 256 * warp;
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = h[lane * 8 + q]; tot += c[q]; }
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t excl = incl - tot, r = st[2 + warp];
            const bool mine = (r >= excl) && (r < incl);
            if (mine) {
                uint32_t acc = excl;
                int q = 0;
                for (; q < 7; ++q) {
                    if (acc + c[q] > r) break;
                    acc += c[q];
                }
                st[2 + warp] = r - acc;
                st[warp] |= (uint32_t)(lane * 8 + q) << shift;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) st[4] |= 255u << shift;
        __syncthreads();
    }
    const float a = key2f(st[0]), b = key2f(st[1]);
    float med = (count & 1) ? a : (a + b) / 2.0f;
    if (st[5]) med = __int_as_float(0x7fc00000);
    __syncthreads();  // ws may be reused
    return med;
}

// global median, per-row reference medians, their median, and the f64 row correction
__global__ void k0_final_kernel(const SelState* __restrict__ st, const float* __restrict__ rowA,
                                const float* __restrict__ rowB, int n, int npow2, double slope,
                                double* __restrict__ rowcorr, float* __restrict__ gmed_out) {
    extern __shared__ float sv[];
    float* refm = sv + npow2;
    __shared__ uint32_t ws[520];
    const int g = blockIdx.x;
    const float kA = key2f(st[g].prefix[0]), kB = key2f(st[g].prefix[1]);
    const float gmed = (kA + kB) / 2.0f;
    if (threadIdx.x == 0 && gmed_out) gmed_out[g] = gmed;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        float v = INFINITY;
        if (i < n) {
            const float a = rowA[(long)g * n + i] - gmed, b = rowB[(long)g * n + i] - gmed;
            v = (a + b) / 2.0f;
            refm[i] = v;
        }
        sv[i] = v;
    }
    __syncthreads();
    const float ctr = block_median_radix(refm, n, ws);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float dm = refm[i] - ctr;
        rowcorr[(long)g * n + i] = slope * (double)dm;
    }
}

// per (group, channel): medians of the 4 bottom / 4 top reference rows of the row-corrected (data - dark), and
// the line through (1.5, bottom), (n-2.5, top)
__global__ void k0_chan_kernel(const uint16_t* __restrict__ raw, const float* __restrict__ dark, int n,
                               const double* __restrict__ rowcorr, double* __restrict__ chan_m,
                               double* __restrict__ chan_c, double* __restrict__ chan_line /*[G,32,n]*/) {
    __shared__ float sv[2][512];
    __shared__ float meds[2];
    __shared__ double line_mc[2];
    __shared__ uint32_t ws[520];
    const int ch = blockIdx.x, g = blockIdx.y;
    const long npl = (long)n * n;
    for (int side = 0; side < 2; ++side) {
        const int rbase = side ? (n - 4) : 0;
        for (int i = threadIdx.x; i < 512; i += blockDim.x) {
            const int row = rbase + (i >> 7), x = ch * 128 + (i & 127);
            const long p = (long)row * n + x;
            float v = (float)raw[(long)g * npl + p] - dark[(long)g * npl + p];
            v = (float)((double)v - rowcorr[(long)g * n + row]);
            sv[side][i] = v;
        }
    }
    __syncthreads();
    for (int side = 0; side < 2; ++side) {
        const float md = block_median_radix(sv[side], 512, ws);
        if (threadIdx.x == 0) meds[side] = md;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double b = (double)meds[0], t = (double)meds[1];
        const double x0 = 1.5, x1 = (double)n - 2.5;
        const double m = (t - b) / (x1 - x0);
        chan_m[g * 32 + ch] = m;
        chan_c[g * 32 + ch] = b - m * x0;
        line_mc[0] = m;
        line_mc[1] = b - m * x0;
    }
    __syncthreads();
    // the line itself, tabulated for the fused v2 kernel (same unfused f64 expression as v1: m * row + c)
    const double m = line_mc[0], c = line_mc[1];
    for (int row = threadIdx.x; row < n; row += blockDim.x) chan_line[((long)g * 32 + ch) * n + row] = m * (double)row + c;
}

}  // namespace rip

using namespace rip;

// =========================================================================================================
// the handle
// =========================================================================================================
static void run_k0(rip_caldir* h, rip_caldir::K0Work& W, const uint16_t* d_raw, const uint16_t* d_amp33, int G, cudaStream_t st) {
    const int n = h->n;
    RIP_REQUIRE(h->has_amp33, "reference-pixel correction needs amp33 statistics in the read file (the reference's np.polyfit fails without them: SURVEY 7)");
    RIP_REQUIRE(n % 128 == 0 && n <= 4096 && n >= 128, "reference-pixel correction needs a frame side that is a multiple of 128 in 128..4096 (got %d)", n);
    RIP_REQUIRE(d_amp33 != nullptr, "reference-pixel correction needs the amp33 cube");
    const long M = (long)n * 128;
    const int nch = n / 128;
    if (W.hist.n < (size_t)G * 4096) {
        W.hist.alloc((size_t)RIP_GMAX * 4096);
        W.sel.alloc(RIP_GMAX);
        W.k0_ticket.alloc(RIP_GMAX);
        W.k0_ticket.zero(st);
        W.hist.zero(st);  // (histograms and tickets are left zeroed by every pass)
        W.rowA.alloc((size_t)RIP_GMAX * n);
        W.rowB.alloc((size_t)RIP_GMAX * n);
        W.gmed.alloc(RIP_GMAX);
        W.rowcorr.alloc((size_t)RIP_GMAX * n);
        W.chan_m.alloc((size_t)RIP_GMAX * 32);
        W.chan_c.alloc((size_t)RIP_GMAX * 32);
        W.chan_line.alloc((size_t)RIP_GMAX * 32 * n);
    }
    RIP_LAUNCH(k0_rows_kernel, dim3((n + K0_ROWS - 1) / K0_ROWS, G), 256, 0, st, d_amp33, h->amp_med.p, n, W.rowA.p, W.rowB.p, W.sel.p,
               W.hist.p, W.k0_ticket.p);
    const int nblk = (int)std::min<long>(64, (M + 4095) / 4096);
    for (int pass = 1; pass < 3; ++pass)
        RIP_LAUNCH(k0_hist_kernel, dim3(nblk, G), 256, 0, st, d_amp33, h->amp_med.p, M, pass, W.sel.p, W.hist.p, W.k0_ticket.p);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    RIP_LAUNCH(k0_final_kernel, G, 1024, (size_t)2 * np2 * sizeof(float), st, W.sel.p, W.rowA.p, W.rowB.p, n, np2,
               h->refout_slope, W.rowcorr.p, W.gmed.p);
    RIP_LAUNCH(k0_chan_kernel, dim3(nch, G), 256, 0, st, d_raw, h->dark_cube.p, n, W.rowcorr.p, W.chan_m.p, W.chan_c.p, W.chan_line.p);
}

static double derive_refout_slope(const rip_caldir_desc* d) {
    // gen_cal_image.py:542-553 (the Python layer normally supplies this with the reference's own expression)
    const size_t m = (size_t)d->n * 128;
    std::vector<float> v(d->amp33_std, d->amp33_std + m);
    std::nth_element(v.begin(), v.begin() + m / 2, v.end());
    const float hi = v[m / 2];
    const float lo = *std::max_element(v.begin(), v.begin() + m / 2);
    const float med = (m & 1) ? hi : (lo + hi) / 2.0f;
    const float t = (med * med) / 128.0f;
    const double cvar = d->c_pink * d->c_pink;
    return d->m_pink * cvar / (d->m_pink * d->m_pink * cvar + d->ru_pink * d->ru_pink + (double)t / std::log(4096.0));
}

extern "C" int rip_caldir_create(int device, const rip_caldir_desc* d, rip_caldir** out) {
    RIP_API_BEGIN
    RIP_REQUIRE(d && out, "rip_caldir_create: null argument");
    RIP_REQUIRE(d->n >= 16 && d->nb >= 1 && d->n > 2 * d->nb + 4, "rip_caldir_create: bad geometry n=%d nb=%d", d->n, d->nb);
    RIP_REQUIRE(d->P >= 1 && d->P <= RIP_PMAX, "rip_caldir_create: P=%d outside 1..%d", d->P, RIP_PMAX);
    RIP_REQUIRE(d->lin_coefs && d->Smin && d->Smax && d->Sref && d->lin_dq, "rip_caldir_create: linearitylegendre planes missing");
    RIP_REQUIRE(d->sat_thresh && d->sat_dq, "rip_caldir_create: saturation planes missing");
    RIP_REQUIRE(d->gain && d->read && d->dark_slope && d->flat, "rip_caldir_create: gain/read/dark_slope/flat missing");
    RIP_REQUIRE(d->gain_dtype == RIP_F32 || d->gain_dtype == RIP_F64, "rip_caldir_create: gain dtype");
    RIP_REQUIRE(!d->ipc || d->ipc_dtype == RIP_F32 || d->ipc_dtype == RIP_F64, "rip_caldir_create: ipc dtype");
    use_device(device);
    std::unique_ptr<rip_caldir> h(new rip_caldir);
    h->device = device;
    h->d = *d;
    const int n = h->n = d->n, nb = h->nb = d->nb, na = h->na = d->n - 2 * d->nb;
    h->P = d->P;
    const size_t npl = (size_t)n * n, npa = (size_t)na * na;
    RIP_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
        int pr_lo = 0, pr_hi = 0;
        RIP_CUDA(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
        // LOWEST priority: the look-ahead's CTAs then fill the slots the fused kernel leaves free in its last wave instead
        // of displacing its CTAs (measured: step 1.418 ms without look-ahead, 1.398 ms with a high-priority side stream,
        // 1.353 ms with this one -- profiles/r02/ab_refpix_lookahead.log)
        (void)pr_hi;
        RIP_CUDA(cudaStreamCreateWithPriority(&h->s_k0, cudaStreamNonBlocking, pr_lo));
        for (auto& w : h->k0w) {
            RIP_CUDA(cudaEventCreateWithFlags(&w.ev_k0, cudaEventDisableTiming));
            RIP_CUDA(cudaEventCreateWithFlags(&w.ev_used, cudaEventDisableTiming));
        }
    }
    cudaStream_t st = h->stream;
    h->coefs.upload(d->lin_coefs, (size_t)d->P * npl, st);
    h->Smin.upload(d->Smin, npl, st);
    h->Smax.upload(d->Smax, npl, st);
    h->Sref.upload(d->Sref, npl, st);
    h->lin_dq.upload(d->lin_dq, npl, st);
    h->sat_thr.upload(d->sat_thresh, npl, st);
    h->read.upload(d->read, npl, st);
    if (d->resetnoise) h->resetnoise.upload(d->resetnoise, npl, st);
    if (d->dark_cube && d->n_dark > 0) h->dark_cube.upload(d->dark_cube, (size_t)d->n_dark * npl, st);
    h->dark_slope.upload(d->dark_slope, npl, st);
    h->flat.upload(d->flat, npl, st);
    h->gain.upload(d->gain, npl * dtype_size(d->gain_dtype), st);
    h->has_ipc = d->ipc != nullptr;
    if (h->has_ipc) h->ipc.upload(d->ipc, 9 * npa * dtype_size(d->ipc_dtype), st);
    h->has_bias = d->biascorr != nullptr && d->n_bias > 0;
    if (h->has_bias) h->biascorr.upload(d->biascorr, (size_t)d->n_bias * npa, st);
    h->has_amp33 = d->has_amp33 && d->amp33_med != nullptr;
    if (h->has_amp33) {
        h->amp_med.upload(d->amp33_med, (size_t)n * 128, st);
        if (d->amp33_std) h->amp_std.upload(d->amp33_std, (size_t)n * 128, st);
        h->refout_slope = (d->refout_slope == d->refout_slope) ? d->refout_slope
                          : (d->amp33_std ? derive_refout_slope(d) : std::numeric_limits<double>::quiet_NaN());
        RIP_REQUIRE(h->refout_slope == h->refout_slope, "rip_caldir_create: cannot derive the reference-output slope (amp33 std missing)");
    }
    // ---- static products ----
    DevBuf<uint32_t> sat_dq, mask_dq, dark_dq;
    sat_dq.upload(d->sat_dq, npl, st);
    if (d->mask_dq) mask_dq.upload(d->mask_dq, npl, st);
    if (d->dark_dq) dark_dq.upload(d->dark_dq, npl, st);
    h->thr_eff.alloc(npl);
    h->aux.alloc(npl);
    h->sdq.alloc(npl);
    h->sdq.zero(st);
    h->flat_ipc.alloc(npl);
    h->dslope_ipc.alloc(npl);
    // flat first: its flags are OR'ed into sdq (utils/flatutils.py:55-68)
    launch_flat_prepare(h->flat.p, n, nb, h->gain.p, d->gain_dtype, h->sdq.p, h->has_ipc ? 1 : 0, h->flat_ipc.p, st);
    RIP_LAUNCH(static_planes_kernel, (unsigned)((npl + 255) / 256), 256, 0, st, n, nb, h->sat_thr.p, sat_dq.p, h->lin_dq.p,
               d->mask_dq ? mask_dq.p : nullptr, d->dark_dq ? dark_dq.p : nullptr, h->thr_eff.p, h->aux.p, h->sdq.p);
    RIP_CUDA(cudaMemcpyAsync(h->dslope_ipc.p, h->dark_slope.p, npl * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (h->has_ipc) {
        DevRaw tmp;
        tmp.alloc(npa * 8);
        launch_ipc_rev_dn(h->flat_ipc.p, n, n, nb, h->ipc.p, d->ipc_dtype, h->gain.p, d->gain_dtype, true, 0.1f, tmp.p, st);
        launch_ipc_rev_dn(h->dslope_ipc.p, n, n, nb, h->ipc.p, d->ipc_dtype, h->gain.p, d->gain_dtype, false, 0.f, tmp.p, st);
        RIP_CUDA(cudaStreamSynchronize(st));
    }
    RIP_CUDA(cudaStreamSynchronize(st));
    // pointers in the stored desc are the caller's: never dereference them again
    *out = h.release();
    RIP_API_END
}

extern "C" void rip_caldir_destroy(rip_caldir* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->s_k0) { cudaStreamSynchronize(h->s_k0); cudaStreamDestroy(h->s_k0); }
    if (h->stream) cudaStreamDestroy(h->stream);
    for (auto& w : h->k0w) {
        if (w.ev_k0) cudaEventDestroy(w.ev_k0);
        if (w.ev_used) cudaEventDestroy(w.ev_used);
    }
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    delete h;
}

extern "C" int rip_caldir_get_static(rip_caldir* h, float* dark_slope_ipc, float* flat_ipc, uint32_t* static_dq, double* refout_slope) {
    RIP_API_BEGIN
    RIP_REQUIRE(h, "rip_caldir_get_static: null handle");
    use_device(h->device);
    const size_t npl = (size_t)h->n * h->n;
    if (dark_slope_ipc) h->dslope_ipc.download(dark_slope_ipc, npl, h->stream);
    if (flat_ipc) h->flat_ipc.download(flat_ipc, npl, h->stream);
    if (static_dq) h->sdq.download(static_dq, npl, h->stream);
    if (refout_slope) *refout_slope = h->refout_slope;
    RIP_CUDA(cudaStreamSynchronize(h->stream));
    RIP_API_END
}

// =========================================================================================================
// fused L1 -> L2
// =========================================================================================================
// default organisation of the fused kernel (RIP_FUSED_VARIANT overrides, for A/B measurements): see rip_v2.cu
static int fused_default_variant() {
    static const int v = [] {
        const char* e = getenv("RIP_FUSED_VARIANT");
        return e ? atoi(e) : 4;  // 4 = v6 (five CTAs per SM) where it is instantiated, else v2; 0 forces v2
    }();
    return v;
}

static void l1_to_l2_dev_impl(rip_caldir* h, const uint16_t* d_raw, const uint16_t* d_amp33, const void* d_area,
                              const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                              const rip_l2_out* o, cudaStream_t st) {
    RIP_REQUIRE(h && d_raw && prm && plan && o, "rip_l1_to_l2: null argument");
    RIP_REQUIRE(o->slope && o->err_read && o->err_poisson && o->pdq, "rip_l1_to_l2: slope/err_read/err_poisson/pdq outputs are required");
    const int G = prm->G, n = h->n;
    RIP_REQUIRE(G == plan->G, "rip_l1_to_l2: params.G=%d but plan.G=%d", G, plan->G);
    RIP_REQUIRE(G >= 3 && G <= RIP_GMAX, "rip_l1_to_l2: G=%d outside 3..%d", G, RIP_GMAX);
    RIP_REQUIRE((plan->start != 0) == (prm->exclude_first != 0), "rip_l1_to_l2: plan.start and exclude_first disagree");
    RIP_REQUIRE(!prm->do_refpix || (int)(h->dark_cube.n / ((size_t)n * n)) >= G, "rip_l1_to_l2: dark cube has fewer groups than the exposure");
    RIP_REQUIRE(!h->has_bias || h->d.n_bias >= G, "rip_l1_to_l2: biascorr cube has fewer groups than the exposure");
    RIP_REQUIRE(prm->sat_backup >= 0 && prm->sat_backup < RIP_GMAX, "rip_l1_to_l2: bad SATURATION_BACKUP %d", prm->sat_backup);
    const double* dw = plan_to_device(h->device, plan, w_exact, st);
    // a look-ahead (rip_caldir_prefetch_refpix) is valid for the call that immediately follows it and names the same cube
    int pre = -1;
    for (int k = 0; k < 2; ++k) {
        if (prm->do_refpix && h->k0w[k].key && h->k0w[k].key == (const void*)d_raw) pre = k;
        h->k0w[k].key = nullptr;
    }
    if (pre >= 0) h->k0_cur = pre;
    rip_caldir::K0Work* W = &h->k0w[h->k0_cur];
    if (prm->do_refpix) {
        if (W->busy) {  // work of the side stream on this set (the awaited statistics, or an unused look-ahead)
            RIP_CUDA(cudaStreamWaitEvent(st, W->ev_k0, 0));
            W->busy = false;
        }
        if (pre < 0) run_k0(h, *W, d_raw, d_amp33, G, st);
    }
    // v2 (rip_v2_core.cuh) for the common all-f32 configuration; params.threads > 0 selects the generic v1 tile kernel
    // (threads < 0: development selector of the fused-kernel variant, -1 = v2, -2 / -3 = v3 without / with stage b in role X,
    //  -4 = v2t, -5 = v6; default: v6 where supported (G = 8, float32 ipc4d), else v2)
    const int variant = prm->threads < 0 ? -prm->threads - 1 : fused_default_variant();
    const bool k64 = h->has_ipc && h->d.ipc_dtype == RIP_F64;
    const bool use_v2 = prm->threads <= 0 && h->has_ipc && h->d.gain_dtype == RIP_F32 &&
                        h->nb == 4 && n % 8 == 0 && n >= 16 && v2_supported(G, h->P, k64) &&
                        (prm->area_dtype == RIP_F32 || prm->area_dtype == RIP_F64);
    if (!use_v2 && prm->threads <= 0 && !h->warned_generic) {  // say so once per handle: the generic kernel is ~5x slower
        h->warned_generic = true;
        fprintf(stderr, "librip_b200: this configuration (G=%d, P=%d, gain %s, ipc4d %s, n=%d) runs the generic fused kernel, about 5x "
                        "slower than the throughput kernel (float32 gain; G = 8 or 16; P <= 11; float64 ipc4d only with G = 8)\n",
                G, h->P, h->d.gain_dtype == RIP_F64 ? "float64" : "float32", !h->has_ipc ? "absent" : (k64 ? "float64" : "float32"), n);
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profile) {
        if (h->prof_used + 2 > h->prof_ev.size()) {
            cudaEvent_t a, b;
            RIP_CUDA(cudaEventCreate(&a));
            RIP_CUDA(cudaEventCreate(&b));
            h->prof_ev.push_back(a);
            h->prof_ev.push_back(b);
        }
        e0 = h->prof_ev[h->prof_used];
        e1 = h->prof_ev[h->prof_used + 1];
        h->prof_used += 2;
    }
    const float* bias_p = h->has_bias ? h->biascorr.p + (size_t)(h->d.n_bias - G) * h->na * h->na : nullptr;  // gen_cal_image.py:561-562
    if (use_v2) {
        v2_pack(h, G, st);
        v2::Args V;
        memset(&V, 0, sizeof V);
        V.n = n; V.ntile = v2::ntiles(n);
        const int cps = (variant == 4 && v6_supported(G, h->P)) ? (k64 ? 4 : 5) : (k64 || variant == 1 || variant == 2) ? ((G <= 8) ? 3 : 2) : 0;
        V.band_rows = prm->band_rows > 0 ? prm->band_rows : v2_default_band_rows(h->device, n, G, cps);
        V.do_refpix = prm->do_refpix; V.do_not_flag_first = prm->do_not_flag_first; V.exclude_first = prm->exclude_first;
        V.sat_backup = prm->sat_backup; V.area_dtype = prm->area_dtype;
        V.negzero = -0.0f;
        V.raw = d_raw; V.area = d_area;
        V.rowcorr = W->rowcorr.p; V.chan_m = W->chan_m.p; V.chan_c = W->chan_c.p;
        V.rec1 = v2_rec1_row0(h, G); V.recK = v2_recK_row0(h); V.thr = h->thr_eff.p;
        V.chan_line = W->chan_line.p;
        V.w_exact = dw;
        V.slope = o->slope; V.err_read = o->err_read; V.err_poisson = o->err_poisson; V.pdq = o->pdq;
        V.endslice = o->endslice; V.rdq = o->rdq; V.lincube = o->lin_cube;
        if (e0) RIP_CUDA(cudaEventRecord(e0, st));
        RIP_REQUIRE(((uintptr_t)d_raw & 15) == 0, "rip_l1_to_l2: the raw cube must be 16-byte aligned");
        if (k64 && variant == 4 && v6_supported(G, h->P)) launch_cal_fused_v6k64(V, G, h->P, st);
        else if (k64) launch_cal_fused_v2k64(V, G, h->P, st);
        else if (variant == 4 && v6_supported(G, h->P)) launch_cal_fused_v6(V, G, h->P, st);
        else if (variant && variant != 4) launch_cal_fused_v3(V, G, h->P, variant, st);
        else launch_cal_fused_v2(V, G, h->P, st);
        if (e1) RIP_CUDA(cudaEventRecord(e1, st));
        if (prm->do_refpix) { RIP_CUDA(cudaEventRecord(W->ev_used, st)); W->used_recorded = true; }
        return;
    }
    CalArgs A;
    memset(&A, 0, sizeof A);
    A.n = n; A.nb = h->nb; A.G = G; A.P = h->P;
    A.band_rows = prm->band_rows > 0 ? prm->band_rows : 128;
    A.do_refpix = prm->do_refpix; A.do_not_flag_first = prm->do_not_flag_first; A.exclude_first = prm->exclude_first;
    A.sat_backup = prm->sat_backup; A.area_dtype = prm->area_dtype;
    A.raw = d_raw; A.area = d_area;
    A.rowcorr = W->rowcorr.p; A.chan_m = W->chan_m.p; A.chan_c = W->chan_c.p;
    A.dark = h->dark_cube.p;
    A.bias = bias_p;
    A.coefs = h->coefs.p; A.Smin = h->Smin.p; A.Smax = h->Smax.p; A.Sref = h->Sref.p;
    A.aux = h->aux.p; A.sdq = h->sdq.p; A.thr = h->thr_eff.p;
    A.gain = h->gain.p; A.ipc = h->has_ipc ? h->ipc.p : nullptr; A.read = h->read.p;
    A.dslope = h->dslope_ipc.p; A.flat = h->flat_ipc.p; A.w_exact = dw;
    A.slope = o->slope; A.err_read = o->err_read; A.err_poisson = o->err_poisson; A.pdq = o->pdq;
    A.endslice = o->endslice; A.rdq = o->rdq; A.lincube = o->lin_cube;
    int threads = prm->threads > 0 ? prm->threads : 128;
    if (e0) RIP_CUDA(cudaEventRecord(e0, st));
    launch_cal_fused(A, h->d.gain_dtype, h->has_ipc ? h->d.ipc_dtype : RIP_F32, threads, st);
    if (e1) RIP_CUDA(cudaEventRecord(e1, st));
    if (prm->do_refpix) { RIP_CUDA(cudaEventRecord(W->ev_used, st)); W->used_recorded = true; }
}

extern "C" int rip_profile_enable(rip_caldir* h, int on) {
    RIP_API_BEGIN
    RIP_REQUIRE(h, "rip_profile_enable: null handle");
    h->profile = on != 0;
    h->prof_used = 0;
    RIP_API_END
}

extern "C" int rip_profile_fetch(rip_caldir* h, double* fused_ms_sum, int* launches) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && fused_ms_sum && launches, "rip_profile_fetch: null argument");
    use_device(h->device);
    double sum = 0.0;
    for (size_t k = 0; k + 1 < h->prof_used; k += 2) {
        RIP_CUDA(cudaEventSynchronize(h->prof_ev[k + 1]));
        float ms = 0.f;
        RIP_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[k], h->prof_ev[k + 1]));
        sum += ms;
    }
    *fused_ms_sum = sum;
    *launches = (int)(h->prof_used / 2);
    h->prof_used = 0;
    RIP_API_END
}

extern "C" int rip_l1_to_l2_dev(rip_caldir* h, const uint16_t* d_raw, const uint16_t* d_amp33, const void* d_area,
                                const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                                const rip_l2_out* d_out, void* stream) {
    RIP_API_BEGIN
    RIP_REQUIRE(h, "rip_l1_to_l2_dev: null handle");
    use_device(h->device);
    l1_to_l2_dev_impl(h, d_raw, d_amp33, d_area, prm, plan, w_exact, d_out, (cudaStream_t)stream);
    RIP_API_END
}

extern "C" int rip_caldir_prefetch_refpix(rip_caldir* h, const uint16_t* d_raw_next, const uint16_t* d_amp33_next, int G) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && d_raw_next && d_amp33_next, "rip_caldir_prefetch_refpix: null argument");
    RIP_REQUIRE(G >= 3 && G <= RIP_GMAX, "rip_caldir_prefetch_refpix: G=%d outside 3..%d", G, RIP_GMAX);
    RIP_REQUIRE((int)(h->dark_cube.n / ((size_t)h->n * h->n)) >= G, "rip_caldir_prefetch_refpix: dark cube has fewer groups than the exposure");
    use_device(h->device);
    rip_caldir::K0Work& W = h->k0w[1 - h->k0_cur];  // the set the most recent fused launch does NOT read
    // the last fused kernel that read this set (two exposures back) must have finished before it is overwritten
    if (W.used_recorded) RIP_CUDA(cudaStreamWaitEvent(h->s_k0, W.ev_used, 0));
    run_k0(h, W, d_raw_next, d_amp33_next, G, h->s_k0);
    RIP_CUDA(cudaEventRecord(W.ev_k0, h->s_k0));
    W.key = (const void*)d_raw_next;
    W.busy = true;
    RIP_API_END
}

extern "C" int rip_l1_to_l2_host(rip_caldir* h, const uint16_t* raw, const uint16_t* amp33, const void* area,
                                 const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                                 const rip_l2_out* out) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && raw && prm && plan && out, "rip_l1_to_l2_host: null argument");
    use_device(h->device);
    cudaStream_t st = h->stream;
    const int G = prm->G, n = h->n, na = h->na;
    RIP_REQUIRE(G >= 3 && G <= RIP_GMAX, "rip_l1_to_l2_host: G=%d outside 3..%d", G, RIP_GMAX);
    const size_t npl = (size_t)n * n;
    h->w_raw.upload(raw, (size_t)G * npl, st);
    if (prm->do_refpix) {
        RIP_REQUIRE(amp33, "rip_l1_to_l2_host: do_refpix needs the amp33 cube");
        h->w_amp.upload(amp33, (size_t)G * n * 128, st);
    }
    if (area) h->w_area.upload(area, npl * dtype_size(prm->area_dtype), st);
    if (h->w_slope.n < npl) { h->w_slope.alloc(npl); h->w_er.alloc(npl); h->w_ep.alloc(npl); h->w_pdq.alloc(npl); }
    rip_l2_out o{};
    o.slope = h->w_slope.p; o.err_read = h->w_er.p; o.err_poisson = h->w_ep.p; o.pdq = h->w_pdq.p;
    if (out->endslice) { if (h->w_end.n < (size_t)na * na) h->w_end.alloc((size_t)na * na); o.endslice = h->w_end.p; }
    if (out->rdq) { if (h->w_rdq.n < (size_t)G * npl) h->w_rdq.alloc((size_t)G * npl); o.rdq = h->w_rdq.p; }
    if (out->lin_cube) { if (h->w_lin.n < (size_t)G * npl) h->w_lin.alloc((size_t)G * npl); o.lin_cube = h->w_lin.p; }
    l1_to_l2_dev_impl(h, h->w_raw.p, prm->do_refpix ? h->w_amp.p : nullptr, area ? h->w_area.p : nullptr, prm, plan, w_exact, &o, st);
    h->w_slope.download(out->slope, npl, st);
    h->w_er.download(out->err_read, npl, st);
    h->w_ep.download(out->err_poisson, npl, st);
    h->w_pdq.download(out->pdq, npl, st);
    if (out->endslice) h->w_end.download(out->endslice, (size_t)na * na, st);
    if (out->rdq) h->w_rdq.download(out->rdq, (size_t)G * npl, st);
    if (out->lin_cube) h->w_lin.download(out->lin_cube, (size_t)G * npl, st);
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}

extern "C" int rip_refpix_stats_host(rip_caldir* h, const uint16_t* raw, const uint16_t* amp33, int G, double* rowcorr,
                                     double* chan_m, double* chan_c, float* gmed) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && raw && amp33, "rip_refpix_stats_host: null argument");
    RIP_REQUIRE(G >= 1 && G <= RIP_GMAX, "rip_refpix_stats_host: G=%d outside 1..%d", G, RIP_GMAX);
    use_device(h->device);
    cudaStream_t st = h->stream;
    const int n = h->n;
    h->w_raw.upload(raw, (size_t)G * n * n, st);
    h->w_amp.upload(amp33, (size_t)G * n * 128, st);
    RIP_CUDA(cudaStreamSynchronize(h->s_k0));  // a pending look-ahead may own set 0
    rip_caldir::K0Work& W0 = h->k0w[0];
    W0.key = nullptr;
    W0.busy = false;
    if (W0.used_recorded) RIP_CUDA(cudaStreamWaitEvent(st, W0.ev_used, 0));  // a fused kernel on a caller's stream may still read it
    run_k0(h, W0, h->w_raw.p, h->w_amp.p, G, st);
    if (rowcorr) W0.rowcorr.download(rowcorr, (size_t)G * n, st);
    if (chan_m) W0.chan_m.download(chan_m, (size_t)G * 32, st);
    if (chan_c) W0.chan_c.download(chan_c, (size_t)G * 32, st);
    if (gmed) W0.gmed.download(gmed, G, st);
    RIP_CUDA(cudaStreamSynchronize(st));
    RIP_API_END
}


// =========================================================================================================
// Pipelined host entry: exposures in flight on three streams (H2D | K0 + fused kernel | D2H), `depth` slots of
// device buffers.  For callers that process a stream of exposures of one SCA from (pinned) host memory: the
// PCIe copies of neighbouring exposures overlap the kernels, so the end-to-end rate is bounded by the slower PCIe
// direction instead of the sum of copies and compute.
// =========================================================================================================
struct rip_pipeline {
    rip_caldir* h = nullptr;
    int G = 0, depth = 0;
    bool want_end = false, want_rdq = false;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    struct Slot {
        DevBuf<uint16_t> raw, amp;
        DevRaw area;
        DevBuf<float> slope, er, ep;
        DevBuf<uint32_t> pdq;
        DevBuf<int8_t> end;
        DevBuf<uint8_t> rdq;
        cudaEvent_t in_done = nullptr, run_done = nullptr, out_done = nullptr;
        long ticket = -1;
    };
    std::unique_ptr<Slot[]> slots;
    long next_ticket = 0;
    // AreaFactor plane kept on the device across exposures (rip_pipeline_set_area): submit(area = NULL) then uses it
    DevRaw area_res;
    int area_res_dtype = 0;
    bool area_res_set = false;
};

extern "C" int rip_pipeline_create(rip_caldir* h, int G, int depth, int want_endslice, int want_rdq, rip_pipeline** out) {
    RIP_API_BEGIN
    RIP_REQUIRE(h && out, "rip_pipeline_create: null argument");
    RIP_REQUIRE(G >= 3 && G <= RIP_GMAX, "rip_pipeline_create: G=%d outside 3..%d", G, RIP_GMAX);
    RIP_REQUIRE(depth >= 1 && depth <= 8, "rip_pipeline_create: depth=%d outside 1..8", depth);
    use_device(h->device);
    std::unique_ptr<rip_pipeline> p(new rip_pipeline);
    p->h = h; p->G = G; p->depth = depth; p->want_end = want_endslice != 0; p->want_rdq = want_rdq != 0;
    RIP_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    RIP_CUDA(cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking));
    RIP_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    const size_t npl = (size_t)h->n * h->n, npa = (size_t)h->na * h->na;
    p->slots.reset(new rip_pipeline::Slot[depth]);
    for (int i = 0; i < depth; ++i) {
        auto& s = p->slots[i];
        s.raw.alloc((size_t)G * npl);
        s.amp.alloc((size_t)G * h->n * 128);
        s.area.alloc(npl * 8);
        s.slope.alloc(npl); s.er.alloc(npl); s.ep.alloc(npl); s.pdq.alloc(npl);
        if (p->want_end) s.end.alloc(npa);
        if (p->want_rdq) s.rdq.alloc((size_t)G * npl);
        RIP_CUDA(cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
        RIP_CUDA(cudaEventCreateWithFlags(&s.run_done, cudaEventDisableTiming));
        RIP_CUDA(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
    }
    *out = p.release();
    RIP_API_END
}

extern "C" void rip_pipeline_destroy(rip_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->h->device);
    cudaStreamSynchronize(p->s_in);
    cudaStreamSynchronize(p->s_run);
    cudaStreamSynchronize(p->s_out);
    for (int i = 0; i < p->depth; ++i) {
        auto& s = p->slots[i];
        if (s.in_done) cudaEventDestroy(s.in_done);
        if (s.run_done) cudaEventDestroy(s.run_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    cudaStreamDestroy(p->s_in);
    cudaStreamDestroy(p->s_run);
    cudaStreamDestroy(p->s_out);
    delete p;
}

extern "C" int rip_pipeline_set_area(rip_pipeline* p, const void* area, int area_dtype) {
    RIP_API_BEGIN
    RIP_REQUIRE(p, "rip_pipeline_set_area: null pipeline");
    rip_caldir* h = p->h;
    use_device(h->device);
    // exposures in flight may still read the old plane
    RIP_CUDA(cudaStreamSynchronize(p->s_run));
    if (!area) {
        p->area_res_set = false;
    } else {
        RIP_REQUIRE(area_dtype == RIP_F32 || area_dtype == RIP_F64, "rip_pipeline_set_area: area dtype must be f32 or f64");
        const size_t bytes = (size_t)h->n * h->n * dtype_size(area_dtype);
        p->area_res.upload(area, bytes, p->s_in);
        RIP_CUDA(cudaStreamSynchronize(p->s_in));
        p->area_res_dtype = area_dtype;
        p->area_res_set = true;
    }
    RIP_API_END
}

extern "C" int rip_pipeline_set_area_wcs(rip_pipeline* p, const double* wcs, int nwcs, double inv_omega, int area_dtype) {
    RIP_API_BEGIN
    RIP_REQUIRE(p && wcs, "rip_pipeline_set_area_wcs: null argument");
    rip_caldir* h = p->h;
    use_device(h->device);
    RIP_REQUIRE(area_dtype == RIP_F32 || area_dtype == RIP_F64, "rip_pipeline_set_area_wcs: area dtype must be f32 or f64");
    const size_t bytes = (size_t)h->n * h->n * dtype_size(area_dtype);
    if (p->area_res.bytes < bytes) {
        RIP_CUDA(cudaStreamSynchronize(p->s_run));
        p->area_res.alloc(bytes);
    }
    // on the compute stream: ordered after the exposures already queued (they read the old plane), before the next ones
    launch_pixel_area(wcs, nwcs, h->n, inv_omega, p->area_res.p, area_dtype, p->s_run);
    p->area_res_dtype = area_dtype;
    p->area_res_set = true;
    RIP_API_END
}

extern "C" int rip_pipeline_submit(rip_pipeline* p, const uint16_t* raw, const uint16_t* amp33, const void* area,
                                   const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                                   const rip_l2_out* out, long* ticket) {
    RIP_API_BEGIN
    RIP_REQUIRE(p && raw && prm && plan && out && ticket, "rip_pipeline_submit: null argument");
    RIP_REQUIRE(prm->G == p->G, "rip_pipeline_submit: params.G=%d but the pipeline was created for G=%d", prm->G, p->G);
    RIP_REQUIRE(out->slope && out->err_read && out->err_poisson && out->pdq, "rip_pipeline_submit: slope/err_read/err_poisson/pdq outputs are required");
    RIP_REQUIRE(!out->endslice || p->want_end, "rip_pipeline_submit: pipeline was created without endslice buffers");
    RIP_REQUIRE(!out->rdq || p->want_rdq, "rip_pipeline_submit: pipeline was created without rdq buffers");
    RIP_REQUIRE(!out->lin_cube, "rip_pipeline_submit: lin_cube is not available on the pipelined path");
    rip_caldir* h = p->h;
    use_device(h->device);
    const size_t npl = (size_t)h->n * h->n, npa = (size_t)h->na * h->na;
    const int G = p->G;
    auto& s = p->slots[p->next_ticket % p->depth];
    if (s.ticket >= 0) RIP_CUDA(cudaEventSynchronize(s.out_done));  // slot still draining to the host: wait for it
    // H2D
    RIP_CUDA(cudaMemcpyAsync(s.raw.p, raw, (size_t)G * npl * 2, cudaMemcpyHostToDevice, p->s_in));
    if (prm->do_refpix) {
        RIP_REQUIRE(amp33, "rip_pipeline_submit: do_refpix needs the amp33 cube");
        RIP_CUDA(cudaMemcpyAsync(s.amp.p, amp33, (size_t)G * h->n * 128 * 2, cudaMemcpyHostToDevice, p->s_in));
    }
    if (area) RIP_CUDA(cudaMemcpyAsync(s.area.p, area, npl * dtype_size(prm->area_dtype), cudaMemcpyHostToDevice, p->s_in));
    const void* d_area = area ? s.area.p : nullptr;
    rip_l1l2_params prm_local = *prm;
    if (!area && p->area_res_set) {  // resident plane (rip_pipeline_set_area)
        d_area = p->area_res.p;
        prm_local.area_dtype = p->area_res_dtype;
    }
    RIP_CUDA(cudaEventRecord(s.in_done, p->s_in));
    // compute (one stream: the K0 workspace of the handle is shared by all slots)
    RIP_CUDA(cudaStreamWaitEvent(p->s_run, s.in_done, 0));
    rip_l2_out o{};
    o.slope = s.slope.p; o.err_read = s.er.p; o.err_poisson = s.ep.p; o.pdq = s.pdq.p;
    if (out->endslice) o.endslice = s.end.p;
    if (out->rdq) o.rdq = s.rdq.p;
    l1_to_l2_dev_impl(h, s.raw.p, prm->do_refpix ? s.amp.p : nullptr, d_area, &prm_local, plan, w_exact, &o, p->s_run);
    RIP_CUDA(cudaEventRecord(s.run_done, p->s_run));
    // D2H
    RIP_CUDA(cudaStreamWaitEvent(p->s_out, s.run_done, 0));
    RIP_CUDA(cudaMemcpyAsync(out->slope, s.slope.p, npl * 4, cudaMemcpyDeviceToHost, p->s_out));
    RIP_CUDA(cudaMemcpyAsync(out->err_read, s.er.p, npl * 4, cudaMemcpyDeviceToHost, p->s_out));
    RIP_CUDA(cudaMemcpyAsync(out->err_poisson, s.ep.p, npl * 4, cudaMemcpyDeviceToHost, p->s_out));
    RIP_CUDA(cudaMemcpyAsync(out->pdq, s.pdq.p, npl * 4, cudaMemcpyDeviceToHost, p->s_out));
    if (out->endslice) RIP_CUDA(cudaMemcpyAsync(out->endslice, s.end.p, npa, cudaMemcpyDeviceToHost, p->s_out));
    if (out->rdq) RIP_CUDA(cudaMemcpyAsync(out->rdq, s.rdq.p, (size_t)G * npl, cudaMemcpyDeviceToHost, p->s_out));
    RIP_CUDA(cudaEventRecord(s.out_done, p->s_out));
    s.ticket = p->next_ticket;
    *ticket = p->next_ticket++;
    RIP_API_END
}

extern "C" int rip_pipeline_wait(rip_pipeline* p, long ticket) {
    RIP_API_BEGIN
    RIP_REQUIRE(p, "rip_pipeline_wait: null pipeline");
    RIP_REQUIRE(ticket >= 0 && ticket < p->next_ticket, "rip_pipeline_wait: unknown ticket %ld", ticket);
    use_device(p->h->device);
    auto& s = p->slots[ticket % p->depth];
    // a newer exposure in the same slot implies this ticket completed (submit waited for it)
    if (s.ticket == ticket) RIP_CUDA(cudaEventSynchronize(s.out_done));
    RIP_API_END
}
