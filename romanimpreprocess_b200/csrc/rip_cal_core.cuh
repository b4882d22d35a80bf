// cal_fused: the fused L1->L2 tile kernel body, one march step, host+device.
//
// A CTA owns an output tile of `band` rows x (TPB-6) columns and marches down the rows.  Thread `tid` follows one
// detector column x = c0-3+tid through four pipeline stages that work on *different* rows in the same step, so a
// single __syncthreads() per step is enough (every stage only reads ring slots written in earlier steps):
//
//   a0 (row s)   raw u16 -> ring; cumulative "raw >= saturation" bits + A/D-floor bits       [needs nothing]
//   a1 (row s-2) 3x3 saturation growth + backup -> group flags; reference-pixel + bias correction;
//                Legendre linearisation of all groups (multilin) ; D = lin*gain -> ring        [sat rows s-3..s-1]
//   b  (row s-4) IPC deconvolution pass 1:  O1 = (D + D) - K(*)D -> ring                       [D rows s-5..s-3]
//   c  (row s-6) IPC pass 2: (O1 + D) - K(*)O1, /gain; ramp fit + jump flags + DQ propagation;
//                dark, error split, flat/area; stores                                           [O1 rows s-7..s-5]
//
// Halo: the 2x3x3 IPC stencil reaches +-2 px and saturation growth +-1 px on raw data, so a0 runs on +-3 rows /
// columns, a1 on +-2, b on +-1 (the per-pixel stages are recomputed on the halo; x-halo = 6 of TPB columns).
// Reference semantics: L1_to_L2/gen_cal_image.py:503-629,697-709 and the utils modules cited in rip_math.cuh.
#pragma once
#include "rip_math.cuh"

namespace rip {

template <typename A, typename B> struct Promote { typedef double type; };
template <> struct Promote<float, float> { typedef float type; };

struct CalArgs {
    int n, nb, G, P;
    int band_rows;
    int do_refpix, do_not_flag_first, exclude_first, sat_backup, area_dtype;
    // per exposure
    const uint16_t* raw;     // [G,n,n]
    const void* area;        // [n,n] f32|f64 or null
    const double* rowcorr;   // [G,n]    K0: slope*(ref_med[i]-ctr)
    const double* chan_m;    // [G,32]   K0: per-channel line
    const double* chan_c;
    // CALDIR planes
    const float* dark;       // [>=G,n,n]
    const float* bias;       // [G,na,na] (already offset by the biascorr group offset) or null
    const float* coefs;      // [P,n,n]
    const float* Smin;
    const float* Smax;
    const float* Sref;
    const uint8_t* aux;      // bit0: lin dq has NO_LIN_CORR|REFERENCE_PIXEL; bit1: (mask|lin dq) has REFERENCE_PIXEL
    const uint32_t* sdq;     // merged static dq
    const float* thr;        // saturation threshold, +inf where unchecked
    const void* gain;        // [n,n] TG
    const void* ipc;         // [9,na,na] TK or null
    const float* read;
    const float* dslope;     // IPC-corrected dark slope
    const float* flat;       // get_flat product
    const double* w_exact;   // [nslices, RIP_GMAX]
    // outputs
    float* slope;
    float* err_read;
    float* err_poisson;
    uint32_t* pdq;
    int8_t* endslice;
    uint8_t* rdq;
    float* lincube;
};

// ring depths
constexpr int D_DEPTH = 5, O_DEPTH = 4, S_DEPTH = 4, R_DEPTH = 3, F_DEPTH = 5;

template <int GMAX, typename TIM, typename TI>
struct CalSmem {
    TIM* D;          // [D_DEPTH][GMAX][TPB]
    TI* O1;          // [O_DEPTH][GMAX][TPB]
    uint16_t* raw;   // [R_DEPTH][GMAX][TPB]
    uint32_t* sat;   // [S_DEPTH][TPB]   low16 cumulative saturation bits, high16 A/D floor bits
    uint32_t* flg;   // [F_DEPTH][TPB]   low16 final SATURATED mask, high16 A/D floor
    uint8_t* nlc;    // [F_DEPTH][TPB]   dynamic NO_LIN_CORR
    int tpb;
    RIP_HD static size_t bytes(int tpb) {
        return (size_t)tpb * (sizeof(TIM) * D_DEPTH * GMAX + sizeof(TI) * O_DEPTH * GMAX + 2 * R_DEPTH * GMAX +
                              4 * S_DEPTH + 4 * F_DEPTH + F_DEPTH) + 64;
    }
    RIP_HD void carve(unsigned char* base, int tpb_) {
        tpb = tpb_;
        size_t off = 0;
        if (sizeof(TI) >= sizeof(TIM)) {
            O1 = (TI*)(base + off); off += sizeof(TI) * O_DEPTH * GMAX * (size_t)tpb;
            D = (TIM*)(base + off); off += sizeof(TIM) * D_DEPTH * GMAX * (size_t)tpb;
        } else {
            D = (TIM*)(base + off); off += sizeof(TIM) * D_DEPTH * GMAX * (size_t)tpb;
            O1 = (TI*)(base + off); off += sizeof(TI) * O_DEPTH * GMAX * (size_t)tpb;
        }
        sat = (uint32_t*)(base + off); off += 4 * S_DEPTH * (size_t)tpb;
        flg = (uint32_t*)(base + off); off += 4 * F_DEPTH * (size_t)tpb;
        raw = (uint16_t*)(base + off); off += 2 * R_DEPTH * GMAX * (size_t)tpb;
        nlc = (uint8_t*)(base + off);
    }
    RIP_HD TIM& d(int row, int g, int t) { return D[(((row + 15) % D_DEPTH) * GMAX + g) * tpb + t]; }
    RIP_HD TI& o(int row, int g, int t) { return O1[(((row + 16) & (O_DEPTH - 1)) * GMAX + g) * tpb + t]; }
    RIP_HD uint16_t& r(int row, int g, int t) { return raw[(((row + 15) % R_DEPTH) * GMAX + g) * tpb + t]; }
    RIP_HD uint32_t& s(int row, int t) { return sat[((row + 16) & (S_DEPTH - 1)) * tpb + t]; }
    RIP_HD uint32_t& f(int row, int t) { return flg[((row + 15) % F_DEPTH) * tpb + t]; }
    RIP_HD uint8_t& nl(int row, int t) { return nlc[((row + 15) % F_DEPTH) * tpb + t]; }
};

// 9-tap source-indexed IPC convolution at active pixel (ya,xa) (utils/ipc_linearity.py:69-94, that order).
// `val(dy,dx)` reads the image at (ya-dy, xa-dx) from a ring; kq[] holds K[1+dy][1+dx] at the source pixel.
template <typename TI, typename TK>
struct IpcTaps {
    TK k[9];  // order: c, (1,0), (-1,0), (0,1), (0,-1), (1,1), (1,-1), (-1,1), (-1,-1)
    bool ok[9];
    RIP_HD void load(const TK* K, int na, int ya, int xa) {
        const long pl = (long)na * na;
        const int DY[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1};
        const int DX[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int ys = ya - DY[q], xs = xa - DX[q];
            ok[q] = (ys >= 0 && ys < na && xs >= 0 && xs < na);
            k[q] = ok[q] ? K[(long)((1 + DY[q]) * 3 + (1 + DX[q])) * pl + (long)ys * na + xs] : (TK)0;
        }
    }
};

template <int GMAX, int PMAX, typename TG, typename TK>
RIP_HD void cal_step(const CalArgs& A, const RampPlanDev& pl,
                     CalSmem<GMAX, typename Promote<float, TG>::type,
                             typename Promote<typename Promote<float, TG>::type, TK>::type>& sm,
                     const int tid, const int TPB, const int c0, const int r0, const int r1, const int s) {
    typedef typename Promote<float, TG>::type TIM;
    typedef typename Promote<TIM, TK>::type TI;
    const int n = A.n, nb = A.nb, G = A.G, na = n - 2 * nb;
    const long npl = (long)n * n;
    const int x = c0 - 3 + tid;
    const bool xin = (x >= 0 && x < n);
    const uint32_t allg = (1u << G) - 1u;
    const TG* gainp = (const TG*)A.gain;
    const TK* ipcp = (const TK*)A.ipc;

    // ---------------- stage a0 : row s ----------------
    {
        const int row = s;
        uint32_t bits = 0u;
        if (row >= 0 && row < n && xin && row >= r0 - 3 && row < r1 + 3) {
            const long p = (long)row * n + x;
            const float thr = A.thr[p];
            bool cum = false;
#pragma unroll
            for (int g = 0; g < GMAX; ++g) {
                if (g < G) {
                    const uint16_t v = A.raw[(long)g * npl + p];
                    sm.r(row, g, tid) = v;
                    if (g >= 1) {  // saturation_check skips the first resultant (gen_cal_image.py:174-180)
                        const float fv = (float)v;
                        cum = cum || (fv >= thr);
                        if (cum) bits |= 1u << g;
                        if (fv <= 0.0f) bits |= 1u << (16 + g);
                    }
                }
            }
        }
        sm.s(row, tid) = bits;
    }

    // ---------------- stage a1 : row s-2 ----------------
    {
        const int row = s - 2;
        if (row >= 0 && row < n && row >= r0 - 2 && row < r1 + 2 && tid >= 1 && tid <= TPB - 2 && xin) {
            const long p = (long)row * n + x;
            uint32_t grown = 0u;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) grown |= sm.s(row + dy, tid + dx);
            grown &= 0xffffu;
            uint32_t satm = grown;
            for (int b = 1; b <= A.sat_backup; ++b) satm |= grown >> b;
            satm &= allg & ~1u;
            const uint32_t adf = sm.s(row, tid) >> 16;
            const bool active = (row >= nb && row < n - nb && x >= nb && x < n - nb);

            float S[GMAX], phi[GMAX];
            const int ch = x >> 7;
#pragma unroll
            for (int g = 0; g < GMAX; ++g) {
                if (g < G) {
                    float v = (float)sm.r(row, g, tid);
                    if (A.do_refpix) {  // gen_cal_image.py:535-556
                        const float dk = A.dark[(long)g * npl + p];
                        v = v - dk;
                        v = (float)((double)v - A.rowcorr[(long)g * n + row]);
                        const double line = A.chan_m[g * 32 + ch] * (double)row + A.chan_c[g * 32 + ch];
                        v = (float)((double)v - line);
                        v = v + dk;
                    }
                    if (active && A.bias) v = v - A.bias[((long)g * na + (row - nb)) * na + (x - nb)];
                    S[g] = v;
                } else {
                    S[g] = 0.0f;
                }
            }
            float c[PMAX];
#pragma unroll
            for (int L = 0; L < PMAX; ++L) c[L] = (L < A.P) ? A.coefs[(long)L * npl + p] : 0.0f;
            const uint8_t aux = A.aux[p];
            uint32_t dq = (aux & 1u) ? DQ_REFERENCE_PIXEL : 0u;
            multilin_pixel<GMAX, PMAX>(S, G, c, A.P, A.Smin[p], A.Smax[p], A.Sref[p], dq, ~satm, A.do_not_flag_first != 0, phi);
            const TIM g_ = (TIM)gainp[p];
#pragma unroll
            for (int g = 0; g < GMAX; ++g)
                if (g < G) sm.d(row, g, tid) = (active && ipcp) ? (TIM)phi[g] * g_ : (TIM)phi[g];
            sm.f(row, tid) = satm | (adf << 16);
            sm.nl(row, tid) = (dq & DQ_NO_LIN_CORR) ? 1 : 0;
        }
    }

    // ---------------- stage b : row s-4 (IPC pass 1, active pixels only) ----------------
    if (ipcp) {
        const int row = s - 4;
        if (row >= nb && row < n - nb && row >= r0 - 1 && row < r1 + 1 && tid >= 2 && tid <= TPB - 3 && x >= nb && x < n - nb) {
            IpcTaps<TI, TK> T;
            T.load(ipcp, na, row - nb, x - nb);
#pragma unroll
            for (int g = 0; g < GMAX; ++g) {
                if (g < G) {
                    const TIM d0 = sm.d(row, g, tid);
                    TI acc = (TI)d0 * (TI)T.k[0];
                    if (T.ok[1]) acc = acc + (TI)sm.d(row - 1, g, tid) * (TI)T.k[1];
                    if (T.ok[2]) acc = acc + (TI)sm.d(row + 1, g, tid) * (TI)T.k[2];
                    if (T.ok[3]) acc = acc + (TI)sm.d(row, g, tid - 1) * (TI)T.k[3];
                    if (T.ok[4]) acc = acc + (TI)sm.d(row, g, tid + 1) * (TI)T.k[4];
                    if (T.ok[5]) acc = acc + (TI)sm.d(row - 1, g, tid - 1) * (TI)T.k[5];
                    if (T.ok[6]) acc = acc + (TI)sm.d(row - 1, g, tid + 1) * (TI)T.k[6];
                    if (T.ok[7]) acc = acc + (TI)sm.d(row + 1, g, tid - 1) * (TI)T.k[7];
                    if (T.ok[8]) acc = acc + (TI)sm.d(row + 1, g, tid + 1) * (TI)T.k[8];
                    sm.o(row, g, tid) = (TI)(TIM)(d0 + d0) - acc;  // output + image2 - ipc_fwd(output)
                }
            }
        }
    }

    // ---------------- stage c : row s-6 (IPC pass 2, ramp fit, L2 epilogue) ----------------
    {
        const int row = s - 6;
        if (row >= r0 && row < r1 && row < n && tid >= 3 && tid <= TPB - 4 && xin) {
            const long p = (long)row * n + x;
            const bool active = (row >= nb && row < n - nb && x >= nb && x < n - nb);
            const TG gval = gainp[p];
            float d[GMAX];
            if (active && ipcp) {
                IpcTaps<TI, TK> T;
                T.load(ipcp, na, row - nb, x - nb);
#pragma unroll
                for (int g = 0; g < GMAX; ++g) {
                    if (g < G) {
                        const TI o0 = sm.o(row, g, tid);
                        TI acc = o0 * (TI)T.k[0];
                        if (T.ok[1]) acc = acc + sm.o(row - 1, g, tid) * (TI)T.k[1];
                        if (T.ok[2]) acc = acc + sm.o(row + 1, g, tid) * (TI)T.k[2];
                        if (T.ok[3]) acc = acc + sm.o(row, g, tid - 1) * (TI)T.k[3];
                        if (T.ok[4]) acc = acc + sm.o(row, g, tid + 1) * (TI)T.k[4];
                        if (T.ok[5]) acc = acc + sm.o(row - 1, g, tid - 1) * (TI)T.k[5];
                        if (T.ok[6]) acc = acc + sm.o(row - 1, g, tid + 1) * (TI)T.k[6];
                        if (T.ok[7]) acc = acc + sm.o(row + 1, g, tid - 1) * (TI)T.k[7];
                        if (T.ok[8]) acc = acc + sm.o(row + 1, g, tid + 1) * (TI)T.k[8];
                        const TI o2 = (o0 + (TI)sm.d(row, g, tid)) - acc;
                        d[g] = (float)(o2 / (TI)gval);
                    } else {
                        d[g] = 0.0f;
                    }
                }
            } else {
#pragma unroll
                for (int g = 0; g < GMAX; ++g) d[g] = (g < G) ? (float)sm.d(row, g, tid) : 0.0f;
            }
            if (A.lincube) {
#pragma unroll
                for (int g = 0; g < GMAX; ++g)
                    if (g < G) A.lincube[(long)g * npl + p] = d[g];
            }
            const uint32_t fl = sm.f(row, tid);
            GroupFlags gf;
            gf.sat = fl & 0xffffu;
            gf.adf = fl >> 16;
            gf.dnu = gf.adf | (A.exclude_first ? 1u : 0u);
            gf.jump = 0u;
            gf.other_unsat = 0u;
            const uint8_t aux = A.aux[p];
            uint32_t pd = (aux & 2u) ? DQ_REFERENCE_PIXEL : 0u;
            FitResult r = ramp_fit_pixel<GMAX, TG, true>(d, gf, pd, gval, A.read[p], active, pl, A.w_exact);
            const uint32_t pdq = A.sdq[p] | (sm.nl(row, tid) ? DQ_NO_LIN_CORR : 0u) | (pd & ~DQ_REFERENCE_PIXEL);
            float fa = A.flat[p];
            if (A.area) {
                if (A.area_dtype == RIP_F64) fa = (float)((double)fa / ((const double*)A.area)[p]);
                else fa = fa / ((const float*)A.area)[p];
            }
            l2_epilogue(r, active, A.dslope[p], fa);
            A.slope[p] = r.slope;
            A.err_read[p] = r.err_read;
            A.err_poisson[p] = r.err_poisson;
            A.pdq[p] = pdq;
            if (A.endslice && active) {
                int es = -1;
                for (int iend = 1; iend < G; ++iend)
                    if (((gf.sat >> iend) & 1u) && !((gf.sat >> (iend - 1)) & 1u)) es = iend - 1;
                A.endslice[(long)(row - nb) * na + (x - nb)] = (int8_t)es;
            }
            if (A.rdq) {
#pragma unroll
                for (int g = 0; g < GMAX; ++g) {
                    if (g < G) {
                        uint32_t b = 0u;
                        if ((gf.dnu >> g) & 1u) b |= DQ_DO_NOT_USE;
                        if ((gf.sat >> g) & 1u) b |= DQ_SATURATED;
                        if ((gf.jump >> g) & 1u) b |= DQ_JUMP_DET;
                        if ((gf.adf >> g) & 1u) b |= DQ_AD_FLOOR;
                        A.rdq[(long)g * npl + p] = (uint8_t)b;
                    }
                }
            }
        }
    }
}

}  // namespace rip
