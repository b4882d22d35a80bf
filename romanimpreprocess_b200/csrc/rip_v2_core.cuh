// cal_fused v2: the throughput form of the fused L1->L2 tile kernel for the common configuration
// (all-f32 calibration planes, ipc4d present, nb = 4, even G in 4..16, frame side a multiple of 8).
// Everything else runs the generic v1 kernel (rip_cal_core.cuh).  Same arithmetic, same roundings, same flags as
// v1 -- only the organisation differs (profiles/r01: v1 was instruction-issue bound, 40 % of its instructions were
// address arithmetic and 10 % branches):
//
//   * exact (G, P) template parameters: no per-group / per-order predicates;
//   * calibration data repacked once per CALDIR into per-(row, column-tile) records of float4 words
//     (rec1: dark[G] bias[G] Smin Smax Sref gain aux coefs[P];  recK: 9 gathered IPC taps, gain, read, dark slope,
//     flat, static dq) -> every input arrives as a fully coalesced 128-bit load with immediate offsets;
//   * the records of the next march step are loaded into registers behind a SINGLE scoreboard wait per step (ptxas
//     gives all global loads of the loop one scoreboard, so any first use waits for every load in flight: see Regs);
//     thresholds and reference-pixel corrections ride the cp.async group of the raw row instead of registers;
//   * shared-memory rings hold float4 (4 groups of one pixel) -> conflict-free LDS.128 / STS.128;
//   * group pairs are processed with the packed FP32 instructions of sm_100 (FMUL2 / FADD2 / FFMA2: two IEEE
//     single-precision results per issue slot; each lane is rounded exactly like the scalar op, so results are
//     bit-identical to v1);
//   * the 8 divisions by the same denominator (z = .../(Smax-Smin), d = o2/gain) share one refined reciprocal and
//     use the FMA-residual correction that yields the correctly rounded quotient (== IEEE division);
//   * gathered IPC taps are zero where the source pixel is outside the active area and D is stored as 0 for
//     non-active pixels, so the 3x3 stencils are branch-free;
//   * jump significance is tested as delta |delta| > thr^2 var on packed slice pairs (no sqrt / division; FMAs are
//     fine there because it only classifies); anything within the relative band of the threshold is re-evaluated
//     exactly in the reference's op order (same fallback as v1).
//
// Written __host__ __device__ like v1 so that tests/hostcheck can walk the identical source on the CPU.
#pragma once
#include "rip_math.cuh"

#include <string.h>

namespace rip {
namespace v2 {

constexpr int TW = 128;   // columns per tile (= compute threads per CTA)
constexpr int TS = 120;   // tile stride: outputs are tile columns 4..123 (tile 0 also 0..3)
constexpr int RW = TW + 2;  // ring width: 1 pad column on each side
constexpr int KQ = 4;     // float4 words per pixel in recK
constexpr int KQ64 = 6;   // ... when the taps are float64 (K64)
RIP_HD constexpr int kq_of(bool k64) { return k64 ? KQ64 : KQ; }
constexpr int RING = 5;   // depth of the rings addressed with the shared modulo-5 slot (D, raw, thresholds, flags, corrections)
constexpr int O_DEPTH = 4, S_DEPTH = 4;
constexpr int PADR = 10;  // zero rows above and below the packed records: the loaders never clamp (rows -9 .. n+4 are touched)

RIP_HD constexpr int nq1(int G, int P) { return (2 * G + 5 + P + 3) / 4; }
inline int ntiles(int n) { return (n - 4 + TS - 1) / TS; }

struct alignas(8) f2 {
    float x, y;
};
struct alignas(16) f4 {
    float x, y, z, w;
};

// ---- packed arithmetic: two independent IEEE-rounded single-precision operations -------------------------
// NB (measured, CUDA 12.9 ptxas for sm_100a): ptxas contracts `mul.rn.f32x2` feeding `add.rn.f32x2` into one FFMA2
// even with --fmad=false, which would change the rounding.  It does NOT contract a packed multiply feeding SCALAR
// `add.rn.f32` / `sub.rn.f32`.  Hence: multiplies are packed (mul2); additions/subtractions that may consume a
// product are scalar pairs (add2 / sub2); packed add/sub (add2p / sub2p) only where no operand is a product.
// A product that must feed a packed add is computed as fma(a, b, nz) with nz = -0.0f passed at RUN TIME (Args::negzero;
// x + (-0) == x for every x, so this is the IEEE product, and ptxas can neither fold the unknown addend nor contract
// an FMA into the following add): mul2x.  With that, products and sums of the long multiply-add chains (Legendre
// recursion, 3x3 stencils) are all packed.
// `make check-sass` (csrc/Makefile) verifies that the FFMA2 count in SASS equals the fma.rn.f32x2 count in PTX.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return f2{r.x, r.y}; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return f2{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)}; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return f2{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)}; }
__device__ __forceinline__ f2 add2p(f2 a, f2 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return f2{r.x, r.y}; }
__device__ __forceinline__ f2 sub2p(f2 a, f2 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y)); return f2{r.x, r.y}; }
__device__ __forceinline__ f2 mul2x(f2 a, f2 b, float nz) { float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(nz, nz)); return f2{r.x, r.y}; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y)); return f2{r.x, r.y}; }
__device__ __forceinline__ float fma1(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float rcp_approx(float d) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d)); return r; }
#define RIP_COLD __noinline__
#else
inline f2 mul2(f2 a, f2 b) { return f2{a.x * b.x, a.y * b.y}; }
inline f2 add2(f2 a, f2 b) { return f2{a.x + b.x, a.y + b.y}; }
inline f2 sub2(f2 a, f2 b) { return f2{a.x - b.x, a.y - b.y}; }
inline f2 add2p(f2 a, f2 b) { return add2(a, b); }
inline f2 sub2p(f2 a, f2 b) { return sub2(a, b); }
inline f2 mul2x(f2 a, f2 b, float) { return f2{a.x * b.x, a.y * b.y}; }
inline f2 fma2(f2 a, f2 b, f2 c) { return f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float fma1(float a, float b, float c) { return fmaf(a, b, c); }
inline float rcp_approx(float d) { return 1.0f / d; }
#define RIP_COLD
#endif
RIP_HD f2 bc(float a) { return f2{a, a}; }

// np.maximum / np.minimum against a non-NaN, non-negative-zero bound (NaN in x propagates): one instruction on the device
RIP_HD float max_nan(float x, float bound) {
#if defined(__CUDA_ARCH__)
    float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(bound)); return r;
#else
    return np_max<float>(x, bound);
#endif
}
RIP_HD float min_nan(float x, float bound) {
#if defined(__CUDA_ARCH__)
    float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(bound)); return r;
#else
    return np_min<float>(x, bound);
#endif
}
RIP_HD float clip_nan(float x, float lo, float hi) { return min_nan(max_nan(x, lo), hi); }

// Correctly rounded x/d for several numerators sharing one denominator (Newton-refined reciprocal + two
// FMA-residual corrections: the classic IEEE division sequence).  `ok` = denominator in the range where no
// intermediate can over/underflow for |x| < 2^40; otherwise the caller uses true division.
struct SharedDiv {
    float d, r;
    bool ok;
    RIP_HD void init(float den) {
        d = den;
        const float ad = den < 0 ? -den : den;
        ok = (ad > 1.0e-18f) && (ad < 1.0e18f);
        float r0 = rcp_approx(den);
        float e = fma1(-den, r0, 1.0f);
        r = fma1(r0, e, r0);
    }
    RIP_HD f2 div2(f2 x) const {
        const f2 nd = bc(-d), rr = bc(r);
        f2 q = mul2(x, rr);
        f2 rem = fma2(nd, q, x);
        q = fma2(rem, rr, q);
        rem = fma2(nd, q, x);
        q = fma2(rem, rr, q);
        return q;
    }
};

// two u16 -> f32 exactly with one packed subtraction: (0x4B000000 | v) is 2^23 + v
RIP_HD f2 u16_pair_to_f32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return sub2p(f2{__uint_as_float(0x4B000000u | a), __uint_as_float(0x4B000000u | b)}, bc(8388608.0f));
#else
    return f2{(float)a, (float)b};
#endif
}
// u16 -> f32 exactly, on the full-rate pipes: (0x4B000000 | v) is 2^23 + v
RIP_HD float u16_to_f32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(0x4B000000u | v) - 8388608.0f;
#else
    return (float)v;
#endif
}

struct Args {
    int n, ntile, band_rows;
    int do_refpix, do_not_flag_first, exclude_first, sat_backup, area_dtype;
    float negzero;           // -0.0f, deliberately a run-time value (see mul2x)
    int pad_;                // 0, deliberately a run-time value (see step: load ordering)
    const uint16_t* raw;     // [G,n,n]
    const void* area;        // [n,n] f32|f64 or null
    const double* rowcorr;   // [G,n]
    const double* chan_m;    // [G,32]  (K0 products; the kernel reads the tabulated lines below)
    const double* chan_c;
    const f4* rec1;          // [n + 2 PADR][ntile][NQ1][TW], pointing at row 0
    const f4* recK;          // [n + 2 PADR][ntile][KQ][TW],  pointing at row 0
    const float* thr;        // [n,n]
    const double* w_exact;
    float* slope;
    float* err_read;
    float* err_poisson;
    uint32_t* pdq;
    int8_t* endslice;
    uint8_t* rdq;
    float* lincube;
    const double* chan_line; // [G,32,n]  chan_m * row + chan_c (f64, unfused), tabulated by K0
};

// Shared memory: two rings of per-row records, so that one byte offset per ring slot addresses everything of a row.
//   ring5 (depth RING, slot = row mod 5): D f4[G/4][RW] | raw u16[G][TW] | thr f32[TW] | flg u32[TW] | rc f64[G] |
//                                         ln f64[2][G] | nlc u8[TW]
//   ring4 (depth 4, slot = row & 3):      O1 f4[G/4][RW] | sat u32[RW]
// raw / thr / rc / ln are filled by cp.async two steps ahead (rows s-2 .. s+2 live); flg (satm | adf<<16) and nlc
// (bit0 dynamic NO_LIN_CORR, bit2 reference pixel) are thread-private delay lines a1 -> c.
struct alignas(16) d2 {
    double x, y;
};
// K64: ipc4d is float64 (the dtype the DUMMY CALDIR builder writes, runs/summer2025run/make_gain_file.py:138,194): the
// reference then runs both IPC passes in float64, so the O1 ring holds doubles -- G/2 planes of
// (two groups of one pixel) = conflict-free LDS.128 -- and the taps are doubles; D stays float32 (data * gain is f32).
template <int G, bool K64 = false>
struct Smem {
    static constexpr int H = G / 4;
    static constexpr int DEPTH = RING;  // slots of the row-record ring
    static constexpr int C_LAG = 6;     // stage c runs on row s - C_LAG
    static constexpr int OFF_D = 0;
    static constexpr int OFF_RAW = OFF_D + 16 * H * RW;
    static constexpr int OFF_THR = OFF_RAW + 2 * G * TW;
    static constexpr int OFF_FLG = OFF_THR + 4 * TW;
    static constexpr int OFF_RC = OFF_FLG + 4 * TW;
    static constexpr int OFF_LN = OFF_RC + 8 * G;
    static constexpr int OFF_NLC = OFF_LN + 16 * G;
    static constexpr int ROW5 = (OFF_NLC + TW + 15) / 16 * 16;
    static constexpr int OFF_O1 = 0;
    static constexpr int OFF_SAT = (K64 ? 32 : 16) * H * RW;
    static constexpr int ROW4 = (OFF_SAT + 4 * RW + 15) / 16 * 16;
    unsigned char* r5;
    unsigned char* r4;
    RIP_HD static size_t bytes() { return (size_t)RING * ROW5 + (size_t)O_DEPTH * ROW4 + 64; }
    RIP_HD void carve(unsigned char* base) {
        r5 = base;
        r4 = base + (size_t)RING * ROW5;
    }
    // o5 = byte offset of a ring5 slot, o4 = byte offset of a ring4 slot
    RIP_HD f4* D(unsigned o5) const { return (f4*)(r5 + o5 + OFF_D); }
    RIP_HD uint16_t* raw(unsigned o5) const { return (uint16_t*)(r5 + o5 + OFF_RAW); }
    RIP_HD float* thr(unsigned o5) const { return (float*)(r5 + o5 + OFF_THR); }
    RIP_HD uint32_t* flg(unsigned o5) const { return (uint32_t*)(r5 + o5 + OFF_FLG); }
    RIP_HD double* rc(unsigned o5) const { return (double*)(r5 + o5 + OFF_RC); }
    RIP_HD double* ln(unsigned o5) const { return (double*)(r5 + o5 + OFF_LN); }
    RIP_HD uint8_t* nlc(unsigned o5) const { return (uint8_t*)(r5 + o5 + OFF_NLC); }
    RIP_HD f4* O1(int row) const { return (f4*)(r4 + (unsigned)(row & (O_DEPTH - 1)) * ROW4 + OFF_O1); }
    RIP_HD d2* O1d(int row) const { return (d2*)(r4 + (unsigned)(row & (O_DEPTH - 1)) * ROW4 + OFF_O1); }  // [G/2][RW]
    RIP_HD uint32_t* sat(int row) const { return (uint32_t*)(r4 + (unsigned)(row & (S_DEPTH - 1)) * ROW4 + OFF_SAT); }
};

// Registers a thread carries from one march step to the next: the calibration records of each stage, loaded one step
// ahead.  ptxas tracks every global load of the loop with ONE scoreboard (measured: profiles/r02), so the first use of
// any of them waits for ALL loads in flight; the schedule therefore consumes everything at the top of the step (the
// `kb = kbn` copy), when the youngest load is two stages old, and issues nothing before that point.
template <int G, int P>
struct Regs {
    static constexpr int NQ1 = nq1(G, P);
    f4 r1[NQ1];        // stage a1, row s-2
    f4 kb[2];          // stage b,  row s-4 (taps 0..7), valid from the top of the step
    float kb8;         //                   (tap 8)
    f4 kbn[2];         // the same for the NEXT step, in flight during this one
    float kbn8;
    f4 kc[KQ];         // stage c,  row s-6
    // K64 (float64 taps): recK has KQ64 words per pixel: 9 taps as doubles (18 words), gain read dslope flat sdq pad
    double kbd[9], kbnd[9];  // stage b taps (this step / next step)
    f4 kc64[6];              // stage c record
    float area32;
    double area64;
    unsigned orow;     // s * n (pixel offset of detector row s): CTA-uniform, advanced by n per step
};

RIP_HD int mod_pos(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }
// lo <= v < hi in two instructions (lo <= hi)
RIP_HD bool in_range(int v, int lo, int hi) { return (unsigned)(v - lo) < (unsigned)(hi - lo); }
RIP_HD int imax(int a, int b) { return a > b ? a : b; }
RIP_HD int imin(int a, int b) { return a < b ? a : b; }
// a record row the loaders may touch instead of `row`: rows outside [lo, hi) are loaded (unconditional loaders, see
// below) but never used, so they are folded onto the nearest used row -- an L2 hit instead of HBM traffic
RIP_HD int fold_row(int row, int lo, int hi) { return imin(imax(row, lo), hi - 1); }
// byte offset of ring5 slot (f + k) mod 5 given o = f * ROW5: one add + one unsigned min (x - 5 ROW5 wraps above x when x < 5 ROW5)
RIP_HD unsigned wrap5(unsigned x, unsigned ring_bytes) { const unsigned y = x - ring_bytes; return x < y ? x : y; }
// byte offset of the ring5 slot of row s+DK, given o5[k] = offset of row s+k (DK is a compile-time constant)
#define RIP_O5(DK) o5[(((DK) % 5) + 5) % 5]
// the same inside the stage functions, for the ring depth of the shared-memory layout they are instantiated with
#define RIP_OS(DK) o5[(((DK) % SM::DEPTH) + SM::DEPTH) % SM::DEPTH]

// ---- asynchronous global -> shared copies (LDGSTS); the host build copies at once ---------------------------
template <int BYTES>
RIP_HD void cp_async(void* smem_dst, const void* gmem_src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
#else
    memcpy(smem_dst, gmem_src, BYTES);
#endif
}
RIP_HD void cp_async_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
RIP_HD void cp_async_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// L2 prefetch of one contiguous block by the TMA engine (cp.async.bulk.prefetch.L2 = UBLKPF: one instruction, no
// registers, no completion tracking).  The packed records of a (row, tile) are contiguous 16 KB (rec1) / 8 KB (recK)
// blocks: one elected thread asks for the block about one march step before the loaders read it, so the loads that
// follow hit L2 (~300 cycles) instead of waiting for HBM (measured before: 7 % of the warp samples on the step's one
// scoreboard wait).  addr and bytes are multiples of 16.  The host build does nothing.
RIP_HD void l2_prefetch_block(const void* addr, unsigned bytes) {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
#else
    (void)addr; (void)bytes;
#endif
}
// records the loaders of step s+1 will read: rec1 of row s (load_a1 at the end of step s+1), recK of row s-2 (load_bn at
// the top of step s+1; load_c re-reads that row two steps later, from L2).  KQW = float4 words per pixel of recK.
template <int G, int P, int KQW>
RIP_HD void prefetch_records(const Args& A, int s, int tile, int tid, int r0, int r1) {
    if (tid == 0) {
        constexpr int NQ1 = nq1(G, P);
        if (in_range(s, r0 - 2, imin(r1 + 2, A.n)))
            l2_prefetch_block(A.rec1 + ((long)s * A.ntile + tile) * (NQ1 * TW), (unsigned)(NQ1 * TW * 16));
        if (in_range(s - 2, r0 - 1, imin(r1 + 1, A.n)))
            l2_prefetch_block(A.recK + ((long)(s - 2) * A.ntile + tile) * (KQW * TW), (unsigned)(KQW * TW * 16));
    }
}

// Everything stage a0 / a1 need of detector row `row` that is not a packed record -> ring slot `slot`:
//   raw resultants  G x 128 columns x u16 = G x 16 chunks of 16 bytes, one (G = 8) or two (G = 16) per thread;
//   the saturation threshold of the thread's own column (4 bytes);
//   row correction [G] and the two channel lines [2][G] of the row (threads 0 .. 3G-1, 8 bytes each).
// Every thread commits one group per step (possibly empty) so that wait_group counts steps.
template <int G, int P, class SM>
RIP_HD void row_async(const Args& A, SM& sm, const Regs<G, P>& R, int row, int row_off, unsigned slot_o5, int tile, int tid,
                      int lo, int hi) {  // row_off: rows relative to row s (R.orow)
    if (in_range(row, imax(lo, 0), imin(hi, A.n))) {
        const unsigned npl = (unsigned)A.n * (unsigned)A.n;
        const unsigned obase = R.orow + (unsigned)(row_off * A.n) + (unsigned)(tile * TS);  // uniform
        const int c = tid & 15, g0 = tid >> 4;
        if (tile * TS + 8 * c + 8 <= A.n) {
#pragma unroll
            for (int k = 0; k < G / 8; ++k)
                cp_async<16>(sm.raw(slot_o5) + (g0 + 8 * k) * TW + 8 * c,
                             A.raw + (obase + (unsigned)(g0 + 8 * k) * npl + (unsigned)(8 * c)));
        }
        const int x = tile * TS + tid;
        cp_async<4>(sm.thr(slot_o5) + tid, A.thr + (obase + (unsigned)(x < A.n ? tid : -tile * TS)));
        if (A.do_refpix && tid < 3 * G) {
            const int g = tid % G, which = tid / G;
            if (which == 0) {
                cp_async<8>(sm.rc(slot_o5) + g, A.rowcorr + ((unsigned)(g * A.n) + (unsigned)row));
            } else {
                int ch = ((tile * TS) >> 7) + (which - 1);
                if (ch > 31) ch = 31;
                cp_async<8>(sm.ln(slot_o5) + (which - 1) * G + g, A.chan_line + ((unsigned)((g * 32 + ch) * A.n) + (unsigned)row));
            }
        }
    }
    cp_async_commit();
}

// ---- record loads (registers).  Unconditional (measured: predicating them on "the stage has a row next step" makes
// ptxas spill 176 bytes and scatter scoreboard waits through the step): the records are padded with PADR zero rows on both sides, so rows
// outside the frame are loaded but never used (a conditional load would keep the old register contents live
// around the whole loop).
// 128-bit load of a record word.  Every word is read once per launch by exactly one thread, so it is loaded without an
// L1 allocation (SASS LDG.E.NA.128.CONSTANT): with the shared-memory carve-out at its maximum only ~28 KB of L1 remain
// per SM, which the 24 KB of records per CTA and step would otherwise sweep -- evicting the 32-byte stack frames and
// the uniform float64 tables of the cold paths.  Measured: 1.329 -> 1.297 ms (profiles/r02/ab_rec_no_l1.log).
// (G = 16: the v2 organisation with its smaller carve-out measured 1.7 % slower with it -- plain loads there.)
template <int G>
RIP_HD f4 ld_rec(const f4* p) {
#if defined(__CUDA_ARCH__)
    if (G > 8) return *p;
    f4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}
RIP_HD float ld_rec1(const float* p) {
#if defined(__CUDA_ARCH__)
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
// the L2 planes are written once and never read back by the kernel: streaming stores (evict-first)
RIP_HD void st_out(float* p, float v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}
RIP_HD void st_out(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}
template <int G, int P>
RIP_HD void load_a1(const Args& A, Regs<G, P>& R, int row, int tile, int tid) {
    const f4* p = A.rec1 + ((long)row * A.ntile + tile) * (Regs<G, P>::NQ1 * TW);
#pragma unroll
    for (int q = 0; q < Regs<G, P>::NQ1; ++q) R.r1[q] = ld_rec<G>(p + q * TW + tid);
}
template <int G, int P>
RIP_HD void load_bn(const Args& A, Regs<G, P>& R, int row, int tile, int tid, unsigned dep) {
    // dep: a run-time zero derived from a register of the loads in flight (see step): the address depends on it, so
    // these loads cannot be issued before the step's single scoreboard wait
    const f4* p = A.recK + ((long)row * A.ntile + tile) * (KQ * TW) + dep;
    R.kbn[0] = ld_rec<G>(p + tid);
    R.kbn[1] = ld_rec<G>(p + TW + tid);
    R.kbn8 = ((const float*)(p + 2 * TW))[4 * tid];  // .x of the third word
}
RIP_HD double f2_as_double(float lo, float hi) {
    double d;
    float w[2] = {lo, hi};
    memcpy(&d, w, 8);
    return d;
}
// K64 forms of the two recK loaders: taps of stage b as doubles (words 0..17), whole record of stage c (6 words of 4)
template <int G, int P>
RIP_HD void load_bn64(const Args& A, Regs<G, P>& R, int row, int tile, int tid, unsigned dep) {
    const f4* p = A.recK + ((long)row * A.ntile + tile) * (KQ64 * TW) + dep;
    f4 w[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) w[q] = ld_rec<G>(p + q * TW + tid);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        R.kbnd[2 * q] = f2_as_double(w[q].x, w[q].y);
        R.kbnd[2 * q + 1] = f2_as_double(w[q].z, w[q].w);
    }
    R.kbnd[8] = f2_as_double(w[4].x, w[4].y);
}
template <int G, int P>
RIP_HD void load_c64(const Args& A, Regs<G, P>& R, int row, int tile, int tid, int x, bool xin) {
    const f4* p = A.recK + ((long)row * A.ntile + tile) * (KQ64 * TW);
#pragma unroll
    for (int q = 0; q < KQ64; ++q) R.kc64[q] = ld_rec<G>(p + q * TW + tid);
    if (A.area) {
        const int rr = row < 0 ? 0 : (row >= A.n ? A.n - 1 : row);
        const unsigned o = (unsigned)rr * (unsigned)A.n + (unsigned)(xin ? x : 0);
        if (A.area_dtype == RIP_F64) R.area64 = ((const double*)A.area)[o];
        else R.area32 = ((const float*)A.area)[o];
    }
}

template <int G, int P>
RIP_HD void load_c(const Args& A, Regs<G, P>& R, int row, int tile, int tid, int x, bool xin) {
    const f4* p = A.recK + ((long)row * A.ntile + tile) * (KQ * TW);
#pragma unroll
    for (int q = 0; q < KQ; ++q) R.kc[q] = ld_rec<G>(p + q * TW + tid);
    if (A.area) {
        const int rr = row < 0 ? 0 : (row >= A.n ? A.n - 1 : row);
        const unsigned o = (unsigned)rr * (unsigned)A.n + (unsigned)(xin ? x : 0);
        if (A.area_dtype == RIP_F64) R.area64 = ((const double*)A.area)[o];
        else R.area32 = ((const float*)A.area)[o];
    }
}

RIP_HD float f4_get(const f4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
RIP_HD uint32_t f_as_u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RIP_HD float abs_f(float f) {
#if defined(__CUDA_ARCH__)
    return fabsf(f);
#else
    return f < 0 ? -f : (f == 0 ? 0.0f : f);
#endif
}

// word w of the rec1 record held in registers (w is a compile-time constant after unrolling)
template <int NQ>
RIP_HD float r1w(const f4 (&r)[NQ], int w) { return f4_get(r[w >> 2], w & 3); }

// 9-tap stencil over a float4 ring for one half (4 groups) -> two packed pairs.  Tap order = the reference's
// accumulation order (utils/ipc_linearity.py:69-94): c, (1,0), (-1,0), (0,1), (0,-1), (1,1), (1,-1), (-1,1), (-1,-1);
// tap (dy,dx) reads the image at (row-dy, col-dx).
RIP_HD void stencil9(const f4* ring_m, const f4* ring_0, const f4* ring_p, int col, const float (&k)[9], float nz, f2& lo, f2& hi) {
    // ring_m = row-1 (dy=+1), ring_0 = row, ring_p = row+1 (dy=-1); col already includes the pad offset
    const f4 c = ring_0[col];
    lo = mul2x(f2{c.x, c.y}, bc(k[0]), nz);
    hi = mul2x(f2{c.z, c.w}, bc(k[0]), nz);
#define RIP_TAP(V, Q)                                     \
    {                                                     \
        const f4 t = (V);                                 \
        lo = add2p(lo, mul2x(f2{t.x, t.y}, bc(k[Q]), nz)); \
        hi = add2p(hi, mul2x(f2{t.z, t.w}, bc(k[Q]), nz)); \
    }
    RIP_TAP(ring_m[col], 1)
    RIP_TAP(ring_p[col], 2)
    RIP_TAP(ring_0[col - 1], 3)
    RIP_TAP(ring_0[col + 1], 4)
    RIP_TAP(ring_m[col - 1], 5)
    RIP_TAP(ring_m[col + 1], 6)
    RIP_TAP(ring_p[col - 1], 7)
    RIP_TAP(ring_p[col + 1], 8)
#undef RIP_TAP
}

// all slices of variant v in the reference's exact op order (fitting.py:225-251); cold
template <int G>
struct Ramp {
    float v[G];
};
template <int G>
RIP_HD_COLD uint32_t jump_exact(const Ramp<G> rd, int v, float slope, float dvardt, float sig2read, const RampPlanDev& pl,
                                const double* w_all) {
    const float (&d)[G] = rd.v;
    const int ngrp = pl.var_ngrp[v], start = pl.start;
    const double thr_exact = jump_threshold(slope, pl);
    int s = pl.var_slice_off[v];
    uint32_t mask = 0u;
    for (int i = start; i < ngrp - 1; ++i) {
        const int dimax = (i == ngrp - 2 || ngrp - 1 - start == 2) ? 1 : 2;
        for (int di = 1; di <= dimax; ++di) {
            const RampSlice& sl = pl.slices[s];
            float dhi = d[0], dlo = d[0];
#pragma unroll
            for (int t = 0; t < G; ++t) {  // register-array selects instead of dynamic indexing
                if (t == i + di) dhi = d[t];
                if (t == i) dlo = d[t];
            }
            const float sme = smap_exact<float>((dhi - dlo) / sl.dt - slope, ngrp, dvardt, sig2read, pl, w_all + (long)s * RIP_GMAX);
            if ((double)sme > thr_exact) mask |= 1u << i;
            ++s;
        }
    }
    return mask;
}

// approximate jump threshold^2 bounds for a slope (fast classification; the band absorbs the ~1e-6 of __logf)
struct ThrBand {
    float hi2, lo2;
    bool ok;
};
RIP_HD ThrBand thr_band(float slope, const RampPlanDev& pl) {
    const float x = clip_nan(slope, pl.IthreshA_f, pl.IthreshB_f);
#if defined(__CUDA_ARCH__)
    const float thr = pl.thrA_f + pl.thrK_f * __logf(x * pl.invIA_f);
#else
    const float thr = pl.thrA_f + pl.thrK_f * logf(x * pl.invIA_f);
#endif
    ThrBand t;
    t.ok = (thr > 0.0f) && (pl.band < 0.5f);
    const float hi = thr * (1.0f + pl.band), lo = thr * (1.0f - pl.band);
    t.hi2 = hi * hi * (1.0f + 4.0e-7f);
    t.lo2 = lo * lo * (1.0f - 4.0e-7f);
    return t;
}

// jump_detect for one pixel and plan variant v >= 1 (saturation-truncated refits: rare), scalar.
// Same decisions as rip::jump_detect_pixel<.., FAST=true>; the sure-flag / sure-clear tests are done on squares.
template <int G>
RIP_HD_COLD FitResult jump_fast_var(const Ramp<G> rdv, int v, float gain, float read, const RampPlanDev& pl, const double* w_all) {
    const float (&d)[G] = rdv.v;
    FitResult r;
    const int ngrp = pl.var_ngrp[v];
    const int start = pl.start;
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < G; ++t)
        if (t < ngrp) acc = acc + pl.var_K[v][t] * (d[t] - d[1]);
    r.slope = acc;
    const float gc = np_clip<float>(gain, 1e-4f, 1e4f);
    const float dvardt = np_max<float>(r.slope / gc, 0.0f);
    r.err_poisson = sqrtf(np_max<float>(pl.var_coef[v] * dvardt, 0.0f));
    r.err_read = read * pl.var_rfac[v];
    const float sig2read = read * read;
    const ThrBand tb = thr_band(r.slope, pl);
    // sure-flag: delta > 0 and delta^2 > thr_hi^2 var; sure-clear: delta <= 0 (thr > 0) or delta^2 < thr_lo^2 var.
    // var = dvardt*A + read^2*B >= 0 (A, B >= 0 checked by build_plan); NaNs make every comparison false -> unsure.
    bool unsure = !tb.ok;
    uint32_t mask = 0u;
    int s = pl.var_slice_off[v];
#pragma unroll
    for (int i = 0; i < G - 1; ++i) {
        if (i >= start && i < ngrp - 1) {
            const int dimax = (i == ngrp - 2 || ngrp - 1 - start == 2) ? 1 : 2;
#pragma unroll
            for (int di = 1; di <= 2; ++di) {
                if (di <= dimax) {
                    const RampSlice& sl = pl.slices[s];
                    const float diff = d[(i + di < G) ? (i + di) : (G - 1)] - d[i];
                    const float var = dvardt * sl.A + sig2read * sl.B;
                    const float delta = diff * sl.inv_dt - r.slope;
                    const float l2 = delta * delta;
                    const bool pos = delta > 0.0f;
                    const bool sure_set = pos && (l2 > tb.hi2 * var);
                    const bool sure_clr = (delta <= 0.0f) || (l2 < tb.lo2 * var);
                    mask |= sure_set ? (1u << i) : 0u;
                    unsure = unsure || !(sure_set || sure_clr);  // borderline or NaN
                    ++s;
                }
            }
        }
    }
    if (unsure) mask = jump_exact<G>(rdv, v, r.slope, dvardt, sig2read, pl, w_all);
    r.jump_mask = mask;
    return r;
}

// Pair constants of the full-ramp jump test, derived from the plan (host: make_fast_tab; device: __constant__).
// Lane x / y of entry [di-1][k] belong to the slices (i, di) with i = 2k / 2k+1; lanes without a slice hold zeros.
struct FastTab {
    f2 inv_dt[2][RIP_GMAX / 2], A[2][RIP_GMAX / 2], B[2][RIP_GMAX / 2];
    f2 K[RIP_GMAX / 2];  // full-ramp weights (plan variant 0)
    float ratio;         // lower/upper bound of the squared significance band, rounded down
    float pad_;
};
RIP_HD bool full_slice_valid(int G, int start, int i, int di) {
    if (i < start || i + di > G - 1) return false;
    if (di == 2 && (G - 1 - start == 2)) return false;
    return true;
}
inline FastTab make_fast_tab(const RampPlanDev& pl) {
    FastTab ft;
    memset(&ft, 0, sizeof ft);
    const int G = pl.G, start = pl.start;
    for (int di = 1; di <= 2; ++di)
        for (int i = 0; i < G && i < RIP_GMAX; ++i) {
            if (!full_slice_valid(G, start, i, di)) continue;
            const RampSlice& sl = pl.slices[(i - start) * 2 + (di - 1)];
            float* f;
            f = &ft.inv_dt[di - 1][i >> 1].x; f[i & 1] = sl.inv_dt;
            f = &ft.A[di - 1][i >> 1].x; f[i & 1] = sl.A;
            f = &ft.B[di - 1][i >> 1].x; f[i & 1] = sl.B;
        }
    for (int t = 0; t < G && t < RIP_GMAX; ++t) (&ft.K[t >> 1].x)[t & 1] = pl.var_K[0][t];
    const double b = (double)pl.band;
    double r = (b < 0.5) ? ((1.0 - b) / (1.0 + b)) * ((1.0 - b) / (1.0 + b)) * (1.0 - 4.0e-7) / (1.0 + 4.0e-7) : 0.0;
    ft.ratio = (float)(r * (1.0 - 1.0e-6));
    return ft;
}

// one pair of slices: lanes (i0, di), (i0+1, di); V0 / V1 = the lane has a slice (compile time)
template <bool V0, bool V1>
RIP_HD void jump_pair(const f2 diff, const f2 inv_dt, const f2 Ac, const f2 Bc, const float nslope, const f2 hd, const f2 hr,
                      const float ratio, const int i0, const bool lane0_live, uint32_t& mask, bool& unsure) {
    // (classification only: rounding is irrelevant here, anything inside the band is re-evaluated exactly)
    const f2 delta = fma2(diff, inv_dt, bc(nslope));
    const f2 hv = fma2(Ac, hd, mul2(Bc, hr));
    const f2 lv = mul2(hv, bc(ratio));
    if (V0) {
        const float l2s = delta.x * abs_f(delta.x);  // signed square: <= 0 is a sure clear (threshold > 0)
        const bool set = lane0_live && (l2s > hv.x), clr = !lane0_live || (l2s < lv.x);
        mask |= set ? (1u << i0) : 0u;
        unsure = unsure || !(set || clr);
    }
    if (V1) {
        const float l2s = delta.y * abs_f(delta.y);
        const bool set = l2s > hv.y, clr = l2s < lv.y;
        mask |= set ? (2u << i0) : 0u;
        unsure = unsure || !(set || clr);
    }
}

// The pair-wise classification of the full ramp (sure set / sure clear / unsure -> exact), out of line: only pixels
// that fail the all-clear test of jump_full come here (cosmic-ray hits, bright sources, NaNs).
// (TAG = the kernel's P: one copy per kernel instantiation, so that `make check-sass` can compare PTX and SASS counts)
template <int G, int TAG>
RIP_HD_COLD uint32_t jump_classify(const Ramp<G> rd, const int start, const float slope, const float dvardt, const float sig2read,
                                   const ThrBand tb, const RampPlanDev& pl, const FastTab& ft, const double* w_all) {
    f2 q[G / 2];
#pragma unroll
    for (int j = 0; j < G / 2; ++j) q[j] = f2{rd.v[2 * j], rd.v[2 * j + 1]};
    bool unsure = !tb.ok;
    uint32_t mask = 0u;
    const f2 hd = bc(tb.hi2 * dvardt), hr = bc(tb.hi2 * sig2read);
    const float ns = -slope;
    // start = 1 (EXCLUDE_FIRST) only removes the two slices that begin at group 0: lane x of the pairs k = 0
#pragma unroll
    for (int k = 0; k < G / 2; ++k) {
        const bool live0 = (k > 0) || (start == 0);
        // di = 1: (d[2k+1] - d[2k], d[2k+2] - d[2k+1])
        {
            const bool v0 = full_slice_valid(G, 0, 2 * k, 1), v1 = full_slice_valid(G, 0, 2 * k + 1, 1);
            if (v0 || v1) {
                const f2 up = f2{q[k].y, q[(k + 1 < G / 2) ? k + 1 : k].x};  // (the last lane pair has no second slice)
                const f2 diff = sub2p(up, q[k]);
                if (v0 && v1) jump_pair<true, true>(diff, ft.inv_dt[0][k], ft.A[0][k], ft.B[0][k], ns, hd, hr, ft.ratio, 2 * k, live0, mask, unsure);
                else if (v0) jump_pair<true, false>(diff, ft.inv_dt[0][k], ft.A[0][k], ft.B[0][k], ns, hd, hr, ft.ratio, 2 * k, live0, mask, unsure);
                else jump_pair<false, true>(diff, ft.inv_dt[0][k], ft.A[0][k], ft.B[0][k], ns, hd, hr, ft.ratio, 2 * k, live0, mask, unsure);
            }
        }
        // di = 2: (d[2k+2] - d[2k], d[2k+3] - d[2k+1])
        if (k + 1 < G / 2) {
            const bool v0 = full_slice_valid(G, 0, 2 * k, 2), v1 = full_slice_valid(G, 0, 2 * k + 1, 2);
            if (v0 || v1) {
                const f2 diff = sub2p(q[(k + 1 < G / 2) ? k + 1 : k], q[k]);
                if (v0 && v1) jump_pair<true, true>(diff, ft.inv_dt[1][k], ft.A[1][k], ft.B[1][k], ns, hd, hr, ft.ratio, 2 * k, live0, mask, unsure);
                else if (v0) jump_pair<true, false>(diff, ft.inv_dt[1][k], ft.A[1][k], ft.B[1][k], ns, hd, hr, ft.ratio, 2 * k, live0, mask, unsure);
                else jump_pair<false, true>(diff, ft.inv_dt[1][k], ft.A[1][k], ft.B[1][k], ns, hd, hr, ft.ratio, 2 * k, live0, mask, unsure);
            }
        }
    }
    if (unsure) mask = jump_exact<G>(rd, 0, slope, dvardt, sig2read, pl, w_all);
    return mask;
}

// all-clear test of one slice pair: folds  (clear bound) - delta^2  of the live lanes into a running NaN-propagating
// minimum.  The bound is the lower edge of the significance band (times var), so a positive minimum means every slice
// is a sure clear.
template <bool V0, bool V1>
RIP_HD void clear_pair(const f2 diff, const f2 inv_dt, const f2 Ac, const f2 Bc, const float nslope, const f2 ld, const f2 lr,
                       const bool lane0_live, float& tmin) {
    const f2 delta = fma2(diff, inv_dt, bc(nslope));
    const f2 lv = fma2(Ac, ld, mul2(Bc, lr));
    // lo^2 var - delta^2 (classification only: any rounding will do).  A strongly NEGATIVE delta also fails this test
    // although it is a sure clear; that only sends the pixel to jump_classify, which knows (such pixels are almost
    // always neighbours of a jump in the same ramp anyway).
    const f2 t = fma2(f2{-delta.x, -delta.y}, delta, lv);
    if (V0) tmin = min_nan(tmin, lane0_live ? t.x : tmin);
    if (V1) tmin = min_nan(tmin, t.y);
}

// jump_detect of the whole ramp (plan variant 0) for one active pixel; q[j] = groups (2j, 2j+1); start = plan.start (0 / 1).
// slope / errors: the reference's op order (fitting.py:187-212).  Flags: one branch-free pass proves "no slice is
// significant" for almost every pixel (minimum over the slices of  lo^2 var - delta^2  > 0); the rest are
// classified pair-wise out of line and pixels with any slice inside the relative band of the threshold (or NaN)
// redo all slices in the reference's exact f64 op order => the flags are those of the exact path.
template <int G, int TAG>
RIP_HD FitResult jump_full(const f2 (&q)[G / 2], const int start, float gain, float read, const RampPlanDev& pl, const FastTab& ft,
                           const double* w_all) {
    static_assert(G >= 6 && (G & 1) == 0, "pair indexing of the full-ramp specialisation");
    FitResult r;
    const float d1 = q[0].y;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < G / 2; ++j) {
        const f2 pr = mul2(ft.K[j], sub2p(q[j], bc(d1)));  // K_t * (d_t - d_1); the sum stays sequential
        acc = acc + pr.x;
        acc = acc + pr.y;
    }
    r.slope = acc;
    const float gc = clip_nan(gain, 1e-4f, 1e4f);
    const float dvardt = max_nan(r.slope / gc, 0.0f);
    r.err_poisson = sqrtf(max_nan(pl.var_coef[0] * dvardt, 0.0f));
    r.err_read = read * pl.var_rfac[0];
    const float sig2read = read * read;
    const ThrBand tb = thr_band(r.slope, pl);
    const float lo2c = tb.lo2 * (1.0f - 2.0e-6f);  // (absorbs the roundings of the products below)
    const f2 ld = bc(lo2c * dvardt), lr = bc(lo2c * sig2read);
    const float ns = -r.slope;
    float tmin = 3.0e38f;
#pragma unroll
    for (int k = 0; k < G / 2; ++k) {
        const bool live0 = (k > 0) || (start == 0);
        {
            const bool v0 = full_slice_valid(G, 0, 2 * k, 1), v1 = full_slice_valid(G, 0, 2 * k + 1, 1);
            if (v0 || v1) {
                const f2 up = f2{q[k].y, q[(k + 1 < G / 2) ? k + 1 : k].x};
                const f2 diff = sub2p(up, q[k]);
                if (v0 && v1) clear_pair<true, true>(diff, ft.inv_dt[0][k], ft.A[0][k], ft.B[0][k], ns, ld, lr, live0, tmin);
                else if (v0) clear_pair<true, false>(diff, ft.inv_dt[0][k], ft.A[0][k], ft.B[0][k], ns, ld, lr, live0, tmin);
                else clear_pair<false, true>(diff, ft.inv_dt[0][k], ft.A[0][k], ft.B[0][k], ns, ld, lr, live0, tmin);
            }
        }
        if (k + 1 < G / 2) {
            const bool v0 = full_slice_valid(G, 0, 2 * k, 2), v1 = full_slice_valid(G, 0, 2 * k + 1, 2);
            if (v0 || v1) {
                const f2 diff = sub2p(q[(k + 1 < G / 2) ? k + 1 : k], q[k]);
                if (v0 && v1) clear_pair<true, true>(diff, ft.inv_dt[1][k], ft.A[1][k], ft.B[1][k], ns, ld, lr, live0, tmin);
                else if (v0) clear_pair<true, false>(diff, ft.inv_dt[1][k], ft.A[1][k], ft.B[1][k], ns, ld, lr, live0, tmin);
                else clear_pair<false, true>(diff, ft.inv_dt[1][k], ft.A[1][k], ft.B[1][k], ns, ld, lr, live0, tmin);
            }
        }
    }
    uint32_t mask = 0u;
    if (!(tb.ok && tmin > 0.0f)) {  // (NaN anywhere -> tmin is NaN -> not clear)
        Ramp<G> rd;
#pragma unroll
        for (int j = 0; j < G / 2; ++j) { rd.v[2 * j] = q[j].x; rd.v[2 * j + 1] = q[j].y; }
        mask = jump_classify<G, TAG>(rd, start, r.slope, dvardt, sig2read, tb, pl, ft, w_all);
    }
    r.jump_mask = mask;
    return r;
}

template <int G, int TAG>
RIP_HD FitResult ramp_fit_fast(const f2 (&q)[G / 2], GroupFlags& gf, uint32_t& pdq, float gain, float read,
                               const RampPlanDev& pl, const FastTab& ft, const double* w_all) {
    FitResult r = jump_full<G, TAG>(q, pl.start, gain, read, pl, ft, w_all);
    const bool unsat = ((gf.sat >> (G - 1)) & 1u) == 0u;
    if (unsat) gf.jump |= r.jump_mask;
    if (gf.sat) {  // truncated refits only where some group is saturated
        Ramp<G> rd;
#pragma unroll
        for (int j = 0; j < G / 2; ++j) { rd.v[2 * j] = q[j].x; rd.v[2 * j + 1] = q[j].y; }
        for (int iend = G - 1; iend > 2 + pl.start; --iend) {
            const bool layer = ((gf.sat >> iend) & 1u) && !((gf.sat >> (iend - 1)) & 1u);
            if (layer) {
                FitResult t = jump_fast_var<G>(rd, G - iend, gain, read, pl, w_all);
                r.slope = t.slope;
                r.err_read = t.err_read;
                r.err_poisson = t.err_poisson;
                gf.jump |= t.jump_mask;
            }
        }
    }
    const uint32_t allg = (1u << G) - 1u;
    const uint32_t unsat_g = ~gf.sat & allg;
    uint32_t pdq2 = 0u;
    if (gf.jump & unsat_g) pdq2 |= DQ_JUMP_DET;
    if (gf.adf & unsat_g) pdq2 |= DQ_AD_FLOOR;
    if ((gf.dnu & allg) == allg) pdq2 |= DQ_DO_NOT_USE;
    if ((gf.sat >> (1 + pl.start)) & 1u) pdq2 |= DQ_DO_NOT_USE;
    if (gf.sat & allg) pdq2 |= DQ_SATURATED;
    if ((pdq & DQ_REFERENCE_PIXEL) == 0u) pdq |= pdq2;
    return r;
}

// the extrapolating evaluation (some |z| > 1 or NaN): scalar, group by group, exactly as multilin_pixel of v1.  Cold:
// arguments and results travel by value so that the caller's arrays stay in registers.
template <int G, int P>
struct ExtrapIn {
    float z[G];
    float c[P];
};
template <int G>
struct ExtrapOut {
    float phi[G];
    uint32_t dq;
};
template <int G, int P>
RIP_HD_COLD ExtrapOut<G> phi_extrap(const ExtrapIn<G, P> in, uint32_t satm, bool do_not_flag_first) {
    ExtrapOut<G> o;
    o.dq = 0u;
    for (int g = 0; g < G; ++g) {
        bool ex;
        o.phi[g] = legendre_eval<float, P, true>(in.z[g], in.c, P, ex);
        const bool first = (g == 0) && do_not_flag_first;
        if (!first && ex && !((satm >> g) & 1u)) o.dq |= DQ_NO_LIN_CORR;
    }
    return o;
}

// ---- the four stages of a march step ----------------------------------------------------------------------------
struct StepCtx {  // what every stage derives from (tile, tid, band, step); all cheap / CTA-uniform
    int n, tid, tile, x, col, r0, r1, s;
    bool xin, xact;
    unsigned o5[5];  // byte offsets of the ring5 slots of rows s, s+1, .. s+4 (== s-5 .. s-1)
};

// stage a1 : row s-2 (saturation growth, refpix, bias, multilin, D = lin * gain)
template <int G, int P, class SM>
RIP_HD void stage_a1(const Args& A, SM& sm, const Regs<G, P>& R, const StepCtx& C) {
    constexpr int H = G / 4, NQ1 = Regs<G, P>::NQ1;
    const int n = C.n, nb = 4, tid = C.tid, col = C.col, x = C.x, r0 = C.r0, r1 = C.r1;
    const unsigned (&o5)[5] = C.o5;
    const uint32_t allg = (1u << G) - 1u;
    const int row = C.s - 2;
    const bool rowin = in_range(row, imax(r0 - 2, 0), imin(r1 + 2, n));
    f4* dst = sm.D(RIP_OS(-2));
    if (rowin && C.xin) {  // (all columns: the two edge columns of the tile compute values nobody reads)
        uint32_t grown = 0u;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const uint32_t* sr = sm.sat(row + dy) + col;
            grown |= sr[-1] | sr[0] | sr[1];
        }
        const uint32_t own = sm.sat(row)[col];
        grown &= 0xffffu;
        uint32_t satm = grown;
        if (grown) {
            for (int b = 1; b <= A.sat_backup; ++b) satm |= grown >> b;
        }
        satm &= allg & ~1u;
        const uint32_t adf = own >> 16;
        const bool active = C.xact && in_range(row, nb, n - nb);
        // raw u16 -> f32 as group pairs: (2^23 | v) - 2^23 is exact, so the packed subtraction equals the conversion
        f2 Sp[G / 2];
        {
            const uint16_t* rq = sm.raw(RIP_OS(-2)) + tid;
#pragma unroll
            for (int j = 0; j < G / 2; ++j) Sp[j] = u16_pair_to_f32((uint32_t)rq[(2 * j) * TW], (uint32_t)rq[(2 * j + 1) * TW]);
        }
        if (A.do_refpix) {  // gen_cal_image.py:535-556: f64 subtractions, f32 stores
            const int chsel = ((x >> 7) != ((C.tile * TS) >> 7)) ? 1 : 0;
            const d2* rc = (const d2*)sm.rc(RIP_OS(-2));               // one LDS.128 per group pair
            const d2* ln = (const d2*)(sm.ln(RIP_OS(-2)) + chsel * G);
#pragma unroll
            for (int j = 0; j < G / 2; ++j) {
                const f2 dk = f2{r1w<NQ1>(R.r1, 2 * j), r1w<NQ1>(R.r1, 2 * j + 1)};
                const f2 v0 = sub2p(Sp[j], dk);  // (no product feeds these packed additions)
                const d2 rcv = rc[j], lnv = ln[j];
                float vx = (float)((double)v0.x - rcv.x), vy = (float)((double)v0.y - rcv.y);
                vx = (float)((double)vx - lnv.x);
                vy = (float)((double)vy - lnv.y);
                Sp[j] = add2p(f2{vx, vy}, dk);
            }
        }
        // biascorr (embedded with zeros outside the active region: v - 0 == v)
        f2 S2[G / 2];
#pragma unroll
        for (int j = 0; j < G / 2; ++j)
            S2[j] = sub2p(Sp[j], f2{r1w<NQ1>(R.r1, G + 2 * j), r1w<NQ1>(R.r1, G + 2 * j + 1)});
        const float Smin = r1w<NQ1>(R.r1, 2 * G), Smax = r1w<NQ1>(R.r1, 2 * G + 1), Sref = r1w<NQ1>(R.r1, 2 * G + 2);
        const float gain = r1w<NQ1>(R.r1, 2 * G + 3);
        const uint32_t aux = f_as_u(r1w<NQ1>(R.r1, 2 * G + 4));
        float c[P];
#pragma unroll
        for (int L = 0; L < P; ++L) c[L] = r1w<NQ1>(R.r1, 2 * G + 5 + L);
        // z = -1 + (2 (S - Smin)) / (Smax - Smin)      (ipc_linearity.py:330)
        SharedDiv sd;
        const float den = Smax - Smin;
        sd.init(den);
        const bool div_ok = sd.ok && (Smin > -1.0e18f) && (Smin < 1.0e18f);
        f2 z2[G / 2];
#pragma unroll
        for (int j = 0; j < G / 2; ++j) {
            const f2 num = mul2(bc(2.0f), sub2p(S2[j], bc(Smin)));
            f2 q;
            if (div_ok) q = sd.div2(num);
            else q = f2{num.x / den, num.y / den};
            z2[j] = add2p(bc(-1.0f), q);
        }
        if (A.do_not_flag_first) z2[0].x = np_clip<float>(z2[0].x, -1.0f, 1.0f);
        // |z| > 1 anywhere (or NaN) -> the extrapolating scalar evaluation of v1 for this pixel (rare).  On the bit
        // patterns: (bits & 0x7fffffff) > bits(1.0f) is true for |z| > 1, infinities and NaNs alike.
        uint32_t zmax = 0u;
#pragma unroll
        for (int j = 0; j < G / 2; ++j) {
            const uint32_t ax = f_as_u(z2[j].x) & 0x7fffffffu, ay = f_as_u(z2[j].y) & 0x7fffffffu;
            zmax = zmax > ax ? zmax : ax;
            zmax = zmax > ay ? zmax : ay;
        }
        const bool anyex = zmax > 0x3f800000u;
        uint32_t dq = (aux & 1u) ? DQ_REFERENCE_PIXEL : 0u;
        f2 phi2[G / 2];
        if (!anyex) {
            f2 prev[G / 2], cur[G / 2];
#pragma unroll
            for (int j = 0; j < G / 2; ++j) { phi2[j] = bc(c[0]); prev[j] = bc(1.0f); cur[j] = z2[j]; }
#pragma unroll
            for (int L = 1; L < P; ++L) {
                const float a = (float)((2 * L + 1) / (double)(L + 1)), b = (float)(L / (double)(L + 1));
#pragma unroll
                for (int j = 0; j < G / 2; ++j) {
                    phi2[j] = add2p(phi2[j], mul2x(bc(c[L]), cur[j], A.negzero));
                    if (L + 1 < P) {  // the recursion value of the last order is never used
                        // (L = 1: prev is exactly 1, b * 1 == b)
                        const f2 bp = (L == 1) ? bc(b) : mul2x(bc(b), prev[j], A.negzero);
                        const f2 nxt = sub2p(mul2x(mul2(bc(a), z2[j]), cur[j], A.negzero), bp);
                        prev[j] = cur[j];
                        cur[j] = nxt;
                    }
                }
            }
        } else {
            ExtrapIn<G, P> ein;
#pragma unroll
            for (int j = 0; j < G / 2; ++j) { ein.z[2 * j] = z2[j].x; ein.z[2 * j + 1] = z2[j].y; }
#pragma unroll
            for (int L = 0; L < P; ++L) ein.c[L] = c[L];
            const ExtrapOut<G> eo = phi_extrap<G, P>(ein, satm, A.do_not_flag_first != 0);
#pragma unroll
            for (int j = 0; j < G / 2; ++j) phi2[j] = f2{eo.phi[2 * j], eo.phi[2 * j + 1]};
            dq |= eo.dq;
        }
        if (aux & 1u) {  // lin dq has NO_LIN_CORR | REFERENCE_PIXEL: S - Sref instead (ipc_linearity.py:334-336)
#pragma unroll
            for (int j = 0; j < G / 2; ++j) phi2[j] = sub2p(S2[j], bc(Sref));
        }
        if (active) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const f2 a = mul2(phi2[2 * h], bc(gain)), b = mul2(phi2[2 * h + 1], bc(gain));
                dst[h * RW + col] = f4{a.x, a.y, b.x, b.y};
            }
        } else {
#pragma unroll
            for (int h = 0; h < H; ++h) dst[h * RW + col] = f4{0.f, 0.f, 0.f, 0.f};
            if (A.lincube && row >= r0 && row < r1 && (tid >= 4 || C.tile == 0) && tid < TW - 4) {
                const unsigned npl = (unsigned)n * (unsigned)n;
#pragma unroll
                for (int g = 0; g < G; ++g)
                    A.lincube[(unsigned)g * npl + (unsigned)row * (unsigned)n + (unsigned)x] = (g & 1) ? phi2[g >> 1].y : phi2[g >> 1].x;
            }
        }
        sm.flg(RIP_OS(-2))[tid] = satm | (adf << 16);
        sm.nlc(RIP_OS(-2))[tid] = (uint8_t)(((dq & DQ_NO_LIN_CORR) ? 1u : 0u) | ((aux & 2u) ? 4u : 0u));
    }
    // (rows of the band range outside the frame, columns beyond the frame: D / flags are never read there)
}

// stage b : row s-4 (IPC pass 1:  O1 = (D + D) - K (*) D)
template <int G, int P, class SM>
RIP_HD void stage_b(const Args& A, SM& sm, const Regs<G, P>& R, const StepCtx& C) {
    constexpr int H = G / 4;
    const int n = C.n, nb = 4, tid = C.tid, col = C.col, r0 = C.r0, r1 = C.r1;
    const unsigned (&o5)[5] = C.o5;
    const int row = C.s - 4;
    const bool rowok = in_range(row, imax(r0 - 1, nb), imin(r1 + 1, n - nb));
    f4* o = sm.O1(row);
    if (rowok) {  // (all columns; the outermost two on each side of the tile compute values nobody reads)
        const float k[9] = {R.kb[0].x, R.kb[0].y, R.kb[0].z, R.kb[0].w, R.kb[1].x, R.kb[1].y, R.kb[1].z, R.kb[1].w, R.kb8};
        const f4* dm = sm.D(RIP_OS(-5));
        const f4* d0 = sm.D(RIP_OS(-4));
        const f4* dp = sm.D(RIP_OS(-3));
#pragma unroll
        for (int h = 0; h < H; ++h) {
            f2 lo, hi;
            stencil9(dm + h * RW, d0 + h * RW, dp + h * RW, col, k, A.negzero, lo, hi);
            const f4 dc = d0[h * RW + col];
            const f2 clo{dc.x, dc.y}, chi{dc.z, dc.w};
            const f2 rlo = sub2p(add2p(clo, clo), lo), rhi = sub2p(add2p(chi, chi), hi);  // output + image2 - ipc_fwd(output)
            o[h * RW + col] = C.xact ? f4{rlo.x, rlo.y, rhi.x, rhi.y} : f4{0.f, 0.f, 0.f, 0.f};  // reference columns stay 0
        }
    } else if (in_range(row, r0 - 1, r1 + 1)) {  // reference-pixel rows: zeros for the stencil of stage c
#pragma unroll
        for (int h = 0; h < H; ++h) o[h * RW + col] = f4{0.f, 0.f, 0.f, 0.f};
    }
}

// stage b, float64 taps (K64): O1 = f64(f32(D + D)) - K (*) D with every product and sum a separate float64 operation in
// the reference's accumulation order (utils/ipc_linearity.py:69-94; the image is float32, the kernel float64 -> NumPy
// promotes the products).  D is read as float4 (4 groups) and converted on the fly.
#define RIP_TAP64(V, Q)                  \
    {                                    \
        const f4 t = (V);                \
        const double kq = k[Q];          \
        a0 = a0 + (double)t.x * kq;      \
        a1 = a1 + (double)t.y * kq;      \
        a2 = a2 + (double)t.z * kq;      \
        a3 = a3 + (double)t.w * kq;      \
    }
template <int G, int P, class SM>
RIP_HD void stage_b64(const Args& A, SM& sm, const Regs<G, P>& R, const StepCtx& C) {
    constexpr int H = G / 4;
    const int n = C.n, nb = 4, col = C.col, r0 = C.r0, r1 = C.r1;
    const unsigned (&o5)[5] = C.o5;
    const int row = C.s - 4;
    const bool rowok = in_range(row, imax(r0 - 1, nb), imin(r1 + 1, n - nb));
    d2* o = sm.O1d(row);
    if (rowok) {
        const double (&k)[9] = R.kbd;
        const f4* dm = sm.D(RIP_OS(-5));
        const f4* d0 = sm.D(RIP_OS(-4));
        const f4* dp = sm.D(RIP_OS(-3));
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const f4 c = d0[h * RW + col];
            double a0 = (double)c.x * k[0], a1 = (double)c.y * k[0], a2 = (double)c.z * k[0], a3 = (double)c.w * k[0];
            RIP_TAP64(dm[h * RW + col], 1)
            RIP_TAP64(dp[h * RW + col], 2)
            RIP_TAP64(d0[h * RW + col - 1], 3)
            RIP_TAP64(d0[h * RW + col + 1], 4)
            RIP_TAP64(dm[h * RW + col - 1], 5)
            RIP_TAP64(dm[h * RW + col + 1], 6)
            RIP_TAP64(dp[h * RW + col - 1], 7)
            RIP_TAP64(dp[h * RW + col + 1], 8)
            // output + image2 (both float32 arrays: float32 sum) - ipc_fwd(output) (float64)
            const double r0_ = (double)(c.x + c.x) - a0, r1_ = (double)(c.y + c.y) - a1;
            const double r2_ = (double)(c.z + c.z) - a2, r3_ = (double)(c.w + c.w) - a3;
            o[(2 * h) * RW + col] = C.xact ? d2{r0_, r1_} : d2{0.0, 0.0};
            o[(2 * h + 1) * RW + col] = C.xact ? d2{r2_, r3_} : d2{0.0, 0.0};
        }
    } else if (in_range(row, r0 - 1, r1 + 1)) {
#pragma unroll
        for (int j = 0; j < G / 2; ++j) o[j * RW + col] = d2{0.0, 0.0};
    }
}
#undef RIP_TAP64

// stage c : row s-6 (IPC pass 2, /gain; ramp fit, jump flags, DQ propagation; dark, error split, flat/area; stores)
struct NoHook {
    RIP_HD void operator()() const {}
};
// `reload` runs once per call, after the ramp fit (the register peak of the stage), when the staged record R.kc /
// R.area* lives on only in a few locals: role Y of v3 issues the loads of the NEXT row there, so that they fly during
// the epilogue and the wait at the step barrier instead of being consumed right after their issue.
// everything of stage c behind the IPC correction: ramp fit, flags, dark / error split / flat, stores
template <int G, int P, class SM, typename Hook>
RIP_HD void stage_c_tail(const Args& A, const RampPlanDev& pl, const FastTab& ft, SM& sm, const StepCtx& C, const f2 (&q)[G / 2],
                         const uint32_t fl, const uint32_t nlc, const float gval, const float readv, const float dsl, const float flat,
                         const uint32_t sdq, const float area32, const double area64, const unsigned p, const bool active, Hook reload) {
    const int n = C.n, nb = 4, na = n - 8, x = C.x;
    const uint32_t allg = (1u << G) - 1u;
    const unsigned npl = (unsigned)n * (unsigned)n;
    const int row = C.s - SM::C_LAG;
    GroupFlags gf;
    gf.sat = fl & 0xffffu;
    gf.adf = fl >> 16;
    gf.dnu = gf.adf | (A.exclude_first ? 1u : 0u);
    gf.jump = 0u;
    gf.other_unsat = 0u;
    uint32_t pd = (nlc & 4u) ? DQ_REFERENCE_PIXEL : 0u;
    FitResult r;
    if (active) {
        r = ramp_fit_fast<G, P>(q, gf, pd, gval, readv, pl, ft, A.w_exact);
    } else {
        // reference pixels / phantom border: the fit result is zeroed by the packaging step
        // (gen_cal_image.py:470-472); only the flag propagation of ramp_fit matters (fitting.py:340-353)
        r.slope = 0.0f; r.err_read = 0.0f; r.err_poisson = 0.0f; r.jump_mask = 0u;
        const uint32_t unsat_g = ~gf.sat & allg;
        uint32_t pdq2 = 0u;
        if (gf.adf & unsat_g) pdq2 |= DQ_AD_FLOOR;
        if ((gf.dnu & allg) == allg) pdq2 |= DQ_DO_NOT_USE;
        if ((gf.sat >> (1 + pl.start)) & 1u) pdq2 |= DQ_DO_NOT_USE;
        if (gf.sat & allg) pdq2 |= DQ_SATURATED;
        if ((pd & DQ_REFERENCE_PIXEL) == 0u) pd |= pdq2;
    }
    reload();  // (after the ramp fit, the register peak of the stage; before the epilogue)
    const uint32_t pdq = sdq | ((nlc & 1u) ? DQ_NO_LIN_CORR : 0u) | (pd & ~DQ_REFERENCE_PIXEL);
    float fa = flat;
    if (A.area) {
        if (A.area_dtype == RIP_F64) fa = (float)((double)fa / area64);
        else fa = fa / area32;
    }
    l2_epilogue(r, active, dsl, fa);
    st_out(A.slope + p, r.slope);
    st_out(A.err_read + p, r.err_read);
    st_out(A.err_poisson + p, r.err_poisson);
    st_out(A.pdq + p, pdq);
    if (A.endslice && active) {
        // group where SATURATED first appears, minus one (gen_cal_image.py:703-708: the last 0->1 transition wins)
        const uint32_t tr = gf.sat & ~(gf.sat << 1) & ~1u & allg;
        int es = -1;
        if (tr) {
#if defined(__CUDA_ARCH__)
            es = 30 - __clz((int)tr);
#else
            int hb = 0;
            for (int g = 0; g < G; ++g) if ((tr >> g) & 1u) hb = g;
            es = hb - 1;
#endif
        }
        A.endslice[(unsigned)(row - nb) * (unsigned)na + (unsigned)(x - nb)] = (int8_t)es;
    }
    if (A.rdq) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            uint32_t b = 0u;
            if ((gf.dnu >> g) & 1u) b |= DQ_DO_NOT_USE;
            if ((gf.sat >> g) & 1u) b |= DQ_SATURATED;
            if ((gf.jump >> g) & 1u) b |= DQ_JUMP_DET;
            if ((gf.adf >> g) & 1u) b |= DQ_AD_FLOOR;
            A.rdq[(unsigned)g * npl + p] = (uint8_t)b;
        }
    }
}


// `reads_done` runs exactly once per call, after the stage's last read of the shared-memory rows other threads will
// overwrite (O1 rows and D of the stage's row): v6 arrives at its split end-of-step barrier there.
template <int G, int P, class SM, typename Hook = NoHook, typename Hook2 = NoHook>
RIP_HD void stage_c(const Args& A, const RampPlanDev& pl, const FastTab& ft, SM& sm, const Regs<G, P>& R, const StepCtx& C,
                    Hook reload = Hook(), Hook2 reads_done = Hook2()) {
    constexpr int H = G / 4;
    const int n = C.n, nb = 4, na = n - 8, tid = C.tid, col = C.col, x = C.x, r0 = C.r0, r1 = C.r1;
    const unsigned (&o5)[5] = C.o5;
    const uint32_t allg = (1u << G) - 1u;
    const unsigned npl = (unsigned)n * (unsigned)n;
    const int row = C.s - SM::C_LAG;
    const bool out_col = (tid >= 4 || C.tile == 0) && tid < TW - 4 && C.xin;
    const bool c_on = in_range(row, r0, r1) && out_col;
    if (!c_on) {
        reads_done();
        reload();
        return;
    }
    const unsigned p = R.orow - (unsigned)SM::C_LAG * (unsigned)n + (unsigned)x;
    const bool active = C.xact && in_range(row, nb, n - nb);
    const uint32_t fl = sm.flg(RIP_OS(-SM::C_LAG))[tid];
    const uint32_t nlc = sm.nlc(RIP_OS(-SM::C_LAG))[tid];
    const float gval = R.kc[2].y, readv = R.kc[2].z, dsl = R.kc[2].w, flat = R.kc[3].x;
    const uint32_t sdq = f_as_u(R.kc[3].y);
    const float area32 = R.area32;
    const double area64 = R.area64;
    f2 q[G / 2];
    if (active) {
        const float k[9] = {R.kc[0].x, R.kc[0].y, R.kc[0].z, R.kc[0].w, R.kc[1].x, R.kc[1].y, R.kc[1].z, R.kc[1].w, R.kc[2].x};
        SharedDiv sd;
        sd.init(gval);
        const f4* om = sm.O1(row - 1);
        const f4* o0 = sm.O1(row);
        const f4* op = sm.O1(row + 1);
        const f4* dd = sm.D(RIP_OS(-SM::C_LAG));
        f2 t[G / 2];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            f2 lo, hi;
            stencil9(om + h * RW, o0 + h * RW, op + h * RW, col, k, A.negzero, lo, hi);
            const f4 oc = o0[h * RW + col], dc = dd[h * RW + col];
            t[2 * h] = sub2p(add2p(f2{oc.x, oc.y}, f2{dc.x, dc.y}), lo);  // (output + image2) - ipc_fwd(output)
            t[2 * h + 1] = sub2p(add2p(f2{oc.z, oc.w}, f2{dc.z, dc.w}), hi);
        }
        bool slow = !sd.ok;
        if (!slow) {
            f2 chk = f2{0.f, 0.f};
#pragma unroll
            for (int j = 0; j < G / 2; ++j) { q[j] = sd.div2(t[j]); chk = add2p(chk, q[j]); }
            const float tt = chk.x + chk.y;
            slow = !(tt == tt);  // a NaN from the correction steps (infinite numerator) -> true division
        }
        if (slow) {
#pragma unroll
            for (int j = 0; j < G / 2; ++j) q[j] = f2{t[j].x / gval, t[j].y / gval};
        }
        if (A.lincube) {
#pragma unroll
            for (int g = 0; g < G; ++g) A.lincube[(unsigned)g * npl + p] = (g & 1) ? q[g >> 1].y : q[g >> 1].x;
        }
    } else {
#pragma unroll
        for (int j = 0; j < G / 2; ++j) q[j] = f2{0.f, 0.f};  // unused: every output of a non-active pixel is flag-only
    }
    reads_done();
    stage_c_tail<G, P>(A, pl, ft, sm, C, q, fl, nlc, gval, readv, dsl, flat, sdq, area32, area64, p, active, reload);
}

// stage c with float64 taps (K64): IPC pass 2 and the division by the gain in float64, the float32 store of the
// reference's `data[j] = ...` (utils/ipc_linearity.py:185-186), then the common tail.
template <int G, int P, class SM, typename Hook = NoHook>
RIP_HD void stage_c64(const Args& A, const RampPlanDev& pl, const FastTab& ft, SM& sm, const Regs<G, P>& R, const StepCtx& C,
                      Hook reload = Hook()) {
    constexpr int H = G / 4;
    const int n = C.n, nb = 4, tid = C.tid, col = C.col, x = C.x, r0 = C.r0, r1 = C.r1;
    const unsigned (&o5)[5] = C.o5;
    const unsigned npl = (unsigned)n * (unsigned)n;
    const int row = C.s - SM::C_LAG;
    const bool out_col = (tid >= 4 || C.tile == 0) && tid < TW - 4 && C.xin;
    const bool c_on = in_range(row, r0, r1) && out_col;
    if (!c_on) {
        reload();
        return;
    }
    const unsigned p = R.orow - (unsigned)SM::C_LAG * (unsigned)n + (unsigned)x;
    const bool active = C.xact && in_range(row, nb, n - nb);
    const uint32_t fl = sm.flg(RIP_OS(-SM::C_LAG))[tid];
    const uint32_t nlc = sm.nlc(RIP_OS(-SM::C_LAG))[tid];
    // record: words 0..17 taps (doubles), 18 gain, 19 read, 20 dark slope, 21 flat, 22 static dq
    const float gval = R.kc64[4].z, readv = R.kc64[4].w, dsl = R.kc64[5].x, flat = R.kc64[5].y;
    const uint32_t sdq = f_as_u(R.kc64[5].z);
    const float area32 = R.area32;
    const double area64 = R.area64;
    f2 q[G / 2];
    if (active) {
        double k[9];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            k[2 * t] = f2_as_double(R.kc64[t].x, R.kc64[t].y);
            k[2 * t + 1] = f2_as_double(R.kc64[t].z, R.kc64[t].w);
        }
        k[8] = f2_as_double(R.kc64[4].x, R.kc64[4].y);
        const d2* om = sm.O1d(row - 1);
        const d2* o0 = sm.O1d(row);
        const d2* op = sm.O1d(row + 1);
        const f4* dd = sm.D(RIP_OS(-SM::C_LAG));
        const double gd = (double)gval;
#pragma unroll
        for (int j = 0; j < G / 2; ++j) {
            const d2 c = o0[j * RW + col];
            double a0 = c.x * k[0], a1 = c.y * k[0];
#define RIP_TAPD(V, Q)              \
    {                               \
        const d2 t = (V);           \
        a0 = a0 + t.x * k[Q];       \
        a1 = a1 + t.y * k[Q];       \
    }
            RIP_TAPD(om[j * RW + col], 1)
            RIP_TAPD(op[j * RW + col], 2)
            RIP_TAPD(o0[j * RW + col - 1], 3)
            RIP_TAPD(o0[j * RW + col + 1], 4)
            RIP_TAPD(om[j * RW + col - 1], 5)
            RIP_TAPD(om[j * RW + col + 1], 6)
            RIP_TAPD(op[j * RW + col - 1], 7)
            RIP_TAPD(op[j * RW + col + 1], 8)
#undef RIP_TAPD
            const f4 dc = dd[(j >> 1) * RW + col];
            const double dx = (double)((j & 1) ? dc.z : dc.x), dy = (double)((j & 1) ? dc.w : dc.y);
            const double o2x = (c.x + dx) - a0, o2y = (c.y + dy) - a1;  // (output + image2) - ipc_fwd(output)
            q[j] = f2{(float)(o2x / gd), (float)(o2y / gd)};
        }
        if (A.lincube) {
#pragma unroll
            for (int g = 0; g < G; ++g) A.lincube[(unsigned)g * npl + p] = (g & 1) ? q[g >> 1].y : q[g >> 1].x;
        }
    } else {
#pragma unroll
        for (int j = 0; j < G / 2; ++j) q[j] = f2{0.f, 0.f};
    }
    (void)H;
    stage_c_tail<G, P>(A, pl, ft, sm, C, q, fl, nlc, gval, readv, dsl, flat, sdq, area32, area64, p, active, reload);
}

// stage a0 : row s (raw -> cumulative saturation / A-D floor bits)
template <int G, int P, class SM>
RIP_HD void stage_a0(const Args& A, SM& sm, const StepCtx& C) {
    const int n = C.n, tid = C.tid, col = C.col, r0 = C.r0, r1 = C.r1;
    const unsigned (&o5)[5] = C.o5;
    const int row = C.s;
    uint32_t bits = 0u;
    const bool rowin = in_range(row, imax(r0 - 3, 0), imin(r1 + 3, n));
    if (rowin && C.xin) {
        const uint16_t* rq = sm.raw(RIP_OS(0)) + tid;
        const float thr = sm.thr(RIP_OS(0))[tid];
        uint32_t rv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) rv[g] = rq[g * TW];
        // fast exit: no group (>= 1; saturation_check skips the first resultant, gen_cal_image.py:174-180) reaches
        // the threshold or the A/D floor
        uint32_t mx = rv[1], mn = rv[1];
#pragma unroll
        for (int g = 2; g < G; ++g) { mx = mx > rv[g] ? mx : rv[g]; mn = mn < rv[g] ? mn : rv[g]; }
        if (!(u16_to_f32(mx) < thr) || mn == 0u) {  // (NaN thresholds were replaced by +inf when the CALDIR was loaded)
            bool cum = false;
#pragma unroll
            for (int g = 1; g < G; ++g) {
                const float fv = u16_to_f32(rv[g]);
                cum = cum || (fv >= thr);
                if (cum) bits |= 1u << g;
                if (fv <= 0.0f) bits |= 1u << (16 + g);
            }
        }
    }
    sm.sat(row)[col] = bits;
}

// ---- one march step --------------------------------------------------------------------------------------------
// Stage rows: a0 row s, a1 row s-2, b row s-4, c row s-6; every stage only reads ring slots written in earlier steps,
// so one barrier per step suffices and the stages may run in any order.  Order and load placement (see Regs):
//     [kb <- kbn: the one scoreboard wait]  [cp.async row s+2; Lb(next)]  a1  c  [Lc(next), L1(next)]  b  a0  barrier
// o5s = (s mod RING) * ROW5, the byte offset of row s's ring5 slot, carried by the caller (next_o5).
template <int G, int P, bool K64 = false>
RIP_HD void step(const Args& A, const RampPlanDev& pl, const FastTab& ft, Smem<G, K64>& sm, Regs<G, P>& R, const int tid,
                 const int tile, const int r0, const int r1, const int s, const unsigned o5s) {
    StepCtx C;
    C.n = A.n; C.tid = tid; C.tile = tile; C.r0 = r0; C.r1 = r1; C.s = s;
    C.x = tile * TS + tid;
    C.col = tid + 1;
    C.xin = C.x < A.n;
    C.xact = in_range(C.x, 4, A.n - 4);
    constexpr unsigned RB = Smem<G>::ROW5, RING_B = RING * Smem<G>::ROW5;
    C.o5[0] = o5s; C.o5[1] = wrap5(o5s + RB, RING_B); C.o5[2] = wrap5(o5s + 2 * RB, RING_B);
    C.o5[3] = wrap5(o5s + 3 * RB, RING_B); C.o5[4] = wrap5(o5s + 4 * RB, RING_B);
    const unsigned (&o5)[5] = C.o5;

    // The one scoreboard wait of the step: `dep` is zero at run time (Args::pad_), but ptxas cannot know, so the loads
    // of load_bn -- and with them everything below -- are ordered after the arrival of all loads in flight.
    if (K64) {
        const unsigned dep = f_as_u(R.kc64[0].x) & (unsigned)A.pad_;
#pragma unroll
        for (int t = 0; t < 9; ++t) R.kbd[t] = R.kbnd[t];
        row_async<G, P>(A, sm, R, s + 2, 2, RIP_O5(2), tile, tid, r0 - 3, r1 + 3);
        prefetch_records<G, P, KQ64>(A, s, tile, tid, r0, r1);
        load_bn64<G, P>(A, R, fold_row(s - 3, r0 - 1, r1 + 1), tile, tid, dep);
        stage_a1<G, P>(A, sm, R, C);
        stage_c64<G, P>(A, pl, ft, sm, R, C);
        load_c64<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
        load_a1<G, P>(A, R, fold_row(s - 1, r0 - 2, r1 + 2), tile, tid);
        stage_b64<G, P>(A, sm, R, C);
        stage_a0<G, P>(A, sm, C);
        R.orow += (unsigned)A.n;
        cp_async_wait<1>();
        return;
    }
    const unsigned dep = f_as_u(R.kc[0].x) & (unsigned)A.pad_;
    R.kb[0] = R.kbn[0]; R.kb[1] = R.kbn[1]; R.kb8 = R.kbn8;
    row_async<G, P>(A, sm, R, s + 2, 2, RIP_O5(2), tile, tid, r0 - 3, r1 + 3);
    prefetch_records<G, P, KQ>(A, s, tile, tid, r0, r1);
    load_bn<G, P>(A, R, fold_row(s - 3, r0 - 1, r1 + 1), tile, tid, dep);

    stage_a1<G, P>(A, sm, R, C);
    stage_c<G, P>(A, pl, ft, sm, R, C);
    load_c<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
    load_a1<G, P>(A, R, fold_row(s - 1, r0 - 2, r1 + 2), tile, tid);
    stage_b<G, P>(A, sm, R, C);
    stage_a0<G, P>(A, sm, C);

    R.orow += (unsigned)A.n;
    cp_async_wait<1>();  // the rows issued in the previous step (row s+1) have landed; the caller's barrier publishes them
}

// prologue: offsets, the cp.async rows of the first two steps, the records the first step consumes, ring pads
template <int G, int P, bool K64 = false>
RIP_HD void prologue(const Args& A, Smem<G, K64>& sm, Regs<G, P>& R, int tid, int tile, int r0, int r1) {
    const int x = tile * TS + tid;
    const bool xin = x < A.n;
    const int s0 = r0 - 3;
    constexpr unsigned RB = Smem<G>::ROW5, RING_B = RING * Smem<G>::ROW5;
    const unsigned o5s = (unsigned)mod_pos(s0, RING) * RB;
    R.orow = (unsigned)(s0 * A.n);  // (mod 2^32 for s0 < 0: only ever used after adding back a non-negative row offset)
    row_async<G, P>(A, sm, R, s0, 0, o5s, tile, tid, r0 - 3, r1 + 3);
    row_async<G, P>(A, sm, R, s0 + 1, 1, wrap5(o5s + RB, RING_B), tile, tid, r0 - 3, r1 + 3);
    if (K64) {
        load_c64<G, P>(A, R, fold_row(s0 - 6, r0 - 1, r1 + 1), tile, tid, x, xin);
        load_bn64<G, P>(A, R, fold_row(s0 - 4, r0 - 1, r1 + 1), tile, tid, 0u);
    } else {
        load_c<G, P>(A, R, fold_row(s0 - 6, r0 - 1, r1 + 1), tile, tid, x, xin);
        load_bn<G, P>(A, R, fold_row(s0 - 4, r0 - 1, r1 + 1), tile, tid, 0u);
    }
    load_a1<G, P>(A, R, fold_row(s0 - 2, r0 - 2, r1 + 2), tile, tid);
    // ring pads and the slots stage a1 / b read before anything was written there
    // (D of every ring5 slot and the whole ring4; never the cp.async targets)
    for (int k = 0; k < RING; ++k)
        for (int i = tid; i < Smem<G>::H * RW; i += TW) sm.D((unsigned)k * RB)[i] = f4{0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < O_DEPTH * Smem<G, K64>::ROW4 / 16; i += TW) ((f4*)sm.r4)[i] = f4{0.f, 0.f, 0.f, 0.f};
    cp_async_wait<0>();
}

// ring5 offset of the next row (the march loop's carried variable)
template <int G>
RIP_HD unsigned next_o5(unsigned o5s) { return (o5s == (RING - 1) * Smem<G>::ROW5) ? 0u : o5s + Smem<G>::ROW5; }
template <int G>
RIP_HD unsigned first_o5(int r0) { return (unsigned)mod_pos(r0 - 3, RING) * Smem<G>::ROW5; }

// =================================================================================================================
// v6: the v2 stages re-scheduled for FIVE resident CTAs per SM (96 registers, 43.8 KB of rings at G = 8).
//
//   * stage c follows stage b within the step (row s-5 instead of s-6) behind a SECOND barrier: the D ring shrinks from 5
//     to 4 rows and the O1 ring from 4 to 3; raw rows are copied ONE row ahead (they are L2-prefetched two ahead), so every
//     per-row record (D | raw | thr | flg | rc | ln | nlc | sat) lives in one depth-4 ring addressed by (row & 3);
//   * the records are no longer staged a whole step ahead in registers (66 of v2's 128): the IPC taps of stage b are
//     loaded at the top of the step, the record of stage c right before stage b (it flies during b and the barrier), the
//     record of a1 after stage c as before -- every load hits L2 thanks to prefetch_records (which runs one step ahead
//     of the loads: rec1 of row s, recK of row s-2 at step s).
//   step order:  [async row s+1; prefetch]  a1  b  [Lc]  | barrier |  c  [L1(next) Lb(next)]  a0  | barrier |
// =================================================================================================================
RIP_HD int mod3_pos(int a) { int r = a % 3; return r < 0 ? r + 3 : r; }

template <int G, bool K64 = false>
struct Smem6 {
    static constexpr int H = G / 4;
    static constexpr int DEPTH = 4;
    static constexpr int C_LAG = 5;
    static constexpr int OFF_D = 0;
    static constexpr int OFF_RAW = OFF_D + 16 * H * RW;
    static constexpr int OFF_THR = OFF_RAW + 2 * G * TW;
    static constexpr int OFF_FLG = OFF_THR + 4 * TW;
    static constexpr int OFF_RC = OFF_FLG + 4 * TW;
    static constexpr int OFF_LN = OFF_RC + 8 * G;
    static constexpr int OFF_NLC = OFF_LN + 16 * G;
    static constexpr int OFF_SAT = (OFF_NLC + TW + 15) / 16 * 16;
    static constexpr int ROW5 = (OFF_SAT + 4 * RW + 15) / 16 * 16;  // bytes of one row record
    static constexpr int ROWO = (K64 ? 32 : 16) * H * RW;           // bytes of one O1 row (float64 taps: O1 in doubles)
    unsigned char* r5;
    unsigned char* ro;
    RIP_HD static size_t bytes() { return (size_t)DEPTH * ROW5 + (size_t)3 * ROWO + 64; }
    RIP_HD void carve(unsigned char* base) {
        r5 = base;
        ro = base + (size_t)DEPTH * ROW5;
    }
    RIP_HD f4* D(unsigned o5) const { return (f4*)(r5 + o5 + OFF_D); }
    RIP_HD uint16_t* raw(unsigned o5) const { return (uint16_t*)(r5 + o5 + OFF_RAW); }
    RIP_HD float* thr(unsigned o5) const { return (float*)(r5 + o5 + OFF_THR); }
    RIP_HD uint32_t* flg(unsigned o5) const { return (uint32_t*)(r5 + o5 + OFF_FLG); }
    RIP_HD double* rc(unsigned o5) const { return (double*)(r5 + o5 + OFF_RC); }
    RIP_HD double* ln(unsigned o5) const { return (double*)(r5 + o5 + OFF_LN); }
    RIP_HD uint8_t* nlc(unsigned o5) const { return (uint8_t*)(r5 + o5 + OFF_NLC); }
    RIP_HD uint32_t* sat(int row) const { return (uint32_t*)(r5 + (unsigned)(row & 3) * ROW5 + OFF_SAT); }
    RIP_HD f4* O1(int row) const { return (f4*)(ro + (unsigned)mod3_pos(row) * ROWO); }
    RIP_HD d2* O1d(int row) const { return (d2*)(ro + (unsigned)mod3_pos(row) * ROWO); }  // [G/2][RW]
};

template <int G>
RIP_HD void make_ctx6(StepCtx& C, const Args& A, int tid, int tile, int r0, int r1, int s) {
    C.n = A.n; C.tid = tid; C.tile = tile; C.r0 = r0; C.r1 = r1; C.s = s;
    C.x = tile * TS + tid;
    C.col = tid + 1;
    C.xin = C.x < A.n;
    C.xact = in_range(C.x, 4, A.n - 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) C.o5[k] = (unsigned)((s + k) & 3) * (unsigned)Smem6<G>::ROW5;
    C.o5[4] = 0u;
}

// L2 prefetch of the raw rows / thresholds the NEXT step copies (row s+2): G segments of 2 TW bytes + 4 TW bytes
template <int G>
RIP_HD void prefetch_raw(const Args& A, int row, int tile, int tid, int lo, int hi) {
    if (tid <= G && in_range(row, imax(lo, 0), imin(hi, A.n))) {
        const int x0 = tile * TS;
        const unsigned ncol = (unsigned)imin(TW, A.n - x0);
        const size_t npl = (size_t)A.n * (size_t)A.n, o = (size_t)row * (size_t)A.n + (size_t)x0;
        if (tid < G) l2_prefetch_block(A.raw + (size_t)tid * npl + o, ncol * 2u);
        else l2_prefetch_block(A.thr + o, ncol * 4u);
    }
}

// IPC taps of stage b straight into R.kb (single buffer: b consumes them before the next load is issued)
template <int G, int P>
RIP_HD void load_b6(const Args& A, Regs<G, P>& R, int row, int tile, int tid) {
    const f4* p = A.recK + ((long)row * A.ntile + tile) * (KQ * TW);
    R.kb[0] = ld_rec<G>(p + tid);
    R.kb[1] = ld_rec<G>(p + TW + tid);
    R.kb8 = ld_rec1((const float*)(p + 2 * TW) + 4 * tid);
}

// Split-phase CTA barrier on an mbarrier (count = TW): arrive where the thread's contribution is published, wait where
// the others' is needed, independent work in between (v6 SCHED 3).  The host walk needs neither.
RIP_HD void sp_arrive(uint64_t* b) {
#if defined(__CUDA_ARCH__)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
#else
    (void)b;
#endif
}
RIP_HD void sp_wait(uint64_t* b, unsigned parity) {
#if defined(__CUDA_ARCH__)
    unsigned ok = 0, spins = 0;
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // (a lost arrival must be an error the host sees, not a hang)
    } while (!ok);
#else
    (void)b; (void)parity;
#endif
}
struct ArriveHook {
    uint64_t* b;
    RIP_HD void operator()() const {
        cp_async_wait<0>();  // this thread's part of row s+1 has landed: published by the arrival
        sp_arrive(b);
    }
};

// the same for float64 taps (words 0..17 of the K64 record) into R.kbd
template <int G, int P>
RIP_HD void load_b6_64(const Args& A, Regs<G, P>& R, int row, int tile, int tid) {
    const f4* p = A.recK + ((long)row * A.ntile + tile) * (KQ64 * TW);
    f4 w[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) w[q] = ld_rec<G>(p + q * TW + tid);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        R.kbd[2 * q] = f2_as_double(w[q].x, w[q].y);
        R.kbd[2 * q + 1] = f2_as_double(w[q].z, w[q].w);
    }
    R.kbd[8] = f2_as_double(w[4].x, w[4].y);
}

// first half of a march step (before the mid-step barrier): a1 (row s-2) and b (row s-4)
// SCHED (development A/B): 0 = a1 b [Lc] | c [L1 Lb] a0 ;  1 = a0 moved into the first half ;  2 = Lc issued before b ;
// 3 = 2 with SPLIT-PHASE barriers (mbar[0] mid-step, mbar[1] end of step; `it` = steps done, gives the wait parity):
//     wait-end(prev)  a1  [Lc]  b  arrive-mid  a0  wait-mid  c: stencil, arrive-end, ramp fit ...  [L1 Lb]
template <int G, int P, int SCHED = 0, bool K64 = false>
RIP_HD void step6a(const Args& A, Smem6<G, K64>& sm, Regs<G, P>& R, const int tid, const int tile, const int r0, const int r1, const int s) {
    using SM = Smem6<G, K64>;
    StepCtx C;
    make_ctx6<G>(C, A, tid, tile, r0, r1, s);
    const unsigned (&o5)[5] = C.o5;
    row_async<G, P>(A, sm, R, s + 1, 1, RIP_OS(1), tile, tid, r0 - 3, r1 + 3);
    prefetch_records<G, P, K64 ? KQ64 : KQ>(A, s, tile, tid, r0, r1);
    prefetch_raw<G>(A, s + 2, tile, tid, r0 - 3, r1 + 3);
    stage_a1<G, P>(A, sm, R, C);
    if (K64) {  // float64 taps: same schedule (record of stage c before stage b)
        load_c64<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
        stage_b64<G, P>(A, sm, R, C);
        return;
    }
    if (SCHED == 2) load_c<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
    stage_b<G, P>(A, sm, R, C);
    // record of stage c: issued only now (ptxas puts every global load of the loop on one scoreboard, so an earlier issue
    // would make stage b's first use of its taps wait for it); it flies while the warps gather at the barrier
    if (SCHED != 2) load_c<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
    if (SCHED == 1) stage_a0<G, P>(A, sm, C);
}

// second half (after the barrier that publishes O1 of row s-4): c (row s-5), the record of the next a1, a0 (row s)
template <int G, int P, int SCHED = 0, bool K64 = false>
RIP_HD void step6b(const Args& A, const RampPlanDev& pl, const FastTab& ft, Smem6<G, K64>& sm, Regs<G, P>& R, const int tid, const int tile,
                   const int r0, const int r1, const int s) {
    StepCtx C;
    make_ctx6<G>(C, A, tid, tile, r0, r1, s);
    if (K64) stage_c64<G, P>(A, pl, ft, sm, R, C);
    else stage_c<G, P>(A, pl, ft, sm, R, C);
    // records of the next step's a1 (row s-1) and b (row s-3): in flight during a0 and the barrier
    load_a1<G, P>(A, R, fold_row(s - 1, r0 - 2, r1 + 2), tile, tid);
    if (K64) load_b6_64<G, P>(A, R, fold_row(s - 3, r0 - 1, r1 + 1), tile, tid);
    else load_b6<G, P>(A, R, fold_row(s - 3, r0 - 1, r1 + 1), tile, tid);
    if (SCHED != 1) stage_a0<G, P>(A, sm, C);
    R.orow += (unsigned)A.n;
    cp_async_wait<0>();  // row s+1 has landed; the caller's barrier publishes it
}

template <int G, int P, bool K64 = false>
RIP_HD void prologue6(const Args& A, Smem6<G, K64>& sm, Regs<G, P>& R, int tid, int tile, int r0, int r1) {
    using SM = Smem6<G, K64>;
    const int s0 = r0 - 3;
    R.orow = (unsigned)(s0 * A.n);
    row_async<G, P>(A, sm, R, s0, 0, (unsigned)(s0 & 3) * (unsigned)SM::ROW5, tile, tid, r0 - 3, r1 + 3);
    load_a1<G, P>(A, R, fold_row(s0 - 2, r0 - 2, r1 + 2), tile, tid);
    if (K64) load_b6_64<G, P>(A, R, fold_row(s0 - 4, r0 - 1, r1 + 1), tile, tid);
    else load_b6<G, P>(A, R, fold_row(s0 - 4, r0 - 1, r1 + 1), tile, tid);
    // ring pads and the slots the stages read before anything was written there (never the cp.async targets)
    for (int k = 0; k < SM::DEPTH; ++k) {
        for (int i = tid; i < SM::H * RW; i += TW) sm.D((unsigned)k * SM::ROW5)[i] = f4{0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < RW; i += TW) sm.sat(k)[i] = 0u;
    }
    for (int i = tid; i < 3 * SM::ROWO / 16; i += TW) ((f4*)sm.ro)[i] = f4{0.f, 0.f, 0.f, 0.f};
    cp_async_wait<0>();
}

// =================================================================================================================
// v3: the same four stages, ROLE-SPLIT over two warp groups of one CTA and fed by the TMA engine.
//
//   * CTA = 256 threads on one 128-column tile.  Warps 0-3 (role X) run a0 + a1 (+ b when BX), warps 4-7 (role Y) run
//     (b +) c on the SAME columns; the shared-memory rings are the hand-off, one __syncthreads() per march step as in
//     v2 (every stage still reads only ring slots written in earlier steps).  Each role carries only its own staged
//     records (X: rec1, Y: recK), so both fit in 80 registers -> 3 CTAs x 8 warps = 24 warps/SM instead of 16, and
//     each warp's instruction footprint halves.
//   * The raw resultant rows and the saturation thresholds arrive by 1-D bulk copies (cp.async.bulk = UBLKCP, the TMA
//     engine) issued by ONE elected thread two rows ahead, completion through an mbarrier per ring slot
//     (expect_tx / try_wait.parity); the 128 LDGSTS per row of v2 and their address arithmetic are gone.
//   * Row / channel corrections (24 doubles per row) keep riding a per-thread cp.async group.
// Device only (the host check walks the v2 `step`, which calls the identical stage functions).
// =================================================================================================================
#if defined(__CUDACC__)
namespace tma {
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (bytes and both addresses multiples of 16), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                 "l"(src), "r"(bytes), "r"(s32(b))
                 : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(ok)
        : "r"(s32(b)), "r"(parity)
        : "memory");
    return ok != 0u;
}
// (try_wait suspends the thread in hardware for a bounded time; the spin count only guards against a lost copy:
// a trap is an error the host sees, a silent hang is not)
__device__ __forceinline__ void wait_parity(uint64_t* b, unsigned parity) {
    unsigned spins = 0;
    while (!try_wait(b, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}
}  // namespace tma

// shared memory of v3 = the v2 rings | 2 slots of IPC taps for stage b (words 0..2 of recK: 3 x TW float4) |
// RING + 2 mbarriers ("raw row landed", "taps landed")
constexpr int TAPS_BYTES = 3 * TW * 16;
template <int G>
RIP_HD constexpr size_t v3_off_taps() { return (Smem<G>::bytes() + 127) / 128 * 128; }
template <int G>
RIP_HD constexpr size_t v3_off_mbar() { return v3_off_taps<G>() + 2 * TAPS_BYTES; }
// (with BX only the first tap slot is used: the mbarriers then sit behind it)
template <int G>
RIP_HD constexpr size_t v3_smem_bytes(bool bx) { return v3_off_mbar<G>() - (bx ? TAPS_BYTES : 0) + 8 * (RING + 2); }
template <int G>
__device__ __forceinline__ uint64_t* v3_mbar(unsigned char* base, bool bx) { return (uint64_t*)(base + v3_off_mbar<G>() - (bx ? TAPS_BYTES : 0)); }
template <int G>
__device__ __forceinline__ f4* v3_taps(unsigned char* base, int slot) { return (f4*)(base + v3_off_taps<G>() + (size_t)slot * TAPS_BYTES); }

// Bulk copies of detector row `row` into ring slot (index k5, byte offset slot_o5), issued by the first warp of role X:
// lane 0 arms the mbarrier, lanes 0 .. G-1 copy one resultant each, lane G the thresholds (one instruction sequence
// for the warp instead of G + 1 in a single lane).  complete_tx before expect_tx is harmless: the phase cannot
// complete before the one pending arrival, which carries the expected byte count.
template <int G>
__device__ __forceinline__ void tma_row(const Args& A, Smem<G>& sm, uint64_t* mbar, int row, int k5, unsigned slot_o5, int tile, int lo,
                                        int hi, int lane) {
    if (!in_range(row, imax(lo, 0), imin(hi, A.n))) return;
    const int x0 = tile * TS;
    const unsigned ncol = (unsigned)imin(TW, A.n - x0);  // multiple of 8 (n % 8 == 0, TS % 8 == 0)
    const size_t npl = (size_t)A.n * (size_t)A.n, o = (size_t)row * (size_t)A.n + (size_t)x0;
    uint64_t* b = mbar + k5;
    if (lane == 0) tma::expect_tx(b, (unsigned)G * ncol * 2u + ncol * 4u);
    if (lane < G) tma::bulk_g2s(sm.raw(slot_o5) + lane * TW, A.raw + (size_t)lane * npl + o, ncol * 2u, b);
    else if (lane == G) tma::bulk_g2s(sm.thr(slot_o5), A.thr + o, ncol * 4u, b);
}

// row / channel corrections of `row` (threads 0 .. 3G-1 of role X, 8 bytes each), one cp.async group per step
template <int G>
__device__ __forceinline__ void corr_async(const Args& A, Smem<G>& sm, int row, unsigned slot_o5, int tile, int tid, int lo, int hi) {
    if (A.do_refpix && tid < 3 * G && in_range(row, imax(lo, 0), imin(hi, A.n))) {
        const int g = tid % G, which = tid / G;
        if (which == 0) {
            cp_async<8>(sm.rc(slot_o5) + g, A.rowcorr + ((unsigned)(g * A.n) + (unsigned)row));
        } else {
            int ch = ((tile * TS) >> 7) + (which - 1);
            if (ch > 31) ch = 31;
            cp_async<8>(sm.ln(slot_o5) + (which - 1) * G + g, A.chan_line + ((unsigned)((g * 32 + ch) * A.n) + (unsigned)row));
        }
    }
    cp_async_commit();
}

__device__ __forceinline__ void make_ctx(StepCtx& C, const Args& A, int tid, int tile, int r0, int r1, int s, unsigned o5s, unsigned RB,
                                         unsigned RING_B) {
    C.n = A.n; C.tid = tid; C.tile = tile; C.r0 = r0; C.r1 = r1; C.s = s;
    C.x = tile * TS + tid;
    C.col = tid + 1;
    C.xin = C.x < A.n;
    C.xact = in_range(C.x, 4, A.n - 4);
    C.o5[0] = o5s; C.o5[1] = wrap5(o5s + RB, RING_B); C.o5[2] = wrap5(o5s + 2 * RB, RING_B);
    C.o5[3] = wrap5(o5s + 3 * RB, RING_B); C.o5[4] = wrap5(o5s + 4 * RB, RING_B);
}

// rows for which stage b runs (IPC pass 1): the taps of exactly these rows are copied
template <int G>
__device__ __forceinline__ bool v3_b_row(int row, int r0, int r1, int n) { return in_range(row, imax(r0 - 1, 4), imin(r1 + 1, n - 4)); }
// taps of stage b (words 0..2 of the recK record of `row`) -> tap slot `slot`, by one thread
template <int G>
__device__ __forceinline__ void tma_taps(const Args& A, unsigned char* smem_raw, uint64_t* mbar, int row, int slot, int tile, int r0, int r1) {
    if (!v3_b_row<G>(row, r0, r1, A.n)) return;
    uint64_t* b = mbar + RING + slot;
    tma::expect_tx(b, (unsigned)TAPS_BYTES);
    tma::bulk_g2s(v3_taps<G>(smem_raw, slot), A.recK + ((long)row * A.ntile + tile) * (KQ * TW), (unsigned)TAPS_BYTES, b);
}

// v2 with the TMA-fed raw ring ("v2t"): one role, 128 threads, the march step of v2 -- only the raw rows and thresholds
// arrive by bulk copies (lanes of warp 0) instead of one LDGSTS pair per thread.
template <int G, int P>
__device__ __forceinline__ void v2t_body(const Args& A, const RampPlanDev& pl, const FastTab& ft, unsigned char* smem_raw) {
    constexpr unsigned RB = Smem<G>::ROW5, RING_B = RING * Smem<G>::ROW5;
    Smem<G> sm;
    sm.carve(smem_raw);
    uint64_t* mbar = (uint64_t*)(smem_raw + (size_t)RING * Smem<G>::ROW5 + (size_t)O_DEPTH * Smem<G>::ROW4);  // the 64 spare bytes of Smem::bytes()
    static_assert(8 * RING <= 64, "mbarriers must fit behind the rings");
    Regs<G, P> R;
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = imin(r0 + A.band_rows, A.n);
    const int x = tile * TS + tid;
    const bool xin = x < A.n;
    const int s0 = r0 - 3;
    unsigned o5s = first_o5<G>(r0);
    int k5 = mod_pos(s0, RING);
    unsigned nwait = 0;
    R.orow = (unsigned)(s0 * A.n);
    if (tid == 0) {
        for (int k = 0; k < RING; ++k) tma::mbar_init(mbar + k, 1u);
        tma::fence_mbar_init();
    }
    for (int k = 0; k < RING; ++k)
        for (int i = tid; i < Smem<G>::H * RW; i += TW) sm.D((unsigned)k * RB)[i] = f4{0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < O_DEPTH * Smem<G>::ROW4 / 16; i += TW) ((f4*)sm.r4)[i] = f4{0.f, 0.f, 0.f, 0.f};
    __syncthreads();
    if (tid < 32) {
        tma_row<G>(A, sm, mbar, s0, k5, o5s, tile, r0 - 3, r1 + 3, tid);
        tma_row<G>(A, sm, mbar, s0 + 1, (k5 + 1) % RING, wrap5(o5s + RB, RING_B), tile, r0 - 3, r1 + 3, tid);
    }
    corr_async<G>(A, sm, s0, o5s, tile, tid, r0 - 3, r1 + 3);
    corr_async<G>(A, sm, s0 + 1, wrap5(o5s + RB, RING_B), tile, tid, r0 - 3, r1 + 3);
    load_c<G, P>(A, R, s0 - 6, tile, tid, x, xin);
    load_a1<G, P>(A, R, s0 - 2, tile, tid);
    load_bn<G, P>(A, R, s0 - 4, tile, tid, 0u);
    cp_async_wait<0>();
    __syncthreads();
    for (int s = s0; s <= r1 + 5; ++s) {
        StepCtx C;
        make_ctx(C, A, tid, tile, r0, r1, s, o5s, RB, RING_B);
        const unsigned (&o5)[5] = C.o5;
        const unsigned dep = f_as_u(R.kc[0].x) & (unsigned)A.pad_;
        R.kb[0] = R.kbn[0]; R.kb[1] = R.kbn[1]; R.kb8 = R.kbn8;
        if (tid < 32) tma_row<G>(A, sm, mbar, s + 2, (k5 + 2) % RING, RIP_O5(2), tile, r0 - 3, r1 + 3, tid);
        corr_async<G>(A, sm, s + 2, RIP_O5(2), tile, tid, r0 - 3, r1 + 3);
        prefetch_records<G, P, KQ>(A, s, tile, tid, r0, r1);
        load_bn<G, P>(A, R, fold_row(s - 3, r0 - 1, r1 + 1), tile, tid, dep);
        stage_a1<G, P>(A, sm, R, C);
        stage_c<G, P>(A, pl, ft, sm, R, C);
        load_c<G, P>(A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin);
        load_a1<G, P>(A, R, fold_row(s - 1, r0 - 2, r1 + 2), tile, tid);
        stage_b<G, P>(A, sm, R, C);
        if (in_range(s, imax(r0 - 3, 0), imin(r1 + 3, A.n))) {
            tma::wait_parity(mbar + k5, (nwait / RING) & 1u);
            ++nwait;
        }
        stage_a0<G, P>(A, sm, C);
        R.orow += (unsigned)A.n;
        cp_async_wait<1>();
        o5s = next_o5<G>(o5s);
        k5 = (k5 == RING - 1) ? 0 : k5 + 1;
        __syncthreads();
    }
}

// role X: a0 (row s), a1 (row s-2) and, with BX, b (row s-4).  With BX the taps of stage b arrive in ONE slot, copied
// at the top of the step that uses them (stage a1 runs in between: about a third of a step).
template <int G, int P, bool BX>
__device__ __forceinline__ void v3_role_x(const Args& A, Smem<G>& sm, unsigned char* smem_raw, uint64_t* mbar, const int tid, const int tile,
                                          const int r0, const int r1) {
    constexpr unsigned RB = Smem<G>::ROW5, RING_B = RING * Smem<G>::ROW5;
    Regs<G, P> R;
    const int s0 = r0 - 3;
    unsigned o5s = first_o5<G>(r0);
    int k5 = mod_pos(s0, RING);  // ring slot index of row s
    unsigned nwait = 0;          // rows of this band waited for so far (they arrive in row order, slot after slot)
    unsigned nb_wait = 0;        // stage-b rows waited for so far
    load_a1<G, P>(A, R, s0 - 2, tile, tid);
    if (tid < 32) {
        tma_row<G>(A, sm, mbar, s0, k5, o5s, tile, r0 - 3, r1 + 3, tid);
        tma_row<G>(A, sm, mbar, s0 + 1, (k5 + 1) % RING, wrap5(o5s + RB, RING_B), tile, r0 - 3, r1 + 3, tid);
    }
    corr_async<G>(A, sm, s0, o5s, tile, tid, r0 - 3, r1 + 3);
    corr_async<G>(A, sm, s0 + 1, wrap5(o5s + RB, RING_B), tile, tid, r0 - 3, r1 + 3);
    cp_async_wait<0>();
    __syncthreads();  // (pairs with the prologue barrier of role Y)
    for (int s = s0; s <= r1 + 5; ++s) {
        StepCtx C;
        make_ctx(C, A, tid, tile, r0, r1, s, o5s, RB, RING_B);
        const unsigned (&o5)[5] = C.o5;
        if (tid < 32) {
            tma_row<G>(A, sm, mbar, s + 2, (k5 + 2) % RING, RIP_O5(2), tile, r0 - 3, r1 + 3, tid);
            if (BX && tid == G + 1) tma_taps<G>(A, smem_raw, mbar, s - 4, 0, tile, r0, r1);
        }
        corr_async<G>(A, sm, s + 2, RIP_O5(2), tile, tid, r0 - 3, r1 + 3);
        prefetch_records<G, P, KQ>(A, s, tile, tid, r0, r1);
        stage_a1<G, P>(A, sm, R, C);
        load_a1<G, P>(A, R, fold_row(s - 1, r0 - 2, r1 + 2), tile, tid);
        if (BX) {
            if (v3_b_row<G>(s - 4, r0, r1, A.n)) {
                tma::wait_parity(mbar + RING, nb_wait & 1u);
                ++nb_wait;
                const f4* kt = v3_taps<G>(smem_raw, 0);
                R.kb[0] = kt[tid]; R.kb[1] = kt[TW + tid]; R.kb8 = ((const float*)(kt + 2 * TW))[4 * tid];
            }
            stage_b<G, P>(A, sm, R, C);
        }
        if (in_range(s, imax(r0 - 3, 0), imin(r1 + 3, A.n))) {  // row s has a copy in flight (or landed): same predicate as tma_row
            tma::wait_parity(mbar + k5, (nwait / RING) & 1u);
            ++nwait;
        }
        stage_a0<G, P>(A, sm, C);
        cp_async_wait<1>();
        o5s = next_o5<G>(o5s);
        k5 = (k5 == RING - 1) ? 0 : k5 + 1;
        __syncthreads();
    }
}

// role Y: c (row s-6) and, without BX, b (row s-4; taps double-buffered in shared memory, copied one step ahead by the
// role's first thread).  The record of stage c is re-loaded for the next row from INSIDE stage c, right after its last
// use (see stage_c): the loads have two thirds of a step to land instead of the few instructions before the barrier.
template <int G, int P>
struct ReloadC {
    const Args& A;
    Regs<G, P>& R;
    int row, tile, tid, x;
    bool xin;
    __device__ __forceinline__ void operator()() const { load_c<G, P>(A, R, row, tile, tid, x, xin); }
};
template <int G, int P, bool BX>
__device__ __forceinline__ void v3_role_y(const Args& A, const RampPlanDev& pl, const FastTab& ft, Smem<G>& sm, unsigned char* smem_raw,
                                          uint64_t* mbar, const int tid, const int tile, const int r0, const int r1) {
    constexpr unsigned RB = Smem<G>::ROW5, RING_B = RING * Smem<G>::ROW5;
    Regs<G, P> R;
    const int s0 = r0 - 3;
    const int x = tile * TS + tid;
    const bool xin = x < A.n;
    unsigned o5s = first_o5<G>(r0);
    unsigned nb_wait = 0;  // stage-b rows waited for so far (consecutive rows alternate between the two tap slots)
    R.orow = (unsigned)(s0 * A.n);
    load_c<G, P>(A, R, s0 - 6, tile, tid, x, xin);
    if (!BX && tid == 0) tma_taps<G>(A, smem_raw, mbar, s0 - 4, (s0 - 4) & 1, tile, r0, r1);
    __syncthreads();
    for (int s = s0; s <= r1 + 5; ++s) {
        StepCtx C;
        make_ctx(C, A, tid, tile, r0, r1, s, o5s, RB, RING_B);
        if (!BX && tid == 0) tma_taps<G>(A, smem_raw, mbar, s - 3, (s - 3) & 1, tile, r0, r1);  // next step's row
        stage_c<G, P>(A, pl, ft, sm, R, C, ReloadC<G, P>{A, R, fold_row(s - 5, r0 - 1, r1 + 1), tile, tid, C.x, C.xin});
        if (!BX) {
            if (v3_b_row<G>(s - 4, r0, r1, A.n)) {
                tma::wait_parity(mbar + RING + ((s - 4) & 1), (nb_wait >> 1) & 1u);
                ++nb_wait;
                const f4* kt = v3_taps<G>(smem_raw, (s - 4) & 1);
                R.kb[0] = kt[tid]; R.kb[1] = kt[TW + tid]; R.kb8 = ((const float*)(kt + 2 * TW))[4 * tid];
            }
            stage_b<G, P>(A, sm, R, C);
        }
        R.orow += (unsigned)A.n;
        o5s = next_o5<G>(o5s);
        __syncthreads();
    }
}

// XR / YR: register budgets of the two roles (setmaxnreg; XR + YR <= 2 x the launch allocation), 0 = leave alone
template <int G, int P, bool BX, int XR, int YR>
__device__ __forceinline__ void v3_body(const Args& A, const RampPlanDev& pl, const FastTab& ft, unsigned char* smem_raw) {
    Smem<G> sm;
    sm.carve(smem_raw);
    uint64_t* mbar = v3_mbar<G>(smem_raw, BX);
    const int t = threadIdx.x, tile = blockIdx.x;
    const int r0 = blockIdx.y * A.band_rows;
    const int r1 = imin(r0 + A.band_rows, A.n);
    constexpr unsigned RB = Smem<G>::ROW5;
    // ring pads and the slots stage a1 / b read before anything was written there (never the bulk-copy targets)
    for (int k = 0; k < RING; ++k)
        for (int i = t; i < Smem<G>::H * RW; i += 2 * TW) sm.D((unsigned)k * RB)[i] = f4{0.f, 0.f, 0.f, 0.f};
    for (int i = t; i < O_DEPTH * Smem<G>::ROW4 / 16; i += 2 * TW) ((f4*)sm.r4)[i] = f4{0.f, 0.f, 0.f, 0.f};
    if (t == 0) {
        for (int k = 0; k < RING + 2; ++k) tma::mbar_init(mbar + k, 1u);
        tma::fence_mbar_init();
    }
    __syncthreads();
    if (t < TW) {
        if (XR > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XR));
        v3_role_x<G, P, BX>(A, sm, smem_raw, mbar, t, tile, r0, r1);
    } else {
        if (YR > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(YR));
        v3_role_y<G, P, BX>(A, pl, ft, sm, smem_raw, mbar, t - TW, tile, r0, r1);
    }
}
#endif  // __CUDACC__

// ---- packed calibration records (built once per CALDIR and group count) -------------------------------------
struct PackSrc {
    int n, nb, G, P;
    const float* dark;     // [>=G,n,n]
    const float* bias;     // [G,na,na] (group offset applied) or null
    const float* coefs;    // [P,n,n]
    const float* Smin;
    const float* Smax;
    const float* Sref;
    const float* gain;     // [n,n] f32
    const uint8_t* aux;    // [n,n]
    const float* ipc;      // [9,na,na] f32
    const float* read;
    const float* dslope;   // IPC-corrected dark slope
    const float* flat;     // get_flat product
    const uint32_t* sdq;   // merged static dq
};

RIP_HD float u_as_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// word w of rec1 at detector pixel (row, x); x >= n or a padding row -> 0
RIP_HD float rec1_word(const PackSrc& S, int row, int x, int w) {
    if (x >= S.n || row < 0 || row >= S.n) return 0.0f;
    const long npl = (long)S.n * S.n, p = (long)row * S.n + x;
    const int G = S.G, na = S.n - 2 * S.nb;
    if (w < G) return S.dark[(long)w * npl + p];
    if (w < 2 * G) {
        const int ya = row - S.nb, xa = x - S.nb;
        if (!S.bias || ya < 0 || ya >= na || xa < 0 || xa >= na) return 0.0f;
        return S.bias[((long)(w - G) * na + ya) * na + xa];
    }
    w -= 2 * G;
    if (w == 0) return S.Smin[p];
    if (w == 1) return S.Smax[p];
    if (w == 2) return S.Sref[p];
    if (w == 3) return S.gain[p];
    if (w == 4) return u_as_f((uint32_t)S.aux[p]);
    w -= 5;
    if (w < S.P) return S.coefs[(long)w * npl + p];
    return 0.0f;
}

// word w of recK: 0..8 gathered IPC taps K[1+dy][1+dx][y-dy][x-dx] in the reference's accumulation order (zero when
// the source pixel is outside the active area or (row, x) is not active), 9 gain, 10 read, 11 dark slope (IPC
// corrected), 12 flat, 13 static dq bits, 14..15 zero.
RIP_HD float recK_word(const PackSrc& S, int row, int x, int w) {
    if (x >= S.n || row < 0 || row >= S.n) return 0.0f;
    const long p = (long)row * S.n + x;
    const int na = S.n - 2 * S.nb;
    if (w < 9) {
        const int DY[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1};
        const int DX[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
        const int ya = row - S.nb, xa = x - S.nb;
        if (ya < 0 || ya >= na || xa < 0 || xa >= na) return 0.0f;
        const int ys = ya - DY[w], xs = xa - DX[w];
        if (ys < 0 || ys >= na || xs < 0 || xs >= na) return 0.0f;
        return S.ipc[((long)((1 + DY[w]) * 3 + (1 + DX[w])) * na + ys) * na + xs];
    }
    if (w == 9) return S.gain[p];
    if (w == 10) return S.read[p];
    if (w == 11) return S.dslope[p];
    if (w == 12) return S.flat[p];
    if (w == 13) return u_as_f(S.sdq[p]);
    return 0.0f;
}

// word w of the float64-tap record (K64): 0..17 the 9 gathered taps as doubles (low word first), 18 gain, 19 read,
// 20 dark slope (IPC corrected), 21 flat, 22 static dq bits, 23 zero.  S.ipc then points at float64 data.
RIP_HD float recK64_word(const PackSrc& S, int row, int x, int w) {
    if (x >= S.n || row < 0 || row >= S.n) return 0.0f;
    const long p = (long)row * S.n + x;
    const int na = S.n - 2 * S.nb;
    if (w < 18) {
        const int t = w >> 1;
        const int DY[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1};
        const int DX[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
        const int ya = row - S.nb, xa = x - S.nb;
        double v = 0.0;
        if (ya >= 0 && ya < na && xa >= 0 && xa < na) {
            const int ys = ya - DY[t], xs = xa - DX[t];
            if (ys >= 0 && ys < na && xs >= 0 && xs < na)
                v = ((const double*)S.ipc)[((long)((1 + DY[t]) * 3 + (1 + DX[t])) * na + ys) * na + xs];
        }
        float parts[2];
        memcpy(parts, &v, 8);
        return parts[w & 1];
    }
    if (w == 18) return S.gain[p];
    if (w == 19) return S.read[p];
    if (w == 20) return S.dslope[p];
    if (w == 21) return S.flat[p];
    if (w == 22) return u_as_f(S.sdq[p]);
    return 0.0f;
}

}  // namespace v2
}  // namespace rip
