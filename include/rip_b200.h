/* rip_b200.h -- C ABI of the B200 (sm_100a) implementation of romanimpreprocess's per-pixel hot path.
 *
 * The reference (Roman-HLIS-Cosmology-PIT/romanimpreprocess) has no FFI: the path sits behind plain Python
 * functions on NumPy arrays.  Each entry point below replaces one of those functions and cites it
 * (paths relative to the reference checkout).  The Python modules of `romanimpreprocess_b200/` bind these with
 * ctypes and keep the reference's names, argument meaning, in-place semantics and exceptions; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; rip_last_error() gives the message (thread local)
 *   - no exceptions cross the boundary, the caller owns every buffer passed in
 *   - "host" entry points take host pointers and are synchronous (copies inside);
 *     "_dev" entry points take device pointers + a cudaStream_t (as void*) and are asynchronous
 *   - frames are row-major [ny][nx]; cubes [ngroup][ny][nx]; ipc4d kernels [3][3][nya][nxa]
 *   - dtype tags: the precision of the reference's arithmetic depends on the dtype of each calibration plane
 *     (SURVEY App. A0), so gain / ipc4d / area planes carry a tag
 */
#ifndef RIP_B200_H
#define RIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIP_ABI_VERSION 1
#define RIP_GMAX 16      /* max resultants (groups) per ramp on the fused path            */
#define RIP_MAXVAR 16    /* full fit + truncated-fit variants in a ramp plan              */
#define RIP_MAXSLICE 256 /* total single/double differences over all variants             */
#define RIP_PMAX 16      /* max Legendre coefficients (order+1)                           */

enum rip_dtype { RIP_F32 = 0, RIP_F64 = 1, RIP_I32 = 2, RIP_U16 = 3 };

/* ---- ramp plan: every scalar of utils/fitting.py:jump_detect that does not depend on the pixel ------------
 * Built by the host with the reference's own NumPy expressions (romanimpreprocess_b200/utils/fitting.py
 * build_plan) so each value carries the reference's rounding.  Variant 0 = full ramp; variant v>=1 = ramp
 * truncated at iend = G - v with two-point weights (reference utils/fitting.py:165-169, 326). */
typedef struct rip_ramp_slice {
    int32_t i, di;  /* difference data[i+di]-data[i]                    (fitting.py:225-231) */
    float dt;       /* tbar[i+di]-tbar[i] as float32                                          */
    float inv_dt;   /* fast path only                                                         */
    float A, B;     /* fast path: var(delta slope) ~= dvardt*A + read^2*B     */
} rip_ramp_slice;

typedef struct rip_ramp_plan {
    int32_t G, start, nvar, reserved;
    float tbar[RIP_GMAX], tau[RIP_GMAX], nreads[RIP_GMAX];
    int32_t var_ngrp[RIP_MAXVAR];
    float var_K[RIP_MAXVAR][RIP_GMAX]; /* weights per variant (fitting.py:162-169)                    */
    float var_coef[RIP_MAXVAR];        /* Poisson variance coefficient, f32 accumulation (:196-200)   */
    float var_rfac[RIP_MAXVAR];        /* sqrt(sum K^2/N) as float32 (:209)                           */
    int32_t var_slice_off[RIP_MAXVAR + 1];
    rip_ramp_slice slices[RIP_MAXSLICE];
    float IthreshA_f, IthreshB_f;
    double SthreshA, SthreshB, logIratio; /* (:172-184, 215-217)                                      */
    float band;                           /* relative half-width of the exact-recheck band            */
    /* fast path only: thr ~= thrA_f + thrK_f * log(clip(slope) * invIA_f),  thrK = (SthreshB-SthreshA)/log(IB/IA) */
    float thrA_f, thrK_f, invIA_f;
} rip_ramp_plan;

/* ---- one SCA's calibration reference data (CALDIR), host pointers ---------------------------*/
typedef struct rip_caldir_desc {
    int32_t n;          /* frame side incl. reference pixels (4096)                         */
    int32_t nb;         /* reference-pixel border (4)                                       */
    int32_t P;          /* Legendre coefficients = order+1                                  */
    int32_t n_dark;     /* groups in the dark cube                                          */
    int32_t n_bias;     /* groups in the biascorr cube, 0 = no biascorr                     */
    int32_t gain_dtype; /* RIP_F32 | RIP_F64                                                */
    int32_t ipc_dtype;  /* RIP_F32 | RIP_F64                                                */
    int32_t has_amp33;  /* read file carries amp33 statistics                               */
    const float* lin_coefs; /* [P,n,n]   linearitylegendre.data                             */
    const float* Smin;      /* [n,n]                                                        */
    const float* Smax;
    const float* Sref;
    const uint32_t* lin_dq;     /* [n,n]                                                    */
    const uint32_t* mask_dq;    /* [n,n] or NULL                                            */
    const float* sat_thresh;    /* [n,n]   saturation.data                                  */
    const uint32_t* sat_dq;     /* [n,n]                                                    */
    const void* gain;           /* [n,n]   gain.data (f32|f64)                              */
    const void* ipc;            /* [3,3,n-2nb,n-2nb] ipc4d.data (f32|f64) or NULL           */
    const float* read;          /* [n,n]   read.data                                        */
    const float* resetnoise;    /* [n,n]   read.resetnoise (forward model) or NULL          */
    const float* dark_cube;     /* [n_dark,n,n] dark.data                                   */
    const float* dark_slope;    /* [n,n]   dark.dark_slope                                  */
    const uint32_t* dark_dq;    /* [n,n]                                                    */
    const float* biascorr;      /* [n_bias,n-2nb,n-2nb] or NULL                             */
    const float* flat;          /* [n,n]   flat (pflat).data                                */
    const float* amp33_med;     /* [n,128] or NULL                                          */
    const float* amp33_std;     /* [n,128] or NULL                                          */
    double biascorr_t0, m_pink, ru_pink, c_pink, u_pink;
    double refout_slope; /* optimal reference-output coefficient (gen_cal_image.py:542-553), computed by the
                            host with the reference's expression; NaN = derive it inside the library          */
} rip_caldir_desc;

typedef struct rip_caldir rip_caldir; /* opaque: device-resident planes + static products of one SCA */

/* ---- parameters / outputs of the fused L1->L2 call ---------------------------------------------------------*/
typedef struct rip_l1l2_params {
    int32_t G;                  /* resultants in the L1 cube                                          */
    int32_t exclude_first;      /* EXCLUDE_FIRST (gen_cal_image.py:142,434)                           */
    int32_t sat_backup;         /* SATURATION_BACKUP (gen_cal_image.py:504)                           */
    int32_t do_not_flag_first;  /* read_pattern[0]==[0] (gen_cal_image.py:583)                        */
    int32_t do_refpix;          /* run the reference-pixel loop (needs amp33; gen_cal_image.py:531)   */
    int32_t area_dtype;         /* RIP_F32 | RIP_F64 for the AreaFactor plane (may be NULL -> 1)      */
    int32_t threads;            /* 0 = default tile width                                             */
    int32_t band_rows;          /* 0 = default rows per tile                                          */
} rip_l1l2_params;

typedef struct rip_l2_out {
    float* slope;        /* [n,n] DN/s, flat-fielded (gen_cal_image.py:627)         */
    float* err_read;     /* [n,n]                     (:613,628)                    */
    float* err_poisson;  /* [n,n]                     (:612,629)                    */
    uint32_t* pdq;       /* [n,n]                                                   */
    int8_t* endslice;    /* [n-2nb,n-2nb] or NULL     (:697-709)                    */
    uint8_t* rdq;        /* [G,n,n] or NULL  group dq after ramp_fit                */
    float* lin_cube;     /* [G,n,n] or NULL  data after multilin + correct_cube     */
} rip_l2_out;

/* ---- library ------------------------------------------------------------------------------------------------*/
const char* rip_last_error(void);
int rip_abi_version(void);
int rip_device_count(int* count);
int rip_device_sync(int device);
int rip_host_alloc(void** p, size_t bytes); /* pinned host memory for full-speed async copies */
int rip_host_free(void* p);
int rip_dev_alloc(int device, void** p, size_t bytes);
int rip_dev_free(int device, void* p);
int rip_copy_h2d(int device, void* dst_dev, const void* src_host, size_t bytes, void* stream);
int rip_copy_d2h(int device, void* dst_host, const void* src_dev, size_t bytes, void* stream);
int rip_stream_sync(int device, void* stream);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long rip_launch_count(void);
/* sizeof() of the ABI structs as compiled (binding self-check): 0 rip_ramp_slice, 1 rip_ramp_plan, 2 rip_caldir_desc,
 * 3 rip_l1l2_params, 4 rip_l2_out, 5 rip_fwd_params; -1 otherwise */
long rip_struct_size(int which);

/* ---- stage entry points: one per reference function, HOST pointers, synchronous ----------------------------*/

/* _lin (utils/ipc_linearity.py:192): phi = sum_L c_L P_L(z) (+ linear extrapolation), exflag = |z|>1.
 * z f32|f64 [npix]; coefs f32 [P,npix]; phi f32 [npix]. */
int rip_lin_eval(int device, const void* z, int z_dtype, const float* coefs, int P, long npix, int linextrap,
                 float* phi, uint8_t* exflag);

/* multilin (utils/ipc_linearity.py:276) and, with single_frame=1, linearity (:234).
 * S f32 [G,npix]; attempt u8 [G,npix] or NULL (=all true); phi f32 [G,npix]; dq_out u32 [npix]. */
int rip_multilin(int device, const float* S, int G, long npix, const float* coefs, int P, const float* Smin,
                 const float* Smax, const float* Sref, const uint32_t* lin_dq, const uint8_t* attempt,
                 int do_not_flag_first, int single_frame, float* phi, uint32_t* dq_out);

/* invlinearity (utils/ipc_linearity.py:347): 24-step bisection.  Slin, S_out f32|f64 [npix]. */
int rip_invlinearity(int device, const void* Slin, int dtype, long npix, const float* coefs, int P,
                     const float* Smin, const float* Smax, void* S_out, uint8_t* exflag);

/* ipc_fwd (utils/ipc_linearity.py:37) / ipc_rev (:102).  image [ny,nx] img_dtype; kernel [3,3,ny,nx] k_dtype;
 * gain [ny,nx] g_dtype or NULL.  out has NumPy's promoted dtype (f64 if any input plane is f64), reported in
 * *out_dtype; out must hold ny*nx doubles in the worst case. */
int rip_ipc_fwd(int device, const void* image, int img_dtype, int ny, int nx, const void* kernel, int k_dtype,
                const void* gain, int g_dtype, void* out, int* out_dtype);
int rip_ipc_rev(int device, const void* image, int img_dtype, int ny, int nx, const void* kernel, int k_dtype,
                int order, const void* gain, int g_dtype, void* out, int* out_dtype);

/* correct_cube (utils/ipc_linearity.py:145): in-place on the active region of data f32 [G,ny,nx];
 * kernel [3,3,ny-2nb,nx-2nb]; gain_full [ny,nx] or NULL. */
int rip_correct_cube(int device, float* data, int G, int ny, int nx, const void* kernel, int k_dtype, int nya,
                     int nxa, const void* gain_full, int g_dtype);

/* jump_detect (utils/fitting.py:89) for plan variant `variant` (0 = truncate_ramp None).
 * data f32 [G,ny,nx]; rdq u8 [G,ny,nx] in/out; smap f32 [nslices,ny,nx] or NULL. */
int rip_jump_detect(int device, const float* data, uint8_t* rdq, int ny, int nx, int nb, const rip_ramp_plan* plan,
                    const double* w_exact, int variant, const void* gain, int g_dtype, const float* read,
                    float* slope, float* err_read, float* err_poisson, float* smap);

/* ramp_fit (utils/fitting.py:258).  rdq u8 [G,ny,nx] and pdq u32 [ny,nx] in/out.  fast=1 uses the factorised
 * variance with exact re-evaluation of borderline slices (same flags), fast=0 the exact op order throughout. */
int rip_ramp_fit(int device, const float* data, uint8_t* rdq, uint32_t* pdq, int ny, int nx, int nb,
                 const rip_ramp_plan* plan, const double* w_exact, const void* gain, int g_dtype,
                 const float* read, int fast, float* slope, float* err_read, float* err_poisson);

/* ref_subtraction_row (utils/reference_subtraction.py:77) in two calls (np.polyfit stays on the host):
 * medians: image f32 [n, ncols]; ref_med[n] = per-row median of the reference output (use_ref_channel) or of the
 * 8 side reference pixels; sci_med[n] (optional) = per-row median of columns 4..n-5. */
int rip_row_medians(int device, const float* image, int n, int ncols, int use_ref_channel, float* ref_med,
                    float* sci_med);
/* apply: image[i,:] = f32(f64(image[i,:]) - m_med*(ref_med[i]-ctr)), ctr = median(ref_med) (:115-123) */
int rip_refsub_row_apply(int device, float* image, int n, int ncols, double m_med, const float* ref_med);

/* ref_subtraction_channel (utils/reference_subtraction.py:16): nchan channels of 128 columns. */
int rip_refsub_channel(int device, float* image, int n, int ncols, int nchan);

/* get_flat (utils/flatutils.py:20).  pdq u32 [n,n] in/out or NULL; out f32 [n,n]. */
int rip_get_flat(int device, const float* flat, int n, int nb, const void* gain, int g_dtype, const void* kernel,
                 int k_dtype, uint32_t* pdq, int ipc_deconvolve, float* out);

/* saturation_check -> romancal flag_saturation (L1_to_L2/gen_cal_image.py:148-185; restated).
 * raw u16 [G,n,n]; rdq u8 [G,n,n] in/out; pdq u32 [n,n] in/out. */
int rip_flag_saturation(int device, const uint16_t* raw, int G, int n, const float* sat_thresh,
                        const uint32_t* sat_dq, int backup, int skip_firstn, uint8_t* rdq, uint32_t* pdq);

/* ---- CALDIR handle + fused path ------------------------------------------------------------------------------*/
int rip_caldir_create(int device, const rip_caldir_desc* desc, rip_caldir** out);
void rip_caldir_destroy(rip_caldir* h);
/* static products: IPC-corrected dark slope (gen_cal_image.py:217-221), get_flat output and its
 * flags, merged static dq; copied to host buffers (any may be NULL) for inspection/tests. */
int rip_caldir_get_static(rip_caldir* h, float* dark_slope_ipc, float* flat_ipc, uint32_t* static_dq, double* refout_slope);

/* calibrateimage numerics, L1 arrays -> L2 arrays (L1_to_L2/gen_cal_image.py:503-629, 697-709).
 * raw u16 [G,n,n]; amp33 u16 [G,n,128] (NULL iff !do_refpix); area [n,n] or NULL. */
int rip_l1_to_l2_host(rip_caldir* h, const uint16_t* raw, const uint16_t* amp33, const void* area,
                      const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                      const rip_l2_out* out);
int rip_l1_to_l2_dev(rip_caldir* h, const uint16_t* d_raw, const uint16_t* d_amp33, const void* d_area,
                     const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                     const rip_l2_out* d_out, void* stream);
/* Look-ahead of the reference-pixel statistics (the "K0" kernels: utils/reference_subtraction_util.py:14-136 restated)
 * for the NEXT exposure of a device-resident stream: computed on a low-priority side stream of the handle into the
 * second of two workspace sets while the fused kernel of the current exposure runs.  If the rip_l1_to_l2_dev call that
 * FOLLOWS names the same d_raw, it waits for that result instead of running the statistics itself (identical results:
 * same kernels, same inputs); any other call in between discards the look-ahead.  Call it AFTER the rip_l1_to_l2_dev of
 * the current exposure.  The cubes must be complete in device memory when this is called (the call is not ordered
 * against any user stream) and must not change until that rip_l1_to_l2_dev call has been issued.
 * Optional: without it rip_l1_to_l2_dev computes the statistics in its own stream, as before. */
int rip_caldir_prefetch_refpix(rip_caldir* h, const uint16_t* d_raw_next, const uint16_t* d_amp33_next, int G);
/* Pipelined host entry for a stream of exposures of one SCA: `depth` exposures in flight on three CUDA streams
 * (H2D | reference-pixel statistics + fused kernel | D2H), so PCIe copies overlap the kernels.  submit() returns
 * immediately (it blocks only when all slots are busy) and hands back a ticket; the host output buffers named in
 * `out` are valid after wait(ticket).  Host buffers should be page-locked (rip_host_alloc) for real overlap; the
 * input buffers must stay untouched until wait() returns.  Same arithmetic as rip_l1_to_l2_host. */
typedef struct rip_pipeline rip_pipeline;
int rip_pipeline_create(rip_caldir* h, int G, int depth, int want_endslice, int want_rdq, rip_pipeline** out);
void rip_pipeline_destroy(rip_pipeline* p);
int rip_pipeline_submit(rip_pipeline* p, const uint16_t* raw, const uint16_t* amp33, const void* area,
                        const rip_l1l2_params* prm, const rip_ramp_plan* plan, const double* w_exact,
                        const rip_l2_out* out, long* ticket);
int rip_pipeline_wait(rip_pipeline* p, long ticket);
/* Keep one AreaFactor plane (gen_cal_image.py:618-622) resident on the device: exposures submitted with area = NULL
 * use it instead of "no area division", so a stream of exposures that share the plane uploads it once instead of
 * 67 MB per exposure.  area = NULL drops the resident plane.  Waits for the exposures in flight. */
int rip_pipeline_set_area(rip_pipeline* p, const void* area, int area_dtype);

/* Per-launch device timing of the fused kernel (bench.py's roofline): while enabled, every rip_l1_to_l2_* call on
 * this handle brackets the fused kernel with CUDA events on its launch stream; rip_profile_fetch waits for them,
 * returns the summed kernel time [ms] and the number of launches, and resets the counter. */
int rip_profile_enable(rip_caldir* h, int on);
int rip_profile_fetch(rip_caldir* h, double* fused_ms_sum, int* launches);
/* the K0 statistics alone (tests): rowcorr f64 [G,n], chan_m/chan_c f64 [G,32] to host */
int rip_refpix_stats_host(rip_caldir* h, const uint16_t* raw, const uint16_t* amp33, int G, double* rowcorr,
                          double* chan_m, double* chan_c, float* gmed);

/* ---- forward model (from_sim/sim_to_isim.py) ------------------------------------------------------------------*/

/* IL.apply (utils/ipc_linearity.py:461-513) on explicit planes of a window [ny,nx] (HOST pointers): counts
 * (i32|f32|f64) + start_e (f32|f64 plane, or NULL -> start_scalar) -> ipc_fwd (kernel [3,3,ny,nx] or NULL) ->
 * /gain (if electrons) -> 24-step bisection; with electrons_out: gain*(S-Sref).  Arithmetic follows NumPy's dtype
 * promotion (int32+f32 -> f64: SURVEY App. A10); out receives f64 values, *out_dtype the NumPy result dtype. */
int rip_il_apply_planes(int device, const void* counts, int c_dtype, int ny, int nx, const void* start_e, int s_dtype,
                        double start_scalar, const void* kernel, int k_dtype, const void* gain, int g_dtype,
                        const float* coefs, int P, const float* Smin, const float* Smax, const float* Sref,
                        int electrons, int electrons_out, double* out, int* out_dtype);

/* the same on the resident planes of a CALDIR handle: counts [na,na], start_e f32 [na,na] or NULL; out f64. */
int rip_il_apply(rip_caldir* h, const void* counts, int c_dtype, const float* start_e, int electrons,
                 int electrons_out, double* out);

typedef struct rip_fwd_params {
    int32_t G;
    int32_t n_reads;                   /* total reads listed in the pattern                              */
    int32_t reads_per_group[RIP_GMAX];
    int32_t read_index[64];            /* flattened read indices (time = read_time*index)                */
    double read_time;
    uint64_t seed;
    int32_t add_read_noise, add_reset_noise, add_biascorr, quantize;
    /* cosmic rays (romanisim.cr.simulate_crs, run per read by romanisim's apportioning when crparam is not None; the
     * reference passes crparam={} = these defaults, from_sim/sim_to_isim.py:233-242).  0 = off. */
    int32_t cr_enable;
    int32_t pad_;
    double cr_flux;                    /* events / cm^2 / s                                  (8)      */
    double cr_area;                    /* detector area [cm^2]                               (16.8)   */
    double cr_conversion_factor;       /* eV per electron                                    (0.5)    */
    double cr_pixel_size;              /* [um]                                               (10)     */
    double cr_pixel_depth;             /* [um]                                               (5)      */
} rip_fwd_params;

/* make_l1_fullcal (from_sim/sim_to_isim.py:163-262) with romanisim's apportioning / read noise restated
 * (SURVEY App. D): counts i32 [na,na] total electrons of the exposure -> resultants f32 [G,na,na] (DN).
 * Reset noise, binomial apportioning of the counts to the reads, IL.apply per read (float64), group mean,
 * read noise, + biascorr, round.  Counter-based Philox RNG (seed, pixel, purpose).
 * cum_counts i32 [n_reads,na,na] (host entry only, tests): externally apportioned cumulative counts, replaces
 * the binomial draws so the deterministic arithmetic can be checked exactly. */
int rip_make_l1_host(rip_caldir* h, const int32_t* counts, const int32_t* cum_counts, const rip_fwd_params* prm,
                     float* resultants);
int rip_make_l1_dev(rip_caldir* h, const int32_t* d_counts, const rip_fwd_params* prm, float* d_resultants,
                    void* stream);
/* Cosmic-ray group bits of the last rip_make_l1_* call of the handle that ran with cr_enable: u32 [na,na], bit g set
 * = a cosmic ray deposited electrons during group g (the per-group JUMP_DET flags of romanisim's dq cube, which
 * make_l1_fullcal returns at from_sim/sim_to_isim.py:262).  _dev returns the handle's own device buffer. */
int rip_fwd_cr_groups_host(rip_caldir* h, uint32_t* groups);
int rip_fwd_cr_groups_dev(rip_caldir* h, const uint32_t** d_groups);
/* Cumulative electrons per read i32 [n_reads,na,na] of the handle's last forward ramp (apportioning + cosmic rays), for
 * the statistical tests of the samplers (only when the ramp drew them itself or ran with cr_enable). */
int rip_fwd_cum_counts_host(rip_caldir* h, int n_reads, int32_t* cum);

/* ---- forward path either side of make_l1_fullcal (a15, a20) ----------------------------------------
 * Image2D.simulate (from_sim/sim_to_isim.py:615-648).  rip_sim_calprep returns the two calibration planes of the
 * scene on the active array [na,na]: this_dark = clip(ipc_rev(dark_slope*gain), -0.1*this_flat, .) in e/s and
 * this_flat = clip(ipc_rev(flat, gain=g), 0, 2-2^-21) (exact against the reference functions for f32 planes).
 * rip_sim_counts_*: counts (+)= Poisson(clip(cnorm*t_exp*g/g_ideal * image * this_flat/area, 0)) and, when
 * t_dark > 0, + Poisson(this_dark*t_dark) (romanisim's dark term, restated: SURVEY App. D).  image f32 [na,na]
 * (e/s per ideal pixel), area [na,na] = pixel area / Omega_ideal (f32|f64) or NULL; counts i32 [na,na];
 * rate_out f64 [na,na] or NULL receives the Poisson mean (for the statistical tests). */
int rip_sim_calprep(rip_caldir* h, float* this_dark, float* this_flat);
int rip_sim_counts_dev(rip_caldir* h, const float* d_image, const void* d_area, int area_dtype, double t_exp,
                       double cnorm, double g_ideal, double t_dark, uint64_t seed, int32_t* d_counts, int accumulate,
                       void* stream);
int rip_sim_counts_host(rip_caldir* h, const float* image, const void* area, int area_dtype, double t_exp, double cnorm,
                        double g_ideal, double t_dark, uint64_t seed, int32_t* counts, int accumulate, double* rate_out);

/* noise_1f_frame (from_sim/sim_to_isim.py:265-303): nframes blocks [nside, nside/32] f32 of 1/f noise (nside a power
 * of two, 32..4096).  draws f64 [nframes, 2*m] (m = 2*nside*nside/32): the N(0,1) stream of the reference call (real
 * parts, then imaginary parts) -- given, the result is deterministic (test path); NULL = Philox(seed, frame). */
int rip_noise_1f_frames_host(int device, int nside, int nframes, uint64_t seed, const double* draws, float* out);

/* fill_in_refdata_and_1f (from_sim/sim_to_isim.py:306-402), in place on the u16 cube [G,n,n]: reference pixels =
 * round(N*read/sqrt(N_g) + N*resetnoise + dark cube), active pixels kept, + 1/f banding (one common and 32 channel
 * frames per group, odd channels mirrored) on every pixel, clip to u16; amp33 u16 [G,n,n/32] (or NULL) = reference
 * output from the read file's amp33 statistics. */
int rip_fill_refdata_1f_dev(rip_caldir* h, uint16_t* d_im, uint16_t* d_amp33, int G, const int32_t* reads_per_group,
                            uint64_t seed, int fill_in_banding, void* stream);
int rip_fill_refdata_1f_host(rip_caldir* h, uint16_t* im, uint16_t* amp33, int G, const int32_t* reads_per_group,
                             uint64_t seed, int fill_in_banding);

/* ---- many-realisations bookkeeping (validation_tests/many_realizations.py:58-106; BASELINE configs[4]) ---------
 * rip_mask_build_host: CombinedMask.build (utils/maskhandling.py:82-117): grow32[b] in {0,1,5,9,25} = pixels affected
 * by bit b (1 = itself, 5 = cross, 9 = 3x3, 25 = 5x5; zero padding at the edges); mask u8 [ny,nx], 1 = masked.
 * rip_moments_accumulate_dev: on the active window of full-frame slope f32 / pdq u32 [n,n]:
 * moments[0] += w, [1] += w*data, [2] += w*data^2 with w = !mask (f32 accumulators [3,na,na], realisation order).
 * rip_moments_finalize_dev: mean, std, -1000 sentinel (:80-83).  rip_stack_median_dev: np.median over R <= 128 planes. */
int rip_mask_build_host(int device, const uint32_t* dq, int ny, int nx, const uint8_t* grow32, uint8_t* mask);
int rip_moments_accumulate_dev(int device, const float* d_slope, const uint32_t* d_pdq, int n, int nb,
                               const uint8_t* grow32, float* d_moments, void* stream);
int rip_moments_finalize_dev(int device, float* d_moments, long npix, void* stream);
int rip_stack_median_dev(int device, const float* d_stack, int R, long npix, float* d_out, void* stream);

/* glue of one realisation: resultants f32 [G,na,na] -> active window of the u16 L1 cube [G,n,n] (romanisim make_asdf
 * restated: clip to 0..65535, zero border), and the per-realisation planes of many_realizations.py:69-73
 * (diffs = last - second group, images = slope, err = hypot(err_read, err_poisson); full frames, zero border). */
int rip_l1_embed_dev(int device, const float* d_resultants, int G, int n, int nb, uint16_t* d_im, void* stream);
int rip_realization_record_dev(int device, const uint16_t* d_im, int G, int n, int nb, const float* d_slope,
                               const float* d_err_read, const float* d_err_poisson, float* d_diffs, float* d_images,
                               float* d_err, void* stream);

/* ---- sky model (utils/sky.py:98-190 medfit; called at gen_cal_image.py:645-647 and by every
 * production noise layer) ---------------------------------------------------------------------------------------
 * rip_block_nanmedian_dev / rip_medfit_host: np.nanmedian of the N x N regions of ky x kx = (ny/N) x (nx/N) pixels
 * starting at ((ny%N)/2, (nx%N)/2) -> meds f32 [N,N] (exact order statistics by radix select; float32 mean of the
 * two middle values for even counts; NaN for an all-NaN region).  The (order+1)(order+2)/2-coefficient normal
 * equations are solved by the caller in float64 (Python mirror: the reference's own NumPy lines).
 * rip_medfit_eval_dev: model[y,x] = float32(sum_k coef[k] * (LPY[j,y] * LPX[i,x])) in the reference's term order
 * (coef, LPX [order+1,nx], LPY [order+1,ny]: host float64); written to d_model (or NULL) and/or subtracted in place
 * from d_arr (row pitch in elements; or NULL). */
int rip_block_nanmedian_dev(int device, const float* d_arr, long pitch, int ny, int nx, int N, float* d_meds, void* stream);
/* The pixel-independent part of medfit (utils/sky.py:137-175), host float64: from meds f32 [N,N] (NaN = region without
 * a finite median) the (order+1)(order+2)/2 coefficients in the reference's ordering, and (if not NULL) the Legendre
 * polynomials on the pixel grid LPX [order+1,nx], LPY [order+1,ny] for rip_medfit_eval_dev.  Legendre values follow
 * scipy.special.legendre_p bit for bit; the normal equations are accumulated in the reference's order and solved by LU
 * with partial pivoting (LAPACK's result agrees to a few ulp of float64: the order of its updates is not defined). */
int rip_medfit_solve(int ny, int nx, int N, int order, const float* meds, double* coef, double* LPX, double* LPY);
/* smooth_mode (utils/sky.py:46-93) as called at gen_cal_image.py:641: rip_bin_masked_dev = binkxk(where(~mask, arr,
 * nan), k) on device planes (mask u8 [ny,nx] or NULL; out f32 [ny/k, nx/k]); rip_gauss_hist_dev = the smoothed
 * histogram sums[j] = sum_i exp(-0.5 ((z[j] - a_i) / width)^2) over the non-NaN elements, float64, nz <= 32. */
int rip_bin_masked_dev(int device, const float* d_arr, const uint8_t* d_mask, int ny, int nx, int k, float* d_out, void* stream);
int rip_gauss_hist_dev(int device, const float* d_arr, long count, const double* z, int nz, double width, double* sums,
                       void* stream);
int rip_medfit_host(int device, const float* arr, int ny, int nx, int N, float* meds);
int rip_medfit_eval_dev(int device, int ny, int nx, int order, const double* coef, const double* LPX, const double* LPY,
                        float* d_model, float* d_arr, long pitch, void* stream);

/* ---- noise layers (L1_to_L2/gen_noise_image.py:60-331 make_noise_cube, directive "R") --------
 * rip_dark_as_l1_dev: the dark cube's last G groups cast to the u16 L1 cube (:101-109, the "not adding" branch).
 * rip_add_read_noise_dev: white read noise N(0,1)*read/sqrt(N_k) on the active pixels, round(clip(.,0,65535)) (:121-135).
 * (the correlated part is rip_fill_refdata_1f_dev, the calibration rip_l1_to_l2_dev, the sky mode removal
 *  rip_block_nanmedian_dev + rip_medfit_eval_dev.)  rip_active_diff_dev: layer = L2(noisy) - L2(reference) on the
 * active window of two full-frame float32 planes -> dense [na,na] (:157-163). */
int rip_dark_as_l1_dev(rip_caldir* h, int G, uint16_t* d_data, void* stream);
int rip_add_read_noise_dev(rip_caldir* h, uint16_t* d_data, int G, const int32_t* reads_per_group, uint64_t seed, void* stream);
int rip_active_diff_dev(int device, const float* d_a, const float* d_b, int n, int nb, float* d_out, void* stream);

/* order statistics (0-based ranks, ascending, NaNs excluded; K <= 16 ranks, one CTA each) of a flat float32 device array
 * -> host; n_valid = number of non-NaN elements.  With rip_clip_dev (np.clip, NaN bounds propagate) this is the z clip of
 * the noise layers (gen_noise_image.py:165-171: np.percentile's linear interpolation is done by the caller). */
int rip_order_stats_dev(int device, const float* d_arr, long count, int K, const long* ranks, float* out, long* n_valid,
                        void* stream);
int rip_clip_dev(int device, float* d_arr, long count, float lo, float hi, void* stream);

/* noise directive "P" with flag r (gen_noise_image.py:258-321): per active pixel, n_samp Poisson draws with mean
 * e = clip(skylevel*gain*frame_time, 0), (draw - e)/gain accumulated sample by sample, averaged into the resultants
 * (group_of_read[i] = group containing sample i or -1) and contracted with the ramp-fit weights of the pixel's ramp end
 * (weights f32 [G,G]: row es, w_defined[es] = row exists; es = endslice > 0 ? endslice : G-1): d_diff [na,na] += . */
int rip_poisson_resample_dev(rip_caldir* h, const float* d_skylevel, const int8_t* d_endslice, int G, int n_samp,
                             const int32_t* group_of_read, const float* weights, const uint8_t* w_defined, double frame_time,
                             uint64_t seed, float* d_diff, void* stream);
/* Noise directive "O" (L1_to_L2/gen_noise_image.py:173-227 -> GalPoisson/draw_with_tilnus.py:12, find_tilnus.py:46):
 * d_diff f32 [na,na] += Pearson draw / clip(gain, 1e-4, 1e4) with I = max(gain * d_withsky, 0.01) and the moments
 * tilnu[i] = (nu21, nu31, nu41) [e/s units] of the ramp ending at group i = (endslice > 0 ? endslice : G-1); rows with
 * defined[i] == 0 and i <= start draw nothing.  All types the reference dispatches are generated on the device: I (Beta),
 * VI (beta prime), IV (Devroye's log-concave rejection in the angle, Heinrich 2004), and on the exact III / V lines the
 * Gamma / inverse Gamma.  Pixels whose parameters are invalid (where the reference raises ValueError) are counted in
 * *d_unsupported (device int32, zeroed by the caller) and draw 0. */
int rip_pearson_noise_dev(rip_caldir* h, const float* d_withsky, const int8_t* d_endslice, int G, int start,
                          const double* tilnu, const uint8_t* defined, uint64_t seed, float* d_diff,
                          int32_t* d_unsupported, void* stream);
/* log k(m, nu) of the Pearson IV density in the angle (Heinrich 2004; GalPoisson/draw_with_tilnus.py:296-306 with a = 1),
 * as the Type IV sampler of rip_pearson_noise_dev evaluates it (complex log-gamma by recurrence + Stirling): test hook. */
double rip_pearson4_logk_host(double m, double nu);

/* ---- pixel area from the WCS (utils/coordutils.py:17-82 pixelarea, used at L1_to_L2/gen_cal_image.py:618-622 with the
 * FITSWCS header of :82-83).  wcs = 211 doubles describing a zenithal FITS WCS with SIP:
 * CRPIX1,2 CRVAL1,2 CD1_1 CD1_2 CD2_1 CD2_2 LONPOLE proj(0 TAN, 1 STG) sip_order, A[p][q] (10x10), B[p][q] (10x10)
 * (romanimpreprocess_b200.utils.coordutils.FitsWCS.pack).  out [N,N] f32|f64 = pixel solid angle [sr] * inv_omega
 * (inv_omega = 1/pars.Omega_ideal gives AreaFactor).  Same construction as the reference, float64. */
int rip_pixel_area_dev(int device, const double* wcs, int nwcs, int N, double inv_omega, void* d_out, int out_dtype,
                       void* stream);
int rip_pixel_area_host(int device, const double* wcs, int nwcs, int N, double inv_omega, void* out, int out_dtype);
/* The same plane computed straight into the pipeline's resident AreaFactor buffer (see rip_pipeline_set_area): the
 * per-exposure AreaFactor then costs a 1.7 kB upload and a 0.3 ms kernel instead of a 67 MB copy. */
int rip_pipeline_set_area_wcs(rip_pipeline* p, const double* wcs, int nwcs, double inv_omega, int area_dtype);

#ifdef __cplusplus
}
#endif
#endif /* RIP_B200_H */
