// Internal device-pointer launchers shared between translation units (not part of the public ABI).
#pragma once
#include "rip_cal_core.cuh"
#include "rip_rt.h"

namespace rip {

// rip_stage.cu ---------------------------------------------------------------------------------------------
// In-place "DN-space" IPC deconvolution of the active region of one f32 plane [ny,nx] (border nb):
//   plane_act <- f32( ipc_rev(plane_act * g, K, order=2) / g ),  g = gain_act (optionally clipped below at g_lo)
// == one group of correct_cube (utils/ipc_linearity.py:185-186) == the IPC part of get_flat (flatutils.py:72-74).
// tmp must hold nya*nxa elements of double.  gain may be null (g = 1).
void launch_ipc_rev_dn(float* plane, int ny, int nx, int nb, const void* K, int k_dtype, const void* gain_full,
                       int g_dtype, bool clip_gain, float g_lo, void* tmp, cudaStream_t st);

// flat pad/flag/clip (utils/flatutils.py:44-69); pdq may be null
void launch_flat_prepare(const float* flat, int n, int nb, const void* gain, int g_dtype, uint32_t* pdq,
                         int ipc_deconvolve, float* out, cudaStream_t st);

// rip_fit.cu -----------------------------------------------------------------------------------------------
// The fused K1 kernel.  gain/ipc dtypes select the instantiation.  Plan must already be on the device.
void launch_cal_fused(const CalArgs& A, int g_dtype, int k_dtype, int threads, cudaStream_t st);
size_t cal_fused_smem_bytes(int G, int g_dtype, int k_dtype, int threads);

// rip_v2.cu ------------------------------------------------------------------------------------------------
namespace v2 { struct Args; struct f4; }
bool v2_supported(int G, int P, bool k64 = false);
// The throughput kernel is instantiated for P = 4 and P = 11 Legendre coefficients; a CALDIR with fewer runs the next
// larger instantiation on records padded with zero coefficients (phi + 0 * P_L(z) == phi for |z| <= 1; the extrapolating
// evaluation for |z| > 1 sums the same zeros): same values as the exact-P evaluation.
inline int v2_pad_P(int P) { return P <= 4 ? 4 : (P <= 11 ? 11 : P); }
void launch_cal_fused_v2k64(const v2::Args& A, int G, int P, cudaStream_t st);
bool v6_supported(int G, int P);
void launch_cal_fused_v6(const v2::Args& A, int G, int P, cudaStream_t st);
void launch_cal_fused_v6k64(const v2::Args& A, int G, int P, cudaStream_t st);
int v2_default_band_rows(int device, int n, int G, int ctas_per_sm = 0);
void launch_cal_fused_v3(const v2::Args& A, int G, int P, int variant, cudaStream_t st);
void v2_plan_to_device(const rip_ramp_plan* plan, cudaStream_t st);
void v3_plan_to_device(const rip_ramp_plan* plan, const void* fast_tab, cudaStream_t st);  // rip_v3.cu
void launch_cal_fused_v2(const v2::Args& A, int G, int P, cudaStream_t st);
void v2_pack(rip_caldir* h, int G, cudaStream_t st);
v2::f4* v2_rec1_row0(rip_caldir* h, int G);  // record of detector row 0 inside the padded allocations
v2::f4* v2_recK_row0(rip_caldir* h);

// rip_area.cu ----------------------------------------------------------------------------------------------
void launch_pixel_area(const double* wcs, int nwcs, int N, double inv_omega, void* d_out, int out_dtype, cudaStream_t st);

}  // namespace rip
