// The CALDIR handle (opaque in the public ABI), shared by rip_caldir.cu and rip_fwd.cu.
#pragma once
#include "rip_rt.h"

namespace rip {
struct SelState { uint32_t prefix[2]; uint32_t rank[2]; };
}
using rip::DevBuf;
using rip::DevRaw;
using rip::SelState;

struct rip_caldir {
    int device = 0;
    rip_caldir_desc d{};  // scalar fields only are meaningful after creation
    int n = 0, nb = 0, na = 0, P = 0;
    bool has_ipc = false, has_bias = false, has_amp33 = false;
    double refout_slope = 0.0;
    // CALDIR planes
    DevBuf<float> coefs, Smin, Smax, Sref, sat_thr, read, resetnoise, dark_cube, dark_slope, biascorr, flat, amp_med;
    DevBuf<uint32_t> lin_dq;
    DevRaw gain, ipc;
    // static products
    DevBuf<float> thr_eff, dslope_ipc, flat_ipc;
    DevBuf<uint8_t> aux;
    DevBuf<uint32_t> sdq;
    // cal_fused v2: packed per-(row, tile) records (rip_v2_core.cuh), built lazily per group count
    DevBuf<float> v2_rec1, v2_recK;
    int v2_G = 0;
    // K0 workspace
    // K0 (reference-pixel statistics) workspaces: two sets, so that the statistics of the NEXT exposure can be computed on
    // the handle's side stream while the fused kernel of the current one runs (rip_caldir_prefetch_refpix)
    struct K0Work {
        DevBuf<uint32_t> hist, k0_ticket;
        DevBuf<SelState> sel;
        DevBuf<float> rowA, rowB, gmed;
        DevBuf<double> rowcorr, chan_m, chan_c, chan_line;
        cudaEvent_t ev_k0 = nullptr;    // statistics complete (recorded on the side stream)
        cudaEvent_t ev_used = nullptr;  // last fused kernel that read this set has finished
        bool used_recorded = false;
        const void* key = nullptr;      // raw cube the set was precomputed for (null: none pending)
        bool busy = false;              // the side stream has written this set since the user stream last waited for it
    };
    K0Work k0w[2];
    int k0_cur = 0;                     // set read by the most recent fused launch
    cudaStream_t s_k0 = nullptr;        // low-priority side stream of the look-ahead
    // host-entry workspace
    DevBuf<uint16_t> w_raw, w_amp;
    DevRaw w_area;
    DevBuf<float> w_slope, w_er, w_ep, w_lin;
    DevBuf<uint32_t> w_pdq;
    DevBuf<int8_t> w_end;
    DevBuf<uint8_t> w_rdq;
    // forward path (rip_sim.cu): amp33 noise plane, scene calibration planes (lazy), 1/f frame workspace (lazy)
    DevBuf<float> amp_std, sim_dark, sim_flat, f_frames, f_work;
    DevBuf<double> f_sums;
    // forward ramp (rip_fwd.cu): certificate planes of the fast inverse (lazy, once per CALDIR), per-call workspace
    DevBuf<float> lin_A, lin_m, f_start;
    DevBuf<int32_t> f_cum;
    DevBuf<uint32_t> f_crg;              // cosmic-ray group bits of the last forward ramp
    DevBuf<float> cr_len_cdf, cr_dedx_cdf;  // samplers of romanisim.cr (built on first use)
    bool f_crg_valid = false;
    bool warned_generic = false;          // the "generic kernel" notice was printed for this handle
    cudaStream_t stream = nullptr;
    // optional per-launch timing of the fused kernel (rip_profile_enable): event pairs recorded on the launch stream
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;  // [2*k] start, [2*k+1] stop
    size_t prof_used = 0;
};

