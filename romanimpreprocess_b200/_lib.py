"""ctypes binding of ``librip_b200.so`` (C ABI declared in ``include/rip_b200.h``).

The library is the product: there is no CPU fallback.  If the shared object is missing or cannot be loaded the
first call raises ``RuntimeError`` with build instructions.
"""

import ctypes as C
import os

import numpy as np

RIP_GMAX = 16
RIP_MAXVAR = 16
RIP_MAXSLICE = 256
RIP_PMAX = 16
RIP_F32, RIP_F64, RIP_I32, RIP_U16 = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RIP_B200_LIB") or os.path.join(_HERE, "librip_b200.so")  # (override: development A/B builds)


class RampSlice(C.Structure):
    _fields_ = [("i", C.c_int32), ("di", C.c_int32), ("dt", C.c_float), ("inv_dt", C.c_float), ("A", C.c_float),
                ("B", C.c_float)]  # fmt: skip


class RampPlan(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("start", C.c_int32), ("nvar", C.c_int32), ("reserved", C.c_int32),
        ("tbar", C.c_float * RIP_GMAX), ("tau", C.c_float * RIP_GMAX), ("nreads", C.c_float * RIP_GMAX),
        ("var_ngrp", C.c_int32 * RIP_MAXVAR),
        ("var_K", (C.c_float * RIP_GMAX) * RIP_MAXVAR),
        ("var_coef", C.c_float * RIP_MAXVAR), ("var_rfac", C.c_float * RIP_MAXVAR),
        ("var_slice_off", C.c_int32 * (RIP_MAXVAR + 1)),
        ("slices", RampSlice * RIP_MAXSLICE),
        ("IthreshA_f", C.c_float), ("IthreshB_f", C.c_float),
        ("SthreshA", C.c_double), ("SthreshB", C.c_double), ("logIratio", C.c_double),
        ("band", C.c_float), ("thrA_f", C.c_float), ("thrK_f", C.c_float), ("invIA_f", C.c_float),
    ]  # fmt: skip


class CaldirDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("nb", C.c_int32), ("P", C.c_int32), ("n_dark", C.c_int32), ("n_bias", C.c_int32),
        ("gain_dtype", C.c_int32), ("ipc_dtype", C.c_int32), ("has_amp33", C.c_int32),
        ("lin_coefs", C.c_void_p), ("Smin", C.c_void_p), ("Smax", C.c_void_p), ("Sref", C.c_void_p),
        ("lin_dq", C.c_void_p), ("mask_dq", C.c_void_p), ("sat_thresh", C.c_void_p), ("sat_dq", C.c_void_p),
        ("gain", C.c_void_p), ("ipc", C.c_void_p), ("read", C.c_void_p), ("resetnoise", C.c_void_p),
        ("dark_cube", C.c_void_p), ("dark_slope", C.c_void_p), ("dark_dq", C.c_void_p), ("biascorr", C.c_void_p),
        ("flat", C.c_void_p), ("amp33_med", C.c_void_p), ("amp33_std", C.c_void_p),
        ("biascorr_t0", C.c_double), ("m_pink", C.c_double), ("ru_pink", C.c_double), ("c_pink", C.c_double),
        ("u_pink", C.c_double), ("refout_slope", C.c_double),
    ]  # fmt: skip


class L1L2Params(C.Structure):
    _fields_ = [("G", C.c_int32), ("exclude_first", C.c_int32), ("sat_backup", C.c_int32),
                ("do_not_flag_first", C.c_int32), ("do_refpix", C.c_int32), ("area_dtype", C.c_int32),
                ("threads", C.c_int32), ("band_rows", C.c_int32)]  # fmt: skip


class L2Out(C.Structure):
    _fields_ = [("slope", C.c_void_p), ("err_read", C.c_void_p), ("err_poisson", C.c_void_p), ("pdq", C.c_void_p),
                ("endslice", C.c_void_p), ("rdq", C.c_void_p), ("lin_cube", C.c_void_p)]  # fmt: skip


class FwdParams(C.Structure):
    _fields_ = [("G", C.c_int32), ("n_reads", C.c_int32), ("reads_per_group", C.c_int32 * RIP_GMAX),
                ("read_index", C.c_int32 * 64), ("read_time", C.c_double), ("seed", C.c_uint64),
                ("add_read_noise", C.c_int32), ("add_reset_noise", C.c_int32), ("add_biascorr", C.c_int32),
                ("quantize", C.c_int32), ("cr_enable", C.c_int32), ("pad_", C.c_int32), ("cr_flux", C.c_double),
                ("cr_area", C.c_double), ("cr_conversion_factor", C.c_double), ("cr_pixel_size", C.c_double),
                ("cr_pixel_depth", C.c_double)]  # fmt: skip


_SIGS = {
    "rip_last_error": (C.c_char_p, []),
    "rip_abi_version": (C.c_int, []),
    "rip_launch_count": (C.c_longlong, []),
    "rip_struct_size": (C.c_long, [C.c_int]),
    "rip_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rip_device_sync": (C.c_int, [C.c_int]),
    "rip_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "rip_host_free": (C.c_int, [C.c_void_p]),
    "rip_dev_alloc": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_size_t]),
    "rip_dev_free": (C.c_int, [C.c_int, C.c_void_p]),
    "rip_copy_h2d": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rip_copy_d2h": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rip_stream_sync": (C.c_int, [C.c_int, C.c_void_p]),
    "rip_lin_eval": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_void_p,
                               C.c_void_p]),
    "rip_multilin": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_invlinearity": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "rip_ipc_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                              C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "rip_ipc_rev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "rip_correct_cube": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p, C.c_int]),
    "rip_jump_detect": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(RampPlan),
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "rip_ramp_fit": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.POINTER(RampPlan), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "rip_row_medians": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_refsub_row_apply": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "rip_refsub_channel": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "rip_get_flat": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_void_p]),
    "rip_flag_saturation": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p]),
    "rip_caldir_create": (C.c_int, [C.c_int, C.POINTER(CaldirDesc), C.POINTER(C.c_void_p)]),
    "rip_caldir_destroy": (None, [C.c_void_p]),
    "rip_caldir_get_static": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "rip_l1_to_l2_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(L1L2Params),
                                    C.POINTER(RampPlan), C.c_void_p, C.POINTER(L2Out)]),
    "rip_caldir_prefetch_refpix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "rip_l1_to_l2_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(L1L2Params),
                                   C.POINTER(RampPlan), C.c_void_p, C.POINTER(L2Out), C.c_void_p]),
    "rip_pipeline_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "rip_pipeline_destroy": (None, [C.c_void_p]),
    "rip_pipeline_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(L1L2Params),
                                      C.POINTER(RampPlan), C.c_void_p, C.POINTER(L2Out), C.POINTER(C.c_long)]),
    "rip_pipeline_wait": (C.c_int, [C.c_void_p, C.c_long]),
    "rip_medfit_solve": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rip_bin_masked_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_gauss_hist_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "rip_pipeline_set_area": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "rip_pipeline_set_area_wcs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "rip_pixel_area_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p]),
    "rip_pixel_area_host": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int]),
    "rip_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "rip_profile_fetch": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "rip_refpix_stats_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "rip_il_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rip_il_apply_planes": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "rip_make_l1_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FwdParams), C.c_void_p]),
    "rip_make_l1_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(FwdParams), C.c_void_p, C.c_void_p]),
    "rip_fwd_cr_groups_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rip_fwd_cr_groups_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "rip_fwd_cum_counts_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "rip_sim_calprep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rip_sim_counts_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    "rip_sim_counts_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    "rip_noise_1f_frames_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "rip_fill_refdata_1f_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int,
                                          C.c_void_p]),
    "rip_fill_refdata_1f_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64,
                                           C.c_int]),
    "rip_mask_build_host": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_moments_accumulate_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "rip_moments_finalize_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_long, C.c_void_p]),
    "rip_l1_embed_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_realization_record_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rip_block_nanmedian_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_medfit_host": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rip_medfit_eval_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_long, C.c_void_p]),
    "rip_dark_as_l1_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_add_read_noise_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]),
    "rip_active_diff_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rip_order_stats_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_long),
                                      C.c_void_p]),
    "rip_clip_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_long, C.c_float, C.c_float, C.c_void_p]),
    "rip_pearson4_logk_host": (C.c_double, [C.c_double, C.c_double]),
    "rip_pearson_noise_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rip_poisson_resample_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_double, C.c_uint64, C.c_void_p, C.c_void_p]),
    "rip_stack_median_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p]),
}  # fmt: skip

EXPORTED_SYMBOLS = sorted(_SIGS)

_lib = None


class RipError(RuntimeError):
    """An error reported by librip_b200 (message from rip_last_error())."""


def lib():
    """Load (once) and return the ctypes handle; fail loudly if the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built.  Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C romanimpreprocess_b200/csrc`).  There is no CPU fallback."
        )
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as e:
        raise RuntimeError(f"cannot load {LIB_PATH}: {e}.  There is no CPU fallback.") from e
    for name, (res, args) in _SIGS.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc):
    """Turn a non-zero status into RipError."""
    if rc != 0:
        msg = lib().rip_last_error()
        raise RipError(msg.decode() if msg else f"librip_b200 error {rc}")


def ptr(a):
    """Host pointer of a C-contiguous ndarray (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


def as_c(a, dtype):
    """C-contiguous array of the given dtype (no copy when already so)."""
    return np.ascontiguousarray(a, dtype=dtype)


def float_tag(a):
    """dtype tag of a floating plane; non-f64 inputs are treated as f32 (the caller converts)."""
    return RIP_F64 if a.dtype == np.float64 else RIP_F32


def as_float_plane(a):
    """Keep f64 planes as f64 (the reference's arithmetic follows the plane dtype: SURVEY A0), else f32."""
    a = np.asarray(a)
    if a.dtype == np.float64:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a, dtype=np.float32)


def device_count():
    n = C.c_int(0)
    check(lib().rip_device_count(C.byref(n)))
    return n.value


def pinned_empty(shape, dtype):
    """An ndarray over page-locked host memory (rip_host_alloc); freed when the last view is garbage-collected."""
    import weakref  # noqa: PLC0415

    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    p = C.c_void_p()
    check(lib().rip_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * nbytes).from_address(p.value)
    base = np.frombuffer(buf, dtype=dtype, count=count)
    weakref.finalize(base, _free_pinned, p.value)
    return base.reshape(shape)


def _free_pinned(address):
    if _lib is not None:
        _lib.rip_host_free(C.c_void_p(address))
