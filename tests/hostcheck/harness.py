"""TEST INFRASTRUCTURE ONLY: run the fused-kernel per-pixel source on the CPU (see hostcheck.cpp).

Builds ``libhostcheck.so`` with g++ from ``csrc/rip_cal_core.cuh`` (the same source nvcc compiles into the CUDA
kernel) and drives it with the march-step schedule of the kernel.  The static products and reference-pixel
statistics, which the CUDA library computes in its own kernels, are supplied here by the oracle.
"""

import ctypes as C
import os
import subprocess

import numpy as np

from oracle import rip_oracle as orc
from romanimpreprocess_b200 import _lib
from romanimpreprocess_b200.utils import fitting

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libhostcheck.so")
SRC = os.path.join(HERE, "hostcheck.cpp")
CSRC = os.path.join(ROOT, "romanimpreprocess_b200", "csrc")


class CalArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "nb", "G", "P", "band_rows", "do_refpix", "do_not_flag_first",
                                       "exclude_first", "sat_backup", "area_dtype")] + \
               [(k, C.c_void_p) for k in ("raw", "area", "rowcorr", "chan_m", "chan_c", "dark", "bias", "coefs", "Smin",
                                          "Smax", "Sref", "aux", "sdq", "thr", "gain", "ipc", "read", "dslope", "flat",
                                          "w_exact", "slope", "err_read", "err_poisson", "pdq", "endslice", "rdq",
                                          "lincube")]  # fmt: skip


class V2Args(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "ntile", "band_rows", "do_refpix", "do_not_flag_first", "exclude_first",
                                       "sat_backup", "area_dtype")] + [("negzero", C.c_float), ("pad_", C.c_int)] + \
               [(k, C.c_void_p) for k in ("raw", "area", "rowcorr", "chan_m", "chan_c", "rec1", "recK", "thr", "w_exact",
                                          "slope", "err_read", "err_poisson", "pdq", "endslice", "rdq", "lincube",
                                          "chan_line")]  # fmt: skip


class PackSrc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "nb", "G", "P")] + \
               [(k, C.c_void_p) for k in ("dark", "bias", "coefs", "Smin", "Smax", "Sref", "gain", "aux", "ipc", "read",
                                          "dslope", "flat", "sdq")]  # fmt: skip


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("rip_cal_core.cuh", "rip_math.cuh", "rip_v2_core.cuh")] + [
        os.path.join(ROOT, "include", "rip_b200.h")
    ]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
           f"-I{os.path.join(ROOT, 'include')}", f"-I{CSRC}", SRC, "-o", SO]  # fmt: skip
    subprocess.run(cmd, check=True)
    return SO


_h = None


def lib():
    global _h
    if _h is None:
        _h = C.CDLL(build())
        _h.hostcheck_cal_fused.restype = C.c_int
        _h.hostcheck_cal_fused.argtypes = [C.POINTER(CalArgs), C.POINTER(_lib.RampPlan), C.c_int, C.c_int, C.c_int]
        assert _h.hostcheck_sizeof_calargs() == C.sizeof(CalArgs)
        _h.hostcheck_cal_fused_v2.restype = C.c_int
        _h.hostcheck_cal_fused_v2.argtypes = [C.POINTER(V2Args), C.POINTER(PackSrc), C.POINTER(_lib.RampPlan)]
        _h.hostcheck_cal_fused_v6.restype = C.c_int
        _h.hostcheck_cal_fused_v6.argtypes = [C.POINTER(V2Args), C.POINTER(PackSrc), C.POINTER(_lib.RampPlan)]
        _h.hostcheck_cal_fused_v6k64.restype = C.c_int
        _h.hostcheck_cal_fused_v6k64.argtypes = [C.POINTER(V2Args), C.POINTER(PackSrc), C.POINTER(_lib.RampPlan)]
        _h.hostcheck_cal_fused_v2k64.restype = C.c_int
        _h.hostcheck_cal_fused_v2k64.argtypes = [C.POINTER(V2Args), C.POINTER(PackSrc), C.POINTER(_lib.RampPlan)]
        assert _h.hostcheck_sizeof_v2args() == C.sizeof(V2Args)
        assert _h.hostcheck_sizeof_packsrc() == C.sizeof(PackSrc)
        _h.hostcheck_shared_div.restype = C.c_long
        _h.hostcheck_shared_div.argtypes = [C.c_long, C.c_uint, C.c_float, C.c_float, C.c_float]
    return _h


def static_products(c, nb=4):
    """thr_eff, aux, sdq, dslope_ipc, flat_ipc as rip_caldir_create builds them (csrc/rip_caldir.cu), via the oracle."""
    lin = c["linearitylegendre"]
    n = lin["Sref"].shape[0]
    sat_dq = c["saturation"]["dq"]
    thr = c["saturation"]["data"].astype(np.float32).copy()
    thr[((sat_dq & orc.NO_SAT_CHECK) != 0) | np.isnan(thr)] = np.inf
    ld = lin["dq"]
    md = orc.expand_gw(c["mask"]["dq"]) if "mask" in c else np.zeros((n, n), np.uint32)  # (do_dqinit, expand_gw_flagging=1)
    aux = np.zeros((n, n), np.uint8)
    aux |= np.where(ld & (orc.NO_LIN_CORR | orc.REFERENCE_PIXEL) != 0, 1, 0).astype(np.uint8)
    aux |= np.where((ld | md) & orc.REFERENCE_PIXEL != 0, 2, 0).astype(np.uint8)
    sdq = np.zeros((n, n), np.uint32)
    has_ipc = "ipc4d" in c
    flat = orc.get_flat(c["flat"]["data"], c["gain"]["data"], c["ipc4d"]["data"] if has_ipc else None, nb, sdq,
                        ipc_deconvolve=has_ipc)  # fmt: skip
    sdq |= md | (sat_dq & orc.NO_SAT_CHECK) | ld
    sdq[nb:-nb, nb:-nb] |= c["dark"]["dq"][nb:-nb, nb:-nb]
    ds = np.array(c["dark"]["dark_slope"], dtype=np.float32)[None].copy()
    if has_ipc:
        orc.correct_cube(ds, c["ipc4d"]["data"], c["gain"]["data"])
    return thr, aux, sdq, ds[0], flat


def refpix_stats(data_u16, amp33_u16, c):
    """rowcorr f64 [G,n], chan_m/chan_c f64 [G,32] as K0 computes them; restates gen_cal_image.py:531-555."""
    G, n, _ = data_u16.shape
    dark = c["dark"]["data"]
    read = c["read"]
    slope = orc.optimal_refout_slope(read)
    rowcorr = np.zeros((G, n))
    cm = np.zeros((G, 32))
    cc = np.zeros((G, 32))
    for j in range(G):
        ro = amp33_u16[j].astype(np.float32) - read["amp33"]["med"]
        ro = ro - np.median(ro)
        ref_med = np.median(ro, axis=1)
        ctr = np.median(ref_med)
        rowcorr[j] = slope * (ref_med - ctr)
        img = data_u16[j].astype(np.float32) - dark[j]
        rows = np.r_[0:4, n - 4 : n]
        sub = (img[rows] - rowcorr[j][rows][:, None]).astype(np.float32)
        for ch in range(n // 128):
            b = np.median(sub[0:4, 128 * ch : 128 * (ch + 1)])
            t = np.median(sub[4:8, 128 * ch : 128 * (ch + 1)])
            m = (float(t) - float(b)) / ((n - 2.5) - 1.5)
            cm[j, ch] = m
            cc[j, ch] = float(b) - m * 1.5
    return rowcorr, cm, cc


def run_fused(cal, data_u16, amp33_u16, read_pattern, frame_time, area, config=None, do_refpix=False, threads=64,
              band_rows=16, want_rdq=True, want_lin=True, v2=False, v6=False):  # fmt: skip
    """Host emulation of rip_l1_to_l2 (same argument meaning as the oracle's l1_to_l2); ``v2`` selects the v2 kernel
    source (rip_v2_core.cuh: all-f32 planes, G in {8,16}, P in {4,11})."""
    config = config or {}
    c = {k: v["roman"] for k, v in cal.items()}
    nb = 4
    G, n, _ = data_u16.shape
    na = n - 2 * nb
    exclude_first = config.get("EXCLUDE_FIRST", True)
    meta = orc.make_meta(read_pattern, frame_time)
    uopt = config.get("RAMP_OPT_PARS", {"slope": 0.4, "gain": 1.8, "sigma_read": 6.5})
    u_ = float(uopt["slope"]) / float(uopt["gain"]) / float(uopt["sigma_read"]) ** 2
    meta["K"] = fitting.construct_weights(u_, meta, exclude_first=exclude_first)
    if "JUMP_DETECT_PARS" in config:
        meta["jump_detect_pars"] = config["JUMP_DETECT_PARS"]
    plan, w_exact = fitting.build_plan(meta, exclude_first)
    thr, aux, sdq, dslope, flat = static_products(c, nb)
    keep = []

    def p(a, dt=None):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    if v2 or v6:
        return _run_v2(c, data_u16, amp33_u16, read_pattern, area, config, do_refpix, band_rows, want_rdq, want_lin,
                       plan, w_exact, (thr, aux, sdq, dslope, flat), meta, p, v6=v6)
    A = CalArgs()
    A.n, A.nb, A.G, A.P = n, nb, G, c["linearitylegendre"]["data"].shape[0]
    A.band_rows = band_rows
    A.do_refpix = 1 if do_refpix else 0
    A.do_not_flag_first = 1 if list(read_pattern[0]) == [0] else 0
    A.exclude_first = 1 if exclude_first else 0
    A.sat_backup = config.get("SATURATION_BACKUP", 1)
    gain = _lib.as_float_plane(c["gain"]["data"])
    ipc = _lib.as_float_plane(c["ipc4d"]["data"]) if "ipc4d" in c else None
    A.area_dtype = _lib.RIP_F32
    if area is not None:
        area = _lib.as_float_plane(area)
        A.area_dtype = _lib.float_tag(area)
    A.raw = p(data_u16, np.uint16)
    A.area = p(area)
    if do_refpix:
        rc, cm, cc = refpix_stats(data_u16, amp33_u16, c)
        A.rowcorr, A.chan_m, A.chan_c = p(rc), p(cm), p(cc)
    A.dark = p(c["dark"]["data"], np.float32)
    if "biascorr" in c:
        bc = c["biascorr"]["data"]
        A.bias = p(bc[bc.shape[0] - G :], np.float32)
    A.coefs = p(c["linearitylegendre"]["data"], np.float32)
    A.Smin = p(c["linearitylegendre"]["Smin"], np.float32)
    A.Smax = p(c["linearitylegendre"]["Smax"], np.float32)
    A.Sref = p(c["linearitylegendre"]["Sref"], np.float32)
    A.aux, A.sdq, A.thr = p(aux), p(sdq), p(thr)
    A.gain, A.ipc = p(gain), p(ipc)
    A.read = p(c["read"]["data"], np.float32)
    A.dslope, A.flat = p(dslope, np.float32), p(flat, np.float32)
    A.w_exact = p(w_exact)
    out = {
        "slope": np.full((n, n), np.nan, np.float32),
        "err_read": np.full((n, n), np.nan, np.float32),
        "err_poisson": np.full((n, n), np.nan, np.float32),
        "pdq": np.full((n, n), 0xDEADBEEF, np.uint32),
        "endslice": np.full((na, na), 99, np.int8),
    }
    A.slope, A.err_read, A.err_poisson = p(out["slope"]), p(out["err_read"]), p(out["err_poisson"])
    A.pdq, A.endslice = p(out["pdq"]), p(out["endslice"])
    if want_rdq:
        out["rdq"] = np.full((G, n, n), 0xEE, np.uint8)
        A.rdq = p(out["rdq"])
    if want_lin:
        out["ipc"] = np.full((G, n, n), np.nan, np.float32)
        A.lincube = p(out["ipc"])
    rc = lib().hostcheck_cal_fused(C.byref(A), C.byref(plan), _lib.float_tag(gain),
                                   _lib.RIP_F32 if ipc is None else _lib.float_tag(ipc), threads)  # fmt: skip
    assert rc == 0
    out["K"] = meta["K"]
    return out


def _run_v2(c, data_u16, amp33_u16, read_pattern, area, config, do_refpix, band_rows, want_rdq, want_lin, plan, w_exact,  # noqa: PLR0913
            static, meta, p, v6=False):  # fmt: skip
    thr, aux, sdq, dslope, flat = static
    G, n, _ = data_u16.shape
    na = n - 8
    P = c["linearitylegendre"]["data"].shape[0]
    k64 = c["ipc4d"]["data"].dtype == np.float64
    assert c["gain"]["data"].dtype == np.float32
    S = PackSrc()
    S.n, S.nb, S.G, S.P = n, 4, G, P
    S.dark = p(c["dark"]["data"], np.float32)
    if "biascorr" in c:
        bc = c["biascorr"]["data"]
        S.bias = p(bc[bc.shape[0] - G :], np.float32)
    S.coefs = p(c["linearitylegendre"]["data"], np.float32)
    S.Smin = p(c["linearitylegendre"]["Smin"], np.float32)
    S.Smax = p(c["linearitylegendre"]["Smax"], np.float32)
    S.Sref = p(c["linearitylegendre"]["Sref"], np.float32)
    S.gain = p(c["gain"]["data"], np.float32)
    S.aux, S.sdq = p(aux), p(sdq)
    S.ipc = p(c["ipc4d"]["data"], np.float64 if k64 else np.float32)
    S.read = p(c["read"]["data"], np.float32)
    S.dslope, S.flat = p(dslope, np.float32), p(flat, np.float32)
    A = V2Args()
    A.n, A.band_rows = n, band_rows
    A.do_refpix = 1 if do_refpix else 0
    A.do_not_flag_first = 1 if list(read_pattern[0]) == [0] else 0
    A.exclude_first = 1 if config.get("EXCLUDE_FIRST", True) else 0
    A.sat_backup = config.get("SATURATION_BACKUP", 1)
    A.negzero = -0.0
    A.area_dtype = _lib.RIP_F32
    if area is not None:
        area = _lib.as_float_plane(area)
        A.area_dtype = _lib.float_tag(area)
    A.raw, A.area = p(data_u16, np.uint16), p(area)
    if do_refpix:
        rc, cm, cc = refpix_stats(data_u16, amp33_u16, c)
        A.rowcorr, A.chan_m, A.chan_c = p(rc), p(cm), p(cc)
    A.thr, A.w_exact = p(thr), p(w_exact)
    out = {
        "slope": np.full((n, n), np.nan, np.float32),
        "err_read": np.full((n, n), np.nan, np.float32),
        "err_poisson": np.full((n, n), np.nan, np.float32),
        "pdq": np.full((n, n), 0xDEADBEEF, np.uint32),
        "endslice": np.full((na, na), 99, np.int8),
    }
    A.slope, A.err_read, A.err_poisson = p(out["slope"]), p(out["err_read"]), p(out["err_poisson"])
    A.pdq, A.endslice = p(out["pdq"]), p(out["endslice"])
    if want_rdq:
        out["rdq"] = np.full((G, n, n), 0xEE, np.uint8)
        A.rdq = p(out["rdq"])
    if want_lin:
        out["ipc"] = np.full((G, n, n), np.nan, np.float32)
        A.lincube = p(out["ipc"])
    if v6:
        fn = lib().hostcheck_cal_fused_v6k64 if k64 else lib().hostcheck_cal_fused_v6
    else:
        fn = lib().hostcheck_cal_fused_v2k64 if k64 else lib().hostcheck_cal_fused_v2
    rc = fn(C.byref(A), C.byref(S), C.byref(plan))
    assert rc == 0, "v2 host check: unsupported (G, P)"
    out["K"] = meta["K"]
    return out
