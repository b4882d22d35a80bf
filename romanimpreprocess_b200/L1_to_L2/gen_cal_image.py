"""L1 -> L2 calibration on the GPU: drop-in for the numerics of ``romanimpreprocess.L1_to_L2.gen_cal_image``.

The reference driver ``calibrateimage`` (reference L1_to_L2/gen_cal_image.py:480-739) runs, per exposure,
dq-init -> saturation flagging -> reference-pixel loop -> bias correction -> ``multilin`` -> ``correct_cube`` ->
``do_ramp_fit`` -> ``subtract_dark_current`` -> error split -> ``get_flat``/area division -> (sky, packaging,
I/O) -> endslice.  Everything from the raw ``uint16`` cube to the flat-fielded slope / error / DQ / endslice
arrays (reference lines 503-629 and 697-709) is one call into ``librip_b200.so`` here
(``rip_l1_to_l2_host`` / ``rip_l1_to_l2_dev``): a reference-pixel statistics pass and ONE fused kernel that
streams the cube through HBM once.  YAML configuration, CALDIR files and the ASDF data models are unchanged.

Classes
-------
CalDir
    One SCA's calibration reference data resident on a GPU (the reference re-opens the ASDF files inside every
    routine; here they are uploaded once and the exposure-independent products are precomputed).

Functions
---------
exposure_meta
    ``N``, ``tbar``, ``tau`` of a read pattern (reference gen_cal_image.py:123-140).
calibrate_arrays
    L1 arrays -> L2 arrays, host buffers (copies inside).
calibrate_device
    The same on device pointers (asynchronous; for callers that keep exposures resident).
calibrateimage
    The reference's entry point: config dict -> L2 ASDF file (needs the ASDF / roman_datamodels stack for I/O).
"""

import ctypes as C

import numpy as np

from .. import _lib, pars
from ..caltree import open_tree
from ..dqflags import pixel
from ..utils import fitting


def exposure_meta(read_pattern, frame_time):
    """``meta`` dictionary of an exposure exactly as ``initializationstep`` builds it (gen_cal_image.py:123-140)."""
    ngrp = len(read_pattern)
    meta = {"frame_time": frame_time, "read_pattern": [list(g) for g in read_pattern], "ngrp": ngrp}
    meta["tbar"] = np.zeros(ngrp, dtype=np.float32)
    meta["tau"] = np.zeros(ngrp, dtype=np.float32)
    meta["N"] = np.zeros(ngrp, dtype=np.int16)
    for i in range(ngrp):
        meta["N"][i] = len(read_pattern[i])
        t0 = read_pattern[i][0]
        meta["tbar"][i] = (t0 + (meta["N"][i] - 1) / 2.0) * frame_time
        meta["tau"][i] = (t0 + (meta["N"][i] - 1) * (2 * meta["N"][i] - 1) / (6.0 * meta["N"][i])) * frame_time
    meta["nborder"] = pars.nborder
    return meta


def ramp_setup(meta, config):
    """Weights and jump parameters as ``do_ramp_fit`` sets them (gen_cal_image.py:434-444); returns the ramp plan."""
    exclude_first = config.get("EXCLUDE_FIRST", True)
    uopt = {"slope": 0.4, "gain": 1.8, "sigma_read": 6.5}
    if "RAMP_OPT_PARS" in config:
        uopt = config["RAMP_OPT_PARS"]
    u_ = float(uopt["slope"]) / float(uopt["gain"]) / float(uopt["sigma_read"]) ** 2
    meta["K"] = fitting.construct_weights(u_, meta, exclude_first=exclude_first)
    meta["ramp_opt_pars"] = uopt
    if "JUMP_DETECT_PARS" in config:
        meta["jump_detect_pars"] = config["JUMP_DETECT_PARS"]
    return fitting.build_plan(meta, exclude_first)


class CalDir:
    """
    Calibration reference data of one SCA resident on one GPU (``rip_caldir_create``).

    Parameters
    ----------
    caldir : dict
        The ``CALDIR`` dictionary of the configuration: keys ``linearitylegendre``, ``saturation``, ``gain``,
        ``read``, ``dark``, ``flat`` (required), ``mask``, ``ipc4d``, ``biascorr`` (optional).  Values are ASDF
        file names (as in the reference) or in-memory trees.
    device : int
        CUDA device ordinal.
    """

    def __init__(self, caldir, device=0):
        self.device = device
        self.source = caldir  # the CALDIR it was built from (file names or in-memory trees), for late host-side reads
        self._h = C.c_void_p()
        d = _lib.CaldirDesc()
        keep = []

        def put(field, arr, dtype=None):
            a = _lib.as_float_plane(arr) if dtype is None else _lib.as_c(arr, dtype)
            keep.append(a)
            setattr(d, field, a.ctypes.data)
            return a

        with open_tree(caldir["linearitylegendre"]) as f:
            r = f["roman"]
            coefs = put("lin_coefs", r["data"], np.float32)
            put("Smin", r["Smin"], np.float32)
            put("Smax", r["Smax"], np.float32)
            put("Sref", r["Sref"], np.float32)
            ldq = put("lin_dq", r["dq"], np.uint32)
        self.P, self.n = coefs.shape[0], coefs.shape[-1]
        self.nb = pars.nborder
        self.na = self.n - 2 * self.nb
        d.n, d.nb, d.P = self.n, self.nb, self.P
        self.lin_dq_active = ldq[self.nb : self.n - self.nb, self.nb : self.n - self.nb].copy()  # IL.set_dq
        if "mask" in caldir:
            with open_tree(caldir["mask"]) as f:
                put("mask_dq", f["roman"]["dq"], np.uint32)
        with open_tree(caldir["saturation"]) as f:
            put("sat_thresh", f["roman"]["data"], np.float32)
            put("sat_dq", f["roman"]["dq"], np.uint32)
        with open_tree(caldir["gain"]) as f:
            g = put("gain", f["roman"]["data"])
            d.gain_dtype = _lib.float_tag(g)
        self.has_ipc = "ipc4d" in caldir
        if self.has_ipc:
            with open_tree(caldir["ipc4d"]) as f:
                k = put("ipc", f["roman"]["data"])
                d.ipc_dtype = _lib.float_tag(k)
        self.refout_slope = None
        d.refout_slope = float("nan")
        with open_tree(caldir["read"]) as f:
            r = f["roman"]
            put("read", r["data"], np.float32)
            if "resetnoise" in r:
                put("resetnoise", r["resetnoise"], np.float32)
            if "anc" in r:
                d.c_pink, d.u_pink = float(r["anc"]["C_PINK"]), float(r["anc"]["U_PINK"])
            self.has_amp33 = "amp33" in r
            if self.has_amp33:
                a = r["amp33"]
                put("amp33_med", a["med"], np.float32)
                put("amp33_std", a["std"], np.float32)
                d.has_amp33 = 1
                d.m_pink, d.ru_pink = float(a["M_PINK"]), float(a["RU_PINK"])
                # the reference's own expression for the optimal reference-output coefficient (gen_cal_image.py:542-553)
                cvar = r["anc"]["C_PINK"] ** 2
                self.refout_slope = a["M_PINK"] * cvar / (
                    a["M_PINK"] ** 2 * cvar + a["RU_PINK"] ** 2 + np.median(a["std"]) ** 2 / 128 / np.log(4096)
                )
                d.refout_slope = float(self.refout_slope)
        with open_tree(caldir["dark"]) as f:
            r = f["roman"]
            dk = put("dark_cube", r["data"], np.float32)
            d.n_dark = dk.shape[0]
            put("dark_slope", np.array(r["dark_slope"], dtype=np.float32), np.float32)
            put("dark_dq", r["dq"], np.uint32)
        self.n_dark = d.n_dark
        self.has_bias = "biascorr" in caldir
        self.biascorr_t0 = 0.0
        if self.has_bias:
            with open_tree(caldir["biascorr"]) as f:
                b = put("biascorr", f["roman"]["data"], np.float32)
                d.n_bias = b.shape[0]
                if "t0" in f["roman"]:
                    self.biascorr_t0 = d.biascorr_t0 = float(f["roman"]["t0"])
        with open_tree(caldir["flat"]) as f:
            put("flat", f["roman"]["data"], np.float32)
        with open_tree(caldir["gain"]) as f:
            self.medgain = np.median(f["roman"]["data"])  # gen_cal_image.py:632-633
        _lib.check(_lib.lib().rip_caldir_create(device, C.byref(d), C.byref(self._h)))

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("CalDir already closed")
        return self._h

    def static_products(self):
        """IPC-corrected dark slope, ``get_flat`` product and merged static DQ (host copies; tests/inspection)."""
        n = self.n
        ds, fl, dq = np.empty((n, n), np.float32), np.empty((n, n), np.float32), np.empty((n, n), np.uint32)
        s = C.c_double(0.0)
        _lib.check(_lib.lib().rip_caldir_get_static(self.handle, _lib.ptr(ds), _lib.ptr(fl), _lib.ptr(dq), C.byref(s)))
        return {"dark_slope_ipc": ds, "flat": fl, "static_dq": dq, "refout_slope": s.value}

    def close(self):
        if self._h:
            _lib.lib().rip_caldir_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def _params(cal, meta, config, do_refpix, threads=0, band_rows=0, area=None):
    prm = _lib.L1L2Params()
    prm.G = int(meta["ngrp"])
    prm.exclude_first = 1 if config.get("EXCLUDE_FIRST", True) else 0
    prm.sat_backup = int(config.get("SATURATION_BACKUP", 1))
    prm.do_not_flag_first = 1 if list(meta["read_pattern"][0]) == [0] else 0  # gen_cal_image.py:583
    prm.do_refpix = 1 if do_refpix else 0
    prm.area_dtype = _lib.RIP_F32 if area is None else _lib.float_tag(area)
    prm.threads, prm.band_rows = int(threads), int(band_rows)
    return prm


def calibrate_arrays(cal, data, amp33, read_pattern, frame_time, area_factor=None, config=None, do_refpix=True,
                     want_rdq=False, want_lin_cube=False, want_endslice=None, threads=0, band_rows=0, out=None,
                     dplan=None, sky_step=False):  # fmt: skip
    """
    The numerics of ``calibrateimage`` from L1 arrays to L2 arrays (reference gen_cal_image.py:503-629, 697-709).

    Parameters
    ----------
    cal : CalDir
    data : np.ndarray, uint16 (G, n, n)
        L1 resultants.
    amp33 : np.ndarray, uint16 (G, n, 128), or None
        Reference-output resultants (needed for the reference-pixel correction).
    read_pattern : list of list of int
    frame_time : float
    area_factor : np.ndarray (n, n), float32 or float64, optional
        Pixel area / ``pars.Omega_ideal`` (reference gen_cal_image.py:618-621; a host WCS product).
    config : dict, optional
        The reference's configuration keys ``EXCLUDE_FIRST``, ``SATURATION_BACKUP``, ``RAMP_OPT_PARS``,
        ``JUMP_DETECT_PARS``, ``SLICEOUT``.
    do_refpix : bool
        Run the reference-pixel loop (the reference always does; needs ``amp33`` and a frame side multiple of 128).
    out : dict, optional
        Preallocated output arrays to fill (e.g. page-locked buffers from ``_lib.pinned_empty`` so that the
        device-to-host copies run at full PCIe speed); missing entries are allocated.
    dplan : DevicePlan, optional
        Reuse the pixel-independent plan of a previous call with the same read pattern and configuration.
    sky_step : bool
        Also do the step that follows the hot path in ``calibrateimage`` (reference gen_cal_image.py:639-651): keep
        ``slope_withsky``, and with ``SKYORDER`` in ``config`` fit and subtract the sky model on the active array
        (``utils.sky.medfit`` on the GPU) -> ``skycoefs``, ``skyorder``.

    Returns
    -------
    dict
        ``slope``, ``err_read``, ``err_poisson`` (float32, flat-fielded DN/s, full frame), ``pdq`` (uint32),
        ``endslice`` (int8, active region; if SLICEOUT), ``rdq`` / ``lin_cube`` if asked, ``meta`` (with ``K``).
    """
    config = config or {}
    d = _lib.as_c(data, np.uint16)
    G, n, _ = d.shape
    if n != cal.n:
        raise ValueError(f"L1 frame side {n} does not match the CALDIR ({cal.n})")
    if G >= 128 and config.get("SLICEOUT", False):
        raise ValueError("too many groups")  # gen_cal_image.py:699-700
    if dplan is not None:
        meta, plan, w_exact = dplan.meta, dplan.plan, dplan.w_exact
    else:
        meta = exposure_meta(read_pattern, frame_time)
        plan, w_exact = ramp_setup(meta, config)
    if meta["ngrp"] != G:
        raise ValueError("read pattern and data cube disagree on the number of resultants")
    a33 = None
    if do_refpix:
        if amp33 is None:
            raise ValueError("the reference-pixel correction needs the amp33 cube")
        a33 = _lib.as_c(amp33, np.uint16)
    area = None if area_factor is None else _lib.as_float_plane(area_factor)
    prm = _params(cal, meta, config, do_refpix, threads, band_rows, area)
    if want_endslice is None:
        want_endslice = bool(config.get("SLICEOUT", False))
    out = {} if out is None else out

    def buf(key, shape, dtype):
        a = out.get(key)
        if a is None or a.shape != shape or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
            a = out[key] = np.empty(shape, dtype)
        return a

    buf("slope", (n, n), np.float32)
    buf("err_read", (n, n), np.float32)
    buf("err_poisson", (n, n), np.float32)
    buf("pdq", (n, n), np.uint32)
    o = _lib.L2Out()
    o.slope, o.err_read, o.err_poisson, o.pdq = (out[k].ctypes.data for k in ("slope", "err_read", "err_poisson", "pdq"))
    if want_endslice:
        buf("endslice", (cal.na, cal.na), np.int8)
        o.endslice = out["endslice"].ctypes.data
    if want_rdq:
        buf("rdq", (G, n, n), np.uint8)
        o.rdq = out["rdq"].ctypes.data
    if want_lin_cube:
        buf("lin_cube", (G, n, n), np.float32)
        o.lin_cube = out["lin_cube"].ctypes.data
    _lib.check(
        _lib.lib().rip_l1_to_l2_host(cal.handle, _lib.ptr(d), _lib.ptr(a33), _lib.ptr(area), C.byref(prm),
                                     C.byref(plan), _lib.ptr(w_exact), C.byref(o))
    )  # fmt: skip
    out["meta"] = meta
    if sky_step:
        from ..utils import sky  # noqa: PLC0415

        nb = cal.nb
        out["slope_withsky"] = np.copy(out["slope"])  # version before sky subtraction (gen_cal_image.py:640)
        if "SKYORDER" in config:
            out["skyorder"] = int(config["SKYORDER"])
            act = np.ascontiguousarray(out["slope"][nb:-nb, nb:-nb])
            out["skycoefs"], skymodel = sky.medfit(act, order=out["skyorder"], device=cal.device)
            out["slope"][nb:-nb, nb:-nb] -= skymodel
        else:
            out["skycoefs"], out["skyorder"] = np.array([]).astype(np.float32), -1
    return out


class DevicePlan:
    """Pixel-independent inputs of the device entry point, built once per (read pattern, config)."""

    def __init__(self, cal, read_pattern, frame_time, config=None, do_refpix=True, area_dtype=None, threads=0,
                 band_rows=0):  # fmt: skip
        config = config or {}
        self.meta = exposure_meta(read_pattern, frame_time)
        self.plan, self.w_exact = ramp_setup(self.meta, config)
        self.prm = _params(cal, self.meta, config, do_refpix, threads, band_rows)
        if area_dtype is not None:
            self.prm.area_dtype = _lib.RIP_F64 if np.dtype(area_dtype) == np.float64 else _lib.RIP_F32


def calibrate_device(cal, dplan, d_raw, d_amp33, d_area, d_slope, d_err_read, d_err_poisson, d_pdq, d_endslice=0,
                     d_rdq=0, d_lin_cube=0, stream=0):  # fmt: skip
    """
    ``calibrate_arrays`` on DEVICE pointers (integers, e.g. ``tensor.data_ptr()``), asynchronous on ``stream``
    (``rip_l1_to_l2_dev``).  The caller owns every buffer and synchronises.
    """
    o = _lib.L2Out()
    o.slope, o.err_read, o.err_poisson, o.pdq = d_slope, d_err_read, d_err_poisson, d_pdq
    o.endslice, o.rdq, o.lin_cube = d_endslice or None, d_rdq or None, d_lin_cube or None
    _lib.check(
        _lib.lib().rip_l1_to_l2_dev(cal.handle, C.c_void_p(d_raw), C.c_void_p(d_amp33 or None),
                                    C.c_void_p(d_area or None), C.byref(dplan.prm), C.byref(dplan.plan),
                                    _lib.ptr(dplan.w_exact), C.byref(o), C.c_void_p(stream or None))
    )  # fmt: skip


def prefetch_refpix_device(cal, d_raw_next, d_amp33_next, ngrp):
    """
    Start the reference-pixel statistics of the NEXT exposure (device pointers; the cubes must already be complete) on
    the handle's side stream, so that they run beside the fused kernel of the current one; the ``calibrate_device`` call
    with the same ``d_raw`` then only waits for them (``rip_caldir_prefetch_refpix``).  Same results either way.
    """
    _lib.check(_lib.lib().rip_caldir_prefetch_refpix(cal.handle, C.c_void_p(d_raw_next), C.c_void_p(d_amp33_next), int(ngrp)))


class Pipeline:
    """
    A stream of exposures of one SCA through the GPU with the PCIe copies overlapped (``rip_pipeline_*``): ``depth``
    exposures are in flight on three CUDA streams (host->device | reference-pixel statistics + fused kernel |
    device->host).  Same arithmetic as ``calibrate_arrays``.

    >>> pipe = Pipeline(cal, read_pattern, frame_time, config)            # doctest: +SKIP
    >>> tickets = [pipe.submit(data, amp33, area) for data, amp33, area in exposures]
    >>> results = [pipe.result(t) for t in tickets]

    For real overlap the host arrays must be page-locked (``_lib.pinned_empty``); input arrays must not be modified
    until ``result()`` of their ticket has returned.
    """

    def __init__(self, cal, read_pattern, frame_time, config=None, do_refpix=True, depth=3, want_rdq=False,
                 want_endslice=None, area_dtype=np.float32):  # fmt: skip
        config = config or {}
        self.cal, self.config, self.do_refpix = cal, config, do_refpix
        self.want_rdq = bool(want_rdq)
        self.want_endslice = bool(config.get("SLICEOUT", False)) if want_endslice is None else bool(want_endslice)
        self.dplan = DevicePlan(cal, read_pattern, frame_time, config, do_refpix, area_dtype)
        self.G = int(self.dplan.meta["ngrp"])
        if self.G >= 128 and config.get("SLICEOUT", False):
            raise ValueError("too many groups")  # gen_cal_image.py:699-700
        self._p = C.c_void_p()
        _lib.check(_lib.lib().rip_pipeline_create(cal.handle, self.G, int(depth), int(self.want_endslice),
                                                  int(self.want_rdq), C.byref(self._p)))  # fmt: skip
        self._pending = {}

    def set_area(self, area_factor):
        """Keep ``area_factor`` (float32 / float64 (n, n), or ``None`` to drop it) resident on the device: exposures
        submitted with ``area_factor=None`` are divided by it.  For streams of exposures that share the plane (the pixel
        area in detector coordinates is set by the optical distortion of the SCA): one 67 MB upload instead of one per
        exposure."""
        if area_factor is None:
            _lib.check(_lib.lib().rip_pipeline_set_area(self._p, None, _lib.RIP_F32))
            return
        area = _lib.as_float_plane(area_factor)
        if area.shape != (self.cal.n, self.cal.n):
            raise ValueError(f"area plane must be ({self.cal.n},{self.cal.n})")
        _lib.check(_lib.lib().rip_pipeline_set_area(self._p, _lib.ptr(area), _lib.float_tag(area)))

    def set_area_wcs(self, wcs, dtype=np.float32):
        """Compute the AreaFactor plane of the next exposures ON THE DEVICE from their WCS (``utils.coordutils.FitsWCS``,
        header text or dict; reference gen_cal_image.py:618-621): pixel solid angle / ``pars.Omega_ideal`` at the n x n
        pixel positions of the full frame, written into the pipeline's resident area buffer (``rip_pipeline_set_area_wcs``).
        Asynchronous on the compute stream: exposures already submitted keep their plane."""
        from ..utils import coordutils  # noqa: PLC0415

        w = wcs if isinstance(wcs, coordutils.FitsWCS) else coordutils.FitsWCS(wcs)
        v = w.pack()
        tag = _lib.RIP_F64 if np.dtype(dtype) == np.float64 else _lib.RIP_F32
        _lib.check(_lib.lib().rip_pipeline_set_area_wcs(self._p, _lib.ptr(v), int(v.size), 1.0 / pars.Omega_ideal, tag))

    def submit(self, data, amp33, area_factor=None, out=None):
        """Queue one exposure; returns a ticket.  ``out``: optional dict of preallocated (pinned) output arrays."""
        cal, n, G = self.cal, self.cal.n, self.G
        d = _lib.as_c(data, np.uint16)
        if d.shape != (G, n, n):
            raise ValueError(f"expected a ({G},{n},{n}) uint16 cube, got {d.shape}")
        a33 = None
        if self.do_refpix:
            if amp33 is None:
                raise ValueError("the reference-pixel correction needs the amp33 cube")
            a33 = _lib.as_c(amp33, np.uint16)
        area = None if area_factor is None else _lib.as_float_plane(area_factor)
        prm = _lib.L1L2Params()
        C.memmove(C.byref(prm), C.byref(self.dplan.prm), C.sizeof(prm))
        prm.area_dtype = _lib.RIP_F32 if area is None else _lib.float_tag(area)
        out = {} if out is None else out

        def buf(key, shape, dtype):
            a = out.get(key)
            if a is None or a.shape != shape or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
                a = out[key] = np.empty(shape, dtype)
            return a

        o = _lib.L2Out()
        o.slope = buf("slope", (n, n), np.float32).ctypes.data
        o.err_read = buf("err_read", (n, n), np.float32).ctypes.data
        o.err_poisson = buf("err_poisson", (n, n), np.float32).ctypes.data
        o.pdq = buf("pdq", (n, n), np.uint32).ctypes.data
        if self.want_endslice:
            o.endslice = buf("endslice", (cal.na, cal.na), np.int8).ctypes.data
        if self.want_rdq:
            o.rdq = buf("rdq", (G, n, n), np.uint8).ctypes.data
        t = C.c_long(-1)
        _lib.check(_lib.lib().rip_pipeline_submit(self._p, _lib.ptr(d), _lib.ptr(a33), _lib.ptr(area), C.byref(prm),
                                                  C.byref(self.dplan.plan), _lib.ptr(self.dplan.w_exact), C.byref(o),
                                                  C.byref(t)))  # fmt: skip
        out["meta"] = self.dplan.meta
        self._pending[t.value] = (out, d, a33, area)  # keep the buffers alive until the copies are done
        return t.value

    def result(self, ticket):
        """Block until the exposure of ``ticket`` is complete; returns its output dict."""
        _lib.check(_lib.lib().rip_pipeline_wait(self._p, C.c_long(ticket)))
        return self._pending.pop(ticket)[0]

    def close(self):
        if self._p:
            _lib.lib().rip_pipeline_destroy(self._p)
            self._p = C.c_void_p()
            self._pending.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


_CAL_CACHE = {}


def _cached_caldir(caldir, device):
    """One resident CalDir (and one pipeline per group count) per (device, CALDIR file set): the reference re-opens the
    calibration files for every exposure (gen_cal_image.py:112,172,213,531,536,560,632); here they are uploaded once."""
    key = (device, tuple(sorted((k, v if isinstance(v, str) else id(v)) for k, v in caldir.items())))
    if key not in _CAL_CACHE:
        _CAL_CACHE[key] = {"cal": CalDir(caldir, device), "pipes": {}}
    return _CAL_CACHE[key]


def clear_caldir_cache():
    """Release the resident CALDIRs of ``calibrateimage`` (device memory)."""
    for ent in _CAL_CACHE.values():
        for pipe in ent["pipes"].values():
            pipe.close()
        ent["cal"].close()
    _CAL_CACHE.clear()


class ProcessLog:
    """String log of the processing steps (reference utils/processlog.py:12-56; stored as ``processinfo["log"]``)."""

    def __init__(self):
        self.output = ""
        self.reffiles = {}

    def append(self, newoutput):
        self.output += newoutput


def read_l1(path_or_tree):
    """The L1 inputs of ``initializationstep`` (gen_cal_image.py:116-126): data u16 [G,n,n], amp33 u16 [G,n,128] or None,
    read pattern, frame time and the ``meta`` branch; ``reference_read`` + ``data_encoding_offset`` are folded back into
    the data as ``do_dqinit`` does for EXTRACT_REF files (from_sim/sim_to_isim.py:711-730; restated)."""
    with open_tree(path_or_tree) as f:
        r = f["roman"]
        data = np.ascontiguousarray(np.asarray(r["data"]), dtype=np.uint16)
        amp33 = np.ascontiguousarray(np.asarray(r["amp33"]), dtype=np.uint16) if "amp33" in r else None
        meta = r["meta"]
        read_pattern = [[int(x) for x in g] for g in meta["exposure"]["read_pattern"]]
        frame_time = float(meta["exposure"]["frame_time"])
        if "reference_read" in r:
            inst = meta["instrument"] if "instrument" in meta else {}
            off = int(inst["data_encoding_offset"]) if "data_encoding_offset" in inst else 0
            data = (data.astype(np.int32) + np.asarray(r["reference_read"]).astype(np.int32) - off).astype(np.uint16)
        border = {
            "amp33": None if amp33 is None else np.copy(amp33),
            "border_ref_pix_left": data[:, :, :4].astype(np.float32),
            "border_ref_pix_right": data[:, :, -4:].astype(np.float32),
            "border_ref_pix_top": data[:, -4:, :].astype(np.float32),
            "border_ref_pix_bottom": data[:, :4, :].astype(np.float32),
        }  # oututils.add_in_ref_data (L1_to_L2/oututils.py:43-49)
        meta_copy = _plain_copy(meta)
    return data, amp33, read_pattern, frame_time, meta_copy, border


def _plain_copy(node):
    """Deep copy of a metadata branch that outlives its file (arrays materialised, tagged nodes kept)."""
    if isinstance(node, dict) or hasattr(node, "items"):
        out = type(node)() if isinstance(node, dict) else {}
        if hasattr(node, "tag") and hasattr(out, "tag"):
            out.tag = node.tag
        for k, v in node.items():
            out[k] = _plain_copy(v)
        return out
    if isinstance(node, (list, tuple)):
        out = type(node)(_plain_copy(v) for v in node) if isinstance(node, list) else [_plain_copy(v) for v in node]
        if hasattr(node, "tag") and hasattr(out, "tag"):
            out.tag = node.tag
        return out
    if hasattr(node, "__array__") and not isinstance(node, (str, bytes)):
        return np.array(node)
    return node


def calibrateimage(config, verbose=True, device=0):
    """
    Main routine to run the specified calibrations from a config file: the reference's entry point
    (L1_to_L2/gen_cal_image.py:480-739) with the same configuration keys.  Reads ``config["IN"]`` (L1 ASDF), writes
    ``config["OUT"]`` (L2 ASDF: ``roman`` + ``processinfo`` trees, arrays as the reference packs them at :653-723).

    What runs where: the per-pixel chain (dq-init ... endslice, :503-629, 697-709) is the fused CUDA path; the pixel
    area comes from the ``FITSWCS`` header on the device (``utils.coordutils``); mask growth, binning, sky mode and the
    ``SKYORDER`` fit (:639-651) are device kernels; reading, packaging and writing are host code here
    (``asdf`` if installed, else ``io.asdf_lite``).

    Not implemented (raise ``NotImplementedError`` instead of silently differing from the reference): ``dark_decay`` in
    CALDIR (:567-570), ``correct_wfi18_transient`` (:572-575), ``romancal_ramp_fit`` (:415-432).  ``FITSOUT`` (:725-736) is
    written with ``io/fits_lite.py``.
    Metadata of the L1 file is passed through; romanisim's ``make_asdf`` bookkeeping (photometry, cal_step, WCS object)
    is not reproduced: the FITS header text is stored under ``processinfo["fitswcs"]`` instead.
    """
    from ..caltree import write_tree  # noqa: PLC0415
    from ..utils import coordutils, maskhandling, sky  # noqa: PLC0415

    caldir = config["CALDIR"]
    if "dark_decay" in caldir:
        raise NotImplementedError("CALDIR['dark_decay'] (romancal dark-decay step, gen_cal_image.py:567-570) is not implemented on the GPU path")
    for key in ("correct_wfi18_transient", "romancal_ramp_fit"):
        if config.get(key, False):
            raise NotImplementedError(f"config['{key}'] is not implemented on the GPU path (reference gen_cal_image.py)")
    mylog = ProcessLog()
    wcs = coordutils.wcs_from_config(config)
    if wcs is None:
        raise ValueError("Unrecognized WCS")  # (the reference cannot package an exposure without FITSWCS either)
    ent = _cached_caldir(caldir, device)
    cal = ent["cal"]
    data, amp33, read_pattern, frame_time, l1meta, border = read_l1(config["IN"])
    mylog.append("Initialized data\n")
    G, n, nb = data.shape[0], data.shape[1], cal.nb
    do_refpix = cal.has_amp33 and amp33 is not None
    pkey = (G, tuple(tuple(g) for g in read_pattern), frame_time, repr(sorted((k, repr(v)) for k, v in config.items() if k in
            ("EXCLUDE_FIRST", "SATURATION_BACKUP", "RAMP_OPT_PARS", "JUMP_DETECT_PARS", "SLICEOUT"))), do_refpix)  # fmt: skip
    if pkey not in ent["pipes"]:
        ent["pipes"][pkey] = Pipeline(cal, read_pattern, frame_time, config, do_refpix=do_refpix, depth=1, want_rdq=True,
                                      area_dtype=np.float64)  # fmt: skip
    pipe = ent["pipes"][pkey]
    # AreaFactor = pixel area / Omega_ideal at the n x n pixel positions (gen_cal_image.py:618-621), float64 on the device
    pipe.set_area_wcs(wcs, dtype=np.float64)
    out = pipe.result(pipe.submit(data, amp33 if do_refpix else None, None))
    meta = out["meta"]
    mylog.append("Saturation check complete\n" + ("Reference pixel correction complete\n" if do_refpix else "") +
                 "Linearity correction complete\nRamp fitting complete\nDark current subtracted\nacquired flat field\n")  # fmt: skip
    slope, pdq, rdq = out["slope"], out["pdq"], out["rdq"]
    medgain = cal.medgain
    mylog.append(f"median gain = {medgain:8.5f} e/DN\n")
    # sky information (gen_cal_image.py:639-651)
    slope_withsky = np.copy(slope)
    m = maskhandling.PixelMask1.build(pdq, device=device)
    medsky, _ = sky.smooth_mode(sky.binkxk(slope, 4, mask=m, device=device), device=device)
    if "SKYORDER" in config:
        skyorder = int(config["SKYORDER"])
        skycoefs, skymodel = sky.medfit(np.ascontiguousarray(slope[nb:-nb, nb:-nb]), order=skyorder, device=device)
        slope[nb:-nb, nb:-nb] -= skymodel
    else:
        skycoefs, skyorder = np.array([]).astype(np.float32), -1
    # packaging (rimage.make_asdf arrays :653-666, oututils.add_in_ref_data :674, typefix dummy fields)
    act = np.s_[nb:-nb, nb:-nb]
    var_r = out["err_read"][act] ** 2
    var_p = out["err_poisson"][act] ** 2
    im2 = {
        "meta": l1meta,
        "data": np.ascontiguousarray(slope[act]),
        "dq": np.ascontiguousarray(pdq[act]),
        "var_poisson": var_p,
        "var_rnoise": var_r,
        "var_flat": np.zeros_like(var_r),
        "err": np.sqrt(var_r + var_p),
        "dq_border_ref_pix_left": np.copy(pdq[:, :4]), "dq_border_ref_pix_right": np.copy(pdq[:, -4:]),
        "dq_border_ref_pix_top": np.copy(pdq[-4:, :]), "dq_border_ref_pix_bottom": np.copy(pdq[:4, :]),
        "data_withsky": np.ascontiguousarray(slope_withsky[act]),
    }  # fmt: skip
    for k, v in border.items():
        if v is not None:
            im2[k] = v
    for fld in ("chisq", "dumo"):  # utils/typefix.py:22-29
        im2[fld] = np.zeros(im2["data"].shape, dtype=np.float16)
    im2["meta"].setdefault("dummyfields", [])
    im2["meta"]["dummyfields"] = list(im2["meta"]["dummyfields"]) + ["roman.chisq", "roman.dumo"]
    im2["meta"]["calibration_software_name"] = "gen_cal_image / HLWAS PIT (romanimpreprocess_b200)"  # oututils.py:102-106
    from .. import __version__ as _ver  # noqa: PLC0415

    im2["meta"]["calibration_software_version"] = str(_ver)
    if "exposure" in im2["meta"]:
        im2["meta"]["exposure"]["read_pattern"] = [list(g) for g in read_pattern]
    processinfo = {
        "medsky": float(medsky), "medgain": float(medgain), "skyorder": skyorder, "skycoefs": np.asarray(skycoefs),
        "ramp_opt_pars": {k: float(v) for k, v in meta["ramp_opt_pars"].items()},
        "meta": {"frame_time": float(frame_time), "read_pattern": [list(g) for g in read_pattern], "ngrp": int(G),
                 "tbar": meta["tbar"], "tau": meta["tau"], "N": meta["N"], "nborder": int(nb), "K": meta["K"]},
        "weights": meta["K"], "config": _plain_copy(config), "log": mylog.output,
        "exclude_first": bool(config.get("EXCLUDE_FIRST", True)),
        "fitswcs": open(config["FITSWCS"]).read(),
    }  # fmt: skip
    if config.get("SLICEOUT", False):
        processinfo["endslice"] = out["endslice"]
    write_tree(config["OUT"], {"roman": im2, "processinfo": processinfo})
    if config.get("FITSOUT", False):  # (gen_cal_image.py:725-736; saturated pixels are accepted in this step, as there)
        from ..io import fits_lite  # noqa: PLC0415

        good = ~maskhandling.PixelMask1.build(im2["dq"], device=device)
        fits_lite.write_hdus(config["OUT"][:-5] + "_asdf_to.fits",
                             [(im2["data"], None), (im2["dq"], None), (np.where(good, im2["data"], -1000).astype(np.float32), None)])  # fmt: skip
    if verbose:
        print(mylog.output)


__all__ = ["CalDir", "DevicePlan", "Pipeline", "calibrate_arrays", "calibrate_device", "calibrateimage", "exposure_meta",
           "ramp_setup", "pixel"]  # fmt: skip
