"""Ramp fitting on the GPU: drop-in for the reference's ``romanimpreprocess.utils.fitting``.

Functions
---------
construct_weights
    Weight vector for slope fitting (host: a (G-1)x(G-1) float64 inverse; reference utils/fitting.py:20-86).
build_plan
    Every pixel-independent scalar of ``jump_detect`` for all truncations, evaluated with the reference's own
    NumPy expressions so that each carries the reference's rounding (consumed by the CUDA kernels).
jump_detect
    Slope, errors, jump significance and JUMP_DET flags (reference utils/fitting.py:89-255).
ramp_fit
    Full fit + saturation-truncated refits + flag propagation (reference utils/fitting.py:258-355).

Same signatures and in-place semantics as the reference (``rdq``, ``pdq`` are updated in place).  ``caldir``
entries may be ASDF file names or in-memory trees (see ``caltree``).
"""

import ctypes as C

import numpy as np

from .. import _lib
from ..caltree import open_tree

DEFAULT_BAND = 1.0e-4  # relative half-width of the band in which the fast kernel re-evaluates a slice exactly


def construct_weights(u, meta, exclude_first=True):
    """
    Makes a numpy array of weights for slope fitting (Casertano et al. 2022, fixed u).

    Parameters
    ----------
    u : float
        Poisson to read noise ratio, unit 1/(e*s).
    meta : dict
        Must contain ``ngrp``, ``N``, ``tbar``, ``tau``.
    exclude_first : bool, optional
        Give zero weight to the first (reset-read) resultant.

    Returns
    -------
    K : np.array of float32, length ``meta['ngrp']``.
    """
    ngrp_all = meta["ngrp"]
    first = 1 if exclude_first else 0
    m = ngrp_all - first
    tbar = meta["tbar"][first:].astype(np.float64)
    tau = meta["tau"][first:].astype(np.float64)
    cov = np.zeros((m, m))
    for a in range(m):
        cov[a, a] = 1.0 / meta["N"][first + a] + u * tau[a]
        for b in range(a):
            cov[a, b] = cov[b, a] = u * tbar[b]
    w = np.linalg.inv(cov)
    w_col = np.sum(w, axis=0)
    w_t = w @ tbar
    f0 = np.sum(w)
    f1 = np.sum(w_t)
    f2 = np.dot(tbar, w_t)
    det = f0 * f2 - f1**2
    out = np.zeros(ngrp_all)
    out[first:] = (f0 * w_t - f1 * w_col) / det
    return out.astype(np.float32)


def _variant_scalars(meta, K_full, start, ngrp):
    """K, coef, rfac and the slice list of one jump_detect call (reference utils/fitting.py:162-169,196-241)."""
    if ngrp == meta["ngrp"]:
        K = K_full
    else:
        K = np.zeros(ngrp, dtype=np.float32)
        K[-1] = 1.0 / (meta["tbar"][ngrp - 1] - meta["tbar"][start])
        K[start] = -K[-1]
    coef = 0.0
    for i in range(start, ngrp):
        coef += K[i] ** 2 * meta["tau"][i]
        for j in range(start, i):
            coef += 2.0 * K[i] * K[j] * meta["tbar"][j]
    rfac = np.sqrt(np.sum(K**2 / np.array(meta["N"][:ngrp])))
    slices = []
    for i in range(start, ngrp - 1):
        dimax = 2
        if i == ngrp - 2 or ngrp - 1 - start == 2:
            dimax = 1
        for di in range(1, 1 + dimax):
            dt = meta["tbar"][i + di] - meta["tbar"][i]
            w = np.zeros(ngrp)
            w[i + di] = 1.0 / dt
            w[i] = -1.0 / dt
            w -= K
            # factorised variance for the fast path: var = dvardt*A + sig2read*B  
            A = 0.0
            B = 0.0
            for a in range(ngrp):
                A += w[a] ** 2 * float(meta["tau"][a])
                B += w[a] ** 2 / float(meta["N"][a])
                for b in range(a):
                    A += 2 * w[a] * w[b] * float(meta["tbar"][b])
            slices.append((i, di, dt, w, A, B))
    return K, coef, rfac, slices


def build_plan(meta, exclude_first=True, band=DEFAULT_BAND):
    """Pack the ramp plan for the C ABI.  Returns ``(RampPlan, w_exact float64 [nslices, RIP_GMAX])``."""
    G = int(meta["ngrp"])
    if G < 3 or G > _lib.RIP_GMAX:
        raise ValueError(f"the GPU ramp fitter supports 3..{_lib.RIP_GMAX} groups, got {G}")
    start = 1 if exclude_first else 0
    K_full = np.asarray(meta["K"], dtype=np.float32)
    plan = _lib.RampPlan()
    plan.G, plan.start = G, start
    for g in range(G):
        plan.tbar[g] = meta["tbar"][g]
        plan.tau[g] = meta["tau"][g]
        plan.nreads[g] = float(meta["N"][g])
    SthreshA, SthreshB, IthreshA, IthreshB = 5.5, 4.5, 1.0, 1000.0
    jp = meta.get("jump_detect_pars", {})
    if "SthreshA" in jp:
        SthreshA = float(jp["SthreshA"])
    if "SthreshB" in jp:
        SthreshB = float(jp["SthreshB"])
    if "IthreshA" in jp:
        IthreshA = float(jp["IthreshA"])
    if "IthreshB" in jp:
        IthreshB = float(jp["IthreshB"])
    plan.IthreshA_f, plan.IthreshB_f = IthreshA, IthreshB
    plan.SthreshA, plan.SthreshB = SthreshA, SthreshB
    plan.logIratio = float(np.log(IthreshB / IthreshA))
    plan.band = band
    plan.thrA_f = SthreshA
    plan.thrK_f = (SthreshB - SthreshA) / plan.logIratio
    plan.invIA_f = 1.0 / IthreshA
    variants = [G] + list(range(G - 1, 2 + start, -1))  # full ramp, then iend = G-1 ... start+3
    if len(variants) > _lib.RIP_MAXVAR:
        raise ValueError("too many truncation variants")
    plan.nvar = len(variants)
    w_rows = []
    off = 0
    for v, ngrp in enumerate(variants):
        K, coef, rfac, slices = _variant_scalars(meta, K_full, start, ngrp)
        plan.var_ngrp[v] = ngrp
        for g in range(ngrp):
            plan.var_K[v][g] = K[g]
        plan.var_coef[v] = coef
        plan.var_rfac[v] = rfac
        plan.var_slice_off[v] = off
        for i, di, dt, w, A, B in slices:
            if off >= _lib.RIP_MAXSLICE:
                raise ValueError("too many jump-detection slices for the plan")
            s = plan.slices[off]
            if not (A >= 0.0 and B >= 0.0):  # cannot happen for a valid read pattern; keeps the squared test sound
                plan.band = float("inf")
            s.i, s.di, s.dt, s.inv_dt, s.A, s.B = i, di, dt, 1.0 / float(dt), A, B
            row = np.zeros(_lib.RIP_GMAX)
            row[:ngrp] = w
            w_rows.append(row)
            off += 1
    plan.var_slice_off[len(variants)] = off
    w_exact = np.ascontiguousarray(np.array(w_rows, dtype=np.float64).reshape(-1, _lib.RIP_GMAX))
    if w_exact.size == 0:
        w_exact = np.zeros((1, _lib.RIP_GMAX))
    return plan, w_exact


def _variant_index(meta, exclude_first, truncate_ramp):
    G = meta["ngrp"]
    if truncate_ramp is None or truncate_ramp == G:
        return 0
    start = 1 if exclude_first else 0
    if not (2 + start < truncate_ramp < G):
        raise ValueError(f"truncate_ramp={truncate_ramp} outside {3 + start}..{G - 1}")
    return G - truncate_ramp


def _gain_read(caldir):
    with open_tree(caldir["gain"]) as f:
        gain = _lib.as_float_plane(f["roman"]["data"])
    with open_tree(caldir["read"]) as f:
        read = _lib.as_c(f["roman"]["data"], np.float32)
    return gain, read


def jump_detect(data, rdq, pdq, meta, caldir, mylog, exclude_first=True, truncate_ramp=None, device=0):
    """
    Searches for a jump (affected pixels are flagged in ``rdq``, not corrected).

    Same contract as the reference's ``jump_detect``; returns ``slope, slope_err_read, slope_err_poisson, smap``.
    """
    ngrp_all = int(meta["ngrp"])
    ny, nx = np.shape(pdq)
    plan, w_exact = build_plan(meta, exclude_first)
    v = _variant_index(meta, exclude_first, truncate_ramp)
    nsl = plan.var_slice_off[v + 1] - plan.var_slice_off[v]
    gain, read = _gain_read(caldir)
    d = _lib.as_c(data[:ngrp_all], np.float32)
    q = _lib.as_c(rdq, np.uint8)
    slope = np.empty((ny, nx), np.float32)
    er = np.empty((ny, nx), np.float32)
    ep = np.empty((ny, nx), np.float32)
    smap = np.zeros((max(nsl, 0), ny, nx), np.float32)
    _lib.check(
        _lib.lib().rip_jump_detect(device, _lib.ptr(d), _lib.ptr(q), ny, nx, int(meta["nborder"]), C.byref(plan),
                                   _lib.ptr(w_exact), v, _lib.ptr(gain), _lib.float_tag(gain), _lib.ptr(read),
                                   _lib.ptr(slope), _lib.ptr(er), _lib.ptr(ep), _lib.ptr(smap) if nsl > 0 else None)
    )  # fmt: skip
    if q is not rdq:
        rdq[...] = q
    if mylog is not None:
        mylog.append(f"truncate at {truncate_ramp}, K = {np.array(plan.var_K[v][: plan.var_ngrp[v]])}\n")
    return slope, er, ep, smap


def ramp_fit(data, rdq, pdq, meta, caldir, mylog, exclude_first=True, device=0, fast=True):
    """
    Ramp fitting with saturation-truncated refits and flag propagation; ``rdq`` and ``pdq`` updated in place.

    Same contract as the reference's ``ramp_fit``; returns ``slope, slope_err_read, slope_err_poisson``.
    ``fast=False`` evaluates every jump significance in the reference's exact op order (slower, same flags).
    """
    ny, nx = np.shape(pdq)
    G = int(meta["ngrp"])
    plan, w_exact = build_plan(meta, exclude_first)
    gain, read = _gain_read(caldir)
    d = _lib.as_c(data[:G], np.float32)
    q = _lib.as_c(rdq, np.uint8)
    p = _lib.as_c(pdq, np.uint32)
    slope = np.empty((ny, nx), np.float32)
    er = np.empty((ny, nx), np.float32)
    ep = np.empty((ny, nx), np.float32)
    _lib.check(
        _lib.lib().rip_ramp_fit(device, _lib.ptr(d), _lib.ptr(q), _lib.ptr(p), ny, nx, int(meta["nborder"]),
                                C.byref(plan), _lib.ptr(w_exact), _lib.ptr(gain), _lib.float_tag(gain),
                                _lib.ptr(read), 1 if fast else 0, _lib.ptr(slope), _lib.ptr(er), _lib.ptr(ep))
    )  # fmt: skip
    if q is not rdq:
        rdq[...] = q
    if p is not pdq:
        pdq[...] = p
    if mylog is not None:
        mylog.append(f"ramp fit on device {device}: {G} groups, {plan.nvar - 1} truncated variants\n")
    return slope, er, ep
