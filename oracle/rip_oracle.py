"""CPU ORACLE for the L1->L2 calibration hot path and the forward ramp model.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm.  It exists only so that the CUDA path can be
checked; nothing under ``romanimpreprocess_b200/`` may import it (only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do).

Every function cites the reference lines it follows (paths relative to the reference checkout).  The op order
and dtypes follow SURVEY.md Appendix A so that, run under the same NumPy, results are bit-identical to the
reference's own functions; this is *pinned*: ``tests/golden/make_golden.py`` imports the unmodified reference
modules (with ``asdf`` / ``roman_datamodels`` stubbed) and stores their outputs, ``tests/test_oracle_golden.py``
compares.  Calibration data are passed as arrays / in-memory trees instead of ASDF file names.

PARITY UNPINNED for the third-party steps whose source is not in the reference tree (romancal ``do_dqinit``,
``flag_saturation`` -> stcal ``flag_saturated_pixels``, ``subtract_dark_current``, ``_create_image_model``,
romanisim ``apportion_counts_to_resultants`` / ``add_read_noise_to_resultants``): functions ``dq_init``,
``flag_saturation``, ``subtract_dark_current`` and ``make_l1_fullcal`` (apportioning + read noise) below restate the
behaviour recorded in SURVEY.md Appendix D (from the reference's call sites and docs); the reference holds no
golden vector for them.
"""

import numpy as np

DO_NOT_USE = np.uint32(1)
SATURATED = np.uint32(2)
JUMP_DET = np.uint32(4)
AD_FLOOR = np.uint32(64)
GW_AFFECTED_DATA = np.uint32(16)
NO_FLAT_FIELD = np.uint32(2**18)
NO_GAIN_VALUE = np.uint32(2**19)
NO_LIN_CORR = np.uint32(2**20)
NO_SAT_CHECK = np.uint32(2**21)
REFERENCE_PIXEL = np.uint32(2**31)

NSIDE = 4096
CHANNELWIDTH = 128

# ---------------------------------------------------------------------------------------------------------
# IPC   (reference src/romanimpreprocess/utils/ipc_linearity.py:37-186)
# ---------------------------------------------------------------------------------------------------------


def ipc_fwd(image, kernel, gain=None):
    """9-tap source-indexed IPC convolution; accumulation order of ipc_linearity.py:69-99."""
    im = image if gain is None else gain * image
    out = im * kernel[1, 1]
    out[1:, :] += im[:-1, :] * kernel[2, 1, :-1, :]
    out[:-1, :] += im[1:, :] * kernel[0, 1, 1:, :]
    out[:, 1:] += im[:, :-1] * kernel[1, 2, :, :-1]
    out[:, :-1] += im[:, 1:] * kernel[1, 0, :, 1:]
    out[1:, 1:] += im[:-1, :-1] * kernel[2, 2, :-1, :-1]
    out[1:, :-1] += im[:-1, 1:] * kernel[2, 0, :-1, 1:]
    out[:-1, 1:] += im[1:, :-1] * kernel[0, 2, 1:, :-1]
    out[:-1, :-1] += im[1:, 1:] * kernel[0, 0, 1:, 1:]
    if gain is not None:
        out /= gain
    return out


def ipc_rev(image, kernel, order=2, gain=None):
    """Iterative deconvolution out <- out + image - K*out (ipc_linearity.py:134-142)."""
    im2 = image if gain is None else gain * image
    out = np.copy(im2)
    for _ in range(order):
        out = out + im2 - ipc_fwd(out, kernel)
    if gain is not None:
        out /= gain
    return out


def correct_cube(data, kernel, gain_full=None):
    """In-place IPC correction of the active region of every group (ipc_linearity.py:176-186)."""
    ngrp, ny, nx = data.shape
    nb = (8192 + (nx - kernel.shape[-1]) // 2) % 16
    g = 1.0 if gain_full is None else np.copy(gain_full[nb : ny - nb, nb : nx - nb])
    for i in range(ngrp):
        data[i, nb : ny - nb, nb : nx - nb] = ipc_rev(data[i, nb : ny - nb, nb : nx - nb] * g, kernel) / g


# ---------------------------------------------------------------------------------------------------------
# Legendre linearity (ipc_linearity.py:192-392)
# ---------------------------------------------------------------------------------------------------------


def lin_eval(z, coefs, linextrap=True):
    """phi = sum_L c_L P_L(z), linear extrapolation for |z|>1 (ipc_linearity.py:215-231)."""
    ex = np.abs(z) > 1
    phi = np.copy(coefs[0])
    prev = np.ones_like(phi)
    cur = np.copy(z)
    for L in range(1, coefs.shape[0]):
        if linextrap:
            phi += coefs[L] * np.where(ex, np.sign(z) ** L * (1 + L * (L + 1) / 2.0 * (np.abs(z) - 1)), cur)
        else:
            phi += coefs[L] * cur
        nxt = (2 * L + 1) / (L + 1) * z * cur - L / (L + 1) * prev
        prev = cur
        cur = nxt
    return phi, ex


def linearity(S, lin, origin=(0, 0)):
    """Single-frame linearity (ipc_linearity.py:261-273); ``lin`` is the ``roman`` branch of the lin file."""
    dy, dx = S.shape
    y0, x0 = origin[1], origin[0]
    Smin = lin["Smin"][y0 : y0 + dy, x0 : x0 + dx]
    Smax = lin["Smax"][y0 : y0 + dy, x0 : x0 + dx]
    phi, ex = lin_eval(-1 + 2 * (S - Smin) / (Smax - Smin), lin["data"][:, y0 : y0 + dy, x0 : x0 + dx])
    dq = np.copy(lin["dq"][y0 : y0 + dy, x0 : x0 + dx])
    dq |= np.where(ex, NO_LIN_CORR, 0).astype(np.uint32)
    return phi, dq


def multilin(S, lin, origin=(0, 0), do_not_flag_first=True, attempt_corr=None):
    """Multi-group linearity with DQ (ipc_linearity.py:313-344)."""
    ngrp, dy, dx = S.shape
    y0, x0 = origin[1], origin[0]
    if attempt_corr is None:
        attempt_corr = np.ones((ngrp, dy, dx), dtype=bool)
    phi = np.zeros(S.shape, dtype=np.float32)
    Smin = lin["Smin"][y0 : y0 + dy, x0 : x0 + dx]
    Smax = lin["Smax"][y0 : y0 + dy, x0 : x0 + dx]
    Sref = lin["Sref"][y0 : y0 + dy, x0 : x0 + dx]
    coefs = lin["data"][:, y0 : y0 + dy, x0 : x0 + dx]
    dq = np.copy(lin["dq"][y0 : y0 + dy, x0 : x0 + dx])
    for j in range(ngrp):
        z = -1 + 2 * (S[j] - Smin) / (Smax - Smin)
        if j == 0 and do_not_flag_first:
            z = np.clip(z, -1, 1)
        p, ex = lin_eval(z, coefs)
        phi[j] = p
        phi[j] = np.where(dq & (NO_LIN_CORR | REFERENCE_PIXEL) == 0, phi[j], S[j] - Sref)
        if not (j == 0 and do_not_flag_first):
            dq |= np.where(np.logical_and(ex, attempt_corr[j]), NO_LIN_CORR, 0).astype(np.uint32)
    return phi, dq


def invlinearity(Slin, lin, origin=(0, 0)):
    """24-step bisection inverse of the Legendre map (ipc_linearity.py:374-392)."""
    dy, dx = Slin.shape
    y0, x0 = origin[1], origin[0]
    coefs = lin["data"][:, y0 : y0 + dy, x0 : x0 + dx]
    z = np.zeros_like(Slin)
    for j in range(1, 25):
        phi, ex = lin_eval(z, coefs, linextrap=False)
        z += np.where(phi < Slin, 1 / 2**j, -1 / 2**j)
    Smin = lin["Smin"][y0 : y0 + dy, x0 : x0 + dx]
    Smax = lin["Smax"][y0 : y0 + dy, x0 : x0 + dx]
    return Smin + (Smax - Smin) / 2.0 * (1 + z), ex


def il_apply(counts, lin, gain_full, ipc_kernel, start_e=0.0, electrons=True, electrons_out=False):
    """IL.apply: IPC -> /gain -> inverse linearity (ipc_linearity.py:461-513)."""
    conv = ipc_fwd(counts + start_e, ipc_kernel) if ipc_kernel is not None else counts + start_e
    nyc, nxc = counts.shape
    g_in = 1.0
    g_out = 1.0
    if electrons or electrons_out:
        g = gain_full
        if g.shape[0] > nyc:
            b = (g.shape[0] - nyc) // 2
            g = g[b:-b, b:-b]
        if electrons:
            g_in = g
        if electrons_out:
            g_out = g
    nb = (8192 - nyc // 2) % 16
    S, _ = invlinearity(conv / g_in, lin, origin=(nb, nb))
    if not electrons_out:
        return S
    return g_out * (S - lin["Sref"][nb : nb + nyc, nb : nb + nxc])


# ---------------------------------------------------------------------------------------------------------
# Reference-pixel correction (utils/reference_subtraction.py:16-125, L1_to_L2/gen_cal_image.py:531-556)
# ---------------------------------------------------------------------------------------------------------


def ref_subtraction_row(image, use_ref_channel=False, slope=None):
    """Row correction from per-row reference medians (reference_subtraction.py:104-125).  In place.

    The reference hard-codes the 4096-pixel geometry (pars.nside); here the frame side is ``image.shape[0]`` so
    that small frames can be used in tests -- identical at 4096 (pinned by tests/golden/refsub_4096.npz).
    """
    ns = image.shape[0]
    if use_ref_channel:
        ref_med = np.median(image[:, ns : ns + CHANNELWIDTH], axis=1)
    else:
        ref_med = np.median(np.hstack((image[:, 0:4], image[:, ns - 4 : ns])), axis=1)
    ref_med = np.asarray(ref_med)
    if slope is None:
        sci_med = np.median(image[:, 4 : ns - 4], axis=1)
        m_med, _ = np.polyfit(ref_med, sci_med, 1)
    else:
        m_med = slope
    ctr = np.median(ref_med)
    image[:, :] = image - (m_med * (ref_med - ctr))[:, None]
    return image


def ref_subtraction_channel(image, use_ref_channel=False):
    """Per-128-column-channel line through bottom/top reference medians (reference_subtraction.py:45-74)."""
    ns = image.shape[0]
    nch = ns // CHANNELWIDTH + (1 if use_ref_channel else 0)
    rows = np.arange(ns)
    for c in range(nch):
        ch = image[:, 128 * c : 128 * (c + 1)]
        bottom = np.median(ch[0:4, :])
        top = np.median(ch[ns - 4 : ns, :])
        A = np.vstack([(1.5, ns - 2.5), np.ones(2)]).T
        m_cor, c_cor = np.linalg.lstsq(A, (bottom, top), rcond=None)[0]
        ch[:, :] = ch - (m_cor * rows + c_cor)[:, None]
    return image


def optimal_refout_slope(read):
    """The once-per-exposure reference-output coefficient (gen_cal_image.py:542-553); np.float64."""
    a = read["amp33"]
    cvar = read["anc"]["C_PINK"] ** 2
    return a["M_PINK"] * cvar / (a["M_PINK"] ** 2 * cvar + a["RU_PINK"] ** 2 + np.median(a["std"]) ** 2 / 128 / np.log(4096))


def refpix_loop(data, amp33, dark_cube, read):
    """The per-group reference-pixel loop of calibrateimage (gen_cal_image.py:530-556).  data f32 in place."""
    ngrp = data.shape[0]
    slope = optimal_refout_slope(read)
    for j in range(ngrp):
        ns = data.shape[1]
        image = np.zeros((ns, ns + CHANNELWIDTH), dtype=np.float32)
        image[:, :ns] = data[j] - dark_cube[j]
        image[:, -CHANNELWIDTH:] = amp33[j] - read["amp33"]["med"]
        image[:, -CHANNELWIDTH:] -= np.median(image[:, -CHANNELWIDTH:])
        image = ref_subtraction_row(image, use_ref_channel=True, slope=slope)
        image = ref_subtraction_channel(image, use_ref_channel=True)
        data[j] = image[:, :ns] + dark_cube[j]
    return data


# ---------------------------------------------------------------------------------------------------------
# Ramp fitting (utils/fitting.py:20-355)
# ---------------------------------------------------------------------------------------------------------


def construct_weights(u, meta, exclude_first=True):
    """Fixed optimal weights (fitting.py:63-86)."""
    K = np.zeros(meta["ngrp"])
    start = 1 if exclude_first else 0
    ngrp = meta["ngrp"] - start
    tbar = meta["tbar"][start:].astype(np.float64)
    tau = meta["tau"][start:].astype(np.float64)
    C = np.zeros((ngrp, ngrp))
    for i in range(ngrp):
        C[i, i] = 1.0 / meta["N"][start + i] + u * tau[i]
        for j in range(i):
            C[i, j] = C[j, i] = u * tbar[j]
    W = np.linalg.inv(C)
    Ws = np.sum(W, axis=0)
    Wt = W @ tbar
    F0 = np.sum(W)
    F1 = np.sum(Wt)
    F2 = np.dot(tbar, Wt)
    D = F0 * F2 - F1**2
    K[start:] = (F0 * Wt - F1 * Ws) / D
    return K.astype(np.float32)


def jump_detect(data, rdq, pdq, meta, gain, read, exclude_first=True, truncate_ramp=None):
    """Slope, errors and Sharma-Casertano jump flags (fitting.py:157-255).  rdq updated in place."""
    ngrp = meta["ngrp"]
    ny, nx = pdq.shape
    start = 1 if exclude_first else 0
    K = meta["K"]
    if truncate_ramp is not None:
        ngrp = truncate_ramp
        K = np.zeros(ngrp, dtype=np.float32)
        K[-1] = 1.0 / (meta["tbar"][ngrp - 1] - meta["tbar"][start])
        K[start] = -K[-1]
    SthreshA, SthreshB, IthreshA, IthreshB = 5.5, 4.5, 1.0, 1000.0
    jp = meta.get("jump_detect_pars", {})
    SthreshA = float(jp.get("SthreshA", SthreshA))
    SthreshB = float(jp.get("SthreshB", SthreshB))
    IthreshA = float(jp.get("IthreshA", IthreshA))
    IthreshB = float(jp.get("IthreshB", IthreshB))

    slope = np.einsum("t,tij->ij", K, data[:ngrp] - data[1][None]).astype(np.float32)
    smap = np.zeros((2 * (ngrp - start) - 3, ny, nx), dtype=np.float32)
    coef = 0.0
    for i in range(start, ngrp):
        coef += K[i] ** 2 * meta["tau"][i]
        for j in range(start, i):
            coef += 2.0 * K[i] * K[j] * meta["tbar"][j]
    dvardt = np.clip(slope / np.clip(gain, 1e-4, 1e4), 0.0, None)
    err_p = np.sqrt(np.clip(coef * dvardt, 0, None)).astype(np.float32)
    sig2read = read**2
    err_r = (read * np.sqrt(np.sum(K**2 / np.array(meta["N"][:ngrp])))).astype(np.float32)

    x = np.clip(slope, IthreshA, IthreshB)
    x = np.log(x / IthreshA) / np.log(IthreshB / IthreshA)
    sthresh = SthreshA + (SthreshB - SthreshA) * x
    nb = meta["nborder"]
    sl = 0
    for i in range(start, ngrp - 1):
        dimax = 1 if (i == ngrp - 2 or ngrp - 1 - start == 2) else 2
        for di in range(1, 1 + dimax):
            dt = meta["tbar"][i + di] - meta["tbar"][i]
            dslope = (data[i + di] - data[i]) / dt - slope
            w = np.zeros(ngrp)
            w[i + di] = 1.0 / dt
            w[i] = -1.0 / dt
            w -= K
            var = np.zeros((ny, nx))
            for a in range(ngrp):
                var += w[a] ** 2 * (dvardt * meta["tau"][a] + sig2read / np.array(meta["N"][a]))
                for b in range(a):
                    var += 2 * w[a] * w[b] * dvardt * meta["tbar"][b]
            smap[sl] = dslope / np.sqrt(var).astype(np.float32)
            rdq[i, nb : ny - nb, nb : nx - nb] |= np.where(
                smap[sl, nb : ny - nb, nb : nx - nb] > sthresh[nb : ny - nb, nb : nx - nb], JUMP_DET, 0
            ).astype(np.uint32)
            sl += 1
    return slope, err_r, err_p, smap


def ramp_fit(data, rdq, pdq, meta, gain, read, exclude_first=True):
    """Full fit + saturation-truncated refits + DQ propagation (fitting.py:310-355).  rdq, pdq in place."""
    loc = np.zeros_like(rdq)
    slope, err_r, err_p, _ = jump_detect(data, loc, pdq, meta, gain, read, exclude_first, None)
    unsat = ~rdq[-1] & SATURATED != 0
    rdq |= np.where(unsat[None], loc, 0)
    start = 1 if exclude_first else 0
    for iend in range(meta["ngrp"] - 1, 2 + start, -1):
        layer = rdq[iend] & ~rdq[iend - 1] & SATURATED != 0
        loc[:, :, :] = 0
        s_, r_, p_, _ = jump_detect(data, loc, pdq, meta, gain, read, exclude_first, iend)
        slope = np.where(layer, s_, slope)
        err_r = np.where(layer, r_, err_r)
        err_p = np.where(layer, p_, err_p)
        rdq |= np.where(layer[None], loc, 0)
    pdq2 = np.zeros_like(pdq)
    dnu = np.uint32(DO_NOT_USE)
    pdq2 |= np.bitwise_or.reduce(np.where(~rdq & SATURATED, rdq, 0), axis=0) & ~dnu
    pdq2 |= np.where(np.bitwise_and.reduce(rdq & DO_NOT_USE != 0, axis=0), dnu, 0).astype(np.uint32)
    pdq2 |= np.where(rdq[1 + start] & SATURATED != 0, DO_NOT_USE, 0).astype(np.uint32)
    pdq2 |= np.bitwise_or.reduce(rdq & SATURATED, axis=0)
    pdq |= np.where(~pdq & REFERENCE_PIXEL, pdq2, 0)
    return slope, err_r, err_p


# ---------------------------------------------------------------------------------------------------------
# Flat (utils/flatutils.py:44-76)
# ---------------------------------------------------------------------------------------------------------


def get_flat(flat_full, gain_full, ipc_kernel, nborder, pdq, ipc_deconvolve=True):
    """Padded, clipped, flagged, IPC-deconvolved flat in DN units (flatutils.py:44-76).  pdq in place."""
    ny, nx = flat_full.shape
    nb = nborder
    f = np.ones((ny, nx), dtype=np.float32)
    f[nb : ny - nb, nb : nx - nb] = flat_full[nb : ny - nb, nb : nx - nb]
    if pdq is not None:
        pdq |= np.where(np.logical_or(f < 0.1, f > 10), NO_FLAT_FIELD, 0).astype(np.uint32)
    f = np.clip(f, 0.1, 10)
    if ipc_deconvolve:
        g = gain_full[nb : ny - nb, nb : nx - nb]
        if pdq is not None:
            pdq[nb : ny - nb, nb : nx - nb] |= np.where(g <= 0.1, NO_GAIN_VALUE, 0).astype(np.uint32)
            g = np.clip(g, 0.1, None)
        f[nb : ny - nb, nb : nx - nb] = ipc_rev(f[nb : ny - nb, nb : nx - nb], ipc_kernel, gain=g)
    return f


# ---------------------------------------------------------------------------------------------------------
# Third-party steps -- RESTATEMENTS, parity unpinned
# ---------------------------------------------------------------------------------------------------------


def expand_gw(mask_dq, iterations=1):
    """do_dqinit(..., expand_gw_flagging=1) (gen_cal_image.py:118): GW_AFFECTED_DATA grown by ``iterations`` pixels
    (binary dilation with the 4-connected structure).  Restated from upstream, parity unpinned."""
    out = mask_dq.astype(np.uint32).copy()
    gw = (out & np.uint32(GW_AFFECTED_DATA)) != 0
    for _ in range(iterations):
        g = gw.copy()
        g[1:, :] |= gw[:-1, :]
        g[:-1, :] |= gw[1:, :]
        g[:, 1:] |= gw[:, :-1]
        g[:, :-1] |= gw[:, 1:]
        gw = g
    out[gw] |= np.uint32(GW_AFFECTED_DATA)
    return out


def dq_init(data_u16, mask_dq, exclude_first=True, expand_gw_flagging=1):
    """romancal do_dqinit as used at gen_cal_image.py:117-143: f32 data, pixeldq = mask dq (GW_AFFECTED_DATA grown by
    one pixel), groupdq u8."""
    data = data_u16.astype(np.float32)
    pdq = np.zeros(data.shape[1:], dtype=np.uint32) if mask_dq is None else expand_gw(mask_dq, expand_gw_flagging)
    rdq = np.zeros(data.shape, dtype=np.uint8)
    if exclude_first:
        rdq[0] |= np.uint8(DO_NOT_USE)
    return data, rdq, pdq


def _grow3(flag):
    """3x3 box dilation of a boolean map (n_pix_grow_sat=1)."""
    out = flag.copy()
    out[1:, :] |= flag[:-1, :]
    out[:-1, :] |= flag[1:, :]
    h = out.copy()
    out[:, 1:] |= h[:, :-1]
    out[:, :-1] |= h[:, 1:]
    return out


def flag_saturation(data, rdq, pdq, sat_thresh, sat_dq, backup=1, skip_firstn=1):
    """saturation_check -> romancal flag_saturation(n_pix_grow_sat=1, backup) (gen_cal_image.py:148-185).

    Restated (docs/L1_to_L2_README.rst:87): on the raw groups ``skip_firstn..G-1``: thresholds
    with NO_SAT_CHECK or NaN are never exceeded (pixeldq |= NO_SAT_CHECK); ``data >= thresh`` sets SATURATED in
    that and all later groups, grown by one pixel (3x3); ``data <= 0`` sets AD_FLOOR|DO_NOT_USE in that group
    only; then SATURATED is also set in the ``backup`` groups preceding a saturated group (never in the skipped
    leading groups).  The group-averaging look-back passes of stcal only add DO_NOT_USE to single groups, which
    the reference's fitter ignores unless every group has it; they are not restated.
    """
    G = data.shape[0]
    thr = sat_thresh.astype(np.float32).copy()
    nocheck = (sat_dq & NO_SAT_CHECK) != 0
    nocheck |= np.isnan(thr)
    thr[nocheck] = np.inf
    pdq |= np.where((sat_dq & NO_SAT_CHECK) != 0, NO_SAT_CHECK, 0).astype(np.uint32)
    cum = np.zeros(data.shape[1:], dtype=bool)
    sat = np.zeros(data.shape, dtype=bool)
    for g in range(skip_firstn, G):
        cum |= data[g] >= thr
        sat[g] = _grow3(cum)
        rdq[g] |= np.where(data[g] <= 0, np.uint8(AD_FLOOR | DO_NOT_USE), np.uint8(0))
    if backup > 0:
        s2 = sat.copy()
        for g in range(skip_firstn, G):
            for b in range(1, backup + 1):
                if g + b < G:
                    s2[g] |= sat[g + b]
        sat = s2
    rdq |= np.where(sat, np.uint8(SATURATED), np.uint8(0))
    return rdq, pdq


def subtract_dark_current(slope_full, dq_full, dark_slope_ipc, dark_dq, nb=4):
    """romancal subtract_dark_current on the active region (gen_cal_image.py:223-229): data -= dark; dq |= dark dq."""
    slope_full[nb:-nb, nb:-nb] -= dark_slope_ipc[nb:-nb, nb:-nb]
    dq_full[nb:-nb, nb:-nb] |= dark_dq[nb:-nb, nb:-nb]


# ---------------------------------------------------------------------------------------------------------
# The whole L1 -> L2 numerics chain (gen_cal_image.py:480-629, 697-709)
# ---------------------------------------------------------------------------------------------------------


def make_meta(read_pattern, frame_time):
    """N, tbar, tau (gen_cal_image.py:123-140)."""
    ngrp = len(read_pattern)
    meta = {"frame_time": frame_time, "read_pattern": read_pattern, "ngrp": ngrp, "nborder": 4}
    meta["tbar"] = np.zeros(ngrp, dtype=np.float32)
    meta["tau"] = np.zeros(ngrp, dtype=np.float32)
    meta["N"] = np.zeros(ngrp, dtype=np.int16)
    for i in range(ngrp):
        meta["N"][i] = len(read_pattern[i])
        t0 = read_pattern[i][0]
        meta["tbar"][i] = (t0 + (meta["N"][i] - 1) / 2.0) * frame_time
        meta["tau"][i] = (t0 + (meta["N"][i] - 1) * (2 * meta["N"][i] - 1) / (6.0 * meta["N"][i])) * frame_time
    return meta


def apply_refpix_corrections(data, dark_cube, rowcorr, chan_m, chan_c, row0=0):
    """The subtractions of the reference-pixel loop (gen_cal_image.py:534-556) with the statistics GIVEN: per group
    ``image = data - dark``; ``image[i,:] -= rowcorr[i]`` (reference_subtraction.py:122-123, float64 then float32 store);
    ``image[j, channel] -= m*j + c`` (:64-68, same); ``data = image + dark``.  ``row0`` = detector row of ``data[:, 0]``
    (lets a row band of the frame be corrected with the statistics of the whole frame)."""
    G, ny, nx = data.shape
    jj = np.arange(row0, row0 + ny, dtype=np.float64)
    nch = nx // CHANNELWIDTH
    for g in range(G):
        v = data[g] - dark_cube[g]
        v = (v - rowcorr[g][:, None]).astype(np.float32)
        line = chan_m[g, :nch][:, None] * jj[None, :] + chan_c[g, :nch][:, None]  # [nch, rows]
        v = (v - np.repeat(line.T, CHANNELWIDTH, axis=1)).astype(np.float32)
        data[g] = v + dark_cube[g]
    return data


def l1_to_l2(data_u16, amp33_u16, cal, read_pattern, frame_time, area_factor, config=None, do_refpix=True,
             return_intermediates=False, refpix_corr=None, fns=None):  # fmt: skip
    """calibrateimage numerics from L1 arrays to L2 arrays (gen_cal_image.py:503-629,697-709).

    ``cal`` maps CALDIR keys to the ``roman`` branches.  Returns a dict with slope, err_read, err_poisson, pdq
    (all full frame [n,n]), rdq [G,n,n] u8, endslice i8 [n-8,n-8], K, and optionally the intermediate cubes.
    ``do_refpix=False`` skips the 4096-only reference-pixel loop (small-frame tests).
    ``refpix_corr = (rowcorr [G,rows], chan_m [G,32], chan_c [G,32], row0)``: apply these reference-pixel corrections
    instead of deriving them from the frame -- the frame may then be a ROW BAND (all planes of ``cal`` cut to the same
    rows, ipc4d / biascorr to the matching active rows): the outer 4 rows act as a fake border whose influence ends 6
    rows in (saturation growth 1 + IPC order 2), so the interior of a band with a 6-row halo equals the whole-frame
    result.  tests/fullframe.py tiles a 4096^2 frame this way over a process pool.
    ``fns``: replacements for the reference-owned steps (refpix_loop, multilin, correct_cube, construct_weights, ramp_fit,
    get_flat) -- oracle/ref_chain.py passes the reference's own unmodified functions here.
    """
    config = config or {}
    fns = fns or {}
    _refpix_loop = fns.get("refpix_loop", refpix_loop)
    _multilin = fns.get("multilin", multilin)
    _correct_cube = fns.get("correct_cube", correct_cube)
    _construct_weights = fns.get("construct_weights", construct_weights)
    _ramp_fit = fns.get("ramp_fit", ramp_fit)
    _get_flat = fns.get("get_flat", get_flat)
    nb = 4
    exclude_first = config.get("EXCLUDE_FIRST", True)
    backup = config.get("SATURATION_BACKUP", 1)
    meta = make_meta(read_pattern, frame_time)
    mask_dq = cal["mask"]["dq"] if "mask" in cal else None
    data, rdq, pdq = dq_init(data_u16, mask_dq, exclude_first)
    flag_saturation(data, rdq, pdq, cal["saturation"]["data"], cal["saturation"]["dq"], backup=backup)
    ngrp = data.shape[0]
    inter = {}
    if refpix_corr is not None:
        rowcorr, chan_m, chan_c, row0 = refpix_corr
        apply_refpix_corrections(data, cal["dark"]["data"], rowcorr, chan_m, chan_c, row0)
    elif do_refpix:
        _refpix_loop(data, amp33_u16, cal["dark"]["data"], cal["read"])
    if return_intermediates:
        inter["refcorr"] = data.copy()
    if "biascorr" in cal:
        bc = cal["biascorr"]["data"]
        de = bc.shape[0] - ngrp
        data[:, nb:-nb, nb:-nb] -= bc[de:]
    data, dq_lin = _multilin(
        data,
        cal["linearitylegendre"],
        do_not_flag_first=(list(read_pattern[0]) == [0]),
        attempt_corr=~rdq & SATURATED,
    )
    pdq |= dq_lin
    if return_intermediates:
        inter["lin"] = data.copy()
    if "ipc4d" in cal:
        _correct_cube(data, cal["ipc4d"]["data"], cal["gain"]["data"])
    if return_intermediates:
        inter["ipc"] = data.copy()
    uopt = config.get("RAMP_OPT_PARS", {"slope": 0.4, "gain": 1.8, "sigma_read": 6.5})
    u_ = float(uopt["slope"]) / float(uopt["gain"]) / float(uopt["sigma_read"]) ** 2
    meta["K"] = _construct_weights(u_, meta, exclude_first=exclude_first)
    if "JUMP_DETECT_PARS" in config:
        meta["jump_detect_pars"] = config["JUMP_DETECT_PARS"]
    slope, err_r, err_p = _ramp_fit(data, rdq, pdq, meta, cal["gain"]["data"], cal["read"]["data"], exclude_first)
    # do_ramp_fit packaging (gen_cal_image.py:458-475): err, var_poisson, border zeroed
    err = np.hypot(err_r, err_p)
    varp = err_p**2

    def embed(a):
        full = np.zeros(a.shape, dtype=np.float32)
        full[nb:-nb, nb:-nb] = a[nb:-nb, nb:-nb]
        return full

    slope = embed(slope)
    varp = embed(varp)
    err = embed(err)
    # dark current (gen_cal_image.py:212-229)
    dslope = np.array(cal["dark"]["dark_slope"], dtype=np.float32)[None]
    if "ipc4d" in cal:
        _correct_cube(dslope, cal["ipc4d"]["data"], cal["gain"]["data"])
    subtract_dark_current(slope, pdq, dslope[0], cal["dark"]["dq"], nb)
    # unpack + error split (gen_cal_image.py:607-613)
    err_p = np.sqrt(varp)
    err_r = np.sqrt(np.clip(err**2 - err_p**2, 0.0, None))
    # flat + area (gen_cal_image.py:616-629)
    flat = _get_flat(cal["flat"]["data"], cal["gain"]["data"], cal["ipc4d"]["data"], nb, pdq)
    flat = (flat / area_factor).astype(np.float32)
    slope /= flat
    err_r /= flat
    err_p /= flat
    # endslice (gen_cal_image.py:697-709)
    endslice = np.zeros(rdq[0, nb:-nb, nb:-nb].shape, dtype=np.int8) - 1
    for iend in range(1, ngrp):
        endslice = np.where(
            rdq[iend, nb:-nb, nb:-nb] & ~rdq[iend - 1, nb:-nb, nb:-nb] & SATURATED != 0, iend - 1, endslice
        )
    out = {"slope": slope, "err_read": err_r, "err_poisson": err_p, "pdq": pdq, "rdq": rdq,
           "endslice": endslice.astype(np.int8), "K": meta["K"], "flat": flat, "meta": meta}  # fmt: skip
    out.update(inter)
    return out


# ---------------------------------------------------------------------------------------------------------
# Forward model (from_sim/sim_to_isim.py:163-262 + romanisim restatements, parity unpinned)
# ---------------------------------------------------------------------------------------------------------


def read_pattern_to_tij(read_pattern, read_time=3.04):
    """romanisim.l1.read_pattern_to_tij as used at sim_to_isim.py:204: time of each read = read_time*index."""
    return [[read_time * r for r in grp] for grp in read_pattern]


def forward_deterministic(mean_counts_per_read, cal, read_pattern, start_e):
    """Noise-free forward ramp: per read IL.apply(cumulative electrons), group mean, + biascorr, round.

    ``mean_counts_per_read`` [n_reads_total, na, na] cumulative electrons at each read (already apportioned).
    Restates the deterministic part of make_l1_fullcal (sim_to_isim.py:222-260): used to check the CUDA
    kernel's IL/bisection/group-average arithmetic exactly, with the RNG draws supplied from outside.
    """
    lin = cal["linearitylegendre"]
    res = []
    k = 0
    for grp in read_pattern:
        acc = None
        for _ in grp:
            s = il_apply(mean_counts_per_read[k], lin, cal["gain"]["data"], cal["ipc4d"]["data"], start_e=start_e)
            acc = s if acc is None else acc + s
            k += 1
        res.append((acc / len(grp)).astype(np.float32))
    return np.stack(res)


def make_l1_fullcal(counts, cal, read_pattern, rng, read_time=3.04, add_reset_noise=True, add_read_noise=True,
                    quantize=True, cum_counts_out=None, crparam=None, cr_groups_out=None):  # fmt: skip
    """make_l1_fullcal (from_sim/sim_to_isim.py:195-260) with romanisim's apportioning and read noise RESTATED
    (parity unpinned): NumPy ``rng`` (a ``np.random.Generator``) instead of GalSim deviates.

    counts int32 [na,na] total electrons of the exposure.  Reset noise N(0,1)*resetnoise*gain - t0*dark_slope/gain
    (:195-215); per read, electrons so far ~ sequential Binomial(remaining, dt/(t_last - t_prev)); IL.apply per read
    (float64); resultant = mean over the group's reads (float32); + N(0,1)*read/sqrt(N) (:246-253); + biascorr
    (:256-258); round (:260).  Returns float32 [G,na,na].
    """
    nb = 4
    lin = cal["linearitylegendre"]
    gain = cal["gain"]["data"]
    g_act = gain[nb:-nb, nb:-nb]
    na = counts.shape[0]
    start_e = np.zeros((na, na), dtype=np.float32)
    if add_reset_noise:
        start_e = rng.standard_normal((na, na), dtype=np.float32)
        start_e *= cal["read"]["resetnoise"][nb:-nb, nb:-nb]
        start_e = (start_e * g_act).astype(np.float32)
    if "biascorr" in cal:
        tbias = float(cal["biascorr"]["t0"])
        start_e = (start_e - tbias * cal["dark"]["dark_slope"][nb:-nb, nb:-nb] / g_act).astype(np.float32)
    tij = read_pattern_to_tij(read_pattern, read_time)
    t_last = tij[-1][-1]
    remaining = np.clip(counts, 0, 2000000000).astype(np.int64)
    cum = np.zeros((na, na), dtype=np.int64)
    t_prev = 0.0
    res = []
    k = 0
    for grp in tij:
        acc = None
        for t in grp:
            if t > t_prev:
                p = (t - t_prev) / (t_last - t_prev) if t_last > t_prev else 1.0
                d = rng.binomial(remaining, min(p, 1.0))
                cum += d
                remaining -= d
                if crparam is not None:  # (restated romanisim: cosmic rays of the interval since the previous read)
                    before = cum.copy()
                    simulate_crs(cum, t - t_prev, rng, **crparam)
                    if cr_groups_out is not None:
                        cr_groups_out[len(res)] |= cum != before
            t_prev = t
            if cum_counts_out is not None:
                cum_counts_out[k] = cum
            k += 1
            s = il_apply(cum.astype(np.int32), lin, gain, cal["ipc4d"]["data"] if "ipc4d" in cal else None,
                         start_e=start_e)  # fmt: skip
            acc = s if acc is None else acc + s
        res.append((acc / len(grp)).astype(np.float32))
    res = np.stack(res)
    if add_read_noise:
        for g, grp in enumerate(tij):
            res[g] += rng.standard_normal((na, na), dtype=np.float32) * (
                cal["read"]["data"][nb:-nb, nb:-nb] / np.float32(np.sqrt(np.float32(len(grp))))
            )
    if "biascorr" in cal:
        bc = cal["biascorr"]["data"]
        res += bc[bc.shape[0] - len(tij) :]
    if quantize:
        res = np.round(res)
    return res.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------
# Cosmic rays: romanisim.cr (romanisim 0.x; third-party, absent from /root/reference) RESTATED from its published
# algorithm -- parity unpinned.  The reference switches it on with crparam={} (from_sim/sim_to_isim.py:233-242):
# romanisim.l1.apportion_counts_to_resultants calls cr.simulate_crs(electrons_so_far, read_time, **crparam) once per
# read and flags the pixels it changed with JUMP_DET in that resultant's dq.
# ---------------------------------------------------------------------------------------------------------
def _cr_sampler(pdf, x):
    """romanisim.cr.create_sampler: inverse-transform sampler of a tabulated pdf (linear interpolation)."""
    y = pdf(x)
    cdf = np.cumsum(y) - y[0]
    cdf /= cdf.max()
    return lambda u: np.interp(u, cdf, x)


def cr_sample_params(n_samples, n_i, n_j, rng, min_dedx=10, max_dedx=10000, min_cr_len=10, max_cr_len=2000, grid_size=10000):
    """romanisim.cr.sample_cr_params: positions [pix], direction [rad], projected length [um], dE/dx [eV/um]."""
    cr_i, cr_j = (rng.random(size=(n_samples, 2)) * (n_i, n_j)).transpose()
    cr_phi = rng.random(n_samples) * 2 * np.pi
    len_grid = np.linspace(min_cr_len, max_cr_len, grid_size)
    cr_length = _cr_sampler(lambda x: np.power(x, -4.33), len_grid)(rng.random(n_samples))
    dedx_grid = np.linspace(min_dedx, max_dedx, grid_size)

    def moyal(x, location=120, scale=50):
        xs = (x - location) / scale
        return np.exp(-(xs + np.exp(-xs)) / 2)

    cr_dedx = _cr_sampler(moyal, dedx_grid)(rng.random(n_samples))
    return cr_i, cr_j, cr_phi, cr_length, cr_dedx


def cr_traverse(start, end, n_i, n_j):
    """romanisim.cr.traverse: pixels crossed by the segment start -> end (pixel centres at integers, borders at
    half-integers) and the path length inside each [pixels]; pixels outside the array are dropped."""
    (i0, j0), (i1, j1) = start, end
    di, dj = i1 - i0, j1 - j0
    ts = [0.0, 1.0]
    for a0, d in ((i0, di), (j0, dj)):
        if d != 0:
            lo, hi = min(a0, a0 + d), max(a0, a0 + d)
            b = np.arange(np.ceil(lo - 0.5) + 0.5, hi, 1.0)
            tt = (b - a0) / d
            ts.extend(tt[(tt > 0) & (tt < 1)].tolist())
    ts = np.unique(np.array(ts))
    tm = 0.5 * (ts[1:] + ts[:-1])
    ii = np.rint(i0 + tm * di).astype(int)
    jj = np.rint(j0 + tm * dj).astype(int)
    length = (ts[1:] - ts[:-1]) * np.hypot(di, dj)
    ok = (ii >= 0) & (ii < n_i) & (jj >= 0) & (jj < n_j) & (length > 0)
    return ii[ok], jj[ok], length[ok]


def simulate_crs(image, time, rng, flux=8, area=16.8, conversion_factor=0.5, pixel_size=10, pixel_depth=5):
    """romanisim.cr.simulate_crs: adds cosmic-ray electrons to ``image`` in place; returns (image, number of events)."""
    n_i, n_j = image.shape
    n_samples = rng.poisson(flux * area * time)
    ci, cj, phi, length, dedx = cr_sample_params(n_samples, n_i, n_j, rng)
    length = length / pixel_size
    i1 = (ci + length * np.cos(phi)).clip(-0.5, n_i + 0.5)
    j1 = (cj + length * np.sin(phi)).clip(-0.5, n_j + 0.5)
    counts_per_pix = dedx * pixel_size / conversion_factor
    for a0, b0, a1, b1, cpp in zip(ci, cj, i1, j1, counts_per_pix):
        ii, jj, l2 = cr_traverse((a0, b0), (a1, b1), n_i, n_j)
        l3 = ((pixel_depth / pixel_size) ** 2 + l2**2) ** 0.5
        image[ii, jj] += rng.poisson(cpp * l3).astype(image.dtype)
    return image, n_samples


# ---------------------------------------------------------------------------------------------------------
# Scene counts (a15) and reference-pixel / 1/f fill (a20): from_sim/sim_to_isim.py:265-402, 615-662
# ---------------------------------------------------------------------------------------------------------
G_IDEAL = 1.458  # pars.g_ideal (reference pars.py:21)


class NormalStream:
    """Stand-in for ``galsim.GaussianDeviate(rng).generate(array)``: float64 normals from a NumPy generator, cast to
    the array's dtype, consumed in call order.  tests/golden/make_golden.py feeds the UNMODIFIED reference functions
    from an identical stream, which pins the draw order and every deterministic step of the restatements below."""

    def __init__(self, seed):
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def generate(self, array):
        array[...] = self.rng.standard_normal(array.size).reshape(array.shape)


def noise_1f_frame(stream, nside=4096, channelwidth=128):
    """One (nside, channelwidth) block of 1/f noise, S(f) = 1/f (sim_to_isim.py:265-303): complex white spectrum of
    2*nside*channelwidth points scaled by |k|^-1/2 (k = 0 removed), forward FFT, real part of the first half / sqrt 2,
    mean removed, float32."""
    m = 2 * nside * channelwidth
    draws = np.zeros(2 * m)
    stream.generate(draws)
    k = np.linspace(0, 1 - 1.0 / m, m)
    k[m // 2 :] -= 1.0
    amp = (1.0e-99 + np.abs(k * m)) ** (-0.5)
    amp[0] = 0.0
    spec = np.zeros((m,), dtype=np.complex128)
    spec[:] = draws[:m]
    spec[:] += 1j * draws[m:]
    spec *= amp
    block = np.fft.fft(spec).real[: m // 2] / np.sqrt(2.0)
    block -= np.mean(block)
    return block.reshape((nside, channelwidth)).astype(np.float32)


def fill_in_refdata_and_1f(im, cal, stream, tij, fill_in_banding=True, amp33=None, nborder=4):
    """fill_in_refdata_and_1f (sim_to_isim.py:306-402), in place on ``im`` [G,n,n] (and ``amp33`` [G,n,n/32]).

    Reference pixels = N(0,1)*read/sqrt(N_g) + N(0,1)*resetnoise (one layer for all groups) + dark cube (:341-352);
    active pixels keep ``im`` (:356-358); per group one common and 32 per-channel 1/f frames, odd channels mirrored,
    added to EVERY pixel as (u_pink*frame + c_pink*common)/sqrt(N_g) (:376-389); reference output = med +
    (N(0,1)*std + RU_PINK*frame + M_PINK*c_pink*common)/sqrt(N_g) cast to the cube dtype (:392-399); finally
    clip(round(.), 0, 65535) (:402).  Draw order as in the reference.
    """
    G, ny, nx = im.shape
    cw = nx // 32
    rd = cal["read"]
    noise = np.zeros((G + 1, ny, nx), dtype=np.float32)
    stream.generate(noise)
    noise[:-1] *= rd["data"][None]
    noise[-1] *= rd["resetnoise"]
    for j in range(len(tij)):
        noise[j] /= len(tij[j]) ** 0.5
    noise[:-1] += noise[-1][None]
    dk = cal["dark"]["data"]
    noise[:-1] += np.copy(dk[dk.shape[0] - G :])
    nb = nborder
    noise[:-1, nb : ny - nb, nb : nx - nb] = im[:, nb : ny - nb, nb : nx - nb].astype(noise.dtype)
    a33 = {"valid": False}
    if amp33 is not None and "amp33" in rd:
        a33 = rd["amp33"]
    if fill_in_banding:
        u_pink, c_pink = float(rd["anc"]["U_PINK"]), float(rd["anc"]["C_PINK"])
        for j in range(len(tij)):
            rn = len(tij[j]) ** 0.5
            common = noise_1f_frame(stream, ny, cw) * c_pink
            for ch in range(32):
                pink = noise_1f_frame(stream, ny, cw) * u_pink + common
                if ch % 2 == 1:
                    pink = pink[:, ::-1]
                noise[j, :, cw * ch : cw * (ch + 1)] += (pink / rn).astype(noise.dtype)
            if a33["valid"]:
                white = np.zeros((ny, cw), dtype=np.float32)
                stream.generate(white)
                white *= a33["std"]
                pink = a33["RU_PINK"] * noise_1f_frame(stream, ny, cw) + a33["M_PINK"] * common
                amp33[j] = (a33["med"] + (white + pink) / rn).astype(amp33.dtype)
    im[...] = np.clip(np.round(noise[:-1]), 0, 2**16 - 1).astype(im.dtype)


def sim_calprep(cal, nborder=4):
    """Calibration planes of Image2D.simulate (sim_to_isim.py:615-633): dark rate in e/s and flat, both IPC-deconvolved
    on the active array, with the reference's clips.  Returns (this_dark, this_flat, g)."""
    nb = nborder
    this_dark = cal["dark"]["dark_slope"][nb:-nb, nb:-nb]
    this_flat = cal["flat"]["data"][nb:-nb, nb:-nb]
    this_dark = this_dark * cal["gain"]["data"][nb:-nb, nb:-nb]
    g = np.copy(cal["gain"]["data"][nb:-nb, nb:-nb])
    K = cal["ipc4d"]["data"]
    this_dark = ipc_rev(this_dark, K)
    this_flat = ipc_rev(this_flat, K, gain=g)
    this_flat = np.clip(this_flat, 0.0, 2 - 2**-21)
    this_dark = np.clip(this_dark, -0.1 * this_flat, None)
    return this_dark, this_flat, g


def scene_rate(image, this_flat, g, area_ratio, t, cnorm=1.0):
    """Poisson mean of the scene electrons (sim_to_isim.py:636-648): clip(C t g/g_ideal image flat/area, 0)."""
    flat_witharea = this_flat / area_ratio
    return np.clip(cnorm * t * g / G_IDEAL * image * flat_witharea, 0, None)


# ---------------------------------------------------------------------------------------------------------
# Mask growth and the moment sums of the many-realisations protocol
# (utils/maskhandling.py:82-117, 152-178; validation_tests/many_realizations.py:74-83)
# ---------------------------------------------------------------------------------------------------------
PIXELMASK1 = {0: 1, 2: 5, 3: 25, 4: 1, 5: 1, 6: 5, 8: 1, 9: 1, 10: 9, 11: 9, 12: 1, 13: 9, 15: 1, 18: 9, 19: 9, 20: 9,
              21: 9, 22: 1, 23: 9, 24: 9, 25: 9, 28: 9, 30: 9}  # bit -> pixels affected  # fmt: skip


def mask_build(dq, grow=None):
    """CombinedMask.build (maskhandling.py:82-117): OR over bits of the bit plane dilated by a cross (5), a 3x3 (9)
    or a 5x5 (25) box, zero padded (scipy.signal.convolve mode='same' of a 0/2 layer with a 0/1 kernel, >= 1)."""
    grow = PIXELMASK1 if grow is None else grow
    ny, nx = dq.shape
    mask = np.zeros((ny, nx), dtype=bool)
    foot = {
        5: [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)],
        9: [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)],
        25: [(a, b) for a in range(-2, 3) for b in range(-2, 3)],
    }
    for bit, g in grow.items():
        if g == 0:
            continue
        layer = (dq & np.uint32(1 << bit)) != 0
        if g == 1:
            mask |= layer
            continue
        pad = np.zeros((ny + 4, nx + 4), dtype=bool)
        pad[2:-2, 2:-2] = layer
        for a, b in foot[g]:
            mask |= pad[2 + a : 2 + a + ny, 2 + b : 2 + b + nx]
    return mask


def moments_accumulate(moments, data, dq, grow=None):
    """many_realizations.py:74-77 on one realisation (float32 accumulators [3,ny,nx], in place)."""
    w = np.logical_not(mask_build(dq, grow))
    moments[0] += np.where(w, 1, 0.0)
    moments[1] += np.where(w, data, 0.0)
    moments[2] += np.where(w, data**2, 0.0)


def moments_finalize(moments):
    """many_realizations.py:80-83 (in place): mean, std, -1000 where no realisation was unmasked."""
    moments[1:] /= moments[0] + 1e-25
    moments[2] = np.sqrt(np.clip(moments[2] - moments[1] ** 2, 0, None))
    moments[1:] = np.where(moments[0][None] > 0.1, moments[1:], -1000.0)


# ---------------------------------------------------------------------------------------------------------
# Sky model: utils/sky.py:98-190 (medfit)
# ---------------------------------------------------------------------------------------------------------
def medfit(arr, N=8, order=2):
    """Low-order 2D Legendre fit to the nan-medians of N x N regions; returns (coef, model as arr.dtype)."""
    from scipy.special import legendre_p

    (ny, nx) = np.shape(arr)
    kx, ky = nx // N, ny // N
    px, py = (nx % N) // 2, (ny % N) // 2
    u_ = 2 * (px - 0.5 + kx * np.linspace(0.5, N - 0.5, N)) / nx - 1
    v_ = 2 * (py - 0.5 + ky * np.linspace(0.5, N - 0.5, N)) / ny - 1
    u, v = np.meshgrid(u_, v_)
    meds = np.nanmedian(arr[py : py + N * ky, px : px + N * kx].reshape((N, ky, N, kx)), axis=(1, 3))
    nc = (order + 1) * (order + 2) // 2
    basis = np.zeros((nc, N, N))
    k = 0
    for i in range(order + 1):
        temp = legendre_p(i, u)
        for j in range(order + 1 - i):
            basis[k] = temp * legendre_p(j, v)
            k += 1
    A = np.zeros((nc, nc))
    b = np.zeros(nc)
    for ipix in range(N):
        for jpix in range(N):
            if not np.isnan(meds[jpix, ipix]):
                A += np.outer(basis[:, jpix, ipix], basis[:, jpix, ipix])
                b += meds[jpix, ipix] * basis[:, jpix, ipix]
    x = np.linalg.solve(A, b)
    LPX = np.array([legendre_p(i, np.linspace(-1, 1 - 2 / nx, nx)) for i in range(order + 1)])
    LPY = np.array([legendre_p(j, np.linspace(-1, 1 - 2 / ny, ny)) for j in range(order + 1)])
    arrmed = np.zeros((ny, nx))
    k = 0
    for i in range(order + 1):
        for j in range(order + 1 - i):
            arrmed += x[k] * np.outer(LPY[j], LPX[i])
            k += 1
    return x, arrmed.astype(arr.dtype), meds
