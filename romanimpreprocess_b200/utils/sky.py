"""Sky estimation on the GPU: drop-in for ``romanimpreprocess.utils.sky.medfit`` (reference utils/sky.py:98-190).

``medfit`` fits a low-order 2D Legendre polynomial to the medians of N x N regions.  The two passes over the image run
on the device: the region medians (``rip_block_nanmedian_dev``: exact order statistics by radix select, NaNs ignored)
and the evaluation of the model on the pixel grid (``rip_medfit_eval_dev``: float64 sum in the reference's term order,
cast to the image dtype).  The normal equations in between (64 numbers in, 6 out) and the Legendre polynomials on
the pixel grid are computed by the library (``rip_medfit_solve``, host float64: same accumulation order as the
reference, Legendre values bit-identical to ``scipy.special.legendre_p``, LU with partial pivoting).  Region medians
are exact; coefficients agree with the reference's ``np.linalg.solve`` to a few float64 ulp, the float32 model to 1 ulp.

``smooth_mode`` (reference utils/sky.py:46-93): the mode of the smoothed histogram of the 4 x 4-binned, masked slope
image (``medsky`` of the L2 file, gen_cal_image.py:641): percentiles by exact order statistics on the device, the
3 x 19 Gaussian-weighted sums by one reduction kernel each.
"""

import ctypes as C

import numpy as np

from .. import _lib


def _solve(meds, N, nx, ny, order, want_grid=True):
    """Coefficients (and grid polynomials) from the region medians: ``rip_medfit_solve`` (host float64 in the library;
    reference utils/sky.py:137-175)."""
    nc = (order + 1) * (order + 2) // 2
    coef = np.empty(nc, np.float64)
    LPX = np.empty((order + 1, nx), np.float64) if want_grid else None
    LPY = np.empty((order + 1, ny), np.float64) if want_grid else None
    m = np.ascontiguousarray(meds, dtype=np.float32)
    _lib.check(_lib.lib().rip_medfit_solve(ny, nx, N, order, _lib.ptr(m), _lib.ptr(coef), _lib.ptr(LPX), _lib.ptr(LPY)))
    return coef, LPX, LPY


_GRID_CACHE = {}


def _solve_cached_grid(meds, N, nx, ny, order):
    """The grid polynomials depend on the shape only: computed once per (nx, ny, order)."""
    key = (nx, ny, order)
    if key in _GRID_CACHE:
        coef, _, _ = _solve(meds, N, nx, ny, order, want_grid=False)
        return (coef, *_GRID_CACHE[key])
    coef, LPX, LPY = _solve(meds, N, nx, ny, order)
    if len(_GRID_CACHE) > 8:
        _GRID_CACHE.clear()
    _GRID_CACHE[key] = (LPX, LPY)
    return coef, LPX, LPY


def medfit(arr, N=8, order=2, device=0):
    """
    Fits a low-order polynomial to a 2D array (medians of N x N regions; see the reference for the coefficient order).

    Returns ``(coef, arrmed)`` with ``arrmed`` of the shape and dtype of ``arr``.
    """
    a = np.ascontiguousarray(arr, dtype=np.float32)
    if arr.dtype != np.float32:
        raise TypeError("medfit on the GPU takes float32 images (the L2 slope and the noise differences are float32)")
    ny, nx = a.shape
    lib = _lib.lib()
    meds = np.empty((N, N), np.float32)
    _lib.check(lib.rip_medfit_host(device, _lib.ptr(a), ny, nx, N, _lib.ptr(meds)))
    x, LPX, LPY = _solve_cached_grid(meds, N, nx, ny, order)
    model = np.empty((ny, nx), np.float32)
    dm = C.c_void_p()
    _lib.check(lib.rip_dev_alloc(device, C.byref(dm), model.nbytes))
    try:
        _lib.check(lib.rip_medfit_eval_dev(device, ny, nx, order, _lib.ptr(np.ascontiguousarray(x)), _lib.ptr(LPX),
                                           _lib.ptr(LPY), dm, None, nx, None))  # fmt: skip
        _lib.check(lib.rip_copy_d2h(device, _lib.ptr(model), dm, model.nbytes, None))
        _lib.check(lib.rip_device_sync(device))
    finally:
        lib.rip_dev_free(device, dm)
    return x, model


def medfit_device(d_arr, pitch, ny, nx, N=8, order=2, device=0, stream=None, subtract=True, d_model=None):
    """``medfit`` on a device-resident float32 window (pointer ``d_arr``, row pitch in elements); subtracts the model in
    place (``slope[nb:-nb, nb:-nb] -= skymodel``, gen_cal_image.py:645-647) and/or writes it to ``d_model``."""
    lib = _lib.lib()
    meds = np.empty((N, N), np.float32)
    dmeds = C.c_void_p()
    _lib.check(lib.rip_dev_alloc(device, C.byref(dmeds), meds.nbytes))
    try:
        _lib.check(lib.rip_block_nanmedian_dev(device, C.c_void_p(d_arr), pitch, ny, nx, N, dmeds, C.c_void_p(stream or None)))
        _lib.check(lib.rip_copy_d2h(device, _lib.ptr(meds), dmeds, meds.nbytes, C.c_void_p(stream or None)))
        _lib.check(lib.rip_stream_sync(device, C.c_void_p(stream or None)))
    finally:
        lib.rip_dev_free(device, dmeds)
    x, LPX, LPY = _solve_cached_grid(meds, N, nx, ny, order)
    _lib.check(lib.rip_medfit_eval_dev(device, ny, nx, order, _lib.ptr(np.ascontiguousarray(x)), _lib.ptr(LPX), _lib.ptr(LPY),
                                       C.c_void_p(d_model or None), C.c_void_p(d_arr if subtract else None), pitch,
                                       C.c_void_p(stream or None)))  # fmt: skip
    return x


def percentiles_device(d_arr, count, qs, device=0, stream=None):
    """``np.percentile(arr, q)`` (linear interpolation) for a flat device-resident float32 array: the two bracketing
    order statistics of every q come from ``rip_order_stats_dev``; NaN anywhere gives NaN, as in NumPy."""
    lib = _lib.lib()
    ranks, fr = [], []
    for q in qs:
        vi = (count - 1) * (float(q) / 100.0)
        lo = int(np.floor(vi))
        ranks += [lo, min(lo + 1, count - 1)]
        fr.append(vi - lo)
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    out = np.empty(len(ranks), np.float32)
    nv = C.c_long(0)
    _lib.check(lib.rip_order_stats_dev(device, C.c_void_p(d_arr), count, len(ranks), _lib.ptr(r), _lib.ptr(out), C.byref(nv),
                                       C.c_void_p(stream or None)))  # fmt: skip
    if nv.value != count:
        return [np.float32(np.nan)] * len(qs)
    res = []
    for i, t in enumerate(fr):
        a, b = np.float64(out[2 * i]), np.float64(out[2 * i + 1])
        v = a + (b - a) * t if t < 0.5 else b - (b - a) * (1 - t)  # numpy's _lerp
        res.append(np.float32(v))
    return res


def binkxk(arr, k, mask=None, device=0):
    """Bin-averaging utility for 2D arrays, k x k (reference utils/sky.py:20-43), on the GPU; remainder pixels are
    ignored.  ``mask`` (bool, same shape): pixels to treat as NaN, i.e. ``binkxk(np.where(~mask, arr, nan), k)`` as the
    driver calls it (gen_cal_image.py:641)."""
    a = np.ascontiguousarray(arr, dtype=np.float32)
    ny, nx = a.shape
    lib = _lib.lib()
    out = np.empty((ny // k, nx // k), np.float32)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    bufs = []
    try:
        da, dm, do = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(lib.rip_dev_alloc(device, C.byref(da), a.nbytes)); bufs.append(da)  # noqa: E702
        _lib.check(lib.rip_copy_h2d(device, da, _lib.ptr(a), a.nbytes, None))
        if m is not None:
            _lib.check(lib.rip_dev_alloc(device, C.byref(dm), m.nbytes)); bufs.append(dm)  # noqa: E702
            _lib.check(lib.rip_copy_h2d(device, dm, _lib.ptr(m), m.nbytes, None))
        _lib.check(lib.rip_dev_alloc(device, C.byref(do), out.nbytes)); bufs.append(do)  # noqa: E702
        _lib.check(lib.rip_bin_masked_dev(device, da, dm if m is not None else None, ny, nx, int(k), do, None))
        _lib.check(lib.rip_copy_d2h(device, _lib.ptr(out), do, out.nbytes, None))
        _lib.check(lib.rip_device_sync(device))
    finally:
        for b in bufs:
            lib.rip_dev_free(device, b)
    return out


def _norm_ppf(p):
    """Inverse of the standard normal distribution function (Acklam's rational approximation refined by one Halley
    step: |error| < 1e-15 on (0, 1)); ``scipy.stats.norm.ppf`` in the reference (utils/sky.py:74)."""
    import math  # noqa: PLC0415

    if not 0.0 < p < 1.0:
        return math.nan
    a = (-3.969683028665376e01, 2.209460984245205e02, -2.759285104469687e02, 1.383577518672690e02, -3.066479806614716e01, 2.506628277459239e00)  # fmt: skip
    b = (-5.447609879822406e01, 1.615858368580409e02, -1.556989798598866e02, 6.680131188771972e01, -1.328068155288572e01)
    c = (-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e00, -2.549732539343734e00, 4.374664141464968e00, 2.938163982698783e00)  # fmt: skip
    d = (7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e00, 3.754408661907416e00)
    if p < 0.02425:
        q = math.sqrt(-2 * math.log(p))
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1)
    elif p > 1 - 0.02425:
        q = math.sqrt(-2 * math.log(1 - p))
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1)
    else:
        q = p - 0.5
        r = q * q
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1)
    e = 0.5 * math.erfc(-x / math.sqrt(2)) - p
    u = e * math.sqrt(2 * math.pi) * math.exp(x * x / 2)
    return x - u / (1 + x * u / 2)


def smooth_mode(arr, pc=25.0, pksmooth=0.5, niter=3, device=0):
    """
    Mode of the smoothed histogram of ``arr`` (NaNs ignored): returns ``(mode, width of the weighting function)`` like the
    reference (utils/sky.py:46-93).  Start: centre = median, sigma = interquantile range / its Gaussian value; then
    ``niter`` times: Gaussian-weighted counts at 19 points within +-sigma of the centre (``rip_gauss_hist_dev``), parabola
    through the highest one and its neighbours, centre <- its vertex.
    """
    a = np.ascontiguousarray(arr, dtype=np.float32).ravel()
    lib = _lib.lib()
    da = C.c_void_p()
    _lib.check(lib.rip_dev_alloc(device, C.byref(da), a.nbytes))
    try:
        _lib.check(lib.rip_copy_h2d(device, da, _lib.ptr(a), a.nbytes, None))
        c1, c2, c3 = nanpercentiles_device(da.value, a.size, [pc, 50.0, 100.0 - pc], device)
        ctr = float(c2)
        sigma = (float(c3) - float(c1)) / (_norm_ppf((100.0 - pc) / 100.0) * 2)
        nz = 21
        for _ in range(niter):
            z = ctr + np.linspace(-1, 1, nz) * sigma
            sums = np.zeros(nz - 2, np.float64)
            zin = np.ascontiguousarray(z[1:-1])
            _lib.check(lib.rip_gauss_hist_dev(device, da, a.size, _lib.ptr(zin), nz - 2, float(pksmooth * sigma), _lib.ptr(sums), None))
            hist = np.zeros(nz)
            hist[1:-1] = sums
            ip = int(np.argmax(hist))
            lo, hi = hist[ip - 1], hist[ip + 1]
            slope, curv = (hi - lo) / 2.0, (hi + lo) / 2.0 - hist[ip]
            ctr = z[ip] + (z[1] - z[0]) * (-slope / 2.0 / curv)
    finally:
        lib.rip_dev_free(device, da)
    return (ctr, sigma * pksmooth)


def nanpercentiles_device(d_arr, count, qs, device=0, stream=None):
    """``np.nanpercentile(arr, q)`` for a flat device-resident float32 array (linear interpolation between the bracketing
    order statistics of the non-NaN elements; all-NaN input gives NaN)."""
    lib = _lib.lib()
    probe = np.zeros(1, np.int64)
    out1 = np.empty(1, np.float32)
    nv = C.c_long(0)
    _lib.check(lib.rip_order_stats_dev(device, C.c_void_p(d_arr), count, 1, _lib.ptr(probe), _lib.ptr(out1), C.byref(nv),
                                       C.c_void_p(stream or None)))  # fmt: skip
    n = nv.value
    if n == 0:
        return [np.float32(np.nan)] * len(qs)
    ranks, fr = [], []
    for q in qs:
        vi = (n - 1) * (float(q) / 100.0)
        lo = int(np.floor(vi))
        ranks += [lo, min(lo + 1, n - 1)]
        fr.append(vi - lo)
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    out = np.empty(len(ranks), np.float32)
    _lib.check(lib.rip_order_stats_dev(device, C.c_void_p(d_arr), count, len(ranks), _lib.ptr(r), _lib.ptr(out), C.byref(nv),
                                       C.c_void_p(stream or None)))  # fmt: skip
    res = []
    for i, t in enumerate(fr):
        lo_v, hi_v = np.float64(out[2 * i]), np.float64(out[2 * i + 1])
        res.append(np.float32(lo_v + (hi_v - lo_v) * t if t < 0.5 else hi_v - (hi_v - lo_v) * (1 - t)))
    return res
