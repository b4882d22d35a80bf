"""Sky estimation on the GPU: drop-in for ``romanimpreprocess.utils.sky.medfit`` (reference utils/sky.py:98-190).

``medfit`` fits a low-order 2D Legendre polynomial to the medians of N x N regions.  The two passes over the image run
on the device: the region medians (``rip_block_nanmedian_dev``: exact order statistics by radix select, NaNs ignored)
and the evaluation of the model on the pixel grid (``rip_medfit_eval_dev``: float64 sum in the reference's term order,
cast to the image dtype).  The (order+1)(order+2)/2 normal equations in between are the reference's own NumPy lines
(64 numbers in, 6 out).  Bit-exact against the reference for float32 images.
"""

import ctypes as C

import numpy as np
from scipy.special import legendre_p

from .. import _lib


def binkxk(arr, k):
    """Bin-averaging utility for 2D array, kxk (reference utils/sky.py:20-42; host NumPy: one cheap reduction)."""
    (ny, nx) = np.shape(arr)
    nyo, nxo = ny // k, nx // k
    return np.mean(arr[: k * nyo, : k * nxo].reshape((nyo, k, nxo, k)), axis=(1, 3))


def _normal_equations(meds, N, nx, ny, order):
    """Reference utils/sky.py:137-165, verbatim arithmetic: centres of the regions, basis, A x = b."""
    kx, ky = nx // N, ny // N
    px, py = (nx % N) // 2, (ny % N) // 2
    u_ = 2 * (px - 0.5 + kx * np.linspace(0.5, N - 0.5, N)) / nx - 1
    v_ = 2 * (py - 0.5 + ky * np.linspace(0.5, N - 0.5, N)) / ny - 1
    u, v = np.meshgrid(u_, v_)
    nc = (order + 1) * (order + 2) // 2
    basis = np.zeros((nc, N, N))
    k = 0
    for i in range(order + 1):
        temp = legendre_p(i, u)
        for j in range(order + 1 - i):
            basis[k, :, :] = temp * legendre_p(j, v)
            k += 1
    A = np.zeros((nc, nc))
    b = np.zeros(nc)
    for ipix in range(N):
        for jpix in range(N):
            if not np.isnan(meds[jpix, ipix]):
                A += np.outer(basis[:, jpix, ipix], basis[:, jpix, ipix])
                b += meds[jpix, ipix] * basis[:, jpix, ipix]
    return np.linalg.solve(A, b)


_GRID_CACHE = {}


def _grid_polynomials(nx, ny, order):
    """Legendre polynomials on the pixel grid (reference utils/sky.py:167-175); they depend on the shape only."""
    key = (nx, ny, order)
    if key not in _GRID_CACHE:
        if len(_GRID_CACHE) > 8:
            _GRID_CACHE.clear()
        _GRID_CACHE[key] = _grid_polynomials_uncached(nx, ny, order)
    return _GRID_CACHE[key]


def _grid_polynomials_uncached(nx, ny, order):
    LPX = np.zeros((order + 1, nx))
    LPY = np.zeros((order + 1, ny))
    u_ = np.linspace(-1, 1 - 2 / nx, nx)
    v_ = np.linspace(-1, 1 - 2 / ny, ny)
    for i in range(order + 1):
        LPX[i, :] = legendre_p(i, u_)
    for j in range(order + 1):
        LPY[j, :] = legendre_p(j, v_)
    return LPX, LPY


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
    x = _normal_equations(meds, N, nx, ny, order)
    LPX, LPY = _grid_polynomials(nx, ny, order)
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
    x = _normal_equations(meds, N, nx, ny, order)
    LPX, LPY = _grid_polynomials(nx, ny, order)
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
