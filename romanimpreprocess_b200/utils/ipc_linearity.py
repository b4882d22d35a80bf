"""IPC and linearity utilities on the GPU: drop-in for ``romanimpreprocess.utils.ipc_linearity``.

Classes
-------
IL
    IPC + inverse linearity forward model object (romanisim ``inv_linearity`` protocol: ``apply``, ``dq``, ``set_dq``).

Functions
---------
ipc_fwd, ipc_rev, correct_cube
    3x3-per-pixel IPC convolution, order-n deconvolution, in-place cube correction
    (reference utils/ipc_linearity.py:37,102,145).
_lin, linearity, multilin, invlinearity
    Legendre linearity, its multi-group form with DQ, and the 24-step bisection inverse
    (reference utils/ipc_linearity.py:192,234,276,347).

Every function keeps the reference's signature, return values, dtype promotion and in-place behaviour; the
arithmetic runs in hand-written CUDA kernels (``csrc/rip_stage.cu``) through the C ABI.  There is no CPU path.
"""

import ctypes as C
import sys

import numpy as np

from .. import _lib
from ..caltree import open_tree
from ..dqflags import pixel


def _float_in(a):
    """float32/float64 C-contiguous view of an input image (ints follow NumPy: int32+f32 -> f64)."""
    a = np.asarray(a)
    if a.dtype in (np.float32, np.float64):
        return np.ascontiguousarray(a)
    if a.dtype.kind in "iu" and a.dtype.itemsize >= 4:
        return np.ascontiguousarray(a, dtype=np.float64)
    return np.ascontiguousarray(a, dtype=np.float32)


def _np_dtype(tag):
    return np.float64 if tag == _lib.RIP_F64 else np.float32


## IPC utilities ##


def ipc_fwd(image, kernel, gain=None, device=0):
    """
    Carries out an IPC operation on the image: ``out[y,x] = sum_{dy,dx} in[y-dy,x-dx] K[1+dy,1+dx,y-dy,x-dx]``.

    ``image`` (ny,nx); ``kernel`` (3,3,ny,nx); optional ``gain`` (ny,nx) to work in DN (g^-1 K g).
    """
    img = _float_in(image)
    K = _lib.as_float_plane(kernel)
    g = None if gain is None else _lib.as_float_plane(np.broadcast_to(gain, img.shape))
    ny, nx = img.shape
    out = np.empty((ny, nx), np.float64)
    tag = C.c_int(0)
    _lib.check(
        _lib.lib().rip_ipc_fwd(device, _lib.ptr(img), _lib.float_tag(img), ny, nx, _lib.ptr(K), _lib.float_tag(K),
                               _lib.ptr(g), 0 if g is None else _lib.float_tag(g), _lib.ptr(out), C.byref(tag))
    )  # fmt: skip
    dt = _np_dtype(tag.value)
    return out.view(dt).reshape(-1)[: ny * nx].reshape(ny, nx).copy() if dt == np.float32 else out


def ipc_rev(image, kernel, order=2, gain=None, device=0):
    """
    Inverse IPC operation to the given order (footprint ``(2*order+1)^2``); with ``gain`` does g^-1 K^-1 g.
    """
    img = _float_in(image)
    K = _lib.as_float_plane(kernel)
    g = None if gain is None else _lib.as_float_plane(np.broadcast_to(gain, img.shape))
    ny, nx = img.shape
    out = np.empty((ny, nx), np.float64)
    tag = C.c_int(0)
    _lib.check(
        _lib.lib().rip_ipc_rev(device, _lib.ptr(img), _lib.float_tag(img), ny, nx, _lib.ptr(K), _lib.float_tag(K),
                               int(order), _lib.ptr(g), 0 if g is None else _lib.float_tag(g), _lib.ptr(out),
                               C.byref(tag))
    )  # fmt: skip
    dt = _np_dtype(tag.value)
    return out.view(dt).reshape(-1)[: ny * nx].reshape(ny, nx).copy() if dt == np.float32 else out


def correct_cube(data, ipc_file, mylog, gain_file=None, device=0):
    """
    IPC corrects a full data cube (``data``, shape (ngrp,ny,nx), float32) in place.

    Operates in electrons if ``gain_file`` is None, in DN if it is provided.  Reference pixels are untouched.
    """
    if ipc_file is None:
        if mylog is not None:
            mylog.append("No IPC file specified, skipping ...\n")
        return
    with open_tree(ipc_file) as F:
        kernel = _lib.as_float_plane(F["roman"]["data"])
    if mylog is not None:
        mylog.append(f"IPC kernel center range --> {np.amin(kernel[1,1,:,:]):f},{np.amax(kernel[1,1,:,:]):f}\n")
    (ngrp, ny, nx) = np.shape(data)
    nb = (8192 + (nx - np.shape(kernel)[-1]) // 2) % 16
    if mylog is not None:
        mylog.append(f" ..., {ngrp:d} groups, excluding {nb:d} border pixels\n")
    g = None
    if gain_file is not None:
        with open_tree(gain_file) as G:
            g = _lib.as_float_plane(G["roman"]["data"])
    if data.dtype != np.float32:
        raise TypeError("correct_cube works in place on float32 cubes")
    d = data if data.flags["C_CONTIGUOUS"] else np.ascontiguousarray(data)
    _lib.check(
        _lib.lib().rip_correct_cube(device, _lib.ptr(d), ngrp, ny, nx, _lib.ptr(kernel), _lib.float_tag(kernel),
                                    kernel.shape[-2], kernel.shape[-1], _lib.ptr(g),
                                    0 if g is None else _lib.float_tag(g))
    )  # fmt: skip
    if d is not data:
        data[...] = d


## LINEARITY UTILITIES ##


def _lin(z, coefs, linextrap=True, device=0):
    """
    Evaluates ``phi = sum_l coefs_l P_l(z)`` with linear extrapolation beyond |z|=1; returns ``(phi, exflag)``.

    ``z`` (ny,nx) float32/float64, ``coefs`` (p_order+1,ny,nx).  ``phi`` is float32 (the accumulator of the
    reference has the coefficient dtype; the calibration files store float32).
    """
    zz = _float_in(z)
    c = _lib.as_c(coefs, np.float32)
    npix = zz.size
    phi = np.empty(zz.shape, np.float32)
    ex = np.empty(zz.shape, np.uint8)
    _lib.check(
        _lib.lib().rip_lin_eval(device, _lib.ptr(zz), _lib.float_tag(zz), _lib.ptr(c), c.shape[0], npix,
                                1 if linextrap else 0, _lib.ptr(phi), _lib.ptr(ex))
    )  # fmt: skip
    return phi, ex.astype(bool)


def _lin_planes(F, ymin, ymax, xmin, xmax):
    r = F["roman"]
    return (
        _lib.as_c(r["data"][:, ymin:ymax, xmin:xmax], np.float32),
        _lib.as_c(r["Smin"][ymin:ymax, xmin:xmax], np.float32),
        _lib.as_c(r["Smax"][ymin:ymax, xmin:xmax], np.float32),
        _lib.as_c(r["Sref"][ymin:ymax, xmin:xmax], np.float32),
        _lib.as_c(r["dq"][ymin:ymax, xmin:xmax], np.uint32),
    )


def linearity(S, linearity_file, origin=(0, 0), device=0):
    """
    Performs a linearity correction of one 2D frame (DN_raw -> DN_lin); returns ``(Slin, dq)``.

    ``origin`` is the (x,y) position of the lower-left corner of ``S`` in the convention of the file.
    """
    (dy, dx) = np.shape(S)
    ymin, xmin = origin[1], origin[0]
    with open_tree(linearity_file) as F:
        coefs, Smin, Smax, Sref, dq0 = _lin_planes(F, ymin, ymin + dy, xmin, xmin + dx)
    s = _lib.as_c(S, np.float32)
    phi = np.empty((dy, dx), np.float32)
    dq = np.empty((dy, dx), np.uint32)
    _lib.check(
        _lib.lib().rip_multilin(device, _lib.ptr(s), 1, dy * dx, _lib.ptr(coefs), coefs.shape[0], _lib.ptr(Smin),
                                _lib.ptr(Smax), _lib.ptr(Sref), _lib.ptr(dq0), None, 0, 1, _lib.ptr(phi),
                                _lib.ptr(dq))
    )  # fmt: skip
    return phi, dq


def multilin(S, linearity_file, origin=(0, 0), do_not_flag_first=True, attempt_corr=None, device=0):
    """
    Performs a linearity correction with multiple groups; returns ``(Slin (ngrp,ny,nx) float32, dq (ny,nx) uint32)``.

    ``attempt_corr``: array like ``S`` that is truthy where an out-of-range group should be flagged NO_LIN_CORR
    (the driver passes ``~rdq & SATURATED``).  ``do_not_flag_first``: clip and never flag the reset-read group.
    """
    (ngrp, dy, dx) = np.shape(S)
    if ngrp > _lib.RIP_GMAX:
        raise ValueError(f"multilin on the GPU supports up to {_lib.RIP_GMAX} groups")
    ymin, xmin = origin[1], origin[0]
    with open_tree(linearity_file) as F:
        coefs, Smin, Smax, Sref, dq0 = _lin_planes(F, ymin, ymin + dy, xmin, xmin + dx)
    s = _lib.as_c(S, np.float32)
    att = None if attempt_corr is None else np.ascontiguousarray(np.asarray(attempt_corr) != 0, dtype=np.uint8)
    phi = np.empty((ngrp, dy, dx), np.float32)
    dq = np.empty((dy, dx), np.uint32)
    _lib.check(
        _lib.lib().rip_multilin(device, _lib.ptr(s), ngrp, dy * dx, _lib.ptr(coefs), coefs.shape[0],
                                _lib.ptr(Smin), _lib.ptr(Smax), _lib.ptr(Sref), _lib.ptr(dq0), _lib.ptr(att),
                                1 if do_not_flag_first else 0, 0, _lib.ptr(phi), _lib.ptr(dq))
    )  # fmt: skip
    return phi, dq


def invlinearity(Slin, linearity_file, origin=(0, 0), device=0):
    """
    Calculates the inverse linearity (DN_lin -> DN_raw) by 24 bisection steps; returns ``(S, exflag)``.

    The result has the dtype of ``Slin`` (float64 inside ``IL.apply``), as in the reference.
    """
    (dy, dx) = np.shape(Slin)
    ymin, xmin = origin[1], origin[0]
    with open_tree(linearity_file) as F:
        coefs, Smin, Smax, _, _ = _lin_planes(F, ymin, ymin + dy, xmin, xmin + dx)
    s = _float_in(Slin)
    out = np.empty((dy, dx), s.dtype)
    ex = np.empty((dy, dx), np.uint8)
    _lib.check(
        _lib.lib().rip_invlinearity(device, _lib.ptr(s), _lib.float_tag(s), dy * dx, _lib.ptr(coefs), coefs.shape[0],
                                    _lib.ptr(Smin), _lib.ptr(Smax), _lib.ptr(out), _lib.ptr(ex))
    )  # fmt: skip
    return out, ex.astype(bool)


class IL:
    """
    IPC + inverse linearity forward model, API-compatible with the reference's ``IL`` so that it can be handed to
    ``romanisim.l1.apportion_counts_to_resultants(..., inv_linearity=IL(...))``.

    Parameters
    ----------
    linearity_file, gain_file : str or tree
        Calibration reference files (ASDF names or in-memory trees).
    ipc_file : str, tree or None
        ipc4d file; None skips the IPC.
    start_e : np.array or float, optional
        Electrons already in the well (reset noise).
    """

    def __init__(self, linearity_file, gain_file, ipc_file, start_e=0.0, device=0):
        self.linearity_file = linearity_file
        self.gain_file = gain_file
        self.ipc_file = ipc_file
        self.start_e = start_e
        self.device = device
        with open_tree(self.linearity_file) as f:
            self._dq = np.copy(f["roman"]["dq"])

    def set_dq(self, ngroup=1, nborder=4):
        """Sets the 3D data quality flags ``self.dq`` (ngroup, ny-2*nborder, nx-2*nborder)."""
        (ny, nx) = np.shape(self._dq)
        self.dq = np.zeros((ngroup, ny - 2 * nborder, nx - 2 * nborder), dtype=np.uint32)
        self.dq[:, :, :] = self._dq[None, nborder : ny - nborder, nborder : nx - nborder]

    def apply(self, counts, electrons=False, electrons_out=False):
        """
        Converts a linearized signal to a non-linear, IPC-convolved signal (DN_raw, or electrons if
        ``electrons_out``).  ``counts`` 2D; ``electrons`` says whether the input is electrons or DN_lin.

        One GPU call (``rip_il_apply_planes``): counts + start_e -> ipc_fwd -> /gain -> 24-step bisection, with
        NumPy's dtype promotion (int32 counts + float32 start_e run in float64: SURVEY App. A10).
        """
        print("apply", electrons, electrons_out, np.shape(counts))
        sys.stdout.flush()
        cnt = np.asarray(counts)
        if cnt.dtype.kind in "iu":
            cnt = np.ascontiguousarray(cnt, dtype=np.int32)
            ctag = _lib.RIP_I32
        else:
            cnt = _lib.as_float_plane(cnt)
            ctag = _lib.float_tag(cnt)
        (nyc, nxc) = cnt.shape
        start = None
        start_scalar = 0.0
        stag = _lib.RIP_F32
        if np.ndim(self.start_e) == 0:
            start_scalar = float(self.start_e)
        else:
            start = _lib.as_float_plane(np.broadcast_to(self.start_e, cnt.shape))
            stag = _lib.float_tag(start)
        kernel = None
        if self.ipc_file is not None:
            with open_tree(self.ipc_file) as f:
                kernel = _lib.as_float_plane(f["roman"]["data"])
        g = None
        if electrons or electrons_out:
            with open_tree(self.gain_file) as f:
                g = np.asarray(f["roman"]["data"])
                (nyg, nxg) = np.shape(g)
                if nyg > nyc:
                    nbg = (nyg - nyc) // 2
                    g = g[nbg:-nbg, nbg:-nbg]
            g = _lib.as_float_plane(g)
        nb = (8192 - nyc // 2) % 16
        with open_tree(self.linearity_file) as F:
            coefs, Smin, Smax, Sref, _ = _lin_planes(F, nb, nb + nyc, nb, nb + nxc)
        out = np.empty((nyc, nxc), np.float64)
        tag = C.c_int(0)
        _lib.check(
            _lib.lib().rip_il_apply_planes(self.device, _lib.ptr(cnt), ctag, nyc, nxc, _lib.ptr(start), stag,
                                           start_scalar, _lib.ptr(kernel),
                                           0 if kernel is None else _lib.float_tag(kernel), _lib.ptr(g),
                                           0 if g is None else _lib.float_tag(g), _lib.ptr(coefs), coefs.shape[0],
                                           _lib.ptr(Smin), _lib.ptr(Smax), _lib.ptr(Sref), 1 if electrons else 0,
                                           1 if electrons_out else 0, _lib.ptr(out), C.byref(tag))
        )  # fmt: skip
        return out if tag.value == _lib.RIP_F64 else out.astype(np.float32)


__all__ = ["IL", "ipc_fwd", "ipc_rev", "correct_cube", "_lin", "linearity", "multilin", "invlinearity", "pixel"]
