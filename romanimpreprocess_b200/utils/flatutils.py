"""Flat-field utilities on the GPU: drop-in for ``romanimpreprocess.utils.flatutils``.

Functions
---------
get_flat
    Flat field in DN units, padded/clipped/flagged and IPC-deconvolved (reference utils/flatutils.py:20-76).
"""

import numpy as np

from .. import _lib
from ..caltree import open_tree


def get_flat(caldir, meta, pdq, ipc_deconvolve=True, device=0):
    """
    Gets the flat field in DN, including IPC deconvolution if requested.

    ``pdq`` (uint32, 2D) is updated in place with NO_FLAT_FIELD / NO_GAIN_VALUE; it may be None.
    Returns the float32 flat image.
    """
    nborder = int(meta["nborder"])
    with open_tree(caldir["flat"]) as f:
        flat = _lib.as_c(f["roman"]["data"], np.float32)
    n = flat.shape[0]
    gain = kernel = None
    if ipc_deconvolve:
        with open_tree(caldir["gain"]) as f:
            gain = _lib.as_float_plane(f["roman"]["data"])
        with open_tree(caldir["ipc4d"]) as f:
            kernel = _lib.as_float_plane(f["roman"]["data"])
    p = None if pdq is None else _lib.as_c(pdq, np.uint32)
    out = np.empty((n, n), np.float32)
    _lib.check(
        _lib.lib().rip_get_flat(device, _lib.ptr(flat), n, nborder, _lib.ptr(gain),
                                0 if gain is None else _lib.float_tag(gain), _lib.ptr(kernel),
                                0 if kernel is None else _lib.float_tag(kernel), _lib.ptr(p),
                                1 if ipc_deconvolve else 0, _lib.ptr(out))
    )  # fmt: skip
    if pdq is not None and p is not pdq:
        pdq[...] = p
    return out
