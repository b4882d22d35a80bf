"""Reference-pixel subtraction on the GPU: drop-in for ``romanimpreprocess.utils.reference_subtraction``.

Functions
---------
ref_subtraction_channel
    Channel-based reference subtraction (reference utils/reference_subtraction.py:16-74).
ref_subtraction_row
    Row-based reference subtraction (reference utils/reference_subtraction.py:77-125).

Both update ``image`` (float32, shape (n, n [+128])) in place and return it, like the reference.  The medians,
the selection of the centre value and the float64 subtraction run in CUDA kernels; only ``np.polyfit`` on the two
n-element median vectors (needed when ``slope`` is None) stays on the host.
"""

import numpy as np

from .. import _lib


def _as_image(image):
    if image.dtype != np.float32 or not image.flags["C_CONTIGUOUS"]:
        raise TypeError("image must be a C-contiguous float32 array (it is updated in place)")
    return image


def ref_subtraction_channel(image, channel_start=0, channel_end=128, use_ref_channel=False, device=0):
    """
    Fits a line through the median of the 4 bottom and 4 top reference rows of each 128-column channel and
    subtracts it from every row of the channel.  ``use_ref_channel`` also treats the reference output (channel 33).
    """
    if channel_start != 0 or channel_end != 128:
        raise ValueError("only the reference's default first channel [0:128] is supported")
    img = _as_image(image)
    n, ncols = img.shape
    n_channels = 33 if use_ref_channel else 32
    _lib.check(_lib.lib().rip_refsub_channel(device, _lib.ptr(img), n, ncols, n_channels))
    return image


def ref_subtraction_row(image, use_ref_channel=False, slope=None, device=0):
    """
    Subtracts ``m * (ref_median[row] - median(ref_medians))`` from every row, where the reference median is taken
    over the reference output (``use_ref_channel``) or the 8 side reference pixels, and ``m`` is ``slope`` or a
    fit of the science-row medians against the reference medians.
    """
    img = _as_image(image)
    n, ncols = img.shape
    ref_med = np.empty(n, np.float32)
    sci_med = np.empty(n, np.float32) if slope is None else None
    _lib.check(_lib.lib().rip_row_medians(device, _lib.ptr(img), n, ncols, 1 if use_ref_channel else 0,
                                          _lib.ptr(ref_med), _lib.ptr(sci_med)))  # fmt: skip
    if slope is None:
        m_med, _ = np.polyfit(ref_med, sci_med, 1)
    else:
        m_med = slope
    _lib.check(_lib.lib().rip_refsub_row_apply(device, _lib.ptr(img), n, ncols, float(m_med), _lib.ptr(ref_med)))
    return image
