"""Growing a bitmask into a boolean pixel mask on the GPU: drop-in for ``romanimpreprocess.utils.maskhandling``.

``CombinedMask.build`` (reference utils/maskhandling.py:82-117) ORs, for every flagged bit, the bit plane dilated by
1 / 5 (cross) / 9 (3x3) / 25 (5x5) pixels with zero padding at the edges (``scipy.signal.convolve(mode="same")``).
One kernel (``rip_mask_build_host``) does all bits at once: bits are grouped by footprint and each pixel ORs the
selected bits of its 5x5 neighbourhood.  Integer work: bit-exact against the reference.
"""

import numpy as np

from .. import _lib
from ..dqflags import pixel


class CombinedMask:
    """Boolean mask from multiple flags; ``maskdict`` maps a flag name (or bit number) to 1, 5, 9 or 25."""

    def __init__(self, maskdict):
        self.array = np.zeros(32, dtype=np.uint8)
        for d in maskdict:
            if isinstance(d, (int, np.integer)):
                whichbit = int(d)
            else:
                e = int(getattr(pixel, str(d).upper()))
                whichbit = 0
                for x in range(32):
                    if e >> x == 1:
                        whichbit = x
            if int(maskdict[d]) not in (0, 1, 5, 9, 25):
                raise KeyError(int(maskdict[d]))  # the reference fails with KeyError in its kernel dictionary
            self.array[whichbit] = int(maskdict[d])

    def build(self, dq, device=0):
        """2D uint32 data-quality array -> boolean mask (True = masked)."""
        dq = np.ascontiguousarray(dq, dtype=np.uint32)
        ny, nx = dq.shape
        out = np.empty((ny, nx), np.uint8)
        _lib.check(_lib.lib().rip_mask_build_host(device, _lib.ptr(dq), ny, nx, _lib.ptr(self.array), _lib.ptr(out)))
        return out.astype(bool)

    def convert_file(self, file_in, file_mask, device=0):
        """Stand-alone function to make a mask from an L2 file (reference utils/maskhandling.py:119-149; the third call of
        the production loop, runs/summer2025run/OpenUniverse_to_L1L2.py:165).  ``.asdf``: the boolean array under ``mask``;
        ``.fits``: a masked image (HDU0: data with -1000 where masked, for display) and an int8 version (HDU1, ``MASK``)."""
        from ..caltree import open_tree, write_tree  # noqa: PLC0415
        from ..io import fits_lite  # noqa: PLC0415

        with open_tree(file_in) as f_in:
            dq = np.asarray(f_in["roman"]["dq"])
            data = np.asarray(f_in["roman"]["data"])
        locmask = self.build(dq, device=device)
        if file_mask[-5:] == ".asdf":
            write_tree(file_mask, {"mask": locmask})
        elif file_mask[-5:] == ".fits":
            fits_lite.write_hdus(file_mask, [(np.where(locmask, -1000.0, data).astype(np.float32), None),
                                             (np.where(locmask, 1, 0).astype(np.int8), {"EXTNAME": "MASK"})])  # fmt: skip


# reference utils/maskhandling.py:152-178
PixelMask1 = CombinedMask(
    {
        "DO_NOT_USE": 1, "JUMP_DET": 5, "DROPOUT": 25, "GW_AFFECTED_DATA": 1, "PERSISTENCE": 1, "AD_FLOOR": 5,
        "UNRELIABLE_ERROR": 1, "NON_SCIENCE": 1, "DEAD": 9, "HOT": 9, "WARM": 1, "LOW_QE": 9, "TELEGRAPH": 1,
        "NO_FLAT_FIELD": 9, "NO_GAIN_VALUE": 9, "NO_LIN_CORR": 9, "NO_SAT_CHECK": 9, "UNRELIABLE_BIAS": 1,
        "UNRELIABLE_DARK": 9, "UNRELIABLE_SLOPE": 9, "UNRELIABLE_FLAT": 9, "UNRELIABLE_RESET": 9, "OTHER_BAD_PIXEL": 9,
    }
)  # fmt: skip
