"""DQ bit values used on the hot path.

Stand-in for ``roman_datamodels.dqflags.pixel`` (whose members are ``np.uint32``).
If ``roman_datamodels`` is installed its enum is used instead so the values can never drift.
"""

import numpy as np

try:  # pragma: no cover - only where the real stack is installed
    from roman_datamodels.dqflags import pixel  # type: ignore
except Exception:  # noqa: BLE001

    class pixel:  # noqa: N801
        """Subset of roman_datamodels.dqflags.pixel needed by the calibration path."""

        GOOD = np.uint32(0)
        DO_NOT_USE = np.uint32(2**0)
        SATURATED = np.uint32(2**1)
        JUMP_DET = np.uint32(2**2)
        DROPOUT = np.uint32(2**3)
        GW_AFFECTED_DATA = np.uint32(2**4)
        PERSISTENCE = np.uint32(2**5)
        AD_FLOOR = np.uint32(2**6)
        OUTLIER = np.uint32(2**7)
        UNRELIABLE_ERROR = np.uint32(2**8)
        NON_SCIENCE = np.uint32(2**9)
        DEAD = np.uint32(2**10)
        HOT = np.uint32(2**11)
        WARM = np.uint32(2**12)
        LOW_QE = np.uint32(2**13)
        TELEGRAPH = np.uint32(2**15)
        NONLINEAR = np.uint32(2**16)
        BAD_REF_PIXEL = np.uint32(2**17)
        NO_FLAT_FIELD = np.uint32(2**18)
        NO_GAIN_VALUE = np.uint32(2**19)
        NO_LIN_CORR = np.uint32(2**20)
        NO_SAT_CHECK = np.uint32(2**21)
        UNRELIABLE_BIAS = np.uint32(2**22)
        UNRELIABLE_DARK = np.uint32(2**23)
        UNRELIABLE_SLOPE = np.uint32(2**24)
        UNRELIABLE_FLAT = np.uint32(2**25)
        RESERVED_5 = np.uint32(2**26)
        RESERVED_6 = np.uint32(2**27)
        UNRELIABLE_RESET = np.uint32(2**28)
        RESERVED_7 = np.uint32(2**29)
        OTHER_BAD_PIXEL = np.uint32(2**30)
        REFERENCE_PIXEL = np.uint32(2**31)
