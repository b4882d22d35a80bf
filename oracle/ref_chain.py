"""The L1->L2 chain with the REFERENCE'S OWN functions (``oracle/_ref``, built by ``oracle/build_ref.py``).

TEST INFRASTRUCTURE ONLY: used by ``bench.py --impl reference`` / ``cpu_baseline`` and by ``tests/``.

``calibrateimage`` itself (L1_to_L2/gen_cal_image.py) cannot be imported in this image (romancal, stcal,
roman_datamodels, gwcs, asdf are absent), but every function it calls that the reference OWNS can:
``reference_subtraction.ref_subtraction_row/_channel``, ``ipc_linearity.multilin / correct_cube``,
``fitting.construct_weights / ramp_fit`` (with ``jump_detect`` inside), ``flatutils.get_flat``.  ``l1_to_l2`` below is
``rip_oracle.l1_to_l2`` with those steps swapped for the reference's unmodified code; only the glue lines of
``calibrateimage`` between them and the three third-party steps (dq-init, saturation flagging, dark subtraction: romancal
/ stcal, restated in rip_oracle.py) stay restatements.  ``asdf.open`` is a stub over in-memory trees, as in
``tests/golden/make_golden.py``.
"""

import contextlib
import os
import sys
import types

import numpy as np

from . import rip_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_TREES = {}
_MODS = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "romanimpreprocess", "utils", "fitting.py"))


def modules():
    """Import the reference modules from oracle/_ref (once), with the asdf / roman_datamodels stubs in place."""
    global _MODS
    if _MODS is not None:
        return _MODS
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if "asdf" not in sys.modules:
        asdf = types.ModuleType("asdf")

        @contextlib.contextmanager
        def _open(name, *a, **k):
            yield _TREES[name]

        asdf.open = _open
        sys.modules["asdf"] = asdf
    elif not hasattr(sys.modules["asdf"], "_rip_trees"):
        real_open = sys.modules["asdf"].open

        @contextlib.contextmanager
        def _open2(name, *a, **k):  # a real asdf is installed: in-memory names are served here, files by asdf itself
            if name in _TREES:
                yield _TREES[name]
            else:
                with real_open(name, *a, **k) as f:
                    yield f

        sys.modules["asdf"].open = _open2
    sys.modules["asdf"]._rip_trees = _TREES
    if "roman_datamodels.dqflags" not in sys.modules:
        rdm, dqf = types.ModuleType("roman_datamodels"), types.ModuleType("roman_datamodels.dqflags")

        class pixel:  # noqa: N801  (the members the four modules use)
            GOOD, DO_NOT_USE, SATURATED, JUMP_DET = np.uint32(0), np.uint32(1), np.uint32(2), np.uint32(4)
            DROPOUT, GW_AFFECTED_DATA, PERSISTENCE, AD_FLOOR = np.uint32(8), np.uint32(16), np.uint32(32), np.uint32(64)
            OUTLIER, UNRELIABLE_ERROR, NON_SCIENCE, DEAD = np.uint32(128), np.uint32(256), np.uint32(512), np.uint32(1024)
            HOT, WARM, LOW_QE, TELEGRAPH = np.uint32(2048), np.uint32(4096), np.uint32(8192), np.uint32(32768)
            NONLINEAR, BAD_REF_PIXEL, NO_FLAT_FIELD, NO_GAIN_VALUE = np.uint32(65536), np.uint32(131072), np.uint32(262144), np.uint32(524288)
            NO_LIN_CORR, NO_SAT_CHECK, UNRELIABLE_BIAS, UNRELIABLE_DARK = np.uint32(1048576), np.uint32(2097152), np.uint32(4194304), np.uint32(8388608)
            UNRELIABLE_SLOPE, UNRELIABLE_FLAT, RESERVED_5, RESERVED_6 = np.uint32(16777216), np.uint32(33554432), np.uint32(67108864), np.uint32(134217728)
            UNRELIABLE_RESET, RESERVED_7, OTHER_BAD_PIXEL, REFERENCE_PIXEL = np.uint32(268435456), np.uint32(536870912), np.uint32(1073741824), np.uint32(2147483648)

        dqf.pixel = pixel
        rdm.dqflags = dqf
        sys.modules["roman_datamodels"], sys.modules["roman_datamodels.dqflags"] = rdm, dqf
    sys.path.insert(0, REF_DIR)
    try:
        from romanimpreprocess.utils import fitting, flatutils, ipc_linearity, reference_subtraction  # noqa: PLC0415
    finally:
        sys.path.remove(REF_DIR)
    _MODS = types.SimpleNamespace(fitting=fitting, flatutils=flatutils, ipc_linearity=ipc_linearity,
                                  reference_subtraction=reference_subtraction)  # fmt: skip
    return _MODS


class _Log:
    output = ""

    def append(self, s):
        pass


def _register(cal, tag):
    """CALDIR trees ({"key": roman-branch}) -> names the stubbed asdf.open serves."""
    names = {}
    for k, v in cal.items():
        names[k] = f"{tag}:{k}"
        _TREES[names[k]] = {"roman": v}
    return names


def _release(names):
    for v in names.values():
        _TREES.pop(v, None)


def l1_to_l2(data_u16, amp33_u16, cal, read_pattern, frame_time, area_factor, config=None, do_refpix=True):
    """``rip_oracle.l1_to_l2`` with every reference-owned step run by the reference's unmodified functions."""
    m = modules()
    names = _register(cal, f"ref{id(cal)}")
    try:

        def refpix_loop(data, amp33, dark_cube, read):  # glue of gen_cal_image.py:530-556 around the reference's two functions
            slope = orc.optimal_refout_slope(read)
            ns = data.shape[1]
            for j in range(data.shape[0]):
                image = np.zeros((ns, ns + orc.CHANNELWIDTH), dtype=np.float32)
                image[:, :ns] = data[j] - dark_cube[j]
                image[:, -orc.CHANNELWIDTH :] = amp33[j] - read["amp33"]["med"]
                image[:, -orc.CHANNELWIDTH :] -= np.median(image[:, -orc.CHANNELWIDTH :])
                image = m.reference_subtraction.ref_subtraction_row(image, use_ref_channel=True, slope=slope)
                image = m.reference_subtraction.ref_subtraction_channel(image, use_ref_channel=True)
                data[j] = image[:, :ns] + dark_cube[j]
            return data

        def multilin(data, lin, do_not_flag_first=True, attempt_corr=None):
            return m.ipc_linearity.multilin(data, names["linearitylegendre"], do_not_flag_first=do_not_flag_first,
                                            attempt_corr=attempt_corr)  # fmt: skip

        def correct_cube(data, kernel, gain_full=None):
            # (the dark-slope call of gen_cal_image.py:217-221 passes the same two files)
            m.ipc_linearity.correct_cube(data, names["ipc4d"], _Log(), gain_file=names["gain"])
            return data

        def construct_weights(u, meta, exclude_first=True):
            return m.fitting.construct_weights(u, meta, exclude_first=exclude_first)

        def ramp_fit(data, rdq, pdq, meta, gain, read, exclude_first=True):
            # do_ramp_fit (gen_cal_image.py:444-452) hands fitting.ramp_fit the CALDIR (file names) and the log
            return m.fitting.ramp_fit(data, rdq, pdq, meta, names, _Log(), exclude_first=exclude_first)

        def get_flat(flat_full, gain_full, ipc_kernel, nborder, pdq, ipc_deconvolve=True):
            caldir = {k: names[k] for k in ("flat", "gain", "ipc4d") if k in names}
            return m.flatutils.get_flat(caldir, {"nborder": nborder}, pdq)

        fns = {"multilin": multilin, "correct_cube": correct_cube, "construct_weights": construct_weights,
               "ramp_fit": ramp_fit, "get_flat": get_flat}  # fmt: skip
        if data_u16.shape[-1] == 4096:
            # the reference's two subtraction functions index rows / columns 0..4095 literally (reference_subtraction.py:
            # 107): smaller frames (bounded samples) run the size-agnostic restatement rip_oracle.refpix_loop, which the
            # refsub_4096 golden pins bit for bit to these functions
            fns["refpix_loop"] = refpix_loop
        return orc.l1_to_l2(data_u16, amp33_u16, cal, read_pattern, frame_time, area_factor, config, do_refpix, fns=fns)
    finally:
        _release(names)
