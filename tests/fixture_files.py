"""Fixture writer for the file-level tests: a synthetic CALDIR (``synth.make_caldir``), an L1 exposure and its FITSWCS
header written as real files in the layouts of SURVEY App. B (ASDF 1.x, one file per CALDIR key, ``roman`` branch),
with ``io.asdf_lite.write_file``.  Everything the reference's ``calibrateimage(config)`` reads."""

import os

import numpy as np

from romanimpreprocess_b200 import synth
from romanimpreprocess_b200.io import asdf_lite

L1_TAG = "asdf://stsci.edu/datamodels/roman/tags/wfi_science_raw-1.0.0"


def fits_header_text(header):
    """80-column card stream as ``astropy.io.fits.Header.tofile`` writes it (from_sim/sim_to_isim.py:987)."""
    cards = []
    for k, v in header.items():
        val = f"'{v:<8}'" if isinstance(v, str) else (f"{v:>20}" if isinstance(v, int) else f"{v!r:>20}")
        cards.append(f"{k:<8}= {val}".ljust(80)[:80])
    cards += ["COMMENT truth wcs from sim_to_isim".ljust(80), "END".ljust(80)]
    text = "".join(cards)
    return text + " " * (-len(text) % 2880)


def sim_header(n_active):
    """The WCS of the reference's test scene (tests/romanimpreprocess/test_workflow.py:62-83), scaled to the frame."""
    return {"CTYPE1": "RA---TAN-SIP", "CTYPE2": "DEC--TAN-SIP", "CRPIX1": (n_active + 1) / 2.0, "CRPIX2": (n_active + 1) / 2.0,
            "CD1_1": 3.0555555555555554e-05, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 3.0555555555555554e-05, "CRVAL1": 37.0,
            "CRVAL2": -20.0, "LONPOLE": 215.0, "A_ORDER": 2, "A_0_2": 2.0e-6, "A_1_1": -1.0e-6, "A_2_0": 3.0e-6,
            "B_ORDER": 2, "B_0_2": 1.4e-5, "B_1_1": -1.0e-5, "B_2_0": 3.0e-7}  # fmt: skip


def write_exposure(tmp, n=256, seed=31, read_pattern=None, p_order=10, ipc_dtype=np.float64, extract_ref=False):
    """Write CALDIR files + L1 file + FITSWCS text under ``tmp``; returns (config, cal, data, amp33, read_pattern)."""
    rp = read_pattern or synth.README_PATTERN
    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=p_order, gain_dtype=np.float32, ipc_dtype=ipc_dtype,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data, amp33, _ = synth.make_l1(cal, rp, seed=seed + 1, n_sources=25, cr_frac=0.01, bright=3.0)
    caldir = {}
    for key, tree in cal.items():
        fn = os.path.join(tmp, f"roman_wfi_{key}_TEST_SCA01.asdf")
        asdf_lite.write_file(fn, tree)
        caldir[key] = fn
    meta = asdf_lite.TaggedDict({"exposure": {"read_pattern": [list(g) for g in rp], "frame_time": synth.FRAME_TIME,
                                              "ma_table_name": "synthetic"},
                                 "instrument": {"name": "WFI", "detector": "WFI01", "optical_element": "F184"},
                                 "model_type": "ScienceRawModel"})  # fmt: skip
    roman = asdf_lite.TaggedDict({"meta": meta, "data": data, "amp33": amp33}, tag=L1_TAG)
    if extract_ref:  # EXTRACT_REF files (from_sim/sim_to_isim.py:711-730): data stored relative to a reference read
        off = 5000
        ref = data[0].astype(np.int32)
        roman["data"] = np.clip(data.astype(np.int32) - ref[None] + off, 0, 65535).astype(np.uint16)
        roman["reference_read"] = data[0].copy()
        meta["instrument"]["data_encoding_offset"] = off
    l1 = os.path.join(tmp, "sim_L1_F184_1_1.asdf")
    asdf_lite.write_file(l1, {"roman": roman})
    wcsfn = os.path.join(tmp, "sim_L1_F184_1_1_asdf_wcshead.txt")
    with open(wcsfn, "w") as f:
        f.write(fits_header_text(sim_header(n - 8)))
    config = {"IN": l1, "OUT": os.path.join(tmp, "sim_L2_F184_1_1.asdf"), "FITSWCS": wcsfn, "CALDIR": caldir,
              "RAMP_OPT_PARS": {"slope": 0.4, "gain": 1.8, "sigma_read": 7.0}, "SLICEOUT": True, "SKYORDER": 2}  # fmt: skip
    return config, cal, data, amp33, rp
