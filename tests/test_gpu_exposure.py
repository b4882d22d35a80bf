"""GPU: whole exposures through ``L1_to_L2.exposure_driver.ExposureCalibrator`` (BASELINE configs[3]: the per-SCA loop of
the reference's runs/summer2025run/OpenUniverse_to_L1L2.py:155-169, one resident CALDIR + pipeline per SCA of a rank):
every (exposure, SCA) item equals the synchronous single-call path on the same arrays, for both ranks of a 2-rank split."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_exposure_driver_matches_the_single_call_path():
    from romanimpreprocess_b200 import _lib, pars, sharding, synth
    from romanimpreprocess_b200.L1_to_L2 import exposure_driver as xd
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
    from romanimpreprocess_b200.utils import coordutils

    n, rp = 256, synth.README_PATTERN
    scas = (1, 2, 3)
    cals = {s: synth.make_caldir(n=n, seed=40 + s, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32,
                                 sprinkle_flags=True, biascorr_amp=3.0) for s in scas}  # fmt: skip
    l1 = {(e, s): synth.make_l1(cals[s], rp, seed=100 + 10 * e + s, n_sources=9, cr_frac=0.01, bright=3.0)[:2]
          for e in range(2) for s in scas}  # fmt: skip
    hdr = {"CTYPE1": "RA---TAN-SIP", "CTYPE2": "DEC--TAN-SIP", "CRPIX1": (n - 7) / 2.0, "CRPIX2": (n - 7) / 2.0,
           "CD1_1": 3.0555555555555554e-05, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 3.0555555555555554e-05, "CRVAL1": 37.0,
           "CRVAL2": -20.0, "LONPOLE": 215.0, "A_ORDER": 2, "A_2_0": 3.0e-6, "B_ORDER": 2, "B_0_2": 1.4e-5}  # fmt: skip
    wcs = {it: coordutils.FitsWCS(dict(hdr, CRVAL1=37.0 + 0.01 * it[0], CRVAL2=-20.0 + 0.02 * it[1])) for it in l1}
    cfg = {"SLICEOUT": True}
    items = sorted(l1)
    got = {}
    for rank in range(2):  # both halves of a 2-rank job, one after the other on this GPU
        outs = [{"slope": _lib.pinned_empty((n, n), np.float32), "err_read": _lib.pinned_empty((n, n), np.float32),
                 "err_poisson": _lib.pinned_empty((n, n), np.float32), "pdq": _lib.pinned_empty((n, n), np.uint32),
                 "endslice": _lib.pinned_empty((n - 8, n - 8), np.int8)} for _ in range(3)]  # fmt: skip
        with xd.ExposureCalibrator(cals, items, rp, synth.FRAME_TIME, cfg, rank=rank, world=2, device=0, depth=2) as drv:
            assert drv.items == sharding.assign_items_balanced(items, rank, 2)

            def sink(e, s, out):
                got[(e, s)] = {k: np.array(out[k]) for k in ("slope", "err_read", "err_poisson", "pdq", "endslice")}

            assert drv.run(lambda e, s: (*l1[(e, s)], wcs[(e, s)]), sink, outs) == len(drv.items)
    assert sorted(got) == items
    for (e, s), out in got.items():
        area = coordutils.pixelarea_device(wcs[(e, s)], N=n, inv_omega=1.0 / pars.Omega_ideal, dtype=np.float32)
        with gci.CalDir(cals[s]) as cd:
            ref = gci.calibrate_arrays(cd, *l1[(e, s)], rp, synth.FRAME_TIME, area, cfg, do_refpix=True, want_endslice=True)
        for k in ("pdq", "endslice"):
            assert np.array_equal(out[k], ref[k]), (e, s, k)
        for k in ("slope", "err_read", "err_poisson"):
            assert np.array_equal(out[k], ref[k], equal_nan=True), (e, s, k)
