"""File-level drop-in: ``gen_cal_image.calibrateimage(config)`` (reference L1_to_L2/gen_cal_image.py:480-739) on CALDIR /
L1 / FITSWCS files written by tests/fixture_files.py.  CPU: the host side (file reading, refusal of unimplemented
switches).  GPU: the L2 file against the oracle run on the same arrays."""

import numpy as np
import pytest
import yaml
from fixture_files import write_exposure

from oracle import rip_oracle as orc
from romanimpreprocess_b200 import pars
from romanimpreprocess_b200.caltree import open_tree
from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
from romanimpreprocess_b200.utils import coordutils


def test_read_l1_and_caldir_files(tmp_path):
    config, cal, data, amp33, rp = write_exposure(str(tmp_path), n=128)
    d, a, rpat, ft, meta, border = gci.read_l1(config["IN"])
    assert np.array_equal(d, data) and np.array_equal(a, amp33) and rpat == [list(g) for g in rp] and ft == 3.04
    assert meta["instrument"]["detector"] == "WFI01"
    assert np.array_equal(border["border_ref_pix_left"], data[:, :, :4].astype(np.float32))
    with open_tree(config["CALDIR"]["ipc4d"]) as f:
        k = np.asarray(f["roman"]["data"])
        assert k.dtype == np.float64 and np.array_equal(k, cal["ipc4d"]["roman"]["data"])  # (the DUMMY builder's dtype)
    with open_tree(config["CALDIR"]["read"]) as f:
        assert float(f["roman"]["anc"]["C_PINK"]) == cal["read"]["roman"]["anc"]["C_PINK"]
        assert np.array_equal(np.asarray(f["roman"]["amp33"]["med"]), cal["read"]["roman"]["amp33"]["med"])
    # the configuration survives a YAML round trip (python -m ... cfg.yaml, gen_cal_image.py:742-746)
    assert yaml.safe_load(yaml.safe_dump(config)) == config


def test_extract_ref_file_is_reconstituted(tmp_path):
    config, cal, data, amp33, rp = write_exposure(str(tmp_path), n=128, extract_ref=True)
    d, *_ = gci.read_l1(config["IN"])
    ok = data.astype(np.int32) - data[0].astype(np.int32)[None] + 5000
    sel = (ok >= 0) & (ok <= 65535)  # (values the encoding could represent)
    assert np.array_equal(d[sel], data[sel]) and sel.mean() > 0.99


def test_unimplemented_switches_raise(tmp_path):
    config, *_ = write_exposure(str(tmp_path), n=128)
    for key in ("correct_wfi18_transient", "romancal_ramp_fit"):
        with pytest.raises(NotImplementedError, match=key):
            gci.calibrateimage(dict(config, **{key: True}))
    bad = dict(config, CALDIR=dict(config["CALDIR"], dark_decay="x.asdf"))
    with pytest.raises(NotImplementedError, match="dark_decay"):
        gci.calibrateimage(bad)
    nowcs = {k: v for k, v in config.items() if k != "FITSWCS"}
    with pytest.raises(ValueError, match="Unrecognized WCS"):
        gci.calibrateimage(nowcs)


@pytest.mark.gpu
@pytest.mark.parametrize("n,ipc_dtype", [(256, np.float64), (256, np.float32)])
def test_calibrateimage_writes_the_l2_file(tmp_path, n, ipc_dtype):
    from romanimpreprocess_b200.utils import maskhandling

    config, cal, data, amp33, rp = write_exposure(str(tmp_path), n=n, ipc_dtype=ipc_dtype)
    gci.calibrateimage(config, verbose=False)
    gci.calibrateimage(dict(config, FITSOUT=True), verbose=False)  # second exposure on the cached CALDIR / pipeline
    from romanimpreprocess_b200.io import fits_lite

    (fd, _), (fq, hq), (fm, _) = fits_lite.read_hdus(config["OUT"][:-5] + "_asdf_to.fits")
    maskhandling.PixelMask1.convert_file(config["OUT"], config["OUT"][:-5] + "_mask.fits")
    (md, _), (mm, hm) = fits_lite.read_hdus(config["OUT"][:-5] + "_mask.fits")
    gci.clear_caldir_cache()
    c = {k: v["roman"] for k, v in cal.items()}
    area = coordutils.pixelarea(coordutils.wcs_from_config(config), N=n) / pars.Omega_ideal
    ref = orc.l1_to_l2(data, amp33, c, rp, 3.04, area, config, do_refpix=True)
    with open_tree(config["OUT"]) as f:
        r, pi = f["roman"], f["processinfo"]
        act = np.s_[4:-4, 4:-4]
        assert np.array_equal(np.asarray(r["dq"]), ref["pdq"][act])
        assert np.array_equal(np.asarray(pi["endslice"]), ref["endslice"])
        # FITS side products (gen_cal_image.py:725-736, maskhandling.py:145-149)
        grown = maskhandling.PixelMask1.build(ref["pdq"][act])
        assert np.array_equal(fq, ref["pdq"][act]) and fq.dtype == np.uint32 and hq["BZERO"] == 2147483648
        assert np.array_equal(fd, np.asarray(r["data"])) and np.array_equal(fm, np.where(~grown, fd, -1000).astype(np.float32))
        assert np.array_equal(mm, grown.astype(np.int8)) and hm["EXTNAME"] == "MASK" and np.array_equal(md, np.where(grown, -1000.0, fd).astype(np.float32))
        np.testing.assert_allclose(np.asarray(r["data_withsky"]), ref["slope"][act], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.asarray(r["var_rnoise"]), ref["err_read"][act] ** 2, rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(np.asarray(r["var_poisson"]), ref["err_poisson"][act] ** 2, rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(np.asarray(r["err"]), np.hypot(ref["err_read"], ref["err_poisson"])[act], rtol=2e-5, atol=1e-7)
        coef, model, _ = orc.medfit(np.ascontiguousarray(ref["slope"][act]), order=2)
        np.testing.assert_allclose(np.asarray(pi["skycoefs"]), coef, rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(np.asarray(r["data"]), ref["slope"][act] - model, rtol=1e-5, atol=2e-6)
        assert pi["skyorder"] == 2 and abs(pi["medgain"] - np.median(c["gain"]["data"])) < 1e-6
        assert np.array_equal(np.asarray(pi["weights"]), ref["K"])
        assert np.array_equal(np.asarray(r["amp33"]), amp33)
        assert np.array_equal(np.asarray(r["border_ref_pix_top"]), data[:, -4:, :].astype(np.float32))
        assert np.array_equal(np.asarray(r["dq_border_ref_pix_left"]), ref["pdq"][:, :4])
        assert np.asarray(r["chisq"]).dtype == np.float16 and r["meta"]["instrument"]["detector"] == "WFI01"
        assert r["meta"]["exposure"]["read_pattern"] == [list(g) for g in rp]
        assert "Ramp fitting complete" in pi["log"] and pi["config"]["SKYORDER"] == 2
        # medsky: mode of the smoothed histogram of the masked, 4x4-binned slope (gen_cal_image.py:641)
        m = maskhandling.PixelMask1.build(ref["pdq"])
        binned = np.mean(np.where(~m, ref["slope"], np.nan)[: n // 4 * 4, : n // 4 * 4].reshape(n // 4, 4, n // 4, 4), axis=(1, 3))
        assert abs(pi["medsky"] - np.nanmedian(binned)) < 0.2 * np.nanstd(binned)
