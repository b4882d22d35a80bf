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


@pytest.mark.gpu
def test_run_config_then_calibrateimage_round_trip(tmp_path):
    """The production loop on files (reference runs/summer2025run/OpenUniverse_to_L1L2.py:155-165): ``sim_to_isim.run_config``
    turns a truth FITS image into an L1 ASDF file + FITSWCS header text, ``calibrateimage`` turns that into the L2 file,
    ``PixelMask1.convert_file`` writes the mask: the calibrated slope recovers the (flipped) truth scene."""
    from fixture_files import sim_header

    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i
    from romanimpreprocess_b200.io import fits_lite
    from romanimpreprocess_b200.utils import maskhandling

    n = 256
    na = n - 8
    config2, cal, _, _, rp = write_exposure(str(tmp_path), n=n, ipc_dtype=np.float32)
    yy, xx = np.mgrid[0:na, 0:na]
    exptime = 139.8
    rate = 40.0 + 0.1 * xx + 0.05 * yy  # e/s per ideal pixel, asymmetric so that a wrong flip shows
    truth = str(tmp_path / "Roman_Test_truth_F184_7_1.fits")  # (obsid 7, SCA 1: vertical flip)
    hdr = dict(sim_header(na), EXPTIME=exptime, FILTER="F184", RA_TARG=37.0, DEC_TARG=-20.0, PA_OBSY=35.0)
    hdr["DATE-OBS"] = "2026-03-01 00:00:00.000"
    fits_lite.write_hdus(truth, [((rate * exptime).astype(np.float32), hdr)])
    reads = [v for g in rp for v in (g[0], g[-1] + 1)]
    config1 = {"IN": truth, "OUT": str(tmp_path / "sim_L1_F184_7_1.asdf"), "READS": reads, "CALDIR": config2["CALDIR"],
               "CNORM": 1.0, "SEED": 500, "FITSOUT": True}  # fmt: skip
    s2i.run_config(config1)
    d, a33, rpat, ft, meta, _ = gci.read_l1(config1["OUT"])
    assert d.shape == (len(rp), n, n) and d.dtype == np.uint16 and rpat == [list(g) for g in rp] and ft == 3.04
    assert meta["instrument"]["detector"] == "WFI01" and meta["instrument"]["optical_element"] == "F184"
    assert a33 is not None and a33.shape == (len(rp), n, 128)
    ((fimg, _),) = fits_lite.read_hdus(config1["OUT"][:-5] + "_asdf_to.fits")
    assert fimg.shape == (len(rp), n, n + 128) and np.array_equal(fimg[:, :, :n], d)
    w = coordutils.FitsWCS(open(config1["OUT"][:-5] + "_asdf_wcshead.txt").read())
    assert w.crpix[1] == pytest.approx(na + 1 - (na + 1) / 2.0 - 1) and w.cd[1, 1] == pytest.approx(-3.0555555555555554e-05)
    config2 = dict(config2, IN=config1["OUT"], OUT=str(tmp_path / "sim_L2_F184_7_1.asdf"),
                   FITSWCS=config1["OUT"][:-5] + "_asdf_wcshead.txt")  # fmt: skip
    gci.calibrateimage(config2, verbose=False)
    maskhandling.PixelMask1.convert_file(config2["OUT"], config2["OUT"][:-5] + "_mask.fits")
    gci.clear_caldir_cache()
    with open_tree(config2["OUT"]) as f:
        slope, dq = np.asarray(f["roman"]["data_withsky"]), np.asarray(f["roman"]["dq"])
    expected = (rate[::-1, :] / pars.g_ideal).astype(np.float32)  # SCA 1 is flipped vertically (sim_to_isim.py:489-493)
    good = dq == 0
    ratio = slope[good] / expected[good]
    assert good.mean() > 0.7 and abs(np.median(ratio) - 1.0) < 0.03, (good.mean(), np.median(ratio))
    wrong = slope[good] / (rate / pars.g_ideal)[good]
    assert np.std(wrong) > 2 * np.std(ratio)  # (the unflipped scene does not fit)
    assert synth.FRAME_TIME == ft
