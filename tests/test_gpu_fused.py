"""GPU: the fused L1->L2 path (rip_l1_to_l2_host/_dev through ``calibrate_arrays``) against the oracle's restatement
of ``calibrateimage`` (reference L1_to_L2/gen_cal_image.py:503-629,697-709) on the same seeded inputs.

DQ (pdq, rdq), endslice: bit-exact.  slope / err_read / err_poisson / linearised+IPC-corrected cube: RTOL=1e-5
(tests/parity.py).  At full size (4096^2) the oracle takes minutes, so parity there is checked through the K0
statistics (cheap to restate), a row band of the frame, and tiling invariance.
"""

import numpy as np
import pytest
from conftest import SMALL_CASES, build_small_case
from parity import assert_bits_equal, assert_float_close, band_check, compare_l2

from oracle import rip_oracle as orc

pytestmark = pytest.mark.gpu

CFG7 = {"RAMP_OPT_PARS": {"slope": 0.4, "gain": 1.8, "sigma_read": 7.0}, "SLICEOUT": True}


def _run(cal, data, amp33, rp, area, cfg, do_refpix, **kw):
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    with gci.CalDir(cal) as cd:
        return gci.calibrate_arrays(cd, data, amp33, rp, 3.04, area, cfg, do_refpix=do_refpix, want_rdq=True,
                                    want_lin_cube=True, want_endslice=True, **kw)  # fmt: skip


@pytest.mark.parametrize("tag", list(SMALL_CASES))
def test_small_cases(tag):
    from romanimpreprocess_b200 import synth

    cal, data_u16, amp33_u16, meta, rp = build_small_case(tag)
    n = data_u16.shape[1]
    area = synth.make_area_factor(n)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data_u16, amp33_u16, c, rp, 3.04, area, CFG7, do_refpix=False, return_intermediates=True)
    # threads = 0: automatic kernel choice (the v2 throughput kernel where eligible: all-f32 planes, G in {8,16},
    # P in {4,11}); threads > 0: the generic v1 tile kernel with that tile width
    for threads, band in ((0, 0), (0, 5), (32, 16), (64, 7)):
        out = _run(cal, data_u16, amp33_u16, rp, area, CFG7, False, threads=threads, band_rows=band)
        stats = compare_l2(out, ref)
        assert np.array_equal(out["meta"]["K"], ref["K"])
        print(tag, threads, band, "values not bit-identical:", stats)


MEDIUM = [
    (256, "README_PATTERN", 10, np.float32, np.float32, 21, {}, 1.0, np.float64),
    (256, "LONG16_PATTERN", 10, np.float32, np.float32, 22, {"EXCLUDE_FIRST": False, "SATURATION_BACKUP": 2}, 4.0, np.float32),
    (384, "TEST_READ_PATTERN", 3, np.float64, np.float64, 23,
     {"JUMP_DETECT_PARS": {"SthreshA": 4.0, "SthreshB": 3.5, "IthreshA": 0.6, "IthreshB": 600.0}}, 2.0, np.float64),
    (256, "README_PATTERN", 10, np.float32, np.float64, 24, {"SATURATION_BACKUP": 0}, 8.0, None),
    (512, "README_PATTERN", 15, np.float64, np.float32, 25, {}, 3.0, np.float64),
    (384, "README_PATTERN", 3, np.float32, np.float32, 26,
     {"SATURATION_BACKUP": 0, "JUMP_DETECT_PARS": {"SthreshA": 4.0, "SthreshB": 3.5}}, 8.0, np.float32),
    (128, "LONG16_PATTERN", 3, np.float32, np.float32, 27, {}, 6.0, np.float64),
    (256, "README_PATTERN", 6, np.float32, np.float32, 28, {}, 4.0, np.float32),   # P = 7: the P = 11 kernel on zero-padded records
    (256, "README_PATTERN", 7, np.float32, np.float64, 29, {}, 4.0, np.float64),   # P = 8 with float64 ipc4d
]  # fmt: skip


@pytest.mark.parametrize("case", MEDIUM, ids=[f"n{c[0]}_{c[1]}_s{c[5]}" for c in MEDIUM])
def test_medium_cases_with_refpix(case):
    from romanimpreprocess_b200 import synth

    n, rpname, po, gdt, kdt, seed, cfg, bright, adt = case
    rp = getattr(synth, rpname)
    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=po, gain_dtype=gdt, ipc_dtype=kdt,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=seed + 1, n_sources=25, cr_frac=0.01, bright=bright)
    area = None if adt is None else synth.make_area_factor(n, adt)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data_u16, amp33_u16, c, rp, 3.04, 1.0 if area is None else area, cfg, do_refpix=True,
                       return_intermediates=True)  # fmt: skip
    for threads in (0, 128):  # v2 where eligible, and v1
        out = _run(cal, data_u16, amp33_u16, rp, area, cfg, True, threads=threads)
        stats = compare_l2(out, ref)
        print(case[:2], threads, "values not bit-identical:", stats)
    assert np.count_nonzero(ref["pdq"] & orc.SATURATED) > 50
    assert np.count_nonzero(ref["pdq"] & orc.JUMP_DET) > 50
    assert len(np.unique(ref["endslice"])) >= 4


def test_static_products_against_oracle():
    """K2: IPC-corrected dark slope (gen_cal_image.py:217-221), get_flat product + flags (flatutils.py:44-76)."""
    from hostcheck import harness
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    cal, *_ = build_small_case("small_sat_g64k64")
    c = {k: v["roman"] for k, v in cal.items()}
    thr, aux, sdq, dslope, flat = harness.static_products(c)
    with gci.CalDir(cal) as cd:
        sp = cd.static_products()
        assert sp["refout_slope"] == float(orc.optimal_refout_slope(c["read"]))
    assert_float_close(sp["dark_slope_ipc"], dslope, "dark_slope_ipc")
    assert_float_close(sp["flat"], flat, "flat")
    assert_bits_equal(sp["static_dq"], sdq, "static_dq")


@pytest.fixture(scope="module")
def full_case():
    from romanimpreprocess_b200 import synth

    rp = synth.README_PATTERN
    cal = synth.make_caldir(n=4096, seed=1000, read_pattern=rp, p_order=10, gain_dtype=np.float32,
                            ipc_dtype=np.float32, sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=200, n_sources=25, cr_frac=1e-3, bright=3.0)
    area = synth.make_area_factor(4096, np.float64)
    return cal, data_u16, amp33_u16, rp, area


def test_full_size_refpix_stats(full_case):
    """K0 at 4096^2 x 8: per-row / global / per-channel medians and the f64 corrections vs a NumPy restatement of
    gen_cal_image.py:531-555 + reference_subtraction.py (medians only -- seconds on the CPU)."""
    import ctypes as C

    from hostcheck import harness
    from romanimpreprocess_b200 import _lib
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    cal, data_u16, amp33_u16, rp, area = full_case
    c = {k: v["roman"] for k, v in cal.items()}
    G, n, _ = data_u16.shape
    rc_ref, cm_ref, cc_ref = harness.refpix_stats(data_u16, amp33_u16, c)
    rc, cm, cc, gm = np.empty((G, n)), np.empty((G, 32)), np.empty((G, 32)), np.empty(G, np.float32)
    with gci.CalDir(cal) as cd:
        _lib.check(_lib.lib().rip_refpix_stats_host(cd.handle, _lib.ptr(data_u16), _lib.ptr(amp33_u16), G, _lib.ptr(rc),
                                                    _lib.ptr(cm), _lib.ptr(cc), _lib.ptr(gm)))  # fmt: skip
    assert np.array_equal(rc, rc_ref), np.abs(rc - rc_ref).max()
    assert np.array_equal(cm, cm_ref) and np.array_equal(cc, cc_ref)
    for j in range(G):
        ro = amp33_u16[j].astype(np.float32) - c["read"]["amp33"]["med"]
        assert gm[j] == np.median(ro)


def test_full_size_band_and_tiling_invariance(full_case):
    """4096^2 x 8, P=11 (BASELINE configs[1]/[0] shape): (1) a 64-row band + halo is cut out and run through the
    oracle with the full-frame K0 statistics -> must match the band of the full-frame CUDA result;
    (2) results are invariant under the tile geometry (threads, band_rows)."""
    from hostcheck import harness

    cal, data_u16, amp33_u16, rp, area = full_case
    cfg = dict(CFG7)
    out = _run(cal, data_u16, amp33_u16, rp, area, cfg, True)
    out2 = _run(cal, data_u16, amp33_u16, rp, area, cfg, True, threads=256, band_rows=512)  # v1 kernel, other tiling
    out3 = _run(cal, data_u16, amp33_u16, rp, area, cfg, True, band_rows=96)  # v2 kernel, other band height
    for k in ("slope", "err_read", "err_poisson", "lin_cube"):
        assert np.array_equal(out[k], out3[k], equal_nan=True), k
    for k in ("pdq", "rdq", "endslice"):
        assert np.array_equal(out[k], out3[k]), k
    for k in ("slope", "err_read", "err_poisson", "lin_cube"):
        assert np.array_equal(out[k], out2[k], equal_nan=True), k
    for k in ("pdq", "rdq", "endslice"):
        assert np.array_equal(out[k], out2[k]), k
    # border semantics (gen_cal_image.py:470-475): science border zero, reference-pixel flag kept
    assert np.all(out["slope"][:4] == 0) and np.all(out["slope"][:, -4:] == 0)
    static_ref = (cal["mask"]["roman"]["dq"] | cal["linearitylegendre"]["roman"]["dq"]) & orc.REFERENCE_PIXEL
    assert np.all(out["pdq"][:4] & orc.REFERENCE_PIXEL) and np.array_equal(out["pdq"] & orc.REFERENCE_PIXEL, static_ref)
    assert np.count_nonzero(out["pdq"] & orc.SATURATED) > 1000 and np.count_nonzero(out["pdq"] & orc.JUMP_DET) > 5000

    # band check against the oracle
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    with gci.CalDir(cal) as cd:
        sp = cd.static_products()
    band_check(out, cal, data_u16, amp33_u16, rp, area, sp["flat"], sp["dark_slope_ipc"])


def _compare_full(out, ref, tag):
    stats = {}
    for k in ("slope", "err_read", "err_poisson"):
        stats[k] = assert_float_close(out[k], ref[k], f"{tag} {k}")
    for k in ("pdq", "rdq", "endslice"):
        assert_bits_equal(out[k], ref[k], f"{tag} {k}")
    return stats


def test_full_frame_every_pixel(full_case):
    """BASELINE metric configuration (4096^2 x 8 resultants, P = 11, reference-pixel correction on): EVERY pixel of the
    CUDA result against ``oracle.l1_to_l2`` of the whole frame (row bands in a process pool with the global K0
    statistics: tests/fullframe.py).  DQ / rdq / endslice bit-exact, floats rtol 1e-5 (the reference's own "same answer"
    bar is <= 2 differing pixels, tests/romanimpreprocess/test_workflow.py:870-874; ours must be 0)."""
    import fullframe

    cal, data_u16, amp33_u16, rp, area = full_case
    cfg = dict(CFG7)
    ref = fullframe.oracle_full_frame(cal, data_u16, amp33_u16, rp, area, cfg)
    out = _run(cal, data_u16, amp33_u16, rp, area, cfg, True)
    stats = _compare_full(out, ref, "4096^2 x 8")
    print("4096^2 x 8, values not bit-identical:", stats)
    assert np.count_nonzero(ref["pdq"] & orc.SATURATED) > 1000 and np.count_nonzero(ref["pdq"] & orc.JUMP_DET) > 5000
    assert np.count_nonzero(ref["pdq"] & orc.GW_AFFECTED_DATA) > np.count_nonzero(cal["mask"]["roman"]["dq"] & orc.GW_AFFECTED_DATA)


def test_full_frame_kernel_variants_agree_and_repeat(full_case):
    """The default kernel (v6: five CTAs per SM, two barriers per step, rings reused after 3 / 4 rows) run three times on the
    whole 4096^2 frame gives bitwise identical arrays each time (a shared-memory race would not) and the same arrays as the
    v2 kernel (params.threads = -1) and, for the float64-ipc4d path, as its v2 form."""
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    cal, data_u16, amp33_u16, rp, area = full_case
    keys = ("slope", "err_read", "err_poisson", "pdq", "rdq", "endslice")
    with gci.CalDir(cal) as cd:
        runs = [gci.calibrate_arrays(cd, data_u16, amp33_u16, rp, 3.04, area, dict(CFG7), do_refpix=True, want_rdq=True,
                                     want_endslice=True, threads=t) for t in (0, 0, 0, -1)]  # fmt: skip
    for r in runs[1:]:
        for k in keys:
            assert np.array_equal(runs[0][k], r[k], equal_nan=runs[0][k].dtype.kind == "f"), k


def test_full_frame_dummy_caldir_dtypes():
    """The dtypes of the production DUMMY CALDIR (reference runs/summer2025run/make_gain_file.py:88,138,194): gain
    float32, ipc4d FLOAT64 (the reference then runs its IPC stage in float64), at 4096^2 x 8, every pixel."""
    import fullframe
    from romanimpreprocess_b200 import synth

    rp = synth.README_PATTERN
    cal = synth.make_caldir(n=4096, seed=1010, read_pattern=rp, p_order=10, gain_dtype=np.float32,
                            ipc_dtype=np.float64, sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=210, n_sources=25, cr_frac=1e-3, bright=3.0)
    area = synth.make_area_factor(4096, np.float64)
    cfg = dict(CFG7)
    ref = fullframe.oracle_full_frame(cal, data_u16, amp33_u16, rp, area, cfg)
    out = _run(cal, data_u16, amp33_u16, rp, area, cfg, True)
    print("4096^2 x 8 (ipc4d f64), values not bit-identical:", _compare_full(out, ref, "ipc4d f64"))


def test_long_table_1024_with_refpix():
    """BASELINE configs[2]: the 16-resultant table, reference-pixel correction on, at 1024^2 (whole-frame oracle)."""
    import fullframe
    from romanimpreprocess_b200 import synth

    rp = synth.LONG16_PATTERN
    cal = synth.make_caldir(n=1024, seed=1020, read_pattern=rp, p_order=10, gain_dtype=np.float32,
                            ipc_dtype=np.float32, sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=220, n_sources=25, cr_frac=2e-3, bright=4.0)
    area = synth.make_area_factor(1024, np.float32)
    cfg = {"SLICEOUT": True, "SATURATION_BACKUP": 2}
    ref = fullframe.oracle_full_frame(cal, data_u16, amp33_u16, rp, area, cfg, band=128)
    out = _run(cal, data_u16, amp33_u16, rp, area, cfg, True)
    print("1024^2 x 16, values not bit-identical:", _compare_full(out, ref, "long table"))
    assert len(np.unique(ref["endslice"])) >= 6


def test_pipeline_matches_synchronous_path():
    """rip_pipeline_* (three streams, exposures in flight) gives the same arrays as rip_l1_to_l2_host, in order, also
    when more exposures are queued than there are slots."""
    from romanimpreprocess_b200 import _lib, synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    n, rp = 256, synth.README_PATTERN
    cal = synth.make_caldir(n=n, seed=51, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    exposures = [synth.make_l1(cal, rp, seed=60 + k, n_sources=25, cr_frac=0.01, bright=3.0)[:2] for k in range(5)]
    area = synth.make_area_factor(n, np.float32)
    cfg = {"SLICEOUT": True}
    with gci.CalDir(cal) as cd:
        ref = [gci.calibrate_arrays(cd, d, a, rp, 3.04, area, cfg, want_rdq=True) for d, a in exposures]
        with gci.Pipeline(cd, rp, 3.04, cfg, depth=2, want_rdq=True) as pipe:
            pinned = []
            for d, a in exposures:
                pd, pa = _lib.pinned_empty(d.shape, np.uint16), _lib.pinned_empty(a.shape, np.uint16)
                pd[...] = d
                pa[...] = a
                pinned.append((pd, pa))
            tickets = [pipe.submit(pd, pa, area) for pd, pa in pinned]
            outs = [pipe.result(t) for t in tickets]
            with pytest.raises(_lib.RipError, match="unknown ticket"):
                pipe.result(99)
    for o, r in zip(outs, ref):
        for k in ("slope", "err_read", "err_poisson"):
            assert np.array_equal(o[k], r[k], equal_nan=True), k
        for k in ("pdq", "rdq", "endslice"):
            assert np.array_equal(o[k], r[k]), k
    assert not np.array_equal(outs[0]["slope"], outs[1]["slope"])


def test_refpix_lookahead_matches_in_stream_statistics():
    """rip_caldir_prefetch_refpix: the reference-pixel statistics of the next device-resident exposure, computed on the
    handle's side stream beside the current fused kernel, give bitwise the arrays of the in-stream path -- over a rotation
    of exposures (both workspace sets reused), with a look-ahead that is never consumed, and with one for another cube."""
    import torch

    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    n, rp = 512, synth.README_PATTERN
    cal = synth.make_caldir(n=n, seed=52, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    exposures = [synth.make_l1(cal, rp, seed=80 + k, n_sources=25, cr_frac=0.01, bright=3.0)[:2] for k in range(3)]
    area = synth.make_area_factor(n, np.float32)
    cfg = {"SLICEOUT": True}
    dev = torch.device("cuda", 0)
    with gci.CalDir(cal) as cd:
        ref = [gci.calibrate_arrays(cd, d, a, rp, 3.04, area, cfg) for d, a in exposures]
        dplan = gci.DevicePlan(cd, rp, 3.04, cfg, do_refpix=True, area_dtype=np.float32)
        d_raw = [torch.from_numpy(d.view(np.int16)).to(dev) for d, _ in exposures]
        d_amp = [torch.from_numpy(a.view(np.int16)).to(dev) for _, a in exposures]
        d_area = torch.from_numpy(area).to(dev)
        outs = [{"slope": torch.empty((n, n), dtype=torch.float32, device=dev),
                 "err_read": torch.empty((n, n), dtype=torch.float32, device=dev),
                 "err_poisson": torch.empty((n, n), dtype=torch.float32, device=dev),
                 "pdq": torch.empty((n, n), dtype=torch.int32, device=dev)} for _ in range(7)]  # fmt: skip
        torch.cuda.synchronize()
        stream = torch.cuda.current_stream().cuda_stream
        order = [0, 1, 2, 0, 1, 2, 0]
        for i, k in enumerate(order):
            o = outs[i]
            gci.calibrate_device(cd, dplan, d_raw[k].data_ptr(), d_amp[k].data_ptr(), d_area.data_ptr(),
                                 o["slope"].data_ptr(), o["err_read"].data_ptr(), o["err_poisson"].data_ptr(),
                                 o["pdq"].data_ptr(), stream=stream)  # fmt: skip
            if i == 3:    # a look-ahead for a cube that is NOT the next one: must be ignored, not used
                gci.prefetch_refpix_device(cd, d_raw[0].data_ptr(), d_amp[0].data_ptr(), len(rp))
            elif i == 4:  # two look-aheads in a row, the second one is the right one
                gci.prefetch_refpix_device(cd, d_raw[1].data_ptr(), d_amp[1].data_ptr(), len(rp))
                gci.prefetch_refpix_device(cd, d_raw[2].data_ptr(), d_amp[2].data_ptr(), len(rp))
            elif i + 1 < len(order):
                k1 = order[i + 1]
                gci.prefetch_refpix_device(cd, d_raw[k1].data_ptr(), d_amp[k1].data_ptr(), len(rp))
        torch.cuda.synchronize()
        for i, k in enumerate(order):
            for name in ("slope", "err_read", "err_poisson"):
                assert np.array_equal(outs[i][name].cpu().numpy(), ref[k][name], equal_nan=True), (i, name)
            assert np.array_equal(outs[i]["pdq"].cpu().numpy().view(np.uint32), ref[k]["pdq"]), i
    assert not np.array_equal(ref[0]["slope"], ref[1]["slope"], equal_nan=True)


def test_sky_step_after_the_hot_path():
    """calibrate_arrays(sky_step=True): slope_withsky + SKYORDER medfit subtraction (reference gen_cal_image.py:639-651)
    == the oracle's medfit (pinned to the reference's) applied to the oracle's slope (model within one float32 ulp: the
    6 x 6 solve is the library's LU, not LAPACK's)."""
    import warnings

    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    n, rp = 256, synth.README_PATTERN
    cal = synth.make_caldir(n=n, seed=61, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data, amp33, _ = synth.make_l1(cal, rp, seed=62, n_sources=9, cr_frac=0.01)
    area = synth.make_area_factor(n, np.float32)
    cfg = {"SKYORDER": 2}
    with gci.CalDir(cal) as cd:
        out = gci.calibrate_arrays(cd, data, amp33, rp, synth.FRAME_TIME, area, cfg, do_refpix=True, sky_step=True)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data, amp33, c, rp, synth.FRAME_TIME, area, cfg, do_refpix=True)
    assert np.array_equal(out["slope_withsky"], ref["slope"], equal_nan=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        coef, model, _ = orc.medfit(np.ascontiguousarray(ref["slope"][4:-4, 4:-4]), order=2)
    expect = ref["slope"].copy()
    expect[4:-4, 4:-4] -= model
    # (coefficients: the library's LU vs NumPy's LAPACK solve, a few float64 ulp; model float32: at most 1 ulp)
    np.testing.assert_allclose(out["skycoefs"], coef, rtol=0, atol=1e-13 * np.abs(coef).max())
    assert out["skyorder"] == 2
    np.testing.assert_allclose(out["slope"], expect, rtol=0, atol=2e-7 * np.abs(model).max(), equal_nan=True)
    assert np.array_equal(out["slope"][:4], expect[:4])
