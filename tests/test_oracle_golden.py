"""CPU: the oracle (oracle/rip_oracle.py) against the golden vectors produced by the unmodified reference.

Bit-exact (``array_equal``) everywhere: the oracle restates the reference with the same NumPy op order.
"""

import hashlib

import numpy as np
import pytest
from conftest import SMALL_CASES, build_small_case, load_golden

from oracle import rip_oracle as orc


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def test_lin_kat(kats):
    """reference tests/romanimpreprocess/test_linutils.py:7-49 (two-sided here)."""
    z = kats["lin_p3_z"].reshape((1, 31))
    coefs = np.zeros((4, 1, 31))
    coefs[3] = 1.0
    phi, ex = orc.lin_eval(z, coefs)
    assert np.all(np.abs(phi - kats["lin_p3_phi"].reshape(phi.shape)) < 1e-6)
    assert np.array_equal(ex[0], np.abs(kats["lin_p3_z"]) > 1)


def test_weights_kat(kats):
    """reference src/romanimpreprocess/L1_to_L2/denoise_construct.py:219-230."""
    from romanimpreprocess_b200 import synth

    meta = synth.meta_from_pattern(synth.README_PATTERN)
    K = orc.construct_weights(0.4 / 1.8 / 7.0**2, meta, exclude_first=True)
    assert np.array_equal(K, kats["weights_readme_ref"])
    assert np.array_equal(K, kats["weights_readme"].astype(np.float32))
    K2 = orc.construct_weights(0.4 / 1.8 / 6.5**2, meta, exclude_first=False)
    assert np.array_equal(K2, kats["weights_readme_ref_noexcl"])


@pytest.mark.parametrize("tag", list(SMALL_CASES))
def test_small_case_bit_exact(tag):
    g = load_golden(tag)
    cal, data_u16, amp33_u16, meta, rp = build_small_case(tag)
    c = {k: v["roman"] for k, v in cal.items()}
    assert str(g["input_digest"]) == _digest(
        data_u16, c["linearitylegendre"]["data"], c["gain"]["data"], c["ipc4d"]["data"], c["read"]["data"]
    ), "synthetic input generator drifted from the one that made the golden file"
    n = data_u16.shape[1]
    G = len(rp)
    S = data_u16.astype(np.float32)
    rdq0 = g["rdq0"]
    lin = c["linearitylegendre"]
    phi, dq = orc.multilin(S, lin, do_not_flag_first=True, attempt_corr=~rdq0 & orc.SATURATED)
    assert np.array_equal(phi, g["multilin_phi"]) and np.array_equal(dq, g["multilin_dq"])
    phi_b, dq_b = orc.multilin(S, lin, do_not_flag_first=False)
    assert np.array_equal(phi_b, g["multilin_phi_flagfirst"]) and np.array_equal(dq_b, g["multilin_dq_flagfirst"])
    p1, d1 = orc.linearity(S[2, 5:25, 7:30], lin, origin=(7, 5))
    assert np.array_equal(p1, g["linearity_phi"]) and np.array_equal(d1, g["linearity_dq"])

    K = c["ipc4d"]["data"]
    gain = c["gain"]["data"]
    g_act = gain[4:-4, 4:-4]
    img = phi[3, 4:-4, 4:-4].copy()
    assert np.array_equal(orc.ipc_fwd(img, K), g["ipc_fwd"])
    assert np.array_equal(orc.ipc_fwd(img, K, gain=g_act), g["ipc_fwd_gain"])
    assert np.array_equal(orc.ipc_rev(img, K), g["ipc_rev"])
    assert np.array_equal(orc.ipc_rev(img, K, gain=g_act), g["ipc_rev_gain"])
    assert np.array_equal(orc.ipc_rev(img, K, order=3), g["ipc_rev_order3"])
    cube = phi.copy()
    orc.correct_cube(cube, K, gain)
    assert np.array_equal(cube, g["correct_cube"])
    cube_e = phi.copy()
    orc.correct_cube(cube_e, K, None)
    assert np.array_equal(cube_e, g["correct_cube_nogain"])

    m = dict(meta)
    m["K"] = orc.construct_weights(0.4 / 1.8 / 7.0**2, m, exclude_first=True)
    assert np.array_equal(m["K"], g["K"])
    m["jump_detect_pars"] = {"SthreshA": 10.0, "SthreshB": 4.5, "IthreshA": 0.6, "IthreshB": 600.0}
    pdq = c["mask"]["dq"].copy() | dq
    read = c["read"]["data"]
    rdq = np.zeros_like(rdq0)
    s, er, ep, smap = orc.jump_detect(cube, rdq, pdq, m, gain, read, True, None)
    for a, k in ((s, "jd_slope"), (er, "jd_err_read"), (ep, "jd_err_poisson"), (smap, "jd_smap"), (rdq, "jd_rdq")):
        assert np.array_equal(a, g[k], equal_nan=True), k
    rdq = np.zeros_like(rdq0)
    s, er, ep, smap = orc.jump_detect(cube, rdq, pdq, m, gain, read, True, G - 1)
    for a, k in ((s, "jdt_slope"), (er, "jdt_err_read"), (ep, "jdt_err_poisson"), (smap, "jdt_smap"), (rdq, "jdt_rdq")):
        assert np.array_equal(a, g[k], equal_nan=True), k
    rdq = rdq0.copy()
    pdq_rf = pdq.copy()
    s, er, ep = orc.ramp_fit(cube, rdq, pdq_rf, m, gain, read, True)
    for a, k in ((s, "rf_slope"), (er, "rf_err_read"), (ep, "rf_err_poisson"), (rdq, "rf_rdq"), (pdq_rf, "rf_pdq")):
        assert np.array_equal(a, g[k], equal_nan=True), k
    assert np.count_nonzero(g["rf_rdq"] & 4) > 0, "golden case has no jump flags"
    m2 = dict(meta)
    m2["K"] = orc.construct_weights(0.4 / 1.8 / 6.5**2, m2, exclude_first=False)
    assert np.array_equal(m2["K"], g["rf2_K"])
    rdq = rdq0.copy()
    rdq[0] = 0
    pdq_rf2 = pdq.copy()
    s, er, ep = orc.ramp_fit(cube, rdq, pdq_rf2, m2, gain, read, False)
    for a, k in ((s, "rf2_slope"), (er, "rf2_err_read"), (ep, "rf2_err_poisson"), (rdq, "rf2_rdq"), (pdq_rf2, "rf2_pdq")):
        assert np.array_equal(a, g[k], equal_nan=True), k

    pdq_f = c["mask"]["dq"].copy()
    fl = orc.get_flat(c["flat"]["data"], gain, K, 4, pdq_f)
    assert np.array_equal(fl, g["flat"]) and np.array_equal(pdq_f, g["flat_pdq"])
    assert np.array_equal(orc.get_flat(c["flat"]["data"], None, None, 4, None, ipc_deconvolve=False), g["flat_noipc"])

    Sinv, ex = orc.invlinearity(phi[2, 4:-4, 4:-4].astype(np.float64), lin, origin=(4, 4))
    assert np.array_equal(Sinv, g["invlin_S"]) and np.array_equal(ex, g["invlin_ex"])
    Sinv32, _ = orc.invlinearity(phi[2, 4:-4, 4:-4], lin, origin=(4, 4))
    assert np.array_equal(Sinv32, g["invlin_S_f32"])
    out = orc.il_apply(g["il_counts"], lin, gain, K, start_e=g["il_start_e"], electrons=True)
    assert np.array_equal(out, g["il_apply"])
    out = orc.il_apply(g["il_counts"], lin, gain, K, start_e=g["il_start_e"], electrons=True, electrons_out=True)
    assert np.array_equal(out, g["il_apply_eout"])
    assert n - 8 == g["il_dq"].shape[-1]


def test_refsub_4096_bit_exact():
    """reference utils/reference_subtraction.py:16,77 on the hard-coded 4096x4224 geometry."""
    g = load_golden("refsub_4096")
    rng = np.random.RandomState(int(g["seed"]))
    im = (rng.normal(size=(4096, 4224)) * 6.0).astype(np.float32)
    im += (4.0 * np.sin(np.arange(4096) / 50.0)).astype(np.float32)[:, None]
    im[:, 4096:] *= 0.8
    im += (np.arange(4224) // 128).astype(np.float32)[None, :] * 0.37
    slope = np.float64(g["slope"])
    a = orc.ref_subtraction_row(im.copy(), use_ref_channel=True, slope=slope)
    assert np.array_equal(a[::97, ::89], g["row_sample"]) and _digest(a) == str(g["row_digest"])
    b = orc.ref_subtraction_channel(a.copy(), use_ref_channel=True)
    assert np.array_equal(b[::97, ::89], g["chan_sample"]) and _digest(b) == str(g["chan_digest"])
    c = orc.ref_subtraction_row(im.copy(), use_ref_channel=False)
    assert np.array_equal(c[::97, ::89], g["rowfit_sample"]) and _digest(c) == str(g["rowfit_digest"])
    d = orc.ref_subtraction_channel(im.copy(), use_ref_channel=False)
    assert _digest(d) == str(g["chan32_digest"])


def test_ref_row_property():
    """The reference's own property test (tests/romanimpreprocess/test_ref.py:7-21) on the oracle."""
    im = np.zeros((4096, 4224), dtype=np.float32)
    im[:, :] = np.cos(np.linspace(0, 2000, 4096))[:, None]
    im[:, -128:] *= 2.0
    for x in range(4224):
        im[:, x] += np.sin(0.1 * x) * np.sin(np.linspace(0, 2000, 4096)) ** 3
    im[:, :-128] += 1.0
    old = im.copy()
    orc.ref_subtraction_row(im, use_ref_channel=False)
    assert np.std(im) < 0.75 * np.std(old)
    assert 0.4 < np.std(im[:, :-128]) < 0.5
    assert 0.99 < np.mean(im[:, :-128]) < 1.01


@pytest.fixture(scope="module")
def gencal_fixture():
    """The reference's full-size test CALDIR (test_workflow.py:117-332, RandomState(1000))."""
    from romanimpreprocess_b200 import synth

    return synth.make_caldir(n=4096, seed=1000)


def test_il_example_kat(gencal_fixture, kats):
    """IL.apply golden vectors of the reference (tests/romanimpreprocess/test_workflow.py:382-422), tol 0.002."""
    cal = {k: v["roman"] for k, v in gencal_fixture.items()}
    y0, y1, x0, x1 = 252, 270, 132, 150  # crop of the active array (multiples of 3 keep the NE[::3,::3] phase)
    K = cal["ipc4d"]["data"][:, :, y0:y1, x0:x1]
    g = cal["gain"]["data"][4 + y0 : 4 + y1, 4 + x0 : 4 + x1]
    lin = cal["linearitylegendre"]
    for target, fill in ((kats["il_target1"], 0.0), (kats["il_target2"], 2.0e3)):
        NE = np.zeros((4088, 4088), dtype=np.float32)
        if fill:
            NE[::3, ::3] = fill
        ne = NE[y0:y1, x0:x1]
        conv = orc.ipc_fwd(ne + 0.0, K)
        S, _ = orc.invlinearity(conv / g, lin, origin=(4 + x0, 4 + y0))
        val = S[260 - y0 : 262 - y0, 140 - x0 : 143 - x0]
        assert np.all(np.abs(target - val) < 0.002), (val, target)
        assert np.all(np.abs(target - val) < 1e-7)  # in fact reproduced to print precision


def test_forward_backward_lin_ilin(gencal_fixture):
    """reference tests/romanimpreprocess/test_workflow.py:335-379."""
    lin = gencal_fixture["linearitylegendre"]["roman"]
    ymin, ymax, xmin, xmax = 260, 262, 140, 143
    S = lin["Sref"][ymin:ymax, xmin:xmax].copy()
    S += 5000.0 * np.linspace(0, 5, 6).reshape((2, 3))
    Slin, dq = orc.linearity(S, lin, origin=(xmin, ymin))
    Sfwd, ex = orc.invlinearity(Slin, lin, origin=(xmin, ymin))
    assert not np.any(ex)
    assert np.amax(np.abs(Sfwd - S)) < 0.002


def test_gencal_linearity_sanity(gencal_fixture):
    """Fixture self-checks of the reference (test_workflow.py:254-265) on a band of rows."""
    lin = gencal_fixture["linearitylegendre"]["roman"]
    Sref = lin["Sref"][4:260, 4:-4]
    s0, _ = orc.linearity(Sref, lin, origin=(4, 4))
    sp, _ = orc.linearity(Sref + 5, lin, origin=(4, 4))
    sm, _ = orc.linearity(Sref - 5, lin, origin=(4, 4))
    der = (sp - sm) / 10.0
    assert -1.5 < np.amin(s0) and np.amax(s0) < 1.5
    assert 0.99 < np.amin(der) and np.amax(der) < 1.01


def _sim_small_cal(n, G, seed):
    """Same CALDIR as tests/golden/make_golden_sim.py:small_cal."""
    from romanimpreprocess_b200 import synth

    pattern = [[0], [1, 2], [3, 4, 5, 6]][:G]
    cal = synth.make_caldir(n=n, read_pattern=pattern, p_order=3, seed=seed)
    cal = {k: v["roman"] for k, v in cal.items()}
    cw = n // 32
    a33 = cal["read"]["amp33"]
    a33["med"] = np.ascontiguousarray(a33["med"][:, :cw])
    a33["std"] = np.ascontiguousarray(a33["std"][:, :cw])
    return cal, pattern


def test_sim_refdata_golden():
    """a20 / a15: noise_1f_frame, fill_in_refdata_and_1f (3 modes) and the Image2D.simulate calibration planes
    against the unmodified reference functions (reference from_sim/sim_to_isim.py:265-402, 615-633) fed from the
    same normal stream (tests/golden/make_golden_sim.py)."""
    g = load_golden("sim_refdata_n256")
    n, G, seed = int(g["n"]), int(g["G"]), int(g["seed"])
    cal, pattern = _sim_small_cal(n, G, seed)
    tij = orc.read_pattern_to_tij(pattern)
    assert np.array_equal(orc.noise_1f_frame(orc.NormalStream(11), n, n // 32), g["frame_seed11"])
    for tag, banding, with33 in (("full", True, True), ("nobanding", False, True), ("no33", True, False)):
        im = g["im0"].copy()
        a33 = np.zeros((G, n, n // 32), np.uint16) if with33 else None
        orc.fill_in_refdata_and_1f(im, cal, orc.NormalStream(77), tij, fill_in_banding=banding, amp33=a33)
        assert np.array_equal(im, g[f"{tag}_im"]), tag
        if with33:
            assert np.array_equal(a33, g[f"{tag}_amp33"]), tag
    d, fl, _ = orc.sim_calprep(cal)
    assert np.array_equal(d, g["this_dark"]) and np.array_equal(fl, g["this_flat"])


def test_mask_and_moments_golden():
    """CombinedMask.build / PixelMask1 (reference utils/maskhandling.py:82-117,152-178) and the moment sums of
    validation_tests/many_realizations.py:74-83 against the unmodified reference (tests/golden/make_golden_sim.py)."""
    g = load_golden("mask_moments")
    assert np.array_equal(orc.mask_build(g["dq"]), g["mask_pixelmask1"])
    assert np.array_equal(orc.mask_build(g["dq"], {2: 25, 11: 5, 7: 9, 1: 1}), g["mask_custom"])
    mom = np.zeros_like(g["moments_sum"])
    for dat, dq in zip(g["data"], g["dqs"]):
        orc.moments_accumulate(mom, dat, dq)
    assert np.array_equal(mom, g["moments_sum"])
    orc.moments_finalize(mom)
    assert np.array_equal(mom, g["moments_final"])
    assert np.any(mom[1] == -1000.0)


def test_sky_medfit_golden():
    """sky.medfit (reference utils/sky.py:98-190) against the unmodified reference function."""
    import warnings

    from conftest import SKY_CASES, synth_sky_image

    g = load_golden("sky_medfit")
    for tag, (ny, nx, order, nreg) in SKY_CASES.items():
        img = synth_sky_image(ny, nx, 40 + ord(tag))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            coef, model, meds = orc.medfit(img, N=nreg, order=order)
        assert np.array_equal(coef, g[f"{tag}_coef"])
        assert np.array_equal(meds, g[f"{tag}_meds"], equal_nan=True)
        assert np.array_equal(model[::7, ::5], g[f"{tag}_model_sub"])
        assert model.astype(np.float64).sum() == float(g[f"{tag}_model_sum"])


def test_band_tiled_oracle_equals_whole_frame():
    """tests/fullframe.py (the whole-frame checker of the 4096^2 GPU parity tests): oracle.l1_to_l2 run band by band
    with the global reference-pixel statistics == oracle.l1_to_l2 of the whole frame, bit for bit (256^2, all flag
    branches, guide-window growth across band edges)."""
    import fullframe
    from romanimpreprocess_b200 import synth

    rp = synth.README_PATTERN
    n = 256
    cal = synth.make_caldir(n=n, seed=5, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data, amp33, _ = synth.make_l1(cal, rp, seed=6, n_sources=25, cr_frac=0.01, bright=3.0)
    area = synth.make_area_factor(n, np.float64)
    cfg = {"RAMP_OPT_PARS": {"slope": 0.4, "gain": 1.8, "sigma_read": 7.0}, "SLICEOUT": True}
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data, amp33, c, rp, 3.04, area, cfg, do_refpix=True)
    tiled = fullframe.oracle_full_frame(cal, data, amp33, rp, area, cfg, band=48, workers=2)
    for k in ("slope", "err_read", "err_poisson", "pdq", "rdq", "endslice"):
        assert np.array_equal(ref[k], tiled[k], equal_nan=True), k
    assert np.count_nonzero(ref["pdq"] & orc.GW_AFFECTED_DATA) > np.count_nonzero(c["mask"]["dq"] & orc.GW_AFFECTED_DATA)


def test_reference_function_chain_equals_the_oracle_port():
    """oracle/ref_chain.py (the reference's own unmodified functions from oracle/_ref, the CPU baseline of bench.py) and the
    oracle port give identical L2 arrays: pins the port's glue against the reference functions on whole chains."""
    from oracle import ref_chain
    from romanimpreprocess_b200 import synth

    if not ref_chain.available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    for n, rp, po, gdt, kdt, cfg in ((72, synth.README_PATTERN, 10, np.float32, np.float32, {"SLICEOUT": True}),
                                     (64, synth.TEST_READ_PATTERN, 3, np.float64, np.float64, {"EXCLUDE_FIRST": False}),
                                     (64, synth.LONG16_PATTERN, 10, np.float32, np.float64, {"SATURATION_BACKUP": 2})):  # fmt: skip
        cal = synth.make_caldir(n=n, seed=5, read_pattern=rp, p_order=po, gain_dtype=gdt, ipc_dtype=kdt, sprinkle_flags=True,
                                biascorr_amp=3.0)  # fmt: skip
        data, amp33, _ = synth.make_l1(cal, rp, seed=6, n_sources=9, cr_frac=0.01, bright=6.0)
        c = {k: v["roman"] for k, v in cal.items()}
        area = synth.make_area_factor(n, np.float64)
        a = orc.l1_to_l2(data, amp33, c, rp, synth.FRAME_TIME, area, cfg, do_refpix=False)
        b = ref_chain.l1_to_l2(data, amp33, c, rp, synth.FRAME_TIME, area, cfg, do_refpix=False)
        for k in ("slope", "err_read", "err_poisson", "pdq", "rdq", "endslice"):
            assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (n, k)
