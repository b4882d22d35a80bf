"""GPU: the forward-path rows either side of make_l1_fullcal (a15, a20) and the many-realisations
bookkeeping (BASELINE configs[4]).

Deterministic steps are checked EXACTLY (golden vectors made by the unmodified reference functions, or the oracle
on the same inputs); everything driven by random numbers is validated statistically against the oracle's ensembles
(the kernels use Philox, the reference GalSim deviates)."""

import ctypes as C

import numpy as np
import pytest
from conftest import load_golden

from oracle import rip_oracle as orc

pytestmark = pytest.mark.gpu


def _small_cal(n, G, seed):
    from romanimpreprocess_b200 import synth

    pattern = [[0], [1, 2], [3, 4, 5, 6]][:G]
    return synth.make_caldir(n=n, read_pattern=pattern, p_order=3, seed=seed), pattern


def _trees(cal):
    return {k: v["roman"] for k, v in cal.items()}


def test_sim_calprep_exact():
    """this_dark / this_flat of Image2D.simulate (sim_to_isim.py:615-633) == the reference's own ipc_rev + clips."""
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    g = load_golden("sim_refdata_n256")
    cal, _ = _small_cal(int(g["n"]), int(g["G"]), int(g["seed"]))
    d, f = s2i.sim_calprep(cal)
    # the fixture's gain plane is float64, so the reference's planes are float64; the library returns float32
    assert np.array_equal(f, g["this_flat"].astype(np.float32))
    assert np.allclose(d, g["this_dark"], rtol=1e-5, atol=1e-7)  # (the float64 product dark*gain is rounded to float32 first)
    # all-float32 CALDIR: identical to the oracle (itself pinned to the reference on the float64 fixture)
    from romanimpreprocess_b200 import synth

    cal32 = synth.make_caldir(n=128, read_pattern=[[0], [1, 2]], p_order=3, seed=5, gain_dtype=np.float32)
    d32, f32 = s2i.sim_calprep(cal32)
    od, of, _ = orc.sim_calprep(_trees(cal32))
    assert od.dtype == np.float32 and np.array_equal(f32, of) and np.array_equal(d32, od)


@pytest.mark.parametrize("n", [256, 4096])
def test_noise_1f_frame_from_draws(n):
    """Same N(0,1) stream as the reference call -> same block up to the float32 FFT rounding (the reference
    transforms in complex128 and casts the block to float32)."""
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    cw = n // 32
    m = 2 * n * cw
    draws = np.zeros(2 * m)
    orc.NormalStream(11).generate(draws)
    ref = orc.noise_1f_frame(orc.NormalStream(11), n, cw)
    if n == 256:
        assert np.array_equal(ref, load_golden("sim_refdata_n256")["frame_seed11"])
    out = s2i.noise_1f_frame(None, nside=n, draws=draws[None])
    assert out.shape == ref.shape and out.dtype == np.float32
    assert np.abs(out - ref).max() < 2e-5 * ref.std() * np.log2(m)


def test_noise_1f_frame_statistics():
    """Philox-driven blocks: zero mean per block, the variance and the 1/f spectrum of the oracle's ensemble."""
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, nf = 256, 48
    out = s2i.noise_1f_frame(None, nside=n, seed=5, nframes=nf)
    ref = np.array([orc.noise_1f_frame(orc.NormalStream(100 + i), n, n // 32) for i in range(nf)])
    assert np.all(np.abs(out.reshape(nf, -1).mean(axis=1)) < 1e-5)
    assert abs(out.std() / ref.std() - 1.0) < 0.08  # (the lowest modes dominate the variance: large sample scatter)
    assert not np.allclose(out[0], out[1])

    def spectrum(fr):
        p = (np.abs(np.fft.rfft(fr.reshape(nf, -1).astype(np.float64), axis=1)) ** 2).mean(axis=0)[1:]
        edges = np.unique(np.round(np.logspace(0, np.log10(p.size), 14)).astype(int))
        return np.array([p[a:b].mean() for a, b in zip(edges[:-1], edges[1:])])

    ps, pr = spectrum(out), spectrum(ref)
    assert np.all(np.abs(ps / pr - 1.0) < 0.25), ps / pr
    assert ps[0] / ps[-1] > 50  # red spectrum


def test_fill_refdata_deterministic_part_exact():
    """With the white-noise planes set to zero and no banding, reference pixels = round(dark cube) and active pixels
    are untouched: identical to the oracle (and to the reference, via the golden pinning of the oracle)."""
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, G = 256, 3
    cal, pattern = _small_cal(n, G, 4242)
    cal["read"]["roman"]["data"] = np.zeros_like(cal["read"]["roman"]["data"])
    cal["read"]["roman"]["resetnoise"] = np.zeros_like(cal["read"]["roman"]["resetnoise"])
    tij = orc.read_pattern_to_tij(pattern)
    im0 = np.random.RandomState(1).randint(0, 60000, size=(G, n, n)).astype(np.uint16)
    im_o, im_g = im0.copy(), im0.copy()
    orc.fill_in_refdata_and_1f(im_o, _trees(cal), orc.NormalStream(3), tij, fill_in_banding=False)
    s2i.fill_in_refdata_and_1f(im_g, cal, 7, tij, fill_in_banding=False)
    assert np.array_equal(im_g, im_o)
    assert np.array_equal(im_g[:, 4:-4, 4:-4], im0[:, 4:-4, 4:-4])


def test_fill_refdata_statistics():
    """Reference-pixel white noise, banding amplitude / common mode / mirroring, and the reference output, against
    the oracle's ensemble statistics."""
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, G = 256, 3
    cal, pattern = _small_cal(n, G, 4242)
    c = _trees(cal)
    cw = n // 32
    tij = orc.read_pattern_to_tij(pattern)
    base = np.full((G, n, n), 30000, np.uint16)
    # (1) white part of the reference pixels: z = (value - dark) / sigma ~ N(0, 1 + 1/12 rounding)
    im = base.copy()
    s2i.fill_in_refdata_and_1f(im, cal, 11, tij, fill_in_banding=False)
    border = np.ones((n, n), bool)
    border[4:-4, 4:-4] = False
    dk = c["dark"]["data"][-G:]
    for g in range(G):
        sig = np.sqrt(c["read"]["data"] ** 2 / len(pattern[g]) + c["read"]["resetnoise"] ** 2)
        z = ((im[g].astype(np.float64) - dk[g]) / sig)[border]
        assert abs(z.mean()) < 4 / np.sqrt(z.size) + 0.01
        assert abs(z.var() - 1.0) < 0.05
    # the reset layer is common to all groups: differences of groups lose it
    d01 = (im[0].astype(np.float64) - dk[0]) - (im[1].astype(np.float64) - dk[1])
    expect = c["read"]["data"] ** 2 * (1.0 / len(pattern[0]) + 1.0 / len(pattern[1]))
    assert abs((d01[border] ** 2 / expect[border]).mean() - 1.0) < 0.08
    # (2) banding on a constant cube: compare with the oracle ensemble
    nr = 6
    gb, ob = [], []
    for r in range(nr):
        a = base.copy()
        a33 = np.zeros((G, n, cw), np.uint16)
        s2i.fill_in_refdata_and_1f(a, cal, 100 + r, tij, fill_in_banding=True, amp33=a33)
        gb.append(a[:, 4:-4, 4:-4].astype(np.float64) - 30000)
        b = base.copy()
        orc.fill_in_refdata_and_1f(b, c, orc.NormalStream(500 + r), tij, fill_in_banding=True)
        ob.append(b[:, 4:-4, 4:-4].astype(np.float64) - 30000)
    gb, ob = np.array(gb), np.array(ob)
    for g in range(G):
        assert abs(gb[:, g].std() / ob[:, g].std() - 1.0) < 0.15, (g, gb[:, g].std(), ob[:, g].std())

    def chan(a, ch):  # active rows of channel ch (interior channels only: no border columns)
        return a[..., :, cw * ch - 4 : cw * (ch + 1) - 4]

    def corr(x, y):
        x, y = x - x.mean(), y - y.mean()
        return (x * y).mean() / np.sqrt((x * x).mean() * (y * y).mean())

    u, cc = float(c["read"]["anc"]["U_PINK"]), float(c["read"]["anc"]["C_PINK"])
    rho = cc**2 / (cc**2 + u**2)
    same = corr(chan(gb, 2), chan(gb, 4))
    mirrored = corr(chan(gb, 2), chan(gb, 3)[..., ::-1])
    unmirrored = corr(chan(gb, 2), chan(gb, 3))
    assert abs(same - rho) < 0.08 and abs(mirrored - rho) < 0.08, (same, mirrored, rho)
    assert unmirrored < mirrored - 0.05


def test_fill_refdata_amp33_full_frame():
    """4096^2 (128-column channels): the reference output follows med + noise with the oracle's scatter."""
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, G = 4096, 2
    pattern = [[0], [1, 2]]
    cal = synth.make_caldir(n=n, read_pattern=pattern, p_order=3, seed=9)
    c = _trees(cal)
    tij = orc.read_pattern_to_tij(pattern)
    im = np.full((G, n, n), 20000, np.uint16)
    a33 = np.zeros((G, n, 128), np.uint16)
    s2i.fill_in_refdata_and_1f(im, cal, 21, tij, fill_in_banding=True, amp33=a33)
    imo = np.full((G, n, n), 20000, np.uint16)
    a33o = np.zeros((G, n, 128), np.uint16)
    orc.fill_in_refdata_and_1f(imo, c, orc.NormalStream(8), tij, fill_in_banding=True, amp33=a33o)
    med = c["read"]["amp33"]["med"]
    for g in range(G):
        rg, ro = a33[g].astype(np.float64) - med, a33o[g].astype(np.float64) - med
        assert abs(rg.mean() - ro.mean()) < 0.3  # truncation bias -0.5 in both
        assert abs(rg.std() / ro.std() - 1.0) < 0.1, (rg.std(), ro.std())
        bg, bo = im[g, 4:-4, 4:-4].astype(np.float64) - 20000, imo[g, 4:-4, 4:-4].astype(np.float64) - 20000
        assert abs(bg.std() / bo.std() - 1.0) < 0.2, (bg.std(), bo.std())
    # reference output and science channels share the common mode (M_PINK * C_PINK * common)
    row_sci = (im[1, 4:-4, 4:-4].astype(np.float64) - 20000).mean(axis=1)
    row_ref = (a33[1, 4:-4].astype(np.float64) - med[4:-4]).mean(axis=1)
    row_sci_o = (imo[1, 4:-4, 4:-4].astype(np.float64) - 20000).mean(axis=1)
    row_ref_o = (a33o[1, 4:-4].astype(np.float64) - med[4:-4]).mean(axis=1)
    cg, co = np.corrcoef(row_sci, row_ref)[0, 1], np.corrcoef(row_sci_o, row_ref_o)[0, 1]
    assert cg > 0.3 and abs(cg - co) < 0.25, (cg, co)


def test_sim_counts_statistics():
    """Poisson mean == the oracle's scene_rate; counts are Poisson with that mean (both samplers: lam < 10 and PTRS)."""
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, G = 256, 3
    cal, pattern = _small_cal(n, G, 4242)
    c = _trees(cal)
    na = n - 8
    yy, xx = np.mgrid[0:na, 0:na]
    image = (0.01 + 0.02 * (xx % 7) + 40.0 * np.exp(-0.5 * ((xx - 120) ** 2 + (yy - 90) ** 2) / 30.0**2)).astype(np.float32)
    image[:3] = -1.0  # negative scene values clip to zero electrons
    area = (1.0 + 0.01 * np.sin(xx / 40.0)).astype(np.float64)
    counts, rate = s2i.simulate_counts(image, cal, pattern, seed=3, area_ratio=area, cnorm=0.9, return_rate=True)
    this_dark, this_flat, g = orc.sim_calprep(c)
    t = 3.04 * (pattern[-1][-1] - pattern[0][0])
    ref_rate = orc.scene_rate(image, this_flat, g, area, t, cnorm=0.9)
    assert np.allclose(rate, ref_rate, rtol=2e-6, atol=1e-9)
    assert counts.dtype == np.int32 and np.all(counts >= 0) and np.all(counts[:3] == 0)
    for lo, hi in ((0.05, 10.0), (10.0, 1e9)):
        w = (rate > lo) & (rate < hi)
        assert w.sum() > 5000
        z = (counts[w] - rate[w]) / np.sqrt(rate[w])
        assert abs(z.mean()) < 5 / np.sqrt(w.sum()), (lo, z.mean())
        assert abs(z.var() - 1.0) < 0.05, (lo, z.var())
    # dark term + accumulation
    c2 = s2i.simulate_counts(np.zeros_like(image), cal, pattern, seed=4, counts=counts, dark=True)
    d = (c2 - counts).astype(np.float64)
    lam = np.clip(this_dark, 0, None) * t
    assert np.all(d >= 0) and abs(d.mean() - lam.mean()) < 5 * np.sqrt(lam.mean() / d.size) + 1e-3
    # sky background term of romanisim's simulate_counts (restated): Poisson(sky * this_flat * t) on top
    c3 = s2i.simulate_counts(np.zeros_like(image), cal, pattern, seed=5, area_ratio=area, sky=1.5)
    lam3 = 1.5 * this_flat.astype(np.float64) * t
    z3 = (c3 - lam3) / np.sqrt(lam3)
    assert abs(z3.mean()) < 5 / np.sqrt(z3.size) and abs(z3.var() - 1.0) < 0.05, (z3.mean(), z3.var())


def test_mask_build_exact():
    from romanimpreprocess_b200.utils import maskhandling

    g = load_golden("mask_moments")
    assert np.array_equal(maskhandling.PixelMask1.build(g["dq"]), g["mask_pixelmask1"])
    custom = maskhandling.CombinedMask({"jump_det": 25, "hot": 5, 7: 9, "saturated": 1})
    assert np.array_equal(custom.build(g["dq"]), g["mask_custom"])
    big = np.random.RandomState(5).randint(0, 2**32, size=(300, 517), dtype=np.uint64).astype(np.uint32)
    big &= np.random.RandomState(6).randint(0, 2**32, size=big.shape, dtype=np.uint64).astype(np.uint32)
    big &= np.random.RandomState(7).randint(0, 2**32, size=big.shape, dtype=np.uint64).astype(np.uint32)
    big &= np.random.RandomState(8).randint(0, 2**32, size=big.shape, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(maskhandling.PixelMask1.build(big), orc.mask_build(big))


def test_moments_and_median_exact():
    """Moment sums / finalisation bit-exact against the reference's lines; stack median == np.median."""
    import torch

    from romanimpreprocess_b200 import _lib
    from romanimpreprocess_b200.utils import maskhandling

    g = load_golden("mask_moments")
    nm = g["dq"].shape[0]
    n, nb = nm + 8, 4
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    mom = torch.zeros((3, nm, nm), dtype=torch.float32, device=dev)
    grow = np.ascontiguousarray(maskhandling.PixelMask1.array)
    for dat, dq in zip(g["data"], g["dqs"]):
        full = np.full((n, n), np.nan, np.float32)
        full[nb:-nb, nb:-nb] = dat
        fdq = np.full((n, n), 0xFFFFFFFF, np.uint32)  # border flags must not leak into the window
        fdq[nb:-nb, nb:-nb] = dq
        ds = torch.from_numpy(full).to(dev)
        dd = torch.from_numpy(fdq.view(np.int32)).to(dev)
        _lib.check(lib.rip_moments_accumulate_dev(0, C.c_void_p(ds.data_ptr()), C.c_void_p(dd.data_ptr()), n, nb,
                                                  _lib.ptr(grow), C.c_void_p(mom.data_ptr()), None))  # fmt: skip
    torch.cuda.synchronize()
    assert np.array_equal(mom.cpu().numpy(), g["moments_sum"])
    _lib.check(lib.rip_moments_finalize_dev(0, C.c_void_p(mom.data_ptr()), nm * nm, None))
    torch.cuda.synchronize()
    assert np.array_equal(mom.cpu().numpy(), g["moments_final"])
    rng = np.random.RandomState(3)
    for R in (1, 2, 5, 8, 9, 33, 64, 100):
        st = rng.randn(R, 5000).astype(np.float32)
        st[:, 7] = 1.0
        if R > 2:
            st[1, 11] = np.nan
            st[2, 13] = np.inf
        d = torch.from_numpy(st).to(dev)
        o = torch.empty(5000, dtype=torch.float32, device=dev)
        _lib.check(lib.rip_stack_median_dev(0, C.c_void_p(d.data_ptr()), R, 5000, C.c_void_p(o.data_ptr()), None))
        torch.cuda.synchronize()
        with np.errstate(all="ignore"):
            ref = np.median(st, axis=0)
        assert np.array_equal(o.cpu().numpy(), ref, equal_nan=True), R


def _mr_diagnostics(image, cal, rp):
    """Flag census of one fresh realisation (only evaluated when the assertion below fails)."""
    import torch

    from romanimpreprocess_b200.validation_tests import many_realizations as mr

    z = mr.Realizations(image, cal, rp, keep_stacks=1)
    z.step(110)
    torch.cuda.synchronize()
    pdq = z.d_pdq.cpu().numpy().view(np.uint32)[4:-4, 4:-4]
    im = z.d_im.cpu().numpy().view(np.uint16)
    info = {"bits": {b: int(np.count_nonzero(pdq & np.uint32(1 << b))) for b in range(32) if np.any(pdq & np.uint32(1 << b))},
            "counts_mean": float(z.d_counts.float().mean().item()), "res_mean": [float(v) for v in z.d_res.mean(dim=(1, 2)).cpu()],
            "im_act_mean": [float(im[g, 4:-4, 4:-4].mean()) for g in range(im.shape[0])],
            "amp33_mean": float(z.d_amp33.cpu().numpy().view(np.uint16).mean()),
            "slope_mean": float(z.d_slope[8:-8, 8:-8].mean().item()), "m0": float(z.d_moments[0].mean().item())}
    z.close()
    return info


def test_many_realizations_small():
    """The whole protocol at 256^2: scene -> counts -> ramp -> reference pixels + 1/f -> L1->L2 -> moments.
    The mean slope recovers the scene (in DN/s) and the realisation scatter matches the reported error."""
    from romanimpreprocess_b200 import pars, synth
    from romanimpreprocess_b200.validation_tests import many_realizations as mr

    n, R = 256, 12
    rp = synth.README_PATTERN
    cal = synth.make_caldir(n=n, read_pattern=rp, p_order=10, seed=77)
    c = _trees(cal)
    na = n - 8
    yy, xx = np.mgrid[0:na, 0:na]
    image = (20.0 + 0.1 * xx).astype(np.float32)  # e/s
    slope_ideal = np.zeros((n, n), np.float32)
    slope_ideal[4:-4, 4:-4] = image / pars.g_ideal
    # (256^2 frames have 8-column channels: no 128-column reference output, so no reference-pixel correction and
    #  therefore no banding to remove; the full protocol incl. banding + refpix runs at 4096^2 in bench.py --workload realizations)
    out = mr.run(image, cal, rp, R, seed=100, slope_ideal=slope_ideal, fill_in_banding=False)
    assert out.shape == (8, n, n)
    cnt, mean, std, bias, mederr = out[3], out[4], out[5], out[6], out[7]
    act = np.zeros((n, n), bool)
    act[8:-8, 8:-8] = True
    good = act & (cnt >= R - 1)
    assert good.mean() > 0.5 * act.mean(), _mr_diagnostics(image, cal, rp)
    # dark electrons are drawn by the forward model and the dark slope is subtracted by the calibration
    rel = bias[good] / slope_ideal[good]
    assert abs(np.median(rel)) < 0.02, np.median(rel)
    ratio = np.median(std[good] / mederr[good])
    assert 0.8 < ratio < 1.25, ratio
    assert np.all(out[1][act] > 0)  # median group difference: the signal accumulates
    del c


def _ulp_diff32(a, b):
    """difference of two float32 arrays in units in the last place (NaNs must coincide)"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia, ib = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia), np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.where(np.isnan(a), 0, np.abs(ia - ib))


def test_sky_medfit_against_reference():
    """GPU medfit vs the reference's medfit (goldens made by the unmodified reference function): region nan-medians by
    radix select (incl. an all-NaN region and odd sizes) exact; coefficients from the library's normal equations + LU
    (rip_medfit_solve) within 1e-13 of NumPy's LAPACK solve; float32 model within 1 ulp, identical at > 99 % of the pixels."""
    import torch

    from conftest import SKY_CASES, synth_sky_image
    from romanimpreprocess_b200.utils import sky

    g = load_golden("sky_medfit")
    for tag, (ny, nx, order, nreg) in SKY_CASES.items():
        img = synth_sky_image(ny, nx, 40 + ord(tag))
        coef, model = sky.medfit(img, N=nreg, order=order)
        assert model.dtype == np.float32 and model.shape == img.shape
        np.testing.assert_allclose(coef, g[f"{tag}_coef"], rtol=0, atol=1e-13 * np.abs(g[f"{tag}_coef"]).max(), err_msg=tag)
        ulp = _ulp_diff32(model[::7, ::5], g[f"{tag}_model_sub"])
        assert ulp.max() <= 1 and np.count_nonzero(ulp) <= 0.01 * ulp.size, (tag, ulp.max(), np.count_nonzero(ulp))
        assert abs(model.astype(np.float64).sum() - float(g[f"{tag}_model_sum"])) <= 1e-7 * abs(float(g[f"{tag}_model_sum"])), tag
        # device-resident form on a window of a larger plane, subtracting in place (gen_cal_image.py:645-647)
        big = np.full((ny + 8, nx + 8), 7.0, np.float32)
        big[4:-4, 4:-4] = img
        d = torch.from_numpy(big).cuda()
        win = d[4:-4, 4:-4]
        c2 = sky.medfit_device(win.data_ptr(), nx + 8, ny, nx, N=nreg, order=order)
        torch.cuda.synchronize()
        out = d.cpu().numpy()
        assert np.array_equal(c2, coef)
        assert np.array_equal(out[4:-4, 4:-4], img - model, equal_nan=True)
        assert np.all(out[:4] == 7.0) and np.all(out[:, :4] == 7.0)


def test_smooth_mode_and_binning():
    """sky.smooth_mode / sky.binkxk on the GPU vs NumPy restatements of utils/sky.py:20-93 (the `medsky` of the L2 file,
    gen_cal_image.py:641): masked 4 x 4 binning identical up to float32 summation order, mode of the smoothed histogram
    to 1e-6 of its width (percentiles are exact order statistics; the Gaussian sums are float64 reductions)."""
    from scipy.stats import norm

    from romanimpreprocess_b200.utils import sky

    rng = np.random.default_rng(77)
    img = (0.8 + 0.05 * rng.standard_normal((1022, 1030)) + 3.0 * (rng.random((1022, 1030)) < 0.01)).astype(np.float32)
    mask = rng.random(img.shape) < 0.03
    b = sky.binkxk(img, 4, mask=mask)
    bref = np.mean(np.where(~mask, img, np.nan)[:1020, :1028].reshape(255, 4, 257, 4), axis=(1, 3))
    assert b.shape == bref.shape and np.array_equal(np.isnan(b), np.isnan(bref))
    np.testing.assert_allclose(b, bref, rtol=1e-6, equal_nan=True)
    arr = bref.astype(np.float32)

    def ref_smooth_mode(arr, pc=25.0, pksmooth=0.5, niter=3):
        c1, c2, c3 = (np.nanpercentile(arr, q) for q in (pc, 50.0, 100.0 - pc))
        ctr, sigma = c2, (c3 - c1) / (norm.ppf((100.0 - pc) / 100.0) * 2)
        N = 21
        for _ in range(niter):
            hs = np.zeros(N)
            z = ctr + np.linspace(-1, 1, N) * sigma
            for i in range(1, N - 1):
                w = np.exp(-0.5 * ((z[i] - arr) / (pksmooth * sigma)) ** 2)
                hs[i] = np.sum(np.where(np.isnan(w), 0.0, w))
            ip = np.argmax(hs)
            bb, aa = (hs[ip + 1] - hs[ip - 1]) / 2.0, (hs[ip + 1] + hs[ip - 1]) / 2.0 - hs[ip]
            ctr = z[ip] + (z[1] - z[0]) * (-bb / 2.0 / aa)
        return ctr, sigma * pksmooth

    for kw in ({}, {"pc": 10.0, "pksmooth": 0.3, "niter": 4}):
        m, w = sky.smooth_mode(arr, **kw)
        mr, wr = ref_smooth_mode(arr, **kw)
        assert abs(w - wr) <= 1e-6 * wr and abs(m - mr) <= 1e-6 * wr, (m, mr, w, wr)
    assert abs(m - 0.8) < 0.01


def test_noise_layers_against_oracle_composition():
    """gen_noise_image 'R' layers (reference L1_to_L2/gen_noise_image.py:60-163, 322-326) at 256^2: the layer built on
    the GPU has the statistics of the same composition done with the oracle (dark cube -> white noise ->
    fill_in_refdata_and_1f -> L1->L2 -> difference); 'z' clips, 'S' removes the sky modes, 'a' starts from the exposure."""
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.L1_to_L2 import gen_noise_image as gni
    from romanimpreprocess_b200.utils import sky

    n = 256
    rp = synth.README_PATTERN
    G = len(rp)
    cal = synth.make_caldir(n=n, read_pattern=rp, p_order=10, seed=31, gain_dtype=np.float32, ipc_dtype=np.float32)
    c = _trees(cal)
    data, amp33, _ = synth.make_l1(cal, rp, seed=32, n_sources=4, cr_frac=0.0, bright=0.5)
    cfg = {}
    layers = gni.make_noise_cube_arrays(data, amp33, cal, rp, synth.FRAME_TIME, ["R", "Rz3", "RaS2", "RS1C7"], seed=5, config=cfg)
    assert layers.shape == (4, n - 8, n - 8) and layers.dtype == np.float32
    # oracle composition of the plain 'R' layer
    tij = orc.read_pattern_to_tij(rp)
    dk = c["dark"]["data"]
    base = dk[dk.shape[0] - G :].astype(np.uint16)
    one = np.ones((n, n), np.float32)
    ref = orc.l1_to_l2(base, None, c, rp, synth.FRAME_TIME, one, cfg, do_refpix=False)
    rng = np.random.default_rng(3)
    noisy = base.copy()
    for k in range(G):
        res = noisy[k, 4:-4, 4:-4].astype(np.float32)
        im = rng.standard_normal(res.shape).astype(np.float32)
        im *= c["read"]["data"][4:-4, 4:-4] / np.sqrt(len(rp[k]))
        noisy[k, 4:-4, 4:-4] = np.round(np.clip(res + im, 0, 2**16 - 1)).astype(np.uint16)
    orc.fill_in_refdata_and_1f(noisy, c, orc.NormalStream(4), tij, fill_in_banding=True)
    out = orc.l1_to_l2(noisy, None, c, rp, synth.FRAME_TIME, one, cfg, do_refpix=False)
    odiff = (out["slope"] - ref["slope"])[4:-4, 4:-4]
    inner = (slice(8, -8), slice(8, -8))
    g0 = layers[0]
    okpix = np.isfinite(odiff) & np.isfinite(g0)
    assert okpix.mean() > 0.99

    def robust_sigma(a):
        q = np.percentile(a, [25, 75])
        return (q[1] - q[0]) / 1.34896

    so, sg = robust_sigma(odiff[okpix]), robust_sigma(g0[okpix])
    assert abs(sg / so - 1.0) < 0.08, (sg, so)
    assert abs(np.median(g0[okpix])) < 0.2 * sg
    # the white part dominates and matches the reported read-noise error of the fit
    assert 0.7 < sg / np.median(out["err_read"][4:-4, 4:-4][okpix]) < 1.6
    # z clip: nothing beyond 3 "sigma" of the inter-quartile range
    z = layers[1]
    iqr = np.percentile(z, 75) - np.percentile(z, 25)
    assert np.all(np.abs(z - np.percentile(z, 50)) <= 3 * iqr / 1.34896 * (1 + 1e-5))
    # S: the fitted low-order model of the result is (numerically) gone
    for lay, order in ((layers[2], 2), (layers[3], 1)):
        coef, model = sky.medfit(np.ascontiguousarray(lay), order=order)
        assert np.abs(model).max() < 0.02 * robust_sigma(lay[inner])
    # 'a': built on the exposure itself -> same noise level
    assert abs(robust_sigma(layers[2][inner]) / sg - 1.0) < 0.25
    # the second production family runs too (directive 'O': test_noise_layer_pearson_directive_moments)
    lo = gni.make_noise_cube_arrays(data, amp33, cal, rp, synth.FRAME_TIME, ["Rz4OS2C5"], seed=5, config=cfg)
    assert lo.shape == (1, n - 8, n - 8) and np.all(np.isfinite(lo)) and 0.8 < robust_sigma(lo[0][inner]) / sg < 2.0


def test_noise_layer_poisson_resampling():
    """Directive 'Pbr' (reference gen_noise_image.py:187-321): re-sampled Poisson noise of the sky level through the
    ramp-fit weights.  For a full ramp the layer is sum_i' A_i' x_i' with x = (Poisson(e) - e)/gain i.i.d. per sample and
    A_i' = sum_j w_j/N_j #{i in group j, i >= i'}: zero mean and variance e/gain^2 sum A^2, pixel by pixel."""
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
    from romanimpreprocess_b200.L1_to_L2 import gen_noise_image as gni
    from romanimpreprocess_b200.utils import sky

    n = 256
    rp = synth.README_PATTERN
    G = len(rp)
    cal = synth.make_caldir(n=n, read_pattern=rp, p_order=10, seed=31, gain_dtype=np.float32, ipc_dtype=np.float32)
    c = _trees(cal)
    data, amp33, _ = synth.make_l1(cal, rp, seed=32, n_sources=0, cr_frac=0.0, bright=0.5, sky=2.0)
    cfg = {}
    lay = gni.make_noise_cube_arrays(data, amp33, cal, rp, synth.FRAME_TIME, ["Pbr", "Pbr"], seed=9, config=cfg)
    assert not np.array_equal(lay[0], lay[1])
    with gci.CalDir(cal) as cd:
        out = gci.calibrate_arrays(cd, data, None, rp, synth.FRAME_TIME, None, {"SLICEOUT": True}, do_refpix=False)
        K = np.asarray(out["meta"]["K"], np.float64)
    skylevel = sky.medfit(np.ascontiguousarray(out["slope"][4:-4, 4:-4]), order=0)[1]
    gain = np.clip(c["gain"]["data"][4:-4, 4:-4], 1e-4, 1e4).astype(np.float64)
    nsamp = rp[-1][-1] + 1
    A = np.zeros(nsamp)
    for j, grp in enumerate(rp):
        for i in grp:
            A[: i + 1] += K[j] / len(grp)
    var = np.clip(skylevel.astype(np.float64) * gain * synth.FRAME_TIME, 0, None) / gain**2 * np.sum(A**2)
    full = out["endslice"] <= 0
    assert full.mean() > 0.9 and var[full].min() > 0
    z = np.concatenate([(lay[k] / np.sqrt(var))[full] for k in range(2)])
    assert abs(z.mean()) < 5 / np.sqrt(z.size) + 0.01, z.mean()
    assert abs(z.var() - 1.0) < 0.05, z.var()


def test_percentiles_device():
    """Order statistics by radix select + NumPy's linear interpolation == np.percentile (float32 rounding)."""
    import torch

    from romanimpreprocess_b200.utils import sky

    rng = np.random.RandomState(8)
    for count in (1, 2, 1001, 300000):
        a = (rng.randn(count) * 3 - 1).astype(np.float32)
        if count > 10:
            a[5] = -0.0
            a[7] = 1.0e30
        d = torch.from_numpy(a).cuda()
        got = sky.percentiles_device(d.data_ptr(), count, (0, 25, 50, 75, 100))
        ref = np.percentile(a.astype(np.float64), [0, 25, 50, 75, 100])
        assert np.allclose(np.array(got, np.float64), ref, rtol=2e-7, atol=0), (count, got, ref)
    a[3] = np.nan
    d = torch.from_numpy(a).cuda()
    assert all(np.isnan(v) for v in sky.percentiles_device(d.data_ptr(), a.size, (25, 75)))


def test_noise_layer_pearson_directive_moments():
    """Directive 'O' (reference gen_noise_image.py:173-227, GalPoisson/draw_with_tilnus.py): the draws have the moments
    the Pearson family is solved for -- variance nu21 I, third central moment nu31 I, fourth 3 nu21^2 I^2 + nu41 I (in
    electrons; the layer is draw / gain) -- for ramps ending at two different groups, and 0 where the admissibility test
    of the reference fails; Types VI and IV for positive nu41."""
    import torch

    from romanimpreprocess_b200 import _lib, synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
    from romanimpreprocess_b200.L1_to_L2 import gen_noise_image as gni

    n, rp = 520, synth.README_PATTERN
    G, na = len(rp), n - 8
    cal = synth.make_caldir(n=n, read_pattern=rp, p_order=3, seed=33, gain_dtype=np.float32, ipc_dtype=np.float32)
    gain = np.clip(_trees(cal)["gain"]["data"][4:-4, 4:-4], 1e-4, 1e4).astype(np.float64)
    levels = [1e-4, 0.5, 5.0, 50.0, 5000.0]  # electrons/s of gain * data_withsky, one band of rows each
    band = na // len(levels)
    I = np.zeros((na, na))
    for k, v in enumerate(levels):
        I[k * band : (k + 1) * band if k < len(levels) - 1 else na] = v
    withsky = (I / gain).astype(np.float32)
    ends = np.full((na, na), -1, np.int8)  # -1 / 0: the full ramp (row G-1 of the table)
    ends[:, na // 2 :] = 2  # right half: ramps truncated after group 2 (two-point weights)
    meta = gci.exposure_meta(rp, synth.FRAME_TIME)
    plan, _ = gci.ramp_setup(meta, {})
    K = np.array([plan.var_K[0][j] for j in range(G)], np.float32)
    kt = np.zeros(G, np.float32)
    kt[2] = 1.0 / (meta["tbar"][2] - meta["tbar"][1])
    kt[1] = -kt[2]
    tab, defined = np.zeros((G, 3)), np.zeros(G, np.uint8)
    for i, w in ((G - 1, K), (2, kt)):
        t = gni.tilde_nus(rp, w)
        tab[i] = (t[0] * synth.FRAME_TIME, t[1] * synth.FRAME_TIME**2, t[2] * synth.FRAME_TIME**3)
        defined[i] = 1
    dev = torch.device("cuda", 0)
    with gci.CalDir(cal) as cd:
        def run(seed, tab=tab):
            d_ws, d_es = torch.from_numpy(withsky).to(dev), torch.from_numpy(ends).to(dev)
            d_diff = torch.zeros((na, na), dtype=torch.float32, device=dev)
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.check(_lib.lib().rip_pearson_noise_dev(cd.handle, d_ws.data_ptr(), d_es.data_ptr(), G, 1, _lib.ptr(tab),
                                                        _lib.ptr(defined), seed, d_diff.data_ptr(), bad.data_ptr(), None))
            torch.cuda.synchronize()
            return d_diff.cpu().numpy().astype(np.float64) * gain, int(bad.item())

        e, nbad = run(5)
        e2, _ = run(6)
        assert nbad == 0 and not np.array_equal(e, e2)
        for half, row in ((np.s_[:, : na // 2], G - 1), (np.s_[:, na // 2 :], 2)):
            n21, n31, n41 = tab[row]
            for k, v in enumerate(levels):
                x = e[half][k * band + 1 : (k + 1) * band - 1].ravel()
                b1, b2 = n31**2 / (n21**3 * max(v, 0.01)), (3 * n21**2 * max(v, 0.01) + n41) / (n21**2 * max(v, 0.01))
                if not ((b2 > 0) and (b2 > b1 + 1) and (b2 > 0.75 * b1)):
                    assert np.all(x == 0.0), (row, v)  # inadmissible: the reference leaves these pixels at zero
                    continue
                Iv, N = max(v, 0.01), x.size
                m2, m3, m4 = n21 * Iv, n31 * Iv, 3 * n21**2 * Iv**2 + n41 * Iv
                assert abs(x.mean()) < 5 * np.sqrt(m2 / N), (row, v, x.mean())
                assert abs(x.var() / m2 - 1) < 6 * np.sqrt((m4 / m2**2 - 1) / N) + 1e-3, (row, v, x.var(), m2)
                skew, sk_hat = m3 / m2**1.5, np.mean((x - x.mean()) ** 3) / x.var() ** 1.5
                assert abs(sk_hat - skew) < 6 * np.sqrt(6.0 / N) + 0.05 * abs(skew), (row, v, sk_hat, skew)
                kurt, ku_hat = m4 / m2**2, np.mean((x - x.mean()) ** 4) / x.var() ** 2
                assert abs(ku_hat - kurt) < 6 * np.sqrt(24.0 / N) + 0.1 * abs(kurt - 3), (row, v, ku_hat, kurt)
        # every intensity from far below the admissibility boundary (shapes -> 0: two-point limit of the Beta) to bright
        # stars draws a finite value
        withsky[...] = (np.logspace(-3, 6, na * na).reshape(na, na) / gain).astype(np.float32)
        e3, nbad = run(8)
        assert nbad == 0 and np.all(np.isfinite(e3)) and np.count_nonzero(e3) > 0.5 * e3.size
        # Type VI (beta prime): a moderately positive nu41 puts beta_2 between the Type III and Type V lines; same moment
        # targets.  nu41 = 1.6 nu31^2 / nu21 gives beta_2 - 3 = 1.6 beta_1 for every intensity.
        tab6 = tab.copy()
        tab6[:, 2] = np.where(tab6[:, 0] > 0, 1.6 * tab6[:, 1] ** 2 / np.where(tab6[:, 0] > 0, tab6[:, 0], 1.0), 0.0)
        withsky[...] = (I / gain).astype(np.float32)
        e6, nbad = run(9, tab6)
        assert nbad == 0
        n21, n31, n41 = tab6[2]
        for k, v in enumerate(levels[1:4], start=1):
            x = e6[:, na // 2 :][k * band + 1 : (k + 1) * band - 1].ravel()
            m2, m3, m4 = n21 * v, n31 * v, 3 * n21**2 * v**2 + n41 * v
            b1, b2 = m3**2 / m2**3, m4 / m2**2
            assert 1.5 * b1 + 3 < b2 < (48 + 39 * b1 + 6 * (4 + b1) ** 1.5) / (32 - b1)  # the case is Type VI
            N = x.size
            assert abs(x.mean()) < 5 * np.sqrt(m2 / N) and abs(x.var() / m2 - 1) < 6 * np.sqrt((b2 - 1) / N) + 1e-3, (v, x.mean(), x.var(), m2)
            sk_hat = np.mean((x - x.mean()) ** 3) / x.var() ** 1.5
            assert abs(sk_hat - m3 / m2**1.5) < 6 * np.sqrt(6.0 / N) + 0.08 * abs(m3 / m2**1.5), (v, sk_hat, m3 / m2**1.5)
        # Type IV: a large positive nu41 puts every intensity above the Type V line, with m from 2.5 (fourth moment barely
        # finite) to 4e4 (nearly Gaussian).  The draws follow the Pearson IV law the reference solves for
        # (GalPoisson/draw_with_tilnus.py:535-598): Kolmogorov-Smirnov distance to the numerically integrated CDF.
        tab4 = tab.copy()
        tab4[:, 2] = np.abs(tab4[:, 2]) * 5.0
        e4, nbad = run(7, tab4)
        assert nbad == 0 and np.all(np.isfinite(e4))
        th = np.linspace(-np.pi / 2, np.pi / 2, 800001)[1:-1]
        ms = []
        for half, row in ((np.s_[:, : na // 2], G - 1), (np.s_[:, na // 2 :], 2)):
            n21, n31, n41 = tab4[row]
            for k, v in enumerate(levels):
                Iv = max(v, 0.01)
                x = np.sort(e4[half][k * band + 1 : (k + 1) * band - 1].ravel())
                b1, b2 = n31**2 / (n21**3 * Iv), (3 * n21**2 * Iv + n41) / (n21**2 * Iv)
                assert b2 > (48 + 39 * b1 + 6 * (4 + b1) ** 1.5) / (32 - b1) and b1 < 32  # the case is Type IV
                r = 6 * (b2 - b1 - 1) / (2 * b2 - 3 * b1 - 6)
                inner = 16 * (r - 1) - b1 * (r - 2) ** 2
                nu = (-1.0 if n31 >= 0 else 1.0) * r * (r - 2) * np.sqrt(b1) / np.sqrt(inner)
                a, m = np.sqrt(n21 * Iv * inner) / 4, r / 2 + 1
                lam = a * nu / (2 * (m - 1))
                ms.append(m)
                logg = (2 * m - 2) * np.log(np.cos(th)) - nu * th
                c = np.cumsum(np.exp(logg - logg.max()))
                F = np.interp(np.arctan((x - lam) / a), th, c / c[-1])
                N = x.size
                ks = np.max(np.abs(F - (np.arange(N) + 0.5) / N))
                assert ks < 1.95 / np.sqrt(N) + 2e-3, (row, v, m, nu, ks)  # (float32 layer / gain round trip: 2e-3)
                if m > 10:  # moments converge: variance nu21 I, and the skew has the sign of nu31
                    assert abs(x.var() / (n21 * Iv) - 1) < 6 * np.sqrt((b2 - 1) / N) + 1e-3, (row, v, x.var(), n21 * Iv)
                    assert abs(x.mean()) < 5 * np.sqrt(n21 * Iv / N)
        assert min(ms) < 2.6 and max(ms) > 1e4


def test_generate_all_noise_driver(tmp_path):
    """generate_all_noise(config) (reference gen_noise_image.py:334-391) on fixture files: one layer of each production
    family, written to config['NOISE']['OUT']; the 'O' layer has the variance of the ramp-fitted Poisson noise."""
    from fixture_files import write_exposure

    from romanimpreprocess_b200.caltree import open_tree
    from romanimpreprocess_b200.L1_to_L2 import gen_noise_image as gni

    config, cal, data, amp33, rp = write_exposure(str(tmp_path), n=256, ipc_dtype=np.float32)
    config["SKYORDER"] = 2
    config["NOISE"] = {"LAYER": ["Rz4PbrS2C1", "Rz4OS2C5", "O"], "TEMP": str(tmp_path / "temp.asdf"), "SEED": 77,
                       "OUT": str(tmp_path / "noise.asdf")}  # fmt: skip
    config["NOISE_PRECISION"] = 32
    gni.generate_all_noise(config)
    with open_tree(config["NOISE"]["OUT"]) as f:
        noise = np.asarray(f["noise"])
        assert list(f["config"]["NOISE"]["LAYER"]) == config["NOISE"]["LAYER"]
    assert noise.shape == (3, 248, 248) and noise.dtype == np.float32 and np.all(np.isfinite(noise))

    def mad(a):
        return np.median(np.abs(a - np.median(a)))

    assert 0.5 < mad(noise[0]) / mad(noise[1]) < 2.0  # both families: read noise + a Poisson-like term of the same sky
    assert 0 < mad(noise[2]) < mad(noise[1])  # the Pearson term alone is the smaller part of the layer
