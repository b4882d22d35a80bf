"""CPU: the host part of the sky fit that lives in the library (rip_medfit_solve, csrc/rip_sky.cu) against the reference's
NumPy / SciPy lines (utils/sky.py:137-175, restated in oracle.medfit and pinned by tests/golden/sky_medfit.npz)."""

import numpy as np
from scipy.special import legendre_p
from scipy.stats import norm

from romanimpreprocess_b200.utils import sky


def _reference_solve(meds, N, nx, ny, order):
    kx, ky = nx // N, ny // N
    px, py = (nx % N) // 2, (ny % N) // 2
    u_ = 2 * (px - 0.5 + kx * np.linspace(0.5, N - 0.5, N)) / nx - 1
    v_ = 2 * (py - 0.5 + ky * np.linspace(0.5, N - 0.5, N)) / ny - 1
    u, v = np.meshgrid(u_, v_)
    nc = (order + 1) * (order + 2) // 2
    basis = np.zeros((nc, N, N))
    k = 0
    for i in range(order + 1):
        t = legendre_p(i, u)
        for j in range(order + 1 - i):
            basis[k] = t * legendre_p(j, v)
            k += 1
    A, b = np.zeros((nc, nc)), np.zeros(nc)
    for ip in range(N):
        for jp in range(N):
            if not np.isnan(meds[jp, ip]):
                A += np.outer(basis[:, jp, ip], basis[:, jp, ip])
                b += meds[jp, ip] * basis[:, jp, ip]
    return np.linalg.solve(A, b)


def test_medfit_solve_against_numpy():
    rng = np.random.default_rng(1)
    for ny, nx, N, order in [(4088, 4088, 8, 2), (248, 248, 8, 2), (500, 377, 8, 3), (4088, 4088, 8, 0), (4088, 4088, 6, 4)]:
        meds = rng.normal(1.0, 0.1, (N, N)).astype(np.float32)
        meds[1, 2] = np.nan
        coef, LPX, LPY = sky._solve(meds, N, nx, ny, order)
        x = _reference_solve(meds, N, nx, ny, order)
        assert np.max(np.abs(coef - x)) <= 1e-14 * np.abs(x).max()
        # grid polynomials: bit-identical to scipy.special.legendre_p on np.linspace (utils/sky.py:167-175)
        for i in range(order + 1):
            assert np.array_equal(LPX[i], np.ravel(legendre_p(i, np.linspace(-1, 1 - 2 / nx, nx)))), (nx, i)
            assert np.array_equal(LPY[i], np.ravel(legendre_p(i, np.linspace(-1, 1 - 2 / ny, ny)))), (ny, i)


def test_norm_ppf():
    for p in (0.75, 0.9, 0.6, 0.01, 0.999, 0.5):
        assert abs(sky._norm_ppf(p) - norm.ppf(p)) < 1e-14 * max(1.0, abs(norm.ppf(p)))


def test_tilde_nus_against_the_reference_golden():
    """gen_noise_image.tilde_nus == the reference's GalPoisson/find_tilnus.get_tilde_nus (golden: make_golden_tilnus.py)."""
    from conftest import load_golden
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.L1_to_L2 import gen_noise_image as gni

    g = load_golden("tilnus")
    for name in ("README_PATTERN", "TEST_READ_PATTERN", "LONG16_PATTERN"):
        for j in range(3):
            t = gni.tilde_nus(getattr(synth, name), g[f"{name}_w{j}"])
            np.testing.assert_allclose(t, g[f"{name}_t{j}"][:3], rtol=1e-13, atol=0)


def test_pearson4_normalisation_against_scipy():
    """log k(m, nu) of the Pearson IV density (reference GalPoisson/draw_with_tilnus.py:296-306, there through
    scipy.special.loggamma of a complex argument) as the library evaluates it for the Type IV sampler (recurrence +
    Stirling series): 1e-8 relative over m - 1 = 1e-3 .. 1e6, |nu| = 1e-4 .. 1e6."""
    import math

    from scipy.special import loggamma

    from romanimpreprocess_b200 import _lib

    lib = _lib.lib()
    rng = np.random.default_rng(1)
    for _ in range(4000):
        m = 1.0 + 10 ** rng.uniform(-3, 6)
        nu = rng.choice([-1.0, 1.0]) * 10 ** rng.uniform(-4, 6)
        ref = (2 * m - 2) * math.log(2) + 2 * loggamma(m + 0.5j * nu).real - (math.log(math.pi) + loggamma(2 * m - 1).real)
        got = lib.rip_pearson4_logk_host(m, nu)
        assert abs(got - ref) <= 1e-8 * max(1.0, abs(ref)), (m, nu, ref, got)


def test_pearson4_devroye_sampler_host_emulation():
    """The Type IV sampler of rip_pearson_noise_dev (csrc/rip_sim.cu pearson4_draw: Devroye's rejection method for
    log-concave densities on the angle, Heinrich 2004 section 7, as the reference's pt4_rvs_devroye,
    GalPoisson/draw_with_tilnus.py:444-483, but with the exact hat constant rc = 1 / g(mode)) restated with NumPy around the
    library's own log k: the accepted draws follow the Pearson IV law (Kolmogorov-Smirnov against the integrated density)
    and at least a quarter of the proposals are accepted for every (m, nu) -- the bound of the method when rc is exact."""
    from romanimpreprocess_b200 import _lib

    lib = _lib.lib()
    rng = np.random.default_rng(7)
    th = np.linspace(-np.pi / 2, np.pi / 2, 400001)[1:-1]
    for m, nu in ((2.51, 0.13), (2.75, -0.14), (6.7, 0.23), (44.4, -1.77), (421.6, 17.1), (41917.0, -1707.0)):
        b = 2 * m - 2
        M = np.arctan2(-nu, b)
        r_const = b * np.log(b / np.hypot(b, nu)) - nu * M
        rc = np.exp(-r_const - lib.rip_pearson4_logk_host(m, nu))
        n = 60000
        x = 4 * rng.random(n)
        right = x > 2
        x = np.where(right, x - 2, x)
        tail = x > 1
        with np.errstate(divide="ignore"):
            z = np.where(tail, np.log(np.where(tail, x - 1, 1.0)), 0.0)
        x = np.where(tail, 1 - z, x)
        t = np.where(right, M + rc * x, M - rc * x)
        inside = np.abs(t) < np.pi / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = inside & ~(z + np.log(rng.random(n)) > b * np.log(np.cos(np.where(inside, t, 0.0))) - nu * t - r_const)
        acc = ok.mean()
        assert acc > 0.24, (m, nu, acc)
        t = np.sort(t[ok])
        logg = b * np.log(np.cos(th)) - nu * th
        c = np.cumsum(np.exp(logg - logg.max()))
        F = np.interp(t, th, c / c[-1])
        ks = np.max(np.abs(F - (np.arange(t.size) + 0.5) / t.size))
        assert ks < 1.95 / np.sqrt(t.size), (m, nu, ks, acc)
