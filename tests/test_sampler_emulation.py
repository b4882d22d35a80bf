"""Host emulation (NumPy float32, same operation order) of the small-expectation Poisson sampler of
`poisson_resample_kernel` (romanimpreprocess_b200/csrc/rip_sim.cu; noise directive 'P' with flag r, reference
L1_to_L2/gen_noise_image.py:258-321): inversion by sequential search with one 23-bit uniform per draw
(rip_rng.cuh `uniform`: ((x >> 9) + 0.5) * 2^-23, strictly inside (0,1) in float32), stopped in the far
tail once the float32 CDF no longer grows.  Checks that the draws have the Poisson mean and variance over the range the
kernel uses it for (0 < e < 10 electrons per sample) and that the search always terminates."""
import numpy as np
import pytest


def _draw(lam, u):
    lamf = np.float32(lam)
    term = np.full(u.shape, np.exp(-lamf), np.float32)
    cdf = term.copy()
    k = np.zeros(u.shape, np.int32)
    act = u > cdf
    steps = 0
    while act.any():
        steps += 1
        assert steps < 200, "search did not terminate"
        k[act] += 1
        term[act] = term[act] * (lamf / k[act].astype(np.float32))
        nxt = (cdf + term).astype(np.float32)
        stalled = act & (nxt == cdf)
        cdf = np.where(act, nxt, cdf)
        act = act & ~stalled & (u > cdf)
    return k


@pytest.mark.parametrize("lam", [0.05, 0.7, 3.0, 6.08, 9.99])
def test_inversion_sampler_moments(lam):
    rng = np.random.default_rng(int(lam * 100))
    n = 1_000_000
    u = ((rng.integers(0, 1 << 23, n) + 0.5) / 8388608.0).astype(np.float32)
    assert u.min() > 0.0 and u.max() < 1.0
    k = _draw(lam, u)
    # mean: 5 sigma of the sample mean; variance: Var(s^2) ~ (lam + 2 lam^2)/n for a Poisson variate
    assert abs(k.mean() - lam) < 5 * np.sqrt(lam / n)
    assert abs(k.var() - lam) < 5 * np.sqrt((lam + 2 * lam * lam) / n)


def test_inversion_sampler_extreme_uniforms():
    # the largest and smallest uniforms the generator can return: the search stops (far tail) and returns a plausible count
    u = np.array([(0 + 0.5) / 8388608.0, (8388607 + 0.5) / 8388608.0], np.float32)
    assert 0.0 < u[0] and u[1] < 1.0  # (the 24-bit form this replaced rounded its top value to exactly 1.0f)
    for lam in (1e-6, 0.5, 9.99):
        k = _draw(lam, u)
        assert k[0] == 0 and 0 <= k[1] < 64
