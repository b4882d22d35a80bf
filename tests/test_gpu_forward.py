"""GPU: the forward ramp generator (rip_make_l1_host through ``sim_to_isim.make_l1_fullcal``).

Deterministic arithmetic (IPC -> /gain -> 24-step float64 bisection -> group mean -> + biascorr -> round) is checked
EXACTLY against the oracle by handing both the same externally apportioned cumulative counts.  The stochastic parts
(reset noise, binomial apportioning, read noise; romanisim restatements, parity unpinned) are validated
statistically against ensembles of the oracle's NumPy restatement, following the protocol of the reference's
validation_tests/many_realizations.py (per-pixel mean and variance over realisations).
"""

import numpy as np
import pytest

from oracle import rip_oracle as orc

pytestmark = pytest.mark.gpu


def _case(n, rp, p_order, gdt, kdt, seed):
    from romanimpreprocess_b200 import synth

    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=p_order, gain_dtype=gdt, ipc_dtype=kdt,
                            biascorr_amp=3.0)  # fmt: skip
    rng = np.random.default_rng(seed)
    na = n - 8
    yy, xx = np.mgrid[0:na, 0:na]
    mean = 200.0 + 60000.0 * np.exp(-0.5 * ((xx - na / 2) ** 2 + (yy - na / 3) ** 2) / 6.0**2) + 40.0 * xx
    counts = rng.poisson(mean).astype(np.int32)
    return cal, counts


@pytest.mark.parametrize("cfg", [(64, "TEST_READ_PATTERN", 3, np.float64, np.float32), (128, "README_PATTERN", 10, np.float32, np.float32),
                                 (64, "README_PATTERN", 10, np.float32, np.float64)])  # fmt: skip
def test_deterministic_chain_exact(cfg):
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, rpname, po, gdt, kdt = cfg
    rp = getattr(synth, rpname)
    cal, counts = _case(n, rp, po, gdt, kdt, 41)
    c = {k: v["roman"] for k, v in cal.items()}
    na = n - 8
    nreads = sum(len(g) for g in rp)
    cum = np.zeros((nreads, na, na), np.int64)
    ref = orc.make_l1_fullcal(counts, c, rp, np.random.default_rng(5), add_reset_noise=False, add_read_noise=False,
                              quantize=True, cum_counts_out=cum)  # fmt: skip
    out, dq = s2i.make_l1_fullcal(counts, rp, cal, seed=1, cum_counts=cum, add_reset_noise=False, add_read_noise=False)
    assert out.dtype == np.float32 and out.shape == ref.shape
    ndiff = np.count_nonzero(out != ref)
    # float64 chain, identical op order: the rounded DN must agree everywhere (a .5 tie flipping is the only escape)
    assert ndiff == 0, f"{ndiff} resultant values differ, max {np.abs(out - ref).max()}"
    assert np.array_equal(dq, np.broadcast_to(c["linearitylegendre"]["dq"][4:-4, 4:-4], dq.shape))
    ref_u, _ = orc.make_l1_fullcal(counts, c, rp, np.random.default_rng(5), add_reset_noise=False, add_read_noise=False,
                                   quantize=False), None  # fmt: skip
    out_u, _ = s2i.make_l1_fullcal(counts, rp, cal, seed=1, cum_counts=cum, add_reset_noise=False, add_read_noise=False,
                                   quantize=False)  # fmt: skip
    # (different binomial draws in ref_u: compare only the final group, whose cumulative count is the total)
    assert np.allclose(out_u[-1], ref_u[-1], rtol=0, atol=0.51 if len(rp[-1]) > 1 else 2e-3)


def test_statistics_against_oracle_ensemble():
    """R realisations of one scene: per-pixel mean / variance of every resultant, GPU (Philox) vs oracle (NumPy RNG)."""
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    n, rp, R = 64, synth.README_PATTERN, 64
    cal, counts = _case(n, rp, 10, np.float32, np.float32, 43)
    c = {k: v["roman"] for k, v in cal.items()}
    na = n - 8
    G = len(rp)
    gpu = np.empty((R, G, na, na), np.float32)
    cpu = np.empty((R, G, na, na), np.float32)
    rng = np.random.default_rng(99)
    with gci.CalDir(cal) as cd:
        for r in range(R):
            gpu[r], _ = s2i.make_l1_fullcal(counts, rp, cd, seed=100 + 10 * r)
            cpu[r] = orc.make_l1_fullcal(counts, c, rp, rng)
        again, _ = s2i.make_l1_fullcal(counts, rp, cd, seed=100)
    assert np.array_equal(again, gpu[0]), "same seed must reproduce the same realisation"
    assert np.array_equal(gpu, np.round(gpu)) and gpu.min() >= 0 and gpu.max() <= 65535 + 1000
    assert not np.array_equal(gpu[0], gpu[1])
    mg, mc = gpu.mean(0, dtype=np.float64), cpu.mean(0, dtype=np.float64)
    vg, vc = gpu.var(0, ddof=1, dtype=np.float64), cpu.var(0, ddof=1, dtype=np.float64)
    # z-score of the difference of means, per pixel and resultant; over 8*56*56 = 25k values it must look N(0,1)
    z = (mg - mc) / np.sqrt((vg + vc) / R + 1e-12)
    assert abs(z.mean()) < 0.05, z.mean()
    assert 0.9 < z.std() < 1.1, z.std()
    assert np.abs(z).max() < 6.0
    # variance ratio: log-ratio of two chi^2_{R-1}/(R-1) has std sqrt(4/(R-1)); mean over pixels ~ 0
    lr = np.log((vg + 1e-9) / (vc + 1e-9))
    assert abs(lr.mean()) < 0.02, lr.mean()
    assert abs(lr.std() - np.sqrt(4.0 / (R - 1))) < 0.05, lr.std()
    # analytic check of the first difference in the faint region: Var[R1-R0] ~ read^2 (1/N1 + 1/N0) + shot
    assert vg[0].mean() > 0


@pytest.mark.timeout(120)
def test_apportioning_terminates_in_every_regime():
    """The float32 inversion sampler must hand over to BTRS before P(X = 0) = (1-q)^n leaves the float32 range (a zero
    start value never terminates; seen once at n q = 128): sweep the electrons per pixel over five decades so that every
    read meets n (1-q)^n from 1 down to far below 1e-38, and check the apportioned ramp against its expectation."""
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i

    n, rp = 128, synth.README_PATTERN
    cal = synth.make_caldir(n=n, seed=3, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32)
    na = n - 8
    counts = np.round(np.logspace(0, 5.3, na * na)).astype(np.int32).reshape(na, na)
    out, _ = s2i.make_l1_fullcal(counts, rp, cal, seed=11, add_reset_noise=False, add_read_noise=False, add_biascorr=False)
    assert np.all(np.isfinite(out))
    # the last group's single read holds all electrons: the signal above the first group grows with the counts
    gain = cal["gain"]["roman"]["data"][4:-4, 4:-4]
    lin = (out[-1] - out[0]) * gain
    big = counts > 2000
    ratio = lin[big] / counts[big]
    assert 0.6 < np.median(ratio) < 1.2, np.median(ratio)
    # intermediate groups sit between the first and the last (cumulative counts are monotone)
    assert np.all(out[3][big] <= out[-1][big] + 1) and np.all(out[3][big] >= out[0][big] - 1)
