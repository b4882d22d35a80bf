"""CPU: the fused kernel's per-pixel source (csrc/rip_cal_core.cuh, csrc/rip_math.cuh), compiled for the host and
walked through the kernel's tile / march schedule (tests/hostcheck), against the oracle.

This is a logic check of the shared __host__ __device__ source in the GPU-less container (ring depths, halos, tile
seams, DQ propagation, every truncation branch); the product parity gate is tests/test_gpu_fused.py.
"""

import numpy as np
import pytest
from conftest import SMALL_CASES, build_small_case
from hostcheck import harness
from parity import compare_l2

from oracle import rip_oracle as orc
from romanimpreprocess_b200 import synth

CFG7 = {"RAMP_OPT_PARS": {"slope": 0.4, "gain": 1.8, "sigma_read": 7.0}}


@pytest.mark.parametrize("tag", list(SMALL_CASES))
def test_small_cases(tag):
    cal, data_u16, amp33_u16, meta, rp = build_small_case(tag)
    n = data_u16.shape[1]
    area = synth.make_area_factor(n)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data_u16, amp33_u16, c, rp, 3.04, area, CFG7, do_refpix=False, return_intermediates=True)
    for threads, band in ((32, 16), (64, 7), (128, 40)):
        out = harness.run_fused(cal, data_u16, amp33_u16, rp, 3.04, area, CFG7, do_refpix=False, threads=threads,
                                band_rows=band)  # fmt: skip
        stats = compare_l2(out, ref, lin_key="ipc")
        assert all(v == 0 for v in stats.values()), stats  # in fact bit-identical


MEDIUM = [
    # n, pattern, order, gain dtype, ipc dtype, seed, config, threads, band, bright
    (256, "README_PATTERN", 10, np.float32, np.float32, 21, {}, 64, 32, 1.0),
    (256, "LONG16_PATTERN", 10, np.float32, np.float32, 22, {"EXCLUDE_FIRST": False, "SATURATION_BACKUP": 2}, 128, 100, 4.0),
    (384, "TEST_READ_PATTERN", 3, np.float64, np.float64, 23,
     {"JUMP_DETECT_PARS": {"SthreshA": 4.0, "SthreshB": 3.5, "IthreshA": 0.6, "IthreshB": 600.0}}, 96, 7, 2.0),
    (256, "README_PATTERN", 10, np.float32, np.float64, 24, {"SATURATION_BACKUP": 0}, 256, 256, 8.0),
]  # fmt: skip


@pytest.mark.parametrize("case", MEDIUM, ids=[f"n{c[0]}_{c[1]}_s{c[5]}" for c in MEDIUM])
def test_medium_cases_with_refpix(case):
    n, rpname, po, gdt, kdt, seed, cfg, threads, band, bright = case
    rp = getattr(synth, rpname)
    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=po, gain_dtype=gdt, ipc_dtype=kdt,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=seed + 1, n_sources=25, cr_frac=0.01, bright=bright)
    area = synth.make_area_factor(n)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data_u16, amp33_u16, c, rp, 3.04, area, cfg, do_refpix=True, return_intermediates=True)
    out = harness.run_fused(cal, data_u16, amp33_u16, rp, 3.04, area, cfg, do_refpix=True, threads=threads,
                            band_rows=band)  # fmt: skip
    stats = compare_l2(out, ref, lin_key="ipc")
    assert all(v == 0 for v in stats.values()), stats
    # the case must exercise what it claims to
    assert np.count_nonzero(ref["pdq"] & orc.SATURATED) > 50
    assert np.count_nonzero(ref["pdq"] & orc.JUMP_DET) > 50
    assert np.count_nonzero(ref["pdq"] & orc.NO_LIN_CORR) > 0
    assert len(np.unique(ref["endslice"])) >= 4


def test_band_check_helper_on_host():
    """The row-band comparison used for the 4096^2 GPU case (tests/parity.py:band_check), exercised here at n=512."""
    from parity import band_check

    rp = synth.README_PATTERN
    n = 512
    cal = synth.make_caldir(n=n, seed=31, read_pattern=rp, p_order=10, gain_dtype=np.float32, ipc_dtype=np.float32,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=32, n_sources=25, cr_frac=1e-2, bright=3.0)
    area = synth.make_area_factor(n, np.float64)
    out = harness.run_fused(cal, data_u16, amp33_u16, rp, 3.04, area, CFG7, do_refpix=True, threads=128, band_rows=128)
    out["lin_cube"] = out["ipc"]
    c = {k: v["roman"] for k, v in cal.items()}
    _, _, _, dslope, flat = harness.static_products(c)
    y0, y1 = band_check(out, cal, data_u16, amp33_u16, rp, area, flat, dslope)
    assert y1 - y0 == 64


V2_CASES = [
    # n, pattern, order, seed, config, band_rows, bright, refpix[, ipc4d dtype]
    (40, "README_PATTERN", 10, 12, {}, 16, 1.0, False),
    (56, "README_PATTERN", 10, 16, {}, 7, 12.0, False, np.float64),   # float64 ipc4d: the K64 form of the kernel
    (256, "README_PATTERN", 10, 28, {"SATURATION_BACKUP": 0}, 100, 8.0, True, np.float64),
    (128, "README_PATTERN", 3, 29, {"EXCLUDE_FIRST": False}, 50, 4.0, True, np.float64),
    (56, "README_PATTERN", 10, 15, {}, 7, 12.0, False),
    (64, "README_PATTERN", 6, 31, {}, 20, 6.0, False),      # P = 7: runs the P = 11 kernel on zero-padded records
    (64, "LONG16_PATTERN", 2, 32, {}, 20, 6.0, False),      # P = 3 -> 4
    (256, "README_PATTERN", 10, 21, {}, 32, 1.0, True),
    (256, "LONG16_PATTERN", 10, 22, {"EXCLUDE_FIRST": False, "SATURATION_BACKUP": 2}, 100, 4.0, True),
    (384, "README_PATTERN", 3, 26, {"SATURATION_BACKUP": 0, "JUMP_DETECT_PARS": {"SthreshA": 4.0, "SthreshB": 3.5}}, 128, 8.0, True),
    (128, "LONG16_PATTERN", 3, 27, {}, 128, 6.0, True),
]  # fmt: skip


@pytest.mark.parametrize("case", V2_CASES, ids=[f"n{c[0]}_{c[1]}_s{c[3]}" + ("_k64" if len(c) > 8 else "") for c in V2_CASES])
def test_v2_kernel_source(case):
    """The throughput kernel (csrc/rip_v2_core.cuh: packed records, float4 rings, packed-pair arithmetic, shared
    reciprocal division, squared jump test) must reproduce the oracle bit for bit, like v1."""
    n, rpname, po, seed, cfg, band, bright, refpix = case[:8]
    kdt = case[8] if len(case) > 8 else np.float32
    rp = getattr(synth, rpname)
    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=po, gain_dtype=np.float32, ipc_dtype=kdt,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=seed + 1, n_sources=25 if n > 100 else 9, cr_frac=0.01,
                                           bright=bright)  # fmt: skip
    area = synth.make_area_factor(n)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data_u16, amp33_u16, c, rp, 3.04, area, cfg, do_refpix=refpix, return_intermediates=True)
    out = harness.run_fused(cal, data_u16, amp33_u16, rp, 3.04, area, cfg, do_refpix=refpix, band_rows=band, v2=True)
    stats = compare_l2(out, ref, lin_key="ipc")
    assert all(v == 0 for v in stats.values()), stats


V6_CASES = [c for c in V2_CASES if c[1] == "README_PATTERN"]  # (G = 8; float32 and float64 ipc4d)


@pytest.mark.parametrize("case", V6_CASES, ids=[f"n{c[0]}_{c[1]}_s{c[3]}" + ("_k64" if len(c) > 8 else "") for c in V6_CASES])
def test_v6_kernel_source(case):
    """The five-CTAs-per-SM schedule of the throughput kernel (rip_v2_core.cuh "v6": depth-4 record ring, stage c one row
    behind stage b behind a second barrier, records loaded from L2 just in time) walks to the same result."""
    n, rpname, po, seed, cfg, band, bright, refpix = case[:8]
    kdt = case[8] if len(case) > 8 else np.float32
    rp = getattr(synth, rpname)
    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=po, gain_dtype=np.float32, ipc_dtype=kdt,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, _ = synth.make_l1(cal, rp, seed=seed + 1, n_sources=25 if n > 100 else 9, cr_frac=0.01,
                                           bright=bright)  # fmt: skip
    area = synth.make_area_factor(n)
    c = {k: v["roman"] for k, v in cal.items()}
    ref = orc.l1_to_l2(data_u16, amp33_u16, c, rp, 3.04, area, cfg, do_refpix=refpix, return_intermediates=True)
    out = harness.run_fused(cal, data_u16, amp33_u16, rp, 3.04, area, cfg, do_refpix=refpix, band_rows=band, v6=True)
    stats = compare_l2(out, ref, lin_key="ipc")
    assert all(v == 0 for v in stats.values()), stats


def test_shared_reciprocal_division_is_ieee_division():
    """rip::v2::SharedDiv (one refined reciprocal + two FMA-residual corrections per numerator) == x / d bit for bit,
    with the hardware reciprocal emulated as a +-1 ulp perturbed 1/d: 2 x 10^7 random pairs in the ranges the kernel
    sees (Smax-Smin with DN numerators; gain with e- numerators)."""
    h = harness.lib()
    assert h.hostcheck_shared_div(10000000, 7, 1e-3, 1e6, 3e5) == 0
    assert h.hostcheck_shared_div(10000000, 9, 0.5, 3.0, 1e7) == 0


@pytest.mark.parametrize("P,hostile", [(11, 0.0), (4, 0.0), (11, 6.0), (16, 0.3)])
def test_fast_inverse_equals_24_step_search(P, hostile):
    """csrc/rip_math.cuh invlin_fast_z (the forward model's certified shortcut) returns the z of the reference's
    24-step search (utils/ipc_linearity.py:381-387) for every signal: ramps, adversarial values sitting exactly on
    the float32 numbers the search compares with, range edges, NaN/inf.  `hostile` inflates the high-order
    coefficients until many pixels lose the monotonicity certificate (those must take the plain search)."""
    import ctypes as C

    h = harness.lib()
    h.hostcheck_invlin_fast.restype = C.c_long
    h.hostcheck_invlin_fast.argtypes = [C.c_long, C.c_uint, C.c_int, C.c_double, C.POINTER(C.c_double)]
    stats = (C.c_double * 3)()
    bad = h.hostcheck_invlin_fast(4000, 1234 + P, P, hostile, stats)
    calls, exact, uncert = stats[0], stats[1], stats[2]
    assert bad == 0, f"{bad} of {calls:.0f} inversions differ"
    assert calls > 1000
    if hostile > 0.0:
        assert uncert > 0
    if hostile == 0.0:
        assert uncert == 0
        assert exact / calls < 8.0, exact / calls  # the point of it: far fewer than 24 evaluations
    print(f"P={P} hostile={hostile}: {calls:.0f} calls, {exact / max(calls, 1):.2f} exact evaluations per call, "
          f"{uncert:.0f} uncertified pixels")
