"""pytest configuration: markers and shared fixtures."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


SMALL_CASES = {
    # tag: (n, read pattern name, p_order, gain dtype, ipc dtype, seed)   -- must match tests/golden/make_golden.py
    "small_p4_f32": (40, "TEST_READ_PATTERN", 3, np.float32, np.float32, 11),
    "small_p11_f32": (40, "README_PATTERN", 10, np.float32, np.float32, 12),
    "small_p4_g64": (40, "TEST_READ_PATTERN", 3, np.float64, np.float32, 13),
    "small_p11_k64": (40, "README_PATTERN", 10, np.float32, np.float64, 14),
    "small_sat_f32": (56, "README_PATTERN", 10, np.float32, np.float32, 15, 12.0),
    "small_sat_g64k64": (56, "TEST_READ_PATTERN", 3, np.float64, np.float64, 16, 12.0),
}


def build_small_case(tag):
    """Regenerate the seeded inputs of one golden case (digest-checked against the golden file)."""
    from romanimpreprocess_b200 import synth

    n, rpname, p_order, gdt, kdt, seed = SMALL_CASES[tag][:6]
    bright = SMALL_CASES[tag][6] if len(SMALL_CASES[tag]) > 6 else 1.0
    rp = getattr(synth, rpname)
    cal = synth.make_caldir(n=n, seed=seed, read_pattern=rp, p_order=p_order, gain_dtype=gdt, ipc_dtype=kdt,
                            sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, meta = synth.make_l1(cal, rp, seed=seed + 1, n_sources=9, cr_frac=0.01, bright=bright)
    return cal, data_u16, amp33_u16, meta, rp


@pytest.fixture(scope="session")
def kats():
    return load_golden("kats")


def synth_sky_image(ny, nx, seed):
    """Same seeded image as tests/golden/make_golden_sim.py:synth_sky_image (medfit goldens)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    img = (0.7 + 0.3 * xx / nx - 0.2 * (yy / ny) ** 2 + 0.05 * rng.randn(ny, nx)).astype(np.float32)
    img[rng.rand(ny, nx) < 0.02] += 30.0
    img[rng.rand(ny, nx) < 0.05] = np.nan
    img[: ny // 8, : nx // 8] = np.nan
    return img


SKY_CASES = {"a": (509, 1022, 2, 8), "b": (300, 257, 0, 8), "c": (412, 412, 3, 4)}
