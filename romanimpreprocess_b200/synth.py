"""Synthetic calibration trees and L1 cubes for tests and benchmarks.

The CALDIR recipe restates the reference's test fixture ``gencal``
(reference tests/romanimpreprocess/test_workflow.py:117-332): same analytic maps, same
``np.random.RandomState`` draw order (dark slope, gain, Smax, c2, read noise, reset noise) so that with
``n=4096, seed=1000`` the reference's ``il_example`` golden vectors (test_workflow.py:402-407) hold.
Everything is returned as in-memory trees ``{"roman": {...}}`` with the CALDIR schema of SURVEY App. B.

Extensions that the reference fixture does not have (all opt-in): arbitrary frame size ``n`` (small frames for
parity tests), Legendre order above 3 (the DUMMY calibration files use order 10,
reference runs/summer2025run/write_linearity_config.pl:60), plane dtypes (gain f32/f64, ipc4d f32/f64; SURVEY
App. A0), and a few DQ bits sprinkled into the cal files so that every flag branch is exercised.
"""

import numpy as np

from .dqflags import pixel

FRAME_TIME = 3.04
README_READS = [0, 1, 1, 2, 2, 4, 4, 10, 10, 26, 26, 32, 32, 34, 34, 35]  # reference README.rst:61
LONG16_READS = [0, 1, 1, 2, 2, 4, 4, 6, 6, 8, 8, 10, 10, 14, 14, 18, 18, 22, 22, 26, 26, 28, 28, 30, 30, 32,
                32, 33, 33, 34, 34, 35]  # fmt: skip  (synthetic 16-group table)
TEST_READ_PATTERN = [[0], [1, 2], [3, 4, 5], [6, 7, 8, 9, 10], [11, 12], [13]]  # test_workflow.py:29


def reads_to_pattern(reads):
    """``READS`` list [a0,b0,a1,b1,...] -> read pattern [[a0..b0-1], ...] (reference sim_to_isim.py:970-974)."""
    return [list(range(int(reads[2 * i]), int(reads[2 * i + 1]))) for i in range(len(reads) // 2)]


README_PATTERN = reads_to_pattern(README_READS)
LONG16_PATTERN = reads_to_pattern(LONG16_READS)


def _legendre_and_derivative(z, order):
    """P_L(z), P_L'(z) for L=0..order in float64."""
    p = [np.ones_like(z), z.copy()]
    d = [np.zeros_like(z), np.ones_like(z)]
    for L in range(1, order):
        p.append(((2 * L + 1) * z * p[L] - L * p[L - 1]) / (L + 1))
        d.append(d[L - 1] + (2 * L + 1) * p[L])
    return p[: order + 1], d[: order + 1]


def make_caldir(
    n=4096,
    seed=1000,
    read_pattern=None,
    p_order=3,
    gain_dtype=np.float64,
    ipc_dtype=np.float32,
    sprinkle_flags=False,
    hi_order_scale=1e-2,
    biascorr_amp=0.0,
):
    """Build a complete synthetic CALDIR as a dict of trees.

    With the defaults this is the reference fixture (gain stored as float64, ipc4d float32, order 3, zero
    biascorr).  ``sprinkle_flags`` adds NO_LIN_CORR / REFERENCE_PIXEL / NO_SAT_CHECK / bad flat / bad gain
    pixels (not in the reference fixture; deterministic positions) to exercise flag branches.
    """
    if read_pattern is None:
        read_pattern = TEST_READ_PATTERN
    rng = np.random.RandomState(seed=seed)
    nb = 4
    na = n - 2 * nb
    x, y = np.meshgrid(np.arange(n), np.arange(n))
    ngrp = len(read_pattern)
    t = np.array([FRAME_TIME * np.mean(np.array(g)) for g in read_pattern])

    cal = {}

    # --- biascorr (test_workflow.py:145-155): zeros + t0 ---
    bc = np.zeros((ngrp, na, na), dtype=np.float32)
    if biascorr_amp != 0.0:
        bc += (biascorr_amp * np.sin(0.37 * x[nb:-nb, nb:-nb] + 0.11 * y[nb:-nb, nb:-nb])).astype(np.float32)[
            None
        ] * (1.0 + 0.25 * np.arange(ngrp, dtype=np.float32))[:, None, None]
    cal["biascorr"] = {"roman": {"data": bc, "t0": float(t[1])}}

    # --- dark (test_workflow.py:157-176) ---
    dark_slope = 0.005 * 10.0 ** rng.normal(loc=0.0, scale=1.0, size=(n, n))
    for sl in (np.s_[:, :nb], np.s_[:, -nb:], np.s_[:nb, :], np.s_[-nb:, :]):
        dark_slope[sl] = 0
    bias = 13000 + 200 * np.cos(2.0 * np.pi * x / 256.0) + 100 * np.sin(2.0 * np.pi * y / 256) ** 3
    dark_dq = np.zeros((n, n), dtype=np.uint32)
    cal["dark"] = {
        "roman": {
            "data": np.clip(
                (bias[None, :, :] + dark_slope[None, :, :] * t[:, None, None]).astype(np.float32), 0.0, 65535.0
            ),
            "dq": dark_dq,
            "dark_slope": dark_slope.astype(np.float32),
            "dark_slope_err": np.zeros((n, n), dtype=np.float32),
        }
    }

    # --- gain (test_workflow.py:178-187): NB stored float64 in the fixture ---
    gain = np.clip(1.5 + 0.03 * rng.normal(loc=0.0, scale=1.0, size=(n, n)), 1.4, 1.6).astype(gain_dtype)
    cal["gain"] = {"roman": {"data": gain, "dq": np.zeros((n, n), dtype=np.uint32)}}

    # --- ipc4d (test_workflow.py:189-208) ---
    K = np.zeros((3, 3, na, na), dtype=np.float32)
    K[0, 1] = K[2, 1] = 0.015
    K[1, 0] = K[1, 2] = 0.013
    K[0, 0] = K[2, 2] = K[0, 2] = K[2, 0] = 0.002
    K[0, :, 0, :] = 0.0
    K[:, 0, :, 0] = 0.0
    K[-1, :, -1, :] = 0.0
    K[:, -1, :, -1] = 0.0
    K[1, 1] = 1.0 - np.sum(K, axis=(0, 1))
    cal["ipc4d"] = {"roman": {"data": K.astype(ipc_dtype), "dq": np.zeros((n, n), dtype=np.uint32)}}

    # --- linearitylegendre (test_workflow.py:210-251) ---
    Smin = 5000 + 500 * np.cos((x + 3 * y) / 100.0)
    Smax = 56000 + 10000 * rng.uniform(size=(n, n))
    Smin = np.clip(Smin, 0.5, 65534.5).astype(np.float32)
    Smax = np.clip(Smax, 0.5, 65534.5).astype(np.float32)
    Sref = (Smin + 300 + 100 * (x % 2)).astype(np.float32)
    pflat = (0.95 + 0.1 * (x / n - 1) - 0.2 * (y / n * (1 - y / n))).astype(np.float32)
    pflat[:nb, :] = 0.0
    pflat[-nb:, :] = 0.0
    pflat[:, :nb] = 0.0
    pflat[:, -nb:] = 0.0
    coefs = np.zeros((p_order + 1, n, n), dtype=np.float32)
    coefs[2] = 20 + 180 * rng.uniform(size=(n, n))
    z = 2 * (Sref - Smin) / (Smax - Smin) - 1
    if p_order == 3:
        coefs[1] = (Smax - Smin) / 2.0 - 3 * coefs[2] * z
        coefs[0] = -coefs[1] * z - coefs[2] * (1.5 * z**2 - 0.5)
    else:
        # higher orders: small deterministic-random coefficients, then solve c1, c0 for Phi(Sref)=0, Phi'(Sref)=1
        hrng = np.random.RandomState(seed=seed + 7919)
        for L in range(3, p_order + 1):
            coefs[L] = (hi_order_scale * coefs[2] * hrng.uniform(-1.0, 1.0, size=(n, n)) / (L - 1)).astype(
                np.float32
            )
        z64 = z.astype(np.float64)
        P, dP = _legendre_and_derivative(z64, p_order)
        rest_d = sum(coefs[L].astype(np.float64) * dP[L] for L in range(2, p_order + 1))
        c1 = (Smax.astype(np.float64) - Smin) / 2.0 - rest_d
        rest_v = sum(coefs[L].astype(np.float64) * P[L] for L in range(2, p_order + 1))
        coefs[1] = c1.astype(np.float32)
        coefs[0] = (-c1 * z64 - rest_v).astype(np.float32)
    lin_dq = np.zeros((n, n), dtype=np.uint32)
    cal["linearitylegendre"] = {
        "roman": {"data": coefs, "dq": lin_dq, "Smin": Smin, "Smax": Smax, "Sref": Sref, "pflat": pflat}
    }

    # --- mask (test_workflow.py:267-279) ---
    mask = np.zeros((n, n), dtype=np.uint32)
    mask[:nb, :] |= 2**31
    mask[-nb:, :] |= 2**31
    mask[:, :nb] |= 2**31
    mask[:, -nb:] |= 2**31
    mask |= np.where(dark_slope > 0.25, np.where(dark_slope > 12.5, 2**11, 2**12), 0).astype(np.uint32)
    cal["mask"] = {"roman": {"dq": mask}}

    # --- flat (file type pflat; test_workflow.py:281-285) ---
    cal["flat"] = {"roman": {"data": pflat.copy(), "dq": np.zeros((n, n), np.uint32)}}

    # --- read (test_workflow.py:287-310) ---
    medband = np.zeros((n, 128), dtype=np.float32) + np.float32(29000.0)
    stdband = np.zeros((n, 128), dtype=np.float32) + np.float32(4.0)
    for i in range(max(n // 256, 1)):
        stdband[256 * i, :] = 5
        medband[256 * i, :] += 30
        if 256 * i + 1 < n:
            medband[256 * i + 1, :] += 15
    cal["read"] = {
        "roman": {
            "anc": {"U_PINK": 0.4, "C_PINK": 0.8},
            "data": (6.0 + 5.0 * rng.uniform(size=(n, n))).astype(np.float32),
            "resetnoise": (25.0 + 5.0 * rng.uniform(size=(n, n))).astype(np.float32),
            "amp33": {"valid": True, "med": medband, "std": stdband, "M_PINK": 0.8, "RU_PINK": 1.0},
        }
    }

    # --- saturation (test_workflow.py:312-321) ---
    sat_dq = np.zeros((n, n), np.uint32)
    cal["saturation"] = {"roman": {"data": np.clip(Smax - 50, 1.5, None).astype(np.float32), "dq": sat_dq}}

    if sprinkle_flags:
        srng = np.random.RandomState(seed=seed + 104729)
        yy = srng.randint(nb, n - nb, size=max(n * n // 400, 8))
        xx = srng.randint(nb, n - nb, size=yy.size)
        k = yy.size // 8
        lin_dq[yy[0:k], xx[0:k]] |= pixel.NO_LIN_CORR
        lin_dq[yy[k : 2 * k], xx[k : 2 * k]] |= pixel.REFERENCE_PIXEL  # (pathological but legal)
        sat_dq[yy[2 * k : 3 * k], xx[2 * k : 3 * k]] |= pixel.NO_SAT_CHECK
        cal["flat"]["roman"]["data"][yy[3 * k : 4 * k], xx[3 * k : 4 * k]] = 0.02
        cal["flat"]["roman"]["data"][yy[4 * k : 5 * k], xx[4 * k : 5 * k]] = 12.0
        gain[yy[5 * k : 6 * k], xx[5 * k : 6 * k]] = 0.05
        dark_dq[yy[6 * k : 7 * k], xx[6 * k : 7 * k]] |= pixel.UNRELIABLE_DARK
        lin_dq[yy[7 * k :], xx[7 * k :]] |= pixel.NONLINEAR
        # guide-window pixels in the mask (do_dqinit grows them by one pixel): a block, scattered pixels, frame corners
        grng = np.random.RandomState(seed=seed + 7919)
        mdq = cal["mask"]["roman"]["dq"]
        gy, gx = grng.randint(0, n, size=max(n * n // 2000, 4)), grng.randint(0, n, size=max(n * n // 2000, 4))
        mdq[gy, gx] |= pixel.GW_AFFECTED_DATA
        mdq[n // 3 : n // 3 + 5, n // 2 : n // 2 + 7] |= pixel.GW_AFFECTED_DATA
        mdq[0, 0] |= pixel.GW_AFFECTED_DATA
        mdq[n - 1, n - 1] |= pixel.GW_AFFECTED_DATA

    return cal


def meta_from_pattern(read_pattern, frame_time=FRAME_TIME):
    """N, tbar, tau per group exactly as the reference computes them (gen_cal_image.py:123-140)."""
    ngrp = len(read_pattern)
    meta = {"frame_time": frame_time, "read_pattern": read_pattern, "ngrp": ngrp}
    meta["tbar"] = np.zeros(ngrp, dtype=np.float32)
    meta["tau"] = np.zeros(ngrp, dtype=np.float32)
    meta["N"] = np.zeros(ngrp, dtype=np.int16)
    for i in range(ngrp):
        meta["N"][i] = len(read_pattern[i])
        t0 = read_pattern[i][0]
        meta["tbar"][i] = (t0 + (meta["N"][i] - 1) / 2.0) * frame_time
        meta["tau"][i] = (t0 + (meta["N"][i] - 1) * (2 * meta["N"][i] - 1) / (6.0 * meta["N"][i])) * frame_time
    meta["nborder"] = 4
    return meta


def make_l1(cal, read_pattern, seed=200, n_sources=25, cr_frac=1e-3, sky=0.6, frame_time=FRAME_TIME, bright=1.0):
    """A synthetic (not forward-modelled) L1 cube ``u16[G,n,n]`` + ``amp33 u16[G,n,128]`` for parity/throughput.

    raw = Sref + (sky + dark + sources)*tbar (mildly compressed near full well) + read noise + CR steps, with
    dark-level reference pixels and a reference output around its calibration median.  Bright sources saturate in
    different groups so that every truncated-fit branch of the ramp fitter fires.
    """
    rng = np.random.RandomState(seed=seed)
    lin = cal["linearitylegendre"]["roman"]
    n = lin["Sref"].shape[0]
    nb = 4
    meta = meta_from_pattern(read_pattern, frame_time)
    G = meta["ngrp"]
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    rate = np.full((n, n), sky, dtype=np.float32) + cal["dark"]["roman"]["dark_slope"]
    texp = float(meta["tbar"][-1])
    for j in range(n_sources):
        cx = 10 + (n - 20) * j / float(n_sources)
        cy = 10 + (n - 20) * ((13 * j) % n_sources) / float(n_sources)
        amp = bright * 4000.0 * j / texp  # DN/s at the peak; ``bright`` > 1 makes sources saturate in early groups
        rate += (amp * np.exp(-0.5 * ((xx - cx) ** 2 + (yy - cy) ** 2) / 2.0**2)).astype(np.float32)
    read = cal["read"]["roman"]["data"]
    cube = np.empty((G, n, n), dtype=np.float32)
    crmask = rng.uniform(size=(n, n)) < cr_frac
    crgrp = rng.randint(1, G, size=(n, n))
    cramp = (200.0 + 3000.0 * rng.uniform(size=(n, n))).astype(np.float32)
    full = lin["Smax"] + 500.0
    for g in range(G):
        s = lin["Sref"] + rate * meta["tbar"][g]
        s = s + np.where(crmask & (crgrp <= g), cramp, 0).astype(np.float32)
        s = s - 2.0e-6 * (s - lin["Sref"]) ** 2  # a little compression
        s = np.minimum(s, full)
        s = s + read / np.sqrt(np.float32(meta["N"][g])) * rng.normal(size=(n, n)).astype(np.float32)
        cube[g] = s
    # reference pixels follow the dark cube (+ noise)
    dcube = cal["dark"]["roman"]["data"]
    de = dcube.shape[0] - G
    for g in range(G):
        d = dcube[de + g] + read * rng.normal(size=(n, n)).astype(np.float32) / np.sqrt(np.float32(meta["N"][g]))
        for sl in (np.s_[:nb, :], np.s_[-nb:, :], np.s_[:, :nb], np.s_[:, -nb:]):
            cube[g][sl] = d[sl]
    # common-mode drift per row (what the reference-pixel step removes) shared with the reference output
    a33 = cal["read"]["roman"]["amp33"]
    amp33 = np.empty((G, n, 128), dtype=np.float32)
    for g in range(G):
        drift = (3.0 * np.sin(np.arange(n) / 37.0 + g) + 1.5 * rng.normal(size=n)).astype(np.float32)
        cube[g] += drift[:, None]
        amp33[g] = (
            a33["med"]
            + a33["M_PINK"] * drift[:, None]
            + a33["std"] * rng.normal(size=(n, 128)).astype(np.float32) / np.sqrt(np.float32(meta["N"][g]))
        )
    data_u16 = np.clip(np.round(cube), 0, 65535).astype(np.uint16)
    amp33_u16 = np.clip(np.round(amp33), 0, 65535).astype(np.uint16)
    return data_u16, amp33_u16, meta


def make_area_factor(n, dtype=np.float64):
    """A smooth pixel-area ratio map (stands in for coordutils.pixelarea(...)/Omega_ideal, gen_cal_image.py:618-621)."""
    u = np.linspace(-1.0, 1.0, n)
    uu, vv = np.meshgrid(u, u)
    return (1.0 + 0.012 * uu - 0.007 * vv + 0.004 * uu * vv - 0.003 * uu**2).astype(dtype)
