"""Shared comparison helpers for the parity tests (CUDA path or host check vs the oracle)."""

import numpy as np

# north_star: DQ / integer outputs bit-exact; slopes, variances, linearised cubes within fp32 relative tolerance.
# The kernels reproduce the reference's op order and roundings, so the observed difference is 0; the stated bound is
# what the tests enforce.
RTOL = 1.0e-5
ATOL = 1.0e-6


def assert_float_close(a, b, name, rtol=RTOL, atol=ATOL):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{name}: NaN pattern differs at {np.count_nonzero(nan_a != nan_b)} pixels"
    inf = np.isinf(a) | np.isinf(b)
    assert np.array_equal(a[inf], b[inf]), f"{name}: infinities differ"
    ok = ~(nan_a | inf)
    d = np.abs(a[ok].astype(np.float64) - b[ok].astype(np.float64))
    lim = atol + rtol * np.abs(b[ok].astype(np.float64))
    bad = d > lim
    assert not np.any(bad), (
        f"{name}: {np.count_nonzero(bad)} of {d.size} values outside rtol={rtol} atol={atol}; "
        f"max abs diff {d.max():.3g}, max rel {np.max(d / (np.abs(b[ok]) + 1e-30)):.3g}"
    )
    return int(np.count_nonzero(d != 0))


def assert_bits_equal(a, b, name):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        idx = np.argwhere(a != b)
        ex = ", ".join(f"{tuple(i)}: {int(a[tuple(i)]):#x} vs {int(b[tuple(i)]):#x}" for i in idx[:5])
        raise AssertionError(f"{name}: {len(idx)} mismatches, e.g. {ex}")


def compare_l2(out, ref, lin_key="lin_cube", check_rdq=True, check_lin=True):
    """``out`` from the CUDA path / host check, ``ref`` from oracle.l1_to_l2(return_intermediates=True)."""
    stats = {}
    for k in ("slope", "err_read", "err_poisson"):
        stats[k] = assert_float_close(out[k], ref[k], k)
    if check_lin and lin_key in out:
        stats["lin_cube"] = assert_float_close(out[lin_key], ref["ipc"], "lin_cube")
    assert_bits_equal(out["pdq"], ref["pdq"], "pdq")
    if "endslice" in out:
        assert_bits_equal(out["endslice"], ref["endslice"], "endslice")
    if check_rdq and "rdq" in out:
        assert_bits_equal(out["rdq"], ref["rdq"], "rdq")
    return stats


def band_check(out, cal, data_u16, amp33_u16, rp, area, flat_ipc, dslope_ipc, nrows=64):
    """Full-frame result ``out`` vs the oracle run on a row band [y0-6, y1+6) as a stand-alone frame, with the
    full-frame reference-pixel statistics supplied (they are global).  Lets the 4096^2 case be checked in seconds."""
    from hostcheck import harness

    from oracle import rip_oracle as orc

    c = {k: v["roman"] for k, v in cal.items()}
    rowcorr, cm, cc = harness.refpix_stats(data_u16, amp33_u16, c)
    G, n, _ = data_u16.shape
    sat_rows = np.argwhere(out["pdq"][4:-4, 4:-4] & orc.SATURATED)[:, 0] + 4
    src_y = int(np.median(sat_rows)) if len(sat_rows) else n // 2
    y0 = min(max(src_y - nrows // 2, 16), n - 16 - nrows)
    y1 = y0 + nrows
    h = 6
    rows = slice(y0 - h, y1 + h)
    data = data_u16[:, rows].astype(np.float32)
    rdq = np.zeros(data.shape, np.uint8)
    rdq[0] |= 1
    pdq = c["mask"]["dq"][rows].copy()
    orc.flag_saturation(data, rdq, pdq, c["saturation"]["data"][rows], c["saturation"]["dq"][rows], backup=1)
    dark = c["dark"]["data"][:, rows]
    jj = np.arange(y0 - h, y1 + h, dtype=np.float64)
    nch = n // 128
    for g in range(G):
        v = data[g] - dark[g]
        v = (v - rowcorr[g, rows][:, None]).astype(np.float32)
        line = cm[g, :nch][:, None] * jj[None, :] + cc[g, :nch][:, None]  # [nch, rows]
        v = (v - np.repeat(line.T, 128, axis=1)).astype(np.float32)
        data[g] = v + dark[g]
    bc = c["biascorr"]["data"]
    data[:, :, 4:-4] -= bc[bc.shape[0] - G :, y0 - h - 4 : y1 + h - 4, :]
    lin = {k: (v[..., rows, :] if hasattr(v, "shape") else v) for k, v in c["linearitylegendre"].items()}
    phi, dq_lin = orc.multilin(data, lin, do_not_flag_first=(list(rp[0]) == [0]), attempt_corr=~rdq & orc.SATURATED)
    K = c["ipc4d"]["data"][:, :, y0 - h - 4 : y1 + h - 4, :]
    gain = c["gain"]["data"][rows]
    cube = phi.copy()
    for g in range(G):
        cube[g, :, 4:-4] = orc.ipc_rev(cube[g, :, 4:-4] * gain[:, 4:-4], K) / gain[:, 4:-4]
    inner = slice(h, h + nrows)
    assert_float_close(out["lin_cube"][:, y0:y1, 4:-4], cube[:, inner, 4:-4], "band lin_cube")
    meta = orc.make_meta(rp, 3.04)
    meta["K"] = orc.construct_weights(0.4 / 1.8 / 7.0**2, meta, exclude_first=True)
    pdq_b = pdq | dq_lin
    # the oracle's ramp_fit treats the outer 4 rows of what it is given as border; the halo (6) covers that
    s, er, ep = orc.ramp_fit(cube, rdq, pdq_b, meta, gain, c["read"]["data"][rows], True)
    assert_bits_equal(out["rdq"][:, y0:y1], rdq[:, inner], "band rdq")
    flat = (flat_ipc[y0:y1, 4:-4] / area[y0:y1, 4:-4]).astype(np.float32)
    ds = dslope_ipc[y0:y1, 4:-4]
    err = np.hypot(er, ep)[inner, 4:-4]
    epo = np.sqrt((ep**2)[inner, 4:-4])
    ero = np.sqrt(np.clip(err**2 - epo**2, 0.0, None))
    assert_float_close(out["slope"][y0:y1, 4:-4], (s[inner, 4:-4] - ds) / flat, "band slope")
    assert_float_close(out["err_read"][y0:y1, 4:-4], ero / flat, "band err_read")
    assert_float_close(out["err_poisson"][y0:y1, 4:-4], epo / flat, "band err_poisson")
    assert np.count_nonzero(rdq[:, inner] & 2) > 0, "band holds no saturated pixel"
    return y0, y1
