"""Whole-frame oracle for the BASELINE-size parity tests: ``oracle.l1_to_l2`` on every row band of a 4096^2 frame in
a process pool (fork: the inputs are shared copy-on-write), with the reference-pixel statistics of the WHOLE frame
(they are global: medians over the reference output and the top/bottom reference rows).  Each band carries a 6-row
halo; only its interior is kept (see ``oracle.l1_to_l2``: ``refpix_corr``).  About a minute of CPU per 4096^2 x 8 frame
spread over the host cores."""

import multiprocessing as mp
import os

import numpy as np

from oracle import rip_oracle as orc

_G = {}
HALO = 6


def _cut_cal(c, ya, yb, n):
    """CALDIR ``roman`` branches cut to detector rows [ya, yb) (active-area planes to the matching active rows)."""
    a0, a1 = ya, yb - 8  # active rows of the band in active coordinates (also at the frame edges)
    out = {}
    for key, tree in c.items():
        t = {}
        for k, v in tree.items():
            if isinstance(v, np.ndarray) and v.ndim >= 2 and v.shape[-2] == n and v.shape[-1] in (n, 128):
                t[k] = v[..., ya:yb, :]
            elif isinstance(v, np.ndarray) and v.ndim >= 2 and v.shape[-2] == n - 8 and v.shape[-1] == n - 8:
                t[k] = v[..., a0:a1, :]
            else:
                t[k] = v
        out[key] = t
    return out


def _band(job):
    y0, y1 = job
    c, data, rp, area, cfg, stats, n = _G["c"], _G["data"], _G["rp"], _G["area"], _G["cfg"], _G["stats"], _G["n"]
    ya, yb = max(y0 - HALO, 0), min(y1 + HALO, n)
    rowcorr, cm, cc = stats
    cb = _cut_cal(c, ya, yb, n)
    ar = area[ya:yb] if isinstance(area, np.ndarray) else area
    r = orc.l1_to_l2(np.ascontiguousarray(data[:, ya:yb]), None, cb, rp, 3.04, ar, cfg, do_refpix=False,
                     refpix_corr=(rowcorr[:, ya:yb], cm, cc, ya))  # fmt: skip
    i0, i1 = y0 - ya, y1 - ya
    out = {k: r[k][i0:i1] for k in ("slope", "err_read", "err_poisson", "pdq")}
    out["rdq"] = r["rdq"][:, i0:i1]
    # endslice is in active coordinates of the band: band active row j <-> detector row ya + 4 + j
    e0, e1 = max(y0, 4) - (ya + 4), min(y1, n - 4) - (ya + 4)
    out["endslice"] = r["endslice"][e0:e1]
    return y0, y1, out


def oracle_full_frame(cal, data_u16, amp33_u16, rp, area, cfg, band=256, workers=None):
    """``oracle.l1_to_l2(..., do_refpix=True)`` of the whole frame, band by band.  Returns slope, err_read, err_poisson,
    pdq [n,n], rdq [G,n,n], endslice [n-8,n-8]."""
    from hostcheck import harness

    c = {k: v["roman"] for k, v in cal.items()}
    G, n, _ = data_u16.shape
    stats = harness.refpix_stats(data_u16, amp33_u16, c)
    _G.update(c=c, data=data_u16, rp=rp, area=area, cfg=cfg, stats=stats, n=n)
    jobs = [(y, min(y + band, n)) for y in range(0, n, band)]
    workers = workers or min(len(jobs), os.cpu_count() or 1, 16)
    ref = {"slope": np.empty((n, n), np.float32), "err_read": np.empty((n, n), np.float32),
           "err_poisson": np.empty((n, n), np.float32), "pdq": np.empty((n, n), np.uint32),
           "rdq": np.empty((G, n, n), np.uint8), "endslice": np.empty((n - 8, n - 8), np.int8)}  # fmt: skip
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for y0, y1, out in pool.imap_unordered(_band, jobs):
            for k in ("slope", "err_read", "err_poisson", "pdq"):
                ref[k][y0:y1] = out[k]
            ref["rdq"][:, y0:y1] = out["rdq"]
            ref["endslice"][max(y0, 4) - 4 : min(y1, n - 4) - 4] = out["endslice"]
    _G.clear()
    return ref
