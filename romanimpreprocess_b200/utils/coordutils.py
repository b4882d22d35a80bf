"""Pixel solid angles from a FITS WCS: drop-in for ``romanimpreprocess.utils.coordutils.pixelarea`` on the inputs the
L1 -> L2 driver feeds it (reference utils/coordutils.py:17-82, called at L1_to_L2/gen_cal_image.py:618-621 with the
``FITSWCS`` header of the configuration, :82-83).

The reference hands the header to astropy / gwcs and differentiates the world coordinates numerically.  None of that
stack exists on the GPU box, and the header is always a zenithal FITS WCS with SIP distortion (the OpenUniverse truth
WCS, from_sim/sim_to_isim.py:986-987; tests/romanimpreprocess/test_workflow.py:62-83), so this module evaluates that
WCS itself -- SIP polynomial, CD matrix, TAN / STG deprojection, spherical rotation with LONPOLE (Calabretta & Greisen
2002; Shupe et al. 2005) -- and then follows the reference's area construction line by line (equal-area reprojection
about the pole of the image's hemisphere, central differences on an (N+2)^2 grid).

Two evaluators share the formulas: NumPy float64 here (``pixelarea``; also the checker of the device version in the
tests) and the CUDA kernel behind ``rip_pixel_area_dev`` (``pixelarea_device``), which writes ``Area / Omega_ideal``
straight into a device plane so that an exposure's AreaFactor costs a 1 kB upload instead of a 67 MB one.
"""

from __future__ import annotations

import re

import numpy as np

MAX_SIP = 9  # SIP order limit of the FITS convention


class FitsWCS:
    """A celestial FITS WCS with optional SIP distortion, parsed from header cards.

    Supported: ``CTYPE`` RA---TAN / DEC--TAN, RA---STG / DEC--STG, each with or without ``-SIP``; ``CD`` matrix or
    ``CDELT`` (+ ``PC``); ``LONPOLE``.  Anything else raises ``ValueError("Unrecognized WCS type")`` like the reference.
    """

    def __init__(self, header):
        h = header if isinstance(header, dict) else parse_header(header)
        ct1, ct2 = str(h.get("CTYPE1", "")).strip(), str(h.get("CTYPE2", "")).strip()
        m1 = re.fullmatch(r"RA---(TAN|STG)(-SIP)?", ct1)
        m2 = re.fullmatch(r"DEC--(TAN|STG)(-SIP)?", ct2)
        if not m1 or not m2 or m1.group(1) != m2.group(1):
            raise ValueError("Unrecognized WCS type")
        self.proj = m1.group(1)
        self.crpix = (float(h["CRPIX1"]), float(h["CRPIX2"]))
        self.crval = (float(h["CRVAL1"]), float(h["CRVAL2"]))
        if "CD1_1" in h:
            self.cd = np.array([[float(h.get("CD1_1", 0.0)), float(h.get("CD1_2", 0.0))],
                                [float(h.get("CD2_1", 0.0)), float(h.get("CD2_2", 0.0))]])  # fmt: skip
        else:
            pc = np.array([[float(h.get("PC1_1", 1.0)), float(h.get("PC1_2", 0.0))],
                           [float(h.get("PC2_1", 0.0)), float(h.get("PC2_2", 1.0))]])  # fmt: skip
            self.cd = np.diag([float(h.get("CDELT1", 1.0)), float(h.get("CDELT2", 1.0))]) @ pc
        # native longitude of the celestial pole: zenithal projections have theta0 = 90 deg
        self.lonpole = float(h["LONPOLE"]) if "LONPOLE" in h else (0.0 if self.crval[1] >= 90.0 else 180.0)
        self.a = np.zeros((MAX_SIP + 1, MAX_SIP + 1))
        self.b = np.zeros((MAX_SIP + 1, MAX_SIP + 1))
        self.a_order = int(h.get("A_ORDER", 0)) if m1.group(2) else 0
        self.b_order = int(h.get("B_ORDER", 0)) if m2.group(2) else 0
        if max(self.a_order, self.b_order) > MAX_SIP:
            raise ValueError("Unrecognized WCS type")
        for key, val in h.items():
            m = re.fullmatch(r"([AB])_(\d)_(\d)", key)
            if m and m1.group(2):
                p, q = int(m.group(2)), int(m.group(3))
                (self.a if m.group(1) == "A" else self.b)[p, q] = float(val)

    def pack(self):
        """The WCS as the flat float64 vector the device kernel takes (``rip_pixel_area_dev``): crpix(2) crval(2) cd(4)
        lonpole proj(0 TAN, 1 STG) order, then A[p][q] and B[p][q] as (MAX_SIP+1)^2 blocks."""
        order = max(self.a_order, self.b_order)
        head = [*self.crpix, *self.crval, *self.cd.ravel(), self.lonpole, 0.0 if self.proj == "TAN" else 1.0, float(order)]
        return np.concatenate([np.array(head), self.a.ravel(), self.b.ravel()]).astype(np.float64)

    def pix2world(self, x, y):
        """0-based pixel coordinates -> (ra, dec) in degrees, float64."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        u = x + 1.0 - self.crpix[0]
        v = y + 1.0 - self.crpix[1]
        order = max(self.a_order, self.b_order)
        f = np.zeros_like(u)
        g = np.zeros_like(u)
        if order > 0:
            # Horner in v inside Horner in u, highest powers first (the device kernel uses the same nesting)
            for p in range(order, -1, -1):
                ca = np.zeros_like(u)
                cb = np.zeros_like(u)
                for q in range(order - p, -1, -1):
                    ca = ca * v + self.a[p, q]
                    cb = cb * v + self.b[p, q]
                f = f * u + ca
                g = g * u + cb
        uu, vv = u + f, v + g
        xi = self.cd[0, 0] * uu + self.cd[0, 1] * vv
        eta = self.cd[1, 0] * uu + self.cd[1, 1] * vv
        deg = np.pi / 180.0
        r = np.hypot(xi, eta) * deg  # radians
        phi = np.arctan2(xi, -eta)
        if self.proj == "TAN":
            theta = np.arctan2(1.0, r)  # R = cot(theta)
        else:
            theta = np.pi / 2.0 - 2.0 * np.arctan(r / 2.0)  # R = 2 tan((90 - theta)/2)
        dp = self.crval[1] * deg
        dphi = phi - self.lonpole * deg
        st, ct = np.sin(theta), np.cos(theta)
        sdp, cdp = np.sin(dp), np.cos(dp)
        cph, sph = np.cos(dphi), np.sin(dphi)
        dec = np.arcsin(np.clip(st * sdp + ct * cdp * cph, -1.0, 1.0))
        ra = self.crval[0] * deg + np.arctan2(-ct * sph, st * cdp - ct * sdp * cph)
        return ra / deg, dec / deg


def parse_header(text):
    """FITS header cards -> dict.  Accepts the 80-column card stream ``astropy.io.fits.Header.tofile`` writes (no line
    breaks; what ``FITSWCS`` files hold, reference from_sim/sim_to_isim.py:987) as well as one card per line."""
    if isinstance(text, bytes):
        text = text.decode("ascii", errors="replace")
    cards = text.splitlines() if "\n" in text.strip() else [text[i : i + 80] for i in range(0, len(text), 80)]
    out = {}
    for card in cards:
        key = card[:8].strip()
        if not key or key in ("COMMENT", "HISTORY", "END") or len(card) < 10 or card[8] != "=":
            if key == "END":
                break
            continue
        body = card[10:]
        if body.lstrip().startswith("'"):
            m = re.match(r"\s*'((?:[^']|'')*)'", body)
            out[key] = m.group(1).replace("''", "'").rstrip() if m else body.strip()
            continue
        val = body.split("/", 1)[0].strip()
        if val in ("T", "F"):
            out[key] = val == "T"
            continue
        try:
            out[key] = int(val)
        except ValueError:
            try:
                out[key] = float(val.replace("D", "E").replace("d", "e"))
            except ValueError:
                out[key] = val
    return out


def wcs_from_config(config):
    """``wcs_from_config`` of the driver (reference L1_to_L2/gen_cal_image.py:64-87): the ``FITSWCS`` header, or None."""
    if "FITSWCS" in config:
        with open(config["FITSWCS"]) as f:
            return FitsWCS(f.read())
    return None


def pixelarea(inwcs, N=4088):
    """(N, N) array of pixel solid angles in steradians (reference utils/coordutils.py:17-82, same construction).

    ``inwcs``: a :class:`FitsWCS`, header text or a header dict.  Other objects raise
    ``ValueError("Unrecognized WCS type")`` as in the reference.
    """
    if isinstance(inwcs, (str, bytes, dict)):
        try:
            inwcs = FitsWCS(inwcs)
        except KeyError as e:
            raise ValueError("Unrecognized WCS type") from e
    if not isinstance(inwcs, FitsWCS):
        raise ValueError("Unrecognized WCS type")
    sp = np.linspace(-1, N, N + 2)
    xx, yy = np.meshgrid(sp, sp)
    deg = np.pi / 180.0
    ra, dec = inwcs.pix2world(xx.ravel(), yy.ravel())
    ra = ra * deg
    dec = dec * deg
    theta = np.pi / 2.0 + dec
    if dec[0] > 0:
        theta = np.pi / 2.0 - dec
    rho = 2.0 * np.sin(theta / 2.0)
    u = (rho * np.cos(ra)).reshape((N + 2, N + 2))
    v = (rho * np.sin(ra)).reshape((N + 2, N + 2))
    del rho
    J11 = (u[1:-1, 2:] - u[1:-1, :-2]) / 2.0
    J12 = (u[2:, 1:-1] - u[:-2, 1:-1]) / 2.0
    J21 = (v[1:-1, 2:] - v[1:-1, :-2]) / 2.0
    J22 = (v[2:, 1:-1] - v[:-2, 1:-1]) / 2.0
    return np.abs(J11 * J22 - J21 * J12)


def pixelarea_device(inwcs, N=4088, inv_omega=1.0, dtype=np.float64, device=0):
    """``pixelarea`` computed by the CUDA kernel (``rip_pixel_area_host``): ``Area * inv_omega`` as a host array.  The
    pipelined driver never brings the plane back: see ``gen_cal_image.Pipeline.set_area_wcs``."""
    from .. import _lib  # noqa: PLC0415

    if not isinstance(inwcs, FitsWCS):
        inwcs = FitsWCS(inwcs)
    w = inwcs.pack()
    out = np.empty((N, N), dtype=np.dtype(dtype))
    _lib.check(_lib.lib().rip_pixel_area_host(device, _lib.ptr(w), int(w.size), int(N), float(inv_omega), _lib.ptr(out),
                                              _lib.float_tag(out)))  # fmt: skip
    return out
