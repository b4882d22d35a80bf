"""Noise realisations ("noise layers") on the GPU: the read-noise directive of
``romanimpreprocess.L1_to_L2.gen_noise_image.make_noise_cube`` (reference L1_to_L2/gen_noise_image.py:60-331).

For a layer command such as ``"Rz4S2C1"`` the reference (i) takes the L1 cube (flag ``a``) or the dark cube cast to the
cube's integer type, (ii) adds white read noise per group and the correlated part (reference pixels, 1/f banding,
reference output: ``fill_in_refdata_and_1f``), (iii) runs the whole L1->L2 calibration on it through temporary ASDF files,
(iv) differences the result with the L2 image of the noiseless cube, (v) optionally clips at ``z`` Gaussian-equivalent
sigmas of the inter-quartile range and (vi) removes the sky modes (``S<order>``: ``sky.medfit``).  Production runs do this
8 times per exposure (runs/summer2025run/OpenUniverse_to_L1L2.py:124-133), i.e. 9+ calibrations per SCA.

Here every step stays in HBM and shares one resident CALDIR:

    rip_dark_as_l1_dev -> rip_add_read_noise_dev -> rip_fill_refdata_1f_dev -> rip_l1_to_l2_dev
    [-> medfit (SKYORDER)] -> rip_active_diff_dev [-> z clip: rip_order_stats_dev + rip_clip_dev] [-> medfit (S)]

Supported directives: ``R`` (flags ``a``, ``z<number>``), ``P`` (flags ``b<order>``, ``r``: re-sampled Poisson noise of
the sky level propagated through the ramp-fit weights of each pixel's ramp end, ``rip_poisson_resample_dev``),
``O`` (pseudo-Poisson draws from the Pearson family with the moments of the ramp-fitted Poisson noise, reference
:173-227 -> ``GalPoisson``; ``rip_pearson_noise_dev``), ``S<order>``, ``C<tag>`` (a label: ignored, as in the reference)
-- i.e. both production layer families ``Rz4PbrS2C*`` and ``Rz4OS2C*`` (runs/summer2025run/OpenUniverse_to_L1L2.py:124-133).
``O`` generates every Pearson type the reference dispatches: Type I (Beta; the type every Roman read pattern tried here
selects, their nu_41 is negative), Type VI (beta prime), Type IV (Devroye's rejection sampler in the angle, exact
normalisation: one sampler where the reference needs two) and, exactly on their lines, Types III / V; pixels with invalid
Pearson parameters raise ``ValueError`` as in the reference.
Random numbers are Philox (the reference: GalSim / NumPy): layers are validated statistically.

``generate_all_noise(config)`` is the reference's driver (:334-391): layers of ``config["NOISE"]["LAYER"]`` for the exposure
``config["IN"]``, written to ``config["NOISE"]["OUT"]`` (ASDF tree ``{"config", "noise"}``).
"""

import re

import numpy as np

from .. import _lib
from ..utils import sky
from . import gen_cal_image as gci


def _get_subscript(arr, ch):
    """Text after the last ``ch`` in ``arr`` up to (not including) the next capital letter:
    ``_get_subscript('RS2Pg4', 'S') -> '2'`` (reference gen_noise_image.py:33-57)."""
    return re.split(r"(?=[A-Z])", arr.split(ch)[-1])[0]


def tilde_nus(read_pattern, weights):
    """(nu~21, nu~31, nu~41) per frame of the weighted MultiAccum combination ``sum_k weights[k] * resultant_k`` of a unit
    Poisson rate (GalPoisson/find_tilnus.py:46-78): with L[k, r] = 1/N_k on the reads of group k, T = the reversed
    cumulative sum of L along the reads (the weight of each read's increment in each resultant) and w = weights . T[:, 1:],
    nu_p1 = sum(w^p), then nu~21 = nu21, nu~31 = nu31 - 3 nu21^2, nu~41 = nu41 - 10 nu21 nu31 - nu21 * 3 nu21^2 + 18 nu21^3."""
    nread = max(g[0] + len(g) for g in read_pattern)
    L = np.zeros((len(read_pattern), nread))
    for k, g in enumerate(read_pattern):
        L[k, g[0] : g[0] + len(g)] = 1.0 / len(g)
    T = np.cumsum(L[:, ::-1], axis=1)[:, ::-1]
    w = np.dot(np.asarray(weights), T[:, 1:])
    nu21, nu31, nu41 = np.sum(w**2), np.sum(w**3), np.sum(w**4)
    nu42 = 3 * nu21**2
    return nu21, nu31 - 3 * nu21**2, nu41 - 10 * nu21 * nu31 - nu21 * nu42 + 18 * nu21**3


def _ptr(t):
    import ctypes as C

    return C.c_void_p(t.data_ptr())


class NoiseLayers:
    """Device-resident noise-layer generator for the exposures of one SCA (one process / GPU)."""

    def __init__(self, caldir, read_pattern, frame_time, config=None, device=0):
        import torch  # noqa: PLC0415  (device memory only)

        self.torch = torch
        self.device = device
        self.dev = torch.device("cuda", device)
        self.cal = caldir if isinstance(caldir, gci.CalDir) else gci.CalDir(caldir, device)
        self.owns_cal = self.cal is not caldir
        self.config = dict(config or {})
        self.read_pattern = read_pattern
        cal = self.cal
        n, G = cal.n, len(read_pattern)
        self.n, self.na, self.G = n, cal.na, G
        self.refpix = bool(cal.has_amp33 and n // 32 == 128)  # the reference output is 128 columns wide
        self.dplan = gci.DevicePlan(cal, read_pattern, frame_time, self.config, do_refpix=self.refpix, area_dtype=np.float32)
        self.rpg = np.ascontiguousarray([len(g) for g in read_pattern], dtype=np.int32)
        z = dict(device=self.dev)
        self.d_work = torch.empty((G, n, n), dtype=torch.uint16, **z)
        self.d_amp33 = torch.zeros((G, n, 128 if self.refpix else max(n // 32, 1)), dtype=torch.uint16, **z)
        self.l2 = {k: [torch.empty((n, n), dtype=torch.float32, **z) for _ in range(3)] for k in ("orig", "ref", "noisy")}
        self.d_pdq = torch.empty((n, n), dtype=torch.int32, **z)
        self.d_diff = torch.empty((self.na, self.na), dtype=torch.float32, **z)
        self.d_withsky = torch.empty((self.na, self.na), dtype=torch.float32, **z)  # data_withsky of the exposure's L2
        self.d_skylevel = torch.empty((self.na, self.na), dtype=torch.float32, **z)
        self.d_endslice = torch.empty((self.na, self.na), dtype=torch.int8, **z)
        self.frame_time = float(frame_time)
        self.have = set()

    def _stream(self):
        import ctypes as C

        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def _calibrate(self, d_cube, d_amp33, d_area, which):
        """L1->L2 of a device cube into slot ``which``; the L2 ``data`` plane is sky-subtracted when SKYORDER is set
        (gen_cal_image.py:643-647), exactly what the reference differences."""
        st = self._stream()
        s, er, ep = self.l2[which]
        gci.calibrate_device(self.cal, self.dplan, d_cube.data_ptr(), d_amp33.data_ptr() if self.refpix else 0,
                             d_area.data_ptr() if d_area is not None else 0, s.data_ptr(), er.data_ptr(), ep.data_ptr(),
                             self.d_pdq.data_ptr(), d_endslice=self.d_endslice.data_ptr() if which == "orig" else 0,
                             stream=st.value or 0)  # fmt: skip
        if which == "orig":  # slope_withsky (gen_cal_image.py:640), active window
            self.d_withsky.copy_(s[self.cal.nb : self.n - self.cal.nb, self.cal.nb : self.n - self.cal.nb])
        if "SKYORDER" in self.config:
            nb, n = self.cal.nb, self.n
            sky.medfit_device(s[nb : n - nb, nb : n - nb].data_ptr(), n, self.na, self.na, order=int(self.config["SKYORDER"]),
                              device=self.device, stream=st.value or 0, subtract=True)  # fmt: skip
        self.have.add(which)

    def set_exposure(self, d_data, d_amp33, d_area=None):
        """The exposure the layers belong to: device tensors u16 [G,n,n], u16 [G,n,128], f32 [n,n] (or None)."""
        self.d_data, self.d_amp33_in, self.d_area = d_data, d_amp33, d_area
        self.have.discard("orig")
        self.have.discard("ref")

    def layer(self, cmd, seed, out=None):
        """One noise layer [na,na] float32 (host) for the directive string ``cmd``.  ``out``: a float32 [na,na] host array
        to receive it (page-locked memory from ``_lib.pinned_empty`` makes the copy run at PCIe speed)."""
        lib, cal, st = _lib.lib(), self.cal, self._stream()
        self.d_diff.zero_()
        flags = ""
        if "R" in cmd:
            flags = _get_subscript(cmd, "R")
            if "a" in flags:
                if "orig" not in self.have:
                    self._calibrate(self.d_data, self.d_amp33_in, self.d_area, "orig")
                self.d_work.copy_(self.d_data)
                if self.refpix:
                    self.d_amp33.copy_(self.d_amp33_in)
                base = "orig"
            else:
                _lib.check(lib.rip_dark_as_l1_dev(cal.handle, self.G, _ptr(self.d_work), st))
                if self.refpix:
                    self.d_amp33.copy_(self.d_amp33_in)
                if "ref" not in self.have:
                    self._calibrate(self.d_work, self.d_amp33, self.d_area, "ref")
                base = "ref"
            seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            _lib.check(lib.rip_add_read_noise_dev(cal.handle, _ptr(self.d_work), self.G, _lib.ptr(self.rpg), seed, st))
            _lib.check(lib.rip_fill_refdata_1f_dev(cal.handle, _ptr(self.d_work), _ptr(self.d_amp33), self.G,
                                                   _lib.ptr(self.rpg), seed, 1, st))  # fmt: skip
            self._calibrate(self.d_work, self.d_amp33, self.d_area, "noisy")
            _lib.check(lib.rip_active_diff_dev(self.device, _ptr(self.l2["noisy"][0]), _ptr(self.l2[base][0]), self.n,
                                               cal.nb, _ptr(self.d_diff), st))  # fmt: skip
            if "z" in flags:
                zclip = float(_get_subscript(flags.upper(), "Z"))
                p25, p50, p75 = sky.percentiles_device(self.d_diff.data_ptr(), self.na * self.na, (25, 50, 75),
                                                       device=self.device, stream=st.value or 0)  # fmt: skip
                iqr, med = p75 - p25, p50
                _lib.check(lib.rip_clip_dev(self.device, _ptr(self.d_diff), self.na * self.na,
                                            float(med - zclip * iqr / 1.34896), float(med + zclip * iqr / 1.34896), st))
        if "O" in cmd:  # pseudo-Poisson draws from the Pearson family (reference :173-227)
            if "orig" not in self.have:
                self._calibrate(self.d_data, self.d_amp33_in, self.d_area, "orig")
            G, meta = self.G, self.dplan.meta
            start = 1 if self.config.get("EXCLUDE_FIRST", True) else 0
            wv = {G - 1: np.array([self.dplan.plan.var_K[0][j] for j in range(G)], np.float32)}  # processinfo weights
            for iend in range(start + 2, G):
                kt = np.zeros(G, dtype=np.float32)
                kt[iend - 1] = 1.0 / (meta["tbar"][iend - 1] - meta["tbar"][start])
                kt[start] = -kt[iend - 1]
                wv[iend - 1] = kt
            tab, defined = np.zeros((G, 3), np.float64), np.zeros(G, np.uint8)
            t_fr = self.frame_time
            for i in range(start + 1, G):
                n21, n31, n41 = tilde_nus(self.read_pattern, wv[i])
                tab[i] = (n21 * t_fr, n31 * t_fr**2, n41 * t_fr**3)  # e/frame -> e/s (reference :212-216)
                defined[i] = 1
            bad = self.torch.zeros(1, dtype=self.torch.int32, device=self.dev)
            _lib.check(lib.rip_pearson_noise_dev(cal.handle, _ptr(self.d_withsky), _ptr(self.d_endslice), G, start,
                                                 _lib.ptr(tab), _lib.ptr(defined), (int(seed) + 13) & 0xFFFFFFFFFFFFFFFF,
                                                 _ptr(self.d_diff), _ptr(bad), st))  # fmt: skip
            nbad = int(bad.item())
            if nbad:
                # (reference GalPoisson/draw_with_tilnus.py:565-566: "Some intensities give invalid Pearson-IV parameters.")
                raise ValueError(f"noise directive {cmd!r}: {nbad} pixels give invalid Pearson parameters")
        if "P" in cmd:
            pflags = _get_subscript(cmd, "P")
            if "orig" not in self.have:
                self._calibrate(self.d_data, self.d_amp33_in, self.d_area, "orig")
            if "b" in pflags:  # background only: the sky model of the given order (reference :191-195)
                b_order = int("0" + _get_subscript(pflags.upper(), "B"))
                sky.medfit_device(self.d_withsky.data_ptr(), self.na, self.na, self.na, order=b_order, device=self.device,
                                  stream=st.value or 0, subtract=False, d_model=self.d_skylevel.data_ptr())  # fmt: skip
            else:
                self.d_skylevel.copy_(self.d_withsky)
            if "r" in pflags:
                G, meta = self.G, self.dplan.meta
                start = 1 if self.config.get("EXCLUDE_FIRST", True) else 0
                w = np.zeros((G, G), np.float32)
                wdef = np.zeros(G, np.uint8)
                w[G - 1] = np.array([self.dplan.plan.var_K[0][j] for j in range(G)], np.float32)  # processinfo weights
                wdef[G - 1] = 1
                for iend in range(start + 2, G):  # two-point weights of the truncated ramps (reference :204-208)
                    kt = np.zeros(G, dtype=np.float32)
                    kt[iend - 1] = 1.0 / (meta["tbar"][iend - 1] - meta["tbar"][start])
                    kt[start] = -kt[iend - 1]
                    w[iend - 1], wdef[iend - 1] = kt, 1
                lastsamp = self.read_pattern[-1][-1]
                gor = np.full(lastsamp + 1, -1, np.int32)
                for j, grp in enumerate(self.read_pattern):
                    for r in grp:
                        if 0 <= r <= lastsamp:
                            gor[r] = j
                _lib.check(lib.rip_poisson_resample_dev(cal.handle, _ptr(self.d_skylevel), _ptr(self.d_endslice), G,
                                                        lastsamp + 1, _lib.ptr(gor), _lib.ptr(w), _lib.ptr(wdef),
                                                        self.frame_time, (int(seed) + 7) & 0xFFFFFFFFFFFFFFFF,
                                                        _ptr(self.d_diff), st))  # fmt: skip
        if "S" in cmd:
            sky_order = int("0" + _get_subscript(cmd, "S"))
            sky.medfit_device(self.d_diff.data_ptr(), self.na, self.na, self.na, order=sky_order, device=self.device,
                              stream=st.value or 0, subtract=True)  # fmt: skip
        if out is not None:
            self.torch.from_numpy(out).copy_(self.d_diff)
            self.torch.cuda.synchronize(self.dev)
            return out
        return self.d_diff.cpu().numpy()

    def close(self):
        if self.owns_cal:
            self.cal.close()


def make_noise_cube_arrays(data, amp33, caldir, read_pattern, frame_time, layers, seed, area_factor=None, config=None, device=0):
    """
    Array-level ``make_noise_cube``: L1 cube u16 [G,n,n] + reference output u16 [G,n,128] -> noise realisations
    float32 [len(layers), n-8, n-8] for the directive strings in ``layers`` (``config["NOISE"]["LAYER"]``).
    """
    import torch  # noqa: PLC0415

    nl = NoiseLayers(caldir, read_pattern, frame_time, config, device)
    try:
        dev = nl.dev
        d_data = torch.from_numpy(np.ascontiguousarray(data).view(np.int16)).to(dev).view(torch.uint16)
        d_amp = torch.from_numpy(np.ascontiguousarray(amp33).view(np.int16)).to(dev).view(torch.uint16)
        d_area = None if area_factor is None else torch.from_numpy(np.ascontiguousarray(area_factor, dtype=np.float32)).to(dev)
        nl.set_exposure(d_data, d_amp, d_area)
        out = np.zeros((len(layers), nl.na, nl.na), np.float32)
        for i, cmd in enumerate(layers):
            out[i] = nl.layer(cmd, int(seed) + 1000 * (i + 1))
        return out
    finally:
        nl.close()


def generate_all_noise(config, device=0):
    """
    Driver for noise generation: the reference's ``generate_all_noise`` (L1_to_L2/gen_noise_image.py:334-391) with the
    same configuration keys -- ``config["NOISE"]`` holds ``LAYER`` (list of directive strings), ``SEED``, ``OUT`` (``TEMP`` is
    accepted and unused: nothing goes through temporary files here).  Reads the L1 exposure ``config["IN"]``; the L2
    quantities the reference takes from ``config["OUT"]`` (``data``, ``data_withsky``, weights, endslice) are recomputed on
    the device from the same inputs.  Writes ``{"config", "noise"}`` (float32, or float16 with ``NOISE_PRECISION: 16``).
    """
    import torch  # noqa: PLC0415

    from ..caltree import write_tree  # noqa: PLC0415
    from ..utils import coordutils  # noqa: PLC0415

    noise = config["NOISE"]
    if config.get("NOISE_PRECISION", 32) not in (16, 32):
        raise ValueError("Unsupported noise precision.")
    data, amp33, read_pattern, frame_time, _, _ = gci.read_l1(config["IN"])
    ent = gci._cached_caldir(config["CALDIR"], device)
    cal = ent["cal"]
    nl = NoiseLayers(cal, read_pattern, frame_time, {**config, "SLICEOUT": True}, device)
    dev = nl.dev
    d_data = torch.from_numpy(np.array(data, dtype=np.uint16).view(np.int16)).to(dev).view(torch.uint16)
    a33 = np.array(amp33, dtype=np.uint16) if amp33 is not None else np.zeros((data.shape[0], cal.n, 128), np.uint16)
    d_amp = torch.from_numpy(np.ascontiguousarray(a33).view(np.int16)).to(dev).view(torch.uint16)
    d_area = None
    wcs = coordutils.wcs_from_config(config)
    if wcs is not None:  # AreaFactor of the exposure (gen_cal_image.py:618-621), as calibrateimage used it
        area = coordutils.pixelarea_device(wcs, N=cal.n, inv_omega=1.0 / gci.pars.Omega_ideal, dtype=np.float32, device=device)
        d_area = torch.from_numpy(np.ascontiguousarray(area, dtype=np.float32)).to(dev)
    nl.set_exposure(d_data, d_amp, d_area)
    layers = list(noise["LAYER"])
    noiseimage = np.zeros((len(layers), nl.na, nl.na), np.float32)
    for i, cmd in enumerate(layers):
        noiseimage[i] = nl.layer(cmd, int(noise["SEED"]) + 1000 * (i + 1))
    print(noiseimage.shape)
    print("percentiles:")
    for q in (5, 25, 50, 75, 95):
        print(q, np.percentile(noiseimage, q, axis=(1, 2)))
    if config.get("NOISE_PRECISION", 32) == 16:
        noiseimage = noiseimage.astype(np.float16)
    write_tree(noise["OUT"], {"config": gci._plain_copy(config), "noise": noiseimage})
    if config.get("FITSOUT", False):
        from ..io import fits_lite  # noqa: PLC0415

        # (FITS has no float16: float32 as in the reference, :386-390)
        fits_lite.write_hdus(noise["OUT"][:-5] + "_asdf_to.fits", [(noiseimage.astype(np.float32), None)])
