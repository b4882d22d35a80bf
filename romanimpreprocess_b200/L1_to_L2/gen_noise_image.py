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
``S<order>``, ``C<tag>`` (a label: ignored, as in the reference) -- i.e. the production layers ``Rz4PbrS2C*``.
``O`` (Pearson pseudo-Poisson draws) is not on the GPU path and raises ``NotImplementedError``.  Random numbers are Philox (the reference: GalSim): layers are validated statistically.
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
        if "O" in cmd:
            raise NotImplementedError(f"noise directive {cmd!r}: the Pearson draws of 'O' are not generated on the GPU")
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
