"""Scene -> L1 forward ramp model on the GPU: drop-in for the numerics of ``romanimpreprocess.from_sim.sim_to_isim``.

Functions
---------
make_l1_fullcal
    Total electrons per pixel -> MultiAccum resultants in DN with the full calibration information (reset noise,
    per-read binomial apportioning, IPC + gain + inverse linearity per read, group averaging, read noise, bias
    correction, rounding): reference from_sim/sim_to_isim.py:163-262, one CUDA kernel (``rip_make_l1_host``).
read_pattern_from_reads
    ``READS`` list of the YAML configuration -> read pattern (reference sim_to_isim.py:970-974).

The reference delegates the apportioning and the read noise to ``romanisim.l1`` (whose source is not part of the
reference) and draws from GalSim deviates.  The kernel restates them with a counter-based Philox
generator, so individual realisations differ from the reference's while their statistics agree (validated as in
the reference's ``validation_tests/many_realizations.py``).  Cosmic-ray injection (``romanisim.cr.simulate_crs``,
switched on by ``crparam={}`` at sim_to_isim.py:238) is restated from its published algorithm in the kernel
``fwd_cr_kernel`` (``crparam`` argument of ``make_l1_fullcal``; the reference's behaviour is ``crparam={}``): the
returned ``dq`` then carries JUMP_DET in the groups a cosmic ray hit, as romanisim's does.
"""

import ctypes as C

import numpy as np

from .. import _lib, pars
from ..L1_to_L2.gen_cal_image import CalDir

READ_TIME = 3.04  # romanisim.parameters.read_time [s]
# defaults of romanisim.cr.simulate_crs (what crparam={} selects)
CR_DEFAULTS = {"flux": 8.0, "area": 16.8, "conversion_factor": 0.5, "pixel_size": 10.0, "pixel_depth": 5.0}


def read_pattern_from_reads(reads):
    """[a0, b0, a1, b1, ...] -> [[a0..b0-1], [a1..b1-1], ...]."""
    return [list(range(int(reads[2 * i]), int(reads[2 * i + 1]))) for i in range(len(reads) // 2)]


def _seed_from(rng, seed):
    if seed is not None:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    if rng is None:
        raise ValueError("make_l1_fullcal needs a random number generator or a seed")  # the reference fails too
    if isinstance(rng, (int, np.integer)):
        return int(rng) & 0xFFFFFFFFFFFFFFFF
    if hasattr(rng, "raw"):  # galsim.BaseDeviate
        return int(rng.raw()) & 0xFFFFFFFFFFFFFFFF
    if hasattr(rng, "integers"):  # np.random.Generator
        return int(rng.integers(0, 2**63 - 1))
    raise TypeError("rng must be an int seed, a galsim deviate or a numpy Generator")


def fwd_params(read_pattern, seed, read_time=READ_TIME, add_read_noise=True, add_reset_noise=True, add_biascorr=True,
               quantize=True, crparam=None):  # fmt: skip
    G = len(read_pattern)
    if G > _lib.RIP_GMAX:
        raise ValueError(f"the GPU forward model supports up to {_lib.RIP_GMAX} resultants")
    flat = [int(r) for grp in read_pattern for r in grp]
    if len(flat) > 64:
        raise ValueError("the GPU forward model supports up to 64 reads per exposure")
    prm = _lib.FwdParams()
    prm.G, prm.n_reads = G, len(flat)
    for g, grp in enumerate(read_pattern):
        prm.reads_per_group[g] = len(grp)
    for k, r in enumerate(flat):
        prm.read_index[k] = r
    prm.read_time = float(read_time)
    prm.seed = int(seed)
    prm.add_read_noise, prm.add_reset_noise = int(bool(add_read_noise)), int(bool(add_reset_noise))
    prm.add_biascorr, prm.quantize = int(bool(add_biascorr)), int(bool(quantize))
    if crparam is not None:  # romanisim: `if crparam is not None: cr.simulate_crs(..., **crparam)`
        unknown = set(crparam) - set(CR_DEFAULTS)
        if unknown:
            raise TypeError(f"simulate_crs() got unexpected keyword arguments {sorted(unknown)}")
        cr = {**CR_DEFAULTS, **crparam}
        prm.cr_enable = 1
        prm.cr_flux, prm.cr_area = float(cr["flux"]), float(cr["area"])
        prm.cr_conversion_factor = float(cr["conversion_factor"])
        prm.cr_pixel_size, prm.cr_pixel_depth = float(cr["pixel_size"]), float(cr["pixel_depth"])
    return prm


def make_l1_fullcal(counts, read_pattern, caldir, rng=None, persistence=None, tstart=None, seed=None, device=0,
                    cum_counts=None, **flags):  # fmt: skip
    """
    Make an L1 image with the full calibration information.

    Parameters
    ----------
    counts : galsim.Image or np.ndarray (int32, [4088, 4088])
        Number of electrons per pixel per exposure.
    read_pattern : list of list of int
        MultiAccum table.
    caldir : dict or CalDir
        Dictionary of the reference files (or an already resident ``CalDir``).
    rng : int, galsim.BaseDeviate or np.random.Generator
        Source of the seed of the counter-based generator of the kernel.
    persistence, tstart
        Accepted for signature compatibility; not used (as in the reference: "not used yet").
    crparam : dict or None (keyword)
        ``None`` (default here): no cosmic rays.  A dict -- ``{}`` is what the reference passes -- switches on
        ``romanisim.cr.simulate_crs`` per read with these overrides of its defaults (``CR_DEFAULTS``).

    Returns
    -------
    l1 : np.ndarray, float32 (ngrp, ny, nx)   resultants in DN (integers after rounding)
    dq : np.ndarray, uint32 (ngrp, ny, nx)
    """
    arr = np.asarray(counts.array if hasattr(counts, "array") else counts)
    cal = caldir if isinstance(caldir, CalDir) else CalDir(caldir, device)
    try:
        if arr.shape != (cal.na, cal.na):
            raise ValueError(f"counts must cover the active array ({cal.na},{cal.na}), got {arr.shape}")
        c = np.ascontiguousarray(np.clip(arr, 0, 2000000000), dtype=np.int32)
        prm = fwd_params(read_pattern, _seed_from(rng, seed), **flags)
        cc = None if cum_counts is None else np.ascontiguousarray(cum_counts, dtype=np.int32)
        out = np.empty((prm.G, cal.na, cal.na), np.float32)
        _lib.check(_lib.lib().rip_make_l1_host(cal.handle, _lib.ptr(c), _lib.ptr(cc), C.byref(prm), _lib.ptr(out)))
        with_dq = np.empty((prm.G, cal.na, cal.na), np.uint32)
        with_dq[...] = cal.lin_dq_active[None]
        if prm.cr_enable:  # romanisim: dq[group, crhits] |= JUMP_DET
            from ..dqflags import pixel  # noqa: PLC0415

            crg = np.empty((cal.na, cal.na), np.uint32)
            _lib.check(_lib.lib().rip_fwd_cr_groups_host(cal.handle, _lib.ptr(crg)))
            for g in range(prm.G):
                with_dq[g] |= np.where((crg >> g) & 1, np.uint32(pixel.JUMP_DET), np.uint32(0))
    finally:
        if cal is not caldir:
            cal.close()
    return out, with_dq


def noise_1f_frame(rng, nside=None, seed=None, device=0, nframes=1, draws=None):
    """
    1/f noise block(s), S(f) = 1/f, shape (nside, nside/32): reference from_sim/sim_to_isim.py:265-303.

    The reference fills a 2*nside*channelwidth-point complex spectrum from ``galsim.GaussianDeviate`` and takes the
    real part of the first half of its forward FFT; here the spectrum is generated in the kernel (Philox) and
    transformed by a four-step shared-memory FFT (``rip_noise_1f_frames_host``).  ``draws`` (float64, [nframes, 2*m])
    replaces the generator by a given N(0,1) stream (deterministic; tests).
    """
    nside = pars.nside if nside is None else int(nside)
    out = np.empty((nframes, nside, nside // 32), np.float32)
    d = None if draws is None else np.ascontiguousarray(draws, dtype=np.float64)
    sd = 0 if draws is not None else _seed_from(rng, seed)
    _lib.check(_lib.lib().rip_noise_1f_frames_host(device, nside, nframes, sd, _lib.ptr(d), _lib.ptr(out)))
    return out[0] if nframes == 1 else out


def fill_in_refdata_and_1f(im, caldir, rng, tij, fill_in_banding=True, amp33=None, seed=None, device=0):
    """
    Fills in reference pixel data and 1/f noise (in place): reference from_sim/sim_to_isim.py:306-402.

    Parameters
    ----------
    im : np.ndarray, uint16 (ngroup, n, n)
        The simulated L1 cube; active pixels are kept, reference pixels are generated, banding is added everywhere.
    caldir : dict or CalDir
    rng : int, galsim.BaseDeviate or np.random.Generator (source of the seed)
    tij : list of list of float
        Read times per group (only the group lengths matter, as in the reference).
    amp33 : np.ndarray, uint16 (ngroup, n, n/32), optional
        Reference output, filled if the read file carries amp33 statistics.
    """
    cal = caldir if isinstance(caldir, CalDir) else CalDir(caldir, device)
    try:
        G = im.shape[0]
        if im.dtype != np.uint16 or not im.flags["C_CONTIGUOUS"]:
            raise TypeError("im must be a C-contiguous uint16 cube")
        rpg = np.ascontiguousarray([len(t) for t in tij], dtype=np.int32)
        if len(rpg) != G:
            raise ValueError("tij must list the reads of every group of im")
        a33 = None
        if amp33 is not None:
            if amp33.dtype != np.uint16 or not amp33.flags["C_CONTIGUOUS"]:
                raise TypeError("amp33 must be a C-contiguous uint16 cube")
            a33 = amp33
        _lib.check(_lib.lib().rip_fill_refdata_1f_host(cal.handle, _lib.ptr(im), _lib.ptr(a33), G, _lib.ptr(rpg),
                                                       _seed_from(rng, seed), int(bool(fill_in_banding))))  # fmt: skip
    finally:
        if cal is not caldir:
            cal.close()


def sim_calprep(caldir, device=0):
    """The scene's calibration planes of ``Image2D.simulate`` (reference sim_to_isim.py:615-633):
    (this_dark [e/s], this_flat), both IPC-deconvolved on the active array and clipped."""
    cal = caldir if isinstance(caldir, CalDir) else CalDir(caldir, device)
    try:
        d = np.empty((cal.na, cal.na), np.float32)
        f = np.empty((cal.na, cal.na), np.float32)
        _lib.check(_lib.lib().rip_sim_calprep(cal.handle, _lib.ptr(d), _lib.ptr(f)))
    finally:
        if cal is not caldir:
            cal.close()
    return d, f


def simulate_counts(image, caldir, read_pattern, rng=None, seed=None, area_ratio=None, cnorm=1.0, read_time=READ_TIME,
                    counts=None, dark=False, return_rate=False, device=0, sky=None):  # fmt: skip
    """
    Electrons per pixel of one exposure from a noiseless scene (reference sim_to_isim.py:636-648):
    ``counts += Poisson(clip(C * t * g / g_ideal * image * this_flat / area_ratio, 0))`` with
    ``t = read_time * (last read - first read)``.

    image : float32 (na, na), e/s per ideal pixel; area_ratio : pixel area / Omega_ideal (na, na) or None;
    counts : int32 (na, na) to accumulate into (the reference adds to romanisim's dark/sky counts) or None;
    dark : also draw the dark electrons Poisson(this_dark * t) (romanisim's term, restated).
    sky : sky background level in e/s/pixel (scalar or (na, na) plane) or None: adds Poisson(sky * this_flat * t), the
        background term of romanisim's ``simulate_counts`` (restated; the LEVEL comes from ``galsim.roman.getSkyLevel``
        -- zodiacal-light tables that ship with GalSim, absent here -- plus stray light and thermal background, so the
        caller supplies it).
    """
    cal = caldir if isinstance(caldir, CalDir) else CalDir(caldir, device)
    try:
        t = float(read_time) * (read_pattern[-1][-1] - read_pattern[0][0])
        img = np.ascontiguousarray(image, dtype=np.float32)
        if img.shape != (cal.na, cal.na):
            raise ValueError(f"image must cover the active array ({cal.na},{cal.na}), got {img.shape}")
        ar = None if area_ratio is None else _lib.as_float_plane(area_ratio)
        acc = counts is not None
        out = np.ascontiguousarray(counts, dtype=np.int32).copy() if acc else np.empty((cal.na, cal.na), np.int32)
        rate = np.empty((cal.na, cal.na), np.float64) if return_rate else None
        _lib.check(_lib.lib().rip_sim_counts_host(cal.handle, _lib.ptr(img), _lib.ptr(ar),
                                                  _lib.float_tag(ar) if ar is not None else _lib.RIP_F32, t, float(cnorm),
                                                  float(pars.g_ideal), t if dark else 0.0, _seed_from(rng, seed),
                                                  _lib.ptr(out), int(acc), _lib.ptr(rate)))  # fmt: skip
        if sky is not None:
            # Poisson(sky * this_flat * t) through the same kernel: its mean is C t g/g_ideal * image * flat / area, so
            # the "image" of the sky term is sky * g_ideal / g * area (C = 1); own random stream (seed + 1)
            from ..caltree import open_tree  # noqa: PLC0415

            with open_tree(cal.source["gain"]) as f:
                g_act = np.asarray(f["roman"]["data"])[cal.nb : cal.n - cal.nb, cal.nb : cal.n - cal.nb].astype(np.float64)
            sk = np.broadcast_to(np.asarray(sky, dtype=np.float64), (cal.na, cal.na)) * float(pars.g_ideal) / g_act
            if ar is not None:
                sk = sk * ar
            sk = np.ascontiguousarray(sk, dtype=np.float32)
            _lib.check(_lib.lib().rip_sim_counts_host(cal.handle, _lib.ptr(sk), _lib.ptr(ar),
                                                      _lib.float_tag(ar) if ar is not None else _lib.RIP_F32, t, 1.0,
                                                      float(pars.g_ideal), 0.0, (_seed_from(rng, seed) + 1) & 0xFFFFFFFFFFFFFFFF,
                                                      _lib.ptr(out), 1, None))  # fmt: skip
    finally:
        if cal is not caldir:
            cal.close()
    return (out, rate) if return_rate else out


# ---------------------------------------------------------------------------------------------------------------
# File-level driver: run_config / Image2D (reference from_sim/sim_to_isim.py:63-160, 405-520, 612-700, 793-812, 947-997)
# ---------------------------------------------------------------------------------------------------------------
L1_TAG = "asdf://stsci.edu/datamodels/roman/tags/wfi_science_raw-1.0.0"


def _sip_flip(header, axis):
    """WCS part of ``hdu_sip_hflip`` (axis 1) / ``hdu_sip_vflip`` (axis 2): reference :63-160, on a header dict in place.
    ``n`` = image size along the flipped axis comes from NAXIS<axis>."""
    n = int(header[f"NAXIS{axis}"])
    header[f"CRPIX{axis}"] = n + 1 - header[f"CRPIX{axis}"]
    for key in (f"CD1_{axis}", f"CD2_{axis}"):
        if key in header:
            header[key] = -header[key]
    try:
        a_order, b_order = int(header["A_ORDER"]), int(header["B_ORDER"])
    except (ValueError, KeyError):
        return
    # the powers of the flipped SIP axis: u^p for axis 1, v^q for axis 2.  A (the u distortion) changes sign for even
    # powers of u / odd powers of v, B (the v distortion) for odd powers of u / even powers of v.
    for name, order, parity in (("A", a_order, 0 if axis == 1 else 1), ("B", b_order, 1 if axis == 1 else 0)):
        for p in range(order + 1):
            for q in range(order + 1 - p):
                if (p if axis == 1 else q) % 2 == parity:
                    key = f"{name}_{p:1d}_{q:1d}"
                    if key in header:
                        header[key] = -float(header[key])


class Image2D:
    """2D scene with its WCS: the reference's ``Image2D("anlsim", fname=...)`` (an OpenUniverse-2024 style "truth" FITS image
    in electrons per exposure), ``simulate`` through a CALDIR and ``L1_write_to``.  Scene electrons, ramps, cosmic rays,
    reference pixels and 1/f noise are generated on the GPU; FITS / ASDF files by ``io/fits_lite.py`` / ``caltree``.

    Not reproduced: romanisim's built-in reference data (``caldir=None``), its metadata bookkeeping (a minimal ``meta`` with
    exposure / instrument / pointing entries is written), the idealised L2 product (``L2_write_to``), and the zodiacal sky
    level of ``galsim.roman.getSkyLevel`` (configure ``SKY_E_PER_S`` [e/s/pixel], default 0)."""

    def __init__(self, intype, **kwargs):
        if intype != "anlsim":
            raise ValueError(f"Image2D: unknown input type {intype!r}")
        self.init_anlsim(kwargs["fname"], flip=kwargs.get("flip", True))

    def init_anlsim(self, fname, flip=True):
        import re  # noqa: PLC0415

        from ..io import fits_lite  # noqa: PLC0415

        m = re.search(r"_(\d+)_(\d+)\.fits", fname)
        self.idsca = (int(m.group(1)), int(m.group(2)))
        data, self.header = fits_lite.read_primary(fname)
        data = np.array(data, dtype=np.float64)
        if flip:  # SCAs are flipped depending on which row of the focal plane they are in (:489-493)
            if self.idsca[1] % 3 == 0:
                data = data[:, ::-1]
                _sip_flip(self.header, 1)
            else:
                data = data[::-1, :]
                _sip_flip(self.header, 2)
        self.image = np.ascontiguousarray(data / float(self.header["EXPTIME"]), dtype=np.float32)  # electrons per second
        self.header["CRPIX1"] -= 1  # offset from FITS -> GWCS convention (:497-498)
        self.header["CRPIX2"] -= 1
        date = str(self.header.get("DATE-OBS", "2025-01-01T00:00:00.000000"))
        self.date = date.replace(" ", "T") + ("Z" if "DATE-OBS" in self.header else "")
        self.filter = str(self.header.get("FILTER", "F184"))[:4]
        self.ra_, self.dec_, self.pa_ = (float(self.header.get(k, 0.0)) for k in ("RA_TARG", "DEC_TARG", "PA_OBSY"))

    def simulate(self, use_read_pattern, caldir=None, config=None, seed=43, includewcs=False, device=0):
        from ..L1_to_L2 import gen_cal_image as gci  # noqa: PLC0415
        from ..io import asdf_lite  # noqa: PLC0415
        from ..utils import coordutils  # noqa: PLC0415

        config = config or {}
        if caldir is None:
            raise NotImplementedError("Image2D.simulate without CALDIR uses romanisim's built-in reference data: not available here")
        cal = gci._cached_caldir(caldir, device)["cal"]
        na, n, nb = cal.na, cal.n, cal.nb
        if self.image.shape != (na, na):
            raise ValueError(f"the scene must cover the active array ({na},{na}), got {self.image.shape}")
        rp = [list(g) for g in use_read_pattern]
        G = len(rp)
        # pixel area / Omega_ideal on the active array (:645; the flat already contains the pixel area)
        wcs_hdr = dict(self.header, CRPIX1=self.header["CRPIX1"] + 1, CRPIX2=self.header["CRPIX2"] + 1)
        area = coordutils.pixelarea_device(coordutils.FitsWCS(wcs_hdr), N=na, inv_omega=1.0 / pars.Omega_ideal,
                                           dtype=np.float64, device=device)  # fmt: skip
        cnorm = float(config.get("CNORM", 1.0))
        counts = simulate_counts(self.image, cal, rp, seed=seed, area_ratio=area, cnorm=cnorm, dark=True,
                                 sky=config.get("SKY_E_PER_S"), device=device)  # fmt: skip
        # cosmic rays as the reference switches them on (crparam={}: romanisim's defaults); the default detector area of
        # 16.8 cm^2 belongs to the full 4088^2 array and scales with the frame for the small frames of the tests
        crparam = {"area": CR_DEFAULTS["area"] * (na / 4088.0) ** 2}
        l1, _ = make_l1_fullcal(counts, rp, cal, seed=(int(seed) + 1) & 0xFFFFFFFFFFFFFFFF, crparam=crparam)
        data = np.zeros((G, n, n), np.uint16)
        data[:, nb : n - nb, nb : n - nb] = np.clip(l1, 0, 65535).astype(np.uint16)
        amp33 = None
        if cal.has_amp33 and not (isinstance(caldir, dict) and caldir.get("NO_AMP33")):
            amp33 = np.zeros((G, n, 128), np.uint16)
        tij = [[READ_TIME * r for r in g] for g in rp]
        fill_in_refdata_and_1f(data, cal, (int(seed) + 2) & 0xFFFFFFFFFFFFFFFF, tij, fill_in_banding=True, amp33=amp33)
        sca = self.idsca[1]
        meta = asdf_lite.TaggedDict({
            "exposure": {"read_pattern": rp, "frame_time": READ_TIME, "nresultants": G, "ma_table_number": 1000000,
                         "start_time": self.date},
            "instrument": {"name": "WFI", "detector": f"WFI{sca:02d}", "optical_element": "F" + self.filter[1:]},
            "observation": {"visit": int(self.idsca[0])},
            "pointing": {"ra_v1": self.ra_, "dec_v1": self.dec_, "pa_v3": self.pa_},
            "model_type": "ScienceRawModel",
        })  # fmt: skip
        im = {"meta": meta, "data": data}
        if amp33 is not None:
            im["amp33"] = amp33
        if "EXTRACT_REF" in config:  # reference read moved out of the cube (:711-735)
            off = int(config["EXTRACT_REF"].get("data_encoding_offset", 0))
            meta["instrument"]["data_encoding_offset"] = off
            meta["exposure"]["read_pattern"] = rp[1:]
            for key, refkey in (("data", "reference_read"), ("amp33", "reference_amp33")):
                if key not in im:
                    continue
                cube = im[key]
                im[refkey] = np.copy(cube[0])
                modref = cube[0].astype(np.int32) - off
                for k in range(1, G):
                    cube[k] = np.clip(cube[k].astype(np.int32) - modref, 0, 65535).astype(np.uint16)
                im[key] = cube[1:]
        self.tree = {"roman": asdf_lite.TaggedDict(im, tag=L1_TAG), "romanisim": {"version": "romanimpreprocess_b200"}}

    def L1_write_to(self, filename):
        from ..caltree import write_tree  # noqa: PLC0415

        if not hasattr(self, "tree"):
            return False
        write_tree(filename, self.tree)
        return True


def run_config(config, device=0):
    """L1 image construction from a configuration dictionary: the reference's ``run_config`` (from_sim/sim_to_isim.py:947-997)
    with the same keys (``IN``: truth FITS image, ``OUT``: L1 ASDF file, ``READS``, ``CALDIR``, ``SEED``, ``CNORM``,
    ``EXTRACT_REF``, ``FITSOUT``).  Also writes the FITS WCS header text ``<OUT>_asdf_wcshead.txt`` that
    ``calibrateimage`` reads back as ``FITSWCS``."""
    from ..io import fits_lite  # noqa: PLC0415

    caldir = config.get("CALDIR", None)
    rp = read_pattern_from_reads(config["READS"])
    seed = int(config.get("SEED", 43))
    x = Image2D("anlsim", fname=config["IN"])
    x.simulate(rp, caldir=caldir, config=config, seed=seed, device=device)
    x.L1_write_to(config["OUT"])
    with open(config["OUT"][:-5] + "_asdf_wcshead.txt", "w") as f:
        f.write(fits_lite.header_text(x.header, comment="truth wcs from sim_to_isim"))
    if config.get("FITSOUT", False):
        r = x.tree["roman"]
        d = np.asarray(r["data"])
        image_out = np.zeros((d.shape[0], d.shape[1], d.shape[2] + 128), np.uint16)
        image_out[:, :, : d.shape[2]] = d
        if "amp33" in r:
            image_out[:, :, d.shape[2] :] = r["amp33"]
        fits_lite.write_hdus(config["OUT"][:-5] + "_asdf_to.fits", [(image_out, None)])


__all__ = ["make_l1_fullcal", "read_pattern_from_reads", "fwd_params", "noise_1f_frame", "fill_in_refdata_and_1f",
           "sim_calprep", "simulate_counts", "pars", "Image2D", "run_config"]  # fmt: skip
