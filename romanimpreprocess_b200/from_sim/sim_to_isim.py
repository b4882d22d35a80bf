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
the reference's ``validation_tests/many_realizations.py``).  Cosmic-ray injection (``romanisim.cr``, switched on by
``crparam={}`` at sim_to_isim.py:238) is not restated: the returned ``dq`` carries the linearity file's flags only.
"""

import ctypes as C

import numpy as np

from .. import _lib, pars
from ..L1_to_L2.gen_cal_image import CalDir

READ_TIME = 3.04  # romanisim.parameters.read_time [s]


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
               quantize=True):  # fmt: skip
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
    finally:
        if cal is not caldir:
            cal.close()
    return out, with_dq


__all__ = ["make_l1_fullcal", "read_pattern_from_reads", "fwd_params", "pars"]
