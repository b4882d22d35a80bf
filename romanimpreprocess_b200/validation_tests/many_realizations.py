"""Many noise realisations of one scene, entirely on the GPU: the protocol of the reference's
``validation_tests/many_realizations.py`` (reference lines 58-106; BASELINE configs[4]).

For realisation j (seed += 10 before each, reference :66-67) the reference runs ``sim_to_isim.run_config`` then
``gen_cal_image.calibrateimage`` through ASDF files and accumulates, per active pixel, the count of unmasked
realisations, the mean and the scatter of the calibrated slope, plus the medians over realisations of the raw group
difference, the slope and its error.  Here every realisation stays in HBM:

    scene electrons  (rip_sim_counts_dev)      -> ramp (rip_make_l1_dev) -> u16 cube (rip_l1_embed_dev)
    reference pixels + 1/f + amp33 (rip_fill_refdata_1f_dev) -> fused L1->L2 (rip_l1_to_l2_dev)
    -> grown mask + moment sums (rip_moments_accumulate_dev), stacks (rip_realization_record_dev)
    -> rip_moments_finalize_dev, rip_stack_median_dev

``torch`` only owns the device buffers.  Realisations are independent: ``ranks``/``rank`` split them across GPUs
(one process per GPU); the three moment planes are then summed over ranks by the caller (``torch.distributed``
all-reduce, the only exchange step of this workload).

Cosmic rays: the reference's forward model injects them (``crparam={}`` at from_sim/sim_to_isim.py:238 = romanisim's
defaults); so does this one by default (``crparam={}``; the detector area scales with the frame for small test frames),
``crparam=None`` switches them off.

Deviations from the reference script, stated: sky subtraction and WCS are outside the hot path (``images`` are the
flat-fielded slopes; the caller supplies the area plane); ``err`` is ``hypot(err_read, err_poisson)``; the reference
maps ``images`` and ``err`` onto the SAME memmap file (lines 55-56), so its planes 2 and 7 are both medians of the
last-written ``err`` -- ``emulate_alias=True`` reproduces that, the default keeps them separate.
"""

import ctypes as C

import numpy as np

from .. import _lib, pars
from ..from_sim import sim_to_isim as s2i
from ..L1_to_L2 import gen_cal_image as gci
from ..utils import maskhandling


def _p(t):
    return C.c_void_p(t.data_ptr())


class Realizations:
    """Device-resident state of a many-realisations run of one SCA (one process / GPU)."""

    def __init__(self, image, caldir, read_pattern, area_ratio=None, config2=None, cnorm=1.0, device=0,
                 keep_stacks=0, dark=True, fill_in_banding=True, read_time=s2i.READ_TIME, crparam={}):  # fmt: skip  # noqa: B006
        import torch  # noqa: PLC0415  (device memory only)

        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.device = device
        self.cal = caldir if isinstance(caldir, gci.CalDir) else gci.CalDir(caldir, device)
        self.owns_cal = self.cal is not caldir
        cal = self.cal
        n, na, G = cal.n, cal.na, len(read_pattern)
        self.n, self.na, self.G = n, na, G
        self.read_pattern = read_pattern
        self.read_time = float(read_time)
        self.cnorm, self.dark, self.banding = float(cnorm), bool(dark), bool(fill_in_banding)
        self.t_exp = self.read_time * (read_pattern[-1][-1] - read_pattern[0][0])
        self.rpg = np.ascontiguousarray([len(g) for g in read_pattern], dtype=np.int32)
        # cosmic rays as in the reference (romanisim defaults; 16.8 cm^2 is the area of the full 4088^2 array)
        self.crparam = None if crparam is None else {"area": s2i.CR_DEFAULTS["area"] * (na / 4088.0) ** 2, **crparam}
        self.grow = np.ascontiguousarray(maskhandling.PixelMask1.array)
        # the reference output is 128 columns wide whatever the frame side (gen_cal_image.py:531-556 and the library's
        # reference-pixel statistics assume it); debugging frames with n/32 != 128 run without that correction
        self.refpix = bool(cal.has_amp33 and n // 32 == 128)
        self.dplan = gci.DevicePlan(cal, read_pattern, self.read_time, config2 or {}, do_refpix=self.refpix,
                                    area_dtype=np.float32)  # fmt: skip
        z = dict(device=self.dev)
        self.d_image = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(self.dev)
        assert tuple(self.d_image.shape) == (na, na), "scene must cover the active array"
        # the area plane enters twice, as in the reference: active window for the scene rate (sim_to_isim.py:645),
        # full frame for the flat division (gen_cal_image.py:618-622)
        full = np.ones((n, n), np.float32)
        if area_ratio is not None:
            ar = np.asarray(area_ratio, dtype=np.float32)
            full = ar if ar.shape == (n, n) else np.pad(ar, cal.nb, mode="edge")
        self.d_area_full = torch.from_numpy(np.ascontiguousarray(full)).to(self.dev)
        self.d_area_act = self.d_area_full[cal.nb : n - cal.nb, cal.nb : n - cal.nb].contiguous()
        self.d_counts = torch.empty((na, na), dtype=torch.int32, **z)
        self.d_res = torch.empty((G, na, na), dtype=torch.float32, **z)
        self.d_im = torch.empty((G, n, n), dtype=torch.uint16, **z)
        self.d_amp33 = torch.zeros((G, n, n // 32), dtype=torch.uint16, **z)
        self.d_slope = torch.empty((n, n), dtype=torch.float32, **z)
        self.d_er = torch.empty((n, n), dtype=torch.float32, **z)
        self.d_ep = torch.empty((n, n), dtype=torch.float32, **z)
        self.d_pdq = torch.empty((n, n), dtype=torch.int32, **z)
        self.d_moments = torch.zeros((3, na, na), dtype=torch.float32, **z)
        self.keep = int(keep_stacks)
        if self.keep:
            self.d_diffs = torch.zeros((self.keep, n, n), dtype=torch.float32, **z)
            self.d_images = torch.zeros((self.keep, n, n), dtype=torch.float32, **z)
            self.d_err = torch.zeros((self.keep, n, n), dtype=torch.float32, **z)
        self.done = 0

    def step(self, seed):
        """One realisation with the given seed, asynchronous on torch's current stream."""
        lib, cal = _lib.lib(), self.cal
        st = C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        _lib.check(lib.rip_sim_counts_dev(cal.handle, _p(self.d_image), _p(self.d_area_act), _lib.RIP_F32, self.t_exp,
                                          self.cnorm, float(pars.g_ideal), self.t_exp if self.dark else 0.0, seed,
                                          _p(self.d_counts), 0, st))  # fmt: skip
        prm = s2i.fwd_params(self.read_pattern, seed, read_time=self.read_time, crparam=self.crparam)
        _lib.check(lib.rip_make_l1_dev(cal.handle, _p(self.d_counts), C.byref(prm), _p(self.d_res), st))
        _lib.check(lib.rip_l1_embed_dev(self.device, _p(self.d_res), self.G, self.n, cal.nb, _p(self.d_im), st))
        _lib.check(lib.rip_fill_refdata_1f_dev(cal.handle, _p(self.d_im), _p(self.d_amp33), self.G, _lib.ptr(self.rpg),
                                               seed, int(self.banding), st))  # fmt: skip
        gci.calibrate_device(cal, self.dplan, self.d_im.data_ptr(), self.d_amp33.data_ptr() if self.refpix else 0,
                             self.d_area_full.data_ptr(), self.d_slope.data_ptr(), self.d_er.data_ptr(),
                             self.d_ep.data_ptr(), self.d_pdq.data_ptr(), stream=st.value or 0)  # fmt: skip
        _lib.check(lib.rip_moments_accumulate_dev(self.device, _p(self.d_slope), _p(self.d_pdq), self.n, cal.nb,
                                                  _lib.ptr(self.grow), _p(self.d_moments), st))  # fmt: skip
        if self.keep:
            assert self.done < self.keep, "more realisations than keep_stacks"
            j = self.done
            _lib.check(lib.rip_realization_record_dev(self.device, _p(self.d_im), self.G, self.n, cal.nb,
                                                      _p(self.d_slope), _p(self.d_er), _p(self.d_ep), _p(self.d_diffs[j]),
                                                      _p(self.d_images[j]), _p(self.d_err[j]), st))  # fmt: skip
        self.done += 1

    def finalize(self, slope_ideal=None, emulate_alias=False):
        """The 8-plane float32 stack of the reference's ``*_many_out.fits`` (reference :86-100), on the host."""
        torch, lib, n, nb = self.torch, _lib.lib(), self.n, self.cal.nb
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        mom = self.d_moments.clone()
        _lib.check(lib.rip_moments_finalize_dev(self.device, _p(mom), self.na * self.na, st))
        out = np.zeros((8, n, n), np.float32)
        if slope_ideal is not None:
            out[0] = slope_ideal
        if self.keep and self.done:
            med = torch.empty((n, n), dtype=torch.float32, device=self.dev)
            stacks = (self.d_diffs, self.d_err if emulate_alias else self.d_images, self.d_err)
            for plane, stack in zip((1, 2, 7), stacks):
                _lib.check(lib.rip_stack_median_dev(self.device, _p(stack), self.done, n * n, _p(med), st))
                out[plane] = med.cpu().numpy()
        out[3:6, nb : n - nb, nb : n - nb] = mom.cpu().numpy()
        out[6] = out[4] - out[0]
        return out

    def close(self):
        if self.owns_cal:
            self.cal.close()


def run(image, caldir, read_pattern, Nrun, seed=100, slope_ideal=None, rank=0, ranks=1, group=None, **kw):
    """Run ``Nrun`` realisations and return the reference's 8-plane stack.  Seeds follow the reference: realisation j
    uses ``seed + 10 (j + 1)``.

    ``ranks`` > 1: this process runs the realisations ``j % ranks == rank`` (one process per GPU) and the partial results
    are combined over ``group`` (an initialised ``torch.distributed`` NCCL group; default: the world) BEFORE they are
    finalised -- the three moment sums by an all-reduce, the per-realisation stacks of the medians by an all-gather --
    so every rank returns the statistics of all ``Nrun`` realisations."""
    mine = [j for j in range(Nrun) if j % ranks == rank]
    emulate_alias = kw.pop("emulate_alias", False)
    if ranks > 1:
        import torch.distributed as dist  # noqa: PLC0415

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("many_realizations.run(ranks > 1) needs an initialised torch.distributed process group: "
                               "moment sums and stacks of the ranks are combined before finalisation")  # fmt: skip
    R = Realizations(image, caldir, read_pattern, keep_stacks=len(mine), **kw)
    try:
        for j in mine:
            R.step(seed + 10 * (j + 1))
        R.torch.cuda.synchronize(R.dev)
        if ranks > 1:
            combine_ranks(R, Nrun, ranks, group)
        return R.finalize(slope_ideal=slope_ideal, emulate_alias=emulate_alias)
    finally:
        R.close()


def combine_ranks(R, Nrun, ranks, group=None):
    """Sum the raw moment planes of all ranks and gather their stacks (the one exchange step of this workload)."""
    import torch.distributed as dist  # noqa: PLC0415

    torch = R.torch
    dist.all_reduce(R.d_moments, group=group)
    kmax = (Nrun + ranks - 1) // ranks
    if not kmax:
        return
    n = R.n
    gathered = []
    for name in ("d_diffs", "d_images", "d_err"):
        mine = getattr(R, name, None)
        pad = torch.zeros((kmax, n, n), dtype=torch.float32, device=R.dev)
        if mine is not None and R.done:
            pad[: R.done] = mine[: R.done]
        parts = [torch.empty_like(pad) for _ in range(ranks)]
        dist.all_gather(parts, pad, group=group)
        # rank r holds the realisations j = r, r + ranks, ...: its first len(range(r, Nrun, ranks)) slices are valid
        gathered.append(torch.cat([parts[r][: len(range(r, Nrun, ranks))] for r in range(ranks)]))
    R.d_diffs, R.d_images, R.d_err = gathered
    R.keep = R.done = Nrun
