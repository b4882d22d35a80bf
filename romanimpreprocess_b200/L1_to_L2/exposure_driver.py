"""Whole WFI exposures (18 SCAs each) through the L1->L2 path on the GPUs of one box.

The reference runs one independent OS process per SCA (Slurm array 1-18, runs/summer2025run/OpenUniverse_to_L1L2.job:4;
the per-exposure loop of runs/summer2025run/OpenUniverse_to_L1L2.py:155-169 calls ``calibrateimage`` once per
(exposure, SCA)).  Here: one process per GPU; the (exposure, SCA) items are dealt to the ranks by
``sharding.assign_items_balanced`` so that each GPU keeps the CALDIRs of only its own SCAs resident (2-3 of the 18 on an
8-GPU box, all 18 = ~100 GB on one GPU), one ``Pipeline`` per resident SCA, and the items of a rank stream through them
with the PCIe copies overlapped.  There is no collective on the data path: ranks never exchange pixels.
"""

import collections

import numpy as np

from .. import sharding
from . import gen_cal_image as gci


class ExposureCalibrator:
    """L1->L2 for the (exposure, SCA) items of one rank.

    Parameters
    ----------
    caldirs : dict
        ``sca -> CALDIR`` (anything ``gen_cal_image.CalDir`` accepts: dict of file names or of in-memory trees).  Only the
        entries of this rank's resident SCAs are opened.
    items : list of (exposure, sca)
        All items of the job (every rank passes the same list).
    read_pattern, frame_time, config : as ``calibrate_arrays``.
    rank, world : this process' place in the job; device : CUDA ordinal.
    depth : exposures in flight per SCA pipeline.
    """

    def __init__(self, caldirs, items, read_pattern, frame_time, config=None, rank=0, world=1, device=0, depth=2,
                 do_refpix=True, area_dtype=np.float32, balanced=True):  # fmt: skip
        assign = sharding.assign_items_balanced if balanced else sharding.assign_items
        self.items = assign(items, rank, world)
        self.scas = sharding.resident_scas(self.items)
        self.rank, self.world, self.device = rank, world, device
        self.area_dtype = area_dtype
        self.cals, self.pipes = {}, {}
        for sca in self.scas:
            self.cals[sca] = gci.CalDir(caldirs[sca], device=device)
            self.pipes[sca] = gci.Pipeline(self.cals[sca], read_pattern, frame_time, config, do_refpix=do_refpix,
                                           depth=depth, want_endslice=bool((config or {}).get("SLICEOUT", False)),
                                           area_dtype=area_dtype)  # fmt: skip
        self.depth = depth

    def run(self, fetch, sink=None, out_buffers=None):
        """Process this rank's items in order.

        ``fetch(exposure, sca) -> (data, amp33, wcs)``: the L1 cube, its amp33 cube (page-locked arrays for real copy
        overlap) and the exposure's WCS (``coordutils.FitsWCS``, header text/dict, or ``None`` for no area division).
        ``sink(exposure, sca, out)`` is called with the result dict of each item (arrays are only valid during the call
        when ``out_buffers`` are recycled).  ``out_buffers``: optional list of preallocated output dicts (pinned), at
        least ``depth`` per resident SCA + 1.  Returns the number of items processed.
        """
        inflight = collections.deque()
        per_sca = collections.Counter()
        free = list(out_buffers) if out_buffers else None
        done = 0

        def retire():
            nonlocal done
            e, sca, t, ob = inflight.popleft()
            out = self.pipes[sca].result(t)
            per_sca[sca] -= 1
            if sink is not None:
                sink(e, sca, out)
            if free is not None:
                free.append(ob)
            done += 1

        for e, sca in self.items:
            # a pipeline slot is reused only after its result was collected; results are collected in submission order
            while per_sca[sca] >= self.depth or (free is not None and not free):
                retire()
            data, amp33, wcs = fetch(e, sca)
            pipe = self.pipes[sca]
            if wcs is not None:
                pipe.set_area_wcs(wcs, dtype=self.area_dtype)
            ob = free.pop() if free is not None else None
            t = pipe.submit(data, amp33, None, out=ob)
            per_sca[sca] += 1
            inflight.append((e, sca, t, ob))
        while inflight:
            retire()
        return done

    def close(self):
        for p in self.pipes.values():
            p.close()
        for c in self.cals.values():
            c.close()
        self.pipes, self.cals = {}, {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
