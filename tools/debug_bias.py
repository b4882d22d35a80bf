"""Development: median (slope - expected) of one faint flat realisation at 4096^2 for a few switches."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from romanimpreprocess_b200 import pars, synth
from romanimpreprocess_b200.validation_tests import many_realizations as mr
n = 4096
for rpname, po in (("TEST_READ_PATTERN", 3),):
    rp = getattr(synth, rpname)
    cal = synth.make_caldir(n=n, seed=71, read_pattern=rp, p_order=po, gain_dtype=np.float32, ipc_dtype=np.float32, biascorr_amp=3.0)
    na = n - 8
    for level in (1.0,):
        image = np.full((na, na), level, np.float32)
        for seed in (200, 210, 220, 230, 240, 250, 260, 270):
            kw = {}
            z = mr.Realizations(image, cal, rp, keep_stacks=0, crparam=None, **kw)
            z.step(seed); torch.cuda.synchronize()
            pdq = z.d_pdq.cpu().numpy().view(np.uint32)[4:-4, 4:-4]
            slope = z.d_slope.cpu().numpy()[4:-4, 4:-4]
            good = pdq == 0
            x = slope - image / np.float32(pars.g_ideal)
            print(rpname, "seed", seed, "refpix", z.refpix, "median x", float(np.median(x[good])), "mean", float(x[good].mean()), "good", good.mean(), flush=True)
            z.close() if z.owns_cal else None
