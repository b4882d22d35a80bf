import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from romanimpreprocess_b200 import pars, synth
from romanimpreprocess_b200.validation_tests import many_realizations as mr
n, R = 256, 3
rp = synth.README_PATTERN
cal = synth.make_caldir(n=n, read_pattern=rp, p_order=10, seed=77)
na = n - 8
yy, xx = np.mgrid[0:na, 0:na]
image = (20.0 + 0.1 * xx).astype(np.float32)
Z = mr.Realizations(image, cal, rp, keep_stacks=R)
for j in range(R):
    Z.step(100 + 10 * (j + 1))
    torch.cuda.synchronize()
    pdq = Z.d_pdq.cpu().numpy().view(np.uint32)
    im = Z.d_im.cpu().numpy().view(np.uint16)
    print("real", j, "counts", Z.d_counts.float().mean().item(), "res last", Z.d_res[-1].mean().item(), "im[-1] act mean", im[-1, 4:-4, 4:-4].mean(), "im[1]", im[1, 4:-4, 4:-4].mean(),
          "slope mean", Z.d_slope[8:-8, 8:-8].mean().item(), "moments0 mean", Z.d_moments[0].mean().item())
    for b in range(32):
        c = np.count_nonzero(pdq[4:-4, 4:-4] & np.uint32(1 << b))
        if c: print("   bit", b, c)
