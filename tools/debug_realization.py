"""development tool: device-time break-down of one realisation at 4096^2 (run under gpurun)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from romanimpreprocess_b200 import _lib, pars, synth
from romanimpreprocess_b200.from_sim import sim_to_isim as s2i
from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
from romanimpreprocess_b200.validation_tests import many_realizations as mr

rp = synth.README_PATTERN
n = 4096
cal, _, area = bench.make_inputs(n, rp, 2, seed=1000)
cd = gci.CalDir(cal, device=0)
na = n - 8
yy, xx = np.mgrid[0:na, 0:na].astype(np.float32)
image = (3.0 + 0.002 * xx + 400.0 * np.exp(-0.5 * (((xx % 512) - 256) ** 2 + ((yy % 512) - 256) ** 2) / 9.0)).astype(np.float32)
rz = mr.Realizations(image, cd, rp, area_ratio=area, config2={"SLICEOUT": True}, device=0, keep_stacks=0)
rz.step(1); rz.step(2)
lib = _lib.lib(); st = C.c_void_p(0); p = mr._p
def T(name, f, reps=3):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); print(f"{name:28s} {(time.perf_counter()-t)/reps*1e3:8.2f} ms")
T("sim_counts (scene + dark)", lambda: _lib.check(lib.rip_sim_counts_dev(cd.handle, p(rz.d_image), p(rz.d_area_act), 0, rz.t_exp, 1.0, float(pars.g_ideal), rz.t_exp, 5, p(rz.d_counts), 0, st)))
T("sim_counts (scene only)", lambda: _lib.check(lib.rip_sim_counts_dev(cd.handle, p(rz.d_image), p(rz.d_area_act), 0, rz.t_exp, 1.0, float(pars.g_ideal), 0.0, 5, p(rz.d_counts), 0, st)))
prm = s2i.fwd_params(rp, 5)
T("make_l1", lambda: _lib.check(lib.rip_make_l1_dev(cd.handle, p(rz.d_counts), C.byref(prm), p(rz.d_res), st)))
T("l1_embed", lambda: _lib.check(lib.rip_l1_embed_dev(0, p(rz.d_res), rz.G, n, 4, p(rz.d_im), st)))
T("fill_refdata_1f", lambda: _lib.check(lib.rip_fill_refdata_1f_dev(cd.handle, p(rz.d_im), p(rz.d_amp33), rz.G, _lib.ptr(rz.rpg), 5, 1, st)))
T("l1_to_l2", lambda: gci.calibrate_device(cd, rz.dplan, rz.d_im.data_ptr(), rz.d_amp33.data_ptr(), rz.d_area_full.data_ptr(), rz.d_slope.data_ptr(), rz.d_er.data_ptr(), rz.d_ep.data_ptr(), rz.d_pdq.data_ptr()))
T("moments_accumulate", lambda: _lib.check(lib.rip_moments_accumulate_dev(0, p(rz.d_slope), p(rz.d_pdq), n, 4, _lib.ptr(rz.grow), p(rz.d_moments), st)))
T("whole step", lambda: rz.step(9))
