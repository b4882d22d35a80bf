"""development tool: wall-clock break-down of one noise layer at 4096^2 (run under gpurun)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from romanimpreprocess_b200 import _lib, synth
from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci, gen_noise_image as gni
from romanimpreprocess_b200.utils import sky

rp = synth.README_PATTERN
cal, exposures, area = bench.make_inputs(4096, rp, 2, seed=1000)
cd = gci.CalDir(cal, device=0)
nl = gni.NoiseLayers(cd, rp, synth.FRAME_TIME, {"SKYORDER": 2}, device=0)
dev = nl.dev
d_data = torch.from_numpy(exposures[0][0].view(np.int16)).to(dev).view(torch.uint16)
d_amp = torch.from_numpy(exposures[0][1].view(np.int16)).to(dev).view(torch.uint16)
d_area = torch.from_numpy(area).to(dev)
nl.set_exposure(d_data, d_amp, d_area)
nl.layer("Rz4S2C1", 1)
lib = _lib.lib()
st = nl._stream()
def T(name, f, reps=3):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); print(f"{name:28s} {(time.perf_counter()-t)/reps*1e3:8.2f} ms")
T("dark_as_l1", lambda: _lib.check(lib.rip_dark_as_l1_dev(cd.handle, nl.G, gni._ptr(nl.d_work), st)))
T("add_read_noise", lambda: _lib.check(lib.rip_add_read_noise_dev(cd.handle, gni._ptr(nl.d_work), nl.G, _lib.ptr(nl.rpg), 5, st)))
T("fill_refdata_1f", lambda: _lib.check(lib.rip_fill_refdata_1f_dev(cd.handle, gni._ptr(nl.d_work), gni._ptr(nl.d_amp33), nl.G, _lib.ptr(nl.rpg), 5, 1, st)))
T("fill_refdata (no banding)", lambda: _lib.check(lib.rip_fill_refdata_1f_dev(cd.handle, gni._ptr(nl.d_work), gni._ptr(nl.d_amp33), nl.G, _lib.ptr(nl.rpg), 5, 0, st)))
T("calibrate + SKYORDER medfit", lambda: nl._calibrate(nl.d_work, nl.d_amp33, nl.d_area, "noisy"))
nl.config.pop("SKYORDER")
T("calibrate only", lambda: nl._calibrate(nl.d_work, nl.d_amp33, nl.d_area, "noisy"))
T("active_diff", lambda: _lib.check(lib.rip_active_diff_dev(0, gni._ptr(nl.l2["noisy"][0]), gni._ptr(nl.l2["ref"][0]), nl.n, 4, gni._ptr(nl.d_diff), st)))
T("percentiles x3", lambda: sky.percentiles_device(nl.d_diff.data_ptr(), nl.na * nl.na, (25, 50, 75)))
T("medfit_device", lambda: sky.medfit_device(nl.d_diff.data_ptr(), nl.na, nl.na, nl.na, order=2, subtract=True))
T("layer to host", lambda: nl.d_diff.cpu().numpy())
