"""Development: cProfile of calibrateimage(config) at full size on tmpfs."""
import cProfile, os, pstats, shutil, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fixture_files import write_exposure
from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
tmp = tempfile.mkdtemp(prefix="rip_prof_", dir="/dev/shm")
try:
    config, *_ = write_exposure(tmp, n=4096, seed=41, p_order=10, ipc_dtype=np.float64)
    gci.calibrateimage(config, verbose=False)
    pr = cProfile.Profile(); pr.enable()
    gci.calibrateimage(config, verbose=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
