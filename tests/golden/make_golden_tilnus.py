"""Golden vectors of the reference's GalPoisson/find_tilnus.get_tilde_nus (numpy only: imported unmodified from
/root/reference).  Run in the container that has /root/reference:  python tests/golden/make_golden_tilnus.py"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src/romanimpreprocess/L1_to_L2/GalPoisson")
from find_tilnus import get_tilde_nus  # noqa: E402

from romanimpreprocess_b200 import synth  # noqa: E402

out = {}
rng = np.random.RandomState(3)
for name in ("README_PATTERN", "TEST_READ_PATTERN", "LONG16_PATTERN"):
    rp = getattr(synth, name)
    n_beta = np.array([len(g) for g in rp])
    a_beta = np.array([g[0] for g in rp])
    for j in range(3):
        w = rng.randn(len(rp)).astype(np.float32)
        w -= w.mean()
        out[f"{name}_w{j}"] = w
        out[f"{name}_t{j}"] = np.array(get_tilde_nus(n_beta, a_beta, w), dtype=np.float64)
np.savez(os.path.join(HERE, "tilnus.npz"), **out)
print({k: v for k, v in out.items() if "_t" in k})
