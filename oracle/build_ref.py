"""Recipe for ``oracle/_ref``: the reference's own hot-path modules, UNMODIFIED, where the GPU box can import them.

TEST INFRASTRUCTURE ONLY (like everything under ``oracle/``): nothing under ``romanimpreprocess_b200/`` imports it.

    python oracle/build_ref.py          # here, in the container that has /root/reference

copies ``utils/{ipc_linearity,fitting,flatutils,reference_subtraction}.py``, ``pars.py`` and empty package markers
from ``/root/reference/src/romanimpreprocess`` into ``oracle/_ref/romanimpreprocess`` byte for byte (git-ignored, so no
reference source enters the history; not gpurun-ignored, so the directory travels to the GPU box with the snapshot) and
writes ``oracle/_ref/MANIFEST.json`` with the sha256 of every copied file.  These four modules import only numpy,
``asdf`` and ``roman_datamodels.dqflags.pixel``; ``oracle/ref_chain.py`` supplies two stub modules for those (an
in-memory ``asdf.open`` and the flag table), exactly as ``tests/golden/make_golden.py`` does.  Everything else of the
reference (gen_cal_image.py and its romancal / stcal / roman_datamodels / gwcs imports) cannot be imported in this image.
"""

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/romanimpreprocess"
FILES = ["pars.py", "utils/ipc_linearity.py", "utils/fitting.py", "utils/flatutils.py", "utils/reference_subtraction.py"]


def build(ref=REF, out=os.path.join(HERE, "_ref")):
    if not os.path.isdir(ref):
        return False
    pkg = os.path.join(out, "romanimpreprocess")
    os.makedirs(os.path.join(pkg, "utils"), exist_ok=True)
    manifest = {}
    for rel in FILES:
        dst = os.path.join(pkg, rel)
        shutil.copyfile(os.path.join(ref, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    for d in (pkg, os.path.join(pkg, "utils")):  # package markers: empty files (the reference's __init__ imports nothing we need)
        open(os.path.join(d, "__init__.py"), "w").close()
    with open(os.path.join(out, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "files": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref built" if ok else "no /root/reference here: nothing built")
    sys.exit(0 if ok else 1)
