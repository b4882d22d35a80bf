"""Calibration-tree access.

The reference opens CALDIR entries by *file name* with ``asdf.open`` inside every numerical routine
(e.g. reference src/romanimpreprocess/utils/ipc_linearity.py:170,267,324,380; utils/fitting.py:201,207;
utils/flatutils.py:47,63,71).  The drop-in modules of this package keep those signatures: a CALDIR entry may be

* a path to an ASDF file (opened with the ``asdf`` package where it is installed, exactly as in the reference, else
  with this package's own reader of the ASDF block layout, ``io/asdf_lite.py``), or
* an already-loaded tree (a mapping that has a ``"roman"`` branch) -- used by the tests, the benchmark and by
  callers that keep their CALDIR resident.

Both are opened through :func:`open_tree`, which is a context manager like ``asdf.open``.
"""

import contextlib
import os


def have_asdf():
    """True if the real ``asdf`` package is importable (then files are opened exactly as the reference opens them)."""
    if os.environ.get("RIP_FORCE_ASDF_LITE"):
        return False
    try:
        import asdf  # noqa: F401, PLC0415
    except ImportError:
        return False
    return True


@contextlib.contextmanager
def open_tree(entry):
    """Yield a mapping with a ``"roman"`` branch for a CALDIR entry or an L1 / L2 file (path or in-memory tree).

    Paths are opened with ``asdf.open`` where that package exists, else with the package's own block reader
    (``io.asdf_lite``: the YAML tree and the binary blocks, tags passed through un-interpreted)."""
    if isinstance(entry, dict) or hasattr(entry, "keys"):
        yield entry
        return
    if have_asdf():
        import asdf  # noqa: PLC0415

        with asdf.open(entry) as f:
            yield f
        return
    from .io import asdf_lite  # noqa: PLC0415

    with asdf_lite.open_file(entry) as f:
        yield f


def write_tree(path, tree):
    """Write ``{"roman": ..., "processinfo": ...}`` to an ASDF file (``asdf.AsdfFile.write_to`` where available, else
    ``io.asdf_lite.write_file``: uncompressed internal blocks, same layout)."""
    if have_asdf():
        import asdf  # noqa: PLC0415

        with asdf.AsdfFile() as af:
            af.tree = tree
            with open(path, "wb") as f:
                af.write_to(f)
        return
    from .io import asdf_lite  # noqa: PLC0415

    # Block checksums are optional in ASDF (all-zero = "not computed"; readers skip the check): MD5 of the ~560 MB of L2
    # arrays costs 0.7 s of the 1.2 s a whole calibrateimage() call takes at 4096^2.  RIP_ASDF_CHECKSUM=1 writes them.
    asdf_lite.write_file(path, tree, checksum=os.environ.get("RIP_ASDF_CHECKSUM", "") not in ("", "0"))
