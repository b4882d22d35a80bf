"""Calibration-tree access.

The reference opens CALDIR entries by *file name* with ``asdf.open`` inside every numerical routine
(e.g. reference src/romanimpreprocess/utils/ipc_linearity.py:170,267,324,380; utils/fitting.py:201,207;
utils/flatutils.py:47,63,71).  The drop-in modules of this package keep those signatures: a CALDIR entry may be

* a path to an ASDF file (needs the ``asdf`` package, exactly as in the reference), or
* an already-loaded tree (a mapping that has a ``"roman"`` branch) -- used by the tests, the benchmark and by
  callers that keep their CALDIR resident.

Both are opened through :func:`open_tree`, which is a context manager like ``asdf.open``.
"""

import contextlib


@contextlib.contextmanager
def open_tree(entry):
    """Yield a mapping with a ``"roman"`` branch for a CALDIR entry (path or in-memory tree)."""
    if isinstance(entry, dict) or hasattr(entry, "keys"):
        yield entry
        return
    try:
        import asdf  # noqa: PLC0415
    except ImportError as e:  # pragma: no cover
        raise ImportError(
            f"CALDIR entry {entry!r} is a file name but the 'asdf' package is not installed; "
            "pass an in-memory tree instead"
        ) from e
    with asdf.open(entry) as f:
        yield f
