"""Minimal FITS writer / reader for the side products of the path (astropy is not a dependency of this package).

The reference writes two small FITS products with ``astropy.io.fits``: the mask image of
``maskhandling.CombinedMask.convert_file`` (utils/maskhandling.py:145-149: a float32 primary image and an int8 ``MASK``
extension) and the ``FITSOUT`` copy of the noise cube (L1_to_L2/gen_noise_image.py:386-390).  This module writes the same
layouts: 2880-byte blocks, 80-column cards, big-endian data, ``BITPIX = 8`` with ``BZERO = -128`` for int8 and ``BITPIX = 32``
with ``BZERO = 2^31`` for uint32 as the FITS standard (and astropy) represent them.  ``calibrateimage``'s ``FITSOUT`` product
(L1_to_L2/gen_cal_image.py:725-736: data, dq, masked data) uses it too.
"""

import numpy as np

_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20}"
    elif isinstance(value, (float, np.floating)):
        v = f"{float(value)!r:>20}"
    else:
        v = f"'{str(value):<8}'"
    c = f"{key:<8}= {v}"
    if comment:
        c += f" / {comment}"
    return c.ljust(80)[:80]


def _hdu_bytes(array, header, primary):
    a = np.asarray(array)
    extra = {}
    if a.dtype == np.int8:  # signed bytes: stored as unsigned with an offset
        a = (a.astype(np.int16) + 128).astype(np.uint8)
        extra = {"BZERO": -128, "BSCALE": 1}
    elif a.dtype == np.uint32:  # unsigned 32-bit: signed with BZERO = 2^31, as astropy writes the dq plane
        a = (a.astype(np.int64) - 2147483648).astype(np.int32)
        extra = {"BZERO": 2147483648, "BSCALE": 1}
    elif a.dtype == np.uint16:
        a = (a.astype(np.int32) - 32768).astype(np.int16)
        extra = {"BZERO": 32768, "BSCALE": 1}
    elif a.dtype == np.bool_:
        a = a.astype(np.uint8)
    kind = a.dtype.kind + str(a.dtype.itemsize)
    if kind not in _BITPIX:
        raise TypeError(f"fits_lite: unsupported dtype {a.dtype}")
    cards = [_card("SIMPLE", True, "conforms to FITS standard")] if primary else [f"{'XTENSION':<8}= {chr(39)}IMAGE   {chr(39)}".ljust(80)]
    cards += [_card("BITPIX", _BITPIX[kind]), _card("NAXIS", a.ndim)]
    cards += [_card(f"NAXIS{i + 1}", d) for i, d in enumerate(a.shape[::-1])]
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards += [_card("PCOUNT", 0), _card("GCOUNT", 1)]
    for k, v in {**extra, **(header or {})}.items():
        cards.append(_card(k, v))
    cards.append("END".ljust(80))
    head = "".join(cards)
    head += " " * (-len(head) % 2880)
    data = np.ascontiguousarray(a, dtype=a.dtype.newbyteorder(">")).tobytes()
    return head.encode("ascii") + data + b"\0" * (-len(data) % 2880)


def write_hdus(path, hdus):
    """``hdus``: list of ``(array, header dict or None)``; the first is the primary HDU, the others IMAGE extensions."""
    with open(path, "wb") as f:
        for i, (a, h) in enumerate(hdus):
            f.write(_hdu_bytes(a, h, i == 0))


def read_hdus(path):
    """Inverse of ``write_hdus`` for files of this module's own layouts: list of ``(array, header dict)``."""
    raw = open(path, "rb").read()
    out, pos = [], 0
    while pos < len(raw):
        hdr = {}
        while True:
            block = raw[pos : pos + 2880].decode("ascii")
            pos += 2880
            done = False
            for i in range(0, 2880, 80):
                c = block[i : i + 80]
                if c.startswith("END"):
                    done = True
                    break
                if c[8:10] == "= ":
                    v = c[10:].split("/")[0].strip()
                    hdr[c[:8].strip()] = v.strip("'").strip() if v.startswith("'") else (v == "T" if v in ("T", "F") else (float(v) if any(ch in v for ch in ".eE") else int(v)))
            if done:
                break
        shape = tuple(int(hdr[f"NAXIS{i}"]) for i in range(int(hdr["NAXIS"]), 0, -1))
        dt = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}[int(hdr["BITPIX"])]
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize if shape else 0
        a = np.frombuffer(raw[pos : pos + nbytes], dtype=dt).reshape(shape)
        pos += nbytes + (-nbytes % 2880)
        if hdr.get("BZERO") == -128 and dt == ">u1":
            a = (a.astype(np.int16) - 128).astype(np.int8)
        elif hdr.get("BZERO") == 2147483648 and dt == ">i4":
            a = (a.astype(np.int64) + 2147483648).astype(np.uint32)
        elif hdr.get("BZERO") == 32768 and dt == ">i2":
            a = (a.astype(np.int32) + 32768).astype(np.uint16)
        else:
            a = a.astype(np.dtype(dt).newbyteorder("="))
        out.append((a, hdr))
    return out


def header_text(header, comment=None):
    """80-column card stream padded to 2880 bytes, as ``astropy.io.fits.Header.tofile`` writes a header (what the
    ``FITSWCS`` files of the reference hold, from_sim/sim_to_isim.py:986-987)."""
    cards = [_card(k, v) for k, v in header.items() if k not in ("COMMENT", "HISTORY", "END", "")]
    if comment:
        cards.append(f"COMMENT {comment}".ljust(80)[:80])
    cards.append("END".ljust(80))
    text = "".join(cards)
    return text + " " * (-len(text) % 2880)


def read_primary(path):
    """Primary HDU of any simple FITS image file: ``(array in native byte order, header dict in card order)``."""
    from ..utils.coordutils import parse_header  # noqa: PLC0415

    raw = open(path, "rb").read()
    pos, text = 0, ""
    while True:
        block = raw[pos : pos + 2880].decode("ascii", errors="replace")
        pos += 2880
        text += block
        if any(block[i : i + 8] == "END     " for i in range(0, 2880, 80)):
            break
        if pos >= len(raw):
            raise ValueError(f"{path}: no END card in the primary header")
    hdr = parse_header(text)
    shape = tuple(int(hdr[f"NAXIS{i}"]) for i in range(int(hdr["NAXIS"]), 0, -1))
    dt = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}[int(hdr["BITPIX"])]
    n = int(np.prod(shape)) if shape else 0
    a = np.frombuffer(raw, dtype=dt, count=n, offset=pos).reshape(shape)
    a = a.astype(np.dtype(dt).newbyteorder("="))
    if "BSCALE" in hdr or "BZERO" in hdr:
        bs, bz = float(hdr.get("BSCALE", 1.0)), float(hdr.get("BZERO", 0.0))
        if bs != 1.0 or bz != 0.0:
            a = a * bs + bz
    return a, hdr
