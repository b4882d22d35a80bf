"""CPU: the ASDF block reader / writer of the host I/O pipeline (romanimpreprocess_b200/io/asdf_lite.py) and
the pixel-area computation from the FITS WCS (utils/coordutils.py; reference utils/coordutils.py:17-82) against the
reference's own known-answer test (tests/romanimpreprocess/test_area.py) and analytic targets.
GPU (marked): the device version of the pixel area against the NumPy one."""

import bz2
import hashlib
import struct
import zlib

import numpy as np
import pytest

from romanimpreprocess_b200.io import asdf_lite as al
from romanimpreprocess_b200.utils import coordutils as cu


# ---- ASDF --------------------------------------------------------------------------------------------------------
def _tree():
    rng = np.random.default_rng(3)
    return {"roman": {"meta": al.TaggedDict({"exposure": {"read_pattern": [[0], [1, 2]], "frame_time": 3.04},
                                              "instrument": {"detector": "WFI07"}},
                                             tag="asdf://stsci.edu/datamodels/roman/tags/common-1.0.0"),
                      "data": rng.integers(0, 65535, (3, 16, 16)).astype(np.uint16),
                      "amp33": rng.integers(0, 65535, (3, 16, 4)).astype(np.uint16),
                      "gain": rng.normal(1.5, 0.1, (16, 16)).astype(np.float32),
                      "ipc": rng.normal(0, 1, (3, 3, 8, 8)),
                      "dq": rng.integers(0, 2**32, (16, 16), dtype=np.uint64).astype(np.uint32),
                      "half": rng.normal(0, 1, (5,)).astype(np.float16),
                      "empty": np.zeros((0,), np.float32), "t0": 3.04, "flag": True, "name": "x"}}  # fmt: skip


def test_asdf_roundtrip(tmp_path):
    p = tmp_path / "a.asdf"
    tree = _tree()
    n = al.write_file(p, tree)
    assert n == p.stat().st_size
    raw = p.read_bytes()
    assert raw.startswith(b"#ASDF 1.0.0\n#ASDF_STANDARD") and b"%YAML 1.1" in raw[:80] and b"#ASDF BLOCK INDEX" in raw
    with al.open_file(p) as f:
        f.verify_checksums()
        r = f["roman"]
        for k in ("data", "amp33", "gain", "ipc", "dq", "half", "empty"):
            a = np.asarray(r[k])
            assert a.dtype == tree["roman"][k].dtype and np.array_equal(a, tree["roman"][k]), k
        assert r["t0"] == 3.04 and r["flag"] is True and r["name"] == "x"
        assert r["meta"]["exposure"]["read_pattern"] == [[0], [1, 2]]
        assert r["meta"].tag == "asdf://stsci.edu/datamodels/roman/tags/common-1.0.0"  # unknown tags survive
        assert np.array_equal(r["data"][1, 2:5], tree["roman"]["data"][1, 2:5])  # leading-axis slicing (gen_cal_image.py:535)
        assert "amp33" in r and r["data"].shape == (3, 16, 16) and len(r["data"]) == 3
        # streaming into a caller buffer (the pinned staging path)
        buf = np.zeros((2, 16, 16), np.uint16)
        assert r["data"].read_into(buf, first=1, count=2) == buf.nbytes
        assert np.array_equal(buf, tree["roman"]["data"][1:])
        with pytest.raises(al.AsdfLiteError):
            r["data"].read_into(np.zeros(3, np.uint16))
        # second generation: what we read can be written again, tags included
        p2 = tmp_path / "b.asdf"
        al.write_file(p2, f.tree)
    with al.open_file(p2) as g:
        assert g["roman"]["meta"].tag == "asdf://stsci.edu/datamodels/roman/tags/common-1.0.0"
        assert np.array_equal(np.asarray(g["roman"]["ipc"]), tree["roman"]["ipc"])


def _handmade(path, payload, compression=b"\0\0\0\0", stored=None, checksum=True, big_endian=False, pad=0):
    """An ASDF file laid out by hand from the standard (independent of write_file): one float32 [2,3] array."""
    stored = payload if stored is None else stored
    head = (b"#ASDF 1.0.0\n#ASDF_STANDARD 1.5.0\n%YAML 1.1\n%TAG ! tag:stsci.edu:asdf/\n--- !core/asdf-1.1.0\n"
            b"asdf_library: !core/software-1.0.0 {author: x, name: y, version: 1.0}\n"
            b"roman:\n  data: !core/ndarray-1.0.0\n    source: 0\n    datatype: float32\n    byteorder: "
            + (b"big" if big_endian else b"little") + b"\n    shape: [2, 3]\n  inline: !core/ndarray-1.0.0\n    data: [1, 2, 3]\n"
            b"    datatype: int16\n    shape: [3]\n...\n")  # fmt: skip
    md5 = hashlib.md5(stored).digest() if checksum else b"\0" * 16  # noqa: S324
    hdr = struct.pack(">I4sQQQ", 0, compression, len(stored) + pad, len(stored), len(payload)) + md5
    with open(path, "wb") as f:
        f.write(head + b"\xd3BLK" + struct.pack(">H", len(hdr)) + hdr + stored + b"\0" * pad)


@pytest.mark.parametrize("comp", ["none", "zlib", "bzp2", "padded", "big"])
def test_asdf_reads_files_laid_out_from_the_standard(tmp_path, comp):
    a = np.arange(6, dtype=np.float32).reshape(2, 3) * 1.5
    p = tmp_path / "h.asdf"
    if comp == "none":
        _handmade(p, a.tobytes())
    elif comp == "padded":
        _handmade(p, a.tobytes(), pad=13, checksum=False)
    elif comp == "big":
        _handmade(p, a.astype(">f4").tobytes(), big_endian=True)
    elif comp == "zlib":
        _handmade(p, a.tobytes(), b"zlib", zlib.compress(a.tobytes()))
    else:
        _handmade(p, a.tobytes(), b"bzp2", bz2.compress(a.tobytes()))
    with al.open_file(p) as f:
        f.verify_checksums()
        assert np.array_equal(np.asarray(f["roman"]["data"]), a)
        assert np.array_equal(f["roman"]["inline"], np.array([1, 2, 3], np.int16))
        assert f["asdf_library"]["name"] == "y"
        out = np.zeros((2, 3), np.float32)
        f["roman"]["data"].read_into(out)
        assert np.array_equal(out, a)


def test_asdf_errors(tmp_path):
    p = tmp_path / "x.asdf"
    p.write_bytes(b"not asdf at all")
    with pytest.raises(al.AsdfLiteError, match="not an ASDF file"):
        al.open_file(p)
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    _handmade(p, a.tobytes(), b"lz4\0", a.tobytes())
    with al.open_file(p) as f:
        with pytest.raises(al.AsdfLiteError, match="not supported"):
            np.asarray(f["roman"]["data"])
    _handmade(p, a.tobytes())
    raw = bytearray(p.read_bytes())
    raw[-3] ^= 0xFF
    p.write_bytes(bytes(raw))
    with al.open_file(p) as f:
        with pytest.raises(al.AsdfLiteError, match="checksum"):
            f.verify_checksums()


# ---- pixel area --------------------------------------------------------------------------------------------------
def test_area_reference_known_answer():
    """tests/romanimpreprocess/test_area.py of the reference, unchanged but for the WCS container: STG projection in both
    hemispheres against the analytic solid angle, |log ratio| < 2e-4; wrong object -> ValueError("Unrecognized WCS type")."""
    for i in range(2):
        N, d = 2000, 0.01
        h = {"CTYPE1": "RA---STG", "CTYPE2": "DEC--STG", "CRPIX1": N / 2.0 + 0.5, "CRPIX2": N / 2.0 + 0.5, "CDELT1": -d,
             "CDELT2": d, "CRVAL1": 25.0, "CRVAL2": 83.0 * (1.0 - 2.0 * i)}  # fmt: skip
        area = cu.pixelarea(h, N=N)
        s = d * (np.linspace(0, N - 1, N) - N / 2.0 - 0.5) * np.pi / 180.0
        x, y = np.meshgrid(s, s)
        area_target = (d * np.pi / 180.0) ** 2 / (1.0 + (x**2 + y**2) / 4.0) ** 2
        assert np.all(np.abs(np.log(area / area_target)) < 2.0e-4)
    with pytest.raises(ValueError, match="Unrecognized WCS type"):
        cu.pixelarea(42, N=64)
    with pytest.raises(ValueError, match="Unrecognized WCS type"):
        cu.pixelarea({"CTYPE1": "RA---ZEA", "CTYPE2": "DEC--ZEA"}, N=64)


SIM_HEADER = {"CTYPE1": "RA---TAN-SIP", "CTYPE2": "DEC--TAN-SIP", "CRPIX1": (4088 + 1) / 2.0, "CRPIX2": (4088 + 1) / 2.0,
              "CD1_1": 3.0555555555555554e-05, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 3.0555555555555554e-05, "CRVAL1": 37.0,
              "CRVAL2": -20.0, "LONPOLE": 215.0, "A_ORDER": 2, "A_0_2": 2.0e-6, "A_1_1": -1.0e-6, "A_2_0": 3.0e-6,
              "B_ORDER": 2, "B_0_2": 1.4e-5, "B_1_1": -1.0e-5, "B_2_0": 3.0e-7}  # (the reference fixture: test_workflow.py:62-83)


def header_cards(h):
    """80-column card stream as astropy's Header.tofile writes it (no line breaks)."""
    cards = []
    for k, v in h.items():
        val = f"'{v:<8}'" if isinstance(v, str) else (f"{v:>20}" if isinstance(v, int) else f"{v!r:>20}")
        cards.append(f"{k:<8}= {val} / comment".ljust(80)[:80])
    cards.append("COMMENT truth wcs from sim_to_isim".ljust(80))
    cards.append("END".ljust(80))
    return "".join(cards)


def test_area_tan_sip_analytic_and_header_text(tmp_path):
    """Gnomonic projection + SIP: solid angle = |det CD| |det d(U,V)/d(u,v)| / (1 + xi^2 + eta^2)^(3/2) exactly."""
    N = 600
    h = dict(SIM_HEADER, CRPIX1=(N + 1) / 2.0, CRPIX2=(N + 1) / 2.0, CD1_2=1.0e-6, CD2_1=-2.0e-6)
    fn = tmp_path / "sim_asdf_wcshead.txt"
    fn.write_text(header_cards(h))
    w = cu.wcs_from_config({"FITSWCS": str(fn)})
    assert w.proj == "TAN" and w.lonpole == 215.0 and w.a[2, 0] == 3.0e-6 and w.b[1, 1] == -1.0e-5
    assert cu.wcs_from_config({}) is None
    area = cu.pixelarea(w, N=N)
    xx, yy = np.meshgrid(np.arange(N, dtype=float), np.arange(N, dtype=float))
    u, v = xx + 1 - w.crpix[0], yy + 1 - w.crpix[1]
    f = 3.0e-6 * u * u - 1.0e-6 * u * v + 2.0e-6 * v * v
    g = 3.0e-7 * u * u - 1.0e-5 * u * v + 1.4e-5 * v * v
    fu, fv = 6.0e-6 * u - 1.0e-6 * v, 4.0e-6 * v - 1.0e-6 * u
    gu, gv = 6.0e-7 * u - 1.0e-5 * v, 2.8e-5 * v - 1.0e-5 * u
    xi = (w.cd[0, 0] * (u + f) + w.cd[0, 1] * (v + g)) * np.pi / 180
    eta = (w.cd[1, 0] * (u + f) + w.cd[1, 1] * (v + g)) * np.pi / 180
    tgt = abs(np.linalg.det(w.cd)) * (np.pi / 180) ** 2 * np.abs((1 + fu) * (1 + gv) - fv * gu) / (1 + xi**2 + eta**2) ** 1.5
    assert np.max(np.abs(area / tgt - 1)) < 1e-8
    # the pointing enters only through rounding: the same detector WCS at another (ra, dec, roll) gives the same area
    area2 = cu.pixelarea(dict(h, CRVAL1=211.0, CRVAL2=64.0, LONPOLE=100.0), N=N)
    assert np.max(np.abs(area2 / area - 1)) < 1e-8
    # round trip through the world coordinates of the reference point
    ra, dec = w.pix2world(w.crpix[0] - 1, w.crpix[1] - 1)
    assert abs(ra - 37.0) < 1e-12 and abs(dec + 20.0) < 1e-12


@pytest.mark.gpu
def test_area_device_matches_numpy():
    """rip_pixel_area_host (CUDA, float64) vs coordutils.pixelarea (NumPy) on the reference fixture's WCS at 4096^2 (the
    size calibrateimage asks for, gen_cal_image.py:619) and on an STG WCS in the northern hemisphere."""
    from romanimpreprocess_b200 import _lib, pars

    for h, N in ((SIM_HEADER, 4096), ({"CTYPE1": "RA---STG", "CTYPE2": "DEC--STG", "CRPIX1": 300.5, "CRPIX2": 280.5,
                                       "CDELT1": -0.01, "CDELT2": 0.01, "CRVAL1": 25.0, "CRVAL2": 83.0}, 512)):  # fmt: skip
        w = cu.FitsWCS(h)
        ref = cu.pixelarea(w, N=N)
        out = cu.pixelarea_device(w, N=N)
        assert out.dtype == np.float64 and np.max(np.abs(out / ref - 1)) < 1e-8
        af = cu.pixelarea_device(w, N=N, inv_omega=1.0 / pars.Omega_ideal, dtype=np.float32)
        assert af.dtype == np.float32 and np.max(np.abs(af / (ref / pars.Omega_ideal) - 1)) < 2e-7
        assert _lib.lib().rip_launch_count() > 0


def test_fits_lite_round_trip(tmp_path):
    """io/fits_lite.py: the two FITS layouts the reference writes (mask image with an int8 MASK extension,
    utils/maskhandling.py:145-149; float32 noise cube, gen_noise_image.py:386-390) -- blocks of 2880 bytes, big-endian
    data, signed bytes as BITPIX 8 + BZERO -128."""
    from romanimpreprocess_b200.io import fits_lite

    rng = np.random.RandomState(4)
    img = rng.randn(37, 53).astype(np.float32)
    msk = (rng.rand(37, 53) < 0.3).astype(np.int8) - (rng.rand(37, 53) < 0.1).astype(np.int8)
    cube = rng.randn(3, 11, 7).astype(np.float32)
    p1, p2 = str(tmp_path / "m.fits"), str(tmp_path / "c.fits")
    fits_lite.write_hdus(p1, [(img, None), (msk, {"EXTNAME": "MASK"})])
    fits_lite.write_hdus(p2, [(cube, None)])
    raw = open(p1, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30] == b"SIMPLE  =                    T" and b"XTENSION= 'IMAGE   '" in raw
    (a0, h0), (a1, h1) = fits_lite.read_hdus(p1)
    assert np.array_equal(a0, img) and h0["BITPIX"] == -32 and h0["NAXIS1"] == 53 and h0["NAXIS2"] == 37
    assert np.array_equal(a1, msk) and a1.dtype == np.int8 and h1["EXTNAME"] == "MASK" and h1["BZERO"] == -128
    ((c0, hc),) = fits_lite.read_hdus(p2)
    assert np.array_equal(c0, cube) and hc["NAXIS"] == 3 and hc["NAXIS3"] == 3
