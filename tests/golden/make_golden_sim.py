"""Golden vectors for the scene / reference-pixel rows of the forward path (a15, a20).

Run HERE (needs /root/reference), never on the GPU box:    python tests/golden/make_golden_sim.py

``from_sim/sim_to_isim.py`` cannot be imported (galsim, romanisim, asdf ... are absent), so the two functions
``noise_1f_frame`` and ``fill_in_refdata_and_1f`` are taken out of the reference file with ``ast`` AT RUN TIME
(nothing is copied into this repository) and executed unmodified in a namespace that provides what they touch:
NumPy, ``copy``, ``warnings``, the ``asdf.open`` stub of make_golden.py, ``pars`` / ``parameters`` objects with a
small frame (nside 256, 32 channels of 8 columns) and a ``galsim.GaussianDeviate`` whose ``generate`` reads from
``oracle.rip_oracle.NormalStream`` -- the same stream class the oracle restatement consumes, so both see identical
draws in identical order and must agree bit for bit.
"""

import ast
import copy
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402

from oracle import rip_oracle as O  # noqa: E402
from romanimpreprocess_b200 import synth  # noqa: E402

REF_FILE = "/root/reference/src/romanimpreprocess/from_sim/sim_to_isim.py"
N, G, SEED = 256, 3, 4242


def reference_functions(nside, channelwidth):
    tree = ast.parse(open(REF_FILE).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("noise_1f_frame", "fill_in_refdata_and_1f")]
    assert len(wanted) == 2
    mod = ast.Module(body=wanted, type_ignores=[])
    galsim = types.SimpleNamespace(GaussianDeviate=lambda stream: stream)  # .generate(array) is the stream's
    ns = {
        "np": np, "copy": copy, "warnings": warnings, "galsim": galsim, "asdf": sys.modules["asdf"],
        "pars": types.SimpleNamespace(nside=nside, channelwidth=channelwidth),
        "parameters": types.SimpleNamespace(nborder=4),
        "print": lambda *a, **k: None,
    }
    exec(compile(mod, REF_FILE, "exec"), ns)
    return ns["noise_1f_frame"], ns["fill_in_refdata_and_1f"]


def small_cal(n, G):
    pattern = [[0], [1, 2], [3, 4, 5, 6]][:G]
    cal = synth.make_caldir(n=n, read_pattern=pattern, p_order=3, seed=SEED)
    cal = {k: v["roman"] for k, v in cal.items()}
    cw = n // 32
    a33 = cal["read"]["amp33"]
    a33["med"] = np.ascontiguousarray(a33["med"][:, :cw])
    a33["std"] = np.ascontiguousarray(a33["std"][:, :cw])
    return cal, pattern


def synth_sky_image(ny, nx, seed):
    """Seeded float32 test image for medfit: smooth gradient + noise + sources + NaN patches (one region all NaN)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    img = (0.7 + 0.3 * xx / nx - 0.2 * (yy / ny) ** 2 + 0.05 * rng.randn(ny, nx)).astype(np.float32)
    img[rng.rand(ny, nx) < 0.02] += 30.0
    img[rng.rand(ny, nx) < 0.05] = np.nan
    img[: ny // 8, : nx // 8] = np.nan
    return img


def main():
    MG.install_stubs()
    ref_frame, ref_fill = reference_functions(N, N // 32)
    cal, pattern = small_cal(N, G)
    names = MG.register({k: {"roman": v} for k, v in cal.items()}, "sim")
    tij = O.read_pattern_to_tij(pattern)
    rng = np.random.RandomState(SEED)
    im0 = rng.randint(0, 60000, size=(G, N, N)).astype(np.uint16)

    # noise_1f_frame alone
    f_ref = ref_frame(O.NormalStream(11))
    f_ora = O.noise_1f_frame(O.NormalStream(11), N, N // 32)
    assert np.array_equal(f_ref, f_ora)

    out = {}
    for tag, banding, with33 in (("full", True, True), ("nobanding", False, True), ("no33", True, False)):
        im_r, im_o = im0.copy(), im0.copy()
        a_r = np.zeros((G, N, N // 32), np.uint16) if with33 else None
        a_o = np.zeros((G, N, N // 32), np.uint16) if with33 else None
        ref_fill(im_r, names, O.NormalStream(77), tij, fill_in_banding=banding, amp33=a_r)
        O.fill_in_refdata_and_1f(im_o, cal, O.NormalStream(77), tij, fill_in_banding=banding, amp33=a_o)
        assert np.array_equal(im_r, im_o), tag
        if with33:
            assert np.array_equal(a_r, a_o), tag
            out[f"{tag}_amp33"] = a_r
        out[f"{tag}_im"] = im_r
    # a15: the calibration planes of Image2D.simulate with the reference's own ipc_rev
    from romanimpreprocess.utils.ipc_linearity import ipc_rev

    nb = 4
    d = cal["dark"]["dark_slope"][nb:-nb, nb:-nb] * cal["gain"]["data"][nb:-nb, nb:-nb]
    g = cal["gain"]["data"][nb:-nb, nb:-nb]
    d = ipc_rev(d, cal["ipc4d"]["data"])
    fl = np.clip(ipc_rev(cal["flat"]["data"][nb:-nb, nb:-nb], cal["ipc4d"]["data"], gain=g), 0.0, 2 - 2**-21)
    d = np.clip(d, -0.1 * fl, None)
    od, ofl, og = O.sim_calprep(cal)
    assert np.array_equal(od, d) and np.array_equal(ofl, fl)
    # CombinedMask.build / PixelMask1 of the unmodified reference (needs an astropy.io.fits stub to import) and the
    # moment sums of validation_tests/many_realizations.py:74-83 executed line by line
    ast_mod, ast_io, ast_fits = types.ModuleType("astropy"), types.ModuleType("astropy.io"), types.ModuleType("astropy.io.fits")
    ast_mod.io, ast_io.fits = ast_io, ast_fits
    sys.modules.update({"astropy": ast_mod, "astropy.io": ast_io, "astropy.io.fits": ast_fits})
    from romanimpreprocess.utils import maskhandling as ref_mh

    mrng = np.random.RandomState(99)
    nm = 96
    dq = np.zeros((nm, nm), np.uint32)
    for bit in range(32):
        hits = mrng.rand(nm, nm) < (0.004 if bit not in (0, 12) else 0.02)
        dq |= np.where(hits, np.uint32(1 << bit), np.uint32(0)).astype(np.uint32)
    dq[0, :7] |= np.uint32(1 << 3)  # a 5x5 grower on the edge
    dq[-1, -1] |= np.uint32(1 << 10)
    ref_mask = ref_mh.PixelMask1.build(dq)
    assert np.array_equal(ref_mask, O.mask_build(dq))
    custom = ref_mh.CombinedMask({"jump_det": 25, "hot": 5, 7: 9, "saturated": 1})
    ref_mask_c = custom.build(dq)
    assert np.array_equal(ref_mask_c, O.mask_build(dq, {2: 25, 11: 5, 7: 9, 1: 1}))
    mom_ref = np.zeros((3, nm, nm), dtype=np.float32)
    mom_ora = np.zeros((3, nm, nm), dtype=np.float32)
    mdata, mdq = [], []
    for j in range(5):
        dat = (mrng.randn(nm, nm) * 3 + 1.5).astype(np.float32)
        dqj = np.where(mrng.rand(nm, nm) < 0.3, dq, np.uint32(0)).astype(np.uint32)
        if j == 0:
            dqj[5:9, 5:9] |= np.uint32(1)  # masked in every realisation -> sentinel
        else:
            dqj[5:9, 5:9] = dqj[5:9, 5:9] | np.uint32(1)
        w = np.logical_not(ref_mh.PixelMask1.build(dqj))
        mom_ref[0, :, :] += np.where(w, 1, 0.0)
        mom_ref[1, :, :] += np.where(w, dat, 0.0)
        mom_ref[2, :, :] += np.where(w, dat**2, 0.0)
        O.moments_accumulate(mom_ora, dat, dqj)
        mdata.append(dat)
        mdq.append(dqj)
    mom_sum = mom_ref.copy()
    mom_ref[1:, :, :] /= mom_ref[0, :, :] + 1e-25
    mom_ref[2, :, :] = np.sqrt(np.clip(mom_ref[2, :, :] - mom_ref[1, :, :] ** 2, 0, None))
    mom_ref[1:, :, :] = np.where(mom_ref[0, :, :][None, :, :] > 0.1, mom_ref[1:, :, :], -1000.0)
    assert np.array_equal(mom_sum, mom_ora)
    O.moments_finalize(mom_ora)
    assert np.array_equal(mom_ref, mom_ora)
    # sky.medfit of the unmodified reference (numpy + scipy only) on a seeded image with NaNs and odd sizes
    from romanimpreprocess.utils import sky as ref_sky

    sky_out = {}
    for tag, (sy, sx, order, nreg) in {"a": (509, 1022, 2, 8), "b": (300, 257, 0, 8), "c": (412, 412, 3, 4)}.items():
        img = synth_sky_image(sy, sx, 40 + ord(tag))
        coef, model = ref_sky.medfit(img, N=nreg, order=order)
        ocoef, omodel, omeds = O.medfit(img, N=nreg, order=order)
        assert np.array_equal(coef, ocoef) and np.array_equal(model, omodel) and model.dtype == np.float32
        sky_out[f"{tag}_coef"], sky_out[f"{tag}_meds"] = coef, omeds
        sky_out[f"{tag}_model_sub"] = model[::7, ::5].copy()
        sky_out[f"{tag}_model_sum"] = np.float64(model.astype(np.float64).sum())
    np.savez_compressed(os.path.join(HERE, "sky_medfit.npz"), **sky_out)
    np.savez_compressed(
        os.path.join(HERE, "mask_moments.npz"), dq=dq, mask_pixelmask1=ref_mask, mask_custom=ref_mask_c,
        data=np.array(mdata), dqs=np.array(mdq), moments_sum=mom_sum, moments_final=mom_ref,
    )
    np.savez_compressed(
        os.path.join(HERE, "sim_refdata_n256.npz"), im0=im0, frame_seed11=f_ref, this_dark=d, this_flat=fl,
        n=N, G=G, seed=SEED, **out,
    )
    print("oracle == reference for noise_1f_frame, fill_in_refdata_and_1f (3 modes), sim_calprep, CombinedMask.build, "
          "moment sums; golden written")


if __name__ == "__main__":
    main()
