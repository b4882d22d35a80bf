"""Generate golden input/output vectors by running the UNMODIFIED reference modules.

Run HERE (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden.py

The reference's hot-path modules (utils/ipc_linearity.py, utils/fitting.py, utils/flatutils.py,
utils/reference_subtraction.py) import only numpy, ``asdf`` and ``roman_datamodels.dqflags.pixel``.  Neither
third-party package is installed here, so two stub modules are pre-seeded in ``sys.modules``:
``asdf.open(name)`` returns a context manager over an in-memory tree registered under ``name`` and
``roman_datamodels.dqflags.pixel`` is a class with ``np.uint32`` members.  The reference source is imported from
where it lies; nothing is copied.  Inputs come from ``romanimpreprocess_b200.synth`` (seeded); the outputs of
the reference functions are written to ``tests/golden/*.npz`` together with a digest of the inputs.
"""

import contextlib
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"

TREES = {}


def install_stubs():
    asdf = types.ModuleType("asdf")

    @contextlib.contextmanager
    def _open(name, *a, **k):
        yield TREES[name]

    asdf.open = _open
    sys.modules["asdf"] = asdf
    rdm = types.ModuleType("roman_datamodels")
    dqf = types.ModuleType("roman_datamodels.dqflags")
    from romanimpreprocess_b200.dqflags import pixel  # plain class with np.uint32 members

    dqf.pixel = pixel
    rdm.dqflags = dqf
    sys.modules["roman_datamodels"] = rdm
    sys.modules["roman_datamodels.dqflags"] = dqf
    sys.path.insert(0, REF_SRC)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


class Log:
    def __init__(self):
        self.output = ""

    def append(self, s):
        self.output += s


def register(cal, tag):
    names = {}
    for k, v in cal.items():
        names[k] = f"{tag}:{k}"
        TREES[names[k]] = v
    return names


def small_case(tag, n, read_pattern, p_order, gain_dtype, ipc_dtype, seed, bright=1.0):
    """Run every size-agnostic reference function on one small synthetic case."""
    from romanimpreprocess.utils import fitting, flatutils, ipc_linearity
    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.dqflags import pixel

    cal = synth.make_caldir(n=n, seed=seed, read_pattern=read_pattern, p_order=p_order, gain_dtype=gain_dtype,
                            ipc_dtype=ipc_dtype, sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data_u16, amp33_u16, meta = synth.make_l1(cal, read_pattern, seed=seed + 1, n_sources=9, cr_frac=0.01, bright=bright)
    names = register(cal, tag)
    out = {}
    out["input_digest"] = np.array(
        digest(data_u16, cal["linearitylegendre"]["roman"]["data"], cal["gain"]["roman"]["data"],
               cal["ipc4d"]["roman"]["data"], cal["read"]["roman"]["data"])  # fmt: skip
    )
    G = len(read_pattern)
    S = data_u16.astype(np.float32)
    # a saturation-like group flag cube to exercise attempt_corr and the truncated fits
    sat = cal["saturation"]["roman"]["data"]
    rdq0 = np.zeros((G, n, n), dtype=np.uint8)
    cum = np.zeros((n, n), dtype=bool)
    for g in range(1, G):
        cum |= S[g] >= sat
        rdq0[g] |= np.where(cum, np.uint8(2), np.uint8(0))
    rdq0[0] |= np.uint8(1)
    out["rdq0"] = rdq0

    # multilin (ipc_linearity.py:276)
    phi, dq = ipc_linearity.multilin(
        S, names["linearitylegendre"], do_not_flag_first=True, attempt_corr=~rdq0 & pixel.SATURATED
    )
    out["multilin_phi"] = phi
    out["multilin_dq"] = dq
    phi_b, dq_b = ipc_linearity.multilin(S, names["linearitylegendre"], do_not_flag_first=False)
    out["multilin_phi_flagfirst"] = phi_b
    out["multilin_dq_flagfirst"] = dq_b
    # linearity single frame (ipc_linearity.py:234), with an origin offset
    p1, d1 = ipc_linearity.linearity(S[2, 5:25, 7:30], names["linearitylegendre"], origin=(7, 5))
    out["linearity_phi"] = p1
    out["linearity_dq"] = d1

    # IPC (ipc_linearity.py:37,102,145)
    K = cal["ipc4d"]["roman"]["data"]
    g_act = cal["gain"]["roman"]["data"][4:-4, 4:-4]
    img = phi[3, 4:-4, 4:-4].copy()
    out["ipc_fwd"] = ipc_linearity.ipc_fwd(img, K)
    out["ipc_fwd_gain"] = ipc_linearity.ipc_fwd(img, K, gain=g_act)
    out["ipc_rev"] = ipc_linearity.ipc_rev(img, K)
    out["ipc_rev_gain"] = ipc_linearity.ipc_rev(img, K, gain=g_act)
    out["ipc_rev_order3"] = ipc_linearity.ipc_rev(img, K, order=3)
    cube = phi.copy()
    ipc_linearity.correct_cube(cube, names["ipc4d"], Log(), gain_file=names["gain"])
    out["correct_cube"] = cube
    cube_e = phi.copy()
    ipc_linearity.correct_cube(cube_e, names["ipc4d"], None)
    out["correct_cube_nogain"] = cube_e

    # ramp fit (fitting.py:20,89,258)
    m = dict(meta)
    m["K"] = fitting.construct_weights(0.4 / 1.8 / 7.0**2, m, exclude_first=True)
    m["jump_detect_pars"] = {"SthreshA": 10.0, "SthreshB": 4.5, "IthreshA": 0.6, "IthreshB": 600.0}
    out["K"] = m["K"]
    caldir = {"gain": names["gain"], "read": names["read"]}
    pdq = cal["mask"]["roman"]["dq"].copy() | dq
    rdq = np.zeros_like(rdq0)
    s, er, ep, smap = fitting.jump_detect(cube, rdq, pdq, m, caldir, Log(), exclude_first=True)
    out["jd_slope"], out["jd_err_read"], out["jd_err_poisson"], out["jd_smap"], out["jd_rdq"] = s, er, ep, smap, rdq
    rdq = np.zeros_like(rdq0)
    s, er, ep, smap = fitting.jump_detect(cube, rdq, pdq, m, caldir, Log(), exclude_first=True, truncate_ramp=G - 1)
    out["jdt_slope"], out["jdt_err_read"], out["jdt_err_poisson"], out["jdt_smap"], out["jdt_rdq"] = (
        s, er, ep, smap, rdq)  # fmt: skip
    rdq = rdq0.copy()
    pdq_rf = pdq.copy()
    s, er, ep = fitting.ramp_fit(cube, rdq, pdq_rf, m, caldir, Log(), exclude_first=True)
    out["rf_slope"], out["rf_err_read"], out["rf_err_poisson"], out["rf_rdq"], out["rf_pdq"] = s, er, ep, rdq, pdq_rf
    # default thresholds, exclude_first False
    m2 = dict(meta)
    m2["K"] = fitting.construct_weights(0.4 / 1.8 / 6.5**2, m2, exclude_first=False)
    rdq = rdq0.copy()
    rdq[0] = 0
    pdq_rf2 = pdq.copy()
    s, er, ep = fitting.ramp_fit(cube, rdq, pdq_rf2, m2, caldir, Log(), exclude_first=False)
    out["rf2_K"] = m2["K"]
    out["rf2_slope"], out["rf2_err_read"], out["rf2_err_poisson"], out["rf2_rdq"], out["rf2_pdq"] = (
        s, er, ep, rdq, pdq_rf2)  # fmt: skip

    # flat (flatutils.py:20)
    pdq_f = cal["mask"]["roman"]["dq"].copy()
    fl = flatutils.get_flat({"flat": names["flat"], "gain": names["gain"], "ipc4d": names["ipc4d"]}, {"nborder": 4}, pdq_f)
    out["flat"], out["flat_pdq"] = fl, pdq_f
    out["flat_noipc"] = flatutils.get_flat({"flat": names["flat"]}, {"nborder": 4}, None, ipc_deconvolve=False)

    # inverse linearity + IL.apply (ipc_linearity.py:347,398)
    Slin = phi[2, 4:-4, 4:-4].astype(np.float64)
    Sinv, ex = ipc_linearity.invlinearity(Slin, names["linearitylegendre"], origin=(4, 4))
    out["invlin_S"], out["invlin_ex"] = Sinv, ex
    Sinv32, _ = ipc_linearity.invlinearity(phi[2, 4:-4, 4:-4], names["linearitylegendre"], origin=(4, 4))
    out["invlin_S_f32"] = Sinv32
    rng = np.random.RandomState(seed + 5)
    counts = rng.poisson(lam=3000.0, size=(n - 8, n - 8)).astype(np.int32)
    start_e = (rng.normal(size=(n - 8, n - 8)) * 40.0).astype(np.float32)
    out["il_counts"], out["il_start_e"] = counts, start_e
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        il = ipc_linearity.IL(names["linearitylegendre"], names["gain"], names["ipc4d"], start_e=start_e)
        il.set_dq(ngroup=G, nborder=4)
        out["il_apply"] = il.apply(counts, electrons=True)
        out["il_apply_eout"] = il.apply(counts, electrons=True, electrons_out=True)
    out["il_dq"] = il.dq
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print("wrote", tag, {k: getattr(v, "shape", None) for k, v in out.items()})


def refsub_case():
    """ref_subtraction_row / _channel need the real 4096x4224 geometry (hard-coded in the reference)."""
    from romanimpreprocess.utils import reference_subtraction as rs

    rng = np.random.RandomState(4242)
    im = (rng.normal(size=(4096, 4224)) * 6.0).astype(np.float32)
    im += (4.0 * np.sin(np.arange(4096) / 50.0)).astype(np.float32)[:, None]
    im[:, 4096:] *= 0.8
    im += (np.arange(4224) // 128).astype(np.float32)[None, :] * 0.37
    slope = np.float64(0.43210987654321)
    out = {"seed": np.array(4242), "slope": np.array(slope)}
    a = rs.ref_subtraction_row(im.copy(), use_ref_channel=True, slope=slope)
    out["row_digest"] = np.array(digest(a))
    out["row_sample"] = a[::97, ::89].copy()
    b = rs.ref_subtraction_channel(a.copy(), use_ref_channel=True)
    out["chan_digest"] = np.array(digest(b))
    out["chan_sample"] = b[::97, ::89].copy()
    c = rs.ref_subtraction_row(im.copy(), use_ref_channel=False)  # polyfit branch (tests/.../test_ref.py)
    out["rowfit_digest"] = np.array(digest(c))
    out["rowfit_sample"] = c[::97, ::89].copy()
    d = rs.ref_subtraction_channel(im.copy(), use_ref_channel=False)
    out["chan32_digest"] = np.array(digest(d))
    np.savez_compressed(os.path.join(HERE, "refsub_4096.npz"), **out)
    print("wrote refsub_4096")


def kats():
    """Literal known answers held by the reference's own tests / sources."""
    out = {}
    # tests/romanimpreprocess/test_linutils.py:14-48
    out["lin_p3_z"] = np.linspace(-1.5, 1.5, 31)
    out["lin_p3_phi"] = np.array(
        [-4.0, -3.4, -2.8, -2.2, -1.6, -1.0, -0.4725, -0.08, 0.1925, 0.36, 0.4375, 0.44, 0.3825, 0.28, 0.1475, 0.0,
         -0.1475, -0.28, -0.3825, -0.44, -0.4375, -0.36, -0.1925, 0.08, 0.4725, 1.0, 1.6, 2.2, 2.8, 3.4, 4.0]
    )  # fmt: skip
    # tests/romanimpreprocess/test_workflow.py:402-407 (IL.apply on the gencal fixture, pixels [260:262,140:143])
    out["il_target1"] = np.array(
        [[4801.0491668, 4900.74928657, 4800.50198393], [4800.30217909, 4900.15392476, 4800.05504147]]
    )
    out["il_target2"] = np.array(
        [[4803.76066256, 4920.3284374, 4803.19832938], [4817.8237426, 6177.69747299, 4817.69985963]]
    )
    # src/romanimpreprocess/L1_to_L2/denoise_construct.py:219-230 (weights for the README table, u=0.4/1.8/7^2)
    out["weights_readme"] = np.array(
        [0.0, -2.1521233e-03, -3.6145949e-03, -6.7949751e-03, 3.0364664e-10, 6.7949742e-03, 3.6145954e-03,
         2.1521233e-03]
    )  # fmt: skip
    # run the reference's own construct_weights for the same table: must print the literal
    from romanimpreprocess.utils import fitting
    from romanimpreprocess_b200 import synth

    meta = synth.meta_from_pattern(synth.README_PATTERN)
    out["weights_readme_ref"] = fitting.construct_weights(0.4 / 1.8 / 7.0**2, meta, exclude_first=True)
    out["weights_readme_ref_noexcl"] = fitting.construct_weights(0.4 / 1.8 / 6.5**2, meta, exclude_first=False)
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **out)
    print("wrote kats", out["weights_readme_ref"])


if __name__ == "__main__":
    install_stubs()
    from romanimpreprocess_b200 import synth

    kats()
    small_case("small_p4_f32", 40, synth.TEST_READ_PATTERN, 3, np.float32, np.float32, 11)
    small_case("small_p11_f32", 40, synth.README_PATTERN, 10, np.float32, np.float32, 12)
    small_case("small_p4_g64", 40, synth.TEST_READ_PATTERN, 3, np.float64, np.float32, 13)
    small_case("small_p11_k64", 40, synth.README_PATTERN, 10, np.float32, np.float64, 14)
    # bright sources: pixels saturate in every group, so each truncated refit of ramp_fit (fitting.py:326-337) fires
    small_case("small_sat_f32", 56, synth.README_PATTERN, 10, np.float32, np.float32, 15, bright=12.0)
    small_case("small_sat_g64k64", 56, synth.TEST_READ_PATTERN, 3, np.float64, np.float64, 16, bright=12.0)
    refsub_case()
