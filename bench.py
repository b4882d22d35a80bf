#!/usr/bin/env python
"""Benchmark of the L1->L2 calibration hot path: calibrated SCA frames/s (4096^2 x 8 resultants).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host cores (oracle port)

A "step" is one pass of the hot path over one SCA exposure per GPU: u16 L1 cube [8,4096,4096] + amp33 -> slope,
err_read, err_poisson, pdq, endslice (reference L1_to_L2/gen_cal_image.py:503-629,697-709), order-10 Legendre
linearity (P=11), f32 calibration planes, reference-pixel correction, IPC deconvolution and jump/saturation flagging
on.  Inputs are synthetic (romanimpreprocess_b200/synth.py: the reference's gencal fixture recipe + a synthetic L1).

Reported on ONE JSON line: ``value`` = SCA/s with the exposure already resident in HBM (device-pointer C ABI,
CUDA-event timed, max over ranks); ``e2e`` = SCA/s through the host-buffer API (``calibrate_arrays`` ->
``rip_l1_to_l2_host``) with the H2D/D2H copies from/to pinned memory inside the timed region; ``roofline`` for the
fused kernel (algorithmic bytes B(G,P)*n^2 / its CUDA-event duration vs the measured HBM copy bandwidth);
``cpu_baseline`` = the oracle (NumPy port of the reference) on a bounded sub-frame on the host.
Multi-GPU: one process per GPU (torchrun), SCAs are independent -> no collective on the data path (weak scaling).
"""

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "calibrated SCA frames/sec (4096^2 x 8 resultants)"
UNIT = "SCA/s"
N_SIDE = 4096
P_ORDER = 10


def bytes_per_pixel(G, P, k64=False):
    """Algorithmic HBM bytes per pixel of the fused L1->L2 pass: B(G,P) = 10.0625 G + 4 P + 105; a float64
    ipc4d (the dtype the DUMMY CALDIR builder writes) doubles the 36 B/px of the nine taps."""
    return 10.0625 * G + 4.0 * P + 105.0 + (36.0 if k64 else 0.0)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region: an NVML polling thread (every
    ~2 ms; the timed region of a default run is tens of ms, too short for `nvidia-smi -lms`), nvidia-smi as fallback."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index, uuid=None):
        self.samples = []  # (t, sm_mhz, reason bits)
        self.smax = None
        self.stop_flag = False
        self.mode = None
        self.proc = None
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand if isinstance(cand, bytes) else cand.encode())
                        break
                    except Exception:  # noqa: BLE001
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
                pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")

            def poll():
                while not self.stop_flag:
                    try:
                        self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(get_reasons(h))))  # fmt: skip
                    except Exception:  # noqa: BLE001
                        pass
                    time.sleep(0.002)

            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            self.mode = "nvml"
        except Exception:  # noqa: BLE001
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
                self.t = threading.Thread(target=self._read_smi, daemon=True)
                self.t.start()
                self.mode = "nvidia-smi"
            except Exception:  # noqa: BLE001
                self.mode = None

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                clk, self.smax = float(f[0]), float(f[1])
            except (ValueError, IndexError):
                continue
            bits = 0
            for bit, v in zip((0x8, 0x40, 0x20, 0x4), f[3:7]):
                if v.lower().startswith("active"):
                    bits |= bit
            self.samples.append((time.perf_counter(), clk, bits))

    def stop(self, t0, t1):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        time.sleep(0.03)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        inside = [(c, b) for t, c, b in self.samples if t0 <= t <= t1]
        if not inside:  # region shorter than one sampling period: nearest samples around it
            inside = [(c, b) for t, c, b in self.samples if t0 - 0.05 <= t <= t1 + 0.05]
        bits = 0
        for _, b in inside:
            bits |= b
        reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
        return {"sm_mhz": float(np.median([c for c, _ in inside])) if inside else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(inside), "source": self.mode}  # fmt: skip


def make_inputs(n, read_pattern, n_exposures, seed=1000, ipc_dtype=np.float32):
    """Synthetic CALDIR + exposures; cached under the temp dir so that repeated invocations on one box (plain run, ncu
    launch list, ncu full capture) do not regenerate them (~1 min of NumPy RNG at 4096^2)."""
    import pickle

    from romanimpreprocess_b200 import synth

    cache = os.path.join(tempfile.gettempdir(), f"rip_bench_inputs_n{n}_G{len(read_pattern)}_e{n_exposures}_s{seed}"
                         f"{'_k64' if ipc_dtype == np.float64 else ''}.pkl")
    if os.path.exists(cache):
        try:
            with open(cache, "rb") as f:
                return pickle.load(f)
        except Exception:  # noqa: BLE001
            pass
    res = _make_inputs(n, read_pattern, n_exposures, seed, ipc_dtype)
    try:
        with open(cache + ".tmp", "wb") as f:
            pickle.dump(res, f, protocol=4)
        os.replace(cache + ".tmp", cache)
    except Exception:  # noqa: BLE001
        pass
    return res


def _make_inputs(n, read_pattern, n_exposures, seed, ipc_dtype=np.float32):
    from romanimpreprocess_b200 import synth

    cal = synth.make_caldir(n=n, seed=seed, read_pattern=read_pattern, p_order=P_ORDER, gain_dtype=np.float32,
                            ipc_dtype=ipc_dtype, sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data, amp33, _ = synth.make_l1(cal, read_pattern, seed=200, n_sources=25, cr_frac=1e-3, bright=3.0)
    rng = np.random.default_rng(seed + 1)
    exposures = [(data, amp33)]
    for _ in range(1, n_exposures):  # further exposures: the same scene with fresh +-3 DN noise (cheap to draw)
        d = (data.astype(np.int32) + rng.integers(-3, 4, size=data.shape, dtype=np.int8)).clip(0, 65535).astype(np.uint16)
        a = (amp33.astype(np.int32) + rng.integers(-3, 4, size=amp33.shape, dtype=np.int8)).clip(0, 65535).astype(np.uint16)
        exposures.append((d, a))
    area = synth.make_area_factor(n, np.float32)
    return cal, exposures, area


# ---------------------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: the oracle (NumPy port of the reference) on host cores
# ---------------------------------------------------------------------------------------------------------------
def _oracle_tile(args):
    """One bounded sample: the full oracle L1->L2 chain on an m x m sub-frame (same G, P, flags)."""
    m, seed = args
    from oracle import rip_oracle as orc
    from romanimpreprocess_b200 import synth

    rp = synth.README_PATTERN
    cal = synth.make_caldir(n=m, seed=seed, read_pattern=rp, p_order=P_ORDER, gain_dtype=np.float32,
                            ipc_dtype=np.float32, sprinkle_flags=True, biascorr_amp=3.0)  # fmt: skip
    data, amp33, _ = synth.make_l1(cal, rp, seed=seed + 1, n_sources=25, cr_frac=1e-3, bright=3.0)
    area = synth.make_area_factor(m, np.float32)
    c = {k: v["roman"] for k, v in cal.items()}
    chain = orc.l1_to_l2
    if _ref_available():  # the reference's own unmodified functions (oracle/_ref, see oracle/ref_chain.py)
        from oracle import ref_chain

        chain = ref_chain.l1_to_l2
    t = time.perf_counter()
    chain(data, amp33, c, rp, synth.FRAME_TIME, area, {"SLICEOUT": True}, do_refpix=True)
    return time.perf_counter() - t


def _ref_available():
    try:
        from oracle import ref_chain

        return ref_chain.available()
    except Exception:  # noqa: BLE001
        return False


def _cpu_kind():
    """(kind, description) of what the CPU legs execute."""
    if _ref_available():
        return "reference", ("the reference's own unmodified functions from oracle/_ref (ref_chain.l1_to_l2: multilin, correct_cube, "
                             "construct_weights, ramp_fit + jump_detect, get_flat; the romancal/stcal steps and -- on sub-frames, whose "
                             "size the reference's row/channel subtraction cannot take -- the reference-pixel loop from the oracle port)")
    return "port", "oracle/rip_oracle.py l1_to_l2 (NumPy port of the reference chain)"


def cpu_baseline_single(m=1024):
    """Oracle on one core on an m^2 sub-frame; returns the cpu_baseline object."""
    dt = _oracle_tile((m, 77))
    frac = (m * m) / float(N_SIDE * N_SIDE)
    kind, what = _cpu_kind()
    return {"value": frac / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{what} on one {m}x{m} sub-frame "
                      f"(1/{int(round(1 / frac))} of an SCA, same G=8, P=11, refpix+IPC+jump flags), {dt:.1f} s, "
                      "scaled by pixel count"}  # fmt: skip


def run_reference(args):
    """--impl reference: the reference's own functions (oracle/_ref; the oracle port if that directory is absent -- the
    reference's calibrateimage itself needs asdf/romancal/stcal, which this image lacks) on all host cores, one
    independent sub-frame per process per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = max(1, min(os.cpu_count() or 1, 32))
    m = 512
    frac = (m * m) / float(N_SIDE * N_SIDE)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_oracle_tile, [(128, 10 + i) for i in range(cores)])  # warm the workers (imports, BLAS)
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_oracle_tile, [(m, 100 + 17 * s + i) for i in range(cores)])
        dt = time.perf_counter() - t0
    value = args.steps * cores * frac / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gen_cal_image L1->L2, one 4096^2 x 8-resultant SCA, P=11, CPU reference path "
                               "(BASELINE configs[0])", "step": f"{cores} independent {m}x{m} sub-frames"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": _cpu_kind()[0],
                         "sample": f"per step: {cores} processes x one {m}x{m} sub-frame through {_cpu_kind()[1]}, "
                                   "scaled by pixel count"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------
def bench_wcs(n_active, k):
    """A TAN-SIP WCS like the reference's test scene (tests/romanimpreprocess/test_workflow.py:62-83), dithered per exposure."""
    from romanimpreprocess_b200.utils import coordutils

    hdr = {"CTYPE1": "RA---TAN-SIP", "CTYPE2": "DEC--TAN-SIP", "CRPIX1": (n_active + 1) / 2.0, "CRPIX2": (n_active + 1) / 2.0,
           "CD1_1": 3.0555555555555554e-05, "CD1_2": 0.0, "CD2_1": 0.0, "CD2_2": 3.0555555555555554e-05,
           "CRVAL1": 37.0 + 0.01 * k, "CRVAL2": -20.0 + 0.005 * k, "LONPOLE": 215.0, "A_ORDER": 2, "A_0_2": 2.0e-6,
           "A_1_1": -1.0e-6, "A_2_0": 3.0e-6, "B_ORDER": 2, "B_0_2": 1.4e-5, "B_1_1": -1.0e-5, "B_2_0": 3.0e-7}  # fmt: skip
    return coordutils.FitsWCS(hdr)


def copy_ceiling(torch, dist, lib, _lib, local, world, dev, h_in, h_area, h_out, d_raw, d_amp, d_area, d_outs, steps):
    """Copy-only microbenchmark on the e2e leg's own buffers: per step one exposure's inputs host->device on one stream
    and one exposure's outputs device->host on another, nothing else; all ranks run it at the same time (barrier before),
    time = max over ranks.  Returns aggregate GB/s (both directions summed) and the SCA/s it would allow."""
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    o_slope, o_er, o_ep, o_pdq, o_end = d_outs
    pairs_out = [(h_out["slope"], o_slope), (h_out["err_read"], o_er), (h_out["err_poisson"], o_ep), (h_out["pdq"], o_pdq),
                 (h_out["endslice"], o_end)]  # fmt: skip

    def one(i):
        d, a = h_in[i % len(h_in)]
        for host, devt in ((d, d_raw), (a, d_amp)) + (((h_area, d_area),) if h_area is not None else ()):
            _lib.check(lib.rip_copy_h2d(local, C.c_void_p(devt.data_ptr()), _lib.ptr(host), C.c_size_t(host.nbytes),
                                        C.c_void_p(s_in.cuda_stream)))  # fmt: skip
        for host, devt in pairs_out:
            _lib.check(lib.rip_copy_d2h(local, _lib.ptr(host), C.c_void_p(devt.data_ptr()), C.c_size_t(host.nbytes),
                                        C.c_void_p(s_out.cuda_stream)))  # fmt: skip

    h2d = sum(x.nbytes for x in (h_in[0][0], h_in[0][1])) + (h_area.nbytes if h_area is not None else 0)
    d2h = sum(h.nbytes for h, _ in pairs_out)
    for i in range(2):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    s_in.synchronize()
    t_in = time.perf_counter() - t0
    s_out.synchronize()
    t = time.perf_counter() - t0
    tt = torch.tensor([t, t_in], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t, t_in = float(tt[0].item()), float(tt[1].item())
    return {"sca_per_s": world * steps / t, "gbs": world * steps * (h2d + d2h) / t / 1e9,
            "h2d_gbs": world * steps * h2d / t_in / 1e9, "d2h_gbs": world * steps * d2h / t / 1e9}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from romanimpreprocess_b200 import _lib, synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    rp = synth.README_PATTERN if args.groups == 8 else synth.LONG16_PATTERN
    G, n = len(rp), args.n
    n_exp = max(2, args.exposures)
    k64 = args.ipc_dtype == "f64"
    cal, exposures, area = make_inputs(n, rp, n_exp, seed=1000 + rank, ipc_dtype=np.float64 if k64 else np.float32)
    cfg = {"SLICEOUT": True}
    cd = gci.CalDir(cal, device=local)
    dplan = gci.DevicePlan(cd, rp, synth.FRAME_TIME, cfg, do_refpix=True, area_dtype=np.float32,
                           threads=args.threads, band_rows=args.band_rows)  # fmt: skip
    na = n - 8

    # ---- device-resident leg -------------------------------------------------------------------------------
    d_raw = [torch.from_numpy(d.view(np.int16)).to(dev) for d, _ in exposures]
    d_amp = [torch.from_numpy(a.view(np.int16)).to(dev) for _, a in exposures]
    d_area = torch.from_numpy(area).to(dev)
    o_slope = torch.empty((n, n), dtype=torch.float32, device=dev)
    o_er = torch.empty_like(o_slope)
    o_ep = torch.empty_like(o_slope)
    o_pdq = torch.empty((n, n), dtype=torch.int32, device=dev)
    o_end = torch.empty((na, na), dtype=torch.int8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    lookahead = not args.no_refpix_lookahead

    def step(i):
        k = i % n_exp
        gci.calibrate_device(cd, dplan, d_raw[k].data_ptr(), d_amp[k].data_ptr(), d_area.data_ptr(), o_slope.data_ptr(),
                             o_er.data_ptr(), o_ep.data_ptr(), o_pdq.data_ptr(), d_endslice=o_end.data_ptr(),
                             stream=stream)  # fmt: skip
        if lookahead:  # reference-pixel statistics of the next resident exposure beside this exposure's fused kernel
            k1 = (i + 1) % n_exp
            gci.prefetch_refpix_device(cd, d_raw[k1].data_ptr(), d_amp[k1].data_ptr(), G)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:  # noqa: BLE001
        uuid = None
    sampler = ClockSampler(local, uuid) if rank == 0 else None
    _lib.check(lib.rip_profile_enable(cd.handle, 1))
    l0 = lib.rip_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    launches = lib.rip_launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    import ctypes as C

    fused_ms, fused_n = C.c_double(0), C.c_int(0)
    _lib.check(lib.rip_profile_fetch(cd.handle, C.byref(fused_ms), C.byref(fused_n)))
    clocks = sampler.stop(t0, t1) if sampler else None
    # the fused kernel on its own (outside the timed region): with the look-ahead, the reference-pixel statistics of the
    # next exposure share the GPU with the kernel in the timed steps and stretch its launch duration a little
    alone_ms, alone_n = C.c_double(0), C.c_int(0)
    serial_ms = 0.0
    if lookahead:
        lookahead = False
        step(args.warmup + args.steps)  # (consumes the look-ahead left by the last timed step)
        torch.cuda.synchronize()
        _lib.check(lib.rip_profile_fetch(cd.handle, C.byref(alone_ms), C.byref(alone_n)))
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for i in range(min(args.steps, 10)):
            step(args.warmup + args.steps + 1 + i)
        eb.record()
        torch.cuda.synchronize()
        serial_ms = ea.elapsed_time(eb)
        _lib.check(lib.rip_profile_fetch(cd.handle, C.byref(alone_ms), C.byref(alone_n)))
        lookahead = True
    _lib.check(lib.rip_profile_enable(cd.handle, 0))
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total_max = float(tmax.item())
    value = world * args.steps / (ms_total_max * 1e-3)

    # ---- end-to-end leg: host buffers in, host buffers out, through the public API --------------------------
    if args.no_e2e:  # profiling runs only (ncu): the printed line is not a bench value
        if rank == 0:
            print(json.dumps({"profiling_run": True, "value": value, "ms_per_step": ms_total_max / args.steps,
                              "fused_ms": fused_ms.value / max(fused_n.value, 1), "gpu_launches": int(launches)}), flush=True)
        cd.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    h_in = []
    for d, a in exposures:
        pd, pa = _lib.pinned_empty(d.shape, np.uint16), _lib.pinned_empty(a.shape, np.uint16)
        pd[...] = d
        pa[...] = a
        h_in.append((pd, pa))
    h_area = _lib.pinned_empty(area.shape, np.float32)
    h_area[...] = area
    h_out = {"slope": _lib.pinned_empty((n, n), np.float32), "err_read": _lib.pinned_empty((n, n), np.float32),
             "err_poisson": _lib.pinned_empty((n, n), np.float32), "pdq": _lib.pinned_empty((n, n), np.uint32),
             "endslice": _lib.pinned_empty((na, na), np.int8)}  # fmt: skip

    # copy-only ceiling of this box for exactly these buffers: the step's H2D bytes on one stream and its D2H bytes on
    # another, concurrently, no kernels -- what a perfect pipeline could reach (all ranks at once; max over ranks)
    # The pixel-area plane of an exposure is a function of its WCS (reference gen_cal_image.py:618-621): the end-to-end
    # path uploads the ~1.7 kB of WCS coefficients and evaluates the plane on the device (rip_pipeline_set_area_wcs), as
    # calibrateimage(config) does; --e2e-area-upload restores the 67 MB host plane per step of the earlier rounds.
    up_area = args.e2e_area_upload
    wcs_list = [bench_wcs(n - 8, k) for k in range(n_exp)]
    ceil = copy_ceiling(torch, dist, lib, _lib, local, world, dev, h_in, h_area if up_area else None, h_out, d_raw[0], d_amp[0], d_area,
                        (o_slope, o_er, o_ep, o_pdq, o_end), steps=max(6, min(args.steps, 20)))

    # one set of pinned output buffers per slot in flight
    depth = 3
    h_outs = [h_out] + [{k: _lib.pinned_empty(v.shape, v.dtype) for k, v in h_out.items()} for _ in range(depth - 1)]
    pipe = gci.Pipeline(cd, rp, synth.FRAME_TIME, cfg, do_refpix=True, depth=depth, want_endslice=True)
    e2e_steps = max(6, min(args.steps, 20))

    def e2e_run(count, first):
        tickets = []
        for i in range(count):
            if len(tickets) >= depth:
                pipe.result(tickets.pop(0))  # the slot's host buffers are about to be reused
            d, a = h_in[(first + i) % n_exp]
            if not up_area:
                pipe.set_area_wcs(wcs_list[(first + i) % n_exp], dtype=np.float32)
            tickets.append(pipe.submit(d, a, h_area if up_area else None, out=h_outs[(first + i) % depth]))
        for t in tickets:
            pipe.result(t)

    e2e_run(3, 0)
    barrier()
    te0 = time.perf_counter()
    e2e_run(e2e_steps, 3)
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    temax = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(temax, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps / float(temax.item())
    # the synchronous single-call API for comparison (not the headline)
    ts0 = 0.0
    for i in range(4):  # first call allocates the handle's staging buffers: not timed; the last one is exposure 0
        if i == 1:
            ts0 = time.perf_counter()
        d, a = h_in[(i + 1) % 4 % n_exp] if i < 3 else h_in[0]
        gci.calibrate_arrays(cd, d, a, rp, synth.FRAME_TIME, h_area, cfg, do_refpix=True, want_endslice=True, out=h_out,
                             dplan=dplan)  # fmt: skip
    e2e_sync = world * 3 / (time.perf_counter() - ts0)
    pipe.close()
    h2d = int(exposures[0][0].nbytes + exposures[0][1].nbytes + (area.nbytes if up_area else wcs_list[0].pack().nbytes))
    d2h = int(sum(v.nbytes for v in h_out.values() if isinstance(v, np.ndarray)))
    checksum = int(h_out["pdq"].astype(np.uint64).sum() % (1 << 32))

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        algo_bytes = bytes_per_pixel(G, P_ORDER + 1, k64) * n * n
        fused_avg_ms = fused_ms.value / max(fused_n.value, 1)
        achieved = algo_bytes / (fused_avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "fused_traffic.json")  # dram bytes per launch from the ncu --set full capture
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(f"G{G}_P{P_ORDER + 1}_n{n}" + ("_k64" if k64 else ""))
            except Exception:  # noqa: BLE001
                traffic = None
        v6 = args.groups == 8 and not k64 and os.environ.get("RIP_FUSED_VARIANT", "4") == "4" and args.threads == 0
        kernel_name = ("cal_fused_v2k64_kernel<8,11>" if k64 else ("cal_fused_v6_kernel<8,11>" if v6 else "cal_fused_v2_kernel<8,11>")) \
            if args.groups == 8 else "cal_fused_v2_kernel<16,11>"
        band_txt = ("auto (49 rows at 4096^2 on 148 SMs: 3.97 waves of 740 resident CTAs)" if v6 else
                    "auto (62 rows at 4096^2 on 148 SMs: 3.96 waves of 592 resident CTAs)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"fused L1->L2 (gen_cal_image chain), one {n}^2 x {G}-resultant SCA per GPU per step, "
                            f"Legendre order {P_ORDER} (P={P_ORDER + 1}), {'f32 CALDIR planes with float64 ipc4d (IPC stage in float64)' if k64 else 'f32 CALDIR planes'}, refpix + IPC + ramp fit + "
                            "jump/saturation flags + dark + flat/area + endslice (BASELINE metric config)",
                "l2_policy": f"inputs larger than L2: each step streams {algo_bytes / 1e9:.2f} GB (126 MB L2) and "
                             f"{n_exp} distinct exposures are rotated",
                "threads": args.threads or 128, "band_rows": args.band_rows or band_txt, "parallelism": f"sca-sharded x{world}",
                "refpix_lookahead": bool(lookahead),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": kernel_name,
                         "kernel_ms": fused_avg_ms, "algorithmic_bytes": algo_bytes,
                         "kernel_ms_alone": (alone_ms.value / alone_n.value) if alone_n.value else fused_avg_ms,
                         "frac_alone": (algo_bytes / ((alone_ms.value / alone_n.value) * 1e-3) / 1e9 / peak) if alone_n.value else achieved / peak,
                         "step_share": fused_ms.value / ms_total,
                         "step_share_serial": (alone_ms.value / serial_ms) if serial_ms else fused_ms.value / ms_total,
                         "step_share_note": ("the reference-pixel statistics (K0) of the next exposure run on a side stream beside the "
                                             "fused kernel (refpix_lookahead): the event-timed kernel spans nearly the whole step; "
                                             "step_share_serial = the kernel's share of a step without the look-ahead (K0, then the "
                                             "kernel, on one stream), the figure to compare with the serialised ncu launch list") if lookahead else None},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "api": f"gen_cal_image.Pipeline.submit/result -> rip_pipeline_* (pinned host buffers, {depth} exposures in "
                           "flight: H2D | kernels | D2H on three streams); "
                           + ("AreaFactor plane uploaded per exposure" if up_area else
                              "AreaFactor plane evaluated on the device from each exposure's FITS WCS (TAN-SIP), inside the timed region"),
                    "sync_api_value": e2e_sync,
                    "ceiling_value": ceil["sca_per_s"], "ceiling_gbs": ceil["gbs"], "ceiling_h2d_gbs": ceil["h2d_gbs"],
                    "ceiling_d2h_gbs": ceil["d2h_gbs"], "frac_of_ceiling": e2e_value / ceil["sca_per_s"],
                    "ceiling": "copy-only: the same pinned buffers, H2D and D2H of one step on two streams, all ranks "
                               "concurrently, no kernels (max over ranks)",
                    "pdq_checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }  # fmt: skip
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_tile)
        print(json.dumps(line), flush=True)
    cd.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_forward(args):
    """Secondary workload (BASELINE configs[1], not the headline line): sim_to_isim forward ramp generation for one
    4096^2 SCA (README table: 8 resultants / 35 reads): reset noise, binomial apportioning per read, IPC + gain +
    24-step float64 bisection of the order-10 Legendre map per read, group means, read noise, biascorr, rounding."""
    import ctypes as C

    import torch

    from romanimpreprocess_b200 import _lib, synth
    from romanimpreprocess_b200.from_sim import sim_to_isim as s2i
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    torch.cuda.set_device(0)
    rp = synth.README_PATTERN
    n = args.n
    cal, _, _ = make_inputs(n, rp, 2, seed=1000)
    cd = gci.CalDir(cal, device=0)
    na = n - 8
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:na, 0:na].astype(np.float32)
    mean = 300.0 + 0.02 * xx + 4.0e4 * np.exp(-0.5 * (((xx % 512) - 256) ** 2 + ((yy % 512) - 256) ** 2) / 9.0)
    counts = rng.poisson(mean).astype(np.int32)
    d_counts = torch.from_numpy(counts).cuda()
    d_out = torch.empty((len(rp), na, na), dtype=torch.float32, device="cuda")
    lib = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        prm = s2i.fwd_params(rp, 1000 + i)
        _lib.check(lib.rip_make_l1_dev(cd.handle, C.c_void_p(d_counts.data_ptr()), C.byref(prm),
                                       C.c_void_p(d_out.data_ptr()), C.c_void_p(stream)))  # fmt: skip

    for i in range(max(args.warmup, 1)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(100 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    nreads = sum(len(g) for g in rp)
    flops = float(na) * na * nreads * (24 * (P_ORDER) * 6 + 18)  # SURVEY 8d estimate (bisection-faithful)
    print(json.dumps({"metric": "forward-model SCA/s (sim_to_isim make_l1_fullcal, 4096^2, 35 reads -> 8 resultants)",
                      "value": 1e3 / ms, "unit": "SCA/s", "ms_per_step": ms, "n_gpus": 1, "steps": args.steps,
                      "dtype": "f64", "data": "synthetic", "secondary_workload": True,
                      "roofline": {"bound": "fp64 ALU (not a measured peak: nominal B200 FP64 ~40 TFLOP/s)",
                                   "achieved": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s"},
                      "mean_DN_last_group": float(d_out[-1].mean().item())}), flush=True)  # fmt: skip
    cd.close()


def run_realizations(args):
    """Secondary workload (BASELINE configs[4], not the headline line): validation_tests/many_realizations protocol,
    `--realizations` noise realisations of one 4096^2 SCA (scene electrons -> forward ramp -> reference pixels + 1/f +
    amp33 -> fused L1->L2 -> grown mask + moment sums), device resident.  Realisations are split over the ranks; the
    three moment planes are summed with one NCCL all-reduce inside the timed region (the only exchange step)."""
    import torch
    import torch.distributed as dist

    from romanimpreprocess_b200 import pars, synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
    from romanimpreprocess_b200.validation_tests import many_realizations as mr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rp = synth.README_PATTERN
    n = args.n
    cal, _, area = make_inputs(n, rp, 2, seed=1000)  # the same SCA on every rank
    cd = gci.CalDir(cal, device=local)
    na = n - 8
    yy, xx = np.mgrid[0:na, 0:na].astype(np.float32)
    image = (3.0 + 0.002 * xx + 400.0 * np.exp(-0.5 * (((xx % 512) - 256) ** 2 + ((yy % 512) - 256) ** 2) / 9.0)).astype(np.float32)
    R = args.realizations
    mine = [j for j in range(R) if j % world == rank]
    rz = mr.Realizations(image, cd, rp, area_ratio=area, config2={"SLICEOUT": True}, device=local, keep_stacks=0)
    for w in range(max(args.warmup, 1)):
        rz.step(7 + w)
    rz.d_moments.zero_()
    rz.done = 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for j in mine:
        rz.step(100 + 10 * (j + 1))
    if world > 1:
        dist.all_reduce(rz.d_moments)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = rz.finalize(slope_ideal=None)
    if rank == 0:
        good = out[3, 8:-8, 8:-8] >= R - 1
        ideal = np.zeros((n, n), np.float32)
        ideal[4:-4, 4:-4] = image / pars.g_ideal
        bias = np.median((out[4] - ideal)[8:-8, 8:-8][good] / ideal[8:-8, 8:-8][good])
        print(json.dumps({"metric": "many-realisations SCA realisations/s (scene -> L1 -> L2 -> moments, 4096^2 x 8 resultants)",
                          "value": R / (float(ms.item()) * 1e-3), "unit": "realisations/s", "n_gpus": world,
                          "realizations": R, "ms_per_realization_per_gpu": float(ms.item()) / max(len(mine), 1),
                          "scaling": "strong", "dtype": "f64+f32", "data": "synthetic", "secondary_workload": True,
                          "exchange": "one NCCL all-reduce of the 3 moment planes (200 MB)" if world > 1 else "none",
                          "check": {"unmasked_in_all_fraction": float(good.mean()), "median_relative_bias": float(bias)}}),
              flush=True)  # fmt: skip
    rz.close()
    cd.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_exposure18(args):
    """Secondary workload (BASELINE configs[3]): whole 18-SCA WFI exposures through L1->L2, the (exposure, SCA) items
    dealt to the ranks by sharding.assign_items_balanced (reference: one Slurm task per SCA,
    runs/summer2025run/OpenUniverse_to_L1L2.job:4), every rank holding only its own SCAs' CALDIRs (all 18 on one GPU,
    2-4 per GPU on eight).  `--exposures18` exposures in total: strong scaling.  No collective on the data path."""
    import torch
    import torch.distributed as dist

    from romanimpreprocess_b200 import _lib, sharding, synth
    from romanimpreprocess_b200.L1_to_L2 import exposure_driver as xd
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    rp = synth.README_PATTERN
    n, G, NSCA, E = args.n, len(synth.README_PATTERN), 18, args.exposures18
    cal, exposures, area = make_inputs(n, rp, 2, seed=1000)  # one synthetic CALDIR content, uploaded once per SCA
    cfg = {"SLICEOUT": True}
    items = [(e, sca) for e in range(E) for sca in range(1, NSCA + 1)]
    t_setup = time.perf_counter()
    drv = xd.ExposureCalibrator({sca: cal for sca in range(1, NSCA + 1)}, items, rp, synth.FRAME_TIME, cfg, rank=rank,
                                world=world, device=local, depth=2)  # fmt: skip
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    na = n - 8
    free_b, total_b = torch.cuda.mem_get_info(local)

    # ---- device-resident leg: this rank's items, inputs in HBM ----------------------------------------------
    d_raw = [torch.from_numpy(d.view(np.int16)).to(dev) for d, _ in exposures]
    d_amp = [torch.from_numpy(a.view(np.int16)).to(dev) for _, a in exposures]
    d_area = torch.from_numpy(area).to(dev)
    o_slope = torch.empty((n, n), dtype=torch.float32, device=dev)
    o_er, o_ep = torch.empty_like(o_slope), torch.empty_like(o_slope)
    o_pdq = torch.empty((n, n), dtype=torch.int32, device=dev)
    o_end = torch.empty((na, na), dtype=torch.int8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    lookahead = not args.no_refpix_lookahead

    def resident_pass():
        for k, (e, sca) in enumerate(drv.items):
            gci.calibrate_device(drv.cals[sca], drv.pipes[sca].dplan, d_raw[k % 2].data_ptr(), d_amp[k % 2].data_ptr(),
                                 d_area.data_ptr(), o_slope.data_ptr(), o_er.data_ptr(), o_ep.data_ptr(), o_pdq.data_ptr(),
                                 d_endslice=o_end.data_ptr(), stream=stream)  # fmt: skip
            if lookahead:  # reference-pixel statistics of the next item (on ITS CalDir handle) beside this fused kernel
                kn = (k + 1) % len(drv.items)
                sca1 = drv.items[kn][1]
                gci.prefetch_refpix_device(drv.cals[sca1], d_raw[kn % 2].data_ptr(), d_amp[kn % 2].data_ptr(), G)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 3)):
        resident_pass()
    barrier()
    l0 = lib.rip_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        resident_pass()
    ev1.record()
    barrier()
    launches = lib.rip_launch_count() - l0
    tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_pass = float(tmax.item()) / args.steps  # one pass = all E exposures of the job
    value = E / (ms_pass * 1e-3)

    # ---- end-to-end leg: pinned host cubes in, pinned host arrays out, WCS -> area on the device --------------
    h_in = []
    for d, a in exposures:
        pd, pa = _lib.pinned_empty(d.shape, np.uint16), _lib.pinned_empty(a.shape, np.uint16)
        pd[...] = d
        pa[...] = a
        h_in.append((pd, pa))
    outs = [{"slope": _lib.pinned_empty((n, n), np.float32), "err_read": _lib.pinned_empty((n, n), np.float32),
             "err_poisson": _lib.pinned_empty((n, n), np.float32), "pdq": _lib.pinned_empty((n, n), np.uint32),
             "endslice": _lib.pinned_empty((na, na), np.int8)} for _ in range(6)]  # fmt: skip
    wcs = {it: bench_wcs(na, it[0] * NSCA + it[1]) for it in drv.items}
    sums = []

    def fetch(e, sca):
        d, a = h_in[(e + sca) % 2]
        return d, a, wcs[(e, sca)]

    def sink(e, sca, out):
        sums.append(int(out["pdq"][n // 2, ::64].astype(np.uint64).sum()))  # the result is read on the host

    drv.run(fetch, sink, list(outs))  # warm-up pass (allocates nothing afterwards)
    barrier()
    te0 = time.perf_counter()
    ndone = drv.run(fetch, sink, list(outs))
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    rate_items, total_items, te_max = sharding.job_throughput(ndone, te)
    e2e_value = rate_items / NSCA
    counts = torch.zeros(world, dtype=torch.float64, device=dev)
    counts[rank] = len(drv.items)
    nres = torch.zeros(world, dtype=torch.float64, device=dev)
    nres[rank] = len(drv.scas)
    if world > 1:
        dist.all_reduce(counts)
        dist.all_reduce(nres)
    if rank == 0:
        h2d = int(exposures[0][0].nbytes + exposures[0][1].nbytes + 8 * 211)
        d2h = int(sum(v.nbytes for v in outs[0].values() if isinstance(v, np.ndarray)))
        print(json.dumps({
            "metric": "18-SCA WFI exposures/sec through L1->L2 (4096^2 x 8 resultants per SCA)", "value": value,
            "unit": "exposures/s", "sca_per_s": value * NSCA, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_pass, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (one CALDIR content uploaded as 18 separate resident handles; 2 exposure cubes rotated)",
            "secondary_workload": True,
            "config": {"refpix_lookahead": bool(lookahead),
                       "workload": f"exposure18: {E} exposures x 18 SCAs = {len(items)} (exposure, SCA) items per pass, dealt by "
                                   "sharding.assign_items_balanced; one resident CALDIR handle + pipeline per SCA of a rank",
                       "items_per_rank": [int(c) for c in counts.tolist()], "resident_caldirs_per_rank": [int(c) for c in nres.tolist()],
                       "imbalance_max_over_mean": float(counts.max().item() / counts.mean().item()),
                       "sca_major_imbalance": sharding.imbalance(items, world, sharding.assign_items),
                       "hbm_used_gb_rank0": (total_b - free_b) / 1e9, "caldir_setup_s_rank0": t_setup,
                       "parallelism": f"sca-sharded x{world}"},
            "e2e": {"value": e2e_value, "unit": "exposures/s", "sca_per_s": rate_items, "h2d_bytes_per_step": h2d * len(items),
                    "d2h_bytes_per_step": d2h * len(items), "items": int(total_items), "seconds_max_over_ranks": te_max,
                    "api": "exposure_driver.ExposureCalibrator.run -> one gen_cal_image.Pipeline per resident SCA (pinned host "
                           "buffers, AreaFactor from each item's WCS on the device)", "checksum": int(sum(sums) % (1 << 32))},
            "gpu_launches": int(launches)}), flush=True)  # fmt: skip
    drv.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_files(args):
    """Secondary workload: the file-level drop-in at full size.  CALDIR files, L1 exposures and FITSWCS
    headers are written once to tmpfs (layouts of SURVEY App. B, tests/fixture_files.py); each step is one
    gen_cal_image.calibrateimage(config): read the L1 ASDF file, L1->L2 on the GPU (resident CALDIR, area from the WCS,
    sky mode + SKYORDER fit), write the L2 ASDF file.  Wall clock, one GPU."""
    import shutil

    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fixture_files import write_exposure

    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci

    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    tmp = tempfile.mkdtemp(prefix="rip_files_", dir=base)
    try:
        t0 = time.perf_counter()
        config, cal, data, amp33, rp = write_exposure(tmp, n=args.n, seed=41, p_order=P_ORDER, ipc_dtype=np.float64)
        t_write = time.perf_counter() - t0
        in_bytes = os.path.getsize(config["IN"])
        gci.calibrateimage(config, verbose=False)  # first exposure: opens the CALDIR files, builds the resident handle
        torch.cuda.synchronize()
        ts = []
        for i in range(max(args.steps, 2)):
            cfg = dict(config, OUT=os.path.join(tmp, f"sim_L2_F184_1_1_{i}.asdf"))
            t = time.perf_counter()
            gci.calibrateimage(cfg, verbose=False)
            ts.append(time.perf_counter() - t)
            out_bytes = os.path.getsize(cfg["OUT"])
            os.remove(cfg["OUT"])
        dt = float(np.median(ts))
        print(json.dumps({"metric": "calibrateimage(config) exposures/s from / to ASDF files on tmpfs (4096^2 x 8 resultants)",
                          "value": 1.0 / dt, "unit": "SCA/s", "s_per_exposure": dt, "n_gpus": 1, "steps": len(ts),
                          "dtype": "f32 (float64 ipc4d: the DUMMY CALDIR dtype)", "data": "synthetic", "secondary_workload": True,
                          "l1_file_bytes": int(in_bytes), "l2_file_bytes": int(out_bytes), "fixture_write_s": t_write,
                          "io": "asdf" if __import__("romanimpreprocess_b200.caltree", fromlist=["have_asdf"]).have_asdf() else "io/asdf_lite.py",
                          "timing": "wall clock per call, median; includes reading the L1 file, packaging and writing the L2 file"}),
              flush=True)  # fmt: skip
        gci.clear_caldir_cache()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_noiselayers(args):
    """Secondary workload (the production call pattern of runs/summer2025run/OpenUniverse_to_L1L2.py:
    124-133): per exposure `--layers` noise layers (half "Rz4PbrS2C<i>", half "Rz4OS2C<i>": the production list) with SKYORDER 2 = (2 + layers) full L1->L2 calibrations,
    white + correlated noise generation (34 x 8 FFTs of 2^20 points per layer), sky-mode fits, differences; device
    resident except the z-clip percentiles and the returned layers (host)."""
    import torch

    from romanimpreprocess_b200 import synth
    from romanimpreprocess_b200.L1_to_L2 import gen_cal_image as gci
    from romanimpreprocess_b200.L1_to_L2 import gen_noise_image as gni

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rp = synth.README_PATTERN
    n = args.n
    cal, exposures, area = make_inputs(n, rp, 2, seed=1000)
    cd = gci.CalDir(cal, device=local)
    cfg = {"SKYORDER": 2}
    nl = gni.NoiseLayers(cd, rp, synth.FRAME_TIME, cfg, device=local)
    d_data = torch.from_numpy(exposures[0][0].view(np.int16)).to(dev).view(torch.uint16)
    d_amp = torch.from_numpy(exposures[0][1].view(np.int16)).to(dev).view(torch.uint16)
    d_area = torch.from_numpy(area).to(dev)
    # the production list: 4 x "Rz4PbrS2C*" + 4 x "Rz4OS2C*" (runs/summer2025run/OpenUniverse_to_L1L2.py:124-133)
    layers = [f"Rz4PbrS2C{i + 1}" if i < args.layers // 2 else f"Rz4OS2C{i + 1}" for i in range(args.layers)]
    nl.set_exposure(d_data, d_amp, d_area)
    nl.layer(layers[0], 1)  # warm-up: workspaces, FFT attributes, the reference calibration of the dark cube
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    from romanimpreprocess_b200 import _lib

    kept = [_lib.pinned_empty((n - 8, n - 8), np.float32) for _ in layers]
    for e in range(args.steps):
        nl.set_exposure(d_data, d_amp, d_area)
        for i, cmd in enumerate(layers):
            nl.layer(cmd, 1000 * e + i + 2, out=kept[i])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    sig = []
    for lay in kept:  # (outside the timed region) robust scatter of each layer of the last exposure
        q = np.percentile(lay[8:-8, 8:-8], [25, 75])
        sig.append(float((q[1] - q[0]) / 1.34896))
    print(json.dumps({"metric": "noise-layer exposures/s (gen_noise_image, 4096^2 x 8 resultants, layers Rz4PbrS2C* / Rz4OS2C*, SKYORDER 2)",
                      "value": 1.0 / dt, "unit": "exposures/s", "layers_per_exposure": len(layers),
                      "ms_per_layer": 1e3 * dt / len(layers), "n_gpus": 1, "steps": args.steps, "dtype": "f32", "data": "synthetic",
                      "secondary_workload": True, "timing": "wall clock incl. the D2H of every layer (pinned buffers) and the host-side normal equations of the sky fits",
                      "layer_sigma_DN_per_s": sig}), flush=True)  # fmt: skip
    nl.close()
    cd.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_SIDE, help="frame side (default 4096; smaller only for debugging)")
    ap.add_argument("--groups", type=int, default=8, choices=[8, 16])
    ap.add_argument("--ipc-dtype", default="f32", choices=["f32", "f64"], help="dtype of the CALDIR's ipc4d (f64: the DUMMY builder's)")
    ap.add_argument("--exposures", type=int, default=3, help="distinct resident exposures rotated through the steps")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--band-rows", type=int, default=0)
    ap.add_argument("--cpu-tile", type=int, default=2048, help="side of the sub-frame the cpu_baseline times (2048: about 12 s of oracle work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident leg only (for ncu captures)")
    ap.add_argument("--no-refpix-lookahead", action="store_true",
                    help="device-resident leg: reference-pixel statistics in the step's own stream (no look-ahead of the next exposure)")
    ap.add_argument("--e2e-area-upload", action="store_true", help="e2e leg: upload the AreaFactor plane per exposure instead of evaluating it on the device from the WCS")
    ap.add_argument("--realizations", type=int, default=64, help="noise realisations (workload realizations)")
    ap.add_argument("--exposures18", type=int, default=4, help="exposures of 18 SCAs each (workload exposure18)")
    ap.add_argument("--layers", type=int, default=8, help="noise layers per exposure (workload noiselayers)")
    ap.add_argument("--workload", default="l1l2", choices=["l1l2", "forward", "realizations", "noiselayers", "exposure18", "files"],
                    help="l1l2 = the headline metric; forward = secondary line for the forward ramp generator")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "forward":
        run_forward(args)
    elif args.workload == "realizations":
        run_realizations(args)
    elif args.workload == "noiselayers":
        run_noiselayers(args)
    elif args.workload == "exposure18":
        run_exposure18(args)
    elif args.workload == "files":
        run_files(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
