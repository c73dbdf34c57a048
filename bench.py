#!/usr/bin/env python
"""bench.py -- DSD->PCM decimation throughput on B200 (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2]

A step is one pass of the hot path (reference DsdDecimator::getSamples, dsd_decimator.cpp:173-222,
as main.cpp:169-191 drives it for one whole file) over one synthetic sigma-delta stream per GPU.
Default workload: BASELINE.json configs[1] -- DSD128 stereo, 60 s -> 24 bit / 176.4 kHz, fp64
tables, TPDF dither off (the bit-exact parity setting).  Weak scaling: every rank converts its
own stream, no data-path collective (shards are independent).

Printed (rank 0, one JSON line):
  value   whole-job DSD Mbit/s, inputs resident in HBM, K steps timed with CUDA events
  e2e     the same through dsdgpu_decimate_host: pinned host bytes in, host PCM out, copies timed
  roofline  the decimation kernel against the shared-memory (LUT gather) roofline + HBM figures
  cpu_baseline  the reference's own decimator (oracle/_ref, or the oracle port) on the host cores
--impl reference times that CPU path alone on the same workload (rank 0 only).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DSD64 = 2822400
WORKLOADS = {
    # name: (dsd_fs, out_fs, nch, seconds, bits, description)
    "c1": (DSD64, 88200, 2, 60, 24, "DSD64 stereo 60 s -> 24-bit/88.2 kHz (BASELINE configs[0])"),
    "c2": (2 * DSD64, 176400, 2, 60, 24, "DSD128 stereo 60 s -> 24-bit/176.4 kHz, fp64 LUT (BASELINE configs[1])"),
    # the other two filters of the reference (dsd_decimator.cpp:57-65): same input, higher output rates
    "d16": (DSD64, 176400, 2, 60, 24, "DSD64 stereo 60 s -> 24-bit/176.4 kHz (divide-by-16 filter, 30 tables)"),
    "d8": (DSD64, 352800, 2, 60, 24, "DSD64 stereo 60 s -> 24-bit/352.8 kHz (divide-by-8 filter, 12 tables)"),
    # EXTENSION filters (no reference counterpart): BASELINE configs[3] as literally stated
    "x64": (8 * DSD64, 352800, 2, 60, 24, "DSD512 stereo 60 s -> 24-bit/352.8 kHz (divide-by-64 EXTENSION filter, 144 tables, "
                                          "two passes of the segmented kernel)"),
    "x128": (8 * DSD64, 176400, 2, 60, 24, "DSD512 stereo 60 s -> 24-bit/176.4 kHz (divide-by-128 EXTENSION filter, 288 tables)"),
    "x256": (8 * DSD64, 88200, 2, 60, 24, "DSD512 stereo 60 s -> 24-bit/88.2 kHz (divide-by-256 EXTENSION filter, 576 tables)"),
}
# Sharded (strong-scaling) workloads, BASELINE configs[2..4]: (description, [(dsd_fs, out_fs, nch, seconds)...], lut, dither)
def sharded_workloads(nfiles=256):
    rng = np.random.default_rng(42)
    c5 = [((DSD64, 88200) if k % 2 == 0 else (2 * DSD64, 176400)) + (2, int(rng.integers(30, 91))) for k in range(nfiles)]
    return {
        "c3": ("DSD256 5.1 60 s -> 24-bit/352.8 kHz, sharded by channel and time (BASELINE configs[2])",
               [(4 * DSD64, 352800, 6, 60)], 64, 0),
        "c4": ("DSD512 stereo 600 s, divide-by-32, time segments with a 71-byte halo (BASELINE configs[3]; the "
               "reference has no divide-by-64 filter for 352.8 kHz)", [(8 * DSD64, 705600, 2, 600)], 64, 0),
        "c4x": ("DSD512 stereo 600 s -> 24-bit/352.8 kHz as BASELINE configs[3] states it: divide-by-64 EXTENSION filter "
                "(the reference rejects this ratio), time segments with a 143-byte halo", [(8 * DSD64, 352800, 2, 600)], 64, 0),
        "c5": (f"{nfiles} files, DSD64->88.2k / DSD128->176.4k alternating, 30-90 s, TPDF dither on, 32-bit tables "
               "(BASELINE configs[4])", c5, 32, 1),
    }


SMEM_BYTES_PER_CLK_PER_SM = 128
L2_BYTES = 126 * 1024 * 1024


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


DATA_KIND = "sdm"            # --data: "sdm" (sigma-delta sines, the default) or "random" (uniform bytes, seed 1234)


def synth_stream(dsd_fs, nch, seconds, lsb_first=1):
    """One second of 2nd-order sigma-delta sine per channel (host/dsd_synth.cpp), tiled; or uniform random
    bytes (shared-memory behaviour depends on the byte statistics, so both are reported)."""
    if DATA_KIND == "random":
        return np.random.default_rng(1234).integers(0, 256, size=(nch, dsd_fs // 8 * seconds), dtype=np.uint8)
    from dsf2flac_b200 import capi
    host = capi.host_lib()
    sec = dsd_fs // 8
    one = np.empty((nch, sec), dtype=np.uint8)
    for c in range(nch):
        host.dsdhost_synth_sdm(one[c].ctypes.data, sec, float(dsd_fs), 1000.0 + 500.0 * c, 0.5, lsb_first)
    return np.ascontiguousarray(np.tile(one, (1, seconds)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for (_, r) in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


# ---- CPU baseline (test infrastructure: oracle/_ref = the unmodified reference decimator) -----

_CPU_DATA = None        # inherited by the forked workers (no pickling of the stream)


def _cpu_worker(args):
    seg_bytes, k, dsd_fs, out_fs, nlut, nstep, scale, clip, reps = args
    data = _CPU_DATA
    from tests.oracle_lib import Oracle, REF_SO
    from dsf2flac_b200 import filters
    orc = Oracle(reference=os.path.exists(REF_SO))
    if dsd_fs // out_fs not in filters.FILTERS:             # extension filter: the reference's arithmetic on it
        orc.set_custom_filter(filters.EXTENSION_FILTERS[dsd_fs // out_fs])
    lo = k * seg_bytes
    sub = np.ascontiguousarray(data[:, lo:lo + seg_bytes])
    nframes = (seg_bytes - nlut) // nstep
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.run_raw(sub, 1 << 40, 1, dsd_fs, out_fs, nlut + 1, nframes, scale, 0.0, clip)
    return reps * nframes * nstep * 8 * data.shape[0], time.perf_counter() - t0


def cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, pool=None, reps=1):
    """The reference's CPU decimator (getSamples loop only, like the GPU arm: no FLAC, no file
    parsing) over the first `sample_seconds` of the stream, `reps` times over, one process per core
    (the reference is single threaded and keeps process-global state, so cores = processes).
    -> (Mbit/s, bits, kind)"""
    from tests.oracle_lib import REF_SO
    nbytes = min(data.shape[1], int(dsd_fs // 8 * sample_seconds))
    seg = nbytes // cores // 4 * 4
    jobs = [(seg, k, dsd_fs, out_fs, fir.nlut, fir.nstep, scale, clip, reps) for k in range(cores)]
    own = pool is None
    if own:
        global _CPU_DATA
        _CPU_DATA = data
        pool = mp.get_context("fork").Pool(cores)
        pool.map(abs, range(cores))                        # workers up before the clock starts
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
    bits = sum(r[0] for r in res)
    return bits / wall / 1e6, bits, ("reference" if os.path.exists(REF_SO) else "port")


def run_reference_arm(args, wl):
    dsd_fs, out_fs, nch, seconds, bits_depth, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dsf2flac_b200 import filters
    fir = filters.pick(dsd_fs, out_fs, extensions=True)
    scale, _, clip = filters.app_params(bits_depth, dither=False)
    cores = len(os.sched_getaffinity(0))
    global _CPU_DATA
    data = _CPU_DATA = synth_stream(dsd_fs, nch, seconds)
    pool = mp.get_context("fork").Pool(cores)
    # a step is the whole file, like the GPU arm's, unless that would push the run past ~2 minutes
    probe, _, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, 1.0, pool)
    budget_s = 100.0 / max(1, args.steps + args.warmup)
    sample_seconds = max(0.25, min(data.shape[1] * 8 / dsd_fs, budget_s * probe * 1e6 / (dsd_fs * nch)))
    for _ in range(args.warmup):
        cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, pool)
    t0 = time.perf_counter()
    total_bits = 0
    for _ in range(args.steps):
        _, b, _ = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, pool)
        total_bits += b
    wall = time.perf_counter() - t0
    pool.close()
    v = total_bits / wall / 1e6
    sample = f"first {sample_seconds:.2f} s of the stream per step, split over {cores} processes"
    print(json.dumps({
        "impl": "reference", "metric": "dsd_mbit_per_s", "value": v, "unit": "Mbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "x_realtime": v * 1e6 / (dsd_fs * nch),
        "config": {"workload": desc, "dither": False, "lut": "fp64"},
        "cpu_baseline": {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def _cpu_file_worker(job):
    """One whole file through the reference's decimator with main.cpp's loop (dither on): what one
    dsf2flac process would do for it, minus container parsing and FLAC."""
    dsd_fs, out_fs, nch, seconds, bits_depth, dither = job
    from tests.oracle_lib import Oracle, REF_SO
    from dsf2flac_b200 import filters
    orc = Oracle(reference=os.path.exists(REF_SO))
    fir = filters.pick(dsd_fs, out_fs, extensions=True)
    if fir.ratio not in filters.FILTERS:
        orc.set_custom_filter(fir)
    one = _CPU_DATA[dsd_fs][:nch]
    data = np.ascontiguousarray(np.tile(one, (1, seconds)))
    scale, amp, clip = filters.app_params(bits_depth, dither=bool(dither))
    t0 = time.perf_counter()
    out = orc.run_app(data, data.shape[1] * 8, 1, dsd_fs, out_fs, scale, amp, clip)
    return data.shape[1] * 8 * nch, time.perf_counter() - t0, len(out)


def run_reference_arm_batch(args, spec):
    """Reference arm of the sharded workloads: the job's files (time segments of one long stream count as
    files of their own) converted by independent reference processes, one per host core -- the reference is
    single threaded and its only batch notion is a serial shell loop (scripts/batch_dsf2flac.sh:16-19).  A step
    is a bounded sample: the first files of the list, as many as keep the run within a few minutes."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    desc, files, lut_bits, dither = spec
    cores = len(os.sched_getaffinity(0))
    global _CPU_DATA
    _CPU_DATA = {}
    for fs in sorted({f[0] for f in files}):
        _CPU_DATA[fs] = synth_stream(fs, max(f[2] for f in files if f[0] == fs), 1)
    # a long stream is cut into 60 s pieces (what the GPU arm's time segments are, without the halo)
    jobs = []
    for fs, out_fs, nch, seconds in files:
        for lo in range(0, seconds, 60):
            jobs.append((fs, out_fs, nch, min(60, seconds - lo), 24, dither))
    pool = mp.get_context("fork").Pool(cores)
    probe = pool.map(_cpu_file_worker, jobs[:cores])
    rate = sum(p[0] for p in probe) / max(p[1] for p in probe)          # bits/s with every core busy
    budget_bits = rate * 100.0 / max(1, args.steps + args.warmup)
    take, acc = 0, 0
    while take < len(jobs) and (take < cores or acc < budget_bits):
        acc += jobs[take][0] * jobs[take][2] * jobs[take][3]
        take += 1
    sample = jobs[:take]
    for _ in range(args.warmup):
        pool.map(_cpu_file_worker, sample, chunksize=1)
    t0 = time.perf_counter()
    bits = 0
    for _ in range(args.steps):
        bits += sum(r[0] for r in pool.map(_cpu_file_worker, sample, chunksize=1))
    wall = time.perf_counter() - t0
    pool.close()
    v = bits / wall / 1e6
    print(json.dumps({
        "impl": "reference", "metric": "dsd_mbit_per_s", "value": v, "unit": "Mbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "files": len(files), "dither": bool(dither), "lut": "fp64 (the reference has no 32-bit variant)"},
        "cpu_baseline": {"value": v, "unit": "Mbit/s", "cores": cores, "kind": "reference" if os.path.exists(
            os.path.join(ROOT, "oracle", "_ref", "libdsdref.so")) else "port",
                         "sample": f"the first {take} of {len(jobs)} files / 60 s pieces per step, one reference process per core"},
        "e2e": {"value": v, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---- GPU arm ----------------------------------------------------------------------------------

def run_gpu_arm(args, wl):
    import torch
    import torch.distributed as dist
    from dsf2flac_b200 import capi, filters

    dsd_fs, out_fs, nch, seconds, bits_depth, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fir = filters.pick(dsd_fs, out_fs, extensions=True)
    lut = filters.build_lut(fir, 1)
    scale, amp, clip = filters.app_params(bits_depth, dither=args.dither)
    data = synth_stream(dsd_fs, nch, seconds)
    nbytes = data.shape[1]
    nframes = filters.app_frames(nbytes * 8, fir)
    bits_per_step = nbytes * 8 * nch                       # DSD bits consumed per step per GPU
    g = capi.DsdGpu(lut, nch, fir.nstep, device=local, lut_precision=args.lut)
    if args.kernel:
        g.set_option("kernel", {"auto": 0, "direct": 1, "ring": 3, "seg": 4}[args.kernel])
    if args.slots:
        g.set_option("slots", args.slots)
    if args.chunk_frames:
        g.set_option("chunk_frames", args.chunk_frames)

    # device-resident ring of distinct input/output buffers, larger than L2 in total, so no step
    # finds its bytes in L2 from the step before
    # The stream sits in its device buffer behind `front` reader-idle bytes (what the reader's ring
    # holds before byte 0, dsd_sample_reader.h:122), chosen so that the frame grid is 4-byte aligned
    # to the buffer ((p0 - in_first + 1) % 4 == 0): the kernel's window loads are then conflict free.
    in_bytes, out_bytes = nch * nbytes, nframes * nch * 4
    ring = max(2, -(-2 * L2_BYTES // (in_bytes + out_bytes)))
    front = (fir.nlut + 1) % 4 if args.align else 0
    front = (4 - front) % 4
    stride = (nbytes + front + 255) // 256 * 256
    d_in = []
    for k in range(ring):
        t = torch.full((nch, stride), 0x69, dtype=torch.uint8)
        t[:, front:front + nbytes] = torch.from_numpy(np.roll(data, 4096 * k, axis=1))
        d_in.append(t.cuda())
    d_out = [torch.empty((nframes, nch), dtype=torch.int32, device="cuda") for _ in range(ring)]
    stream = torch.cuda.Stream()

    def step(k):
        g.decimate_device(d_in[k % ring], stride, -front, nbytes + front, 0x69, fir.nlut, nframes, scale, amp, clip, 0,
                          capi.OUT_INT32, d_out[k % ring], stream=stream.cuda_stream)

    sampler = ClockSampler(local) if rank == 0 else None     # polls every 100 ms from here to the end of the e2e leg
    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            step(k)
    barrier()
    launches0 = g.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for k in range(args.steps):
            step(k)
            ev[k + 1].record(stream)
    stream.synchronize()
    barrier()
    t_wall1 = time.perf_counter()
    launches = g.launches - launches0
    dev_ms = ev[0].elapsed_time(ev[-1])
    per_step = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    total_ms = max_over_ranks(dev_ms)

    # ---- end to end: host bytes -> PCM on the host, through the C ABI with host buffers -------
    pin_in = capi.PinnedBuffer((nch, nbytes), np.uint8)
    pin_in.array[:] = data
    pin_out = capi.PinnedBuffer((nframes, nch), np.int32)

    def e2e_step():
        g.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, nframes, scale, amp, clip, 0, capi.OUT_INT32,
                            pin_out.ptr)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()                                         # synchronous: returns when the PCM is on the host
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    checksum = int(pin_out.array[:: max(1, nframes // 65536)].astype(np.int64).sum())
    # the device-timed region lasts a few ms, shorter than any nvidia-smi poll: report the clocks seen
    # from the first warm-up launch to the end of the end-to-end leg (the GPU is busy throughout)
    clocks = sampler.stop(t_wall0 - 3600.0, time.perf_counter()) if sampler else None

    # ---- what the link allows: pinned copies of the same sizes, one way and both ways at once --
    link = None
    if rank == 0:
        h_in = torch.from_numpy(pin_in.array.reshape(-1))          # views of the pinned buffers
        h_out = torch.from_numpy(pin_out.array.reshape(-1))
        dv_in = torch.empty_like(h_in, device="cuda")
        dv_out = torch.empty_like(h_out, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def timed(fn, reps=5):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        def up():
            with torch.cuda.stream(s1):
                dv_in.copy_(h_in, non_blocking=True)

        def down():
            with torch.cuda.stream(s2):
                h_out.copy_(dv_out, non_blocking=True)

        t_up, t_down = timed(up), timed(down)
        t_both = timed(lambda: (up(), down()))
        link = {"h2d_gbs": in_bytes / t_up / 1e9, "d2h_gbs": out_bytes / t_down / 1e9,
                "duplex_ms": t_both * 1e3, "note": "pinned cudaMemcpyAsync of one step's input / output; duplex = both at once"}
        del dv_in, dv_out

    # ---- the drop-in class itself (reader drained byte by byte, as dsf2flac would run it) -----
    plugin = None
    if rank == 0 and args.plugin_seconds > 0:
        host = capi.host_lib()
        host.dsdhost_set_extension_filters(0 if fir.ratio in filters.FILTERS else 1)
        pb = min(nbytes, int(dsd_fs // 8 * args.plugin_seconds))
        cap = nframes + 2048
        out = capi.PinnedBuffer((cap, nch), np.int32)
        plugin = {}
        # (a) as dsf2flac runs it today: a reader stepped one byte per channel at a time
        # (b) the same class with the file's bytes handed over in bulk (DsdDecimator::setBulkSource)
        for name, fn, count in (("stepped_reader", host.dsdhost_run_app, pb), ("bulk_source", host.dsdhost_run_app_bulk, nbytes)):
            best = None
            for _ in range(3):                              # the first call pays context creation and page-locking
                t0 = time.perf_counter()
                n = fn(pin_in.ptr, nch, nbytes, count * 8, 0x69, 1, dsd_fs, out_fs, scale, amp, clip, args.lut, 0, out.ptr, cap)
                dt = time.perf_counter() - t0
                if n > 0 and (best is None or dt < best):
                    best = dt
            if best:
                plugin[name] = {"value": count * 8 * nch / best / 1e6, "unit": "Mbit/s", "stream_seconds": count * 8 / dsd_fs,
                                "wall_ms": best * 1e3}
        plugin["note"] = ("drop-in DsdDecimator::getSamples in main.cpp's 1024-frame calls, decimator construction "
                          "(tables + GPU context) included")
        out.free()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = bits_per_step * args.steps * world / (total_ms * 1e-3) / 1e6
        kernel_ms = per_step[len(per_step) // 2]
        samples_per_launch = nframes * nch
        lut_elem = 8 if args.lut == 64 else 4
        smem_bytes = samples_per_launch * fir.nlut * lut_elem          # algorithmic LUT bytes per launch
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        smem_peak = n_sm * SMEM_BYTES_PER_CLK_PER_SM * peaks["sm_max_mhz"] * 1e6 / 1e9
        achieved = smem_bytes / (kernel_ms * 1e-3) / 1e9
        hbm_bytes = samples_per_launch * (fir.nstep + 4)
        roof = {"bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
                "traffic": None, "peak_source": f"{n_sm} SMs x 128 B/clk x sm_max_mhz ({peak_src})",
                "algorithmic_bytes_per_sample": fir.nlut * lut_elem, "samples_per_launch": samples_per_launch,
                "kernel_ms": kernel_ms,
                "hbm": {"bound": "hbm", "achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "algorithmic_bytes_per_sample": fir.nstep + 4, "peak_source": peak_src}}
        try:                                               # measured DRAM traffic of the same kernel on the same workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.kernel or "ring"]
            if tr["workload"] == args.workload and fir.ratio == 32 and not args.dither:
                roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roof["traffic_source"] = tr["source"]
                roof["algorithmic_hbm_bytes"] = hbm_bytes
        except Exception:
            pass
        if clocks and clocks.get("sm_mhz"):
            roof["frac_at_sampled_clock"] = achieved / (n_sm * SMEM_BYTES_PER_CLK_PER_SM * clocks["sm_mhz"] * 1e6 / 1e9)
        # CPU baseline: bounded sample on the host cores
        cores = len(os.sched_getaffinity(0))
        cpu = None
        if world == 1 and args.cpu_seconds > 0:
            # about args.cpu_seconds of wall time on every core: the whole file, as often as that takes
            probe, _, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, seconds)
            reps = max(1, int(args.cpu_seconds * probe * 1e6 / bits_per_step))
            v, b, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, seconds, reps=reps)
            v1, b1, _ = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, 1, seconds / 2)
            cpu = {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind,
                   "sample": f"the whole {seconds} s file {reps} times over ({b / 1e9:.1f} Gbit), one process per core "
                             f"(the reference is single threaded), decimator only",
                   "one_core": {"value": v1, "unit": "Mbit/s", "cores": 1, "x_realtime": v1 * 1e6 / (dsd_fs * nch)}}
        e2e_value = bits_per_step * args.steps * world / e2e_s / 1e6
        print(json.dumps({
            "metric": "dsd_mbit_per_s", "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic" if DATA_KIND == "sdm" else "synthetic (uniform random bytes)",
            "per_gpu_mbit_per_s": value / world, "x_realtime": value * 1e6 / (dsd_fs * nch) / 1.0,
            "config": {"workload": desc, "dither": bool(args.dither), "lut": f"fp{args.lut}",
                       "kernel": args.kernel or "auto", "frames_per_step": nframes, "channels": nch,
                       "l2": f"ring of {ring} distinct in/out buffer sets ({(in_bytes + out_bytes) * ring >> 20} MiB) > L2",
                       "parallelism": f"{world} x independent streams, no collective"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mbit/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": e2e_s / args.steps * 1e3, "x_realtime": e2e_value * 1e6 / (dsd_fs * nch),
                    "api": "dsdgpu_decimate_host (pinned host buffers, chunks pipelined over the context's slots)",
                    "checksum": checksum, "link": link},
            "plugin": plugin, "gpu_launches": launches, "clocks": clocks,
        }), flush=True)
    pin_in.free()
    pin_out.free()
    g.close()
    if world > 1:
        dist.destroy_process_group()


def run_sharded_arm(args, spec):
    """Strong scaling: one fixed job (a long stream, a multi-channel stream or a batch of files) cut
    by dsdgpu_plan_shards into equal-cost shards, one process per GPU, no data-path collective.
    value = device-resident throughput of the whole job; e2e = the same with pinned host buffers."""
    import torch
    import torch.distributed as dist
    from dsf2flac_b200 import batch, capi, filters

    desc, files, lut_bits, dither = spec
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    firs = [filters.pick(f[0], f[1], extensions=True) for f in files]
    frames = [filters.app_frames(f[0] * f[3], fir) for f, fir in zip(files, firs)]
    shards = capi.plan_shards(frames, [f[2] for f in files], [fir.nlut for fir in firs], rank, world, 1024)
    base = {}                                              # one second of modulator output per (rate, channel)

    def second(fs, c):
        if (fs, c) not in base:
            base[(fs, c)] = synth_stream(fs, c + 1, 1)[c]
        return base[(fs, c)]

    runner = batch.BatchRunner(device=local)
    work, total_bits = [], sum(f[0] * f[2] * f[3] for f in files)
    for (j, c0, k, f0, nf) in shards:
        fs, out_fs, nch, seconds = files[j]
        fir = firs[j]
        sec = fs // 8
        lo = max(0, fir.nlut + f0 * fir.nstep - (fir.nlut - 1))
        hi = min(sec * seconds, fir.nlut + (f0 + nf - 1) * fir.nstep + 1)
        lo -= (lo - fir.nlut - f0 * fir.nstep - 1) % 16      # frame grid 16-byte aligned to the window
        lo = max(lo, 0)
        pin = capi.PinnedBuffer((k, hi - lo), np.uint8)
        for c in range(k):                                   # the window [lo, hi) of the one-second pattern, tiled
            one, phase = second(fs, c0 + c), lo % sec
            pin.array[c] = np.tile(one, (hi - lo + phase) // sec + 2)[phase:phase + hi - lo]
        job = batch.Job(pin.array, fs, out_fs, msb_flag=True, length_samples=fs * seconds, bits=24, dither=bool(dither),
                        lut_precision=lut_bits, first_byte=lo, extensions=True)
        job_nch = nch
        out = capi.PinnedBuffer((nf, k), np.int32)
        ctx = runner.ctx_for(job, k)
        # device rows on a 256-byte pitch: the kernels stage 16-byte chunks with cp.async when rows are aligned
        d_in = torch.empty((k, (hi - lo + 255) // 256 * 256), dtype=torch.uint8, device="cuda")
        d_in[:, :hi - lo] = torch.from_numpy(pin.array).cuda()
        d_out = torch.empty((nf, k), dtype=torch.int32, device="cuda")
        work.append((job, (j, c0, k, f0, nf), pin, out, ctx, d_in, d_out, job_nch))
    scale_of = lambda job: filters.app_params(job.bits, job.scale_db, job.dither)
    stream = torch.cuda.Stream()

    def device_step():
        for job, (j, c0, k, f0, nf), pin, out, ctx, d_in, d_out, job_nch in work:
            ctx.set_option("dither_stride", job_nch)
            scale, amp, clip = scale_of(job)
            ctx.decimate_device(d_in, d_in.shape[1], job.first_byte, job.in_count, job.idle, job.fir.nlut + f0 * job.fir.nstep,
                                nf, scale, amp, clip, f0 * job_nch + c0, capi.OUT_INT32, d_out, stream=stream.cuda_stream)

    def host_step():
        for job, shard, pin, out, ctx, d_in, d_out, job_nch in work:
            j, c0, k, f0, nf = shard
            ctx.set_option("dither_stride", job_nch)
            scale, amp, clip = scale_of(job)
            ctx.decimate_host_raw(pin.ptr, pin.array.shape[1], job.first_byte, job.in_count, job.idle,
                                  job.fir.nlut + f0 * job.fir.nstep, nf, scale, amp, clip, f0 * job_nch + c0, capi.OUT_INT32, out.ptr)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        device_step()
    stream.synchronize()
    barrier()
    launches0 = runner.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        device_step()
    e1.record(stream)
    stream.synchronize()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = runner.launches - launches0
    for _ in range(args.warmup):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop(0.0, time.perf_counter()) if sampler else None
    h2d = sum(w[2].array.nbytes for w in work)
    d2h = sum(w[3].array.nbytes for w in work)
    if rank == 0:
        value = total_bits * args.steps / (dev_ms * 1e-3) / 1e6
        e2e_value = total_bits * args.steps / e2e_s / 1e6
        audio_s = sum(f[3] for f in files)
        print(json.dumps({
            "metric": "dsd_mbit_per_s", "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if lut_bits == 64 else "i32", "data": "synthetic",
            "per_gpu_mbit_per_s": value / world, "x_realtime": audio_s / (dev_ms * 1e-3 / args.steps),
            "config": {"workload": desc, "files": len(files), "dither": bool(dither), "lut": f"{lut_bits}-bit",
                       "shards_rank0": len(work), "l2": "per-step working set > L2" if h2d + d2h > L2_BYTES else "working set may fit L2",
                       "parallelism": f"{world} ranks, dsdgpu_plan_shards (file / channel / time segment + halo), no collective"},
            "roofline": None, "cpu_baseline": None,
            "e2e": {"value": e2e_value, "unit": "Mbit/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3, "x_realtime": audio_s / (e2e_s / args.steps),
                    "api": "dsdgpu_decimate_host per shard (pinned host buffers)", "note": "bytes are rank 0's"},
            "gpu_launches": launches, "clocks": clocks,
        }), flush=True)
    runner.close()
    if world > 1:
        dist.destroy_process_group()


def run_dop_arm(args):
    """DoP packing: DSD64 stereo 60 s -> 176.4 kHz DoP frames.  An HBM-bound byte
    shuffle: 2 bytes in + 4 bytes out per output sample."""
    import torch
    from dsf2flac_b200 import capi
    from tests.oracle_lib import Oracle, REF_SO
    dsd_fs, nch, seconds = DSD64, 2, args.dop_seconds
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(local)
    data = synth_stream(dsd_fs, nch, seconds)
    nbytes = data.shape[1]
    nframes = nbytes // 2
    g = capi.DsdGpu(np.zeros((1, 256)), nch, 1, device=local)
    # A ring of buffer sets several times the L2 (127 MB each): a step finds none of its input in L2, and the
    # dirty lines it leaves behind are written back during the next step, as the previous step's are during
    # this one -- steady state, DRAM traffic per step = the algorithmic bytes.  (Flushing L2 by WRITING a
    # buffer between steps would charge the write-back of that buffer to the timed kernel.)
    ring = max(2, 6 * 60 // seconds)
    d_in = [torch.from_numpy(np.roll(data, 4096 * k, axis=1)).cuda() for k in range(ring)]
    d_out = [torch.empty((nframes, nch), dtype=torch.int32, device="cuda") for _ in range(ring)]
    stream = torch.cuda.Stream()
    times = []
    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        for k in range(args.warmup + args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.dop_pack_device(d_in[k % ring], nbytes, 0, nbytes, 0x69, 0, nframes, 1, d_out[k % ring], stream=stream.cuda_stream)
            e1.record(stream)
            times.append((e0, e1))
    stream.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in times[args.warmup:])
    kernel_ms = ms[len(ms) // 2]
    pin_in = capi.PinnedBuffer((nch, nbytes), np.uint8)
    pin_in.array[:] = data
    pin_out = capi.PinnedBuffer((nframes, nch), np.int32)
    lib = g.lib
    for _ in range(args.warmup):
        lib.dsdgpu_dop_pack_host(g.h, pin_in.ptr, nbytes, 0, nbytes, 0x69, 0, nframes, 1, pin_out.ptr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lib.dsdgpu_dop_pack_host(g.h, pin_in.ptr, nbytes, 0, nbytes, 0x69, 0, nframes, 1, pin_out.ptr)
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop(0.0, time.perf_counter())
    orc = Oracle(reference=os.path.exists(REF_SO))
    sub = np.ascontiguousarray(data[:, :dsd_fs // 8 * 5])
    t0 = time.perf_counter()
    want = orc.dop_run_raw(sub, 1 << 40, 1, 1, sub.shape[1] // 2 - 1)
    cpu_s = time.perf_counter() - t0
    assert np.array_equal(pin_out.array[:len(want)], want), "DoP frames differ from the CPU packer"
    peaks, peak_src = measured_peaks()
    bits = nbytes * 8 * nch
    hbm_bytes = nframes * nch * 6
    print(json.dumps({
        "metric": "dsd_mbit_per_s", "value": bits / (kernel_ms * 1e-3) / 1e6, "unit": "Mbit/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"DSD64 stereo {seconds} s -> DoP frames at 176.4 kHz (reference dop_packer.cpp)",
                   "l2": f"ring of {ring} distinct in/out buffer sets ({ring * 127 * seconds // 60} MB) > L2"},
        "roofline": {"bound": "hbm", "achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                     "algorithmic_bytes_per_sample": 6, "samples_per_launch": nframes * nch, "peak_source": peak_src},
        "cpu_baseline": {"value": sub.size * 8 / cpu_s / 1e6, "unit": "Mbit/s", "cores": 1, "kind": "reference" if orc.reference else "port",
                         "sample": "first 5 s, DopPacker::pack_buffer over an in-memory reader"},
        "e2e": {"value": bits * args.steps / e2e_s / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": nch * nbytes,
                "d2h_bytes_per_step": nframes * nch * 4, "ms_per_step": e2e_s / args.steps * 1e3, "api": "dsdgpu_dop_pack_host"},
        "gpu_launches": args.steps, "clocks": clocks,
    }), flush=True)
    g.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3", "c4", "c4x", "c5", "dop"])
    ap.add_argument("--lut", type=int, default=64, choices=[64, 32])
    ap.add_argument("--dither", type=int, default=0)
    ap.add_argument("--kernel", default="", choices=["", "auto", "direct", "ring", "seg"])
    ap.add_argument("--align", type=int, default=1, help="0: device buffer starts at stream byte 0 (frame grid off by one byte)")
    ap.add_argument("--files", type=int, default=256, help="files in the c5 batch (about 120 MB of pinned host memory each)")
    ap.add_argument("--slots", type=int, default=0, help="pipelined slots of the host entry point (1..4)")
    ap.add_argument("--chunk-frames", type=int, default=0, help="frames per pipelined host chunk")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall seconds of CPU work in the cpu_baseline sample (0: skip)")
    ap.add_argument("--plugin-seconds", type=float, default=2.0)
    ap.add_argument("--data", default="sdm", choices=["sdm", "random"], help="synthetic input: sigma-delta sines or uniform random bytes")
    ap.add_argument("--seconds", type=int, default=0, help="stream length of the unsharded workloads (default: 60)")
    ap.add_argument("--dop-seconds", type=int, default=600, help="stream length of the dop workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global DATA_KIND
    DATA_KIND = args.data
    if args.workload == "dop":
        return run_dop_arm(args)
    if args.workload in ("c3", "c4", "c4x", "c5"):
        if args.impl == "reference":
            return run_reference_arm_batch(args, sharded_workloads(args.files)[args.workload])
        return run_sharded_arm(args, sharded_workloads(args.files)[args.workload])
    wl = WORKLOADS[args.workload]
    if args.seconds:
        wl = wl[:3] + (args.seconds, wl[4], wl[5].replace(" 60 s", f" {args.seconds} s"))
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
