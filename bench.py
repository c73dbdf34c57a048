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
    # name: (dsd_fs, out_fs, nch, seconds, bits, description, name)
    "c1": (DSD64, 88200, 2, 60, 24, "DSD64 stereo 60 s -> 24-bit/88.2 kHz (BASELINE configs[0])", "c1"),
    "c2": (2 * DSD64, 176400, 2, 60, 24, "DSD128 stereo 60 s -> 24-bit/176.4 kHz, fp64 LUT (BASELINE configs[1])", "c2"),
    # the other two filters of the reference (dsd_decimator.cpp:57-65): same input, higher output rates
    "d16": (DSD64, 176400, 2, 60, 24, "DSD64 stereo 60 s -> 24-bit/176.4 kHz (divide-by-16 filter, 30 tables)", "d16"),
    "d8": (DSD64, 352800, 2, 60, 24, "DSD64 stereo 60 s -> 24-bit/352.8 kHz (divide-by-8 filter, 12 tables)", "d8"),
    # EXTENSION filters (no reference counterpart): BASELINE configs[3] as literally stated
    "x64": (8 * DSD64, 352800, 2, 60, 24, "DSD512 stereo 60 s -> 24-bit/352.8 kHz (divide-by-64 EXTENSION filter, 144 tables, "
                                          "two passes of the ring kernel)", "x64"),
    "x128": (8 * DSD64, 176400, 2, 60, 24, "DSD512 stereo 60 s -> 24-bit/176.4 kHz (divide-by-128 EXTENSION filter, 288 tables)", "x128"),
    "x256": (8 * DSD64, 88200, 2, 60, 24, "DSD512 stereo 60 s -> 24-bit/88.2 kHz (divide-by-256 EXTENSION filter, 576 tables)", "x256"),
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


_SYNTH = None


def synth_stream(dsd_fs, nch, seconds, lsb_first=1):
    """One second of 2nd-order sigma-delta sine per channel (host/dsd_synth.cpp built on its own as
    tests/cpp/_build/libdsdsynth.so -- the input generator is not the product: neither arm maps a product library to
    make its input), tiled; or uniform random bytes (shared-memory behaviour depends on the byte statistics,
    so both are reported)."""
    if DATA_KIND == "random":
        return np.random.default_rng(1234).integers(0, 256, size=(nch, dsd_fs // 8 * seconds), dtype=np.uint8)
    global _SYNTH
    if _SYNTH is None:
        import ctypes as C
        path = os.path.join(ROOT, "tests", "cpp", "_build", "libdsdsynth.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ROOT, "tests/cpp/_build/libdsdsynth.so"], check=True, capture_output=True)
        _SYNTH = C.CDLL(path)
        _SYNTH.dsdhost_synth_sdm.restype = None
        _SYNTH.dsdhost_synth_sdm.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int]
    sec = dsd_fs // 8
    one = np.empty((nch, sec), dtype=np.uint8)
    for c in range(nch):
        _SYNTH.dsdhost_synth_sdm(one[c].ctypes.data, sec, float(dsd_fs), 1000.0 + 500.0 * c, 0.5, lsb_first)
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
    seg_bytes, k, dsd_fs, out_fs, nlut, nstep, scale, amp, clip, reps = args
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
        orc.run_raw(sub, 1 << 40, 1, dsd_fs, out_fs, nlut + 1, nframes, scale, amp, clip)
    return reps * nframes * nstep * 8 * data.shape[0], time.perf_counter() - t0


def cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, pool=None, reps=1, amp=0.0):
    """The reference's CPU decimator (getSamples loop only, like the GPU arm: no FLAC, no file
    parsing) over the first `sample_seconds` of the stream, `reps` times over, one process per core
    (the reference is single threaded and keeps process-global state, so cores = processes).
    -> (Mbit/s, bits, kind)"""
    from tests.oracle_lib import REF_SO
    nbytes = min(data.shape[1], int(dsd_fs // 8 * sample_seconds))
    seg = nbytes // cores // 4 * 4
    jobs = [(seg, k, dsd_fs, out_fs, fir.nlut, fir.nstep, scale, amp, clip, reps) for k in range(cores)]
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


def _reference_case(wl, dither, cores, budget_s):
    """A bounded sample of one unsharded workload on every host core -> (Mbit/s, sample text, kind)."""
    from dsf2flac_b200 import filters
    dsd_fs, out_fs, nch, seconds, bits_depth, desc = wl[:6]
    fir = filters.pick(dsd_fs, out_fs, extensions=True)
    scale, amp, clip = filters.app_params(bits_depth, dither=bool(dither))
    data = synth_stream(dsd_fs, nch, seconds)
    probe, _, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, 0.5, amp=amp)
    sample_seconds = max(0.25, min(seconds, budget_s * probe * 1e6 / (dsd_fs * nch)))
    v, _, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, amp=amp)
    return v, f"first {sample_seconds:.2f} s of the stream, split over {cores} processes", kind


def run_reference_arm(args, wl):
    dsd_fs, out_fs, nch, seconds, bits_depth, desc = wl[:6]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dsf2flac_b200 import filters
    fir = filters.pick(dsd_fs, out_fs, extensions=True)
    scale, amp, clip = filters.app_params(bits_depth, dither=bool(args.dither))
    cores = len(os.sched_getaffinity(0))
    global _CPU_DATA
    data = _CPU_DATA = synth_stream(dsd_fs, nch, seconds)
    pool = mp.get_context("fork").Pool(cores)
    # a step is the whole file, like the GPU arm's, unless that would push the run past ~2 minutes
    probe, _, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, 1.0, pool, amp=amp)
    budget_s = 100.0 / max(1, args.steps + args.warmup)
    sample_seconds = max(0.25, min(data.shape[1] * 8 / dsd_fs, budget_s * probe * 1e6 / (dsd_fs * nch)))
    for _ in range(args.warmup):
        cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, pool, amp=amp)
    t0 = time.perf_counter()
    total_bits = 0
    for _ in range(args.steps):
        _, b, _ = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, sample_seconds, pool, amp=amp)
        total_bits += b
    wall = time.perf_counter() - t0
    pool.close()
    v = total_bits / wall / 1e6
    sample = f"first {sample_seconds:.2f} s of the stream per step, split over {cores} processes"
    line = {
        "impl": "reference", "metric": "dsd_mbit_per_s", "value": v, "unit": "Mbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "x_realtime": v * 1e6 / (dsd_fs * nch),
        "config": {"workload": desc, "dither": bool(args.dither), "lut": "fp64"},
        "cpu_baseline": {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.subs and args.workload == "c2":
        # the same sub-objects the GPU arm prints: BASELINE configs[0] (the config the metric's x realtime is quoted on) with
        # dither off and on (on = the reference's default, main.cpp:240-244), a few seconds of CPU work each
        c1 = WORKLOADS["c1"]
        sub = {}
        for name, dither in (("dither_off", 0), ("dither_on", 1)):
            sv, ssample, skind = _reference_case(c1, dither, cores, 4.0)
            sub[name] = {"value": sv, "unit": "Mbit/s", "x_realtime": sv * 1e6 / (c1[0] * c1[2]), "cores": cores, "kind": skind,
                         "sample": ssample}
        line["configs0"] = {"workload": c1[5], **sub}
        if args.gpus > 1:
            spec = sharded_workloads(SUB_C5_FILES)["c5"]
            sv, take, njobs, skind, _ = reference_batch_rate(spec, cores, 15.0, 1, 0)
            line["sharded"] = {"c5": {"workload": spec[0], "value": sv, "unit": "Mbit/s", "cores": cores, "kind": skind,
                                      "sample": f"the first {take} of {njobs} files, one reference process per core"}}
    print(json.dumps(line), flush=True)


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


def reference_batch_rate(spec, cores, budget_s, steps, warmup):
    """The files of a sharded job (time segments of one long stream count as files of their own) converted by
    independent reference processes, one per host core; a step is a bounded sample: the first files of the list.
    -> (Mbit/s, files per step, files in all, kind, wall seconds of the timed steps)"""
    desc, files, lut_bits, dither = spec
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
    budget_bits = rate * budget_s
    take, acc = 0, 0
    while take < len(jobs) and (take < cores or acc < budget_bits):
        acc += jobs[take][0] * jobs[take][2] * jobs[take][3]
        take += 1
    sample = jobs[:take]
    for _ in range(warmup):
        pool.map(_cpu_file_worker, sample, chunksize=1)
    t0 = time.perf_counter()
    bits = 0
    for _ in range(steps):
        bits += sum(r[0] for r in pool.map(_cpu_file_worker, sample, chunksize=1))
    wall = time.perf_counter() - t0
    pool.close()
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdsdref.so")) else "port"
    return bits / wall / 1e6, take, len(jobs), kind, wall


def run_reference_arm_batch(args, spec):
    """Reference arm of the sharded workloads: the reference is single threaded and its only batch notion is a serial
    shell loop (scripts/batch_dsf2flac.sh:16-19), so: one reference process per host core over the file list."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    desc, files, lut_bits, dither = spec
    cores = len(os.sched_getaffinity(0))
    v, take, njobs, kind, wall = reference_batch_rate(spec, cores, 100.0 / max(1, args.steps + args.warmup), args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": "dsd_mbit_per_s", "value": v, "unit": "Mbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "files": len(files), "dither": bool(dither), "lut": "fp64 (the reference has no 32-bit variant)"},
        "cpu_baseline": {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind,
                         "sample": f"the first {take} of {njobs} files / 60 s pieces per step, one reference process per core"},
        "e2e": {"value": v, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---- GPU arm ----------------------------------------------------------------------------------

SUB_C5_FILES = 32           # files of the c5 slice in the default line's `sharded` sub-object


class Dist:
    """One process per GPU; torch.distributed (NCCL) carries only the barrier and the max-over-ranks of the timings."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the GPU path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def smem_peak_gbs(dist):
    peaks, peak_src = measured_peaks()
    n_sm = dist.torch.cuda.get_device_properties(dist.local).multi_processor_count
    return n_sm * SMEM_BYTES_PER_CLK_PER_SM * peaks["sm_max_mhz"] * 1e6 / 1e9, n_sm, peaks, peak_src


def stream_case(dist, args, wl, dither, lut_bits, steps, warmup, link=False, plugin_seconds=0.0, cpu_seconds=0.0,
                out_kind_name="int32", dither_cache=False):
    """One unsharded workload on every rank (weak scaling): device-resident throughput (CUDA events on the launch stream),
    the same through dsdgpu_decimate_host with pinned host buffers, the roofline of the decimation kernel, and on request
    the link's copy-only floor (rank 0 alone and all ranks at once), the drop-in class and the CPU baseline."""
    torch = dist.torch
    from dsf2flac_b200 import capi, filters
    dsd_fs, out_fs, nch, seconds, bits_depth, desc = wl[:6]
    rank, world, local = dist.rank, dist.world, dist.local
    fir = filters.pick(dsd_fs, out_fs, extensions=True)
    lut = filters.build_lut(fir, 1)
    scale, amp, clip = filters.app_params(bits_depth, dither=bool(dither))
    data = synth_stream(dsd_fs, nch, seconds)
    nbytes = data.shape[1]
    nframes = filters.app_frames(nbytes * 8, fir)
    bits_per_step = nbytes * 8 * nch                       # DSD bits consumed per step per GPU
    g = capi.DsdGpu(lut, nch, fir.nstep, device=local, lut_precision=lut_bits)
    if args.kernel:
        g.set_option("kernel", {"auto": 0, "direct": 1, "ring": 3, "seg": 4}[args.kernel])
    if args.slots:
        g.set_option("slots", args.slots)
    if args.chunk_frames:
        g.set_option("chunk_frames", args.chunk_frames)
    # The context's dither plane cache (the reference's dither stream is a constant: unseeded rand(), one file per process)
    # is OFF for every measurement except the sub-object that names it: with it off each step regenerates the plane of
    # its file, as the first conversion of a process does.
    g.set_option("dither_cache", (1 << 27) if dither_cache else 0)
    out_kind = {"int32": capi.OUT_INT32, "pcm24": capi.OUT_PCM24, "pcm16": capi.OUT_PCM16}[out_kind_name]
    out_elem = {"int32": 4, "pcm24": 3, "pcm16": 2}[out_kind_name]

    # device-resident ring of distinct input/output buffers, larger than L2 in total, so no step
    # finds its bytes in L2 from the step before
    # The stream sits in its device buffer behind `front` reader-idle bytes (what the reader's ring
    # holds before byte 0, dsd_sample_reader.h:122), chosen so that the frame grid is 4-byte aligned
    # to the buffer ((p0 - in_first + 1) % 4 == 0): the kernel's window loads are then conflict free.
    in_bytes, out_bytes = nch * nbytes, nframes * nch * out_elem
    ring = max(2, -(-2 * L2_BYTES // (in_bytes + out_bytes)))
    front = (fir.nlut + 1) % 4 if args.align else 0
    front = (4 - front) % 4
    stride = (nbytes + front + 255) // 256 * 256
    d_in = []
    for k in range(ring):
        t = torch.full((nch, stride), 0x69, dtype=torch.uint8)
        t[:, front:front + nbytes] = torch.from_numpy(np.roll(data, 4096 * k, axis=1))
        d_in.append(t.cuda())
    d_out = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(ring)]
    stream = torch.cuda.Stream()

    def step(k):
        g.decimate_device(d_in[k % ring], stride, -front, nbytes + front, 0x69, fir.nlut, nframes, scale, amp, clip, 0,
                          out_kind, d_out[k % ring], stream=stream.cuda_stream)

    with torch.cuda.stream(stream):
        for k in range(warmup):
            step(k)
    dist.barrier()
    launches0 = g.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for k in range(steps):
            step(k)
            ev[k + 1].record(stream)
    stream.synchronize()
    dist.barrier()
    launches = g.launches - launches0
    dev_ms = ev[0].elapsed_time(ev[-1])
    per_step = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(steps))
    total_ms = dist.max_over_ranks(dev_ms)
    del d_in, d_out

    # ---- end to end: host bytes -> PCM on the host, through the C ABI with host buffers -------
    pin_in = capi.PinnedBuffer((nch, nbytes), np.uint8)
    pin_in.array[:] = data
    pin_out = capi.PinnedBuffer((nframes * nch * out_elem,), np.uint8)

    def e2e_step():
        g.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, nframes, scale, amp, clip, 0, out_kind, pin_out.ptr)

    for _ in range(warmup):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()                                         # synchronous: returns when the PCM is on the host
    torch.cuda.synchronize()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0)
    dist.barrier()
    view = pin_out.array.view(np.int32) if out_elem == 4 else pin_out.array
    checksum = int(view[:: max(1, view.size // 65536)].astype(np.int64).sum())

    # ---- what the link allows: pinned copies of the same sizes, one way and both ways at once --
    link_info, link_n = None, None
    if link:
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

        if rank == 0:
            t_up, t_down = timed(up), timed(down)
            t_both = timed(lambda: (up(), down()))
            link_info = {"h2d_gbs": in_bytes / t_up / 1e9, "d2h_gbs": out_bytes / t_down / 1e9, "duplex_ms": t_both * 1e3,
                         "note": "rank 0 alone: pinned cudaMemcpyAsync of one step's input / output; duplex = both at once"}
        # every rank at once (barrier-aligned): what N ranks' copies get out of this host's memory path together
        dist.barrier()
        t_all = dist.max_over_ranks(timed(lambda: (up(), down()), reps=8))
        dist.barrier()
        link_n = {"ranks": world, "duplex_ms": t_all * 1e3, "aggregate_gbs": world * (in_bytes + out_bytes) / t_all / 1e9,
                  "per_rank_gbs_each_way": (in_bytes + out_bytes) / 2 / t_all / 1e9,
                  "note": "all ranks copy one step's input up and output down at the same time; max over ranks"}
        del dv_in, dv_out

    # ---- the drop-in class itself (reader drained byte by byte, as dsf2flac would run it) -----
    plugin = None
    if rank == 0 and plugin_seconds > 0 and out_elem == 4:
        host = capi.host_lib()
        host.dsdhost_set_extension_filters(0 if fir.ratio in filters.FILTERS else 1)
        pb = min(nbytes, int(dsd_fs // 8 * plugin_seconds))
        cap = nframes + 2048
        out = capi.PinnedBuffer((cap, nch), np.int32)
        plugin = {}
        # (a) as dsf2flac runs it today: a reader stepped one byte per channel at a time
        # (b) the same class with the file's bytes handed over in bulk (DsdDecimator::setBulkSource)
        for name, fn, count in (("stepped_reader", host.dsdhost_run_app, pb), ("bulk_source", host.dsdhost_run_app_bulk, nbytes)):
            best = None
            for _ in range(3):                              # the first call pays context creation and page-locking
                t0 = time.perf_counter()
                # out = NULL: main.cpp's loop as it is -- one 1024-frame buffer, reused, read by its consumer (the harness
                # would otherwise add a copy of the whole PCM stream into a second array, which dsf2flac never makes)
                n = fn(pin_in.ptr, nch, nbytes, count * 8, 0x69, 1, dsd_fs, out_fs, scale, amp, clip, lut_bits, 0, None, cap)
                dt = time.perf_counter() - t0
                if n > 0 and (best is None or dt < best):
                    best = dt
            if best:
                plugin[name] = {"value": count * 8 * nch / best / 1e6, "unit": "Mbit/s", "stream_seconds": count * 8 / dsd_fs,
                                "wall_ms": best * 1e3}
        plugin["note"] = ("drop-in DsdDecimator::getSamples in main.cpp's 1024-frame calls, decimator construction "
                          "(tables + GPU context) included")
        out.free()

    smem_peak, n_sm, peaks, peak_src = smem_peak_gbs(dist)
    value = bits_per_step * steps * world / (total_ms * 1e-3) / 1e6
    kernel_ms = per_step[len(per_step) // 2]
    samples_per_launch = nframes * nch
    lut_elem = 8 if lut_bits == 64 else 4
    smem_bytes = samples_per_launch * fir.nlut * lut_elem          # algorithmic LUT bytes per launch
    achieved = smem_bytes / (kernel_ms * 1e-3) / 1e9
    hbm_bytes = samples_per_launch * (fir.nstep + out_elem)
    roof = {"bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
            "traffic": None, "peak_source": f"{n_sm} SMs x 128 B/clk x sm_max_mhz ({peak_src})",
            "algorithmic_bytes_per_sample": fir.nlut * lut_elem, "samples_per_launch": samples_per_launch,
            "kernel_ms": kernel_ms,
            "hbm": {"bound": "hbm", "achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "algorithmic_bytes_per_sample": fir.nstep + out_elem, "peak_source": peak_src}}
    try:                                               # measured DRAM traffic of the same kernel on the same workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "ring_dither" if dither else (args.kernel or "ring")
        tr = tr[key]
        if tr["workload"] == wl[6] and fir.ratio == 32 and lut_bits == 64 and out_elem == 4:
            roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roof["traffic_source"] = tr["source"]
            roof["algorithmic_hbm_bytes"] = hbm_bytes
    except Exception:
        pass
    cpu = None
    if rank == 0 and world == 1 and cpu_seconds > 0:
        # about cpu_seconds of wall time on every core: the whole file, as often as that takes
        cores = len(os.sched_getaffinity(0))
        probe, _, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, seconds, amp=amp)
        reps = max(1, int(cpu_seconds * probe * 1e6 / bits_per_step))
        v, b, kind = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, cores, seconds, reps=reps, amp=amp)
        v1, b1, _ = cpu_reference_rate(data, dsd_fs, out_fs, fir, scale, clip, 1, seconds / 2, amp=amp)
        cpu = {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind,
               "sample": f"the whole {seconds} s file {reps} times over ({b / 1e9:.1f} Gbit), one process per core "
                         f"(the reference is single threaded), decimator only",
               "one_core": {"value": v1, "unit": "Mbit/s", "cores": 1, "x_realtime": v1 * 1e6 / (dsd_fs * nch)}}
    e2e_value = bits_per_step * steps * world / e2e_s / 1e6
    e2e = {"value": e2e_value, "unit": "Mbit/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
           "ms_per_step": e2e_s / steps * 1e3, "x_realtime": e2e_value * 1e6 / (dsd_fs * nch),
           "api": "dsdgpu_decimate_host (pinned host buffers, chunks pipelined over the context's slots)",
           "checksum": checksum}
    if link:
        e2e["link"] = link_info
        e2e["link_n"] = link_n
        if link_n:
            e2e["frac_of_link_n_floor"] = link_n["duplex_ms"] / e2e["ms_per_step"]
    pin_in.free()
    pin_out.free()
    g.close()
    return {"workload": desc, "dither": bool(dither), "lut": f"fp{lut_bits}" if lut_bits == 64 else "32-bit (Q30)", "out": out_kind_name,
            "value": value, "unit": "Mbit/s", "ms_per_step": total_ms / steps,
            "per_gpu_mbit_per_s": value / world, "x_realtime": value * 1e6 / (dsd_fs * nch), "frames_per_step": nframes, "channels": nch,
            "l2": f"ring of {ring} distinct in/out buffer sets ({(in_bytes + out_bytes) * ring >> 20} MiB) > L2",
            "roofline": roof, "e2e": e2e, "plugin": plugin, "cpu_baseline": cpu, "gpu_launches": launches}


def run_gpu_arm(args, wl):
    dist = Dist()
    sampler = ClockSampler(dist.local) if dist.rank == 0 else None     # polls every 100 ms from here to the end
    t_wall0 = time.perf_counter()
    main = stream_case(dist, args, wl, args.dither, args.lut, args.steps, args.warmup, link=True,
                       plugin_seconds=args.plugin_seconds, cpu_seconds=args.cpu_seconds, out_kind_name=args.out,
                       dither_cache=bool(args.dither_cache))
    extra = {}
    if args.subs and args.workload == "c2":
        # BASELINE configs[0] -- the config the metric's "x realtime" is quoted on -- with dither off (the parity setting)
        # and on (the reference's default, main.cpp:240-244); the same measurement, a sub-object each
        c1 = WORKLOADS["c1"]
        sub = {}
        for name, dither in (("dither_off", 0), ("dither_on", 1)):
            r = stream_case(dist, args, c1, dither, args.lut, args.steps, args.warmup)
            sub[name] = {k: r[k] for k in ("dither", "value", "unit", "ms_per_step", "per_gpu_mbit_per_s", "x_realtime", "roofline", "e2e",
                                           "gpu_launches")}
        extra["configs0"] = {"workload": c1[5], **sub}
        # ... and configs[1] itself with dither on, and with the PCM handed back as packed 24-bit samples (SURVEY 8f-4:
        # the same integers, 3 bytes each over PCIe instead of the 4 of a FLAC__int32)
        r = stream_case(dist, args, wl, 1, args.lut, args.steps, args.warmup)
        extra["dither_on"] = {k: r[k] for k in ("dither", "value", "unit", "ms_per_step", "x_realtime", "roofline", "e2e", "gpu_launches")}
        # ... with the context's dither plane cache on: what every file after the first costs in a batch of one process
        # (the generator then runs in the warm-up steps only; `dither_on` above regenerates the plane in every step)
        r = stream_case(dist, args, wl, 1, args.lut, args.steps, args.warmup, dither_cache=True)
        extra["dither_cached"] = {"note": "dither plane kept by the context (dsdgpu_set_option dither_cache): generated during warm-up, read in place by the timed steps",
                                  **{k: r[k] for k in ("dither", "value", "unit", "ms_per_step", "x_realtime", "roofline", "gpu_launches")}}
        r = stream_case(dist, args, wl, args.dither, args.lut, args.steps, args.warmup, out_kind_name="pcm24")
        extra["packed24"] = {k: r[k] for k in ("out", "value", "unit", "ms_per_step", "roofline", "e2e", "gpu_launches")}
        if dist.world > 1:
            # the channel / time / file splits BASELINE names (strong scaling of one fixed job over the ranks)
            sh = {}
            specs = sharded_workloads(SUB_C5_FILES)
            for name in ("c5", "c4x"):
                r = sharded_case(dist, args, specs[name], max(3, args.steps // 4), 3)
                sh[name] = r
            extra["sharded"] = sh
    clocks = sampler.stop(t_wall0 - 3600.0, time.perf_counter()) if sampler else None
    if dist.rank == 0:
        if clocks and clocks.get("sm_mhz"):
            smem_peak, n_sm, _, _ = smem_peak_gbs(dist)
            main["roofline"]["frac_at_sampled_clock"] = main["roofline"]["achieved"] / (
                n_sm * SMEM_BYTES_PER_CLK_PER_SM * clocks["sm_mhz"] * 1e6 / 1e9)
        line = {
            "metric": "dsd_mbit_per_s", "value": main["value"], "unit": "Mbit/s", "n_gpus": dist.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.lut == 64 else "i32",
            "data": "synthetic" if DATA_KIND == "sdm" else "synthetic (uniform random bytes)",
            "per_gpu_mbit_per_s": main["per_gpu_mbit_per_s"], "x_realtime": main["x_realtime"],
            "config": {"workload": main["workload"], "dither": bool(args.dither), "dither_cache": bool(args.dither_cache), "lut": main["lut"], "out": main["out"],
                       "kernel": args.kernel or "auto", "frames_per_step": main["frames_per_step"], "channels": main["channels"],
                       "l2": main["l2"], "parallelism": f"{dist.world} x independent streams, no collective"},
            "roofline": main["roofline"], "cpu_baseline": main["cpu_baseline"], "e2e": main["e2e"], "plugin": main["plugin"],
            "gpu_launches": main["gpu_launches"], "clocks": clocks,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    dist.close()


def sharded_case(dist, args, spec, steps, warmup):
    """Strong scaling: one fixed job (a long stream, a multi-channel stream or a batch of files) cut
    by dsdgpu_plan_shards into equal-cost shards, one process per GPU, no data-path collective.
    value = device-resident throughput of the whole job; e2e = the same with pinned host buffers."""
    torch = dist.torch
    from dsf2flac_b200 import batch, capi, filters

    desc, files, lut_bits, dither = spec
    rank, world, local = dist.rank, dist.world, dist.local
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
        ctx.set_option("dither_cache", 0)                    # every step regenerates the dither planes of its files (see stream_case)
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

    # End to end: this rank's shards as ONE dsdgpu_run_batch call (the C entry point: chunks of neighbouring shards and
    # files pipelined through the slots, contexts kept across calls).  Shards that hold only some channels of their job
    # (configs[2]) need the job's dither stride, which a job of their own cannot express: those go shard by shard.
    luts = {}
    cjobs, singles = [], []
    for item in work:
        job, (j, c0, k, f0, nf), pin, out, ctx, d_in, d_out, job_nch = item
        if k != job_nch:
            singles.append(item)
            continue
        key = (job.fir.ratio, job.lut_precision)
        if key not in luts:
            luts[key] = np.ascontiguousarray(filters.build_lut(job.fir, job.msb_flag))
        scale, amp, clip = scale_of(job)
        cjobs.append(dict(in_=pin.ptr, in_stride=pin.array.shape[1], in_first=job.first_byte, in_count=job.in_count,
                          source_block_bytes=0, fill=job.idle, nch=k, nlut=job.fir.nlut, nstep=job.fir.nstep,
                          lut_precision=job.lut_precision, lut=luts[key], p0=job.fir.nlut + f0 * job.fir.nstep, nframes=nf,
                          scale=scale, dither_amp=amp, clip=clip, dither_first_sample=f0 * job_nch + c0, out_kind=capi.OUT_INT32,
                          out=out.ptr))

    def host_step():
        if cjobs:
            capi.run_batch(cjobs, devices=(local,))
        for job, shard, pin, out, ctx, d_in, d_out, job_nch in singles:
            j, c0, k, f0, nf = shard
            ctx.set_option("dither_stride", job_nch)
            scale, amp, clip = scale_of(job)
            ctx.decimate_host_raw(pin.ptr, pin.array.shape[1], job.first_byte, job.in_count, job.idle,
                                  job.fir.nlut + f0 * job.fir.nstep, nf, scale, amp, clip, f0 * job_nch + c0, capi.OUT_INT32, out.ptr)

    for _ in range(warmup):
        device_step()
    stream.synchronize()
    dist.barrier()
    launches0 = runner.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        device_step()
    e1.record(stream)
    stream.synchronize()
    dist.barrier()
    dev_ms = dist.max_over_ranks(e0.elapsed_time(e1))
    launches = runner.launches - launches0
    cached_ms = None
    if dither:
        # the same job with the contexts' dither plane cache on: what a batch costs once the plane of its longest file exists
        for w in work:
            w[4].set_option("dither_cache", 1 << 27)
        for _ in range(max(1, warmup)):
            device_step()
        stream.synchronize()
        dist.barrier()
        e0.record(stream)
        for _ in range(steps):
            device_step()
        e1.record(stream)
        stream.synchronize()
        dist.barrier()
        cached_ms = dist.max_over_ranks(e0.elapsed_time(e1))
    for _ in range(warmup):
        host_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        host_step()
    torch.cuda.synchronize()
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0)
    dist.barrier()
    h2d = sum(w[2].array.nbytes for w in work)
    d2h = sum(w[3].array.nbytes for w in work)
    # roofline of the whole job: LUT bytes every output sample of every file needs, against `world` GPUs' shared memory
    lut_elem = 8 if lut_bits == 64 else 4
    smem_bytes = sum(n * f[2] * fir.nlut * lut_elem for n, f, fir in zip(frames, files, firs))
    smem_peak, n_sm, peaks, peak_src = smem_peak_gbs(dist)
    achieved = smem_bytes * steps / (dev_ms * 1e-3) / 1e9
    value = total_bits * steps / (dev_ms * 1e-3) / 1e6
    e2e_value = total_bits * steps / e2e_s / 1e6
    audio_s = sum(f[3] for f in files)
    res = {"workload": desc, "files": len(files), "dither": bool(dither), "lut": f"{lut_bits}-bit", "scaling": "strong",
           "value": value, "unit": "Mbit/s", "ms_per_step": dev_ms / steps, "per_gpu_mbit_per_s": value / world,
           "x_realtime": audio_s / (dev_ms * 1e-3 / steps), "shards_rank0": len(work),
           "l2": "per-step working set > L2" if h2d + d2h > L2_BYTES else "working set may fit L2",
           "roofline": {"bound": "smem", "achieved": achieved, "peak": smem_peak * world, "unit": "GB/s", "frac": achieved / (smem_peak * world),
                        "traffic": None, "peak_source": f"{world} GPUs x {n_sm} SMs x 128 B/clk x sm_max_mhz ({peak_src})",
                        "algorithmic_bytes_per_sample": sorted({fir.nlut * lut_elem for fir in firs}),
                        "note": "whole job: LUT bytes of every output sample / device time of the slowest rank"},
           "e2e": {"value": e2e_value, "unit": "Mbit/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_s / steps * 1e3, "x_realtime": audio_s / (e2e_s / steps),
                   "api": (f"dsdgpu_run_batch: one call per step for {len(cjobs)} of this rank's {len(work)} shards, pipelined across files "
                           f"(pinned host buffers); channel-subset shards through dsdgpu_decimate_host"), "note": "bytes and shard counts are rank 0's"},
           "gpu_launches": launches}
    if cached_ms:
        res["dither_cached"] = {"note": "contexts keep the dither plane (dither_cache on): generated in the warm-up steps, read in place afterwards",
                                "value": total_bits * steps / (cached_ms * 1e-3) / 1e6, "unit": "Mbit/s", "ms_per_step": cached_ms / steps,
                                "roofline_frac": smem_bytes * steps / (cached_ms * 1e-3) / 1e9 / (smem_peak * world)}
    for w in work:
        w[2].free()
        w[3].free()
    del work
    runner.close()
    return res


def run_sharded_arm(args, spec):
    dist = Dist()
    sampler = ClockSampler(dist.local) if dist.rank == 0 else None
    r = sharded_case(dist, args, spec, args.steps, args.warmup)
    clocks = sampler.stop(0.0, time.perf_counter()) if sampler else None
    if dist.rank == 0:
        cpu = None
        if dist.world == 1 and args.cpu_seconds > 0:
            cores = len(os.sched_getaffinity(0))
            v, take, njobs, kind, _ = reference_batch_rate(spec, cores, args.cpu_seconds, 1, 0)
            cpu = {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind,
                   "sample": f"the first {take} of {njobs} files / 60 s pieces, one reference process per core"}
        print(json.dumps({
            "metric": "dsd_mbit_per_s", "value": r["value"], "unit": "Mbit/s", "n_gpus": dist.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if spec[2] == 64 else "i32", "data": "synthetic",
            "per_gpu_mbit_per_s": r["per_gpu_mbit_per_s"], "x_realtime": r["x_realtime"],
            "config": {"workload": r["workload"], "files": r["files"], "dither": r["dither"], "lut": r["lut"],
                       "shards_rank0": r["shards_rank0"], "l2": r["l2"],
                       "parallelism": f"{dist.world} ranks, dsdgpu_plan_shards (file / channel / time segment + halo), no collective"},
            "roofline": r["roofline"], "cpu_baseline": cpu, "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": clocks,
        }), flush=True)
    dist.close()


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
    ap.add_argument("--dither-cache", type=int, default=0, help="1: the main case keeps the context's dither plane across steps (default: regenerated every step)")
    ap.add_argument("--out", default="int32", choices=["int32", "pcm24", "pcm16"], help="PCM layout handed back (int32 = FLAC__int32 as main.cpp:173-177; "
                    "pcm24 / pcm16 = packed little-endian samples, SURVEY 8f-4)")
    ap.add_argument("--subs", type=int, default=1, help="0: skip the configs[0] / dither-on / sharded sub-objects of the default line")
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
        wl = wl[:3] + (args.seconds, wl[4], wl[5].replace(" 60 s", f" {args.seconds} s"), wl[6] + f"_{args.seconds}s")
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
