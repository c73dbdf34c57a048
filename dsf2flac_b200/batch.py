"""One-process-per-GPU driver for a batch of DSD streams.

The reference converts one file at a time on one thread (scripts/batch_dsf2flac.sh:16-19 is its
only batch notion).  Here every rank computes the same shard plan (dsdgpu_plan_shards: by file,
by channel, by time segment with an nlut-1 byte halo), runs its own shards through the C ABI
with host buffers, and keeps the PCM on its host.  Shards are independent
(reference dsd_decimator.cpp:193-198: a sample depends on nlut bytes of its own channel), so
there is no data-path collective; torch.distributed is used by the callers for barriers and
timing only.

A Job describes what main.cpp would run for one file with --onefile: frames i = 0..n-1 at byte
position p = nlut + i*nstep (after the creep loop, main.cpp:169-171), scale/dither/clip from
main.cpp:239-245, dither draws 2m, 2m+1 for output sample m = i*nch + c.
"""
from dataclasses import dataclass, field

import numpy as np

from . import capi, filters


@dataclass
class Job:
    planar: np.ndarray            # uint8 [nch, nbytes], host memory (pinned for real runs)
    dsd_fs: int
    out_fs: int
    msb_flag: bool = True         # reader->msbIsPlayedFirst(): True for .dsf, False for .dff
    length_samples: int = -1      # DSD samples per channel (container header); -1 = 8*nbytes
    bits: int = 24
    dither: bool = False
    scale_db: float = 4.0
    lut_precision: int = 64
    idle: int = 0x69
    first_byte: int = 0           # stream index of planar[:, 0] when only a window of the stream is held
    extensions: bool = False      # accept the ratios the reference rejects (64, 128, 256: extension filters)
    fir: object = field(init=False)

    def __post_init__(self):
        if self.length_samples < 0:
            self.length_samples = 8 * (self.first_byte + self.planar.shape[1])
        self.fir = filters.pick(self.dsd_fs, self.out_fs, extensions=self.extensions)
        if self.fir is None:
            raise ValueError("Sorry, incompatible sample rate combination")

    @property
    def nch(self):
        return self.planar.shape[0]

    @property
    def frames(self):
        return filters.app_frames(self.length_samples, self.fir)

    @property
    def in_count(self):
        # the reader still delivers the byte at index ceil(L/8): availability is tested before
        # its marker moves (dsf_file_reader.cpp:87,101); clipped to what the buffer holds
        return max(0, min(self.first_byte + self.planar.shape[1], -(-self.length_samples // 8) + 1) - self.first_byte)


def plan(jobs, rank, world, align_frames=1024):
    return capi.plan_shards([j.frames for j in jobs], [j.nch for j in jobs], [j.fir.nlut for j in jobs], rank, world,
                            align_frames)


class BatchRunner:
    """Runs this rank's shards of a batch.  Contexts (tables on the device) are cached per
    (ratio, bit order, channel count, precision); outputs are int32 [frame_count, ch_count]."""

    def __init__(self, device=0, make_ctx=capi.DsdGpu):
        self.device = device
        self.make_ctx = make_ctx
        self.ctxs = {}

    def ctx_for(self, job, ch_count):
        key = (job.fir.ratio, bool(job.msb_flag), ch_count, job.lut_precision)
        if key not in self.ctxs:
            lut = filters.build_lut(job.fir, job.msb_flag)
            self.ctxs[key] = self.make_ctx(lut, ch_count, job.fir.nstep, device=self.device,
                                           lut_precision=job.lut_precision)
        return self.ctxs[key]

    def run_shard(self, job, shard, out=None):
        _, c0, k, f0, nf = shard
        ctx = self.ctx_for(job, k)
        ctx.set_option("dither_stride", job.nch)
        scale, amp, clip = filters.app_params(job.bits, job.scale_db, job.dither)
        if out is None:
            out = np.empty((nf, k), dtype=np.int32)
        nbytes = job.planar.shape[1]
        ctx.decimate_host_raw(job.planar.ctypes.data + c0 * nbytes, nbytes, job.first_byte, job.in_count, job.idle,
                              job.fir.nlut + f0 * job.fir.nstep, nf, scale, amp, clip, f0 * job.nch + c0,
                              capi.OUT_INT32, out.ctypes.data)
        return out

    def run(self, jobs, rank, world, align_frames=1024):
        """-> list of (shard, int32 array) for this rank, one synchronous dsdgpu_decimate_host call per shard (each keeps its
        own PCM array: what the tests stitch).  run_batch_c() below is the pipelined way: one dsdgpu_run_batch call, chunks
        of neighbouring shards and files in flight together, PCM written into the jobs' interleaved outputs."""
        return [(s, self.run_shard(jobs[s[0]], s)) for s in plan(jobs, rank, world, align_frames)]

    @property
    def launches(self):
        return sum(c.launches for c in self.ctxs.values())

    def close(self):
        for c in self.ctxs.values():
            c.close()
        self.ctxs = {}


def run_batch_c(jobs, devices=(0,), rank=0, world=1, align_frames=1024, lib=None, out_kind=capi.OUT_INT32):
    """The same batch through the C entry point dsdgpu_run_batch (include/dsdgpu.h): one call, one host thread per device,
    chunks pipelined across files.  -> list of [frames, nch] arrays, one per job (only this rank's shards are filled)."""
    luts, outs, cjobs = {}, [], []
    for j in jobs:
        key = (j.fir.ratio, bool(j.msb_flag))
        if key not in luts:
            luts[key] = np.ascontiguousarray(filters.build_lut(j.fir, j.msb_flag))
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        if out_kind == capi.OUT_PCM24:
            out = np.zeros((j.frames, j.nch, 3), dtype=np.uint8)
        else:
            out = np.zeros((j.frames, j.nch), dtype={capi.OUT_INT32: np.int32, capi.OUT_FLOAT64: np.float64, capi.OUT_PCM16: np.int16}[out_kind])
        outs.append(out)
        cjobs.append(dict(in_=j.planar, in_stride=j.planar.shape[1], in_first=j.first_byte, in_count=j.in_count, source_block_bytes=0,
                          fill=j.idle, nch=j.nch, nlut=j.fir.nlut, nstep=j.fir.nstep, lut_precision=j.lut_precision, lut=luts[key],
                          p0=j.fir.nlut, nframes=j.frames, scale=scale, dither_amp=amp, clip=clip, dither_first_sample=0,
                          out_kind=out_kind, out=out))
    capi.run_batch(cjobs, devices, rank, world, align_frames, lib=lib)
    return outs


def stitch(jobs, shard_results):
    """Host-side reassembly of shard outputs (from all ranks) into one [frames, nch] int32 array
    per job; raises if the shards do not tile the batch exactly once."""
    outs = [np.zeros((j.frames, j.nch), dtype=np.int32) for j in jobs]
    seen = [np.zeros((j.frames, j.nch), dtype=np.uint8) for j in jobs]
    for (job, c0, k, f0, nf), data in shard_results:
        outs[job][f0:f0 + nf, c0:c0 + k] = data
        seen[job][f0:f0 + nf, c0:c0 + k] += 1
    for s in seen:
        if s.size and (s.min() != 1 or s.max() != 1):
            raise ValueError("shards do not tile the batch exactly once")
    return outs
