"""BASELINE.json configurations at their full sizes, on synthetic sigma-delta streams.

config 1 (DSD64 stereo 60 s -> 24 bit / 88.2 kHz) is compared with the oracle in full.
config 2 (DSD128 stereo 60 s -> 176.4 kHz, fp64 tables) is too long for that to stay in the
seconds range, so it is checked through size-independent properties: two independent kernels
(segmented and direct) agree bit for bit over the whole stream, time-segment shards with their
halo stitch to the unsharded result, and windows sampled across the stream equal the oracle."""
import numpy as np
import pytest

from dsf2flac_b200 import batch, capi, filters
from tests.oracle_lib import DSD64

pytestmark = pytest.mark.gpu


def sdm_stream(nch, nbytes, fs, lsb_first=1):
    host = capi.host_lib()
    # one second of modulator output, tiled (the content only has to be real DSD statistics)
    sec = fs // 8
    one = np.empty((nch, sec), dtype=np.uint8)
    for c in range(nch):
        host.dsdhost_synth_sdm(one[c].ctypes.data, sec, float(fs), 1000.0 + 500.0 * c, 0.5, lsb_first)
    reps = -(-nbytes // sec)
    return np.ascontiguousarray(np.tile(one, (1, reps))[:, :nbytes])


def test_config1_full_vs_oracle(oracle):
    nbytes = DSD64 * 60 // 8
    data = sdm_stream(2, nbytes, DSD64)
    fir = filters.FILTERS[32]
    n = filters.app_frames(nbytes * 8, fir)
    assert n == 5291983
    for dither in (False, True):
        scale, amp, clip = filters.app_params(24, dither=dither)
        want = oracle.run_app(data, nbytes * 8, 1, DSD64, 88200, scale, amp, clip)
        with capi.DsdGpu(filters.build_lut(fir, 1), 2, fir.nstep) as g:
            got = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
        assert np.array_equal(got, want)


def test_config2_full_properties(oracle):
    fs = 2 * DSD64
    nbytes = fs * 60 // 8
    data = sdm_stream(2, nbytes, fs)
    fir = filters.pick(fs, 176400)
    assert fir.ratio == 32
    n = filters.app_frames(nbytes * 8, fir)
    assert n == 10583983
    scale, amp, clip = filters.app_params(24, dither=False)
    lut = filters.build_lut(fir, 1)
    outs = {}
    for name, k in (("ring", capi.KERNEL_RING), ("seg", capi.KERNEL_SEGMENTED), ("direct", capi.KERNEL_DIRECT)):
        with capi.DsdGpu(lut, 2, fir.nstep) as g:
            g.set_option("kernel", k)
            outs[name] = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
    assert np.array_equal(outs["seg"], outs["direct"])
    assert np.array_equal(outs["ring"], outs["direct"])
    # 8-way time/channel shards with halos == unsharded
    job = batch.Job(data, fs, 176400, msb_flag=True, bits=24, dither=False)
    runner = batch.BatchRunner()
    parts = []
    for r in range(8):
        parts += runner.run([job], r, 8)
    assert len(parts) == 8
    assert np.array_equal(batch.stitch([job], parts)[0], outs["seg"])
    runner.close()
    # sampled windows against the oracle (start, middle, the very end)
    for f0 in (0, n // 3, n - 50000):
        b0 = max(0, fir.nlut + f0 * fir.nstep - 200)
        b1 = min(nbytes, fir.nlut + (f0 + 50000) * fir.nstep + 8)
        sub = np.ascontiguousarray(data[:, b0:b1])
        pre = fir.nlut + f0 * fir.nstep - b0 + 1
        cnt = min(50000, n - f0)
        want = oracle.run_raw(sub, 1 << 40, 1, fs, 176400, pre, cnt, scale, amp, clip)
        assert np.array_equal(outs["seg"][f0:f0 + cnt], want), f0
    # the signal is what went in: 1 kHz / 1.5 kHz sines at 0.5 * 4 dB
    x = outs["seg"][:176400].astype(np.float64) / 2 ** 23
    assert abs(np.abs(x).max() - 0.5 * 10 ** 0.2) < 0.01


def test_config4_size_device_resident_periodicity(oracle):
    """BASELINE config 4's size: DSD512 stereo, 600 s = 1 693 440 000 bytes per channel, divided
    by 32 (423 359 983 frames x 2 channels, 3.4 GB in and 3.4 GB out, device resident).  The
    stream is one second of modulator output tiled 600 times, so -- the filter being time
    invariant -- frame f + k * (frames per second) must equal frame f bit for bit wherever the
    72-byte window lies inside the data; the first second is checked against the oracle.  This
    walks every index, offset and tile counter past 2^31 / 2^32."""
    import torch
    fs = 8 * DSD64
    sec = fs // 8                                            # 2 822 400 bytes per channel per second
    seconds = 600
    nbytes = sec * seconds
    fir = filters.pick(fs, fs // 32)
    n = filters.app_frames(nbytes * 8, fir)
    assert n == 423359983
    one = sdm_stream(2, sec, fs)
    scale, amp, clip = filters.app_params(24, dither=False)
    front = 3                                                # frame grid 4-byte aligned to the buffer
    stride = (nbytes + front + 255) // 256 * 256
    d_in = torch.full((2, stride), 0x69, dtype=torch.uint8, device="cuda")
    d_in[:, front:front + nbytes] = torch.from_numpy(one).cuda().repeat(1, seconds)
    d_out = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    with capi.DsdGpu(filters.build_lut(fir, 1), 2, fir.nstep) as g:
        g.decimate_device(d_in, stride, -front, nbytes + front, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, d_out)
        torch.cuda.synchronize()
        fps = sec // fir.nstep                               # frames per tiled second
        first = d_out[:fps].clone()
        # the first second against the oracle (its own first frames see the start of the stream)
        want = oracle.run_raw(np.ascontiguousarray(np.tile(one, (1, 2))[:, :sec + 400]), 1 << 40, 1, fs, fs // 32,
                              fir.nlut + 1, fps, scale, amp, clip)
        assert np.array_equal(first.cpu().numpy(), want)
        # every later second repeats second 1 (second 0 starts with frames whose window is not periodic yet)
        base = d_out[fps:2 * fps]
        for k in (2, 77, 300, 598):
            assert torch.equal(d_out[k * fps:(k + 1) * fps], base), k
        tail = d_out[599 * fps:n]
        assert torch.equal(tail[:len(tail) - 32], base[:len(tail) - 32])
        # and the direct kernel agrees on the last 100 000 frames (stream end, past 2^32 bytes of output)
        g.set_option("kernel", capi.KERNEL_DIRECT)
        cnt = 100_000
        d_chk = torch.empty((cnt, 2), dtype=torch.int32, device="cuda")
        g.decimate_device(d_in, stride, -front, nbytes + front, 0x69, fir.nlut + (n - cnt) * fir.nstep, cnt, scale, amp, clip,
                          0, capi.OUT_INT32, d_chk)
        assert torch.equal(d_chk, d_out[n - cnt:])


def test_config4_as_stated_divide_by_64_extension(oracle):
    """BASELINE config 4 as it is written -- DSD512 stereo, 600 s -> 352.8 kHz -- is a ratio (64) the reference
    rejects; with the extension filter it runs as two passes of the ring kernel over a 3.4 GB plane of running
    sums.  Same size-independent checks: the first second against the reference's arithmetic on the extension
    filter (the C restatement, pinned to oracle/_ref by tests/test_extension.py), periodicity of a tiled
    stream across all ten minutes, and the segmented and direct kernels on the last 100 000 frames."""
    import torch
    fs = 8 * DSD64
    sec = fs // 8
    seconds = 600
    nbytes = sec * seconds
    fir = filters.pick(fs, 352800, extensions=True)
    assert (fir.ratio, fir.nlut, fir.nstep) == (64, 144, 8)
    n = filters.app_frames(nbytes * 8, fir)
    assert n == nbytes // 8 - 17 == 211679983
    one = sdm_stream(2, sec, fs)
    scale, amp, clip = filters.app_params(24, dither=False)
    front = 3                                                # (p0 - in_first + 1) % 4 == 0 with p0 = 144
    stride = (nbytes + front + 255) // 256 * 256
    d_in = torch.full((2, stride), 0x69, dtype=torch.uint8, device="cuda")
    d_in[:, front:front + nbytes] = torch.from_numpy(one).cuda().repeat(1, seconds)
    d_out = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    oracle.set_custom_filter(fir)
    try:
        with capi.DsdGpu(filters.build_lut(fir, 1), 2, fir.nstep) as g:
            g.decimate_device(d_in, stride, -front, nbytes + front, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, d_out)
            torch.cuda.synchronize()
            fps = sec // fir.nstep
            want = oracle.run_raw(np.ascontiguousarray(np.tile(one, (1, 2))[:, :sec + 400]), 1 << 40, 1, fs, 352800,
                                  fir.nlut + 1, fps, scale, amp, clip)
            assert np.array_equal(d_out[:fps].cpu().numpy(), want)
            base = d_out[fps:2 * fps]
            for k in (2, 77, 300, 598):
                assert torch.equal(d_out[k * fps:(k + 1) * fps], base), k
            tail = d_out[599 * fps:n]
            assert torch.equal(tail[:len(tail) - 32], base[:len(tail) - 32])
            cnt = 100_000
            for kernel in (capi.KERNEL_SEGMENTED, capi.KERNEL_DIRECT):
                g.set_option("kernel", kernel)
                d_chk = torch.empty((cnt, 2), dtype=torch.int32, device="cuda")
                g.decimate_device(d_in, stride, -front, nbytes + front, 0x69, fir.nlut + (n - cnt) * fir.nstep, cnt, scale, amp,
                                  clip, 0, capi.OUT_INT32, d_chk)
                assert torch.equal(d_chk, d_out[n - cnt:]), kernel
    finally:
        oracle.set_custom_filter(None)


@pytest.mark.parametrize("out_fs", [176400, 88200])
def test_divide_by_128_and_256_extension_full_size(oracle, out_fs):
    """DSD512 stereo, 60 s -> 176.4 / 88.2 kHz (ratios 128 and 256: 4 and 8 passes of the segmented kernel over
    the plane of running sums): the segmented and the direct kernel (96-table passes, a different cut of the
    same sequential sum) agree over the whole stream, and windows at the start, in the middle and at the end
    equal the reference's arithmetic on the extension filter."""
    fs = 8 * DSD64
    nbytes = fs * 60 // 8
    data = sdm_stream(2, nbytes, fs)
    fir = filters.pick(fs, out_fs, extensions=True)
    n = filters.app_frames(nbytes * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=False)
    lut = filters.build_lut(fir, 1)
    outs = {}
    for name, k in (("seg", capi.KERNEL_SEGMENTED), ("direct", capi.KERNEL_DIRECT)):
        with capi.DsdGpu(lut, 2, fir.nstep) as g:
            g.set_option("kernel", k)
            outs[name] = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
    assert np.array_equal(outs["seg"], outs["direct"])
    oracle.set_custom_filter(fir)
    try:
        for f0 in (0, n // 2, n - 20000):
            b0 = max(0, fir.nlut + f0 * fir.nstep - 2 * fir.nlut)
            b1 = min(nbytes, fir.nlut + (f0 + 20000) * fir.nstep + 8)
            sub = np.ascontiguousarray(data[:, b0:b1])
            pre = fir.nlut + f0 * fir.nstep - b0 + 1
            cnt = min(20000, n - f0)
            want = oracle.run_raw(sub, 1 << 40, 1, fs, out_fs, pre, cnt, scale, amp, clip)
            assert np.array_equal(outs["seg"][f0:f0 + cnt], want), f0
    finally:
        oracle.set_custom_filter(None)
