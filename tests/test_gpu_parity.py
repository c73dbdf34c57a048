"""GPU parity, through the C ABI (include/dsdgpu.h), against the CPU oracle on the same seeded
inputs and against the committed golden vectors.  Integer PCM and fp64 sums are bit-exact for
the fp64-table path; the fp32-table variant is held to the tolerance BASELINE.json states:
within +-1 LSB at 24 bit and <= 1e-6 relative error before quantisation."""
import os

import numpy as np
import pytest

from dsf2flac_b200 import batch, capi, filters
from tests.oracle_lib import DSD64, seeded_bytes

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
RATES = {32: 88200, 16: 176400, 8: 352800}
KERNELS = {"direct": capi.KERNEL_DIRECT, "seg": capi.KERNEL_SEGMENTED, "ring": capi.KERNEL_RING, "auto": capi.KERNEL_AUTO}


def make(ratio, msb, nch, kernel="auto", precision=64):
    fir = filters.FILTERS[ratio]
    g = capi.DsdGpu(filters.build_lut(fir, msb), nch, fir.nstep, lut_precision=precision)
    g.set_option("kernel", KERNELS[kernel])
    return fir, g


@pytest.mark.parametrize("kernel", ["direct", "seg", "ring"])
@pytest.mark.parametrize("msb", [0, 1])
@pytest.mark.parametrize("ratio", [32, 16, 8])
def test_app_parity_all_depths(oracle, ratio, msb, kernel):
    fir, g = make(ratio, msb, 2, kernel)
    data = seeded_bytes(2, 40000, 11 * ratio + msb)
    L = 40000 * 8 - 24
    n = filters.app_frames(L, fir)
    for bits, dither in ((24, False), (16, True), (20, True), (24, True)):
        scale, amp, clip = filters.app_params(bits, dither=dither)
        want = oracle.run_app(data, L, msb, DSD64, RATES[ratio], scale, amp, clip)
        assert len(want) == n
        got = g.decimate_host(data[:, :-(-L // 8) + 1], fir.nlut, n, scale, amp, clip)
        assert np.array_equal(got, want), (bits, dither)
    g.close()


@pytest.mark.parametrize("kernel", ["direct", "seg", "ring"])
@pytest.mark.parametrize("ratio", [32, 16, 8])
def test_fp64_sums_bit_exact_incl_idle_prefill(oracle, ratio, kernel):
    fir, g = make(ratio, 1, 3, kernel)
    data = seeded_bytes(3, 9000, ratio)
    # straight after rewind: p0 = -1, the first frames read only the 0x69 pre-fill; the last ones
    # run past the end of the data
    nframes = (9000 + 300) // fir.nstep
    want = oracle.run_raw(data, 9000 * 8, 1, DSD64, RATES[ratio], 0, nframes, 1.0, want="f64")
    got = g.decimate_host(data, -1, nframes, 1.0, out_kind=capi.OUT_FLOAT64)
    assert got.tobytes() == want.tobytes()
    # a different idle byte and a hot scale that clips
    want = oracle.run_raw(data, 9000 * 8, 1, DSD64, RATES[ratio], 5, nframes, 9e7, 0.0, 8388607.0, idle=0x96)
    got = g.decimate_host(data, 4, nframes, 9e7, 0.0, 8388607.0, fill=0x96)
    assert np.array_equal(got, want) and np.abs(got).max() == 8388607
    g.close()


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("ratio", [32, 8])
def test_clip_amplitudes_of_every_kind(oracle, ratio, precision):
    """The ring kernels clip AFTER rounding, on the integer (round() is monotone: round(clamp(x, -c, c)) =
    clamp(round(x), round(-c), round(c))); the reference clips the double and rounds then (dsd_decimator.cpp:207-214).
    Same integers for whole, fractional, half-way, tiny and beyond-int32 clip amplitudes, with and without dither,
    int32 and packed output; the fp64 output keeps the clipped double itself."""
    if precision == 32 and ratio != 32:
        pytest.skip("32-bit tables exist for the divide-by-32 filter")
    fir, g = make(ratio, 1, 2, "ring", precision)
    data = seeded_bytes(2, 6000, 77 + ratio)
    n = 5000 // fir.nstep
    tol = precision == 32
    for scale, clip, amp in ((9e7, 8388607.0, 0.0), (9e7, 8388607.0, 1.0), (3e3, 1000.5, 0.0), (3e3, 1000.49, 1.0),
                             (3e3, 1001.7, 0.0), (40.0, 0.4, 0.0), (40.0, 0.5, 0.0), (40.0, 2.5, 1.0),
                             # clip amplitudes at and beyond int32 (samples stay inside: the reference's cast of a larger
                             # double is undefined behaviour in C++, the kernels saturate)
                             (1e9, 3e9, 0.0), (1e9, 2147483647.2, 0.0), (1.5e9, 2147483646.5, 0.0), (1e9, 0.0, 0.0),
                             (1e9, 2147483648.0, 2.0), (1.5e9, 1.9e9, 0.0)):
        want = oracle.run_raw(data, 6000 * 8, 1, DSD64, RATES[ratio], fir.nlut + 1, n, scale, amp, clip)
        got = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
        if tol:                 # 32-bit tables: a sum may differ in its last bits, a rounded sample by one
            assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= max(1, scale * 4e-8), (scale, clip, amp)
            if clip > 0:
                assert np.abs(got.astype(np.int64)).max() <= np.abs(want.astype(np.int64)).max()
        else:
            assert np.array_equal(got, want), (scale, clip, amp)
        if clip > 0:
            wf = oracle.run_raw(data, 6000 * 8, 1, DSD64, RATES[ratio], fir.nlut + 1, n, scale, amp, clip, want="f64")
            gf = g.decimate_host(data, fir.nlut, n, scale, amp, clip, out_kind=capi.OUT_FLOAT64)
            if not tol:
                assert gf.tobytes() == wf.tobytes(), (scale, clip, amp)
    want = oracle.run_raw(data, 6000 * 8, 1, DSD64, RATES[ratio], fir.nlut + 1, n, 9e7, 1.0, 8388607.0)
    got = g.decimate_host(data, fir.nlut, n, 9e7, 1.0, 8388607.0, out_kind=capi.OUT_PCM24)
    q = got[..., 0].astype(np.int32) | (got[..., 1].astype(np.int32) << 8) | (got[..., 2].astype(np.int8).astype(np.int32) << 16)
    if not tol:
        assert np.array_equal(q, want)
    g.close()


@pytest.mark.parametrize("ratio,msb", [(r, m) for r in RATES for m in (0, 1)])
def test_golden_vectors(ratio, msb):
    tag = f"r{ratio}_m{msb}"
    fir, g = make(ratio, msb, 2)
    data, L = GOLD[f"in_{tag}"], 2048 * 8
    n = filters.app_frames(L, fir)
    assert np.array_equal(g.decimate_host(data, fir.nlut, n, 13295047.713424837, 0.0, 8388607.0), GOLD[f"app24_{tag}"])
    assert np.array_equal(g.decimate_host(data, fir.nlut, n, 51933.78013056577, 1.0, 32767.0), GOLD[f"app16d_{tag}"])
    got = g.decimate_host(data, -1, 96, 1.0, out_kind=capi.OUT_FLOAT64)
    assert got.tobytes() == GOLD[f"raw64_{tag}"].tobytes()
    # in_count models the reader: the byte at ceil(L/8) is still delivered, then the idle byte
    got = g.decimate_host(data[:, :1001], 899, 200, 830940.4820890523, 0.0, 524287.0, fill=0x96)
    assert np.array_equal(got, GOLD[f"tail20_{tag}"])
    g.close()


def test_golden_sigma_delta_both_packings():
    for lsb_first in (1, 0):
        fir, g = make(32, lsb_first, 2)
        got = g.decimate_host(GOLD[f"sine_in_lsb{lsb_first}"], fir.nlut, filters.app_frames(1536 * 8, fir),
                              13295047.713424837, 0.0, 8388607.0)
        assert np.array_equal(got, GOLD[f"sine_app24_lsb{lsb_first}"])
        g.close()


@pytest.mark.parametrize("kernel", ["direct", "seg", "ring"])
def test_ragged_windows_and_tiny_calls(oracle, kernel):
    fir, g = make(32, 0, 1, kernel)
    data = seeded_bytes(1, 30000, 3)
    full = oracle.run_raw(data, 30000 * 8, 0, DSD64, 88200, 0, 7600, 5.0, want="f64")     # frames at p = -1 + 4i
    # the caller holds only a window of the stream (in_first > 0, unaligned), everything else is fill
    for first, count in ((0, 30000), (16, 20000), (3, 12345), (4097, 1), (29990, 10)):
        got = np.empty((7600, 1))
        g.decimate_host_raw(data.ctypes.data + first, 0, first, count, 0x69, -1, 7600, 5.0, 0.0, 0.0, 0, capi.OUT_FLOAT64,
                            got.ctypes.data)
        masked = np.full_like(data, 0x69)
        masked[:, first:first + count] = data[:, first:first + count]
        want = oracle.run_raw(masked, 30000 * 8, 0, DSD64, 88200, 0, 7600, 5.0, want="f64")
        assert got.tobytes() == want.tobytes(), (first, count)
    # any start, any count, including one frame and none
    for p0, n in ((72, 1), (73, 2), (1001, 513), (29000, 400), (50, 0)):
        got = g.decimate_host(data, p0, n, 5.0, out_kind=capi.OUT_FLOAT64)
        want = oracle.run_raw(data, 30000 * 8, 0, DSD64, 88200, p0 + 1, n, 5.0, want="f64")
        assert got.tobytes() == want.tobytes(), (p0, n)
    assert full.shape == (7600, 1)
    g.close()


@pytest.mark.parametrize("nch", [1, 2, 6])
def test_channel_counts_and_chunked_pipeline(oracle, nch):
    fir, g = make(32, 1, nch)
    g.set_option("chunk_frames", 4099)       # many double-buffered chunks, odd size
    data = seeded_bytes(nch, 70000, nch)
    scale, amp, clip = filters.app_params(24, dither=True)
    n = filters.app_frames(70000 * 8, fir)
    want = oracle.run_app(data, 70000 * 8, 1, DSD64, 88200, scale, amp, clip)
    got = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
    assert np.array_equal(got, want)
    g.close()


def test_dither_stream_by_position(oracle):
    import torch
    fir, g = make(32, 1, 2)
    for first, count in ((0, 5000), (12345, 4000), (100_000_000, 3000)):
        d = torch.empty(count, dtype=torch.float64, device="cuda")
        g.dither_fill_device(first, count, d)
        r = oracle.rand(2 * count, skip=2 * first).astype(np.float64) / 2147483647.0
        want = r[0::2] - r[1::2]
        assert d.cpu().numpy().tobytes() == want.tobytes(), first
    # a call that starts mid-stream draws what the reference would have drawn there
    data = seeded_bytes(2, 20000, 8)
    scale, amp, clip = filters.app_params(16, dither=True)
    want = oracle.run_raw(data, 20000 * 8, 1, DSD64, 88200, 73, 3000, scale, amp, clip, rand_skip=2 * 777)
    got = g.decimate_host(data, 72, 3000, scale, amp, clip, dither_first_sample=777)
    assert np.array_equal(got, want)
    g.close()


def test_dither_plane_cache(oracle):
    """The context keeps the dither plane it has generated (dsdgpu_set_option "dither_cache"): calls that start inside it
    read it in place, calls that reach past it generate the rest, a plane that is outgrown is copied, a call far into
    the stream makes its own -- and every one of them draws what the reference draws at that position."""
    fir, g = make(32, 1, 2)
    data = seeded_bytes(2, 60000, 31)
    scale, amp, clip = filters.app_params(24, dither=True)
    nbits = 60000 * 8

    def want(p0, n, first):
        return oracle.run_raw(data, nbits, 1, DSD64, 88200, p0 + 1, n, scale, amp, clip, rand_skip=2 * first)   # p0 = pre_steps - 1

    def got(p0, n, first, kind=capi.OUT_INT32):
        return g.decimate_host(data, p0, n, scale, amp, clip, dither_first_sample=first, out_kind=kind)

    launches = []
    for p0, n, first in ((72, 3000, 0),              # fills samples [0, 6000)
                         (72, 3000, 0),              # read in place: one launch less
                         (72 + 4 * 1000, 2000, 2000),   # starts inside, ends inside
                         (72 + 4 * 3000, 4000, 6000),   # starts at the end of the valid part: generates [6000, 14000)
                         (72 + 4 * 500, 9000, 1001),    # odd first sample (the 16-byte dither loads fall back to 8-byte ones)
                         (72, 14000, 0),                # reaches past the valid part again
                         (72, 14000, 0)):               # ... and now lies inside it
        before = g.launches
        assert np.array_equal(got(p0, n, first), want(p0, n, first)), (p0, n, first)
        launches.append(g.launches - before)
    assert launches[1] == launches[0] - 1 and launches[2] == launches[1] and launches[3] == launches[0]
    assert launches[4] == launches[0] and launches[5] == launches[0] and launches[6] == launches[1]
    # a call far into the stream (a time shard) does not extend the plane
    before = g.launches
    assert np.array_equal(got(72, 3000, 5_000_000), want(72, 3000, 5_000_000))
    assert g.launches - before == launches[0]
    # a small limit: calls beyond it make their own plane; shrinking below the valid part drops the plane
    g.set_option("dither_cache", 4096)
    assert np.array_equal(got(72, 3000, 0), want(72, 3000, 0))
    assert np.array_equal(got(72, 1000, 100), want(72, 1000, 100))
    g.set_option("dither_cache", 0)
    assert np.array_equal(got(72, 3000, 0), want(72, 3000, 0))
    g.close()
    # growth past the first allocation (2^22 samples): the valid part is carried over
    fir, g = make(32, 1, 2)
    n1 = 3000
    assert np.array_equal(got(72, n1, 0), want(72, n1, 0))
    d = seeded_bytes(2, 4 * 2_200_000 + 400, 32)
    big = g.decimate_host(d, 72, 2_200_000, scale, amp, clip, dither_first_sample=0)       # 4.4 M samples > 2^22
    head = oracle.run_raw(d, d.shape[1] * 8, 1, DSD64, 88200, 73, 5000, scale, amp, clip, rand_skip=0)
    assert np.array_equal(big[:5000], head)
    tail0 = 2_195_000
    tail = oracle.run_raw(d, d.shape[1] * 8, 1, DSD64, 88200, 73 + 4 * tail0, 5000, scale, amp, clip, rand_skip=2 * 2 * tail0)
    assert np.array_equal(big[tail0:], tail)
    assert np.array_equal(got(72, n1, 0), want(72, n1, 0))
    g.close()


def test_device_pointers_and_stream(oracle):
    import torch
    fir, g = make(32, 1, 2)
    data = seeded_bytes(2, 50000, 21)
    n = filters.app_frames(50000 * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=False)
    want = oracle.run_app(data, 50000 * 8, 1, DSD64, 88200, scale, amp, clip)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        g.decimate_device(d_in, 50000, 0, 50000, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, d_out,
                          stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)
    before = g.launches
    g.decimate_device(d_in, 50000, 0, 50000, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, d_out)
    assert g.launches == before + 1
    g.close()


def test_submit_wait_slots(oracle):
    fir, g = make(16, 0, 2)
    data = seeded_bytes(2, 30000, 4)
    want = oracle.run_raw(data, 30000 * 8, 0, DSD64, 176400, 31, 14000, 1000.0)
    pin_in = capi.PinnedBuffer((2, 30000), np.uint8)
    pin_in.array[:] = data
    pin_out = capi.PinnedBuffer((14000, 2), np.int32)
    half = 7000
    for s in (0, 1):
        g.submit(s, pin_in.ptr, 30000, 0, 30000, 0x69, 30 + s * half * 2, half, 1000.0, 0.0, 0.0, 0, capi.OUT_INT32,
                 pin_out.ptr + s * half * 2 * 4)
    with pytest.raises(capi.DsdGpuError, match="in flight"):
        g.submit(0, pin_in.ptr, 30000, 0, 30000, 0x69, 30, half, 1000.0, 0.0, 0.0, 0, capi.OUT_INT32, pin_out.ptr)
    g.wait(0)
    g.wait(1)
    with pytest.raises(capi.DsdGpuError, match="nothing in flight"):
        g.wait(1)
    assert np.array_equal(pin_out.array, want)
    pin_in.free(); pin_out.free()
    g.close()


@pytest.mark.parametrize("ratio", [32, 16, 8])
def test_fp32_tables_within_tolerance(oracle, ratio):
    fir, g = make(ratio, 1, 2, precision=32)
    host = capi.host_lib()
    data = np.empty((2, 200000), dtype=np.uint8)
    for c, f in enumerate((1000.0, 1500.0)):
        host.dsdhost_synth_sdm(data[c].ctypes.data, 200000, float(DSD64), f, 0.5, 1)
    n = filters.app_frames(200000 * 8, fir)
    scale, _, clip = filters.app_params(24)
    want_i = oracle.run_app(data, 200000 * 8, 1, DSD64, RATES[ratio], scale, 0.0, clip)
    got_i = g.decimate_host(data, fir.nlut, n, scale, 0.0, clip)
    assert np.abs(got_i.astype(np.int64) - want_i).max() <= 1                       # +-1 LSB at 24 bit
    want_f = oracle.run_raw(data, 200000 * 8, 1, DSD64, RATES[ratio], fir.nlut + 1, n, 1.0, want="f64")
    got_f = g.decimate_host(data, fir.nlut, n, 1.0, out_kind=capi.OUT_FLOAT64)
    assert np.abs(got_f - want_f).max() <= 1e-6 * np.abs(want_f).max()              # pre-quantisation, relative
    g.close()


def test_option_errors():
    fir, g = make(16, 1, 2)
    with pytest.raises(capi.DsdGpuError):
        g.set_option("no_such_knob", 1)
    odd = capi.DsdGpu(np.ones((5, 256)), 1, 3)              # not one of the reference's filter geometries
    odd.set_option("kernel", capi.KERNEL_RING)
    with pytest.raises(capi.DsdGpuError, match="ring kernel needs"):
        odd.decimate_host(seeded_bytes(1, 1000, 1), 30, 10, 1.0)
    odd.set_option("kernel", capi.KERNEL_AUTO)                # ... the direct kernel takes any geometry
    assert odd.decimate_host(seeded_bytes(1, 1000, 1), 30, 10, 1.0).shape == (10, 1)
    odd.close()
    with pytest.raises(capi.DsdGpuError, match="bad kernel"):
        g.set_option("kernel", 2)                             # the retired skew kernel's number
    g.close()


@pytest.mark.parametrize("world", [1, 4])
def test_batch_runner_shards_on_gpu(oracle, world):
    from tests.test_sharding import mixed_jobs
    jobs = mixed_jobs(3)
    runner = batch.BatchRunner()
    results = []
    for r in range(world):
        results += runner.run(jobs, r, world, align_frames=64)
    assert runner.launches >= len(results)
    for j, got in zip(jobs, batch.stitch(jobs, results)):
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        want = oracle.run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip)
        assert np.array_equal(got, want)
    runner.close()


def to_blocks(planar, B):
    """DSF data-chunk layout of a planar [nch, n] stream (n padded with zeros to whole blocks)."""
    nch, n = planar.shape
    nb = -(-n // B)
    padded = np.zeros((nch, nb * B), dtype=np.uint8)
    padded[:, :n] = planar
    return np.ascontiguousarray(padded.reshape(nch, nb, B).transpose(1, 0, 2)).reshape(-1)


@pytest.mark.parametrize("kernel", ["ring", "direct"])
@pytest.mark.parametrize("nch", [1, 2, 3, 6])
def test_container_layouts_consumed_in_place(oracle, nch, kernel):
    """DSF block-planar and DSDIFF byte-interleaved sources give the same PCM as
    planar rows of the same stream: chunked pipeline, ragged last block, frames that start in
    the idle pre-fill and run past the end of the data."""
    fir, g = make(32, 1, nch, kernel)
    g.set_option("chunk_frames", 5000)
    n = 4096 * 7 + 1234
    data = seeded_bytes(nch, n, 40 + nch)
    scale, amp, clip = filters.app_params(24, dither=True)
    nframes = (n + 400) // fir.nstep
    want = oracle.run_raw(data, n * 8, 1, DSD64, 88200, 0, nframes, scale, amp, clip)
    for block, src in ((4096, to_blocks(data, 4096)), (64, to_blocks(data, 64)), (1, np.ascontiguousarray(data.T).reshape(-1))):
        g.set_option("source_block_bytes", block)
        got = np.empty((nframes, nch), dtype=np.int32)
        g.decimate_host_raw(src.ctypes.data, 0, 0, n, 0x69, -1, nframes, scale, amp, clip, 0, capi.OUT_INT32, got.ctypes.data)
        assert np.array_equal(got, want), block
    # a window of the stream that starts on a later block
    g.set_option("source_block_bytes", 4096)
    src = to_blocks(data, 4096)
    first = 2 * 4096
    masked = np.full_like(data, 0x69)
    masked[:, first:] = data[:, first:]
    want = oracle.run_raw(masked, n * 8, 1, DSD64, 88200, 0, nframes, scale, 0.0, clip)
    got = np.empty((nframes, nch), dtype=np.int32)
    g.decimate_host_raw(src.ctypes.data + first * nch, 0, first, n - first, 0x69, -1, nframes, scale, 0.0, clip, 0,
                        capi.OUT_INT32, got.ctypes.data)
    assert np.array_equal(got, want)
    g.set_option("source_block_bytes", 0)
    g.close()


def test_container_layouts_device_pointers(oracle):
    import torch
    fir, g = make(32, 0, 2)
    n = 50000
    data = seeded_bytes(2, n, 77)
    nframes = filters.app_frames(n * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=False)
    want = oracle.run_app(data, n * 8, 0, DSD64, 88200, scale, amp, clip)
    for block, src in ((4096, to_blocks(data, 4096)), (1, np.ascontiguousarray(data.T).reshape(-1))):
        g.set_option("source_block_bytes", block)
        d_in = torch.from_numpy(src).cuda()
        d_out = torch.zeros((nframes, 2), dtype=torch.int32, device="cuda")
        g.decimate_device(d_in, 0, 0, n, 0x69, fir.nlut, nframes, scale, amp, clip, 0, capi.OUT_INT32, d_out)
        assert np.array_equal(d_out.cpu().numpy(), want), block
    with pytest.raises(capi.DsdGpuError):
        g.set_option("source_block_bytes", 48)
    g.close()


def test_32bit_tables_same_bits_from_every_kernel_and_sharding(oracle):
    """lut_precision=32: fixed-point entries summed exactly -- the integer ring kernel, the fp64
    skew and direct kernels and an 8-way sharded run all return the same bits, and those stay
    within BASELINE's tolerance of the fp64 reference (checked tighter: 1e-7 relative)."""
    fir = filters.FILTERS[32]
    nbytes = 400_000
    data = seeded_bytes(2, nbytes, 5)
    n = filters.app_frames(nbytes * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=True)
    outs = {}
    for name in ("ring", "seg", "direct"):
        g = capi.DsdGpu(filters.build_lut(fir, 1), 2, fir.nstep, lut_precision=32)
        g.set_option("kernel", KERNELS[name])
        outs[name] = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
        if name == "ring":
            f64 = g.decimate_host(data, fir.nlut, n, 1.0, out_kind=capi.OUT_FLOAT64)
        g.close()
    assert np.array_equal(outs["ring"], outs["direct"]) and np.array_equal(outs["seg"], outs["direct"])
    job = batch.Job(data, DSD64, 88200, msb_flag=True, bits=24, dither=True, lut_precision=32)
    runner = batch.BatchRunner()
    parts = []
    for r in range(8):
        parts += runner.run([job], r, 8)
    assert np.array_equal(batch.stitch([job], parts)[0], outs["ring"])
    runner.close()
    want_i = oracle.run_app(data, nbytes * 8, 1, DSD64, 88200, scale, amp, clip)
    assert np.abs(outs["ring"].astype(np.int64) - want_i).max() <= 1
    want_f = oracle.run_raw(data, nbytes * 8, 1, DSD64, 88200, fir.nlut + 1, n, 1.0, want="f64")
    assert np.abs(f64 - want_f).max() <= 1e-7 * np.abs(want_f).max()


@pytest.mark.parametrize("ratio", [32, 16, 8])
def test_device_pointers_unaligned_everything(oracle, ratio):
    """Device entry point with nothing aligned: input base and row stride odd, output base only
    4-byte aligned, frame grid off the 4-byte grid, odd channel count -- the staging falls back
    to bytewise assembly and the emitter to scalar stores, the PCM does not change."""
    import torch
    fir, g = make(ratio, 1, 3, "ring")
    n_bytes = 33_333
    data = seeded_bytes(3, n_bytes, 17 + ratio)
    scale, amp, clip = filters.app_params(20, dither=True)
    n = filters.app_frames(n_bytes * 8, fir)
    want = oracle.run_app(data, n_bytes * 8, 1, DSD64, RATES[ratio], scale, amp, clip)
    stride = n_bytes + 7
    raw = torch.zeros(3 * stride + 64, dtype=torch.uint8, device="cuda")
    for c in range(3):
        raw[1 + c * stride: 1 + c * stride + n_bytes] = torch.from_numpy(data[c]).cuda()
    out = torch.zeros(n * 3 + 8, dtype=torch.int32, device="cuda")
    g.decimate_device(raw.data_ptr() + 1, stride, 0, n_bytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32,
                      out.data_ptr() + 4)
    assert np.array_equal(out[1:1 + n * 3].cpu().numpy().reshape(n, 3), want)
    assert int(out[0]) == 0 and int(out[1 + n * 3]) == 0          # nothing written outside
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nch", [2, 4, 6])
def test_even_channels_output_base_only_4_byte_aligned(oracle, nch):
    """Channel pairs are written with 8- / 16-byte stores; a caller's int32 pointer that is only 4-byte aligned (a slice
    of a tensor) must fall back to scalar stores instead of faulting (a misaligned address is a sticky context error)."""
    import torch
    fir, g = make(32, 1, nch, "ring")
    n_bytes = 20_000
    data = seeded_bytes(nch, n_bytes, 40 + nch)
    scale, amp, clip = filters.app_params(24, dither=True)
    n = filters.app_frames(n_bytes * 8, fir)
    want = oracle.run_app(data, n_bytes * 8, 1, DSD64, 88200, scale, amp, clip)
    d_in = torch.from_numpy(data).cuda()
    for off in (1, 2, 3):                                   # 4, 8 and 12 bytes past a 16-byte boundary
        out = torch.zeros(n * nch + 8, dtype=torch.int32, device="cuda")
        g.decimate_device(d_in, n_bytes, 0, n_bytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32,
                          out.data_ptr() + 4 * off)
        assert np.array_equal(out[off:off + n * nch].cpu().numpy().reshape(n, nch), want)
        assert int(out[off - 1]) == 0 and int(out[off + n * nch]) == 0
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["ring", "seg", "direct"])
@pytest.mark.parametrize("nch", [1, 2, 3, 6])
def test_packed_pcm_outputs_carry_the_same_integers(oracle, kernel, nch):
    """DSDGPU_OUT_PCM24 / _PCM16 are the int32 samples narrowed to 3 / 2 little-endian bytes (main.cpp clips to
    2^(bits-1)-1, :245, so nothing is lost); host path (chunked) and device path (any output alignment)."""
    import torch
    for ratio in (32, 16, 8):
        fir, g = make(ratio, 1, nch, kernel)
        n_bytes = 30_011
        data = seeded_bytes(nch, n_bytes, 900 + ratio + nch)
        n = filters.app_frames(n_bytes * 8, fir)
        g.set_option("chunk_frames", 1777)
        for bits, kind in ((24, capi.OUT_PCM24), (20, capi.OUT_PCM24), (16, capi.OUT_PCM16)):
            scale, amp, clip = filters.app_params(bits, dither=True)
            want = oracle.run_app(data, n_bytes * 8, 1, DSD64, RATES[ratio], scale, amp, clip)
            got = g.decimate_host(data, fir.nlut, n, scale, amp, clip, out_kind=kind)
            got = capi.unpack_pcm24(got) if kind == capi.OUT_PCM24 else got.astype(np.int32)
            assert np.array_equal(got, want), (ratio, bits)
            d_in = torch.from_numpy(data).cuda()
            elem = 3 if kind == capi.OUT_PCM24 else 2
            for off in (0, 1, 2, 4):
                raw = torch.full((n * nch * elem + 32,), 0x5A, dtype=torch.uint8, device="cuda")
                g.decimate_device(d_in, n_bytes, 0, n_bytes, 0x69, fir.nlut, n, scale, amp, clip, 0, kind, raw.data_ptr() + off)
                host = raw.cpu().numpy()
                body = host[off:off + n * nch * elem]
                got = capi.unpack_pcm24(body.reshape(n, nch, 3)) if elem == 3 else body.view(np.int16).reshape(n, nch).astype(np.int32) if off % 2 == 0 \
                    else np.frombuffer(body.tobytes(), dtype=np.int16).reshape(n, nch).astype(np.int32)
                assert np.array_equal(got, want), (ratio, bits, off)
                assert (host[:off] == 0x5A).all() and (host[off + n * nch * elem:] == 0x5A).all()     # nothing written outside
        # the packed kinds refuse a clip that does not guarantee the range
        with pytest.raises(capi.DsdGpuError, match="needs 0 < clip"):
            g.decimate_host(data, fir.nlut, 16, 1.0, 0.0, 0.0, out_kind=capi.OUT_PCM24)
        with pytest.raises(capi.DsdGpuError, match="needs 0 < clip"):
            g.decimate_host(data, fir.nlut, 16, 1.0, 0.0, 8388607.0, out_kind=capi.OUT_PCM16)
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 3])
def test_c_batch_runner_on_gpu(oracle, world):
    """dsdgpu_run_batch on the GPU: mixed files (three filters, 1-6 channels, dither on and off, both bit orders), every
    split fills each job's interleaved output exactly once and the whole equals the oracle; int32 and packed 24-bit."""
    from dsf2flac_b200 import batch
    from tests.test_sharding import mixed_jobs
    jobs = mixed_jobs(7)
    want = []
    for j in jobs:
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        want.append(oracle.run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip))
    for kind in (capi.OUT_INT32, capi.OUT_PCM24):
        total = [np.zeros_like(w) for w in want]
        for r in range(world):
            outs = batch.run_batch_c(jobs, devices=(0,), rank=r, world=world, align_frames=64, out_kind=kind)
            for t, o in zip(total, outs):
                o = capi.unpack_pcm24(o) if kind == capi.OUT_PCM24 else o
                assert not ((t != 0) & (o != 0)).any()
                t += o
        for t, w in zip(total, want):
            assert np.array_equal(t, w)



@pytest.mark.gpu
def test_c_batch_runner_all_devices_of_one_process(oracle):
    """dsdgpu_run_batch with every visible GPU in ONE process (one host thread per device): the shape a C++ caller uses on
    a multi-GPU box.  With one GPU this is the single-device path again; on the 2/4/8-GPU boxes it is the real thing."""
    import torch
    from dsf2flac_b200 import batch
    from tests.test_sharding import mixed_jobs
    ndev = torch.cuda.device_count()
    jobs = mixed_jobs(11)
    outs = batch.run_batch_c(jobs, devices=tuple(range(ndev)), rank=0, world=1, align_frames=64)
    for j, got in zip(jobs, outs):
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        want = oracle.run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip)
        assert np.array_equal(got, want)
