"""EXTENSION filters: decimation by 64, 128 and 256 -- DSD512 -> 352.8 / 176.4 / 88.2 kHz and
the like, which the reference rejects (dsd_decimator.cpp:57-68).  No reference parity exists for these
ratios; what is pinned is the reference's OPERATOR on the extension coefficients:
  * tests/golden/golden_ext_v1.npz -- produced by the unmodified reference arithmetic (oracle/_ref with
    the filter installed through its own initLookupTable, tests/golden/make_golden_ext.py),
  * the C restatement with the same filter registered (bit-equal to both),
  * the GPU path (segmented kernel: chain of <= 72-table passes over running fp64 sums; direct kernel in
    96-table passes) bit-equal to the restatement, int32 and fp64.
Without the switch the product answers exactly like the reference: "Sorry, incompatible ..."."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from dsf2flac_b200 import capi, filters
from tests.oracle_lib import DSD64, seeded_bytes

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ext_v1.npz"))
DSD512 = 8 * DSD64
CASES = {64: (DSD512, 352800), 128: (DSD512, 176400), 256: (DSD512, 88200)}
S24, C24 = 13295047.713424837, 8388607.0


@pytest.fixture()
def ext_oracle(oracle):
    """The checker (the reference's own arithmetic where oracle/_ref is present, else the C restatement) with one
    extension filter registered at a time."""
    class Ext:
        def use(self, ratio):
            oracle.set_custom_filter(filters.EXTENSION_FILTERS[ratio])
            return oracle
    yield Ext()
    oracle.set_custom_filter(None)


@pytest.fixture()
def ext_port(port):
    """The C restatement itself (the CPU test that pins it to the golden vectors of the reference's arithmetic)."""
    class Ext:
        def use(self, ratio):
            port.set_custom_filter(filters.EXTENSION_FILTERS[ratio])
            return port
    yield Ext()
    port.set_custom_filter(None)


# ---- CPU -------------------------------------------------------------------------------------

@pytest.mark.parametrize("ratio", list(CASES))
def test_extension_filter_shape_and_response(ratio):
    fir = filters.EXTENSION_FILTERS[ratio]
    k = ratio // 32
    assert (fir.ntaps, fir.tzero, fir.nlut, fir.nstep) == (18 * ratio - 1, 9 * ratio, 18 * ratio // 8, ratio // 8)
    assert fir.nlut / fir.nstep - fir.tzero / fir.ratio == 9.0          # getFirstValidSample as for /32
    h, h32 = fir.coefs, filters.FILTERS[32].coefs
    assert np.array_equal(h, h[::-1]) and abs(h.sum() - h32.sum()) < 1e-15
    # the /32 response against absolute frequency, with the DSD clock k times faster
    n = 1 << 20
    H, H32 = np.abs(np.fft.rfft(h, n)), np.abs(np.fft.rfft(h32, n))
    idx = np.arange(0, n // 2 // k, 16)
    a, b = H[idx], H32[idx * k]
    passband = b > 10 ** (-3 / 20)
    assert np.max(np.abs(20 * np.log10(a[passband] / b[passband]))) < 1e-4
    assert 20 * np.log10(np.max(a[b < 1e-5])) < -99.0                     # the /32 filter's -100 dB stop band
    assert 20 * np.log10(np.max(H[n // 2 // k:])) < -120.0                # nothing comes back above the old Nyquist


def test_reference_behaviour_without_the_switch(oracle):
    for dsd_fs, out_fs in CASES.values():
        assert oracle.info(dsd_fs, out_fs) is None
        assert filters.pick(dsd_fs, out_fs) is None
        assert filters.pick(dsd_fs, out_fs, extensions=True).ratio == dsd_fs // out_fs


@pytest.mark.parametrize("ratio", list(CASES))
def test_restatement_equals_reference_arithmetic_golden(ext_port, ratio):
    dsd_fs, out_fs = CASES[ratio]
    orc = ext_port.use(ratio)
    info = orc.info(dsd_fs, out_fs, 54321 * 8)
    assert [info["ratio"], info["nlut"], info["nstep"], info["tzero"], info["length_pcm"]] == list(GOLD[f"info_{ratio}"])
    assert [info["first_valid"], info["last_valid"]] == list(GOLD[f"valid_{ratio}"])
    for msb in (0, 1):
        tag = f"r{ratio}_m{msb}"
        lut = orc.lut(dsd_fs, out_fs, msb)
        assert hashlib.sha256(lut.tobytes()).digest() == GOLD[f"lut_sha_{tag}"].tobytes()
        assert filters.build_lut(filters.EXTENSION_FILTERS[ratio], msb).tobytes() == lut.tobytes()
        data, L = GOLD[f"in_{tag}"], 40 * ratio * 8
        assert np.array_equal(orc.run_app(data, L, msb, dsd_fs, out_fs, S24, 0.0, C24), GOLD[f"app24_{tag}"])
        assert np.array_equal(orc.run_app(data, L, msb, dsd_fs, out_fs, 51933.78013056577, 1.0, 32767.0), GOLD[f"app16d_{tag}"])
        assert orc.run_raw(data, L, msb, dsd_fs, out_fs, 0, 48, 1.0, want="f64").tobytes() == GOLD[f"raw64_{tag}"].tobytes()
        assert len(GOLD[f"app24_{tag}"]) == filters.app_frames(L, filters.EXTENSION_FILTERS[ratio]) == 40 * 8 - 17


@pytest.mark.parametrize("ratio", list(CASES))
def test_restatement_equals_reference_arithmetic_live(ext_oracle, ref, ratio):
    dsd_fs, out_fs = CASES[ratio]
    orc = ext_oracle.use(ratio)
    ref.set_custom_filter(filters.EXTENSION_FILTERS[ratio])
    try:
        data = seeded_bytes(3, 90 * ratio, ratio)
        for L in (90 * ratio * 8, 90 * ratio * 8 - 13):
            for bits, dither in ((24, 0.0), (16, 1.0)):
                scale, _, clip = filters.app_params(bits)
                assert np.array_equal(orc.run_app(data, L, 1, dsd_fs, out_fs, scale, dither, clip),
                                      ref.run_app(data, L, 1, dsd_fs, out_fs, scale, dither, clip))
        a = orc.run_raw(data, 90 * ratio * 8, 0, dsd_fs, out_fs, 0, 120, 3.0, want="f64")
        b = ref.run_raw(data, 90 * ratio * 8, 0, dsd_fs, out_fs, 0, 120, 3.0, want="f64")
        assert a.tobytes() == b.tobytes()
    finally:
        ref.set_custom_filter(None)


def test_drop_in_class_with_the_switch(ext_oracle):
    """Host logic through the CPU stand-in of the C ABI: rejected by default, valid and bit-equal to the
    restatement with DsdDecimator::setExtensionFilters(true)."""
    from tests.standin import cpu_host_lib
    host = cpu_host_lib()
    r, a, b, n = C.c_int32(), C.c_double(), C.c_double(), C.c_int64()
    assert host.dsdhost_info(DSD512, 352800, 0, C.byref(r), C.byref(a), C.byref(b), C.byref(n)) == 0
    assert b"incompatible sample rate" in host.dsdhost_last_error()
    host.dsdhost_set_extension_filters(1)
    try:
        for ratio, (dsd_fs, out_fs) in CASES.items():
            orc = ext_oracle.use(ratio)
            assert host.dsdhost_info(dsd_fs, out_fs, 54321 * 8, C.byref(r), C.byref(a), C.byref(b), C.byref(n)) == 1
            info = orc.info(dsd_fs, out_fs, 54321 * 8)
            assert (r.value, a.value, b.value, n.value) == (ratio, info["first_valid"], info["last_valid"], info["length_pcm"])
            data = seeded_bytes(2, 60 * ratio, 5 + ratio)
            L = 60 * ratio * 8 - 3
            scale, amp, clip = filters.app_params(24, dither=True)
            cap = L // 8 + 2048
            out = np.zeros((cap, 2), dtype=np.int32)
            got = host.dsdhost_run_app(data.ctypes.data, 2, data.shape[1], L, 0x69, 1, dsd_fs, out_fs, scale, amp, clip, 64,
                                       37, out.ctypes.data, cap)
            assert got >= 0, host.dsdhost_last_error()
            assert np.array_equal(out[:got], orc.run_app(data, L, 1, dsd_fs, out_fs, scale, amp, clip))
    finally:
        host.dsdhost_set_extension_filters(0)


def test_segment_layout_fits_shared_memory():
    """seg_core.h's budget, restated: table image + two byte tiles per CTA within 227 KB for the channel
    counts and frame hops the kernel is offered (KS = samples per thread falls back 4 -> 2 -> 1)."""
    lut_bytes, limit = 128 + 256 * 96 * 8, 232448

    def smem(samples, nch, nstep):
        frames = (samples + nch - 1) // nch + 1
        stride = ((frames - 1) * nstep + 160 + 15) // 16 * 16
        return lut_bytes + 2 * nch * stride
    for nstep in (1, 2, 4, 8, 16, 32):
        for nch in (1, 2, 3, 5, 6, 8):
            assert any(smem(512 * ks, nch, nstep) <= limit for ks in (4, 2, 1)), (nstep, nch)
    assert smem(2048, 2, 8) <= limit and smem(2048, 6, 8) <= limit       # /64 stereo and 5.1 run with KS = 4


# ---- GPU -------------------------------------------------------------------------------------

def make(ratio, msb, nch, kernel, precision=64):
    fir = filters.EXTENSION_FILTERS.get(ratio) or filters.FILTERS[ratio]
    g = capi.DsdGpu(filters.build_lut(fir, msb), nch, fir.nstep, lut_precision=precision)
    g.set_option("kernel", kernel)
    return fir, g


@pytest.mark.gpu
@pytest.mark.parametrize("ratio,msb", [(r, m) for r in CASES for m in (0, 1)])
def test_gpu_golden_extension(ratio, msb):
    tag = f"r{ratio}_m{msb}"
    for kernel in (capi.KERNEL_AUTO, capi.KERNEL_SEGMENTED, capi.KERNEL_DIRECT) + ((capi.KERNEL_RING,) if ratio == 64 else ()):
        fir, g = make(ratio, msb, 2, kernel)
        data, L = GOLD[f"in_{tag}"], 40 * ratio * 8
        n = filters.app_frames(L, fir)
        assert np.array_equal(g.decimate_host(data, fir.nlut, n, S24, 0.0, C24), GOLD[f"app24_{tag}"])
        assert np.array_equal(g.decimate_host(data, fir.nlut, n, 51933.78013056577, 1.0, 32767.0), GOLD[f"app16d_{tag}"])
        got = g.decimate_host(data, -1, 48, 1.0, out_kind=capi.OUT_FLOAT64)
        assert got.tobytes() == GOLD[f"raw64_{tag}"].tobytes()
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nch", [1, 2, 3, 6])
@pytest.mark.parametrize("ratio", list(CASES))
def test_gpu_extension_parity_channels_chunks_dither(ext_oracle, ratio, nch):
    dsd_fs, out_fs = CASES[ratio]
    orc = ext_oracle.use(ratio)
    fir, g = make(ratio, 1, nch, capi.KERNEL_AUTO)
    nbytes = 700 * ratio + 13
    data = seeded_bytes(nch, nbytes, 31 * ratio + nch)
    L = nbytes * 8 - 5
    n = filters.app_frames(L, fir)
    for bits, dither, chunk in ((24, False, 1 << 20), (24, True, 1531), (16, True, 4096)):
        g.set_option("chunk_frames", chunk)
        scale, amp, clip = filters.app_params(bits, dither=dither)
        want = orc.run_app(data, L, 1, dsd_fs, out_fs, scale, amp, clip)
        assert len(want) == n
        got = g.decimate_host(data[:, :-(-L // 8) + 1], fir.nlut, n, scale, amp, clip)
        assert np.array_equal(got, want), (bits, dither, chunk)
    # fp64 sums straight after rewind (idle pre-fill) and past the end of the data, both kernels
    nframes = (nbytes + 2 * fir.nlut) // fir.nstep
    want = orc.run_raw(data, nbytes * 8, 1, dsd_fs, out_fs, 0, nframes, 1.0, want="f64")
    for kernel in (capi.KERNEL_SEGMENTED, capi.KERNEL_DIRECT) + ((capi.KERNEL_RING,) if ratio == 64 else ()):
        g.set_option("kernel", kernel)
        g.set_option("chunk_frames", 2000)
        got = g.decimate_host(data, -1, nframes, 1.0, out_kind=capi.OUT_FLOAT64)
        assert got.tobytes() == want.tobytes(), kernel
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ratio,fs", [(32, 88200), (16, 176400), (8, 352800)])
def test_gpu_segmented_kernel_on_the_reference_filters(oracle, ratio, fs):
    """The any-geometry kernel must also return the reference's bits for the reference's own filters."""
    fir, g = make(ratio, 0, 3, capi.KERNEL_SEGMENTED)
    data = seeded_bytes(3, 50000, ratio)
    nframes = (50000 + 300) // fir.nstep
    want = oracle.run_raw(data, 50000 * 8, 0, DSD64, fs, 0, nframes, 1.0, want="f64")
    got = g.decimate_host(data, -1, nframes, 1.0, out_kind=capi.OUT_FLOAT64)
    assert got.tobytes() == want.tobytes()
    scale, amp, clip = filters.app_params(20, dither=True)
    want = oracle.run_app(data, 50000 * 8, 0, DSD64, fs, scale, amp, clip)
    got = g.decimate_host(data, fir.nlut, len(want), scale, amp, clip)
    assert np.array_equal(got, want)
    g.close()


@pytest.mark.gpu
def test_gpu_extension_source_layouts_and_windows(ext_oracle):
    """DSF block-planar and DSDIFF byte-interleaved sources, and a caller that holds only a window."""
    ratio = 64
    dsd_fs, out_fs = CASES[ratio]
    orc = ext_oracle.use(ratio)
    nch, nbytes = 2, 8 * 4096
    data = seeded_bytes(nch, nbytes, 99)
    fir = filters.EXTENSION_FILTERS[ratio]
    lut = filters.build_lut(fir, 1)
    nframes = (nbytes + 300) // fir.nstep
    want = orc.run_raw(data, nbytes * 8, 1, dsd_fs, out_fs, 0, nframes, 2.0, want="f64")
    for block in (4096, 1):
        g = capi.DsdGpu(lut, nch, fir.nstep)
        g.set_option("source_block_bytes", block)
        if block == 1:
            src = np.ascontiguousarray(data.T).reshape(-1)
        else:
            src = np.ascontiguousarray(data.reshape(nch, nbytes // block, block).transpose(1, 0, 2)).reshape(-1)
        got = np.empty((nframes, nch))
        g.decimate_host_raw(src.ctypes.data, 0, 0, nbytes, 0x69, -1, nframes, 2.0, 0.0, 0.0, 0, capi.OUT_FLOAT64, got.ctypes.data)
        assert got.tobytes() == want.tobytes(), block
        g.close()
    g = capi.DsdGpu(lut, nch, fir.nstep)
    for first, count in ((16, 20000), (3, 12345), (30000, 10)):
        masked = np.full_like(data, 0x69)
        masked[:, first:first + count] = data[:, first:first + count]
        want = orc.run_raw(masked, nbytes * 8, 1, dsd_fs, out_fs, 0, nframes, 2.0, want="f64")
        got = np.empty((nframes, nch))
        g.decimate_host_raw(data.ctypes.data + first, nbytes, first, count, 0x69, -1, nframes, 2.0, 0.0, 0.0, 0,
                            capi.OUT_FLOAT64, got.ctypes.data)
        assert got.tobytes() == want.tobytes(), (first, count)
    g.close()


@pytest.mark.gpu
def test_gpu_drop_in_class_extension(ext_oracle):
    """The drop-in DsdDecimator with the switch on, DSD512 -> 352.8 kHz, main.cpp's loop."""
    host = capi.host_lib()
    host.dsdhost_set_extension_filters(1)
    try:
        orc = ext_oracle.use(64)
        data = seeded_bytes(2, 300000, 64)
        L = 300000 * 8
        scale, amp, clip = filters.app_params(24, dither=True)
        cap = L // 8 + 2048
        out = np.zeros((cap, 2), dtype=np.int32)
        for fn in (host.dsdhost_run_app, host.dsdhost_run_app_bulk):
            got = fn(data.ctypes.data, 2, data.shape[1], L, 0x69, 1, DSD512, 352800, scale, amp, clip, 64, 0, out.ctypes.data, cap)
            assert got >= 0, host.dsdhost_last_error()
            assert np.array_equal(out[:got], orc.run_app(data, L, 1, DSD512, 352800, scale, amp, clip))
    finally:
        host.dsdhost_set_extension_filters(0)


def _ext_jobs():
    from dsf2flac_b200 import batch
    return [batch.Job(seeded_bytes(2, 64 * 700, 1), DSD512, 352800, msb_flag=True, bits=24, dither=True, extensions=True,
                      length_samples=64 * 700 * 8 - 11),
            batch.Job(seeded_bytes(3, 128 * 300, 2), DSD512, 176400, msb_flag=False, bits=16, dither=True, extensions=True),
            batch.Job(seeded_bytes(1, 5000, 3), DSD64, 88200, bits=20, dither=False)]


def _check_stitched(orc_for, jobs, results):
    from dsf2flac_b200 import batch
    for j, got in zip(jobs, batch.stitch(jobs, results)):
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        want = orc_for(j).run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip)
        assert np.array_equal(got, want), j.fir.ratio


@pytest.mark.parametrize("world", [1, 3])
def test_extension_shards_stitch_cpu(ext_oracle, oracle, world):
    """Time segments of a divide-by-64 / -128 stream carry a 143- / 287-byte halo; the plan, the halo and the
    positional dither give every world size the unsharded result (CPU stand-in of the C ABI)."""
    from dsf2flac_b200 import batch
    from tests import standin
    jobs = _ext_jobs()
    runner = batch.BatchRunner(make_ctx=standin.make_ctx)
    results = []
    for r in range(world):
        results += runner.run(jobs, r, world, align_frames=64)
    _check_stitched(lambda j: ext_oracle.use(j.fir.ratio) if j.extensions else oracle, jobs, results)


@pytest.mark.gpu
def test_gpu_extension_shards_stitch(ext_oracle, oracle):
    from dsf2flac_b200 import batch
    jobs = _ext_jobs()
    runner = batch.BatchRunner()
    results = []
    for r in range(4):
        results += runner.run(jobs, r, 4, align_frames=64)
    runner.close()
    _check_stitched(lambda j: ext_oracle.use(j.fir.ratio) if j.extensions else oracle, jobs, results)


@pytest.mark.gpu
def test_gpu_odd_geometries_and_limits():
    """Geometries nobody ships: the any-geometry kernels take them (or decline with a message), never crash.
    100 tables x 3-byte frame hop, 5 channels; 40 channels at a 32-byte hop; 100 channels at a 32-byte hop (does not
    fit the segmented kernel's shared memory: AUTO falls back to the direct kernel, asking for the segmented one is an error)."""
    rng = np.random.default_rng(5)

    def reference_sums(lut, data, nstep, p0, nframes, fill=0x69):
        nch, nbytes = data.shape
        out = np.zeros((nframes, nch))
        for i in range(nframes):
            for c in range(nch):
                s = 0.0
                for t in range(lut.shape[0]):
                    j = p0 + i * nstep - t
                    s = s + lut[t, data[c, j] if 0 <= j < nbytes else fill]
                out[i, c] = s
        return out

    lut = (rng.random((100, 256)) - 0.5) * 0.3
    data = seeded_bytes(5, 700, 8)
    want = reference_sums(lut, data, 3, 40, 230)
    for kernel in (capi.KERNEL_AUTO, capi.KERNEL_SEGMENTED, capi.KERNEL_DIRECT):
        with capi.DsdGpu(lut, 5, 3) as g:
            g.set_option("kernel", kernel)
            got = g.decimate_host(data, 40, 230, 1.0, out_kind=capi.OUT_FLOAT64)
            assert got.tobytes() == want.tobytes(), kernel
    lut = (rng.random((80, 256)) - 0.5) * 0.3
    data = seeded_bytes(40, 1200, 9)
    want = reference_sums(lut, data, 32, 100, 30)
    for kernel in (capi.KERNEL_AUTO, capi.KERNEL_SEGMENTED, capi.KERNEL_DIRECT):        # 40 channels at a 32-byte hop
        with capi.DsdGpu(lut, 40, 32) as g:
            g.set_option("kernel", kernel)
            got = g.decimate_host(data, 100, 30, 1.0, out_kind=capi.OUT_FLOAT64)
            assert got.tobytes() == want.tobytes(), kernel
    data = seeded_bytes(100, 900, 10)
    want = reference_sums(lut, data, 32, 100, 20)
    with capi.DsdGpu(lut, 100, 32) as g:                                                # 100 channels: no room for the tiles
        got = g.decimate_host(data, 100, 20, 1.0, out_kind=capi.OUT_FLOAT64)
        assert got.tobytes() == want.tobytes()
        g.set_option("kernel", capi.KERNEL_SEGMENTED)
        with pytest.raises(capi.DsdGpuError, match="shared memory"):
            g.decimate_host(data, 100, 20, 1.0, out_kind=capi.OUT_FLOAT64)
    for nlut, nstep in ((641, 8), (72, 65), (0, 4)):
        with pytest.raises(capi.DsdGpuError):
            capi.DsdGpu(np.zeros((max(nlut, 1), 256)) if nlut else np.zeros((1, 256))[:0].reshape(0, 256), 2, nstep)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_extension_random_geometry(ext_oracle, seed):
    """Random geometry for the ring passes (and the segmented kernel): 1-6 channels, windows that start before the data,
    odd chunk sizes, dither on and off at random positions of the dither stream (tools/fuzz_more.py runs more of these)."""
    rng = np.random.default_rng(5000 + seed)
    ratio = int(rng.choice([64, 128, 256]))
    fir = filters.EXTENSION_FILTERS[ratio]
    msb = int(rng.integers(0, 2))
    nch = int(rng.integers(1, 7))
    nbytes = int(rng.integers(2000, 120000))
    data = rng.integers(0, 256, size=(nch, nbytes), dtype=np.uint8)
    p0 = int(rng.integers(-5, max(1, nbytes // 2)))
    nframes = int(rng.integers(0, (nbytes - p0 + 500) // fir.nstep))
    dither = bool(rng.integers(0, 2))
    scale, amp, clip = filters.app_params(int(rng.choice([16, 20, 24])), dither=dither)
    fill = int(rng.choice([0x69, 0x96, 0x00]))
    kernel = str(rng.choice(["auto", "auto", "seg"]))
    sample0 = int(rng.integers(0, 1_000_000)) if dither and rng.random() < 0.5 else 0
    want = ext_oracle.use(ratio).run_raw(data, 1 << 40, msb, DSD512, DSD512 // ratio, p0 + 1, nframes, scale, amp, clip, idle=fill,
                                         rand_skip=2 * sample0)
    g = capi.DsdGpu(filters.build_lut(fir, msb), nch, fir.nstep)
    g.set_option("kernel", {"auto": capi.KERNEL_AUTO, "seg": capi.KERNEL_SEGMENTED}[kernel])
    g.set_option("chunk_frames", int(rng.integers(1, 9000)))
    got = np.empty((nframes, nch), dtype=np.int32)
    g.decimate_host_raw(data.ctypes.data, nbytes, 0, nbytes, fill, p0, nframes, scale, amp, clip, sample0, capi.OUT_INT32,
                        got.ctypes.data)
    g.close()
    assert np.array_equal(got, want), dict(ratio=ratio, nch=nch, nbytes=nbytes, p0=p0, nframes=nframes, dither=dither, kernel=kernel)
