"""the containers either side of the hot path.

CPU part (no GPU): the product's DSF / DSDIFF header parser and its model of "which stream bytes
does the reference's reader deliver" are held against the reference's OWN readers + decimator
(oracle/_ref, one conversion per process) on synthetic files covering the awkward ends: data
that stops mid-block, exactly on a block, files that end with the sample chunk (the reference
then drops the tail of the last DSDIFF buffer), files with a chunk / metadata behind it.  The
drop-in's bulk path is also run end to end against a CPU stand-in of the C ABI.

GPU part: the same files through dsdhost_convert_file (file payload uploaded in its own layout,
block-planar addressing / device de-interleave) must equal the reference's output bit for bit.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from dsf2flac_b200 import capi, filters
from tests.dsd_files import write_dsdiff, write_dsf
from tests.oracle_lib import DSD64, REF_SO, Oracle, seeded_bytes
from tests.standin import cpu_host_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(REF_SO)
SCALE, CLIP = filters.app_params(24, dither=False)[0], 8388607.0

# (tag, kind, nbytes per channel, nch, writer kwargs)
CASES = [
    ("dsf_midblock", 1, 5000, 2, {}),
    ("dsf_two_blocks_exact", 1, 8192, 2, {}),
    ("dsf_exact_with_trailer", 1, 8192, 2, {"trailer": bytes(range(256)) * 40}),
    ("dsf_small_blocks_6ch", 1, 3000, 6, {"block": 256}),
    ("dsf_mono", 1, 4100, 1, {}),
    ("dff_tail_dropped", 2, 5000, 2, {}),
    ("dff_followed_by_chunk", 2, 5000, 2, {"trailer_chunk": True}),
    ("dff_buffer_multiple", 2, 5120, 2, {}),
    ("dff_one_short_of_buffer", 2, 2047, 2, {}),
    ("dff_small", 2, 700, 2, {}),
    ("dff_6ch", 2, 4000, 6, {"trailer_chunk": True}),
    ("dff_mono_odd", 2, 3001, 1, {}),
]


def make_file(tmp_path, tag, kind, nbytes, nch, kw):
    data = seeded_bytes(nch, nbytes, len(tag) * 131 + nbytes)
    path = str(tmp_path / (tag + (".dsf" if kind == 1 else ".dff")))
    (write_dsf if kind == 1 else write_dsdiff)(path, data, DSD64, **kw)
    return path, data


def reference_file_run(path, kind, tmp_path, dither=0.0):
    dst = str(tmp_path / "ref.npz")
    r = subprocess.run([sys.executable, "-m", "tests.ref_file_runner", path, str(kind), "88200", repr(SCALE), repr(dither),
                        repr(CLIP), dst], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.load(dst)


def file_info(lib, path, kind=0):
    info = (C.c_int32 * 5)()
    length, delivered = C.c_int64(), C.c_int64()
    ok = lib.dsdhost_file_info(path.encode(), kind, info, C.byref(length), C.byref(delivered))
    return ok, list(info), length.value, delivered.value


@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref (the reference's readers)")
@pytest.mark.parametrize("tag,kind,nbytes,nch,kw", CASES, ids=[c[0] for c in CASES])
def test_container_model_against_reference_readers(tmp_path, tag, kind, nbytes, nch, kw):
    lib = cpu_host_lib()
    path, data = make_file(tmp_path, tag, kind, nbytes, nch, kw)
    ref = reference_file_run(path, kind, tmp_path)
    ok, info, length, delivered = file_info(lib, path)
    assert ok == 1
    assert info[:4] == [int(ref["channels"]), int(ref["fs"]), int(ref["msb"]), kind]
    assert length == int(ref["length"])
    # the bytes the container path will feed the GPU, run through the CPU oracle's decimator
    planar = np.zeros((nch, max(delivered, 1)), dtype=np.uint8)
    for c in range(nch):
        assert lib.dsdhost_file_channel_bytes(path.encode(), 0, c, planar[c].ctypes.data, delivered) == delivered
    want = Oracle().run_app(planar[:, :delivered], length, info[2], DSD64, 88200, SCALE, 0.0, CLIP)
    assert want.shape == ref["pcm"].shape and np.array_equal(want, ref["pcm"])
    # and the drop-in with a bulk source, end to end on the CPU stand-in of the C ABI
    out = np.zeros((len(want) + 8, nch), dtype=np.int32)
    n = lib.dsdhost_convert_file(path.encode(), 0, 88200, SCALE, 0.0, CLIP, 64, 777, out.ctypes.data, len(out))
    assert n == len(want) and np.array_equal(out[:n], ref["pcm"])


def test_container_rejects():
    lib = cpu_host_lib()
    assert file_info(lib, "/nonexistent/file.dsf")[0] == 0
    assert lib.dsdhost_last_error().decode() == "could not open file"
    assert file_info(lib, os.path.join(ROOT, "README.md"))[0] == 0


def test_hostile_headers_are_refused(tmp_path):
    """Header fields of untrusted files: a negative / overflowing DSF sample count, DSDIFF chunk sizes whose +12/+13
    wraps (a top-level chunk of "size" 2^64-12 would never advance the walk).  Each must come back as an error."""
    import struct
    lib = cpu_host_lib()
    data = seeded_bytes(2, 5000, 9)
    for k, count in enumerate((-1, -8 * 4096, (1 << 63) - 1, (1 << 63) - 8)):
        p = str(tmp_path / f"bad{k}.dsf")
        write_dsf(p, data, DSD64, sample_count=count & 0xFFFFFFFFFFFFFFFF)
        assert file_info(lib, p)[0] == 0 and b"sample count" in lib.dsdhost_last_error()
    good = tmp_path / "good.dff"
    write_dsdiff(str(good), data, DSD64)
    raw = good.read_bytes()
    assert file_info(lib, str(good))[0] == 1
    for k, size in enumerate((0xFFFFFFFFFFFFFFF4, 0xFFFFFFFFFFFFFFF3, 1 << 63)):
        p = tmp_path / f"bad{k}.dff"
        p.write_bytes(b"JUNK" + struct.pack(">Q", size) + raw)          # a leading chunk whose size wraps
        assert file_info(lib, str(p), 2)[0] == 0 and b"chunk size" in lib.dsdhost_last_error()
        at = raw.index(b"PROP")
        p.write_bytes(raw[:at + 4] + struct.pack(">Q", size) + raw[at + 12:])
        assert file_info(lib, str(p))[0] == 0


def test_dst_compressed_dsdiff_is_refused(tmp_path):
    import struct
    from tests.dsd_files import _chunk
    prop = b"SND " + _chunk(b"FS  ", struct.pack(">I", DSD64)) + _chunk(b"CHNL", struct.pack(">H", 2) + b"SLFTSRGT") \
        + _chunk(b"CMPR", b"DST " + bytes([11]) + b"DST Encoded")
    body = b"DSD " + _chunk(b"FVER", struct.pack(">I", 0x01050000)) + _chunk(b"PROP", prop) + _chunk(b"DST ", b"\0" * 64)
    p = tmp_path / "x.dff"
    p.write_bytes(b"FRM8" + struct.pack(">Q", len(body)) + body)
    lib = cpu_host_lib()
    assert file_info(lib, str(p))[0] == 0 and b"DST" in lib.dsdhost_last_error()


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref (the reference's readers)")
@pytest.mark.parametrize("tag,kind,nbytes,nch,kw", CASES, ids=[c[0] for c in CASES])
def test_gpu_file_conversion_equals_reference(tmp_path, tag, kind, nbytes, nch, kw):
    host = capi.host_lib()
    path, data = make_file(tmp_path, tag, kind, nbytes, nch, kw)
    for dither in (0.0, 1.0):
        ref = reference_file_run(path, kind, tmp_path, dither)
        out = np.zeros((len(ref["pcm"]) + 8, nch), dtype=np.int32)
        n = host.dsdhost_convert_file(path.encode(), 0, 88200, SCALE, dither, CLIP, 64, 1000, out.ctypes.data, len(out))
        assert n == len(ref["pcm"]), host.dsdhost_last_error()
        assert np.array_equal(out[:n], ref["pcm"])


@pytest.mark.gpu
def test_gpu_large_dsf_file_matches_planar_path(tmp_path):
    """A 20 MB stereo DSF file: the container path (block-planar, chunked pipeline) equals the
    planar path on the same stream."""
    from tests.test_gpu_fullsize import sdm_stream
    fs, nbytes = 2 * DSD64, 10_000_000
    data = sdm_stream(2, nbytes, fs)
    path = str(tmp_path / "big.dsf")
    write_dsf(path, data, fs)
    fir = filters.pick(fs, 176400)
    n = filters.app_frames(nbytes * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=True)
    with capi.DsdGpu(filters.build_lut(fir, 1), 2, fir.nstep) as g:
        padded = np.zeros((2, nbytes + 1), dtype=np.uint8)      # the byte after the data is the block's zero padding
        padded[:, :nbytes] = data
        want = g.decimate_host(padded, fir.nlut, n, scale, amp, clip)
    out = np.zeros((n + 8, 2), dtype=np.int32)
    got = capi.host_lib().dsdhost_convert_file(path.encode(), 0, 176400, scale, amp, clip, 64, 0, out.ctypes.data, len(out))
    assert got == n and np.array_equal(out[:n], want)
