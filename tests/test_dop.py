"""DoP packing (reference src/dop_packer.cpp, dop_track_helper in main.cpp).

CPU: the oracle's restatement against the reference's DopPacker (oracle/_ref) on raw calls, and
the container path + frame arithmetic (dsdhost_dop_file over the stand-in ABI) against the
reference's reader + DopPacker on files.  GPU: the kernel through the device and host entry points
in all three source layouts against the oracle, and whole files against the reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dsf2flac_b200 import capi
from tests.oracle_lib import DSD64, REF_SO, Oracle, seeded_bytes
from tests.standin import cpu_host_lib
from tests.test_containers import CASES, ROOT, make_file

HAVE_REF = os.path.exists(REF_SO)


def reference_dop_file(path, kind, tmp_path):
    dst = str(tmp_path / "refdop.npz")
    r = subprocess.run([sys.executable, "-m", "tests.ref_file_runner", path, str(kind), "0", "0", "0", "0", dst], cwd=ROOT,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.load(dst)


@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref")
@pytest.mark.parametrize("nch,nbytes,length,msb,pre,nframes", [
    (2, 5000, 40000, 1, 0, 2600),       # straight after rewind: every marker 0xFA, runs past the end
    (2, 5000, 40000, 1, 1, 2600),       # after main.cpp's creep step: markers alternate 0x05 / 0xFA
    (3, 700, 5597, 0, 1, 360),
    (1, 100, 800, 1, 5, 70),
    (6, 4096, 32768, 0, 2, 2049),
])
def test_oracle_dop_equals_reference_packer(oracle, ref, nch, nbytes, length, msb, pre, nframes):
    data = seeded_bytes(nch, nbytes, nbytes + pre)
    want = ref.dop_run_raw(data, length, msb, pre, nframes, chunk_frames=7)
    assert np.array_equal(oracle.dop_run_raw(data, length, msb, pre, nframes), want)
    markers = (want.astype(np.int64) >> 16) & 0xFF
    assert set(np.unique(markers)) <= {0x05, 0xFA}
    if pre % 2 == 1:
        assert np.all(markers[0::2] == 0x05) and np.all(markers[1::2] == 0xFA)


@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref")
@pytest.mark.parametrize("tag,kind,nbytes,nch,kw", CASES, ids=[c[0] for c in CASES])
def test_dop_file_path_against_reference(tmp_path, oracle, tag, kind, nbytes, nch, kw):
    lib = cpu_host_lib()
    path, data = make_file(tmp_path, tag, kind, nbytes, nch, kw)
    ref = reference_dop_file(path, kind, tmp_path)
    out = np.zeros((int(ref["frames"]) + 8, nch), dtype=np.int32)
    n = lib.dsdhost_dop_file(path.encode(), 0, out.ctypes.data, len(out))
    assert n == int(ref["frames"]), lib.dsdhost_last_error()
    assert np.array_equal(out[:n], ref["pcm"])


def test_dop_app_loop_matches_oracle_frame_count(oracle):
    lib = cpu_host_lib()
    for nbytes, length in ((5000, 40000), (4096, 32768), (4099, 32790), (300, 2400), (3, 24), (1, 8)):
        data = seeded_bytes(2, nbytes, nbytes)
        want = oracle.dop_run_app(data, length, 1)
        out = np.zeros((len(want) + 8, 2), dtype=np.int32)
        n = lib.dsdhost_dop_run_app(data.ctypes.data, 2, nbytes, length, 0x69, 1, out.ctypes.data, len(out))
        assert n == len(want) and np.array_equal(out[:n], want), (nbytes, length)


@pytest.mark.gpu
@pytest.mark.parametrize("nch", [1, 2, 6])
def test_gpu_dop_kernel_all_layouts(oracle, nch):
    import torch
    from tests.test_gpu_parity import to_blocks
    nbytes = 70_001
    data = seeded_bytes(nch, nbytes, 3 + nch)
    g = capi.DsdGpu(np.zeros((1, 256)), nch, 1)
    g.set_option("chunk_frames", 7777)
    for msb, pre, nframes in ((1, 0, 36000), (0, 1, 35100), (1, 7, 40)):
        want = oracle.dop_run_raw(data, 1 << 40, msb, pre, nframes)
        for block, src in ((0, data), (4096, to_blocks(data, 4096)), (1, np.ascontiguousarray(data.T).reshape(-1))):
            g.set_option("source_block_bytes", block)
            got = g.dop_pack_host(src if block == 0 else src.reshape(1, -1), pre - 1, nframes, msb, in_count=nbytes)
            assert np.array_equal(got, want), (msb, pre, block)
            d_in = torch.from_numpy(np.ascontiguousarray(src)).cuda()
            d_out = torch.zeros((nframes, nch), dtype=torch.int32, device="cuda")
            g.dop_pack_device(d_in, nbytes if block == 0 else 0, 0, nbytes, 0x69, pre - 1, nframes, msb, d_out)
            assert np.array_equal(d_out.cpu().numpy(), want), (msb, pre, block, "device")
    g.set_option("source_block_bytes", 0)
    g.close()


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref")
@pytest.mark.parametrize("tag,kind,nbytes,nch,kw", CASES[:3] + CASES[5:8], ids=[c[0] for c in CASES[:3] + CASES[5:8]])
def test_gpu_dop_file_equals_reference(tmp_path, tag, kind, nbytes, nch, kw):
    host = capi.host_lib()
    path, data = make_file(tmp_path, tag, kind, nbytes, nch, kw)
    ref = reference_dop_file(path, kind, tmp_path)
    out = np.zeros((int(ref["frames"]) + 8, nch), dtype=np.int32)
    n = host.dsdhost_dop_file(path.encode(), 0, out.ctypes.data, len(out))
    assert n == int(ref["frames"]), host.dsdhost_last_error()
    assert np.array_equal(out[:n], ref["pcm"])
