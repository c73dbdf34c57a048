"""Multi-GPU partitioning on CPU: the shard plan tiles any batch exactly
once for every world size, balances cost, and a 2-rank gloo run (each rank computing only its
own shards, through the CPU stand-in of the C ABI) stitches to exactly what the oracle produces
for whole files -- file, channel and time-segment cuts with their halos included."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from dsf2flac_b200 import batch, capi, filters
from tests import standin
from tests.oracle_lib import DSD64, seeded_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mixed_jobs(seed=0):
    rng = np.random.default_rng(seed)
    jobs = []
    for k in range(5):
        fs, out = [(DSD64, 88200), (2 * DSD64, 176400), (DSD64, 176400), (DSD64, 352800)][k % 4]
        nch = [2, 1, 6, 2, 3][k]
        nbytes = int(rng.integers(3000, 9000))
        jobs.append(batch.Job(seeded_bytes(nch, nbytes, 100 + k), fs, out, msb_flag=bool(k % 2), bits=[24, 16, 20][k % 3],
                              dither=(k != 3), length_samples=nbytes * 8 - int(rng.integers(0, 40))))
    return jobs


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plan_tiles_exactly_once_and_balances(world):
    frames, nch, cost = [100000, 7, 250000, 1024, 0, 99999], [2, 6, 1, 2, 2, 3], [72, 30, 12, 72, 72, 72]
    cover = [np.zeros((f, c), dtype=np.int32) for f, c in zip(frames, nch)]
    loads = []
    for r in range(world):
        load = 0
        for job, c0, k, f0, nf in capi.plan_shards(frames, nch, cost, r, world, align_frames=256):
            assert nf > 0 and k > 0
            cover[job][f0:f0 + nf, c0:c0 + k] += 1
            load += nf * k * cost[job]
        loads.append(load)
    assert all((c == 1).all() for c in cover)
    total = sum(f * c * w for f, c, w in zip(frames, nch, cost))
    assert sum(loads) == total
    assert max(loads) - min(loads) <= 2 * 256 * 72 * 6 + 1


def test_plan_shapes():
    # 6 channels over 2 ranks -> 3 + 3 whole channels (config 3); stereo over 8 -> quarter channels (config 4)
    assert capi.plan_shards([1 << 20], [6], [72], 0, 2) == [(0, 0, 3, 0, 1 << 20)]
    assert capi.plan_shards([1 << 20], [6], [72], 1, 2) == [(0, 3, 3, 0, 1 << 20)]
    assert capi.plan_shards([1 << 20], [2], [72], 5, 8) == [(0, 1, 1, 1 << 18, 1 << 18)]
    assert capi.plan_shards([], [], [], 0, 2) == []
    with pytest.raises(capi.DsdGpuError):
        capi.plan_shards([10], [2], [72], 2, 2)


@pytest.mark.parametrize("world", [1, 3])
def test_single_process_stitch_equals_oracle(oracle, world):
    jobs = mixed_jobs()
    runner = batch.BatchRunner(make_ctx=standin.make_ctx)
    results = []
    for r in range(world):
        results += runner.run(jobs, r, world, align_frames=64)
    for j, got in zip(jobs, batch.stitch(jobs, results)):
        from dsf2flac_b200 import filters
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        want = oracle.run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("world,ndev", [(1, 1), (1, 3), (2, 2), (5, 1)])
def test_c_batch_runner_equals_oracle(oracle, world, ndev):
    """dsdgpu_run_batch (the C entry a C++ caller uses: one call, one host thread per device, chunks pipelined across
    files) over the CPU stand-in: every (world, devices per process) split fills each job's interleaved output exactly
    once -- whole files, time segments and channel-subset shards -- and the sum equals the oracle bit for bit."""
    jobs = mixed_jobs(3)
    want = []
    for j in jobs:
        scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
        want.append(oracle.run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip))
    total = [np.zeros_like(w) for w in want]
    for r in range(world):
        outs = batch.run_batch_c(jobs, devices=tuple(range(ndev)), rank=r, world=world, align_frames=64, lib=standin.lib())
        for t, o in zip(total, outs):
            assert not ((t != 0) & (o != 0)).any()            # nobody writes a sample somebody else wrote
            t += o
    for t, w in zip(total, want):
        assert np.array_equal(t, w)


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "gloo_worker.py")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK world=2" in r.stdout
