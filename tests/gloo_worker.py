"""Worker of tests/test_sharding.py::test_two_ranks_gloo (launched by torchrun, backend gloo).
Each rank runs only its own shards; rank 0 gathers them, stitches and compares with the oracle."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsf2flac_b200 import batch, filters  # noqa: E402
from tests import standin  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402
from tests.test_sharding import mixed_jobs  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    jobs = mixed_jobs()
    mine = batch.BatchRunner(make_ctx=standin.make_ctx).run(jobs, rank, world, align_frames=64)
    dist.barrier()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # test-only: the product keeps PCM per rank
    if rank == 0:
        assert all(len(g) > 0 for g in gathered), "every rank must own work"
        outs = batch.stitch(jobs, [x for g in gathered for x in g])
        oracle = Oracle()
        for j, got in zip(jobs, outs):
            scale, amp, clip = filters.app_params(j.bits, j.scale_db, j.dither)
            want = oracle.run_app(j.planar, j.length_samples, j.msb_flag, j.dsd_fs, j.out_fs, scale, amp, clip)
            assert np.array_equal(got, want)
        print(f"GLOO_OK world={world}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
