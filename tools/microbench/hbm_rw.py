"""What HBM sustains on this GPU for the three traffic mixes the path's HBM-bound kernels have:
pure write (dither generator), 1:2 read:write (DoP packing), 1:1 copy (MEASURED_PEAKS' figure).
Plain torch ops, CUDA events, buffers of 2 GiB (>> L2).  Prints one JSON line."""
import json
import torch

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

n = 1 << 31
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
a32, b32 = a.view(torch.int32), b.view(torch.int32)
a16 = a.view(torch.int16)[: n // 4]                        # n/2 bytes read ...
res = {}
res["write_only_memset_gbs"] = n / timed(lambda: a.zero_()) / 1e9
res["write_only_fill_i32_gbs"] = n / timed(lambda: a32.fill_(7)) / 1e9
res["read_only_sum_gbs"] = n / timed(lambda: a32.sum()) / 1e9
res["copy_1r1w_gbs"] = 2 * n / timed(lambda: b.copy_(a)) / 1e9
# 1:2 read:write: int16 -> int32 widening copy (n/2 bytes in, n bytes out)
res["widen_1r2w_gbs"] = 1.5 * n / timed(lambda: b32[: n // 4].copy_(a16)) / 1e9 / 1.0 * (n // 4 * 6) / (1.5 * n)
print(json.dumps(res))
