#!/usr/bin/env python
"""HBM bandwidth at different read:write mixes (torch elementwise kernels, CUDA events, best of 10).
MEASURED_PEAKS.json's hbm_gbs is a 1:1 copy; gs_integral moves 1 byte in for 4 bytes out."""
import torch
def best(fn, nbytes, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); t.append(a.elapsed_time(b))
    return nbytes / (min(t) * 1e-3) / 1e9
n = 1 << 31                                   # 2 Gi elements
u8 = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
i32 = torch.empty((n,), dtype=torch.int32, device="cuda")
i32b = torch.empty((n // 2,), dtype=torch.int32, device="cuda"); i32c = torch.empty_like(i32b)
print("copy 1:1 (int32 -> int32)      %.0f GB/s" % best(lambda: i32c.copy_(i32b), 2 * 4 * (n // 2)))
print("write only (fill int32)        %.0f GB/s" % best(lambda: i32.fill_(7), 4 * n))
print("u8 -> int32 cast (1 in : 4 out) %.0f GB/s" % best(lambda: i32.copy_(u8), 5 * n))
print("read only (sum of int32)       %.0f GB/s" % best(lambda: i32.sum(), 4 * n))
print("int32 -> u8 cast (4 in : 1 out) %.0f GB/s" % best(lambda: u8.copy_(i32), 5 * n))
