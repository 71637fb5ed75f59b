"""What do the copy engines reach over NVLink between two B200s?  One-way and simultaneous two-way peer copies
(cudaMemcpyPeerAsync through torch, each direction issued on a stream of its SOURCE device = push), sizes 64 MB .. 2 GB."""
import time

import torch

assert torch.cuda.device_count() >= 2
for i, j in ((0, 1), (1, 0)):
    with torch.cuda.device(i):
        pass
n = 2 << 30
a0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
b0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
a1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
b1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
s0 = torch.cuda.Stream(device=0)
s1 = torch.cuda.Stream(device=1)


def run(size, both, chunk=None, iters=5):
    chunk = chunk or size
    best = 1e9
    for _ in range(iters):
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        for off in range(0, size, chunk):
            with torch.cuda.stream(s0):
                a1[off:off + chunk].copy_(a0[off:off + chunk], non_blocking=True)        # 0 -> 1, pushed by device 0
            if both:
                with torch.cuda.stream(s1):
                    b0[off:off + chunk].copy_(b1[off:off + chunk], non_blocking=True)    # 1 -> 0, pushed by device 1
        s0.synchronize(); s1.synchronize()
        best = min(best, time.perf_counter() - t0)
    return size / best / 1e9


for size in (64 << 20, 512 << 20, 2 << 30):
    print("size %5d MB: one-way %.0f GB/s, two-way %.0f GB/s each way, two-way in 64 MB chunks %.0f GB/s each way" % (
        size >> 20, run(size, False), run(size, True), run(size, True, 64 << 20)), flush=True)
