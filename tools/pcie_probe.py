"""PCIe probe: pinned-host <-> device copy bandwidth for the e2e path's transfer shapes (run on the GPU box)."""
import time
import torch

dev = torch.device("cuda", 0)
n, K = 1 << 20, 64


def bw(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


h = torch.empty(K * n * 13, dtype=torch.uint8, pin_memory=True)
d = torch.empty(K * n * 13, dtype=torch.uint8, device=dev)
print("D2H contiguous 872 MB: %.1f GB/s" % bw(lambda: h.copy_(d, non_blocking=True), h.numel()))
ha = torch.empty(K * n * 8, dtype=torch.uint8, pin_memory=True)
da = torch.empty(K * n * 8, dtype=torch.uint8, device=dev)
print("H2D contiguous 537 MB: %.1f GB/s" % bw(lambda: da.copy_(ha, non_blocking=True), ha.numel()))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        da.copy_(ha, non_blocking=True)


t = bw(both, h.numel() + ha.numel())
print("duplex (D2H 872 MB || H2D 537 MB): %.1f GB/s total -> %.2f ms per pair" % (t, (h.numel() + ha.numel()) / t / 1e6))
# strided 2-D copies like the env-chunked pipeline: 64 rows of 512 KB out of 4 MB-pitch rows
h2 = torch.empty(K, n, dtype=torch.float32, pin_memory=True)
d2 = torch.empty(K, 1 << 17, dtype=torch.float32, device=dev)
print("D2H 2-D (64 x 512 KB rows, host pitch 4 MB): %.1f GB/s" % bw(lambda: h2[:, : 1 << 17].copy_(d2, non_blocking=True), d2.numel() * 4))
