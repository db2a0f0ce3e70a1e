"""Probe: host->device copy ceilings on this box (pinned memory), to put the end-to-end number in context."""
import time
import torch

n = 512 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timeit(lambda: d.copy_(h, non_blocking=True))
print("one 512 MB copy        : %.1f GB/s" % (n / t / 1e9))
chunk = 2 * 1024 * 1024
def chunks():
    for o in range(0, n, chunk):
        d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
t = timeit(chunks)
print("256 x 2 MB, one stream : %.1f GB/s" % (n / t / 1e9))
streams = [torch.cuda.Stream() for _ in range(4)]
def chunks4():
    for i, o in enumerate(range(0, n, chunk)):
        with torch.cuda.stream(streams[i % 4]):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
t = timeit(chunks4)
print("256 x 2 MB, 4 streams  : %.1f GB/s" % (n / t / 1e9))
chunk = 8 * 1024 * 1024
t = timeit(chunks)
print("64 x 8 MB, one stream  : %.1f GB/s" % (n / t / 1e9))
h2 = torch.empty(64 * 1024 * 1024, dtype=torch.uint8).pin_memory()
d2 = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device="cuda")
t = timeit(lambda: h2.copy_(d2, non_blocking=True))
print("one 64 MB D2H copy     : %.1f GB/s" % (h2.numel() / t / 1e9))
