"""Pinned host -> device bandwidth with one and two concurrent streams, and with 128 MB chunks (what the streaming
ingest issues).  Measurement aid: the floor of the e2e path."""
import time, torch
dev = torch.device("cuda", 0)
n = 1 << 31
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9
print("one copy, 2 GiB: %.1f GB/s" % t(lambda: d.copy_(h, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): d[n // 2:].copy_(h[n // 2:], non_blocking=True)
print("two streams, halves: %.1f GB/s" % t(two))
ch = 128 << 20
def chunks():
    for o in range(0, n, ch): d[o:o + ch].copy_(h[o:o + ch], non_blocking=True)
print("128 MiB chunks, one stream: %.1f GB/s" % t(chunks))
def chunks2():
    for i, o in enumerate(range(0, n, ch)):
        with torch.cuda.stream(s1 if i & 1 else s2): d[o:o + ch].copy_(h[o:o + ch], non_blocking=True)
print("128 MiB chunks, alternating two streams: %.1f GB/s" % t(chunks2))
