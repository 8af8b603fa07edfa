#!/usr/bin/env python
"""Host->device copy rate of this box for the benchmark's capture sizes: one 1.2 GB copy, 32 MB chunks on one
stream, chunks alternating over two streams (pinned memory, CUDA events).  The end-to-end step cannot be
faster than these."""
import torch
n = 1_199_999_988
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def whole(): dev.copy_(host, non_blocking=True)
def chunks(c):
    def f():
        for o in range(0, n, c): dev[o:o + c].copy_(host[o:o + c], non_blocking=True)
    return f
s2 = [torch.cuda.Stream(), torch.cuda.Stream()]
def two_streams(c):
    def f():
        cur = torch.cuda.current_stream()
        for s in s2: s.wait_stream(cur)
        for k, o in enumerate(range(0, n, c)):
            with torch.cuda.stream(s2[k & 1]): dev[o:o + c].copy_(host[o:o + c], non_blocking=True)
        for s in s2: cur.wait_stream(s)
    return f
for name, fn in (("one copy of 1.2 GB", whole), ("32 MB chunks, one stream", chunks(32 << 20)), ("8 MB chunks, one stream", chunks(8 << 20)),
                 ("128 MB chunks, one stream", chunks(128 << 20)), ("32 MB chunks, two streams", two_streams(32 << 20))):
    ms = timed(fn)
    print(f"{name:32s} {ms:7.3f} ms  {n / ms / 1e6:6.2f} GB/s")
