"""H2D bandwidth of the e2e input copy (pinned host -> device): one stream vs the copy split over several streams."""
import json
import torch
dev = torch.device("cuda:0")
n = 268435456 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    sz = n // parts
    def run():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * sz:(i + 1) * sz].copy_(h[i * sz:(i + 1) * sz], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream(dev).wait_stream(s)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"streams": parts, "ms": ms, "GB/s": n * 4 / ms / 1e6}))
