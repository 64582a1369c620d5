"""Measure the pure-WRITE HBM bandwidth of this GPU (cudaMemset via torch fill_, and a
uint8 zero fill) for a 3.7 GB buffer: the ceiling a write-only kernel can reach, next to the
read+write copy figure of MEASURED_PEAKS.json.  Run on the GPU box: python tools/write_bw_probe.py"""
import json

import torch

n = 65536 * 56448
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
src = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, fn in [("fill_u8", lambda: buf.fill_(7)), ("zero_", lambda: buf.zero_()),
                 ("fill_i32", lambda: buf.view(torch.int32).fill_(7)), ("copy_rw", lambda: buf.copy_(src))]:
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    bytes_moved = n * (2 if name == "copy_rw" else 1)
    res[name] = {"ms": best, "GBps": bytes_moved / best / 1e6}
print(json.dumps(res))
