"""Not a test: micro-timings of the aggregation kernel (single band / single sweep) on the GPU box."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import mgm_b200

ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)

def run(W, H, L, K, felz, mask, NDIR=8, reps=3, rows=0):
    VS = ctx.padded_labels(L)
    ctx.set_rows_per_band(rows)
    cc = torch.rand((H, W, VS), device="cuda") * 60
    cc[..., L:] = float("inf")
    torch.cuda.synchronize()
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 2.0 if felz else 8.0, 20000.0 if felz else 32.0, NDIR, K, felz, mask)
            e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    info = ctx.last_launch_info()
    ms = min(ts[1:])
    return ms, info

if __name__ != '__main__':
    CASES = []
else:
  CASES = [
    (2048, 43, 256, 3, 1, 0x01, "1 band sweep0 trunc K3"),
    (2048, 35, 256, 3, 1, 0x10, "1 band sweep4 trunc K3"),
    (2048, 43, 256, 3, 0, 0x01, "1 band sweep0 sgm K3"),
    (2048, 43, 256, 2, 0, 0x01, "1 band sweep0 sgm K2"),
    (2048, 1536, 256, 3, 1, 0x01, "full sweep0 trunc K3"),
    (2048, 1536, 256, 3, 1, 0x10, "full sweep4 trunc K3"),
    (2048, 1536, 256, 3, 1, 0x0F, "full sweeps0-3 trunc K3"),
    (2048, 1536, 256, 3, 1, 0xFF, "full all trunc K3"),
    (2048, 1536, 256, 3, 0, 0xFF, "full all sgm K3"),
    (1920, 1080, 128, 2, 0, 0xFF, "cfg2 all sgm K2"),
    (4096, 4096, 64, 2, 0, 0xFF, "big 4096x4096x64 sgm K2"),
    (4096, 4096, 64, 4, 0, 0xFF, "big 4096x4096x64 sgm K4"),
]
if __name__ == '__main__' and len(sys.argv) > 1:
    CASES = [c for c in CASES if sys.argv[1] in c[6]]
for (W, H, L, K, felz, mask, label) in CASES:
    ms, info = run(W, H, L, K, felz, mask, reps=int(sys.argv[2]) if len(sys.argv) > 2 else 3)
    g_steps = None
    print("%-28s %8.3f ms  rows=%d/%d thr=%d smem=%d" % (label, ms, info["rows_axis"], info["rows_diag"], info["threads_per_cta"], info["smem_bytes"]), flush=True)
    if "1 band" in label:
        sig = 2 if (mask & 0xF0) else 1
        steps = W + sig * (H - 1)
        print("     -> %.2f us/step (%d steps)" % (ms * 1e3 / steps, steps))
