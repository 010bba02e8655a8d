"""Not a test: pure-write and copy bandwidth of the device (torch fill_/copy_), for the cost-volume roofline."""
import torch
n = 2048 * 1536 * 256
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(1.0)); print("fill_ 3.2 GB: %.3f ms = %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: a.zero_()); print("zero_ 3.2 GB: %.3f ms = %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy_ 3.2 GB: %.3f ms = %.0f GB/s read+written" % (ms, 2 * n * 4 / ms / 1e6))
ms = t(lambda: a.sum()); print("sum   3.2 GB: %.3f ms = %.0f GB/s read" % (ms, n * 4 / ms / 1e6))
