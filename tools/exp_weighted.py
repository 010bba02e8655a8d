"""Not a test: the generic kernel on per-edge weights (update_costW with image weights) beside the lean unweighted kernels."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
def t(fn, reps=3):
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream); fn(); e1.record(stream)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
for (W, H, L, K, felz) in [(1920, 1080, 128, 2, 0), (1920, 1080, 128, 4, 0), (2048, 1536, 256, 3, 1)]:
    VS = ctx.padded_labels(L)
    cc = torch.rand((H, W, VS), device="cuda") * 60; cc[..., L:] = float("inf")
    w = torch.rand((8, H, W), device="cuda") * 0.75 + 0.25
    out = torch.empty((H, W), device="cuda"); cost = torch.empty((H, W), device="cuda")
    P1, P2 = (2.0, 20000.0) if felz else (8.0, 32.0)
    a = t(lambda: ctx.aggregate_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, P1, P2, 8, K, felz, 1, "vfit", out.data_ptr(), cost.data_ptr()))
    b = t(lambda: ctx.aggregate_dev(cc.data_ptr(), w.data_ptr(), 1, W, H, -(L - 1), 0, P1, P2, 8, K, felz, 1, "vfit", out.data_ptr(), cost.data_ptr()))
    print("%dx%dx%d K%d felz=%d: unweighted (lean) %.3f ms | per-edge weights (generic) %.3f ms" % (W, H, L, K, felz, a, b), flush=True)
    del cc, w; torch.cuda.empty_cache()
