"""Not a test: run the fused step many times on the same cost volume and compare every result with the first one, bit for
bit (a rare ordering bug in the band hand-off would show as a run that differs)."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
ctx = mgm_b200.Context(0)
cases = [(2048, 1536, 256, 3, 1, 0, 60), (1920, 1080, 128, 2, 0, 0, 150), (640, 480, 128, 2, 0, 8, 400), (640, 480, 128, 3, 1, 8, 300),
         (1242, 375, 192, 4, 0, 0, 200), (900, 700, 64, 2, 0, 0, 300)]
for (W, H, L, K, felz, rows, reps) in cases:
    VS = ctx.padded_labels(L)
    g = torch.Generator(device="cuda"); g.manual_seed(W + L)
    cc = torch.rand((H, W, VS), device="cuda", generator=g) * 60
    cc[..., L:] = float("inf")
    out = torch.empty((H, W), device="cuda"); cost = torch.empty((H, W), device="cuda")
    P1, P2 = (2.0, 20000.0) if felz else (8.0, 32.0)
    ctx.set_option("reset"); ctx.set_rows_per_band(rows)
    ref = None; bad = 0
    for i in range(reps):
        out.fill_(-1); cost.fill_(-1)
        ctx.aggregate_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, P1, P2, 8, K, felz, 1, "vfit", out.data_ptr(), cost.data_ptr())
        ctx.synchronize()
        if ref is None:
            ref = (out.clone(), cost.clone())
        elif not (torch.equal(out, ref[0]) and torch.equal(cost, ref[1])):
            bad += 1
    print("%dx%dx%d K%d felz=%d rows=%d: %d runs, %d differ from the first" % (W, H, L, K, felz, rows, reps, bad), flush=True)
    del cc
    torch.cuda.empty_cache()
ctx.set_rows_per_band(0)
