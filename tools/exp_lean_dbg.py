"""Not a test: per-role cycle counters of the axis bands (option dbg), lean (MGM_LEAN_DBG build) vs generic SGM kernels."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
ctx = mgm_b200.Context(0)
cases = [(1920, 1080, 128, 2)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (W, H, L, K) in cases:
    VS = ctx.padded_labels(L)
    cc = torch.rand((H, W, VS), device="cuda") * 60
    cc[..., L:] = float("inf")
    for opts in ({}, {"no_lean_sgm": 1}):
        for mask in (0x0F, 0x01, 0xFF):
            ctx.set_option("reset")
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 8.0, 32.0, 8, K, 0, mask)
            ctx.synchronize()
            ctx.set_option("dbg", 1)
            print("%dx%dx%d K%d %s mask %02x" % (W, H, L, K, opts, mask), flush=True)
            sys.stderr.flush()
            ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 8.0, 32.0, 8, K, 0, mask)
            ctx.synchronize()
