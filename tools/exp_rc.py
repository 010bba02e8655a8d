"""Not a test: single-sweep latency of the register-chain vs lane-pair kernels vs rows per band (depth-bound regime)."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
W, H, L, K = 2048, 1536, 256, 3
VS = ctx.padded_labels(L)
cc = torch.rand((H, W, VS), device="cuda") * 60
torch.cuda.synchronize()
def t(mask, opts):
    ctx.set_option("reset")
    for k, v in opts.items():
        ctx.set_option(k, v)
    ts = []
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 2.0, 20000.0, 8, K, 1, mask)
            e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    info = ctx.last_launch_info()
    print("mask %02x %-44s %8.3f ms rows=%d/%d thr=%d" % (mask, str(opts), min(ts[1:]), info["rows_axis"], info["rows_diag"], info["threads_per_cta"]), flush=True)
for mask in (0x01, 0x10, 0x11, 0x55, 0xFF):
    for rc in (0, 1):
        for rows in (0, 40, 24, 16, 12, 8, 4):
            o = {"reg_chains": rc}
            if rows:
                o.update(rows_axis=rows, rows_diag=rows)
            t(mask, o)
