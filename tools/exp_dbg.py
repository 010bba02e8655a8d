"""Not a test: phase timing of one axis sweep (register-chain band steps), option dbg."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
ctx = mgm_b200.Context(0)
W, H, L, K = 2048, 1536, 256, 3
VS = ctx.padded_labels(L)
cc = torch.rand((H, W, VS), device="cuda") * 60
torch.cuda.synchronize()
for opts in ({}, {"reg_chains": 1}, {"rows_axis": 28}, {"sgm": 1}, {"sgm": 1, "cc_pf": 0}, {"sgm": 1, "rows_axis": 28}):
    ctx.set_option("reset"); ctx.set_option("dbg", 1)
    sgm = opts.pop("sgm", 0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    for mask in (0x01, 0x0F):
        print("opts", opts, "sgm" if sgm else "trunc", "mask %02x" % mask, flush=True)
        for i in range(2):
            if sgm:
                ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 8.0, 32.0, 8, K, 0, mask)
            else:
                ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 2.0, 20000.0, 8, K, 1, mask)
            ctx.synchronize()
        sys.stderr.flush()
