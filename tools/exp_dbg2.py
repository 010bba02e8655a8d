"""Not a test: per-band phase timing of sweep 0 (option dbg=2)."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
ctx = mgm_b200.Context(0)
W, H, L, K = 2048, 1536, 256, 3
VS = ctx.padded_labels(L)
cc = torch.rand((H, W, VS), device="cuda") * 60
torch.cuda.synchronize()
for felz, mask in ((0, 0x01), (0, 0xFF), (1, 0x01), (1, 0xFF)):
    ctx.set_option("reset"); ctx.set_option("dbg", 2)
    print("felz", felz, "mask %02x" % mask, flush=True)
    for i in range(2):
        if i == 1:
            sys.stderr.flush()
        ctx.set_option("dbg", 2 if i == 1 else 0)
        ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 2.0 if felz else 8.0, 20000.0 if felz else 32.0, 8, K, felz, mask)
        ctx.synchronize()
