import sys
sys.path.insert(0, ".")
import numpy as np, torch
import mgm_b200
from bench import synth_pair
ctx = mgm_b200.Context(0)
W, H, L, NDIR, K, felz, P1, P2 = 640, 480, 256, 8, 3, 1, 2.0, 20000.0
dmin = -(L - 1)
VS = ctx.padded_labels(L)
u, v = synth_pair(W, H, L, seed=5)
du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
dcc = torch.empty((H, W, VS), device="cuda")
torch.cuda.synchronize()
ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, dmin, 0, "census", "census", float("inf"), 3, dcc.data_ptr())
o1 = torch.empty((H, W), device="cuda"); c1 = torch.empty((H, W), device="cuda"); S1 = torch.empty((H, W, L), device="cuda")
ctx.aggregate_dev(dcc.data_ptr(), 0, 0, W, H, dmin, 0, P1, P2, NDIR, K, felz, 1, "vfit", o1.data_ptr(), c1.data_ptr(), S1.data_ptr())
ctx.synchronize()
ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, W, H, dmin, 0, P1, P2, NDIR, K, felz, 0xFF)
pa = torch.empty((H, W, VS), device="cuda"); pb = torch.empty((H, W, VS), device="cuda")
ctx.sum_sweeps_dev(W, H, dmin, 0, 0x55, pa.data_ptr()); ctx.sum_sweeps_dev(W, H, dmin, 0, 0xAA, pb.data_ptr())
ctx.synchronize()
tot = pa + pb
torch.cuda.synchronize()
o2 = torch.empty((H, W), device="cuda"); c2 = torch.empty((H, W), device="cuda")
ctx.finish_sum_dev(tot.data_ptr(), dcc.data_ptr(), W, H, dmin, 0, NDIR, 1, "vfit", 0, H, o2.data_ptr(), c2.data_ptr())
ctx.synchronize()
fin = torch.isfinite(c1) & torch.isfinite(c2)
d = torch.where(fin, (c1 - c2).abs(), torch.zeros_like(c1))
i = int(d.argmax().item()); y, x = i // W, i % W
print("worst", y, x, float(d.max()), "c1", float(c1[y, x]), "c2", float(c2[y, x]), "o1", float(o1[y, x]), "o2", float(o2[y, x]))
o = int(np.floor(float(o1[y, x]) + 0.5)) - dmin
cc = dcc[y, x, :L].cpu().numpy()
S2 = (tot[y, x, :L] - 7 * dcc[y, x, :L]).cpu().numpy()
print("label", o, "S1", S1[y, x, max(0, o - 3):o + 4].cpu().numpy(), "S2", S2[max(0, o - 3):o + 4])
print("nonfinite mismatch", int((torch.isfinite(c1) != torch.isfinite(c2)).sum()))
