"""Not a test: cost volume kernel timings (census 3x3 on the headline shape, AD 3 channels)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from mgm_b200 import Context
ctx = Context(0)
for (W, H, L, nch, dist, win) in [(2048, 1536, 256, 1, "census", 3), (1920, 1080, 128, 1, "census", 5), (2048, 1536, 256, 3, "ad", 3), (4096, 4096, 64, 1, "ncc", 5), (1242, 375, 192, 1, "ad", 3)]:
    u, v = bench.synth_pair(W, H, L, 0)
    u = np.repeat(np.asarray(u, np.float32).reshape(-1, H, W)[:1], nch, 0)
    v = np.repeat(np.asarray(v, np.float32).reshape(-1, H, W)[:1], nch, 0)
    du, dv = torch.from_numpy(np.ascontiguousarray(u)).cuda(), torch.from_numpy(np.ascontiguousarray(v)).cuda()
    VS = ctx.padded_labels(L)
    dcc = torch.empty((H, W, VS), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    def run():
        ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, nch, -(L - 1), 0, "none", dist, float("inf"), win, dcc.data_ptr())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(10): run()
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%dx%dx%d %s nch=%d win=%d: %.3f ms (%.0f GB/s written)" % (W, H, L, dist, nch, win, ms, W * H * VS * 4 / ms / 1e6))
