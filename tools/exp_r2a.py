"""Not a test: round-2 experiment sweep A (one GPU).  Prints one line per measurement.
usage: python tools/exp_r2a.py <group>   groups: sgm, axis, chain"""
import os, sys
sys.path.insert(0, ".")
grp = sys.argv[1] if len(sys.argv) > 1 else "sgm"
import torch
import mgm_b200

ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
_cc = {}

def run(W, H, L, K, felz, mask, reps=2):
    VS = ctx.padded_labels(L)
    key = (W, H, VS)
    if key not in _cc:
        _cc.clear(); torch.cuda.empty_cache()
        cc = torch.rand((H, W, VS), device="cuda") * 60
        cc[..., L:] = float("inf")
        _cc[key] = cc
    cc = _cc[key]
    torch.cuda.synchronize()
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, 2.0 if felz else 8.0, 20000.0 if felz else 32.0, 8, K, felz, mask)
            e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), ctx.last_launch_info()

def line(tag, W, H, L, K, felz, mask, opts):
    ctx.set_option("reset")
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        ms, info = run(W, H, L, K, felz, mask)
        print("%-34s %dx%dx%d K%d %s mask=%02x %-40s %8.3f ms rows=%d/%d thr=%d smem=%d" % (
            tag, W, H, L, K, "trunc" if felz else "sgm", mask, str(opts), ms, info["rows_axis"], info["rows_diag"],
            info["threads_per_cta"], info["smem_bytes"]), flush=True)
    except Exception as e:
        print("%-34s %s FAILED %s" % (tag, opts, str(e)[:100]), flush=True)

if grp == "sgm":
    for (W, H, L, K) in [(1920, 1080, 128, 2), (1242, 375, 192, 4), (4096, 4096, 64, 2), (2048, 1536, 256, 3)]:
        line("sgm default", W, H, L, K, 0, 0xFF, {})
        for pf in (2, 4, 8):
            line("sgm cc_pf", W, H, L, K, 0, 0xFF, {"cc_pf": pf})
        for rows in (40, 28, 20, 14, 10):
            line("sgm rows", W, H, L, K, 0, 0xFF, {"rows_axis": rows, "rows_diag": rows})
            line("sgm rows+pf4", W, H, L, K, 0, 0xFF, {"rows_axis": rows, "rows_diag": rows, "cc_pf": 4})
        if L <= 128:
            line("sgm lanes4", W, H, L, K, 0, 0xFF, {"lanes4": 1})
            line("sgm lanes8", W, H, L, K, 0, 0xFF, {"lanes8": 1})
            line("sgm lanes4 pf4", W, H, L, K, 0, 0xFF, {"lanes4": 1, "cc_pf": 4})
elif grp == "axis":
    # one axis sweep / one diagonal sweep / 2 and 4 sweeps of the headline shape vs rows per band (sweep-sharded layouts)
    W, H, L, K = 2048, 1536, 256, 3
    for mask in (0x01, 0x04, 0x10, 0x11, 0x55):
        for rows in (56, 40, 32, 28, 24, 20, 16):
            line("trunc rows", W, H, L, K, 1, mask, {"rows_axis": rows, "rows_diag": rows})
elif grp == "chain":
    # whole headline workload vs row groups (run with MGMB200_LIBRARY=<chain prefetch variant>)
    W, H, L, K = 2048, 1536, 256, 3
    for g in (1, 2, 3):
        line("trunc groups lib=%s" % os.path.basename(os.environ.get("MGMB200_LIBRARY", "default")), W, H, L, K, 1, 0xFF, {"groups": g})
        line("trunc groups axis lib=%s" % os.path.basename(os.environ.get("MGMB200_LIBRARY", "default")), W, H, L, K, 1, 0x0F, {"groups": g})
