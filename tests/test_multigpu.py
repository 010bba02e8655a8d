"""GPU suite, multi-GPU part (-m gpu; skipped with fewer than two devices): the sweep-sharded layout of north_star on
real hardware -- one process per GPU, NCCL for the barrier / all-gather / all-reduce, CUDA IPC peer mappings for the
message volumes.  The ORDERED exchange (peer stores into the row-slab owner's volumes by the aggregation kernel, local
sweep-ordered finish) must reproduce the single-GPU maps bit for bit on every rank; the ALL-REDUCE exchange (the one
north_star names) may differ by rounding: its label flips and cost differences are counted and bounded.
Run by hand with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu -q` (log in profiles/)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import mgm_b200
    from mgm_b200 import sharding
    from bench import synth_pair
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = mgm_b200.Context(rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    W, H, L, NDIR, K, felz, P1, P2, dist_name, win = cfg
    dmin, dmax = -(L - 1), 0
    VS = ctx.padded_labels(L)
    u, v = synth_pair(W, H, L, seed=5)
    with torch.cuda.stream(stream):
        du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
        dcc = torch.empty((H, W, VS), device="cuda")
        ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, dmin, dmax, "census" if dist_name == "census" else "none",
                           dist_name, float("inf"), win, dcc.data_ptr())
        one_out = torch.empty((H, W), device="cuda"); one_cost = torch.empty((H, W), device="cuda")
        ctx.aggregate_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, P1, P2, NDIR, K, felz, 1, "vfit", one_out.data_ptr(),
                          one_cost.data_ptr())
    ctx.synchronize()
    sw = sharding.SweepSharded(ctx, dist, torch, stream, W, H, dmin, dmax, P1, P2, NDIR, K, felz, "vfit", world, rank)
    sw.setup(dcc)
    res = {}
    neq = lambda a, b: int((~((a == b) | (torch.isnan(a) & torch.isnan(b)))).sum().item())
    for ex in ("ordered", "allreduce", "ordered"):
        out = torch.full((H, W), -5.0, device="cuda"); cost = torch.full((H, W), -5.0, device="cuda")
        torch.cuda.synchronize()
        sw.step(dcc, out, cost, exchange=ex)
        stream.synchronize()
        torch.cuda.synchronize()
        res[ex] = (neq(out, one_out), neq(cost, one_cost))
    # the all-reduce exchange sums the sweeps in another order: WTA labels may flip at near-ties.  Counted on
    # the unrefined maps; the costs of the pixels that keep their label agree to rounding
    with torch.cuda.stream(stream):
        lab1 = torch.empty((H, W), device="cuda"); c1 = torch.empty((H, W), device="cuda")
        ctx.aggregate_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, P1, P2, NDIR, K, felz, 1, "none", lab1.data_ptr(), c1.data_ptr())
    ctx.synchronize()
    sw.refinement = "none"
    lab2 = torch.full((H, W), -5.0, device="cuda"); c2 = torch.full((H, W), -5.0, device="cuda")
    torch.cuda.synchronize()
    sw.step(dcc, lab2, c2, exchange="allreduce")
    stream.synchronize()
    torch.cuda.synchronize()
    flipped = ~((lab1 == lab2) | (torch.isnan(lab1) & torch.isnan(lab2)))
    keep = ~flipped & torch.isfinite(c1)
    rel = float(((c2[keep] - c1[keep]).abs() / c1[keep].abs().clamp(min=1.0)).max().item()) if bool(keep.any()) else 0.0
    res["allreduce_labels"] = (int(flipped.sum().item()), rel)
    dist.barrier()
    sw.close()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [(640, 480, 256, 8, 3, 1, 2.0, 20000.0, "census", 3),     # headline kernel
                                 (500, 375, 64, 16, 2, 0, 8.0, 32.0, "ad", 3)])           # 16 sweeps, 2 per GPU at N=8
def test_sweep_sharded_matches_one_gpu(cfg):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    world = 1 << (world.bit_length() - 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(120)
    npix = cfg[0] * cfg[1]
    for rank, res in got:
        assert res["ordered"] == (0, 0), (rank, res)               # bit-identical to one GPU, on every rank, twice
        flips, rel = res["allreduce_labels"]
        assert flips <= npix // 1000 and rel <= 1e-4, (rank, res)  # within north_star's tolerance, a handful of near-tie flips
    print("sweep-sharded vs 1 GPU, world=%d: %s" % (world, got[0][1]))
