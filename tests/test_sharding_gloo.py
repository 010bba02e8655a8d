"""CPU suite, part 4: the N>1 host logic under a real process group (gloo, world_size 2 and 3).
Each rank "aggregates" its sweeps with the oracle port, the per-sweep volumes are exchanged, every rank sums
its row slab in sweep order and the maps are all-gathered -- the same protocol bench.py / the GPU path run
with NCCL + CUDA IPC.  The assembled result must equal the single-process oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgm_b200 import sharding


def test_partitions():
    for ndir in [1, 2, 4, 8, 16]:
        for world in [1, 2, 3, 4, 8]:
            masks = [sharding.sweep_mask(ndir, world, r) for r in range(world)]
            assert sum(masks) == (1 << ndir) - 1 and all(a & b == 0 for i, a in enumerate(masks) for b in masks[i + 1:])
    for ny in [1, 7, 375, 1536, 4096]:
        for world in [1, 2, 3, 8]:
            s = sharding.row_slabs(ny, world)
            rps = sharding.slab_rows(ny, world)
            assert s[0][0] == 0 and s[-1][1] == ny and all(s[i][1] == s[i + 1][0] for i in range(world - 1))
            # uniform slabs: the aggregation kernel finds the slab of an image row by one division
            assert all((a, b) == (min(ny, r * rps), min(ny, (r + 1) * rps)) for r, (a, b) in enumerate(s))
            assert rps * world >= ny and rps >= 2
            # the kernel's magic-number division (capi.cu: slab_magic) is exact for every row
            magic = ((1 << 32) + rps - 1) // rps
            assert all((y * magic) >> 32 == y // rps for y in range(ny))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as O
    from tests.conftest import synth_volume
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, L, NDIR, K = 31, 22, 9, 8, 3
    cc = synth_volume(nx, ny, L, seed=11, real=True)
    full = O.orc_mgm(cc, None, -(L - 1), 2, 20000, NDIR, K, 1, 1, want_passes=True)
    mask = sharding.sweep_mask(NDIR, world, rank)
    # "my" sweeps (computed locally), zeros elsewhere; exchange = all_gather of the owned volumes
    mine = torch.from_numpy(np.stack([full["passes"][p] if (mask >> p) & 1 else np.zeros_like(cc) for p in range(NDIR)]))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    sweeps = [gathered[sharding.sweep_owner(p, world)][p].numpy() for p in range(NDIR)]
    r0, r1 = sharding.row_slabs(ny, world)[rank]
    S = np.zeros((r1 - r0, nx, L), np.float32)
    for p in range(NDIR):                      # sweep order, like mgm_core.cc:582-587
        S += sweeps[p][r0:r1]
    S = S - np.float32(NDIR - 1) * cc[r0:r1]
    fin = np.isfinite(S)
    Sm = np.where(fin, S, np.inf)
    out = (np.argmin(Sm, axis=2) - (L - 1)).astype(np.float32)
    outs = [None] * world
    dist.all_gather_object(outs, out)
    if rank == 0:
        q.put((np.array_equal(np.concatenate(outs, 0), full["out"]), np.array_equal(S, full["S"][r0:r1], equal_nan=True)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert ok == (True, True)
