"""Host-side logic of the multi-GPU direction sharding.

The NDIR sweeps of mgm() are independent until the sum S = sum_p L_p (mgm_core.cc:582-587) -- the reference's
own mgm_naive_parallelism (mgm_core.cc:632) exploits exactly that.  Here sweep p is aggregated on rank
p mod world; every rank then finishes a slab of image rows: it reads all NDIR sweep volumes for its slab (its
own from local HBM, the others' through CUDA IPC mappings over NVLink), adds them IN SWEEP ORDER inside the
fused WTA kernel (so the result is bit-identical to one GPU whatever the rank count), and the two small
output maps are all-gathered.  The only bytes that cross NVLink are (world-1)/world of the NDIR volumes.
"""


def sweep_owner(p, world):
    return p % world


def sweep_mask(ndir, world, rank):
    """bit p set = this rank aggregates sweep p"""
    m = 0
    for p in range(ndir):
        if sweep_owner(p, world) == rank:
            m |= 1 << p
    return m


def row_slabs(ny, world):
    """[r0, r1) image-row slab finished by each rank; contiguous, covering, balanced to one row"""
    b = [(ny * r) // world for r in range(world + 1)]
    return [(b[r], b[r + 1]) for r in range(world)]


def exchange_sweep_handles(ctx, dist, ndir, world, rank):
    """Every rank exports CUDA IPC handles of the sweep volumes it owns; returns device pointers (local or
    peer-mapped) of all ndir sweeps, in sweep order."""
    handles = [None] * ndir
    for p in range(ndir):
        if sweep_owner(p, world) == rank:
            ptr, nbytes = ctx.sweep_volume(p)
            if not ptr:
                raise RuntimeError("sweep %d has not been aggregated on rank %d yet" % (p, rank))
            handles[p] = ctx.ipc_export(ptr)
    gathered = [None] * world
    dist.all_gather_object(gathered, handles)
    ptrs = []
    for p in range(ndir):
        o = sweep_owner(p, world)
        ptrs.append(ctx.sweep_volume(p)[0] if o == rank else ctx.ipc_open(gathered[o][p]))
    return ptrs
