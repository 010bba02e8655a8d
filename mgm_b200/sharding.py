"""Host-side logic of the multi-GPU direction sharding.

The NDIR sweeps of mgm() are independent until the sum S = sum_p L_p (mgm_core.cc:582-587) -- the reference's
own mgm_naive_parallelism (mgm_core.cc:632) exploits exactly that.  Here sweep p is aggregated on rank
p mod world, and the image rows are cut into `world` slabs of `slab_rows` rows; slab r is finished by rank r.

ordered exchange (default, bit-exact): every rank holds all NDIR message volumes; the rank that aggregates sweep p
    stores the messages of slab r straight into rank r's volume p (CUDA IPC mappings: peer stores over NVLink issued
    by the aggregation kernel while the sweep runs).  After one barrier every rank finds all sweeps of its slab in
    local HBM and adds them IN SWEEP ORDER inside the WTA kernel -- the result is bit-identical to one GPU whatever
    the rank count -- and the two small maps are all-gathered.
all-reduce exchange (the one north_star names): every rank sums its own sweeps, the partial volumes are summed by
    one NCCL all-reduce, every rank finishes its slab of the sum.  The floating-point order differs from the
    reference's; bench.py counts the label / cost differences.
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


def slab_rows(ny, world):
    """rows per slab: uniform (the kernel maps a row to its slab by one division), at least 2; the last slabs may be
    shorter or empty"""
    return max(2, -(-ny // world))


def row_slabs(ny, world):
    """[r0, r1) image-row slab finished by each rank; contiguous, covering (trailing slabs may be empty)"""
    rps = slab_rows(ny, world)
    return [(min(ny, r * rps), min(ny, (r + 1) * rps)) for r in range(world)]


def exchange_volume_handles(ctx, dist, ndir, world, rank):
    """Every rank exports CUDA IPC handles of all its message volumes; returns table[p][r] = device pointer (local or
    peer-mapped) of rank r's volume of sweep p, and the list of peer mappings to close."""
    handles = []
    for p in range(ndir):
        ptr, nbytes = ctx.sweep_volume(p)
        if not ptr:
            raise RuntimeError("message volume %d is not allocated on rank %d (sweeps_alloc)" % (p, rank))
        handles.append(ctx.ipc_export(ptr))
    gathered = [None] * world
    dist.all_gather_object(gathered, handles)
    table, opened = [], []
    for p in range(ndir):
        row = []
        for r in range(world):
            if r == rank:
                row.append(ctx.sweep_volume(p)[0])
            else:
                q = ctx.ipc_open(gathered[r][p])
                opened.append(q)
                row.append(q)
        table.append(row)
    return table, opened


class SweepSharded:
    """One stereo pair, its sweeps sharded over the ranks of a process group (one process per GPU)."""

    def __init__(self, ctx, dist, torch, stream, W, H, dmin, dmax, P1, P2, NDIR, K, felz, refinement, world, rank):
        self.ctx, self.dist, self.torch, self.stream = ctx, dist, torch, stream
        self.W, self.H, self.dmin, self.dmax = W, H, dmin, dmax
        self.P1, self.P2, self.NDIR, self.K, self.felz, self.refinement = P1, P2, NDIR, K, felz, refinement
        self.world, self.rank = world, rank
        self.mask = sweep_mask(NDIR, world, rank)
        self.rps = slab_rows(H, world)
        self.r0, self.r1 = row_slabs(H, world)[rank]
        self.table = None
        self.opened = []

    def setup(self, dcc):
        torch, W, H = self.torch, self.W, self.H
        self.ctx.sweeps_alloc(W, H, self.dmin, self.dmax, self.NDIR)
        self.ctx.synchronize()
        self.table, self.opened = exchange_volume_handles(self.ctx, self.dist, self.NDIR, self.world, self.rank)
        # maps padded to world * slab_rows rows so that the slabs all-gather into one tensor
        rows = self.world * self.rps
        self.gout = torch.empty((rows, W), dtype=torch.float32, device="cuda")
        self.gcost = torch.empty((rows, W), dtype=torch.float32, device="cuda")
        self.lout = torch.zeros((self.rps, W), dtype=torch.float32, device="cuda")
        self.lcost = torch.zeros((self.rps, W), dtype=torch.float32, device="cuda")
        self.full_out = torch.empty((H, W), dtype=torch.float32, device="cuda")
        self.full_cost = torch.empty((H, W), dtype=torch.float32, device="cuda")
        self.token = torch.zeros(1, device="cuda")
        self.partial = None

    def close(self):
        for q in self.opened:
            self.ctx.ipc_close(q)
        self.opened = []

    def describe(self, exchange):
        per = (self.NDIR + self.world - 1) // self.world
        if exchange == "ordered":
            return ("one pair, %d sweep(s) per GPU; messages stored by the aggregation kernel into the row slab owner's "
                    "volumes (peer stores over NVLink), one barrier, sweep-ordered local finish of %d rows per GPU, "
                    "all-gather of the two maps" % (per, self.rps))
        return ("one pair, %d sweep(s) per GPU; per-GPU partial sums, one NCCL all-reduce of the %dx%dx%d volume, finish of "
                "%d rows per GPU, all-gather of the two maps" % (per, self.W, self.H, self.dmax - self.dmin + 1, self.rps))

    def step(self, dcc, dout, dcost, exchange="ordered"):
        """cost volume resident (on every rank) -> the two full maps resident on every rank"""
        ctx, dist, torch, W, H = self.ctx, self.dist, self.torch, self.W, self.H
        a = (W, H, self.dmin, self.dmax, self.P1, self.P2, self.NDIR, self.K, self.felz)
        with torch.cuda.stream(self.stream):
            if exchange == "ordered":
                ctx.aggregate_sweeps_slabs_dev(dcc.data_ptr(), 0, 0, *a, self.mask, self.world, self.rps, self.table)
                dist.all_reduce(self.token)   # barrier on the stream: every rank's peer stores are complete
                if self.r1 > self.r0:
                    ctx.finish_rows_dev([self.table[p][self.rank] for p in range(self.NDIR)], dcc.data_ptr(), W, H,
                                        self.dmin, self.dmax, self.NDIR, 1, self.refinement, self.r0, self.r1,
                                        self.full_out.data_ptr(), self.full_cost.data_ptr())
            else:
                VS = ctx.padded_labels(self.dmax - self.dmin + 1)
                if self.partial is None:
                    self.partial = torch.empty((H, W, VS), dtype=torch.float32, device="cuda")
                ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, *a, self.mask)
                ctx.sum_sweeps_dev(W, H, self.dmin, self.dmax, self.mask, self.partial.data_ptr())
                dist.all_reduce(self.partial)
                if self.r1 > self.r0:
                    ctx.finish_sum_dev(self.partial.data_ptr(), dcc.data_ptr(), W, H, self.dmin, self.dmax, self.NDIR, 1,
                                       self.refinement, self.r0, self.r1, self.full_out.data_ptr(),
                                       self.full_cost.data_ptr())
            n = self.r1 - self.r0
            if n > 0:
                self.lout[:n].copy_(self.full_out[self.r0:self.r1])
                self.lcost[:n].copy_(self.full_cost[self.r0:self.r1])
            dist.all_gather_into_tensor(self.gout, self.lout)
            dist.all_gather_into_tensor(self.gcost, self.lcost)
            for r in range(self.world):
                a0, a1 = min(H, r * self.rps), min(H, (r + 1) * self.rps)
                if a1 > a0:
                    dout[a0:a1].copy_(self.gout[r * self.rps:r * self.rps + (a1 - a0)])
                    dcost[a0:a1].copy_(self.gcost[r * self.rps:r * self.rps + (a1 - a0)])
