"""Host-side partition of a source cloud over the GPUs of one node (contiguous source-point ranges,
full target map replicated, one 29-value exchange per evaluation)."""


def source_range(n_points, rank, world):
    """[lo, hi) of `rank`'s contiguous slice of an n_points cloud: slices differ by at most one 32-point group and
    start on a group boundary (the kernels hand 32-point groups to warps)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    groups = (n_points + 31) // 32
    base, extra = divmod(groups, world)
    g_lo = rank * base + min(rank, extra)
    g_hi = g_lo + base + (1 if rank < extra else 0)
    return min(n_points, g_lo * 32), min(n_points, g_hi * 32)


class ShardedNdt:
    """One source cloud split over the GPUs of a node (one process per GPU, torch.distributed for the plumbing).

    Every rank builds the full target map, uploads its contiguous slice of the source, and all ranks run the same
    persistent solve; the per-evaluation sums are exchanged inside the kernel (P2P mailboxes over NVLink)."""

    def __init__(self, ndt, dist):
        self.ndt = ndt
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self._attached_n = None

    def setInputTarget(self, points, is_dense=True):
        return self.ndt.setInputTarget(points, is_dense)

    def setInputSource(self, points):
        n = len(points)
        lo, hi = source_range(n, self.rank, self.world)
        self.ndt.setInputSource(points[lo:hi])
        if self.world > 1:
            mine = self.ndt.comm_export()
            handles = [None] * self.world
            self.dist.all_gather_object(handles, mine)
            self.ndt.comm_attach(self.rank, self.world, handles, n)
            self.dist.barrier()
        self._attached_n = n

    def align(self, guess=None):
        if self.world > 1:
            self.dist.barrier()   # all ranks launch together (the kernels wait on each other)
        return self.ndt.align(guess)

    def eval_derivatives(self, p, compute_hessian=True):
        if self.world > 1:
            self.dist.barrier()
        return self.ndt.eval_derivatives(p, compute_hessian=compute_hessian)

    def result(self):
        return self.ndt.result()
