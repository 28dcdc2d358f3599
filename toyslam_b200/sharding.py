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


def point_range(n_points, rank, world):
    """[lo, hi) of `rank`'s contiguous slice of an n_points target cloud (sizes differ by at most one point)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_points, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


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

    def setInputTargetSharded(self, points, is_dense=True):
        """Target-map build split over the ranks: every rank sorts / reduces only its contiguous range
        of `points`, one exchange of per-voxel partials {key, count, sum x, sum x x^T}, then every rank merges all
        partials in rank order and ends with the full map (identical bits everywhere).  Collectives (NCCL, device
        tensors): min/max/sum all-reduce of the bounding box, all-gather of the partials."""
        import numpy as np
        import torch
        n = len(points)
        lo, hi = point_range(n, self.rank, self.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        local = torch.ones((hi - lo, 4), dtype=torch.float32, device=dev)
        if hi > lo:
            local[:, :3] = torch.as_tensor(np.ascontiguousarray(points[lo:hi, :3], dtype=np.float32)).to(dev)
        return self.setInputTargetShardedDevice(local, is_dense)

    def setInputTargetShardedDevice(self, local, is_dense=True, mode="owner"):
        """Same, for a slice that already lives on this rank's GPU: `local` is an (n_r, 4) float32 CUDA tensor.

        mode "owner" (default, SURVEY §8e row 3 as designed): the per-voxel partials travel to the rank that OWNS their key
        range (all-to-all, 88 B per partial), the owner merges them in rank order and finalises its voxels, the finished
        records (64 B + 48 B fp64 inverse covariance per voxel) are all-gathered — owners hold ascending key ranges, so
        the concatenation in rank order is the key-sorted map — and every rank builds its voxel index.
        mode "replicated": every rank gathers and merges ALL partials (the round-1 scheme; keeps the moments on every
        rank for the parity dump)."""
        import os
        import time
        import numpy as np
        import torch
        dist, ndt = self.dist, self.ndt
        dev = local.device
        W = self.world
        timing = os.environ.get("NDTB200_SHARD_TIMING") is not None
        marks = []

        def mark(name):
            if timing:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))
        mark("start")
        mn, mx, nf = ndt.cloud_bounds(local.data_ptr(), local.shape[0], is_dense)
        mark("cloud_bounds")
        # one MIN all-reduce for the box (max = -min(-x)), one SUM for the finite count
        t_box = torch.tensor(np.concatenate([mn, -mx]), dtype=torch.float32, device=dev)
        t_nf = torch.tensor([nf], dtype=torch.int64, device=dev)
        if W > 1:
            dist.all_reduce(t_box, op=dist.ReduceOp.MIN)
            dist.all_reduce(t_nf, op=dist.ReduceOp.SUM)
        box = t_box.cpu().numpy()
        gmin, gmax, nf_total = box[:3].copy(), (-box[3:]).copy(), int(t_nf.item())
        mark("bbox_allreduce")
        st, nv = ndt.build_partials(gmin, gmax)
        nv = nv if st == 0 else 0
        mark("build_partials")
        keys = torch.zeros(max(1, nv), dtype=torch.int32, device=dev)
        cnts = torch.zeros(max(1, nv), dtype=torch.int32, device=dev)
        moms = torch.zeros((max(1, nv), 9), dtype=torch.float64, device=dev)
        if nv > 0:
            ndt.copy_partials(keys.data_ptr(), cnts.data_ptr(), moms.data_ptr())
        mark("copy_partials")
        if mode == "replicated" or W == 1:
            return self._merge_replicated(local, keys, cnts, moms, nv, gmin, gmax, nf_total)
        # ---- owners: ascending key ranges balanced by voxel count (splitters from a sample of every rank's sorted keys) ----
        S = 256
        sample = torch.full((S,), 2 ** 31 - 1, dtype=torch.int32, device=dev)
        if nv > 0:
            pos = torch.linspace(0, nv - 1, S, device=dev).long()
            sample = keys[pos]
        all_samples = torch.empty(W * S, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_samples, sample)
        srt = torch.sort(all_samples).values
        upper = [int(srt[(r + 1) * S].item()) for r in range(W - 1)]          # first key NOT owned by rank r
        offs = ndt.partials_split(upper, W)
        mark("splitters")
        send = torch.tensor(np.diff(offs), dtype=torch.int64, device=dev)
        recv = torch.empty(W, dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send)
        send_l, recv_l = [int(x) for x in send.tolist()], [int(x) for x in recv.tolist()]
        n_in = int(sum(recv_l))
        r_keys = torch.empty(max(1, n_in), dtype=torch.int32, device=dev)
        r_cnts = torch.empty(max(1, n_in), dtype=torch.int32, device=dev)
        r_moms = torch.empty((max(1, n_in), 9), dtype=torch.float64, device=dev)
        dist.all_to_all_single(r_keys[:n_in], keys[:nv], recv_l, send_l)
        dist.all_to_all_single(r_cnts[:n_in], cnts[:nv], recv_l, send_l)
        dist.all_to_all_single(r_moms[:n_in], moms[:nv], recv_l, send_l)
        torch.cuda.synchronize()
        mark("all_to_all")
        st2, n_own = ndt.merge_partials(gmin, gmax, nf_total, r_keys.data_ptr(), r_cnts.data_ptr(), r_moms.data_ptr(), n_in)
        if st2 not in (0,):
            n_own = 0
        mark("merge_partials")
        # ---- all-gather the finished records in rank (= key) order ----
        own = torch.tensor([n_own], dtype=torch.int64, device=dev)
        owns = torch.empty(W, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(owns, own)
        owns_l = [int(x) for x in owns.tolist()]
        total = int(sum(owns_l))
        my_off = int(sum(owns_l[:self.rank]))
        # uneven all-gather straight into the final arrays: every owner broadcasts its slice (owners hold ascending key
        # ranges, so the rank-ordered concatenation IS the key-sorted map; no padding, no re-packing)
        g_rec = torch.empty((max(1, total), 16), dtype=torch.int32, device=dev)       # 64-byte records
        g_ic = torch.empty((max(1, total), 6), dtype=torch.float64, device=dev)
        if n_own > 0:
            ndt.copy_records(g_rec[my_off:my_off + n_own].data_ptr(), g_ic[my_off:my_off + n_own].data_ptr())
        off = 0
        works = []
        for r in range(W):
            if owns_l[r] > 0:
                works.append(dist.broadcast(g_rec[off:off + owns_l[r]], src=r, async_op=True))
                works.append(dist.broadcast(g_ic[off:off + owns_l[r]], src=r, async_op=True))
            off += owns_l[r]
        for wk in works:
            wk.wait()
        torch.cuda.synchronize()
        mark("allgather_records")
        self._keep = (local, g_rec, g_ic)
        st3 = ndt.set_map_from_records(gmin, gmax, nf_total, g_rec.data_ptr(), g_ic.data_ptr(), total)
        mark("set_map_from_records")
        if timing and self.rank == 0:
            print("SHARD_TIMING " + " ".join("%s=%.3fms" % (b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks[:-1], marks[1:])), flush=True)
        return st3

    def _merge_replicated(self, local, keys, cnts, moms, nv, gmin, gmax, nf_total):
        import torch
        dist, ndt = self.dist, self.ndt
        dev = local.device
        counts = torch.tensor([nv], dtype=torch.int64, device=dev)
        all_counts = [torch.zeros_like(counts) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(all_counts, counts)
        else:
            all_counts = [counts]
        sizes = [int(c.item()) for c in all_counts]
        cap = max(1, max(sizes))
        if keys.shape[0] < cap:
            pad = cap - keys.shape[0]
            keys = torch.cat([keys, torch.zeros(pad, dtype=keys.dtype, device=dev)])
            cnts = torch.cat([cnts, torch.zeros(pad, dtype=cnts.dtype, device=dev)])
            moms = torch.cat([moms, torch.zeros((pad, 9), dtype=moms.dtype, device=dev)])
        keys, cnts, moms = keys[:cap].contiguous(), cnts[:cap].contiguous(), moms[:cap].contiguous()
        if self.world > 1:
            g_keys = torch.empty(self.world * cap, dtype=torch.int32, device=dev)
            g_cnts = torch.empty(self.world * cap, dtype=torch.int32, device=dev)
            g_moms = torch.empty((self.world * cap, 9), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(g_keys, keys)
            dist.all_gather_into_tensor(g_cnts, cnts)
            dist.all_gather_into_tensor(g_moms, moms)
            sel = torch.cat([torch.arange(r * cap, r * cap + sizes[r], device=dev) for r in range(self.world)])
            keys, cnts, moms = g_keys[sel].contiguous(), g_cnts[sel].contiguous(), g_moms[sel].contiguous()
        else:
            keys, cnts, moms = keys[:sizes[0]].contiguous(), cnts[:sizes[0]].contiguous(), moms[:sizes[0]].contiguous()
        total = int(sum(sizes))
        torch.cuda.synchronize()
        self._keep = (local, keys, cnts, moms)
        return ndt.build_from_partials(gmin, gmax, nf_total, keys.data_ptr(), cnts.data_ptr(), moms.data_ptr(), total)

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

    def getFitnessScore(self, max_range=float("inf")):
        """pcl::Registration::getFitnessScore of the whole (sharded) source: every rank searches the nearest raw target
        point of its slice, the ranks all-reduce {sum of squared distances, count}."""
        import sys
        import torch
        mr = sys.float_info.max if max_range == float("inf") else float(max_range)
        s, c = self.ndt.fitness_sums(mr)
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.tensor([s, float(c)], dtype=torch.float64, device=dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            s, c = float(t[0].item()), int(round(float(t[1].item())))
        return s / c if c > 0 else sys.float_info.max
