"""Multi-GPU form of the path: one process per GPU, the sampled reads sharded
across ranks, every rank scans its shard for ALL query k-mers, and the per-k-mer
count vectors are summed with one small all-reduce.

The reference's only parallelism is an OpenMP team over k-mers sharing one index
(/root/reference/approx_counter.cpp:547-599); reads are independent and the
result is a sum over reads (:589-596), so sharding reads needs no other exchange.
On GPUs the all-reduce is libapc's own (apc_comm_init_rank + apc_scan_allreduce, include/apc.h: one
ncclAllReduce on the context's stream); torch.distributed only carries the communicator's id to the ranks.
CPU tests of the host logic run the same class over gloo with torch's all-reduce.
"""
import numpy as np


def shard_bounds(n_reads, rank, world, align=32):
    """Contiguous block of reads for `rank`: ceil(n_tiles / world) scan tiles
    (32 reads) per rank, so every shard but the last is tile-aligned."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    tiles = (n_reads + align - 1) // align
    per = (tiles + world - 1) // world
    lo = min(n_reads, rank * per * align)
    hi = min(n_reads, (rank + 1) * per * align)
    return lo, hi


def allreduce_counts(counts, group=None):
    """Sum the per-rank count vectors in place.  `counts`: torch int64 tensor
    (CUDA for NCCL, CPU for gloo); uint64 counts are carried as int64 (sums stay
    far below 2^63: at most 3 x #reads per k-mer)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class ShardedApproxCounter:
    """One rank's view: upload this rank's shard, scan on the torch current
    stream into a torch tensor, all-reduce, read back."""

    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist
        from .api import ApproxCounter
        self.torch = torch
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.counter = ApproxCounter(self.device)
        self.counter.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._counts = None
        # the ranks' contexts form one NCCL communicator through the C ABI; rank 0 creates the id
        self.abi_comm = False
        if self.world > 1 and dist.get_backend(group) == "nccl":
            box = [ApproxCounter.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self.counter.comm_init_rank(self.world, self.rank, box[0])
            self.abi_comm = True

    def close(self):
        self.counter.close()

    def upload_sample(self, sample):
        """sample: the WHOLE uint8[n, L] matrix; only this rank's rows are uploaded."""
        lo, hi = shard_bounds(sample.shape[0], self.rank, self.world)
        self.counter.upload_sample(np.ascontiguousarray(sample[lo:hi]))
        return lo, hi

    def upload_shard(self, rows):
        self.counter.upload_sample(rows)

    def set_queries(self, kmers, k):
        self.counter.set_queries(kmers, k)
        n = len(kmers)
        if self._counts is None or self._counts.numel() != n:
            self._counts = self.torch.zeros(max(n, 1), dtype=self.torch.int64, device=f"cuda:{self.device}")[:n]

    def scan_allreduce(self):
        """Asynchronous on the torch current stream; returns the device tensor."""
        self.counter.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)
        if self._counts.numel():
            if self.abi_comm:
                self.counter.scan_allreduce(self._counts.data_ptr())
            else:
                self.counter.scan(self._counts.data_ptr())
                allreduce_counts(self._counts, self.group)
        return self._counts

    def errorCount(self, kmers, k):
        """errorCount (:531-601) over the sharded sample; every rank gets the totals."""
        self.set_queries(kmers, k)
        return self.scan_allreduce().cpu().numpy().astype(np.uint64)
