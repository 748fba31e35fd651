"""Host-side plumbing of the multi-GPU path (one process per GPU, torch.distributed for rendezvous only).

Mirrors what the reference does around its MPI communicator:
  * rank_slice            -- contiguous per-rank slice of the input (ReadSet::appendAllFiles(files, rank, size),
                             src/ReadSet.cpp:186-258, src/ReadFileReader.h:379-398)
  * global_read_offsets   -- setGlobalReadSetConstants (src/DistributedFunctions.h:73-100): every rank learns the
                             global index of its first read
  * estimate_raw_kmers    -- KS::estimateRawKmers(world, reads) (src/DistributedFunctions.h:154-161)
  * init_comm             -- ScopedMPIComm (src/MPIUtils.h:256-391): rank 0 creates the NCCL id, everybody joins

The data path itself (k-mer all-to-all, lookup requests/answers, histogram all-reduce) is inside the CUDA library
(kmn_comm_init / kmn_count_batch / kmn_trim_batch / kmn_histogram) and uses NCCL directly.
"""
import numpy as np


def rank_slice(n_items, rank, size):
    """[lo, hi) of `n_items` owned by `rank` of `size`: contiguous, ordered, covering, sizes differ by at most 1."""
    if size < 1 or not (0 <= rank < size):
        raise ValueError("bad rank %d / %d" % (rank, size))
    lo = (n_items * rank) // size
    hi = (n_items * (rank + 1)) // size
    return lo, hi


def global_read_offsets(n_local_reads, group=None):
    """Returns (my_global_offset, total_reads) -- all_gather of the per-rank read counts."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0, n_local_reads
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([n_local_reads], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    counts = [int(t.item()) for t in allc]
    return sum(counts[:rank]), sum(counts)


def estimate_raw_kmers(n_local_reads, n_local_bases, k, group=None):
    """(avgLen - k + 1) * numReads, minimum 128 per rank, summed over ranks (src/KmerSpectrum.h:573-584)."""
    import torch
    import torch.distributed as dist

    local = 128
    if n_local_reads > 0:
        avg = n_local_bases // n_local_reads
        if avg > k:
            local = max(128, (avg - k + 1) * n_local_reads)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    t = torch.tensor([local], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return int(t.item())


def init_comm(ctx, group=None, make_id=None):
    """Joins `ctx` (a kmernator_b200.Context) to a communicator spanning the torch.distributed group."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [None]
    if rank == 0:
        box[0] = bytes(np.asarray((make_id or type(ctx).comm_unique_id)(), dtype=np.uint8).tobytes())
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(rank, world, np.frombuffer(box[0], dtype=np.uint8))
