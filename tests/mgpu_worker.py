"""Worker of tests/test_multigpu.py: run under torchrun, one rank per GPU.

Owner-sharded count pass + distributed lookup pass against the single-table oracle
(the reference's own criterion: identical results for any rank count, test/runFilterTests.sh:93-116)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import kmernator_b200 as K
    import oracle
    from bench import synth
    from tests.gpu_util import assert_tables_equal, oracle_table

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")

    def make_ctx(k, **kw):
        ctx = K.Context(kmer_size=k, device=local, **kw)
        uid = [K.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        ctx.comm_init(rank, world, uid[0])
        return ctx

    for k, n_reads, uneven in ((31, 12000, False), (63, 6000, False), (31, 5000, True), (21, 4000, False)):
        bases, q, off = synth.reads_numpy(n_reads, 150, 30000, seed=21 + k, err=0.004, lowq=0.002, n_rate=0.001, var_len=True)
        disc = (np.arange(n_reads) % 53 == 0).astype(np.uint8)
        # contiguous read slices per rank (ReadSet::appendAllFiles(files, rank, size), src/ReadSet.cpp:186-258);
        # `uneven`: the last rank gets nothing at all
        bounds = np.linspace(0, n_reads, world + 1).astype(int)
        if uneven:
            bounds = np.linspace(0, n_reads, world).astype(int).tolist() + [n_reads]
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        b0, b1 = int(off[r0]), int(off[r1])
        my_b, my_q = bases[b0:b1], q[b0:b1]
        my_off = np.ascontiguousarray(off[r0:r1 + 1] - off[r0])
        my_disc = np.ascontiguousarray(disc[r0:r1])

        # (a staging size that makes the per-owner overflow list an odd number of records before rounding: the receive
        #  buffers of the peers must still start on sector boundaries)
        ctx = make_ctx(k, table_slots=1 << 21, stage_keys=(1 << 20) + 768 * world)
        # two batches per rank (a rank without reads still has to take part in the collectives)
        n_my = r1 - r0
        half = n_my // 2
        cut = int(my_off[half])
        ctx.count_batch(my_b[:cut], my_q[:cut], np.ascontiguousarray(my_off[:half + 1]), discarded=my_disc[:half])
        ctx.count_batch(my_b[cut:], my_q[cut:], np.ascontiguousarray(my_off[half:] - my_off[half]), discarded=my_disc[half:])
        ctx.count_finish(apply_purge=False)

        g = ctx.export()
        # every exported k-mer is owned by this rank: ((hash>>24)&0x7ffff) % size   (src/Kmer.h:2284-2295)
        for key in g["keys"][:: max(1, len(g["keys"]) // 3000)]:
            assert oracle.owner(oracle.kmer_hash(key.tobytes()), world) == rank
        hist = ctx.histogram()                      # all-reduced (MPIHistogram::reduce)
        st = ctx.stats()
        parts = [None] * world
        dist.all_gather_object(parts, (g["keys"], g["count"], g["dir"], st["raw_kmers"], st["raw_good_kmers"], st["unique_kmers"]))

        osp = oracle_table(bases, q, off, k, disc=disc, threads=4)
        o = osp.export()
        keys = np.concatenate([p[0] for p in parts])
        cnt = np.concatenate([p[1] for p in parts])
        dr = np.concatenate([p[2] for p in parts])
        order = np.lexsort(keys.T[::-1])
        assert_tables_equal(dict(keys=keys[order], count=cnt[order], dir=dr[order]), o)
        ost = osp.stats()
        assert sum(p[3] for p in parts) == ost["raw"] and sum(p[4] for p in parts) == ost["raw_good"]
        assert sum(p[5] for p in parts) == ost["unique"]
        exact = np.bincount(o["count"], minlength=65536).astype(np.uint64)
        assert (hist == exact).all()

        # kmn_lookup with a communicator (collective; getElementIfExists on a distributed map): every rank asks for a
        # different sample of the k-mers, most of them owned by other ranks, plus keys that are in no table
        okeys, ocnt = o["keys"], o["count"]
        sel = np.arange(rank, len(okeys), 5)[:3000]
        absent = np.random.default_rng(100 + rank).integers(0, 256, (40, okeys.shape[1]), dtype=np.uint8)
        absent[:, -1] &= np.uint8((0xFF << (2 * ((4 - k % 4) % 4))) & 0xFF)
        lut = {bytes(x): int(cn) for x, cn in zip(okeys, ocnt)}
        ask = np.concatenate([okeys[sel], absent])
        want = np.concatenate([ocnt[sel], np.array([lut.get(bytes(x), 0) for x in absent], dtype=ocnt.dtype)])
        got_counts = ctx.lookup(ask if rank != world - 1 or not uneven else ask[:0])
        if not (uneven and rank == world - 1):
            assert (got_counts == want).all(), (k, "lookup")

        # lookup pass on this rank's reads, against the oracle on the same reads
        ctx.purge_min_depth(2)
        osp.purge_min_depth(2)
        for scoring in ("MAX", "MEDIAN"):
            got = ctx.trim_batch(my_b, my_off, 2, scoring, discarded=my_disc, n_reads=n_my)
            if n_my:
                exp = osp.trim_reads(my_b, my_off, 2, oracle.SCORING[scoring], my_disc, threads=2)
                for a, b, name in zip(got, exp, ("off", "len", "score", "was")):
                    assert (a == b).all(), (k, scoring, name)
        ctx.close()
        dist.barrier()
        if rank == 0:
            print("mgpu ok: k=%d reads=%d ranks=%d uneven=%s distinct=%d" % (k, n_reads, world, uneven, len(o["count"])), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
