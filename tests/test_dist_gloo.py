"""N>1 host path on CPU: world_size-2 gloo processes.

(1) the rendezvous plumbing of kmernator_b200.dist (slices, global offsets, raw-k-mer estimate, id broadcast);
(2) the sharding contract itself, emulated with the oracle: every rank extracts the k-mers of its read slice, routes
    each to owner = ((hashlittle2 >> 24) & 0x7ffff) % size (src/Kmer.h:2284-2295), counts what it receives, and answers
    the other rank's lookup requests in order (src/DistributedFunctions.h:877-902) -- the merged result must equal
    the single-rank oracle (the reference's own invariance criterion, test/runFilterTests.sh:93-116)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeCtx:
    """stands in for kmernator_b200.Context (no GPU here): records what comm_init received"""
    got = None

    @staticmethod
    def comm_unique_id():
        return np.arange(128, dtype=np.uint8)

    def comm_init(self, rank, nranks, uid):
        self.got = (rank, nranks, bytes(np.asarray(uid, dtype=np.uint8).tobytes()))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from bench import synth
    from kmernator_b200 import dist as kd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k, n_reads = 31, 600
        bases, quals, off = synth.reads_numpy(n_reads, 100, 4000, seed=5, err=0.01, lowq=0.003, n_rate=0.002)
        lo, hi = kd.rank_slice(n_reads, rank, world)
        goff, total = kd.global_read_offsets(hi - lo)
        assert (goff, total) == (lo, n_reads)
        est = kd.estimate_raw_kmers(hi - lo, int(off[hi] - off[lo]), k)
        assert est == sum((100 - k + 1) * (b - a) for a, b in (kd.rank_slice(n_reads, r, world) for r in range(world)))
        ctx = _FakeCtx()
        kd.init_comm(ctx)
        assert ctx.got == (rank, world, bytes(range(128)))

        # ---- count pass: route every good k-mer of my reads to its owner
        outbox = [[] for _ in range(world)]
        my_kmers = []                                       # (read, pos, key bytes) for the lookup pass
        for r in range(lo, hi):
            s = bases[int(off[r]): int(off[r + 1])].tobytes()
            qq = quals[int(off[r]): int(off[r + 1])].tobytes()
            keys, fw, wt, _ = oracle.read_kmers(s, qq, k, with_ext=False)
            for i in range(len(keys)):
                kb = keys[i].tobytes()
                own = oracle.owner(oracle.kmer_hash(kb), world)
                my_kmers.append((r, i, kb, own))
                if wt[i] > np.float32(0.10):
                    outbox[own].append(kb)
        gathered = [None] * world
        dist.all_gather_object(gathered, outbox)            # gathered[src][dst]
        table = {}
        for src in range(world):
            for kb in gathered[src][rank]:
                table[kb] = min(65535, table.get(kb, 0) + 1)
        for kb in table:
            assert oracle.owner(oracle.kmer_hash(kb), world) == rank

        # ---- lookup pass: requests out in order, answers back in the same order
        req = [[kb for (_, _, kb, own) in my_kmers if own == d] for d in range(world)]
        allreq = [None] * world
        dist.all_gather_object(allreq, req)
        ans = [[(table.get(kb, 0) if table.get(kb, 0) >= 2 else 0) for kb in allreq[src][rank]] for src in range(world)]
        allans = [None] * world
        dist.all_gather_object(allans, ans)                 # allans[owner][requester]
        cursor = [0] * world
        vals = {}
        for (r, i, kb, own) in my_kmers:
            vals[(r, i)] = allans[own][rank][cursor[own]]
            cursor[own] += 1

        # ---- single-rank oracle on everything
        osp = oracle.OracleSpectrum(k, est_distinct=1 << 14)
        osp.add_reads(bases.tobytes(), quals, off)
        e = osp.export()
        full = {kk.tobytes(): int(c) for kk, c in zip(e["keys"], e["count"])}
        mine = {kb: c for kb, c in full.items() if oracle.owner(oracle.kmer_hash(kb), world) == rank}
        assert mine == table
        osp.purge_min_depth(2)
        for (r, i, kb, own) in my_kmers:
            assert vals[(r, i)] == osp.lookup(kb)
        sizes = [None] * world
        dist.all_gather_object(sizes, len(table))
        assert sum(sizes) == len(full)
        q.put((rank, "ok", len(table)))
    except Exception as ex:                                 # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_rank_slice_properties():
    from kmernator_b200 import dist as kd
    for n in (0, 1, 7, 1000):
        for size in (1, 2, 3, 8):
            prev = 0
            for r in range(size):
                lo, hi = kd.rank_slice(n, r, size)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
    with pytest.raises(ValueError):
        kd.rank_slice(10, 2, 2)


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", "rank %d: %s" % (rank, info)
    assert all(info > 0 for _, _, info in res)
