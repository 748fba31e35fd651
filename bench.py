#!/usr/bin/env python
"""bench.py -- k-mer insertions/sec (k=31, 150 bp reads) on N B200, vs the CPU restatement of FilterReads-P.

Contract (one JSON line on rank 0):
  value    = whole-job k-mer insertions/s of the count pass (kmn_reset + kmn_count_batch + kmn_count_finish incl. the
             post-build min-depth purge) with the reads already resident in HBM, CUDA events on the library's stream,
             max over ranks.
  e2e      = same metric through the C ABI with HOST (pinned) buffers: H2D of every batch and a D2H read of the
             spectrum counters inside the timed region (kmn_count_batch_2na on TwoBitSequence-packed reads; the ASCII
             entry beside it).  One GPU: on its own context with eight drains per step.
  roofline = the dominant kernel (largest share of the step) with ITS algorithmic bytes per launch over its
             measured launch time vs MEASURED_PEAKS.json hbm_gbs, `traffic` from profiles/r02_traffic.json (ncu);
             `per_kernel` lists the three record passes; `whole_pass` is SURVEY.md 8d's figure for the pass as a whole
             (66.5 B per presented instance) -- the one north_star's ">= 0.50" refers to.
  checks   = full-size invariants + an exact comparison with the CPU oracle at the run's rank count (untimed).
  lookup_pass / exchange = kmn_trim_batch on 20 M reads / bytes pushed between GPUs per step.
  cpu_baseline = the oracle port (reference binary is not buildable: no MPI/Boost) on a bounded sample, all host threads.

`--impl reference` times only that CPU port (the one other place oracle/ may be executed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
READ_LEN = 150
KMERS_PER_READ = READ_LEN - K + 1
ALGO_BYTES_PER_INSERT = 64.0      # table RMW: one 32-B sector in, one out
ALGO_BYTES_PER_INSTANCE = 66.5    # + 2.5 B streamed input per instance (whole count pass)

# BASELINE.json configs as concrete synthetic inputs.  c2 is the configuration the metric is quoted on (and
# what the driver runs); the others are extra bench lines (`--workload`), each also covered by a small GPU-vs-oracle test.
WORKLOADS = {
    "c2": dict(k=31, reads=100_000_000, genome=250_000_000, genome_seed=0x4B6D6572,
               desc="C2: k=31 count + min-depth-2 purge on %d synthetic 150bp reads per GPU (60x coverage, 0.1%% substitutions)"),
    "c4": dict(k=31, reads=100_000_000, genome=0, genome_seed=0x4D455441, meta=dict(n_genomes=1000, min_len=250_000, max_len=4_000_000, sigma=2.0),
               desc="C4 (half scale, fits one GPU): k=31 on %d reads of a skewed metagenome (1000 genomes 0.25-4 Mbp log-uniform, abundances "
                    "log-normal sigma 2), + max-depth-100 / min-depth-2 normalisation on the lookup pass"),
    "c5a": dict(k=63, reads=20_000_000, genome=250_000_000, genome_seed=0x4B6D6572,
                desc="C5a: k=63 (two-word keys) count + min-depth-2 purge on %d synthetic 150bp reads of the C2 genome"),
    "c5b": dict(k=21, reads=50_000_000, genome=5386, genome_seed=0x50484958, min_kmer_quality=0.0, min_quality_score=2, ext=True,
                desc="C5b: k=21 MeraculousCounter-style spectrum (extension counters) on %d reads of a 5386-bp genome: every genomic count saturates"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (0 = the workload's own size; C2: 100M x 150 bp)")
    ap.add_argument("--genome", type=int, default=0, help="genome length per GPU-worth of reads (0 = the workload's own; C2: 250 Mbp, 60x)")
    ap.add_argument("--table-slots", type=int, default=0)
    ap.add_argument("--stage-keys", type=int, default=0)
    ap.add_argument("--slice-mb", type=int, default=64)
    ap.add_argument("--pipe-batches", type=int, default=0, help="drains (multi-GPU: rounds) per step: a staging set holds 1/N of the input; 0 = 2 on one GPU, 8 on several")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per e2e step (0 = same as --reads)")
    ap.add_argument("--e2e-batch", type=int, default=4_000_000)
    ap.add_argument("--e2e-pipe-batches", type=int, default=8, help="staging sets per step of the end-to-end run on one GPU")
    ap.add_argument("--cpu-reads", type=int, default=2_000_000, help="reads of the bounded CPU sample (about 10 s of work for 16 host threads)")
    ap.add_argument("--parity-reads", type=int, default=240_000, help="reads of the untimed exact oracle comparison (all ranks together)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lookup", action="store_true")
    ap.add_argument("--no-checks", action="store_true")
    ap.add_argument("--lookup-reads", type=int, default=20_000_000)
    return ap.parse_args()


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_port_rate(bases_np, quals_np, off_np, threads, k=K, steps=1, warmup=0, min_quality=3, min_kmer_quality=0.10, track_ext=False):
    """oracle port of the count pass (serial/OpenMP T x T build), k-mer instances / s"""
    import oracle
    n_reads = len(off_np) - 1
    inst = int(off_np[-1]) - n_reads * (k - 1)
    times = []
    for it in range(warmup + steps):
        s = oracle.OracleSpectrum(k, min_quality, min_kmer_quality, track_ext=track_ext, threads=threads, est_distinct=max(1 << 16, inst // 3))
        t0 = time.perf_counter()
        s.add_reads(bases_np.tobytes() if not isinstance(bases_np, bytes) else bases_np, quals_np, off_np)
        s.purge_min_depth(2)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        del s
    return inst / (sum(times) / len(times)), sum(times) / len(times)


def cpu_sample(wl, n, genome_len):
    """The bounded CPU sample of a workload: the first n reads of a read set drawn from the workload's full-size genome with
    the same error / quality model, generated with numpy (identical for the reference arm and for cpu_baseline)."""
    import numpy as np
    from bench import synth
    if wl.get("meta"):
        lens, share = synth.metagenome_plan(seed=wl["genome_seed"], **wl["meta"])
        rng = np.random.default_rng(wl["genome_seed"])
        g = rng.integers(0, 4, int(lens.sum()), dtype=np.uint8)
        goff = np.concatenate([[0], np.cumsum(lens)[:-1]])
        gi = np.searchsorted(np.cumsum(share), rng.random(n)).clip(max=len(lens) - 1)
        starts = goff[gi] + (rng.random(n) * (lens[gi] - READ_LEN)).astype(np.int64)
        return synth.reads_numpy(n, READ_LEN, 0, seed=0x5245, genome=g, starts=starts)
    return synth.reads_numpy(n, READ_LEN, genome_len, seed=wl["genome_seed"] & 0x7FFFFFFF)


def run_reference(args, rank, world, emit):
    """--impl reference: the CPU restatement of the reference path on the box's host cores (rank 0 only), on the same
    bounded sample of the workload that the GPU arm's cpu_baseline times."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    k = wl["k"]
    threads = os.cpu_count() or 1
    n = args.cpu_reads
    genome_len = args.genome or wl["genome"]
    bases, quals, off = cpu_sample(wl, n, genome_len)
    rate, dt = cpu_port_rate(bases, quals, off, threads, k=k, steps=max(1, args.steps), warmup=min(args.warmup, 1),
                             min_quality=wl.get("min_quality_score", 3), min_kmer_quality=wl.get("min_kmer_quality", 0.10), track_ext=bool(wl.get("ext")))
    sample = "first %d reads of the workload's synthetic read set (numpy generator, full-size genome, same error/quality model) per step" % n
    line = {
        "impl": "reference", "metric": "k-mer insertions/sec (k=%d, 150bp reads)" % k, "value": rate, "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": (wl["desc"] % (args.reads or wl["reads"])) + " (bounded CPU sample)", "reads_per_step": n},
        "cpu_baseline": {"value": rate, "unit": "kmers/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "restated CPU baseline - reference binary not buildable here (no MPI/Boost)"},
        "e2e": {"value": rate, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def oracle_parity(KM, dist, torch, dev, rank, world, local_rank, n_total, wl):
    """Untimed exact comparison with the CPU oracle at this rank count (the driver's test box has one GPU, so this is
    where N > 1 parity becomes visible): every rank counts its contiguous slice of a fixed read set on a fresh context
    joined to a communicator, the exported tables are gathered, and rank 0 checks keys, counts, the owner rule
    ((hashlittle2 >> 24) & 0x7ffff) % N of every exported key (src/Kmer.h:2284-2295), the all-reduced histogram and the
    trim results of the collective lookup pass against oracle.OracleSpectrum on the same reads
    (the reference's own criterion for the distributed apps, test/runFilterTests.sh:93-116)."""
    import numpy as np
    import oracle
    from bench import synth
    k = wl["k"]
    ext = bool(wl.get("ext"))
    minq, minw = wl.get("min_quality_score", 3), wl.get("min_kmer_quality", 0.10)
    genome = 5386 if wl["genome"] and wl["genome"] < 100_000 else max(20_000, n_total * READ_LEN // 40)
    bases, q, off = synth.reads_numpy(n_total, READ_LEN, genome, seed=77, err=0.003, lowq=0.001, n_rate=0.0005, var_len=True)
    bounds = np.linspace(0, n_total, world + 1).astype(np.int64)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    b0, b1 = int(off[r0]), int(off[r1])
    my_b, my_q = np.ascontiguousarray(bases[b0:b1]), np.ascontiguousarray(q[b0:b1])
    my_off = np.ascontiguousarray(off[r0:r1 + 1] - off[r0])
    inst = int(off[-1])
    ctx = KM.Context(kmer_size=k, min_quality_score=minq, min_kmer_quality=minw, table_slots=max(1 << 20, inst), stage_keys=max(1 << 18, inst // 3),
                     value_kind=KM.capi.KMN_VALUE_DIR_EXT if ext else KM.capi.KMN_VALUE_DIR, device=local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(KM.Context.comm_unique_id().copy())
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, uid.cpu().numpy())
    half = (r1 - r0) // 2
    cut = int(my_off[half])
    ctx.count_batch(my_b[:cut], my_q[:cut], np.ascontiguousarray(my_off[:half + 1]))
    ctx.count_batch(my_b[cut:], my_q[cut:], np.ascontiguousarray(my_off[half:] - my_off[half]))
    ctx.count_finish(apply_purge=False)
    g = ctx.export()
    hist = ctx.histogram()
    st = ctx.stats()
    ctx.purge_min_depth(2)
    n_trim = min(r1 - r0, max(1, 100_000 // world))
    tb0 = int(my_off[n_trim])
    trims = ctx.trim_batch(my_b[:tb0], np.ascontiguousarray(my_off[:n_trim + 1]), 2, "MAX", n_reads=n_trim)
    ctx.close()
    mine = (g["keys"], g["count"], g["dir"], g["ext"], st["raw_kmers"], st["raw_good_kmers"], st["unique_kmers"], r0, n_trim, [np.asarray(t) for t in trims])
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    res = None
    if rank == 0:
        osp = oracle.OracleSpectrum(k, minq, minw, track_ext=ext, threads=os.cpu_count() or 1, est_distinct=max(1 << 16, inst // 2))
        osp.add_reads(bases.tobytes(), q, off)
        o = osp.export()
        keys = np.concatenate([p[0] for p in parts])
        cnt = np.concatenate([p[1] for p in parts])
        dr = np.concatenate([p[2] for p in parts])
        order = np.lexsort(keys.T[::-1])
        ok_keys = keys.shape == o["keys"].shape and bool((keys[order] == o["keys"]).all())
        ok_cnt = ok_keys and bool((cnt[order] == o["count"]).all())
        dd = dr[order].astype(np.int64) - o["dir"].astype(np.int64) if ok_keys else np.array([9])
        ok_dir = bool(((dd == 0) | (dd == 1)).all())
        ok_ext = True
        if ext and ok_keys:
            ok_ext = bool((np.concatenate([p[3] for p in parts])[order] == o["ext"]).all())
        n_owner, ok_owner = 0, True
        for rk, p in enumerate(parts):
            step = max(1, len(p[0]) // 20000)
            for key in p[0][::step]:
                n_owner += 1
                ok_owner = ok_owner and (oracle.owner(oracle.kmer_hash(key.tobytes()), world) == rk)
        ost = osp.stats()
        ok_stats = (sum(p[4] for p in parts), sum(p[5] for p in parts), sum(p[6] for p in parts)) == (ost["raw"], ost["raw_good"], ost["unique"])
        ok_hist = bool((hist == np.bincount(o["count"], minlength=65536).astype(np.uint64)).all())
        osp.purge_min_depth(2)
        ok_trim, n_tr = True, 0
        for p in parts:
            pr0, pn = p[7], p[8]
            if pn == 0:
                continue
            sb0, sb1 = int(off[pr0]), int(off[pr0 + pn])
            exp = osp.trim_reads(bases[sb0:sb1], np.ascontiguousarray(off[pr0:pr0 + pn + 1] - off[pr0]), 2, oracle.SCORING["MAX"], threads=os.cpu_count() or 1)
            for a_, b_ in zip(p[9], exp):
                ok_trim = ok_trim and bool((np.asarray(a_) == b_).all())
            n_tr += pn
        res = {"oracle_parity": bool(ok_keys and ok_cnt and ok_dir and ok_ext and ok_owner and ok_stats and ok_hist and ok_trim),
               "oracle_parity_detail": {"reads": n_total, "ranks": world, "distinct_kmers": int(len(o["count"])), "keys": ok_keys, "counts": ok_cnt,
                                        "direction_bias_pm1": ok_dir, "extension_counters": ok_ext if ext else None, "owner_rule": ok_owner,
                                        "owner_rule_keys_checked": n_owner, "stats": ok_stats, "histogram": ok_hist, "trim_results": ok_trim,
                                        "trim_reads_checked": n_tr}}
    return res


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner) must not write there
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import kmernator_b200 as KM
    from bench import synth

    wl = WORKLOADS[args.workload]
    k = wl["k"]
    kpr = READ_LEN - k + 1
    ext = bool(wl.get("ext"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input, generated in HBM (multi-GPU: each rank holds a contiguous slice of the reads) ----
    n_reads = args.reads or wl["reads"]
    meta = synth.metagenome_plan(seed=wl["genome_seed"], **wl["meta"]) if wl.get("meta") else None
    genome_len = (args.genome or wl["genome"]) * (world if wl["genome"] >= 1_000_000 else 1)
    if meta is not None:
        genome_len = int(meta[0].sum())
    bases, quals, off = synth.reads_torch(n_reads, READ_LEN, genome_len, seed=wl["genome_seed"], device=dev, read_seed=0x5245414453 + rank, meta=meta)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    off_u64 = off  # int64 bit pattern == uint64
    inst_per_rank = n_reads * kpr

    # ---- table / staging plan ----
    # distinct k-mers ~ genome covered + ~k error k-mers per substitution (C2: 0.7 G -> 1.4 G slots at load 0.5)
    if meta is not None:
        cov = meta[1] * n_reads * world * READ_LEN / meta[0]
        genomic = float((meta[0] * (1.0 - np.exp(-cov * kpr / READ_LEN))).sum())
    else:
        cov = n_reads * world * READ_LEN / max(1, genome_len)
        genomic = genome_len * (1.0 - np.exp(-cov * kpr / READ_LEN))
    err_kmers = n_reads * world * READ_LEN * 0.001 * min(k, 3 * genome_len)          # a tiny genome has few distinct error k-mers
    est_distinct = int((genomic * 1.02 + err_kmers * 1.05) / world) + (1 << 16)
    table_slots = args.table_slots or int(est_distinct / 0.5)
    # staging: one set = 1/pipe_batches of the input
    pipe_batches = args.pipe_batches or (2 if world == 1 else 8)
    stage_keys = args.stage_keys or int(inst_per_rank * 1.02 / pipe_batches)
    vk = KM.capi.KMN_VALUE_DIR_EXT if ext else KM.capi.KMN_VALUE_DIR
    ctx = KM.Context(kmer_size=k, est_raw_kmers=inst_per_rank, table_slots=table_slots, stage_keys=stage_keys, value_kind=vk,
                     min_quality_score=wl.get("min_quality_score", 3), min_kmer_quality=wl.get("min_kmer_quality", 0.10),
                     slice_bytes=args.slice_mb << 20, device=local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(KM.Context.comm_unique_id().copy())
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, uid.cpu().numpy())
    kstream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step_device():
        ctx.reset()
        ctx.count_batch(bases, quals, off_u64, n_reads=n_reads)
        ctx.count_finish(apply_purge=True)

    # ---- warm-up ----
    for _ in range(args.warmup):
        step_device()
    ctx.sync()
    ctx.profile_enable(True)
    ctx.profile_read()
    launches0 = ctx.launches

    # ---- timed region: K steps, CUDA events on the library stream, barrier + sync on both sides ----
    # inputs (30 GB) and table (>20 GB) are far larger than the 126 MB L2, and every step starts by clearing the table
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record(kstream)
        for _ in range(args.steps):
            step_device()
        ev1.record(kstream)
        ctx.sync()
        barrier()
    ms = ev0.elapsed_time(ev1)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.launches - launches0
    stats = ctx.stats()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = inst_per_rank * world / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # the kernels are timed inside a long step: prefer a sustained HBM figure, then the plain one, then any HBM key
    def _num(v):
        return float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) and v > 0 else None
    flat = {}
    for k_, v_ in (peaks.items() if isinstance(peaks, dict) else []):
        if isinstance(v_, dict):
            for k2, v2 in v_.items():
                flat["%s.%s" % (k_, k2)] = v2
        else:
            flat[k_] = v_
    cands = [k_ for k_ in flat if "hbm" in k_.lower() and _num(flat[k_])]
    cands.sort(key=lambda k_: (0 if "sustain" in k_.lower() else 1 if k_ == "hbm_gbs" else 2, k_))
    if cands:
        peak, peak_src = _num(flat[cands[0]]), "measured (MEASURED_PEAKS.json %s)" % cands[0]
        if peak < 100:                       # a TB/s figure
            peak *= 1000.0
    else:
        peak, peak_src = 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"
    zero = {"ms": 0.0, "launches": 0, "units": 0}
    ins, par, sub, wgt = prof.get("insert", zero), prof.get("parse", zero), prof.get("subpart", zero), prof.get("weight", zero)
    good_per_step = stats["raw_good_kmers"]              # instances actually inserted in the last step
    # The count pass is a pipeline of three kernels of similar cost; SURVEY.md 8d states its algorithmic traffic for the pass as
    # a whole (64 B of table read-modify-write + 2.5 B of input per instance = 66.5 B), which is `whole_pass` below and the
    # figure north_star's ">= 0.50" refers to.  The dominant kernel is whichever class takes the largest share of the step;
    # it is reported with the bytes IT has to move per launch (stated per kernel in DESIGN.md 3):
    #   k_kmer_scatter      1.25 B of bases + 1/8 B of "counted" bits read, 8 B of record written per instance = 9.375 B
    #   k_slice_split       8 B read + 8 B written per record = 16 B
    #   k_count_slices_*    8 B per record + one sweep of the table per drain (every 16-byte slot read and written = 32 B per slot)
    #   k_insert_staged     (k > 31 / weights / extension counters: no shared-memory slices) the 64 B per instance of 8d
    drains_per_step = max(1, -(-inst_per_rank // max(1, stage_keys)))
    ins_step_bytes = 8.0 * good_per_step + 32.0 * stats["table_slots"] * drains_per_step
    cls = {"insert": (ins, (ins_step_bytes / max(1, good_per_step)) if sub["launches"] else ALGO_BYTES_PER_INSERT,
                      "k_count_slices" if sub["launches"] else "k_insert_staged"),
           "subpart": (sub, 16.0, "k_slice_split"), "parse": (par, 9.375, "k_kmer_scatter")}
    dom = max(cls, key=lambda c_: cls[c_][0]["ms"])
    dk, dbytes, dname = cls[dom]
    d_ms_per_launch = dk["ms"] / max(1, dk["launches"])
    d_units_per_launch = good_per_step * args.steps / max(1, dk["launches"])
    achieved = dbytes * d_units_per_launch / (d_ms_per_launch * 1e-3) / 1e9 if dk["ms"] else 0.0
    traffic, traffic_src = None, None
    try:                                                 # dram bytes per launch of that kernel from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj.get("%s:%s" % (args.workload, dname))
        if ent and world == 1 and n_reads == wl["reads"] and not args.pipe_batches and not args.stage_keys:
            traffic, traffic_src = float(ent["dram_bytes_per_launch"]), ent["source"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dname, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
        "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": dbytes * d_units_per_launch, "peak_source": peak_src,
        "algorithmic_bytes_per_unit": dbytes, "units_per_launch": d_units_per_launch,
        "ms_per_launch": d_ms_per_launch, "launches": dk["launches"],
        "share_of_step": dk["ms"] / ms if ms else None,
        "kernel_classes_ms_per_step": {"weight": wgt["ms"] / args.steps, "parse": par["ms"] / args.steps, "subpart": sub["ms"] / args.steps,
                                       "insert": ins["ms"] / args.steps},
        "per_kernel": {n_: {"kernel": v_[2], "ms_per_step": v_[0]["ms"] / args.steps, "algorithmic_bytes_per_instance": v_[1],
                            "achieved": (v_[1] * good_per_step * args.steps / (v_[0]["ms"] * 1e-3) / 1e9) if v_[0]["ms"] else None,
                            "frac": (v_[1] * good_per_step * args.steps / (v_[0]["ms"] * 1e-3) / 1e9 / peak) if v_[0]["ms"] else None}
                       for n_, v_ in cls.items()},
        "whole_pass": {"algorithmic_bytes_per_instance": ALGO_BYTES_PER_INSTANCE, "definition": "SURVEY.md 8d: 64 B table RMW + 2.5 B input per presented instance",
                       "target_frac": 0.50,
                       "achieved": ALGO_BYTES_PER_INSTANCE * inst_per_rank / (ms_per_step * 1e-3) / 1e9,
                       "frac": ALGO_BYTES_PER_INSTANCE * inst_per_rank / (ms_per_step * 1e-3) / 1e9 / peak},
    }

    # ---- full-size parity properties (untimed): every counted instance is in the table exactly once ----
    checks = None
    if not args.no_checks:
        ctx.reset()
        ctx.count_batch(bases, quals, off_u64, n_reads=n_reads)
        ctx.count_finish(apply_purge=False)
        st0 = ctx.stats()
        hist = ctx.histogram().astype(np.float64)           # all-reduced over ranks; exact in fp64 below 2^53
        sum_counts = float((hist * np.arange(65536, dtype=np.float64)).sum())
        tot = torch.tensor([st0["raw_kmers"], st0["raw_good_kmers"], st0["unique_kmers"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        raw_all, good_all, uniq_all = (float(x) for x in tot.tolist())
        checks = {"instances_presented": raw_all, "instances_counted": good_all, "sum_of_table_counts": sum_counts,
                  "distinct_kmers": uniq_all, "histogram_entries": float(hist.sum()), "saturated_kmers": float(hist[65535]),
                  # exact when nothing saturates; with saturated k-mers (C5b) the table holds at most the counted instances
                  "sum_counts_equals_counted_instances": bool(sum_counts == good_all) if hist[65535] == 0 else None,
                  "sum_counts_le_counted_instances": bool(sum_counts <= good_all),
                  "histogram_entries_equal_distinct": bool(hist.sum() == uniq_all),
                  "presented_equals_reads_x_kmers": bool(raw_all == float(inst_per_rank) * world)}
        step_device()                                        # leave the purged table of a normal step for the lookup pass
        par_res = oracle_parity(KM, dist, torch, dev, rank, world, local_rank, args.parity_reads, wl)
        if par_res is not None:
            checks.update(par_res)

    # ---- multi-GPU: the exchange (records pushed to their owners over NVLink), against 900 GB/s per direction ----
    exchange = None
    rt = prof.get("route")
    if world > 1 and rt and rt["ms"] > 0 and rt["units"] > 0:
        gbps = rt["units"] / (rt["ms"] * 1e-3) / 1e9
        exchange = {"bytes_per_step_per_gpu": rt["units"] / args.steps, "copy_ms_per_step": rt["ms"] / args.steps, "GBps_while_copying": gbps,
                    "frac_of_nvlink_900GBps": gbps / 900.0, "share_of_step": rt["ms"] / ms,
                    "record_bytes_per_step_per_gpu": 8.0 * good_per_step * (world - 1) / world,
                    "note": "bytes_per_step_per_gpu is what the copy engines moved (sub-region capacity incl. slack); record_bytes is the payload"}

    # ---- lookup pass (FilterReads pass 2: per-read min-depth trim + score), device-resident reads ----
    lookup = None
    if not args.no_lookup:
        n_l = min(args.lookup_reads, n_reads)
        lb, lo = bases[: n_l * READ_LEN], off_u64[: n_l + 1]
        outs = (torch.empty(n_l, dtype=torch.int32, device=dev), torch.empty(n_l, dtype=torch.int32, device=dev),
                torch.empty(n_l, dtype=torch.float32, device=dev), torch.empty(n_l, dtype=torch.uint8, device=dev))
        ctx.trim_batch(lb, lo, 2, "MAX", n_reads=n_l, out=outs)            # warm-up (allocates the per-k-mer value buffer)
        ctx.sync()
        ctx.profile_enable(True)
        ctx.profile_read()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0.record(kstream)
        for _ in range(args.steps):
            ctx.trim_batch(lb, lo, 2, "MAX", n_reads=n_l, out=outs)
        l1.record(kstream)
        ctx.sync()
        barrier()
        lt = torch.tensor([l0.elapsed_time(l1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        lms = float(lt.item())
        lprof = ctx.profile_read()
        ctx.profile_enable(False)
        n_lk = n_l * kpr * world
        lookup = {"value": n_lk / (lms * 1e-3), "unit": "kmer lookups/s", "reads_per_gpu": n_l, "ms_per_pass": lms,
                  "algorithmic_bytes_per_lookup": 33.25, "achieved_GBps_per_gpu": 33.25 * n_lk / world / (lms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": 33.25 * n_lk / world / (lms * 1e-3) / 1e9 / peak,
                  "kept_reads_full_length": int((outs[1] == READ_LEN).sum().item()),
                  "profile_ms_per_pass": {k_: v["ms"] / args.steps for k_, v in lprof.items()}}
        if wl.get("meta") is not None:
            # C4: RANDOM normalisation to depth 100 on these scores (ReadSelector::chooseRead: rand() % s <= D), counter-based draws
            D = 100
            sc = outs[2].to(torch.float64)
            passing = sc >= 2
            s_ = sc[passing].clamp(min=1.0).to(torch.int64)
            gen = torch.Generator(device=dev)
            gen.manual_seed(12345)
            draws = torch.randint(0, 2 ** 32, s_.shape, device=dev, generator=gen, dtype=torch.int64)
            kept = ((s_ <= D) | ((draws % s_) <= D)).sum().item()
            expct = torch.where(s_ > D, (D + 1.0) / s_.to(torch.float64), torch.ones_like(s_, dtype=torch.float64))
            lookup["normalisation"] = {"max_depth": D, "passing_reads": int(passing.sum().item()), "kept": int(kept),
                                       "expected_kept": float(expct.sum().item()), "sigma": float((expct * (1 - expct)).sum().sqrt().item())}

    # ---- e2e: host (pinned) buffers through the C ABI, H2D of every batch + D2H of the counters inside the timing ----
    e2e = None
    if not args.no_e2e:
        # multi-GPU: a bounded sample per rank keeps the pinned host memory of N ranks on one box small
        n_e = min(args.e2e_reads or (n_reads if world == 1 else max(n_reads // world, args.e2e_batch)), n_reads)
        hb = torch.empty(n_e * READ_LEN, dtype=torch.uint8, pin_memory=True)
        hq = torch.empty(n_e * READ_LEN, dtype=torch.uint8, pin_memory=True)
        hb.copy_(bases[: n_e * READ_LEN])
        hq.copy_(quals[: n_e * READ_LEN])
        torch.cuda.synchronize()
        hbn, hqn = hb.numpy(), hq.numpy()
        bsz = args.e2e_batch
        hoff_t = torch.empty(bsz + 1, dtype=torch.int64, pin_memory=True)      # offsets are an input too: pinned like the bases
        hoff_t.copy_(torch.arange(bsz + 1, dtype=torch.int64) * READ_LEN)
        hoff = hoff_t.numpy()

        # the same reads in the reference's in-memory form: TwoBitSequence bytes (38 per 150-base read) instead of ASCII bases
        # (Read::_data, src/Sequence.h:372-380); packed on the GPU here, outside the timing, like a ReadSet loaded earlier
        pb = (READ_LEN + 3) // 4
        hp = torch.empty(n_e * pb, dtype=torch.uint8, pin_memory=True)
        lutc = torch.zeros(256, dtype=torch.uint8, device=dev)
        for ch, cv in ((ord("C"), 1), (ord("G"), 2), (ord("T"), 3)):
            lutc[ch] = cv
        for c0 in range(0, n_e, 2_000_000):
            n = min(2_000_000, n_e - c0)
            cd = lutc[bases[c0 * READ_LEN: (c0 + n) * READ_LEN].long()].reshape(n, READ_LEN)
            cd = torch.nn.functional.pad(cd, (0, pb * 4 - READ_LEN)).reshape(n, pb, 4)
            pk = (cd[:, :, 0] << 6) | (cd[:, :, 1] << 4) | (cd[:, :, 2] << 2) | cd[:, :, 3]
            hp[c0 * pb: (c0 + n) * pb].copy_(pk.reshape(-1))
            del cd, pk
        torch.cuda.synchronize()
        hpn = hp.numpy()
        hpoff_t = torch.empty(bsz + 1, dtype=torch.int64, pin_memory=True)
        hpoff_t.copy_(torch.arange(bsz + 1, dtype=torch.int64) * pb)
        hpoff = hpoff_t.numpy()

        # batches arrive over PCIe while earlier ones are counted: what is left when the last batch has landed is the drain of
        # the last staging set, so the end-to-end run wants SMALLER sets than the device-resident one (more drains, each a
        # sweep of the table, but hidden behind the copies).  One GPU: a fresh context with 1/8 of the input per set.
        e2e_pipe = pipe_batches
        if world == 1 and not args.stage_keys and args.e2e_pipe_batches > pipe_batches:
            e2e_pipe = args.e2e_pipe_batches
            ctx.close()
            ctx = KM.Context(kmer_size=k, est_raw_kmers=inst_per_rank, table_slots=table_slots, stage_keys=int(inst_per_rank * 1.02 / e2e_pipe),
                             value_kind=vk, min_quality_score=wl.get("min_quality_score", 3), min_kmer_quality=wl.get("min_kmer_quality", 0.10),
                             slice_bytes=args.slice_mb << 20, device=local_rank)

        def step_e2e(packed):
            ctx.reset()
            for r0 in range(0, n_e, bsz):
                nb = min(bsz, n_e - r0)
                if packed:
                    ctx.count_batch_2na(hpn[r0 * pb: (r0 + nb) * pb], hpoff[: nb + 1], hqn[r0 * READ_LEN: (r0 + nb) * READ_LEN], hoff[: nb + 1], n_reads=nb)
                else:
                    ctx.count_batch(hbn[r0 * READ_LEN: (r0 + nb) * READ_LEN], hqn[r0 * READ_LEN: (r0 + nb) * READ_LEN], hoff[: nb + 1], n_reads=nb)
            ctx.count_finish(apply_purge=True)
            return ctx.stats()      # D2H read of the spectrum counters

        def time_e2e(packed):
            for _ in range(min(2, args.warmup)):
                step_e2e(packed)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                st_ = step_e2e(packed)
            ctx.sync()
            barrier()
            dt_ = (time.perf_counter() - t0) / args.steps
            te = torch.tensor([dt_], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), st_

        note = None if n_e == n_reads else "SAMPLE: %d of %d reads per rank (pinned host memory of %d ranks on one box)" % (n_e, n_reads, world)
        n_off = (n_e // bsz + 1) * (bsz + 1) * 8
        dt, st_e = time_e2e(True)
        e2e = {"value": n_e * kpr * world / dt, "unit": "kmers/s", "h2d_bytes_per_step": n_e * (READ_LEN + pb) + 2 * n_off,
               "d2h_bytes_per_step": 64 + 48, "reads_per_step": n_e, "batch_reads": bsz, "ms_per_step": dt * 1e3,
               "raw_good_kmers": st_e["raw_good_kmers"], "entry": "kmn_count_batch_2na", "workload_note": note, "drains_per_step": e2e_pipe,
               "input_format": "host reads as the reference holds them in memory: TwoBitSequence bytes (4 bases per byte) + one quality byte per base"}
        dt, st_a = time_e2e(False)
        e2e["ascii"] = {"value": n_e * kpr * world / dt, "unit": "kmers/s", "h2d_bytes_per_step": 2 * n_e * READ_LEN + n_off, "ms_per_step": dt * 1e3,
                        "raw_good_kmers": st_a["raw_good_kmers"], "entry": "kmn_count_batch",
                        "input_format": "host reads as ASCII: one byte per base + one quality byte per base"}
        del hp
        del hb, hq

    # ---- CPU baseline on a bounded sample of the same workload (rank 0 only; same sample as --impl reference) ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        n_c = min(args.cpu_reads, n_reads)
        cb, cq, co = cpu_sample(wl, n_c, (args.genome or wl["genome"]))
        threads = os.cpu_count() or 1
        rate, dtc = cpu_port_rate(cb, cq, co, threads, k=k, min_quality=wl.get("min_quality_score", 3),
                                  min_kmer_quality=wl.get("min_kmer_quality", 0.10), track_ext=ext)
        cpu = {"value": rate, "unit": "kmers/s", "cores": threads, "kind": "port",
               "sample": "first %d reads of the workload's synthetic read set (numpy generator, full-size genome, same error/quality model; %.1f s of CPU work)" % (n_c, dtc),
               "note": "restated CPU baseline - reference binary not buildable here (no MPI/Boost)"}

    if rank == 0:
        line = {
            "metric": "k-mer insertions/sec (k=%d, 150bp reads)" % k, "value": value, "unit": "kmers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"] % n_reads, "workload_id": args.workload,
                       "reads_per_gpu": n_reads, "genome_bp": genome_len, "table_slots": stats["table_slots"],
                       "table_partitions": stats["table_partitions"], "stage_keys": stage_keys, "slice_mb": args.slice_mb,
                       "l2_policy": "inputs (%.0f GB) and table (%.0f GB) exceed the 126 MB L2; table cleared every step" % (2e-9 * n_reads * READ_LEN, 16e-9 * stats["table_slots"]),
                       "parallelism": "owner-sharded x%d" % world if world > 1 else "single GPU"},
            "roofline": roofline, "checks": checks, "exchange": exchange, "lookup_pass": lookup, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "stats": {k_: int(v) for k_, v in stats.items()},
            "profile_ms_per_step": {k_: v["ms"] / args.steps for k_, v in prof.items()},
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        # a rank that fails must not leave its peers waiting in a collective until the launcher's timeout
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
