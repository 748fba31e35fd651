#!/usr/bin/env python
"""bench.py -- k-mer insertions/sec (k=31, 150 bp reads) on N B200, vs the CPU restatement of FilterReads-P.

Contract (one JSON line on rank 0):
  value    = whole-job k-mer insertions/s of the count pass (kmn_reset + kmn_count_batch + kmn_count_finish incl. the
             post-build min-depth purge) with the reads already resident in HBM, CUDA events on the library's stream,
             max over ranks.
  e2e      = same metric through the C ABI with HOST (pinned) buffers: H2D of every batch and a D2H read of the
             spectrum counters inside the timed region.
  roofline = dominant kernel (k_insert_staged): 64 algorithmic bytes per staged instance (32-B sector read + 32-B
             write-back of a 16-B slot RMW) / its measured launch time vs MEASURED_PEAKS.json hbm_gbs.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU (config C2: 100M x 150 bp)")
    ap.add_argument("--genome", type=int, default=250_000_000, help="genome length per GPU-worth of reads (60x coverage)")
    ap.add_argument("--table-slots", type=int, default=0)
    ap.add_argument("--stage-keys", type=int, default=0)
    ap.add_argument("--slice-mb", type=int, default=64)
    ap.add_argument("--pipe-batches", type=int, default=8, help="sub-batches of the parse/insert pipeline per step")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per e2e step (0 = same as --reads)")
    ap.add_argument("--e2e-batch", type=int, default=4_000_000)
    ap.add_argument("--cpu-reads", type=int, default=400_000)
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


def cpu_port_rate(bases_np, quals_np, off_np, threads, steps=1, warmup=0):
    """oracle port of the count pass (serial/OpenMP T x T build), k-mer instances / s"""
    import oracle
    n_reads = len(off_np) - 1
    inst = n_reads * KMERS_PER_READ
    times = []
    for it in range(warmup + steps):
        s = oracle.OracleSpectrum(K, threads=threads, est_distinct=max(1 << 16, inst // 3))
        t0 = time.perf_counter()
        s.add_reads(bases_np.tobytes() if not isinstance(bases_np, bytes) else bases_np, quals_np, off_np)
        s.purge_min_depth(2)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        del s
    return inst / (sum(times) / len(times)), sum(times) / len(times)


def run_reference(args, rank, world, emit):
    """--impl reference: the CPU restatement of the reference path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from bench import synth
    threads = os.cpu_count() or 1
    n = args.cpu_reads
    bases, quals, off = synth.reads_numpy(n, READ_LEN, max(10_000, int(n * READ_LEN / 60)), seed=0x5245)
    rate, dt = cpu_port_rate(bases, quals, off, threads, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    sample = "%d synthetic 150bp reads (60x of a %d bp genome, same error/quality model) per step" % (n, max(10_000, int(n * READ_LEN / 60)))
    line = {
        "impl": "reference", "metric": "k-mer insertions/sec (k=31, 150bp reads)", "value": rate, "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "C2: k=31 count + min-depth-2 purge, synthetic 150bp reads (bounded CPU sample)", "reads_per_step": n},
        "cpu_baseline": {"value": rate, "unit": "kmers/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "restated CPU baseline - reference binary not buildable here (no MPI/Boost)"},
        "e2e": {"value": rate, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input, generated in HBM (config C2; multi-GPU: each rank holds a contiguous slice of the reads) ----
    n_reads = args.reads
    genome_len = args.genome * world
    bases, quals, off = synth.reads_torch(n_reads, READ_LEN, genome_len, seed=0x4B6D6572, device=dev, read_seed=0x5245414453 + rank)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    off_u64 = off  # int64 bit pattern == uint64
    inst_per_rank = n_reads * KMERS_PER_READ

    # ---- table / staging plan ----
    # distinct k-mers ~ genome (coverage 60x) + ~30 error k-mers per substitution: 0.7 G for C2 -> 1.4 G slots at load 0.5
    est_distinct = int(genome_len / world * 1.02 + n_reads * READ_LEN * 0.001 * 31 * 1.05)
    table_slots = args.table_slots or int(est_distinct / 0.5)
    # staging: two sets of stage_keys records; one set = one sub-batch of the parse/insert pipeline (default: 1/8 of the input)
    stage_keys = args.stage_keys or int(inst_per_rank * 1.02 / args.pipe_batches)
    ctx = KM.Context(kmer_size=K, est_raw_kmers=inst_per_rank, table_slots=table_slots, stage_keys=stage_keys,
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

    # ---- roofline of the dominant kernel (phase-2 insert) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # the phase-2 kernel is timed inside a long step: prefer a sustained HBM figure, then the plain one, then any HBM key
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
    ins = prof.get("insert", {"ms": 0.0, "launches": 0, "units": 0})
    par = prof.get("parse", {"ms": 0.0, "launches": 0, "units": 0})
    good_per_step = stats["raw_good_kmers"]              # instances actually inserted in the last step
    ins_ms_per_launch = ins["ms"] / max(1, ins["launches"])
    ins_units_per_launch = good_per_step * args.steps / max(1, ins["launches"])
    achieved = ALGO_BYTES_PER_INSERT * ins_units_per_launch / (ins_ms_per_launch * 1e-3) / 1e9 if ins["ms"] else 0.0
    roofline = {
        "bound": "hbm", "kernel": "k_insert_staged", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one k_insert_staged launch of this workload (1.477 G records),
        # ncu --set full, profiles/r01ah_ncu_insert_c2_selected.csv: 31.0 GB read + 9.35 GB written.  Far below the
        # algorithmic 64 B per record because the table group being filled stays L2-resident (the design's point).
        "traffic": 40.35e9 if (world == 1 and n_reads == 100_000_000) else None, "traffic_unit": "bytes per launch",
        "algorithmic_bytes_per_launch": ALGO_BYTES_PER_INSERT * ins_units_per_launch, "peak_source": peak_src,
        "algorithmic_bytes_per_unit": ALGO_BYTES_PER_INSERT, "units_per_launch": ins_units_per_launch,
        "ms_per_launch": ins_ms_per_launch, "launches": ins["launches"],
        "share_of_step": ins["ms"] / ms if ms else None,
        "parse_kernel": {"ms_per_step": par["ms"] / args.steps, "share_of_step": par["ms"] / ms if ms else None,
                         "GBps_input": (2.0 * n_reads * READ_LEN * args.steps) / (par["ms"] * 1e-3) / 1e9 if par["ms"] else None},
        "whole_pass": {"achieved": ALGO_BYTES_PER_INSTANCE * inst_per_rank / (ms_per_step * 1e-3) / 1e9,
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
                  "distinct_kmers": uniq_all, "histogram_entries": float(hist.sum()),
                  "sum_counts_equals_counted_instances": bool(sum_counts == good_all and hist[65535] == 0),
                  "histogram_entries_equal_distinct": bool(hist.sum() == uniq_all),
                  "presented_equals_reads_x_kmers": bool(raw_all == float(inst_per_rank) * world)}
        step_device()                                        # leave the purged table of a normal step for the lookup pass

    # ---- multi-GPU: the exchange (records pushed to their owners over NVLink), against 900 GB/s per direction ----
    exchange = None
    rt = prof.get("route")
    if world > 1 and rt and rt["ms"] > 0 and rt["units"] > 0:
        gbps = rt["units"] / (rt["ms"] * 1e-3) / 1e9
        exchange = {"bytes_per_step_per_gpu": rt["units"] / args.steps, "copy_ms_per_step": rt["ms"] / args.steps, "GBps_while_copying": gbps,
                    "frac_of_nvlink_900GBps": gbps / 900.0, "share_of_step": rt["ms"] / ms,
                    "note": "copy-engine pushes of whole staging parts (capacity, incl. slack), overlapped with phase 1/2 on other streams"}

    # ---- lookup pass (FilterReads pass 2: per-read min-depth trim + score), device-resident reads, single GPU only ----
    lookup = None
    if world == 1 and not args.no_lookup:
        n_l = min(args.lookup_reads, n_reads)
        lb, lo = bases[: n_l * READ_LEN], off_u64[: n_l + 1]
        outs = (torch.empty(n_l, dtype=torch.int32, device=dev), torch.empty(n_l, dtype=torch.int32, device=dev),
                torch.empty(n_l, dtype=torch.float32, device=dev), torch.empty(n_l, dtype=torch.uint8, device=dev))
        ctx.trim_batch(lb, lo, 2, "MAX", n_reads=n_l, out=outs)            # warm-up (allocates the per-k-mer value buffer)
        ctx.sync()
        ctx.profile_enable(True)
        ctx.profile_read()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(kstream)
        for _ in range(args.steps):
            ctx.trim_batch(lb, lo, 2, "MAX", n_reads=n_l, out=outs)
        l1.record(kstream)
        ctx.sync()
        lms = l0.elapsed_time(l1) / args.steps
        lprof = ctx.profile_read()
        ctx.profile_enable(False)
        n_lk = n_l * KMERS_PER_READ
        lookup = {"value": n_lk / (lms * 1e-3), "unit": "kmer lookups/s", "reads": n_l, "ms_per_pass": lms,
                  "algorithmic_bytes_per_lookup": 33.25, "achieved_GBps": 33.25 * n_lk / (lms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": 33.25 * n_lk / (lms * 1e-3) / 1e9 / peak,
                  # scattered 16-byte loads from a 16 GiB table on this GPU (profiles/r01_randacc_microbench.csv): 18.3 G/s
                  "x_measured_random_load_rate": n_lk / (lms * 1e-3) / 18.3e9,
                  "kept_reads_full_length": int((outs[1] == READ_LEN).sum().item()),
                  "profile_ms_per_pass": {k: v["ms"] / args.steps for k, v in lprof.items()}}

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

        def step_e2e():
            ctx.reset()
            for r0 in range(0, n_e, bsz):
                nb = min(bsz, n_e - r0)
                ctx.count_batch(hbn[r0 * READ_LEN: (r0 + nb) * READ_LEN], hqn[r0 * READ_LEN: (r0 + nb) * READ_LEN], hoff[: nb + 1], n_reads=nb)
            ctx.count_finish(apply_purge=True)
            return ctx.stats()      # D2H read of the spectrum counters

        for _ in range(min(2, args.warmup)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st_e = step_e2e()
        ctx.sync()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": n_e * KMERS_PER_READ * world / dt, "unit": "kmers/s", "h2d_bytes_per_step": 2 * n_e * READ_LEN + (n_e // bsz + 1) * (bsz + 1) * 8,
               "d2h_bytes_per_step": 64 + 48, "reads_per_step": n_e, "batch_reads": bsz, "ms_per_step": dt * 1e3,
               "raw_good_kmers": st_e["raw_good_kmers"]}
        del hb, hq

    # ---- CPU baseline on a bounded sample of the same workload (rank 0 only) ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        n_c = min(args.cpu_reads, n_reads)
        cb = bases[: n_c * READ_LEN].cpu().numpy()
        cq = quals[: n_c * READ_LEN].cpu().numpy()
        co = (np.arange(n_c + 1, dtype=np.uint64) * READ_LEN)
        threads = os.cpu_count() or 1
        rate, dtc = cpu_port_rate(cb, cq, co, threads)
        cpu = {"value": rate, "unit": "kmers/s", "cores": threads, "kind": "port",
               "sample": "first %d reads of the same synthetic input (%.1f s of CPU work)" % (n_c, dtc),
               "note": "restated CPU baseline - reference binary not buildable here (no MPI/Boost)"}

    if rank == 0:
        line = {
            "metric": "k-mer insertions/sec (k=31, 150bp reads)", "value": value, "unit": "kmers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C2: k=31 count + min-depth-2 purge on %d synthetic 150bp reads per GPU (60x coverage, 0.1%% substitutions)" % n_reads,
                       "reads_per_gpu": n_reads, "genome_bp": genome_len, "table_slots": stats["table_slots"],
                       "table_partitions": stats["table_partitions"], "stage_keys": stage_keys, "slice_mb": args.slice_mb,
                       "l2_policy": "inputs (30 GB) and table (>20 GB) exceed the 126 MB L2; table cleared every step",
                       "parallelism": "owner-sharded x%d" % world if world > 1 else "single GPU"},
            "roofline": roofline, "checks": checks, "exchange": exchange, "lookup_pass": lookup, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "stats": {k: int(v) for k, v in stats.items()},
            "profile_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
