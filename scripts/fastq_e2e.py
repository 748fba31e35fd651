"""end-to-end wall time of the C++ FilterReads driver on a synthetic FASTQ file (host parse + count + lookup + write):
python scripts/fastq_e2e.py [n_reads]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = 150
bases, quals, off = synth.reads_numpy(n, L, 5_000_000, seed=11, err=0.001, lowq=0.0005)
path = "/tmp/kmn_e2e.fastq"
t0 = time.time()
b2 = bases.reshape(n, L)
q2 = quals.reshape(n, L)
with open(path, "wb") as f:
    for c0 in range(0, n, 100_000):
        c1 = min(n, c0 + 100_000)
        rows = []
        for i in range(c0, c1):
            rows.append(b"@r%d\n" % i + b2[i].tobytes() + b"\n+\n" + q2[i].tobytes() + b"\n")
        f.write(b"".join(rows))
size = os.path.getsize(path)
exe = os.path.join(ROOT, "kmernator_b200", "host", "bin", "FilterReads")
t1 = time.time()
p = subprocess.run([exe, "--skip-artifact-filter", "1", "--min-depth", "2", "--out", "/tmp/kmn_e2e_out", "31", path], capture_output=True, text=True)
dt = time.time() - t1
out_size = sum(os.path.getsize(os.path.join("/tmp", x)) for x in os.listdir("/tmp") if x.startswith("kmn_e2e_out"))
print(json.dumps({"reads": n, "fastq_bytes": size, "write_fastq_s": t1 - t0, "filter_reads_wall_s": dt, "rc": p.returncode,
                  "reads_per_s": n / dt, "kmers_per_s": n * (L - 30) / dt, "output_bytes": out_size, "stderr_tail": p.stderr[-1500:]}))
