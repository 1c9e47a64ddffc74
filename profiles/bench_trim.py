#!/usr/bin/env python
"""Measurement of the posterior-trimming path: reads/s of lva_trim_batch (host
buffers in, host results out, through the C ABI) next to the CPU restatement of the reference's
helper (oracle/trim_oracle.py, pure Python like the reference) on a bounded sample.
usage: python profiles/bench_trim.py [n_reads]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanopore_dna_storage_b200 import trim
from oracle import trim_oracle as to

START, END = "CACCTGTGCTGCGTCAGGCTGTGTC", "GCTGTCCGTTCCGCATTGACACGGC"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(1)
seqs, transs = [], []
for i in range(n):
    body = START + "".join(rng.choice(list("ACGT"), 217)) + END   # 217 bases = C2 oligo (nstate_pos - 1 + ...)
    if i & 1:
        body = to.reverse_complement(body)
    s = "".join(rng.choice(list("ACGT"), int(rng.integers(5, 40)))) + body + "".join(rng.choice(list("ACGT"), int(rng.integers(5, 40))))
    seqs.append(s)
    transs.append(np.cumsum(1 + rng.poisson(3.5, len(s))).astype(np.uint32))
trim.find_barcodes_and_orientation(seqs[:64], transs[:64], START, END, 11, 180)  # warm-up (context, module load)
t0 = time.perf_counter()
got = trim.find_barcodes_and_orientation(seqs, transs, START, END, 11, 180)
gpu_s = time.perf_counter() - t0
m = min(n, 40)
t0 = time.perf_counter()
want = [to.trim_read(s, [int(x) for x in t], START, END, 11, 180) for s, t in zip(seqs[:m], transs[:m])]
cpu_s = time.perf_counter() - t0
assert all((g.start_pos, g.end_pos, g.rc) == (w[0], w[1], w[2]) for g, w in zip(got, want))
print(json.dumps({"metric": "reads trimmed/s (barcode search + orientation, host to host, includes packing the batch in Python)",
                  "value": n / gpu_s, "n_reads": n, "ok_frac": sum(g.status == 0 for g in got) / n,
                  "cpu_baseline": {"value": m / cpu_s, "unit": "reads/s", "cores": 1, "kind": "port",
                                   "sample": f"{m} reads, pure-Python restatement of helper.py:157-209 (the reference is pure Python too)"}}))
