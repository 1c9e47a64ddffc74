#!/usr/bin/env python
"""Developer probe: decode N reads of a workload once or twice through the C ABI and print the
library's own timing record (no torch, no CPU baseline).  usage: quick.py [workload] [reads] [slots] [reps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanopore_dna_storage_b200 import synth  # noqa: E402
from nanopore_dna_storage_b200.decoder import ListDecoder  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
slots = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
w = synth.workloads()[name]
t0 = time.time()
reads = synth.make_reads(w, n)
print(f"{n} reads of {w.name} generated in {time.time() - t0:.1f}s", flush=True)
with ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, max_slots=slots) as dec:
    work = [dec.work(int(r.post.shape[0])) for r in reads]
    for rep in range(reps):
        t0 = time.time()
        got = dec.decode([r.post for r in reads])
        dt = time.time() - t0
        tm = dec.timing()
        ok = sum(1 for r, (b, s) in zip(reads, got) if b.shape[0] and np.array_equal(b[0], r.msg))
        gb = sum(x["state_bytes"] for x in work) / 1e9
        print(json.dumps(dict(rep=rep, wall_s=round(dt, 3), reads_per_s=round(n / dt, 2), top1_ok=ok, slots=dec.slots,
                              alg_GB=round(gb, 2), alg_GBps=round(gb / (tm["kernel_ms"] * 1e-3), 1), **tm)), flush=True)
