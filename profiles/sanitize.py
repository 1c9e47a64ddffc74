#!/usr/bin/env python
"""Developer probe for compute-sanitizer: a few small decodes through the fast kernel (two-stage merge, pop-by-pop rows,
tie fallback) and the wide kernel.  usage: compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/sanitize.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanopore_dna_storage_b200 import synth  # noqa: E402
from nanopore_dna_storage_b200.decoder import ListDecoder  # noqa: E402

for (mem, rate, msg_len, L, q) in [(8, 5, 42, 8, 0.0), (6, 1, 40, 8, 0.5), (6, 5, 40, 24, 0.0), (6, 1, 40, 4, 0.0)]:
    w = synth.Workload("s", mem, rate, msg_len, L, 8, sub=0.01, dele=0.01, ins=0.005, seed=0x5A17 + mem + L, quantum=q)
    reads = synth.make_reads(w, 2)
    with ListDecoder(mem, rate, msg_len, L, 8) as dec:
        got = dec.decode([r.post for r in reads])
        tm = dec.timing()
    print(mem, rate, msg_len, L, q, [int(b.shape[0]) for b, s in got], "fallbacks", tm["exact_fallbacks"], flush=True)
