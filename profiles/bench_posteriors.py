#!/usr/bin/env python
"""Measurement of the flip-flop posterior path: blocks/s and reads/s of
lva_posteriors_batch (host buffers in and out through the C ABI) next to flappie's own function
(oracle/_ref/libflipflop_ref.so when present, else the C restatement) on one host core.
usage: python profiles/bench_posteriors.py [n_reads] [blocks_per_read]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanopore_dna_storage_b200 import posteriors

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
rng = np.random.default_rng(2)
reads = [rng.normal(0, 2, (T, 40)).astype(np.float32) for _ in range(n)]
posteriors.transpost_crf_flipflop(reads[:8])
t0 = time.perf_counter()
got = posteriors.transpost_crf_flipflop(reads)
gpu_s = time.perf_counter() - t0
ref = os.path.join(ROOT, "oracle", "_ref", "libflipflop_ref.so")
kind = "reference" if os.path.exists(ref) else "port"
L = C.CDLL(ref if kind == "reference" else os.path.join(ROOT, "oracle", "_build", "libflipflop_oracle.so"))
fn = L.ffr_transpost if kind == "reference" else L.ffo_transpost
fn.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
m = min(n, 64)
out = np.zeros((T, 40), np.float32)
t0 = time.perf_counter()
for r in reads[:m]:
    fn(r.ctypes.data, T, out.ctypes.data)
cpu_s = time.perf_counter() - t0
print(json.dumps({"metric": "reads/s through lva_posteriors_batch (host to host)", "value": n / gpu_s,
                  "blocks_per_s": n * T / gpu_s, "n_reads": n, "blocks_per_read": T,
                  "max_abs_diff_last_read": float(np.abs(got[m - 1] - out).max()),
                  "cpu_baseline": {"value": m / cpu_s, "unit": "reads/s", "cores": 1, "kind": kind,
                                   "sample": f"{m} reads x {T} blocks, transpost_crf_flipflop on one core"}}))
