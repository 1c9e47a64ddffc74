#!/usr/bin/env python
"""Second batch of golden vectors (tests/golden/reference_lists_r2.json), same record format as
make_golden.py: decoded lists of the UNMODIFIED reference (oracle/_ref, compiled in place from
/root/reference) on seeded synthetic reads.  This batch covers what the first one left open:

  * the bench workload itself: reads 0..63 of synth.workloads()["C2"] -- the whole batch bench.py
    decodes on one GPU (a quarter of them beta = 3 "hard" reads) -- plus three reads whose length
    is a multiple of 29 (fused-band hazard, T = 493 / 522); bench.py compares its GPU output with
    these outside the timed region;
  * C3 at the paper's msg_len 100 (m = 14, 16384 x 8 states per row) and one read at msg_len 180;
  * C5 (L = 16) in BOTH orientations at full size;
  * list sizes above 16 (L = 32, L = 64 at m = 8, the paper's list size for m = 8).

Records are appended as they finish, so an interrupted run keeps what it has.
usage: python tests/golden/make_golden_r2.py [first_case_index]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nanopore_dna_storage_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

CH = dict(sub=0.004, dele=0.0085, ins=0.0005)
C2 = dict(mem_conv=11, rate=5, msg_len=180, list_size=8, max_deviation=20, seed=0x5EED0002, rc_fraction=0.0)
C5 = dict(mem_conv=11, rate=5, msg_len=180, list_size=16, max_deviation=20, seed=0x5EED0008, rc_fraction=0.5)
C3 = dict(mem_conv=14, rate=1, msg_len=100, list_size=8, max_deviation=20, seed=0x5EED0003, rc_fraction=0.0)
C3L = dict(mem_conv=14, rate=1, msg_len=180, list_size=8, max_deviation=20, seed=0x5EED0013, rc_fraction=0.0)
M8L32 = dict(mem_conv=8, rate=2, msg_len=180, list_size=32, max_deviation=20, seed=0x601D0020, rc_fraction=0.0)
M8L64 = dict(mem_conv=8, rate=2, msg_len=180, list_size=64, max_deviation=20, seed=0x601D0021, rc_fraction=0.0)
M8L64RC = dict(mem_conv=8, rate=5, msg_len=180, list_size=64, max_deviation=20, seed=0x601D0022, rc_fraction=1.0)

CASES = [
    ("r2 m8 r2/3 msg180 L32", M8L32, [0, 2]),
    ("r2 m8 r2/3 msg180 L64", M8L64, [0, 2]),
    ("r2 m8 r5/6 msg180 L64 rc", M8L64RC, [2]),
    ("r2 C5 m11 r5/6 msg180 L16 fwd+rc", C5, [0, 1, 2]),
    ("r2 C2 bench reads", C2, list(range(16)) + [169, 214, 342]),
    ("r2 C2 bench reads", C2, list(range(16, 64))),   # the rest of rank 0's bench batch: every benchmarked read has a fixture
    ("r2 C3 m14 r1/2 msg100 L8", C3, [0, 1]),
    ("r2 C3 m14 r1/2 msg180 L8", C3L, [0]),
]


def main():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_lists_r2.json")
    head = dict(generator="tests/golden/make_golden_r2.py", reference="oracle/_ref/libvcc_ref.so "
                "(unmodified /root/reference/viterbi/viterbi_convolutional_code.cpp, -O3 -march=x86-64-v3 -fopenmp)")
    out = []
    if os.path.exists(path):
        out = json.load(open(path))["cases"]
    done = {(c["name"], c["read"]) for c in out}
    for name, kw, idxs in CASES:
        w = synth.Workload(name, **CH, **kw)
        for i in idxs:
            if (name, i) in done:
                continue
            r = synth.make_read(w, i)
            t0 = time.time()
            st, bits = ob.ref_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, r.rc)
            assert st == 0
            out.append(dict(name=name, workload=dict(kw), read=i, rc=bool(r.rc), T=int(r.post.shape[0]),
                            beta=r.beta, post_sha256=hashlib.sha256(r.post.tobytes()).hexdigest(),
                            msg="".join(map(str, r.msg.tolist())),
                            lists=["".join(map(str, b.tolist())) for b in bits]))
            print(f"{name} read {i} (T={r.post.shape[0]}, beta={r.beta}, rc={r.rc}): {len(bits)} entries, "
                  f"top1 correct={len(bits) > 0 and (bits[0] == r.msg).all()}, {time.time() - t0:.1f}s", flush=True)
            with open(path, "w") as f:
                json.dump(dict(head, cases=out), f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
