#!/usr/bin/env python
"""Generates tests/golden/reference_lists.json by running the UNMODIFIED reference (oracle/_ref,
compiled from /root/reference in place) on seeded synthetic reads.  Posteriors are not stored:
they are regenerated from (workload, read index) at test time and checked against the sha256
recorded here, so the fixture stays a few KB.   usage: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nanopore_dna_storage_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

CASES = [
    # name, workload kwargs, read indices, rc
    ("C1 m6 r1/2 msg180 L1 dev20", dict(mem_conv=6, rate=1, msg_len=180, list_size=1, max_deviation=20), [0, 1, 2, 3], False),
    ("C1 m6 r1/2 msg180 L1 full band", dict(mem_conv=6, rate=1, msg_len=180, list_size=1, max_deviation=None), [0], False),
    ("m6 r1/2 msg180 L8 dev20 rc", dict(mem_conv=6, rate=1, msg_len=180, list_size=8, max_deviation=20), [4, 5], True),
    ("C4 m8 r2/3 msg180 L8", dict(mem_conv=8, rate=2, msg_len=180, list_size=8, max_deviation=20), [0, 1], False),
    ("C4 m8 r3/4 msg181 L8", dict(mem_conv=8, rate=3, msg_len=181, list_size=8, max_deviation=20), [0], False),
    ("C4 m8 r4/5 msg180 L8 rc", dict(mem_conv=8, rate=4, msg_len=180, list_size=8, max_deviation=20), [0], True),
    ("C4 m8 r7/8 msg181 L8", dict(mem_conv=8, rate=7, msg_len=181, list_size=8, max_deviation=20), [0], False),
    ("C2 m11 r5/6 msg180 L8 dev20", dict(mem_conv=11, rate=5, msg_len=180, list_size=8, max_deviation=20), [0, 1], False),
    ("C2 m11 r5/6 msg180 L8 dev20 rc", dict(mem_conv=11, rate=5, msg_len=180, list_size=8, max_deviation=20), [2], True),
    ("C5 m11 r5/6 msg180 L16 dev20", dict(mem_conv=11, rate=5, msg_len=180, list_size=16, max_deviation=20), [0], False),
    ("C3 m14 r1/2 msg40 L8 dev20", dict(mem_conv=14, rate=1, msg_len=40, list_size=8, max_deviation=20), [0], False),
    ("m8 r5/6 msg180 L8 quantised ties", dict(mem_conv=8, rate=5, msg_len=180, list_size=8, max_deviation=20, quantum=0.5), [0], False),
]


def main():
    out = []
    for name, kw, idxs, rc in CASES:
        w = synth.Workload(name, sub=0.004, dele=0.0085, ins=0.0005, seed=0x601D0000 + len(out), **kw)
        for i in idxs:
            r = synth.make_read(w, i)
            if rc:  # decode the reverse-complement strand: regenerate the posterior from rc bases
                w_rc = synth.Workload(name, sub=0.004, dele=0.0085, ins=0.0005, seed=w.seed, rc_fraction=1.0, **kw)
                r = synth.make_read(w_rc, i)
            t0 = time.time()
            st, bits = ob.ref_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, rc)
            assert st == 0
            out.append(dict(name=name, workload=dict(kw, seed=w.seed, rc_fraction=1.0 if rc else 0.0),
                            read=i, rc=rc, T=int(r.post.shape[0]),
                            post_sha256=hashlib.sha256(r.post.tobytes()).hexdigest(),
                            msg="".join(map(str, r.msg.tolist())),
                            lists=["".join(map(str, b.tolist())) for b in bits]))
            print(f"{name} read {i}: {len(bits)} entries, top1 correct={len(bits) > 0 and (bits[0] == r.msg).all()}, "
                  f"{time.time() - t0:.1f}s", flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_lists.json")
    with open(path, "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", reference="oracle/_ref/libvcc_ref.so "
                       "(unmodified /root/reference/viterbi/viterbi_convolutional_code.cpp, -O3 -march=x86-64-v3 -fopenmp)",
                       cases=out), f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
