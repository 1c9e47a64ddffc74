"""Golden vectors: decoded lists produced by the UNMODIFIED reference (tests/golden/make_golden.py,
run where /root/reference exists) on seeded synthetic reads.  The posteriors are regenerated here
from the seed and checked against the recorded sha256 before anything is compared."""
import hashlib
import json
import os

import numpy as np
import pytest

from nanopore_dna_storage_b200 import synth

GOLD = []
for _name in ("reference_lists.json", "reference_lists_r2.json"):  # make_golden.py, make_golden_r2.py
    GOLD += json.load(open(os.path.join(os.path.dirname(__file__), "golden", _name)))["cases"]


def _regen(case):
    kw = dict(case["workload"])
    w = synth.Workload(case["name"], sub=0.004, dele=0.0085, ins=0.0005, **kw)
    r = synth.make_read(w, case["read"])
    assert hashlib.sha256(r.post.tobytes()).hexdigest() == case["post_sha256"], \
        "synthetic posterior differs from the one the golden list was made from (numpy version?)"
    return w, r


def _bits(lists, msg_len):
    return np.array([[int(c) for c in s] for s in lists], dtype=np.uint8).reshape(len(lists), msg_len)


# CPU suite: every small-code case up to list size 16, plus one L = 32 and one L = 64 read (the oracle
# needs about half a minute for those)
SMALL = [c for c in GOLD if c["workload"]["mem_conv"] <= 8 and
         (c["workload"]["list_size"] <= 16 or (c["workload"]["list_size"] == 32 and c["read"] == 0) or c["rc"])]


@pytest.mark.parametrize("case", SMALL, ids=[f"{c['name']}#{c['read']}" for c in SMALL])
def test_oracle_reproduces_reference_lists(oracle, case):
    w, r = _regen(case)
    st, bits, _ = oracle.oracle_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, case["rc"])
    assert st == 0
    want = _bits(case["lists"], w.msg_len)
    assert bits.shape == want.shape and np.array_equal(bits, want)


@pytest.mark.gpu
def test_gpu_reproduces_reference_lists():
    """Every golden case through the C ABI: the full-size C2 bench reads (incl. beta = 3 hard reads and
    lengths that are multiples of 29), C3 at msg_len 100 and 180, C5 in both orientations, L = 32 / 64."""
    from nanopore_dna_storage_b200.decoder import ListDecoder
    groups = {}
    for c in GOLD:
        key = (json.dumps(c["workload"], sort_keys=True), c["rc"])
        groups.setdefault(key, []).append(c)
    for (_, rc), cases in groups.items():
        regen = [_regen(c) for c in cases]
        w = regen[0][0]
        with ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, rc=rc) as dec:
            got = dec.decode([r.post for _, r in regen])
        for c, (bits, _) in zip(cases, got):
            want = _bits(c["lists"], w.msg_len)
            assert bits.shape == want.shape and np.array_equal(bits, want), f"{c['name']} read {c['read']}"
