"""Pins the CPU oracle (oracle/lva_oracle.cpp) against the UNMODIFIED reference compiled in place
(oracle/_ref/libvcc_ref.so) on the hazard corpus of SURVEY.md §8c.  Runs wherever oracle/_ref
exists (this container; the GPU box via the shipped .so)."""
import numpy as np
import pytest

from nanopore_dna_storage_b200 import synth
from oracle import binding as ob

pytestmark = pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _same(post, mem, rate, msg_len, L, dev, rc, **kw):
    so, bo, _ = ob.oracle_decode(post, mem, rate, msg_len, L, dev, rc, **kw)
    sr, br = ob.ref_decode(post, mem, rate, msg_len, L, dev, rc, **kw)
    assert so == sr == 0
    assert bo.shape == br.shape and np.array_equal(bo, br)
    return bo


def _reads(mem, rate, msg_len, L, dev, n, seed, **kw):
    w = synth.Workload("t", mem, rate, msg_len, L, dev, sub=0.01, dele=0.01, ins=0.005, seed=seed, **kw)
    return w, [synth.make_read(w, i) for i in range(n)]


# (1) every (m, r) that the tool accepts, forward and rc, banded
@pytest.mark.parametrize("mem", [6, 8])
@pytest.mark.parametrize("rate", [1, 2, 3, 4, 5, 7])
def test_every_code_and_rate(mem, rate):
    msg_len = next(m for m in (40, 41, 42) if synth.positions(mem, rate, m))
    w, reads = _reads(mem, rate, msg_len, 4, 20, 2, 100 + mem * 10 + rate)
    for r in reads:
        for rc in (False, True):
            _same(r.post, mem, rate, msg_len, 4, 20, rc)


@pytest.mark.parametrize("mem,msg_len", [(11, 24), (14, 12)])
def test_large_memory_codes(mem, msg_len):
    rate = 5 if mem == 11 else 1
    assert synth.positions(mem, rate, msg_len)
    w, reads = _reads(mem, rate, msg_len, 2, 10, 1, 300 + mem)
    _same(reads[0].post, mem, rate, msg_len, 2, 10, False)


# (3) band vs default, incl. hard / undecodable reads (stale rows outside the band are not cleared)
@pytest.mark.parametrize("dev", [None, 20, 6, 3])
def test_band_and_stale_rows(dev):
    w, reads = _reads(6, 5, 40, 4, dev, 3, 400, betas=(2.0,), sigma=2.0)
    for r in reads:
        _same(r.post, 6, 5, 40, 4, dev, False)


# (4) list sizes
@pytest.mark.parametrize("L", [1, 2, 3, 4, 8, 16])
def test_list_sizes(L):
    w, reads = _reads(6, 1, 40, L, 20, 2, 500 + L)
    for r in reads:
        _same(r.post, 6, 1, 40, L, 20, False)


# (5) quantised posteriors: massive exact score ties
def test_quantised_posteriors():
    w, reads = _reads(6, 5, 40, 8, 20, 4, 600, quantum=0.5, betas=(3.0, 2.0))
    for r in reads:
        _same(r.post, 6, 5, 40, 8, 20, r.rc)
    w, reads = _reads(8, 2, 40, 8, None, 2, 601, quantum=1.0, betas=(2.0,))
    for r in reads:
        _same(r.post, 8, 2, 40, 8, None, False)


# (6) shortest legal posterior and one block too few
def test_minimum_length_and_too_small():
    P = synth.positions(6, 1, 40)[0]
    rng = np.random.default_rng(7)
    post = np.log(rng.dirichlet(np.ones(40), size=P + 1)).astype(np.float32)
    _same(post, 6, 1, 40, 4, 20, False)
    so, _, _ = ob.oracle_decode(post[:P], 6, 1, 40, 4, 20, False)
    sr, _ = ob.ref_decode(post[:P], 6, 1, 40, 4, 20, False)
    assert so == 21 and sr == 2  # "Too small post matrix" (VCC:600-601) in both


# (7) sync markers
def test_sync_marker():
    w, reads = _reads(6, 1, 40, 4, 20, 2, 700)
    for r in reads:
        _same(r.post, 6, 1, 40, 4, 20, False, sync_marker="110", sync_period=9)
        _same(r.post, 6, 1, 40, 4, 20, True, sync_marker="01", sync_period=5)


# (8) empty list: no finite path reaches the end state (random posteriors + a marker nobody sent)
def test_empty_or_short_lists_agree():
    P = synth.positions(6, 1, 40)[0]
    rng = np.random.default_rng(11)
    post = np.full((P + 1, 40), -np.inf, dtype=np.float32)
    post[:, 0] = 0.0  # only A+ -> A+ is possible: the trellis can never advance
    b = _same(post, 6, 1, 40, 4, 20, False)
    assert b.shape[0] == 0


# (10) thread count does not change the reference's output
def test_reference_is_thread_invariant():
    w, reads = _reads(6, 5, 40, 8, 20, 1, 800)
    _, b1 = ob.ref_decode(reads[0].post, 6, 5, 40, 8, 20, False, num_thr=1)
    _, b8 = ob.ref_decode(reads[0].post, 6, 5, 40, 8, 20, False, num_thr=8)
    assert np.array_equal(b1, b8)


def test_tables_and_encoder_match_reference():
    for mem in (6, 8, 11, 14):
        for rate in (1, 2, 3, 4, 5, 7):
            for msg_len in (40, 41, 180, 181):
                for rc in (False, True):
                    so, to = ob.oracle_tables(mem, rate, msg_len, rc)
                    sr, tr = ob.ref_tables(mem, rate, msg_len, rc)
                    assert (so == 0) == (sr == 0)
                    if so == 0:
                        assert to == tr
    rng = np.random.default_rng(3)
    for mem, rate, msg_len in [(6, 1, 40), (8, 3, 41), (11, 5, 180), (14, 7, 181), (8, 4, 180)]:
        m = rng.integers(0, 2, msg_len, dtype=np.uint8)
        assert np.array_equal(ob.oracle_encode(m, mem, rate)[1], ob.ref_encode(m, mem, rate)[1])


def test_band_expression_is_fused():
    """Hazard 1: the -O3 reference contracts t/T*P - dev into an fma.  For P=116
    and T a multiple of 29 the fused and unfused forms give different band starts at some step;
    the oracle's default (fused) must agree with the reference on such a read."""
    P, dev = 116, 20
    diff = [(T, t) for T in (464, 493, 522) for t in range(T)
            if ob.band_start(t, T, P, dev) != ob.band_start(t, T, P, dev, unfused=True)]
    assert diff, "expected the fused and unfused band to differ somewhere"
    # a short analogue that the reference can decode quickly: find (P, T) with a fused/unfused split
    mem, rate, msg_len = 6, 5, 64
    Pn = synth.positions(mem, rate, msg_len)[0]
    hit = None
    for T in range(Pn + 1, 6 * Pn):
        if any(ob.band_start(t, T, Pn, 8) != ob.band_start(t, T, Pn, 8, unfused=True) for t in range(T)):
            hit = T
            break
    assert hit is not None
    rng = np.random.default_rng(5)
    post = np.log(rng.dirichlet(np.ones(40) * 0.3, size=hit)).astype(np.float32)
    _same(post, mem, rate, msg_len, 4, 8, False)
