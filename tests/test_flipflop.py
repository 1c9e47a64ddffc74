"""Flip-flop transition posteriors: the C restatement pinned bit for bit against the
unmodified flappie function (CPU), lva_posteriors_batch against the restatement (GPU, through the C ABI;
floating point: tolerance stated below), and the chain posteriors -> list decoder on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(HERE, "..", "oracle")


def _lib(path, fn):
    L = C.CDLL(os.path.join(ORACLE, path))
    f = getattr(L, fn)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
    return f


def oracle_transpost(trans):
    f = _lib("_build/libflipflop_oracle.so", "ffo_transpost")
    t = np.ascontiguousarray(trans, dtype=np.float32)
    out = np.zeros_like(t)
    assert f(t.ctypes.data, t.shape[0], out.ctypes.data) == 0
    return out


def _tol(t):
    """8 float32 ulps of the largest magnitude the forward / backward recursions can reach (T * max|w|)
    + 2e-5: the posteriors are differences of values that large, in float32, in the reference too."""
    return 8.0 * float(np.spacing(np.float32(t.shape[0] * float(np.abs(t).max())))) + 2e-5


def _weights(rng, T, scale):
    """CRF-like weights: a favoured transition per block on top of noise."""
    t = rng.normal(0, scale, (T, 40)).astype(np.float32)
    t[np.arange(T), rng.integers(0, 40, T)] += np.float32(2 * scale)
    return t


def test_restatement_is_bit_identical_to_flappie():
    ref_path = os.path.join(ORACLE, "_ref", "libflipflop_ref.so")
    if not os.path.exists(ref_path):
        pytest.skip("compiled reference absent")
    ref = _lib("_ref/libflipflop_ref.so", "ffr_transpost")
    rng = np.random.default_rng(11)
    for T, scale in [(1, 1.0), (2, 3.0), (9, 0.5), (257, 2.0), (1500, 3.0), (4000, 1.0)]:
        t = _weights(rng, T, scale)
        want = np.zeros_like(t)
        assert ref(t.ctypes.data, T, want.ctypes.data) == 0
        got = oracle_transpost(t)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (T, scale)


def test_restatement_rows_are_log_normalised():
    rng = np.random.default_rng(12)
    t = _weights(rng, 300, 2.0)
    p = oracle_transpost(t).astype(np.float64)
    # float32: the row sum is subtracted from values of magnitude ~|log Z| (hundreds here)
    assert np.allclose(np.log(np.exp(p).sum(axis=1)), 0.0, atol=_tol(t))


def test_restatement_matches_float64_forward_backward():
    """Independent check of the algebra: the same posteriors from a float64 forward-backward."""
    rng = np.random.default_rng(13)
    T = 60
    t = _weights(rng, T, 1.5)
    W = np.full((T, 8, 8), -np.inf)            # W[blk, to, from]
    for to in range(4):
        W[:, to, :] = t[:, to * 8:to * 8 + 8]
    for b in range(4, 8):
        W[:, b, b] = t[:, 32 + b]
        W[:, b, b - 4] = t[:, 32 + b - 4]
    lse = lambda a, axis: np.logaddexp.reduce(a, axis=axis)
    fwd = np.zeros((T + 1, 8))
    for k in range(T):
        fwd[k + 1] = lse(W[k] + fwd[k][None, :], 1)
    bwd = np.zeros((T + 1, 8))
    for k in range(T, 0, -1):
        bwd[k - 1] = lse(W[k - 1] + bwd[k][:, None], 0)
    want = np.zeros((T, 40))
    for k in range(T):
        joint = W[k] + fwd[k][None, :] + bwd[k + 1][:, None]
        row = np.concatenate([joint[:4].reshape(-1), np.array([joint[f + 4, f] for f in range(4)] + [joint[b, b] for b in range(4, 8)])])
        want[k] = row - lse(row, 0)
    got = oracle_transpost(t)
    assert np.abs(got - want).max() < 5e-4


@pytest.mark.gpu
def test_gpu_posteriors_match_restatement():
    """Tolerance: every sum is folded in the reference's order, so the GPU differs only through the last
    bits of expf / log1pf; those perturb the forward / backward values, whose magnitude grows like
    |log Z| ~ T, and the posterior is a difference of such values.  Bound (_tol): 8 float32 ulps of
    T * max|w| + 2e-5 absolute."""
    from nanopore_dna_storage_b200 import posteriors
    rng = np.random.default_rng(14)
    reads = [_weights(rng, T, s) for T, s in [(1, 1.0), (2, 2.0), (3, 1.0), (40, 3.0), (333, 2.0), (518, 1.5),
                                              (1200, 2.5), (2500, 1.0), (5000, 2.0)] * 2]
    got = posteriors.transpost_crf_flipflop(reads)
    worst = 0.0
    for t, g in zip(reads, got):
        want = oracle_transpost(t)
        err = float(np.abs(g - want).max())
        worst = max(worst, err / _tol(t))
        assert g.shape == want.shape and err <= _tol(t), (t.shape, err, _tol(t))
        assert np.allclose(np.log(np.exp(g.astype(np.float64)).sum(axis=1)), 0.0, atol=_tol(t))
    print(f"worst error / tolerance = {worst:.3f}")


@pytest.mark.gpu
def test_gpu_posteriors_feed_the_decoder():
    """Decode the GPU posteriors and the CPU-restatement posteriors of the same weights: identical lists,
    and the sent message on top (the weights are a synthetic read's scores, sharpened)."""
    from nanopore_dna_storage_b200 import posteriors, synth
    from nanopore_dna_storage_b200.decoder import ListDecoder
    w = synth.Workload("ff m8 r5/6 L4", 8, 5, 42, 4, 20, sub=0.0, dele=0.0, ins=0.0, seed=0x5EED0F0F)
    reads = synth.make_reads(w, 3)
    weights = [np.float32(1.5) * r.post for r in reads]       # sharpened scores stand in for network outputs
    gpu_post = posteriors.transpost_crf_flipflop(weights)
    cpu_post = [oracle_transpost(x) for x in weights]
    with ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, device=0) as dec:
        a = dec.decode(gpu_post)
        b = dec.decode(cpu_post)
    for r, (ab, _), (bb, _) in zip(reads, a, b):
        assert np.array_equal(ab, bb)
        assert ab.shape[0] >= 1 and np.array_equal(ab[0], r.msg)
