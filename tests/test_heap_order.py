"""The device heap primitives (heap_order.cuh, compiled for the host) must move elements exactly
like libstdc++'s make_heap / pop_heap / push_heap, ties included: the reference's list order
under score ties is whatever those routines produce (VCC:762-795)."""
import ctypes as C
import os
import numpy as np
import pytest

from oracle import binding as ob

SHIM = os.path.join(os.path.dirname(ob.ORACLE_SO), "libheap_shim.so")


def ours(op, s, t):
    lib = C.CDLL(SHIM)
    s = np.ascontiguousarray(s, dtype=np.float32).copy()
    t = np.ascontiguousarray(t, dtype=np.uint32).copy()
    lib.lva_heap_op(C.c_int(op), s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.c_uint32(s.size))
    return s, t


@pytest.mark.parametrize("distinct", [2, 4, 1000])
def test_heap_sequences_match_libstdcxx(distinct):
    ob.oracle_lib()
    rng = np.random.default_rng(distinct)
    pops = 0
    for trial in range(3000):
        n = int(rng.integers(1, 9))
        s = rng.integers(0, distinct, size=n).astype(np.float32)
        t = np.arange(n, dtype=np.uint32)
        a, at = ob.std_heap_op(0, s, t)
        b, bt = ours(0, s, t)
        assert np.array_equal(a, b) and np.array_equal(at, bt), "make_heap layout differs"
        s, t = a, at
        tag = 100
        while len(s) > 0:
            a, at = ob.std_heap_op(1, s, t)
            b, bt = ours(1, s, t)
            assert np.array_equal(a, b) and np.array_equal(at, bt), "pop_heap layout differs"
            pops += 1
            s, t = a[:-1], at[:-1]
            if rng.random() < 0.6 and len(s) < 8:
                s = np.append(s, np.float32(rng.integers(0, distinct)))
                t = np.append(t, np.uint32(tag))
                tag += 1
                a, at = ob.std_heap_op(2, s, t)
                b, bt = ours(2, s, t)
                assert np.array_equal(a, b) and np.array_equal(at, bt), "push_heap layout differs"
                s, t = a, at
    assert pops > 5000
