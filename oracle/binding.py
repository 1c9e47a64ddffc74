"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle and the compiled reference.

  oracle/_build/liblva_oracle.so   restatement (oracle/lva_oracle.cpp)
  oracle/_ref/libvcc_ref.so        the unmodified reference, compiled in place (oracle/ref_shim.cpp)
  oracle/_ref/viterbi_nanopore.out the unmodified reference CLI

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liblva_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvcc_ref.so")
REF_EXE = os.path.join(HERE, "_ref", "viterbi_nanopore.out")
DEFAULT_DEV = 0xFFFFFFFF


def build(quiet: bool = True) -> None:
    """make -C oracle (oracle always; reference only when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Params(C.Structure):
    _fields_ = [("mem_conv", C.c_uint32), ("rate", C.c_uint32), ("msg_len", C.c_uint32),
                ("list_size", C.c_uint32), ("max_deviation", C.c_uint32), ("rc", C.c_int32),
                ("sync_marker", C.c_char_p), ("sync_period", C.c_uint32),
                ("num_thr", C.c_uint32), ("unfused_band", C.c_int32)]


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.lvo_band_start.restype = C.c_int64
        _oracle.lvo_band_start.argtypes = [C.c_uint32] * 4 + [C.c_int]
    return _oracle


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        _ref = C.CDLL(REF_SO)
    return _ref


def have_ref() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_EXE)


def _dev(max_deviation):
    return DEFAULT_DEV if max_deviation is None else int(max_deviation)


def oracle_decode(post: np.ndarray, mem_conv: int, rate: int, msg_len: int, list_size: int = 1,
                  max_deviation=None, rc: bool = False, sync_marker: str = "",
                  sync_period: int = 0, num_thr: int = 0, unfused_band: bool = False):
    """-> (status, bits[n, msg_len] uint8, scores[n] float32)"""
    post = np.ascontiguousarray(post, dtype=np.float32).reshape(-1, 40)
    p = _Params(mem_conv, rate, msg_len, list_size, _dev(max_deviation), int(rc),
                sync_marker.encode(), sync_period, num_thr or (os.cpu_count() or 1),
                int(unfused_band))
    bits = np.zeros((list_size, msg_len), dtype=np.uint8)
    scores = np.zeros(list_size, dtype=np.float32)
    n = C.c_uint32(0)
    st = oracle_lib().lvo_decode(C.byref(p), post.ctypes.data_as(C.c_void_p),
                                 C.c_uint32(post.shape[0]), bits.ctypes.data_as(C.c_void_p),
                                 scores.ctypes.data_as(C.c_void_p), C.byref(n))
    return st, bits[:n.value].copy(), scores[:n.value].copy()


def ref_decode(post: np.ndarray, mem_conv: int, rate: int, msg_len: int, list_size: int = 1,
               max_deviation=None, rc: bool = False, sync_marker: str = "",
               sync_period: int = 0, num_thr: int = 0):
    """The unmodified reference, in process. -> (status, bits[n, msg_len] uint8)"""
    post = np.ascontiguousarray(post, dtype=np.float32).reshape(-1, 40)
    dev = msg_len + mem_conv + 1 if max_deviation is None else int(max_deviation)  # VCC:238-242
    bits = np.zeros((max(list_size, 1), msg_len), dtype=np.uint8)
    n = C.c_uint32(0)
    st = ref_lib().vcc_ref_decode(post.ctypes.data_as(C.c_void_p), C.c_uint32(post.shape[0]),
                                  C.c_uint32(mem_conv), C.c_uint32(rate), C.c_uint32(msg_len),
                                  C.c_uint32(list_size), C.c_uint32(num_thr or (os.cpu_count() or 1)),
                                  C.c_uint32(dev), C.c_int(int(rc)), sync_marker.encode(),
                                  C.c_uint32(sync_period), bits.ctypes.data_as(C.c_void_p),
                                  C.byref(n))
    return st, bits[:n.value].copy()


def _tables(lib, fn, mem_conv, rate, msg_len, rc):
    G = (C.c_uint32 * 2)()
    inf = (C.c_uint32 * 2)()
    punc = (C.c_uint8 * 16)()
    plen = C.c_uint32(0)
    P = C.c_uint32(0)
    p2m = (C.c_uint32 * 256)()
    st = getattr(lib, fn)(C.c_uint32(mem_conv), C.c_uint32(rate), C.c_uint32(msg_len),
                          C.c_int(int(rc)), G, inf, punc, C.byref(plen), C.byref(P), p2m)
    if st:
        return st, None
    return 0, dict(G=(G[0], G[1]), init=inf[0], final=inf[1],
                   punc=list(punc[:plen.value]), nstate_pos=P.value,
                   pos2msg=list(p2m[:min(P.value, 256)]))


def oracle_tables(mem_conv, rate, msg_len, rc=False):
    return _tables(oracle_lib(), "lvo_tables", mem_conv, rate, msg_len, rc)


def ref_tables(mem_conv, rate, msg_len, rc=False):
    return _tables(ref_lib(), "vcc_ref_tables", mem_conv, rate, msg_len, rc)


def _encode(lib, fn, order, msg, mem_conv, rate):
    msg = np.ascontiguousarray(msg, dtype=np.uint8)
    out = np.zeros(2 * (msg.size + mem_conv) + 8, dtype=np.uint8)
    n = C.c_uint32(0)
    if order == "oracle":
        st = getattr(lib, fn)(C.c_uint32(mem_conv), C.c_uint32(rate), msg.ctypes.data_as(C.c_void_p),
                              C.c_uint32(msg.size), out.ctypes.data_as(C.c_void_p), C.byref(n))
    else:
        st = getattr(lib, fn)(msg.ctypes.data_as(C.c_void_p), C.c_uint32(msg.size),
                              C.c_uint32(mem_conv), C.c_uint32(rate),
                              out.ctypes.data_as(C.c_void_p), C.byref(n))
    return st, out[:n.value].copy()


def oracle_encode(msg, mem_conv, rate):
    return _encode(oracle_lib(), "lvo_encode", "oracle", msg, mem_conv, rate)


def ref_encode(msg, mem_conv, rate):
    return _encode(ref_lib(), "vcc_ref_encode", "ref", msg, mem_conv, rate)


def band_start(t, nblk, P, dev, unfused=False) -> int:
    return int(oracle_lib().lvo_band_start(t, nblk, P, dev, int(unfused)))


def std_heap_op(op: int, scores, tags):
    """Run libstdc++ make_heap(0) / pop_heap(1) / push_heap(2) on (score, tag) arrays."""
    s = np.ascontiguousarray(scores, dtype=np.float32).copy()
    g = np.ascontiguousarray(tags, dtype=np.uint32).copy()
    oracle_lib().lvo_std_heap_op(C.c_int(op), s.ctypes.data_as(C.c_void_p),
                                 g.ctypes.data_as(C.c_void_p), C.c_uint32(s.size))
    return s, g
