"""Host-side mirror of the reference decoder interface, over the C ABI.

The reference's decode entry is
    decode_post_conv_parallel_LVA(post, msg_len, list_size, num_thr, max_deviation)
(/root/reference/viterbi/viterbi_convolutional_code.cpp:126-129) after set_conv_params(mem_conv,
rate, msg_len, sync_marker, sync_period) and the global rc_flag (:57-80).  `ListDecoder` keeps
those names and meanings; the only new notion is that a call takes a BATCH of posterior
matrices, because independent reads are what fills a B200.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence
import numpy as np

from . import _lib


class LvaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[lva {code}] {msg}")
        self.code = code


class ListDecoder:
    """One decoder configuration on one GPU.

    decode(posts) -> list of (bits[n_i, msg_len] uint8, scores[n_i] float32) in input order;
    n_i <= list_size, n_i = 0 for an empty list.  A read with fewer than nstate_pos+1 blocks
    raises, like the reference ("Too small post matrix"), unless strict=False.
    """

    def __init__(self, mem_conv: int, rate: int, msg_len: int, list_size: int = 1,
                 max_deviation: int | None = None, rc: bool = False, sync_marker: str = "",
                 sync_period: int = 0, device: int = 0, max_slots: int = 0, exact_only: bool = False,
                 unfused_band: bool = False):
        self._L = _lib.lib()
        self.mem_conv, self.rate, self.msg_len, self.list_size = mem_conv, rate, msg_len, list_size
        self.rc = rc
        p = _lib.LvaParams(mem_conv, rate, msg_len, list_size,
                           _lib.LVA_DEFAULT_DEVIATION if max_deviation is None else max_deviation,
                           int(rc), sync_marker.encode(), sync_period, device, max_slots,
                           (_lib.LVA_FLAG_EXACT_ONLY if exact_only else 0) |
                           (_lib.LVA_FLAG_UNFUSED_BAND if unfused_band else 0))
        st = C.c_int(0)
        err = C.create_string_buffer(512)
        self._h = self._L.lva_create(C.byref(p), C.byref(st), err, 512)
        if not self._h:
            raise LvaError(st.value, err.value.decode(errors="replace").strip())

    def close(self):
        if getattr(self, "_h", None):
            self._L.lva_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def nstate_pos(self) -> int:
        return self._L.lva_nstate_pos(self._h)

    @property
    def slots(self) -> int:
        return self._L.lva_slots(self._h)

    def work(self, T: int) -> dict:
        w = _lib.LvaWork()
        self._L.lva_count_work(self._h, T, C.byref(w))
        return dict(state_updates=w.state_updates, edge_acs=w.edge_acs, state_bytes=w.state_bytes)

    def timing(self) -> dict:
        t = _lib.LvaTiming()
        self._L.lva_get_timing(self._h, C.byref(t))
        return {f: getattr(t, f) for f, _ in t._fields_}

    # -- raw batch interface: concatenated blocks + offsets ---------------------------------
    def decode_packed(self, post, offsets: np.ndarray, device_ptr: int | None = None, strict=True):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        L, M = self.list_size, self.msg_len
        bits = np.zeros((n, L, M), dtype=np.uint8)
        scores = np.full((n, L), -np.inf, dtype=np.float32)
        counts = np.zeros(n, dtype=np.uint32)
        status = np.zeros(n, dtype=np.int32)
        if device_ptr is None:
            post = np.ascontiguousarray(post, dtype=np.float32)
            fn, ptr = self._L.lva_decode_batch, post.ctypes.data
        else:
            fn, ptr = self._L.lva_decode_batch_device, device_ptr
        st = fn(self._h, ptr, offsets.ctypes.data, n, bits.ctypes.data, scores.ctypes.data,
                counts.ctypes.data, status.ctypes.data)
        if st != 0:
            raise LvaError(st, self._L.lva_last_error(self._h).decode(errors="replace"))
        if strict and (status != 0).any():
            bad = int(np.flatnonzero(status)[0])
            raise LvaError(int(status[bad]), f"Too small post matrix (read {bad})")
        return bits, scores, counts, status

    def decode_ranges(self, post, starts, ends, device_ptr: int | None = None, strict=True):
        """Reads given as block ranges [starts[i], ends[i]) of ONE posterior array (host array `post`, or
        device memory at `device_ptr`): what helper.truncate_post_file does with a file copy
        (helper.py:211-224), without moving the posteriors.  -> list of (bits, scores)."""
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        ends = np.ascontiguousarray(ends, dtype=np.uint64)
        n = starts.size
        L, M = self.list_size, self.msg_len
        bits = np.zeros((n, L, M), dtype=np.uint8)
        scores = np.full((n, L), -np.inf, dtype=np.float32)
        counts = np.zeros(n, dtype=np.uint32)
        status = np.zeros(n, dtype=np.int32)
        if n == 0:
            return []
        if device_ptr is None:
            post = np.ascontiguousarray(post, dtype=np.float32)
            ptr, on_dev = post.ctypes.data, 0
        else:
            ptr, on_dev = device_ptr, 1
        st = self._L.lva_decode_ranges(self._h, ptr, on_dev, starts.ctypes.data, ends.ctypes.data, n, bits.ctypes.data,
                                       scores.ctypes.data, counts.ctypes.data, status.ctypes.data)
        if st != 0:
            raise LvaError(st, self._L.lva_last_error(self._h).decode(errors="replace"))
        if strict and (status != 0).any():
            bad = int(np.flatnonzero(status)[0])
            raise LvaError(int(status[bad]), f"Too small post matrix (read {bad})")
        return [(bits[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(n)]

    def decode_with_crc(self, posts: Sequence[np.ndarray], pad: bool, num_oligos: int):
        """decode + helper.decode_list_CRC_index on the GPU: per read (bits, scores, crc_ok[n_i],
        index[n_i], choice) with choice = first entry whose CRC matches and index < num_oligos, or -1."""
        posts = [np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 40) for p in posts]
        n = len(posts)
        if n == 0:
            return []
        offsets = np.zeros(n + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([p.shape[0] for p in posts])
        flat = np.concatenate(posts, axis=0)
        L, M = self.list_size, self.msg_len
        bits = np.zeros((n, L, M), dtype=np.uint8)
        scores = np.full((n, L), -np.inf, dtype=np.float32)
        counts = np.zeros(n, dtype=np.uint32)
        status = np.zeros(n, dtype=np.int32)
        ok = np.zeros((n, L), dtype=np.uint8)
        index = np.zeros((n, L), dtype=np.uint16)
        choice = np.zeros(n, dtype=np.int32)
        st = self._L.lva_decode_batch_crc(self._h, flat.ctypes.data, offsets.ctypes.data, n, int(pad), num_oligos,
                                          bits.ctypes.data, scores.ctypes.data, counts.ctypes.data,
                                          status.ctypes.data, ok.ctypes.data, index.ctypes.data, choice.ctypes.data)
        if st != 0:
            raise LvaError(st, self._L.lva_last_error(self._h).decode(errors="replace"))
        return [(bits[i, :counts[i]].copy(), scores[i, :counts[i]].copy(), ok[i, :counts[i]].astype(bool),
                 index[i, :counts[i]].copy(), int(choice[i])) for i in range(n)]

    def decode(self, posts: Sequence[np.ndarray], strict=True):
        posts = [np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 40) for p in posts]
        offsets = np.zeros(len(posts) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([p.shape[0] for p in posts])
        flat = np.concatenate(posts, axis=0) if posts else np.zeros((0, 40), np.float32)
        bits, scores, counts, _ = self.decode_packed(flat, offsets, strict=strict)
        return [(bits[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(len(posts))]


def conv_encode(msg_bits, mem_conv: int, rate: int) -> np.ndarray:
    """conv_encode of the reference (VCC:450-499) through the C ABI (host only)."""
    L = _lib.lib()
    msg = np.ascontiguousarray(msg_bits, dtype=np.uint8)
    out = np.zeros(2 * (msg.size + mem_conv) + 8, dtype=np.uint8)
    n = C.c_uint32(0)
    st = L.lva_encode(mem_conv, rate, msg.ctypes.data, msg.size, out.ctypes.data, out.size, C.byref(n))
    if st != 0:
        raise LvaError(st, "encode failed")
    return out[:n.value].copy()


def code_tables(mem_conv: int, rate: int, msg_len: int, rc: bool = False, sync_marker: str = "",
                sync_period: int = 0) -> dict:
    """The tables set_conv_params builds (VCC:264-415), plus the per-row validity mask."""
    L = _lib.lib()
    t = _lib.LvaCodeTables()
    st = L.lva_get_code_tables(mem_conv, rate, msg_len, int(rc), sync_marker.encode(), sync_period, C.byref(t))
    if st != 0:
        raise LvaError(st, "bad code parameters")
    P = t.nstate_pos
    return dict(G=(t.G[0], t.G[1]), init=t.initial_state, final=t.final_state, punc=list(t.punc[:t.punc_len]),
                nstate_pos=P, pos2msg=list(t.pos2msg[:P]), valid_mask=list(t.valid_mask[:P]),
                valid_value=list(t.valid_value[:P]), history_words=list(t.history_words[:P]),
                first_run_row=list(t.first_run_row[:P]))


def band(t: int, nblk: int, nstate_pos: int, max_deviation: int, unfused: bool = False):
    L = _lib.lib()
    a, b = C.c_uint32(0), C.c_uint32(0)
    L.lva_band_ex(t, nblk, nstate_pos, max_deviation, _lib.LVA_FLAG_UNFUSED_BAND if unfused else 0, C.byref(a), C.byref(b))
    return a.value, b.value
