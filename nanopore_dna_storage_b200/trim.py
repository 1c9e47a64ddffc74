"""Posterior trimming on the GPU: Python mirror of the reference helpers over the
C ABI (`lva_trim_batch`, include/lva_b200.h).

Names and argument meaning follow /root/reference/helper.py:157-224 and the driver loop of
/root/reference/generate_decoded_lists.py:68-83; the batch call does for many reads what the driver
does for one.  There is no CPU path here: the library must be built and a GPU present.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib

TRIM_OK, TRIM_TOO_SHORT, TRIM_BARCODE_ORDER, TRIM_TOO_FEW_BLOCKS, TRIM_NO_WINDOW = 0, 1, 2, 3, 4
INF = 0xFFFFFFFF  # the reference's np.inf distance


@dataclass
class Trim:
    start_pos: int      # first payload block (inclusive), -1 on failure   helper.py:194
    end_pos: int        # last payload block (inclusive), -1 on failure    helper.py:195
    rc: bool            # generate_decoded_lists.py:70-74
    status: int         # TRIM_*
    dist: tuple         # (start fwd, end fwd, start rc, end rc) edit distances, INF where the search failed


def find_barcodes_and_orientation(basecalls: Sequence[str], trans: Sequence[Sequence[int]], start_barcode: str,
                                  end_barcode: str, mem_conv: int, msg_len: int, device: int = 0) -> list:
    """For every read: helper.find_barcode_pos_in_post for (START, END) and for
    (rc(END), rc(START)), then the driver's choice of orientation and its size check
    (`end_pos - start_pos + 1 < MEM_CONV + MSG_LEN + 1` -> failure)."""
    n = len(basecalls)
    assert len(trans) == n
    if n == 0:
        return []
    base_off = np.zeros(n + 1, dtype=np.uint64)
    base_off[1:] = np.cumsum([len(b) for b in basecalls])
    trans_off = np.zeros(n + 1, dtype=np.uint64)
    trans_off[1:] = np.cumsum([len(t) for t in trans])
    bases = "".join(basecalls).encode("ascii")
    tr = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.uint32) for t in trans])
                              if int(trans_off[-1]) else np.zeros(1, dtype=np.uint32))
    out_start = np.empty(n, dtype=np.int32)
    out_end = np.empty(n, dtype=np.int32)
    out_rc = np.empty(n, dtype=np.int32)
    out_status = np.empty(n, dtype=np.int32)
    out_dist = np.empty((n, 4), dtype=np.uint32)
    rcode = _lib.lib().lva_trim_batch(device, bases, base_off.ctypes.data, tr.ctypes.data, trans_off.ctypes.data, n,
                                      start_barcode.encode("ascii"), end_barcode.encode("ascii"),
                                      mem_conv + msg_len + 1, out_start.ctypes.data, out_end.ctypes.data,
                                      out_rc.ctypes.data, out_status.ctypes.data, out_dist.ctypes.data)
    if rcode != 0:
        raise RuntimeError(f"lva_trim_batch failed with status {rcode}")
    return [Trim(int(out_start[i]), int(out_end[i]), bool(out_rc[i]), int(out_status[i]),
                 tuple(int(x) for x in out_dist[i])) for i in range(n)]


def truncated_block_offsets(block_offsets: np.ndarray, trims: Sequence[Trim]) -> tuple:
    """helper.truncate_post_file (helper.py:211-224) without copying: per successfully trimmed read the
    block range [first + start_pos, first + end_pos + 1) of the concatenated posterior array.  Returns
    (keep, starts, ends): indices of the reads kept and their block ranges."""
    keep, starts, ends = [], [], []
    for i, t in enumerate(trims):
        if t.status != TRIM_OK:
            continue
        first = int(block_offsets[i])
        assert first + t.end_pos + 1 <= int(block_offsets[i + 1]), "trim range outside the read's posterior blocks"
        keep.append(i)
        starts.append(first + t.start_pos)
        ends.append(first + t.end_pos + 1)
    return keep, np.asarray(starts, dtype=np.uint64), np.asarray(ends, dtype=np.uint64)
