"""Flip-flop transition posteriors on the GPU: Python mirror of flappie's
`transpost_crf_flipflop(trans, return_log=true)` (/root/reference/flappie/src/decode.c:377-497) over the
C ABI (`lva_posteriors_batch`, include/lva_b200.h).  No CPU path: the library must be built and a GPU
present."""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import _lib


def transpost_crf_flipflop(trans: Sequence[np.ndarray], device: int = 0) -> list:
    """trans: per read a (T, 40) float32 array of CRF transition weights in the .post order
    ([to*8+from] flip targets, [32+from] flop target of `from`).  Returns per read the (T, 40)
    log-posteriors, each block log-normalised -- what `flappie --post-output-file` writes."""
    n = len(trans)
    if n == 0:
        return []
    arrs = [np.ascontiguousarray(t, dtype=np.float32).reshape(-1, 40) for t in trans]
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([a.shape[0] for a in arrs])
    flat = np.ascontiguousarray(np.concatenate(arrs, axis=0)) if int(offs[-1]) else np.zeros((1, 40), np.float32)
    out = np.empty_like(flat)
    rc = _lib.lib().lva_posteriors_batch(device, flat.ctypes.data, offs.ctypes.data, n, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"lva_posteriors_batch failed with status {rc}")
    return [out[int(offs[i]):int(offs[i + 1])].copy() for i in range(n)]
