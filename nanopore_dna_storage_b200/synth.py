"""Synthetic reads for the list-Viterbi decode path.

Random CRC'd messages -> convolutional encode -> optional reverse complement ->
iid insertion/deletion/substitution channel -> flip-flop state path with random dwell
-> per-block 40 log-posteriors in the flappie transition layout.

Reference conventions restated here (nothing is imported from the reference):
  * message framing   helper.py:27-32, 257-264  (12-bit PRP'd index | payload | CRC-8 | pad)
  * conv. encoder     viterbi/viterbi_convolutional_code.cpp:264-357, 426-499
  * channel           helper.py:34-64           (per step: insert? else delete? else sub?)
  * posterior layout  viterbi_convolutional_code.cpp:553-587, flappie/src/decode.c:104-114
                      index = to*8+from for flip targets to in 0..3, 32+from for the flop
                      target of `from`; values are per-block log-normalised (logsumexp = 0).

Everything is seeded per read by a counter-based generator (Philox key = seed,
counter = read index), so any read of any workload can be regenerated on its own.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

PRP_A, PRP_B, PRP_A_INV, INDEX_LEN, CRC_LEN = 1751, 2532, 3303, 12, 8  # helper.py:27-32

_G = {6: (0o171, 0o133, 0b100101), 8: (0o515, 0o677, 0b10010110),
      11: (0o5537, 0o6131, 0b10010110001), 14: (0o75063, 0o56711, 0b10010110001101)}
_PUNC = {1: [0], 2: [0, 2, 0], 3: [0, 1], 4: [0, 3, 0, 2, 1], 5: [0, 1, 2], 7: [0, 3, 1, 1]}


def crc8(data: bytes) -> int:
    """CRC-8, polynomial 0x07, init 0, no reflection, no final xor (PyPI `crc8` default;
    check value crc8(b'123456789') == 0xF4)."""
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def _bitrev(v: int, n: int) -> int:
    r = 0
    for i in range(n):
        r = (r << 1) | ((v >> i) & 1)
    return r


def positions(mem: int, rate: int, msg_len: int):
    """(nstate_pos, st_pos2msg_pos) or None when msg_len+mem is not hit exactly."""
    pat = _PUNC[rate]
    n, pos = 0, [0]
    while n < msg_len + mem:
        n += 1 if pat[(len(pos) - 1) % len(pat)] == 0 else 2
        pos.append(n)
    return (len(pos), pos) if n == msg_len + mem else None


def conv_encode(msg_bits, mem: int, rate: int) -> np.ndarray:
    """Punctured rate-1/2 convolutional code, terminated in the bit-reversed start state.
    Returns the coded bit array (2 bits per base)."""
    g0, g1, init = _G[mem]
    fin = _bitrev(init, mem)
    msg_bits = [int(b) for b in msg_bits]
    bits = msg_bits + [(fin >> i) & 1 for i in range(mem)]
    state, full = init, []
    for b in bits:
        reg = state | (b << mem)
        full.append(bin(reg & g0).count("1") & 1)
        full.append(bin(reg & g1).count("1") & 1)
        state = reg >> 1
    assert state == fin
    pos = positions(mem, rate, len(msg_bits))
    if pos is None:
        raise ValueError("Output length not even. Try padding with a single 0 at end.")
    pat, out, i = _PUNC[rate], [], 0
    for p in range(pos[0] - 1):
        blk = pat[p % len(pat)]
        if blk == 0:
            out += [full[i], full[i + 1]]; i += 2
        elif blk == 1:
            out += [full[i + 1], full[i + 2]]; i += 4
        elif blk == 2:
            out += [full[i], full[i + 3]]; i += 4
        else:
            out += [full[i + 1], full[i + 3]]; i += 4
    assert i == len(full)
    return np.array(out, dtype=np.uint8)


def frame_message(index: int, payload_bits: np.ndarray, pad: bool) -> np.ndarray:
    """index(12, PRP'd) | payload | CRC-8 over (2 index bytes | payload bytes) | optional 0."""
    idx = (PRP_A * index + PRP_B) % (1 << INDEX_LEN)
    idx_bits = [(idx >> (INDEX_LEN - 1 - i)) & 1 for i in range(INDEX_LEN)]
    payload_bits = np.asarray(payload_bits, dtype=np.uint8)
    assert payload_bits.size % 8 == 0
    payload_bytes = np.packbits(payload_bits).tobytes()
    crc = crc8(idx.to_bytes(2, "big") + payload_bytes)
    crc_bits = [(crc >> (7 - i)) & 1 for i in range(8)]
    bits = idx_bits + payload_bits.tolist() + crc_bits + ([0] if pad else [])
    return np.array(bits, dtype=np.uint8)


def channel(bases: np.ndarray, rng, sub: float, dele: float, ins: float) -> np.ndarray:
    out, i, n = [], 0, len(bases)
    while True:
        if rng.random() < ins:
            out.append(int(rng.integers(4)))
        else:
            if i == n:
                break
            if rng.random() < dele:
                i += 1
            else:
                if rng.random() < sub:
                    out.append(int((bases[i] + 1 + rng.integers(3)) % 4))
                else:
                    out.append(int(bases[i]))
                i += 1
    return np.array(out, dtype=np.uint8)


def posteriors_from_bases(bases: np.ndarray, rng, beta: float, sigma: float,
                          mean_extra_dwell: float = 3.5, quantum: float = 0.0,
                          dwell=None) -> np.ndarray:
    """T x 40 float32 log-posteriors for a base sequence.

    State path: first base enters its flip state; a base equal to the previous one toggles
    flip<->flop; a different base enters its flip state.  Each base dwells 1+Poisson blocks:
    the first block carries the move transition, the rest the stay transition.
    `quantum` > 0 rounds the logits to a coarse grid before normalisation AND snaps the
    normalised values to that grid, forcing exact score ties (hazard corpus item 5)."""
    n = len(bases)
    if dwell is None:
        dwell = 1 + rng.poisson(mean_extra_dwell, size=n)
    T = int(dwell.sum())
    true_idx = np.empty(T, dtype=np.int64)
    t, prev_state = 0, None
    for k in range(n):
        b = int(bases[k])
        if prev_state is None:
            state = b
            frm = (b + 1 + int(rng.integers(3))) % 4  # arbitrary different flip state
        else:
            frm = prev_state
            state = b + 4 if prev_state == b else b   # b+ -> b- ; b- or other -> b+
        # move transition frm -> state
        true_idx[t] = state * 8 + frm if state < 4 else 32 + frm
        # stays
        stay = state * 8 + state if state < 4 else 32 + state
        true_idx[t + 1:t + dwell[k]] = stay
        t += int(dwell[k])
        prev_state = state
    logits = rng.normal(0.0, sigma, size=(T, 40)).astype(np.float32)
    logits[np.arange(T), true_idx] += np.float32(beta)
    if quantum > 0:
        logits = (np.round(logits / quantum) * quantum).astype(np.float32)
    mx = logits.max(axis=1, keepdims=True)
    lse = mx + np.log(np.exp(logits - mx).sum(axis=1, keepdims=True, dtype=np.float32))
    post = (logits - lse).astype(np.float32)
    if quantum > 0:
        post = (np.round(post / quantum) * quantum).astype(np.float32)
    return np.ascontiguousarray(post)


@dataclass
class Workload:
    """One of BASELINE.json's configs, as decoder parameters + channel."""
    name: str
    mem_conv: int
    rate: int
    msg_len: int
    list_size: int
    max_deviation: int | None = 20
    sub: float = 0.0
    dele: float = 0.0
    ins: float = 0.0
    rc_fraction: float = 0.0
    sigma: float = 1.5
    betas: tuple = (8.0, 5.0, 4.0, 3.0)
    seed: int = 0x5EED0000
    quantum: float = 0.0

    @property
    def pad(self) -> bool:
        return (self.msg_len - INDEX_LEN - CRC_LEN) % 8 == 1


@dataclass
class Read:
    index: int
    msg: np.ndarray          # msg_len bits, forward orientation
    rc: bool
    post: np.ndarray         # T x 40 float32
    beta: float


def make_read(w: Workload, i: int) -> Read:
    rng = np.random.Generator(np.random.Philox(key=w.seed, counter=[i, 0, 0, 0]))
    pad = w.pad
    npay = w.msg_len - INDEX_LEN - CRC_LEN - (1 if pad else 0)
    if npay >= 0 and npay % 8 == 0:
        msg = frame_message(i % 4096, rng.integers(0, 2, size=npay, dtype=np.uint8), pad)
    else:  # short test messages: plain random bits, no framing
        msg = rng.integers(0, 2, size=w.msg_len, dtype=np.uint8)
    coded = conv_encode(msg, w.mem_conv, w.rate)
    bases = (2 * coded[0::2] + coded[1::2]).astype(np.uint8)
    rc = bool(rng.random() < w.rc_fraction)
    if rc:
        bases = (3 - bases[::-1]).astype(np.uint8)
    bases = channel(bases, rng, w.sub, w.dele, w.ins)
    beta = float(w.betas[int(rng.integers(len(w.betas)))])
    post = posteriors_from_bases(bases, rng, beta, w.sigma, quantum=w.quantum)
    return Read(i, msg, rc, post, beta)


def make_reads(w: Workload, n: int, start: int = 0):
    return [make_read(w, start + i) for i in range(n)]


def workloads() -> dict:
    """BASELINE.json configs (C1..C5) as Workload objects."""
    ch = dict(sub=0.004, dele=0.0085, ins=0.0005)
    return {
        "C1": Workload("C1 m6 r1/2 msg180 L1", 6, 1, 180, 1, seed=0x5EED0001, **ch),
        "C2": Workload("C2 m11 r5/6 msg180 L8", 11, 5, 180, 8, seed=0x5EED0002, **ch),
        "C3": Workload("C3 m14 r1/2 msg100 L8", 14, 1, 100, 8, seed=0x5EED0003, **ch),
        "C4r2": Workload("C4 m8 r2/3 msg180 L8", 8, 2, 180, 8, seed=0x5EED0004, **ch),
        "C4r3": Workload("C4 m8 r3/4 msg181 L8", 8, 3, 181, 8, seed=0x5EED0005, **ch),
        "C4r4": Workload("C4 m8 r4/5 msg180 L8", 8, 4, 180, 8, seed=0x5EED0006, **ch),
        "C4r7": Workload("C4 m8 r7/8 msg181 L8", 8, 7, 181, 8, seed=0x5EED0007, **ch),
        "C4r2L64": Workload("m8 r2/3 msg180 L64 (the paper's list size for m=8)", 8, 2, 180, 64, seed=0x5EED0009, **ch),
        "C5": Workload("C5 m11 r5/6 msg180 L16 fwd+rc", 11, 5, 180, 16, seed=0x5EED0008,
                       rc_fraction=0.5, **ch),
    }
