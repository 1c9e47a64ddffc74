"""TEST INFRASTRUCTURE ONLY: restatement of the list check the reference's drivers run after decoding
(helper.decode_list_CRC_index, /root/reference/helper.py:371-388, framing helper.py:257-264), the
checker for the GPU CRC-8 / index sidecar (lva_decode_batch_crc).

CRC-8 itself lives in PyPI `crc8` (install_python_packages.sh:3), which is not vendored by the
reference and not installable here (no network, not in the wheelhouse): PARITY UNPINNED against the
real package.  Restated from its published behaviour: polynomial 0x07, initial value 0, no
reflection, no final xor; pinned only by the standard check value CRC-8("123456789") = 0xF4.
"""
import numpy as np

PRP_A, PRP_B, PRP_A_INV, INDEX_LEN = 1751, 2532, 3303, 12  # helper.py:27-32


def crc8(data: bytes) -> int:
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def check_crc_index(bits, pad: bool):
    """helper.decode_list_CRC_index for one candidate: (crc_ok, index)."""
    bits = [int(b) for b in bits]
    if pad:
        bits = bits[:-1]                      # helper.py:374-375: strip the pad bit
    n = (len(bits) + 7) // 8 * 8
    bits = [0] * (n - len(bits)) + bits       # helper.py:376-378: left-pad to whole bytes
    by = np.packbits(np.array(bits, dtype=np.uint8)).tobytes()
    ok = crc8(by[:-1]) == by[-1]              # helper.py:379-381
    idx_raw = int.from_bytes(by[:2], "big") & ((1 << INDEX_LEN) - 1)
    return ok, (PRP_A_INV * (idx_raw - PRP_B)) % (1 << INDEX_LEN)   # helper.py:382-386
