"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's posterior trimming, used by tests/
as the checker of lva_trim_batch.  Nothing under nanopore_dna_storage_b200/ may import this.

Follows /root/reference/helper.py:157-209 (find_barcode_pos_in_post), helper.py:226-229
(reverse_complement) and /root/reference/generate_decoded_lists.py:33-34, 68-83 line by line.
`distance.levenshtein` (PyPI `distance`, not installed here: parity unpinned against the package) is
restated as the textbook unit-cost edit distance; pinned by known answers in tests/test_trim.py.
"""
import math

INF = math.inf


def levenshtein(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def reverse_complement(dna: str) -> str:                       # helper.py:226-229
    complement = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}
    return ''.join(complement[base] for base in dna[::-1])


def find_barcode_pos_in_post(basecall, trans_arr, start_barcode, end_barcode):
    """helper.py:157-209 with the file reads replaced by their contents.  Returns
    (start_pos, end_pos, dist_start, dist_end); raises where the reference raises."""
    basecall_len = len(basecall)
    start_barcode_len, end_barcode_len = len(start_barcode), len(end_barcode)
    if start_barcode_len + end_barcode_len > basecall_len:     # helper.py:177-179
        return (-1, -1, INF, INF)
    start_bc_edit_distance = [levenshtein(start_barcode, basecall[i:i + start_barcode_len])
                              for i in range(basecall_len // 2 + 1 - start_barcode_len)]        # helper.py:181-184
    end_bc_edit_distance = [levenshtein(end_barcode, basecall[i:i + end_barcode_len])
                            for i in range(basecall_len // 2, basecall_len - end_barcode_len)]  # helper.py:186-188
    start_bc_first_base = start_bc_edit_distance.index(min(start_bc_edit_distance))  # raises on an empty list
    end_bc_first_base = basecall_len // 2 + end_bc_edit_distance.index(min(end_bc_edit_distance))
    start_bc_last_base = start_bc_first_base + start_barcode_len - 1
    start_pos = trans_arr[start_bc_last_base + 1] - 1
    end_pos = trans_arr[end_bc_first_base - 1] - 1
    if end_pos < start_pos:                                    # helper.py:207-209
        return (-1, -1, INF, INF)
    return (start_pos, end_pos, min(start_bc_edit_distance), min(end_bc_edit_distance))


def trim_read(basecall, trans_arr, start_barcode, end_barcode, mem_conv, msg_len):
    """generate_decoded_lists.py:33-34, 68-83: (start_pos, end_pos, rc, ok, dists)."""
    start_rc, end_rc = reverse_complement(end_barcode), reverse_complement(start_barcode)
    (start_pos, end_pos, ds, de) = find_barcode_pos_in_post(basecall, trans_arr, start_barcode, end_barcode)
    (start_pos_rc, end_pos_rc, ds_rc, de_rc) = find_barcode_pos_in_post(basecall, trans_arr, start_rc, end_rc)
    rc = False
    if ds + de > ds_rc + de_rc:
        rc = True
        start_pos, end_pos = start_pos_rc, end_pos_rc
    ok = not (start_pos == -1 or end_pos - start_pos + 1 < mem_conv + msg_len + 1)
    return start_pos, end_pos, rc, ok, (ds, de, ds_rc, de_rc)
