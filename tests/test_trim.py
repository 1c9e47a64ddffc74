"""Posterior trimming: oracle pinned by known answers (CPU), lva_trim_batch against
the oracle on seeded reads (GPU, through the C ABI)."""
import numpy as np
import pytest

from oracle import trim_oracle as to

START, END = "CACCTGTGCTGCGTCAGGCTGTGTC", "GCTGTCCGTTCCGCATTGACACGGC"  # 25-mers as used with the real data


def test_levenshtein_known_answers():
    assert to.levenshtein("kitten", "sitting") == 3
    assert to.levenshtein("flaw", "lawn") == 2
    assert to.levenshtein("", "abc") == 3
    assert to.levenshtein("ACGT", "ACGT") == 0
    assert to.levenshtein("AAAA", "TTTT") == 4
    assert to.levenshtein("GATTACA", "GCATGCU") == 4


def test_reverse_complement():
    assert to.reverse_complement("AACGTN") == "NACGTT"


def _read(rng, payload_bases, flank=(10, 10), errors=0, rc=False):
    bases = "ACGT"
    body = START + "".join(rng.choice(list(bases), payload_bases)) + END
    if rc:
        body = to.reverse_complement(body)
    seq = list("".join(rng.choice(list(bases), flank[0])) + body + "".join(rng.choice(list(bases), flank[1])))
    for _ in range(errors):
        k = int(rng.integers(len(seq)))
        op = int(rng.integers(3))
        if op == 0:
            seq[k] = bases[int(rng.integers(4))]
        elif op == 1:
            del seq[k]
        else:
            seq.insert(k, bases[int(rng.integers(4))])
    seq = "".join(seq)
    trans = np.cumsum(1 + rng.poisson(3.5, len(seq))).astype(np.uint32)  # block at which each base is emitted
    return seq, trans


def test_oracle_finds_clean_barcodes():
    rng = np.random.default_rng(7)
    seq, trans = _read(rng, 220, flank=(12, 9))
    sp, ep, ds, de = to.find_barcode_pos_in_post(seq, list(trans), START, END)
    assert (ds, de) == (0, 0)
    first_payload_base = 12 + len(START)
    assert sp == trans[first_payload_base] - 1
    assert ep == trans[12 + len(START) + 220 - 1] - 1
    start_pos, end_pos, rc, ok, _ = to.trim_read(seq, list(trans), START, END, 11, 180)
    assert (start_pos, end_pos, rc, ok) == (sp, ep, False, True)
    seq_rc, trans_rc = _read(rng, 220, rc=True)
    assert to.trim_read(seq_rc, list(trans_rc), START, END, 11, 180)[2] is True


def test_oracle_failures():
    assert to.find_barcode_pos_in_post("ACGT" * 5, list(range(1, 21)), START, END) == (-1, -1, to.INF, to.INF)
    rng = np.random.default_rng(3)
    seq, trans = _read(rng, 40)  # payload too short for m=11, msg_len=180
    assert to.trim_read(seq, list(trans), START, END, 11, 180)[3] is False


def test_abi_exports_trim_symbol():
    from nanopore_dna_storage_b200 import _lib
    assert hasattr(_lib.lib(), "lva_trim_batch")


@pytest.mark.gpu
def test_gpu_trim_matches_oracle():
    from nanopore_dna_storage_b200 import trim
    rng = np.random.default_rng(20261017)
    seqs, transs = [], []
    for i in range(160):
        kind = i % 8
        if kind == 0:
            s, t = "".join(rng.choice(list("ACGT"), int(rng.integers(5, 49)))), None           # shorter than the barcodes
        elif kind == 1:
            s, t = _read(rng, int(rng.integers(20, 60)), errors=int(rng.integers(0, 6)))        # too few blocks
        elif kind == 2:
            s, t = "".join(rng.choice(list("ACGT"), int(rng.integers(60, 400)))), None         # no barcodes at all
        else:
            s, t = _read(rng, int(rng.integers(150, 260)), flank=(int(rng.integers(0, 40)), int(rng.integers(0, 40))),
                         errors=int(rng.integers(0, 25)), rc=bool(i & 1))
        if t is None:
            t = np.cumsum(1 + rng.poisson(3.5, len(s))).astype(np.uint32)
        seqs.append(s)
        transs.append(t)
    got = trim.find_barcodes_and_orientation(seqs, transs, START, END, 11, 180)
    n_ok = 0
    for s, t, g in zip(seqs, transs, got):
        try:
            sp, ep, rc, ok, d = to.trim_read(s, [int(x) for x in t], START, END, 11, 180)
        except ValueError:   # min() of an empty window list: the reference raises
            assert g.status == trim.TRIM_NO_WINDOW and g.start_pos == -1
            continue
        want_d = tuple(trim.INF if x == to.INF else int(x) for x in d)
        assert g.dist == want_d, (s, g, d)
        assert g.rc == rc
        assert (g.status == trim.TRIM_OK) == ok
        assert (g.start_pos, g.end_pos) == (sp, ep)
        n_ok += ok
    assert n_ok >= 60  # the corpus really exercises the success path


@pytest.mark.gpu
def test_gpu_trim_then_decode_uses_block_ranges():
    """The trimmed range is a pair of block offsets: decoding [start, end+1) of a padded posterior array
    gives the list of the un-padded read."""
    from nanopore_dna_storage_b200 import synth, trim
    from nanopore_dna_storage_b200.decoder import ListDecoder
    w = synth.Workload("trim m8 r5/6 L4", 8, 5, 42, 4, 20, sub=0.0, dele=0.0, ins=0.0, seed=0x5EED0B0B)
    reads = synth.make_reads(w, 3)
    rng = np.random.default_rng(5)
    pads = [(int(rng.integers(1, 30)), int(rng.integers(1, 30))) for _ in reads]
    junk = lambda n: np.full((n, 40), np.float32(np.log(1 / 40)), dtype=np.float32)
    padded = [np.concatenate([junk(a), r.post, junk(b)]) for r, (a, b) in zip(reads, pads)]
    block_offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    block_offsets[1:] = np.cumsum([p.shape[0] for p in padded])
    trims = [trim.Trim(a, a + r.post.shape[0] - 1, False, trim.TRIM_OK, (0, 0, 0, 0)) for r, (a, _) in zip(reads, pads)]
    keep, starts, ends = trim.truncated_block_offsets(block_offsets, trims)
    assert keep == [0, 1, 2]
    flat = np.concatenate(padded)
    import torch
    with ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, device=0) as dec:
        want = dec.decode([r.post for r in reads])
        got = dec.decode_ranges(flat, starts, ends)                       # lva_decode_ranges, host posteriors
        dev = torch.from_numpy(flat).cuda()
        got_dev = dec.decode_ranges(None, starts, ends, device_ptr=dev.data_ptr())  # ... posteriors left in HBM
        # ranges may come in any order and overlap
        perm = [2, 0, 1, 0]
        got_perm = dec.decode_ranges(flat, starts[perm], ends[perm])
    for (wb, ws), (gb, gs), (db, ds) in zip(want, got, got_dev):
        assert np.array_equal(wb, gb) and np.array_equal(ws.view(np.uint32), gs.view(np.uint32))
        assert np.array_equal(wb, db) and np.array_equal(ws.view(np.uint32), ds.view(np.uint32))
    for k, (gb, gs) in zip(perm, got_perm):
        assert np.array_equal(want[k][0], gb) and np.array_equal(want[k][1].view(np.uint32), gs.view(np.uint32))
