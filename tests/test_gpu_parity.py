"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on seeded synthetic reads.

Bit-exact decoded lists AND bit-exact path metrics (the reference's arithmetic is single fp32
adds in a fixed order, so the GPU must reproduce it exactly -- tolerance 0; north_star asks for
1e-5 relative)."""
import numpy as np
import pytest

from nanopore_dna_storage_b200 import synth
from oracle import crc_oracle
from nanopore_dna_storage_b200.decoder import ListDecoder

pytestmark = pytest.mark.gpu


def _compare(oracle, w: synth.Workload, n_reads, exact_only, rc=False, quantum=0.0, max_slots=0, tag=""):
    reads = [synth.make_read(w, i) for i in range(n_reads)]
    for r in reads:
        r.rc = rc
    with ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, rc=rc,
                     exact_only=exact_only, max_slots=max_slots) as dec:
        got = dec.decode([r.post for r in reads])
        tm = dec.timing()
    bad = []
    for r, (bits, scores) in zip(reads, got):
        st, obits, oscores = oracle.oracle_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size,
                                                  w.max_deviation, rc)
        assert st == 0
        if bits.shape != obits.shape or not np.array_equal(bits, obits):
            bad.append((r.index, "bits", bits.shape, obits.shape))
        elif not np.array_equal(scores.view(np.uint32), oscores.view(np.uint32)):
            bad.append((r.index, "scores", scores, oscores))
    assert not bad, f"{tag}: {len(bad)}/{n_reads} reads differ: {bad[:3]}"
    return tm


CASES = [
    # mem, rate, msg_len, L, dev
    (6, 1, 40, 1, 20), (6, 1, 40, 1, None), (6, 1, 40, 4, 20), (6, 2, 41, 8, 20), (6, 3, 40, 8, 10),
    (6, 4, 40, 2, 20), (6, 5, 40, 8, 20), (6, 7, 41, 8, 20), (8, 5, 40, 8, 20), (8, 3, 40, 3, 20),
    (11, 5, 29, 8, 20), (6, 1, 60, 16, 20), (6, 5, 40, 5, None),
    # several completed 32-bit chunks of message history behind every entry (chain ids, boundary
    # rows), narrow bands (window slots recycled every few steps, stale rows), list sizes above 16
    (6, 5, 130, 8, 20), (6, 1, 100, 4, 6), (8, 2, 100, 8, 3), (6, 7, 130, 2, 10), (8, 4, 102, 16, 8),
    (6, 1, 40, 24, 20), (6, 5, 70, 64, 20), (8, 5, 40, 32, 4), (6, 2, 130, 17, 5),
]


@pytest.mark.parametrize("mem,rate,msg_len,L,dev", CASES)
@pytest.mark.parametrize("rc", [False, True])
def test_exact_path_matches_oracle(oracle, mem, rate, msg_len, L, dev, rc):
    w = synth.Workload("t", mem, rate, msg_len, L, dev, sub=0.01, dele=0.01, ins=0.005,
                       seed=4000 + mem * 10 + rate)
    _compare(oracle, w, 6, exact_only=True, rc=rc, tag=f"m{mem} r{rate} L{L} dev{dev} rc{rc}")


@pytest.mark.parametrize("mem,rate,msg_len,L,dev", CASES)
@pytest.mark.parametrize("rc", [False, True])
def test_default_path_matches_oracle(oracle, mem, rate, msg_len, L, dev, rc):
    w = synth.Workload("t", mem, rate, msg_len, L, dev, sub=0.01, dele=0.01, ins=0.005,
                       seed=5000 + mem * 10 + rate)
    _compare(oracle, w, 6, exact_only=False, rc=rc, tag=f"m{mem} r{rate} L{L} dev{dev} rc{rc}")


@pytest.mark.parametrize("exact_only", [True, False])
def test_quantised_posteriors_force_ties(oracle, exact_only):
    """Hazard corpus item 5: posteriors on a coarse grid make exact score ties common, so the
    pop order of libstdc++'s heap decides the lists."""
    for (mem, rate, msg_len, L) in [(6, 1, 40, 8), (6, 5, 40, 8), (8, 2, 40, 4)]:
        w = synth.Workload("q", mem, rate, msg_len, L, 20, sub=0.02, dele=0.02, ins=0.01,
                           seed=6000 + mem + rate, quantum=0.5, betas=(3.0, 2.0))
        tm = _compare(oracle, w, 6, exact_only=exact_only, tag=f"quantised m{mem} r{rate} L{L}")
        if not exact_only:  # the fast path must really have met ties and handed them to the exact merge
            assert tm["exact_fallbacks"] > 0, "quantised posteriors produced no tie fallback"


@pytest.mark.parametrize("mem,rate,msg_len,L", [(6, 1, 40, 4), (6, 5, 40, 8), (8, 3, 41, 3)])
def test_flat_posteriors_tie_everywhere(oracle, mem, rate, msg_len, L):
    """Every transition equally likely: ALL metrics of a step tie, three-way ties with duplicates included
    (a copy dropped by the fast merge may be the entry libstdc++'s heap pops first).  The list is then
    decided by heap order alone; the fast path must notice and hand every such state to the exact merge."""
    T = 3 * (msg_len + mem) + 7
    post = np.full((T, 40), np.float32(np.log(1.0 / 40.0)), dtype=np.float32)
    post2 = post.copy()
    post2[::3, 5] += np.float32(0.5)   # a few equal bumps: ties between paths of different shape
    post2[1::4, 33] += np.float32(0.5)
    for rc in (False, True):
        with ListDecoder(mem, rate, msg_len, L, 20, rc=rc) as dec:
            got = dec.decode([post, post2])
            tm = dec.timing()
        assert tm["exact_fallbacks"] > 0
        for pp, (bits, scores) in zip((post, post2), got):
            st, obits, oscores = oracle.oracle_decode(pp, mem, rate, msg_len, L, 20, rc)
            assert st == 0 and bits.shape == obits.shape and np.array_equal(bits, obits)
            assert np.array_equal(scores.view(np.uint32), oscores.view(np.uint32))


def test_unfused_band_flag_follows_the_unfused_oracle(oracle):
    """ADVICE: a reference built without FMA contraction bands some reads one row differently; the
    LVA_FLAG_UNFUSED_BAND handle must then follow the oracle's unfused variant, the default handle the
    fused one, on reads whose length makes the two differ."""
    P = 47   # nstate_pos of (mem 6, rate 1, msg_len 40)
    w = synth.Workload("band", 6, 1, 40, 4, 6, sub=0.01, dele=0.01, ins=0.005, seed=7300, betas=(3.0, 2.0))
    picked = []
    for i in range(400):
        r = synth.make_read(w, i)
        T = r.post.shape[0]
        if any(oracle.band_start(t, T, P, 6) != oracle.band_start(t, T, P, 6, unfused=True) for t in range(T)):
            picked.append(r)
        if len(picked) == 4:
            break
    assert picked, "no read length with a fused/unfused difference found"
    for unfused in (False, True):
        with ListDecoder(6, 1, 40, 4, 6, unfused_band=unfused) as dec:
            got = dec.decode([r.post for r in picked])
        for r, (bits, scores) in zip(picked, got):
            st, obits, oscores = oracle.oracle_decode(r.post, 6, 1, 40, 4, 6, unfused_band=unfused)
            assert st == 0 and bits.shape == obits.shape and np.array_equal(bits, obits)
            assert np.array_equal(scores.view(np.uint32), oscores.view(np.uint32))


def test_chain_table_overflow_is_detected_and_repeated(oracle, monkeypatch):
    """A chain table that is far too small (1024 slots) must never produce a wrong list: the probe chain
    runs away, the kernel raises the flag, the call is repeated with larger tables until the read fits."""
    monkeypatch.setenv("LVA_TAB_LOG2", "10")
    w = synth.Workload("tab", 6, 5, 130, 8, 20, sub=0.01, dele=0.01, ins=0.005, seed=7400, betas=(3.0, 2.0))
    reads = synth.make_reads(w, 4)
    with ListDecoder(6, 5, 130, 8, 20) as dec:
        got = dec.decode([r.post for r in reads])
        tm = dec.timing()
        again = dec.decode([r.post for r in reads])   # the grown tables are kept for the next call
        tm2 = dec.timing()
    assert tm["table_regrows"] >= 1 and tm["chain_table_slots"] > 1024 and tm["chain_keys_max"] > 1024
    assert tm2["table_regrows"] == 0
    for r, (bits, scores), (b2, s2) in zip(reads, got, again):
        st, obits, oscores = oracle.oracle_decode(r.post, 6, 5, 130, 8, 20)
        assert st == 0 and bits.shape == obits.shape and np.array_equal(bits, obits)
        assert np.array_equal(scores.view(np.uint32), oscores.view(np.uint32))
        assert np.array_equal(bits, b2) and np.array_equal(scores.view(np.uint32), s2.view(np.uint32))


def test_zero_and_tiny_deviation(oracle):
    """--max-deviation 0 (empty band: every list empty, like the reference) and 1 (a 3-row window)."""
    w = synth.Workload("dev", 6, 1, 40, 4, 20, sub=0.01, dele=0.01, ins=0.005, seed=7500)
    reads = synth.make_reads(w, 3)
    for dev in (0, 1, 2):
        with ListDecoder(6, 1, 40, 4, dev) as dec:
            got = dec.decode([r.post for r in reads])
        for r, (bits, scores) in zip(reads, got):
            st, obits, oscores = oracle.oracle_decode(r.post, 6, 1, 40, 4, dev)
            assert st == 0 and bits.shape == obits.shape and np.array_equal(bits, obits), dev
            assert np.array_equal(scores.view(np.uint32), oscores.view(np.uint32))


def test_more_reads_than_slots(oracle):
    w = synth.Workload("t", 6, 1, 40, 4, 20, sub=0.01, dele=0.01, ins=0.005, seed=7001)
    _compare(oracle, w, 11, exact_only=False, max_slots=4, tag="11 reads in 4 slots")


def test_too_small_posterior_is_reported():
    from nanopore_dna_storage_b200.decoder import LvaError
    with ListDecoder(6, 1, 40, 1, 20) as dec:
        post = np.zeros((dec.nstate_pos, 40), dtype=np.float32)  # needs nstate_pos + 1 blocks
        with pytest.raises(LvaError) as ei:
            dec.decode([post])
        assert ei.value.code == 21
        out = dec.decode([post], strict=False)
        assert out[0][0].shape[0] == 0


def test_gpu_crc_sidecar_matches_python_restatement():
    """north_star (e): CRC-8 / index check of all list entries on the GPU must equal the
    restatement of helper.decode_list_CRC_index (helper.py:371-388) applied to the emitted lists."""
    for name, rc in (("C4r2", False), ("C4r3", True)):   # msg_len 180 (no pad) and 181 (pad bit)
        w = synth.workloads()[name]
        reads = synth.make_reads(synth.Workload(w.name, w.mem_conv, w.rate, w.msg_len, w.list_size, 20,
                                                sub=0.004, dele=0.0085, ins=0.0005, seed=w.seed,
                                                rc_fraction=1.0 if rc else 0.0), 12)
        with ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, 20, rc=rc) as dec:
            got = dec.decode_with_crc([r.post for r in reads], pad=w.pad, num_oligos=4096)
            plain = dec.decode([r.post for r in reads])
        n_ok = 0
        for r, (bits, scores, ok, index, choice), (pbits, _) in zip(reads, got, plain):
            assert np.array_equal(bits, pbits)
            want_choice = -1
            for e in range(bits.shape[0]):
                wok, widx = crc_oracle.check_crc_index(bits[e], w.pad)
                assert bool(ok[e]) == wok and int(index[e]) == widx
                if want_choice < 0 and wok:
                    want_choice = e
            assert choice == want_choice
            if choice >= 0 and np.array_equal(bits[choice], r.msg):
                assert int(index[choice]) == r.index % 4096
                n_ok += 1
        assert n_ok >= 3, "expected several reads to decode to their framed message"


def test_survivor_mode_list_size_one(oracle):
    """L = 1 runs in survivor mode (metrics + 4-bit decisions + batched traceback, north_star d):
    must equal the oracle and the exact history-carrying path, incl. rc, full band, tight bands with
    stale rows, sync markers, CRC sidecar, and a batch larger than the slots."""
    for (mem, rate, msg_len, dev, rc) in [(6, 1, 60, 20, False), (6, 5, 64, 20, True), (8, 3, 41, 4, False),
                                          (8, 7, 41, None, True), (11, 5, 29, 20, False), (6, 2, 41, 3, False)]:
        w = synth.Workload("s", mem, rate, msg_len, 1, dev, sub=0.02, dele=0.02, ins=0.01, seed=8100 + mem + rate,
                           betas=(8.0, 3.0, 2.0))
        _compare(oracle, w, 9, exact_only=False, rc=rc, max_slots=4, tag=f"survivor m{mem} r{rate} dev{dev} rc{rc}")
    w = synth.workloads()["C1"]
    reads = synth.make_reads(w, 6)
    with ListDecoder(6, 1, 180, 1, 20) as dec:
        got = dec.decode_with_crc([r.post for r in reads], pad=w.pad, num_oligos=4096)
    with ListDecoder(6, 1, 180, 1, 20, exact_only=True) as dec:
        ref = dec.decode_with_crc([r.post for r in reads], pad=w.pad, num_oligos=4096)
    for a, b in zip(got, ref):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and a[4] == b[4]


def test_full_size_headline_config_properties():
    """BASELINE.json configs[1] at full size (m=11, r=5/6, msg_len=180, L=8, band 20), through
    properties that need no CPU decode: clean reads round-trip (encode -> posteriors -> decode gives
    the framed message back on top with a matching CRC and the right index), and the fast path equals
    the exact heap-order path bit for bit (lists and metrics) on noisy reads."""
    w = synth.workloads()["C2"]
    clean = synth.Workload(w.name, 11, 5, 180, 8, 20, seed=4242, betas=(8.0,))
    reads = synth.make_reads(clean, 6)
    with ListDecoder(11, 5, 180, 8, 20) as dec:
        got = dec.decode_with_crc([r.post for r in reads], pad=clean.pad, num_oligos=4096)
        noisy = synth.make_reads(w, 3, start=77)
        fast = dec.decode([r.post for r in noisy])
    for r, (bits, scores, ok, index, choice) in zip(reads, got):
        assert bits.shape[0] == 8 and np.array_equal(bits[0], r.msg)
        assert choice == 0 and ok[0] and int(index[0]) == r.index
        assert np.all(np.diff(scores) <= 0), "list must be sorted by metric"
        assert len({b.tobytes() for b in bits}) == 8, "list entries must be distinct messages"
    with ListDecoder(11, 5, 180, 8, 20, exact_only=True, max_slots=4) as dec:
        exact = dec.decode([r.post for r in noisy])
    for (b1, s1), (b2, s2) in zip(fast, exact):
        assert np.array_equal(b1, b2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))


@pytest.mark.parametrize("mem,rate,msg_len,L,marker,period,rc", [
    (6, 1, 40, 4, "110", 9, False), (6, 5, 40, 8, "01", 7, True), (8, 3, 40, 4, "1011", 12, False),
    (8, 5, 40, 8, "1", 5, True),
])
def test_sync_markers_match_oracle(oracle, mem, rate, msg_len, L, marker, period, rc):
    """--sync-marker / --sync-period pin message bits (is_valid_state, VCC:944-978): same lists and
    metrics as the oracle on posteriors of messages that do NOT carry the marker (hard reads)."""
    w = synth.Workload("sync", mem, rate, msg_len, L, 20, sub=0.01, dele=0.01, ins=0.005, seed=7100 + mem + rate)
    reads = [synth.make_read(w, i) for i in range(4)]
    for exact_only in (True, False):
        with ListDecoder(mem, rate, msg_len, L, 20, rc=rc, sync_marker=marker, sync_period=period,
                         exact_only=exact_only) as dec:
            got = dec.decode([r.post for r in reads])
        for r, (bits, scores) in zip(reads, got):
            st, obits, oscores = oracle.oracle_decode(r.post, mem, rate, msg_len, L, 20, rc, marker, period)
            assert st == 0
            assert bits.shape == obits.shape and np.array_equal(bits, obits), (mem, rate, marker, exact_only)
            assert np.array_equal(scores.view(np.uint32), oscores.view(np.uint32))


@pytest.mark.parametrize("mem,rate,msg_len,L", [(6, 1, 248, 2), (8, 5, 242, 4)])
def test_longest_messages_use_all_history_words(oracle, mem, rate, msg_len, L):
    """msg_len + mem_conv close to the 256-bit limit of the reference's bitset (VCC:604-605): the
    longest chains of interned chunks (7 completed chunks behind an entry of the last rows)."""
    from nanopore_dna_storage_b200.decoder import code_tables
    assert max(code_tables(mem, rate, msg_len)["history_words"]) == 8
    w = synth.Workload("long", mem, rate, msg_len, L, 20, sub=0.004, dele=0.0085, ins=0.0005, seed=7200 + mem)
    _compare(oracle, w, 2, exact_only=False, tag=f"long m{mem} r{rate}")
    _compare(oracle, w, 1, exact_only=True, tag=f"long exact m{mem} r{rate}")


def test_two_stage_merge_precondition_holds(oracle, tmp_path):
    """lva_fast.cuh merge_two_stage() relies on the move-in candidates of a state being pairwise distinct on
    every row the host did not mark (code_tables.cpp first_run_rows).  LVA_DBG bit 3 makes the kernel test
    that claim for every state it merges (the switch is read once per process, hence the subprocess); the
    same run must still reproduce the oracle, on hard reads (beta 3: lists full of near-equal alternatives),
    in both orientations and on a code whose first run is long."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import json, sys
import numpy as np
sys.path.insert(0, %r)
from nanopore_dna_storage_b200 import synth
from nanopore_dna_storage_b200.decoder import ListDecoder
from oracle import binding as ob
out = []
for (mem, rate, msg_len, L, rc) in [(11, 5, 60, 8, False), (11, 5, 60, 8, True), (8, 7, 60, 8, True), (6, 5, 100, 4, False), (8, 2, 100, 2, False), (8, 5, 60, 16, True)]:
    w = synth.Workload("t", mem, rate, msg_len, L, 20, sub=0.02, dele=0.02, ins=0.01, seed=7100 + mem + rate, betas=(3.0, 4.0))
    reads = [synth.make_read(w, i) for i in range(4)]
    with ListDecoder(mem, rate, msg_len, L, 20, rc=rc) as dec:
        got = dec.decode([r.post for r in reads])
        tm = dec.timing()
    same = True
    for r, (bits, scores) in zip(reads, got):
        st, obits, oscores = ob.oracle_decode(r.post, mem, rate, msg_len, L, 20, rc)
        same = same and st == 0 and bits.shape == obits.shape and np.array_equal(bits, obits) \
            and np.array_equal(scores.view(np.uint32), oscores.view(np.uint32))
    out.append(dict(cfg=[mem, rate, msg_len, L, rc], same=bool(same), failures=int(tm["structure_check_failures"])))
print("RESULT " + json.dumps(out))
''' % root
    env = dict(os.environ, LVA_DBG="8")
    p = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for rec in json.loads(line[7:]):
        assert rec["same"], rec
        assert rec["failures"] == 0, rec
