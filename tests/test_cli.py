"""Drop-in contract of the CLI: same argv grammar, stdout text, exit codes and
files as the reference's viterbi_nanopore.out.  Compared side by side with the compiled reference
where oracle/_ref exists; the decode comparison needs a GPU."""
import os
import subprocess

import numpy as np
import pytest

from nanopore_dna_storage_b200 import _lib, synth
from oracle import binding as ob

OURS = _lib.CLI_PATH
REF = ob.REF_EXE
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="compiled reference absent")


def run(exe, args, cwd):
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True)
    return p.returncode, p.stdout, p.stderr


def both(args, cwd):
    return run(OURS, args, cwd), run(REF, args, cwd)


@needs_ref
@pytest.mark.parametrize("args", [
    ["--help"], ["-h"], [], ["-m", "decode"], ["-m", "encode", "-i", "a", "-o", "b"],
    ["-m", "encode", "-i", "a", "-o", "b", "--msg-len", "40"],
    ["-m", "encode", "-i", "a", "-o", "b", "--mem-conv", "6"],
    ["-m", "encode", "-i", "a", "-o", "b", "--mem-conv", "7", "--msg-len", "40"],
    ["-m", "encode", "-i", "a", "-o", "b", "--mem-conv", "6", "--msg-len", "40", "-r", "6"],
    ["-m", "encode", "-i", "a", "-o", "b", "--mem-conv", "8", "--msg-len", "180", "-r", "3"],
    ["-m", "frobnicate", "-i", "a", "-o", "b", "--mem-conv", "6", "--msg-len", "40"],
    ["--mode=encode", "--infile=a", "--outfile=b", "--mem-conv=7", "--msg-len=40"],
    ["-m", "decode", "-i", "a", "-o", "b", "--mem-conv", "6", "--msg-len", "40", "--sync-marker", "1x0",
     "--sync-period", "5"],
    ["-m", "decode", "-i", "a", "-o", "b", "--mem-conv", "6", "--msg-len", "40", "--sync-marker", "110",
     "--sync-period", "2", "--rc"],
])
def test_usage_and_parameter_errors_match_reference(tmp_path, args):
    (rc1, out1, _), (rc2, out2, _) = both(args, tmp_path)
    assert rc1 == rc2 and out1 == out2


@needs_ref
@pytest.mark.parametrize("args", [
    ["--bogus"], ["-m"], ["-x", "1"], ["-m", "encode", "--msg-len", "abc", "-i", "a", "-o", "b"],
    ["-m", "encode", "--mem-conv", "300", "-i", "a", "-o", "b"], ["-m", "encode", "--msg-len", "-4"],
    ["-", "-m"], ["-m", "decode", "-i", "a", "-o", "b", "--mem-conv", "6", "--msg-len", "40", "--sync-marker", "11"],
])
def test_malformed_argv_aborts_like_reference(tmp_path, args):
    (rc1, out1, err1), (rc2, out2, err2) = both(args, tmp_path)
    assert rc2 == -6, "reference is expected to die of SIGABRT on an uncaught exception"
    assert rc1 == rc2 and out1 == out2
    assert err1.splitlines()[:2] == err2.splitlines()[:2]  # same exception type and what()


@needs_ref
def test_encode_matches_reference_byte_for_byte(tmp_path):
    rng = np.random.default_rng(4)
    for mem, rate, msg_len in [(6, 1, 40), (8, 3, 181), (11, 5, 180), (14, 1, 100), (8, 7, 181), (6, 2, 41), (8, 4, 180)]:
        msgs = rng.integers(0, 2, size=(5, msg_len))
        inp = tmp_path / f"in_{mem}_{rate}.txt"
        inp.write_text("".join("".join(map(str, m)) + "\n" for m in msgs) + "0101")  # unterminated tail is dropped
        a, b = tmp_path / "ours.txt", tmp_path / "ref.txt"
        args = ["-m", "encode", "-i", str(inp), "--mem-conv", str(mem), "--msg-len", str(msg_len), "-r", str(rate), "-t", "3", ""]
        assert run(OURS, args + ["-o", str(a)], tmp_path)[0] == 0
        assert run(REF, args + ["-o", str(b)], tmp_path)[0] == 0
        assert a.read_bytes() == b.read_bytes() and len(a.read_text().splitlines()) == 5
    # wrong message length and bad characters
    inp = tmp_path / "bad.txt"
    inp.write_text("0101\n")
    r1, r2 = both(["-m", "encode", "-i", str(inp), "-o", "x", "--mem-conv", "6", "--msg-len", "40"], tmp_path)
    assert r1[:2] == r2[:2] and r1[0] == 255
    inp.write_text("01a1\n")
    r1, r2 = both(["-m", "encode", "-i", str(inp), "-o", "x", "--mem-conv", "6", "--msg-len", "4"], tmp_path)
    assert r1[0] == r2[0] == -6


def test_cli_has_no_cpu_decode(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    post = tmp_path / "p.post"
    np.zeros((200, 40), np.float32).tofile(post)
    rc, out, err = run(OURS, ["-m", "decode", "-i", str(post), "-o", str(tmp_path / "o"), "--mem-conv", "6",
                             "--msg-len", "40"], tmp_path)
    assert rc == 1 and b"no CUDA device" in err and not (tmp_path / "o").exists()


@pytest.mark.gpu
@needs_ref
def test_decode_cli_is_a_drop_in(tmp_path):
    """The argv the Python drivers build (generate_decoded_lists.py:86-90): literal '' when not rc,
    -t N, --max-deviation 20; list files must be byte-identical to the reference's."""
    cases = [(6, 1, 40, 1, False), (6, 5, 40, 8, True), (8, 3, 41, 4, False)]
    for k, (mem, rate, msg_len, L, rc) in enumerate(cases):
        w = synth.Workload("cli", mem, rate, msg_len, L, 20, sub=0.01, dele=0.01, ins=0.005, seed=900 + k,
                           rc_fraction=1.0 if rc else 0.0)
        r = synth.make_read(w, 0)
        post = tmp_path / f"r{k}.post"
        r.post.tofile(post)
        # a trailing partial block, as a truncated file would have
        with open(post, "ab") as f:
            f.write(np.float32([-1.5, -2.5]).tobytes())
        args = ["-m", "decode", "-i", str(post), "--mem-conv", str(mem), "--msg-len", str(msg_len), "-l", str(L),
                "-t", "4", "-r", str(rate), "--rc" if rc else "", "--max-deviation", "20"]
        a, b = tmp_path / f"ours{k}", tmp_path / f"ref{k}"
        r1 = run(OURS, args + ["-o", str(a)], tmp_path)
        r2 = run(REF, args + ["-o", str(b)], tmp_path)
        assert r1[0] == r2[0] == 0 and r1[1] == r2[1]
        assert a.read_bytes() == b.read_bytes()
    # too small a posterior: both abort
    post = tmp_path / "short.post"
    np.zeros((10, 40), np.float32).tofile(post)
    args = ["-m", "decode", "-i", str(post), "-o", str(tmp_path / "s"), "--mem-conv", "6", "--msg-len", "40"]
    assert run(OURS, args, tmp_path)[0] == run(REF, args, tmp_path)[0] == -6


@pytest.mark.gpu
def test_batch_manifest_orientations_devices_and_directory_mode(tmp_path, oracle):
    """The batch host of the CLI: a manifest mixing forward and reverse-complement reads (third column),
    --devices (every GPU of the box; one worker thread and one handle per GPU and orientation), the
    directory mode with the reference driver's <prefix>_<i> naming (generate_decoded_lists.py:85) and the
    per-read info file.  Every list file is compared with the oracle."""
    import torch
    w = synth.Workload("cli", 6, 5, 40, 4, 20, sub=0.01, dele=0.01, ins=0.005, seed=951, rc_fraction=0.5)
    reads = [synth.make_read(w, i) for i in range(9)]
    assert any(r.rc for r in reads) and not all(r.rc for r in reads)
    d = tmp_path / "posts"
    d.mkdir()
    lines = []
    for i, r in enumerate(reads):
        p = d / f"read_{i:03d}.post"
        r.post.tofile(p)
        if r.rc:
            (d / f"read_{i:03d}.post.rc").write_text("")
        lines.append(f"{p} {tmp_path / f'm_{i}'}" + (" rc" if r.rc else ""))
    (tmp_path / "batch.txt").write_text("\n".join(lines) + "\n")
    want = []
    for r in reads:
        st, bits, _ = oracle.oracle_decode(r.post, 6, 5, 40, 4, 20, r.rc)
        want.append("".join("".join(map(str, b)) + "\n" for b in bits))
    common = ["-m", "decode", "--mem-conv", "6", "--msg-len", "40", "-l", "4", "-r", "5", "--max-deviation", "20"]
    ngpu = torch.cuda.device_count()
    for devs in (["--device", "0"], ["--devices", "all"], ["--devices", f"0-{ngpu - 1}"]):
        for f in tmp_path.glob("m_*"):
            f.unlink()
        rc, out, err = run(OURS, common + ["--batch-list", str(tmp_path / "batch.txt")] + devs, tmp_path)
        assert rc == 0, err
        for i in range(len(reads)):
            assert (tmp_path / f"m_{i}").read_text() == want[i], (devs, i)
    rc, out, err = run(OURS, common + ["--post-dir", str(d), "--out-prefix", str(tmp_path / "list"), "--info-file",
                                       str(tmp_path / "info.txt"), "--devices", "all"], tmp_path)
    assert rc == 0, err
    info = (tmp_path / "info.txt").read_text().splitlines()
    assert len(info) == len(reads)
    for i, r in enumerate(reads):
        assert (tmp_path / f"list_{i}").read_text() == want[i]
        cols = info[i].split("\t")
        assert cols[0].endswith(f"read_{i:03d}.post") and cols[2] == ("rc" if r.rc else "fwd")
        assert int(cols[3]) == r.post.shape[0] and int(cols[4]) == want[i].count("\n") and cols[5] == "ok"
    rc, out, err = run(OURS, common + ["--batch-list", str(tmp_path / "batch.txt"), "--devices", "9-3"], tmp_path)
    assert rc == 1 and b"--devices" in err


@pytest.mark.gpu
def test_batch_list_extension(tmp_path, oracle):
    w = synth.Workload("cli", 6, 1, 40, 4, 20, sub=0.01, dele=0.01, ins=0.005, seed=950)
    lines, reads = [], []
    for i in range(5):
        r = synth.make_read(w, i)
        p = tmp_path / f"b{i}.post"
        r.post.tofile(p)
        lines.append(f"{p} {tmp_path / f'list_{i}'}")
        reads.append(r)
    (tmp_path / "batch.txt").write_text("\n".join(lines) + "\n")
    rc, out, err = run(OURS, ["-m", "decode", "--batch-list", str(tmp_path / "batch.txt"), "--mem-conv", "6",
                             "--msg-len", "40", "-l", "4", "--max-deviation", "20"], tmp_path)
    assert rc == 0, err
    for i, r in enumerate(reads):
        st, bits, _ = oracle.oracle_decode(r.post, 6, 1, 40, 4, 20)
        want = "".join("".join(map(str, b)) + "\n" for b in bits)
        assert (tmp_path / f"list_{i}").read_text() == want
