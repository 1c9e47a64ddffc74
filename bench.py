#!/usr/bin/env python
"""bench.py -- reads decoded/s of the list-Viterbi hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU decoder

A "step" is one pass of the hot path over one batch of synthetic reads of the workload
(default C2 = m=11, r=5/6, msg_len=180, L=8, --max-deviation 20: BASELINE.json configs[1]).
`value`  : reads/s with the posteriors already resident in HBM (lva_decode_batch_device).
`e2e`    : reads/s through the reference-facing C ABI call with HOST buffers (lva_decode_batch):
           H2D of the posteriors and D2H of the candidate lists are inside the timed region.
Multi-GPU: one process per GPU (torchrun); the global read set is dealt to the ranks by the
product's own length-balanced plan (nanopore_dna_storage_b200/sharding.py), every rank decodes its
share (weak scaling: reads per GPU fixed, no data-path collective); time = max over ranks.

Outside the timed region the decoded lists of the batch are compared with what the UNMODIFIED
reference produced for the same reads: the committed fixtures (tests/golden/reference_lists_r2.json,
reads 0..15 of this workload) and a bounded live CPU sample (`cpu_baseline`).

Reference arm: K distinct reads of the same batch, throughput mode = as many
single-read processes as the host has cores, each `-t 1` (when K < cores the spare cores go to the
processes as OpenMP threads); a latency-mode figure (one process, all cores) is reported beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nanopore_dna_storage_b200 import synth  # noqa: E402

ISSUE_PEAK_THREAD_INSTR = 148 * 4 * 32  # per clock; x sm_max_mhz = thread-instructions/s


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--reads-per-step", type=int, default=0, help="per GPU; 0 = one full set of slots")
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reads", type=int, default=2, help="reads of the batch the live CPU sample decodes")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max((float(r[3]) for r in self.rows if len(r) > 3), default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


def golden_lists(w: synth.Workload):
    """{read index: [list lines]} of the committed reference fixtures made from exactly this workload."""
    out = {}
    try:
        cases = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_lists_r2.json")))["cases"]
    except Exception:
        return out
    for c in cases:
        k = c["workload"]
        if (k["mem_conv"], k["rate"], k["msg_len"], k["list_size"], k["max_deviation"], k["seed"]) == \
                (w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, w.seed) and \
                abs(k.get("rc_fraction", 0.0) - w.rc_fraction) < 1e-9:
            out[c["read"]] = (c["lists"], c["rc"])
    return out


def lines_of(bits: np.ndarray):
    return ["".join(map(str, b.tolist())) for b in bits]


def workload_config(args, w):
    return {"workload": f"{w.name} max_deviation={w.max_deviation} (BASELINE.json configs[1])"
            if args.workload == "C2" else w.name,
            "mem_conv": w.mem_conv, "rate": w.rate, "msg_len": w.msg_len, "list_size": w.list_size,
            "max_deviation": w.max_deviation, "rc_fraction": w.rc_fraction,
            "channel": {"sub": w.sub, "del": w.dele, "ins": w.ins}, "sigma": w.sigma, "betas": list(w.betas)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference CLI (oracle/_ref/viterbi_nanopore.out, compiled in place from
# /root/reference by oracle/Makefile) on .post files, as its own drivers call it
# ---------------------------------------------------------------------------------------------------
def ref_cli_decode_many(w, reads, procs: int, threads: int, tmp: str):
    """Decode `reads` with `procs` concurrent reference processes of `threads` OpenMP threads each.
    -> (wall seconds, {read index: list lines})"""
    from oracle import binding as ob
    exe = ob.REF_EXE
    jobs = []
    for r in reads:
        pin, pout = os.path.join(tmp, f"r{r.index}.post"), os.path.join(tmp, f"r{r.index}.dec")
        r.post.tofile(pin)
        jobs.append((r, [exe, "-m", "decode", "-i", pin, "-o", pout, "--msg-len", str(w.msg_len), "-l", str(w.list_size),
                         "-t", str(threads), "--mem-conv", str(w.mem_conv), "--rc" if r.rc else "", "--max-deviation",
                         str(w.max_deviation), "-r", str(w.rate)], pout))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    t0 = time.perf_counter()
    running, it, out = [], iter(jobs), {}
    while True:
        while len(running) < procs:
            j = next(it, None)
            if j is None:
                break
            running.append((j, subprocess.Popen(j[1], stdout=subprocess.DEVNULL, env=env)))
        if not running:
            break
        for k, (j, p) in enumerate(running):
            if p.poll() is not None:
                assert p.returncode == 0, f"reference decoder failed on read {j[0].index}"
                with open(j[2]) as f:
                    out[j[0].index] = f.read().split()
                running.pop(k)
                break
        else:
            time.sleep(0.05)
    return time.perf_counter() - t0, out


def oracle_port_decode_many(w, reads, threads):
    from oracle import binding as ob
    t0 = time.perf_counter()
    out = {}
    for r in reads:
        st, bits, _ = ob.oracle_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, r.rc,
                                       num_thr=threads)
        assert st == 0
        out[r.index] = lines_of(bits)
    return time.perf_counter() - t0, out


def reference_arm(args, w, cfg):
    from oracle import binding as ob
    cores = os.cpu_count() or 1
    K = max(1, args.steps)
    reads = synth.make_reads(w, K, start=0)  # the first K reads of rank 0's GPU batch
    kind = "reference" if ob.have_ref() else "port"
    gold = golden_lists(w)
    with tempfile.TemporaryDirectory() as tmp:
        if kind == "reference":
            # throughput mode: waves of single-read processes; spare cores become OpenMP threads of the processes
            t_all, lists, waves, left = 0.0, {}, [], list(reads)
            while left:
                procs = min(len(left), cores)
                thr = max(1, cores // procs)
                dt, out = ref_cli_decode_many(w, left[:procs], procs, thr, tmp)
                waves.append({"processes": procs, "threads_each": thr, "seconds": round(dt, 1)})
                t_all += dt
                lists.update(out)
                left = left[procs:]
            # latency mode beside it: one process, every core, one read
            dt_lat, out_lat = ref_cli_decode_many(w, reads[:1], 1, cores, tmp)
            assert out_lat[reads[0].index] == lists[reads[0].index], "reference lists depend on -t"
            latency = {"value": 1.0 / dt_lat, "unit": "reads/s", "sample": f"1 read, one process -t {cores}, {dt_lat:.1f} s"}
            sample = (f"{K} distinct reads (indices 0..{K - 1} of the GPU batch), throughput mode: "
                      + "; ".join(f"{x['processes']} processes x -t {x['threads_each']} ({x['seconds']} s)" for x in waves))
        else:
            t_all, lists = oracle_port_decode_many(w, reads, cores)
            latency, sample = None, f"{K} distinct reads, oracle port, one process -t {cores}"
    checked = 0
    for r in reads:
        if r.index in gold:
            assert lists[r.index] == gold[r.index][0], f"reference arm disagrees with the committed fixture of read {r.index}"
            checked += 1
    value = K / t_all
    print(json.dumps({"impl": "reference", "metric": "reads decoded/s", "value": value, "unit": "reads/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * t_all / K, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                      "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
                      "latency_mode": latency, "reads_checked_against_fixtures": checked,
                      "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = synth.workloads()[args.workload]
    cfg = workload_config(args, w)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, w, cfg)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    from nanopore_dna_storage_b200.decoder import ListDecoder
    from nanopore_dna_storage_b200.sharding import plan

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # one decoder per orientation present in the workload (C5 decodes forward and reverse-complement reads)
    orients = [False] if w.rc_fraction <= 0 else ([True] if w.rc_fraction >= 1 else [False, True])
    decs = {rc: ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, rc=rc, device=local_rank,
                            max_slots=args.slots) for rc in orients}
    slots = min(d.slots for d in decs.values())
    B = args.reads_per_step or slots
    # global read set 0 .. world*B-1, dealt by the product's length-balanced plan; rank r decodes share r
    all_reads = synth.make_reads(w, B * world, start=0)
    mine = plan([r.post.shape[0] for r in all_reads], world)[rank]
    reads = [all_reads[i] for i in mine]
    groups = {rc: [r for r in reads if r.rc == rc] for rc in orients}
    batches = {}
    for rc, rs in groups.items():
        if not rs:
            continue
        offsets = np.zeros(len(rs) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([r.post.shape[0] for r in rs])
        flat = np.concatenate([r.post for r in rs], axis=0)
        host_pinned = torch.empty(flat.shape, dtype=torch.float32).pin_memory()
        host_pinned.copy_(torch.from_numpy(flat))
        batches[rc] = dict(reads=rs, offsets=offsets, host=host_pinned, host_np=host_pinned.numpy(),
                           dev=host_pinned.cuda(non_blocking=False))
    work = [decs[r.rc].work(int(r.post.shape[0])) for r in reads]
    step_bytes = sum(x["state_bytes"] for x in work)
    step_acs = sum(x["edge_acs"] for x in work)
    step_updates = sum(x["state_updates"] for x in work)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(device_resident: bool, acc: dict):
        outs = {}
        for rc, b in batches.items():
            if device_resident:
                outs[rc] = decs[rc].decode_packed(None, b["offsets"], device_ptr=b["dev"].data_ptr())
            else:
                outs[rc] = decs[rc].decode_packed(b["host_np"], b["offsets"])
            tm = decs[rc].timing()
            for k in ("kernel_ms", "total_ms", "launches", "step_launches", "exact_fallbacks", "h2d_bytes", "d2h_bytes",
                      "chain_inserts"):
                acc[k] = acc.get(k, 0) + tm[k]
            acc["chain_keys_max"] = max(acc.get("chain_keys_max", 0), tm["chain_keys_max"])
            acc["chain_table_slots"] = tm["chain_table_slots"]
        return outs

    for _ in range(args.warmup):
        run(True, {})
    # ---- timed: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_acc = {}
    for _ in range(args.steps):
        out_dev = run(True, dev_acc)
    barrier()
    wall = time.perf_counter() - t0
    # ---- timed: host buffers through the C ABI (e2e) ----
    for _ in range(min(args.warmup, 1)):
        run(False, {})
    barrier()
    t1 = time.perf_counter()
    host_acc = {}
    for _ in range(args.steps):
        out_host = run(False, host_acc)
    barrier()
    wall_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    # ---- outside the timed region: parity of what was just decoded ----
    lists = {}
    for rc, b in batches.items():
        assert np.array_equal(out_dev[rc][0], out_host[rc][0]) and np.array_equal(out_dev[rc][2], out_host[rc][2])
        bits, _, counts, _ = out_dev[rc]
        for k, r in enumerate(b["reads"]):
            lists[r.index] = lines_of(bits[k, :counts[k]])
    gold = golden_lists(w)
    fixture_checked = 0
    for r in reads:
        if r.index in gold:
            assert lists[r.index] == gold[r.index][0], f"GPU list of read {r.index} differs from the reference fixture"
            fixture_checked += 1
    ncorrect = sum(1 for r in reads if lists[r.index] and lists[r.index][0] == "".join(map(str, r.msg.tolist())))

    times = torch.tensor([wall, wall_e2e, dev_acc["kernel_ms"], dev_acc["total_ms"]], dtype=torch.float64, device="cuda")
    counts_t = torch.tensor([len(reads), fixture_checked, ncorrect], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts_t, op=dist.ReduceOp.SUM)
    wall, wall_e2e, kernel_ms, total_ms = [float(x) for x in times.tolist()]
    n_reads_all, fixture_all, ncorrect_all = [int(x) for x in counts_t.tolist()]

    if rank == 0:
        n_total = n_reads_all * args.steps
        value = n_total / wall
        peak, sm_max_mhz, peak_src = measured_peaks()
        achieved = step_bytes * args.steps / (kernel_ms * 1e-3) / 1e9  # per GPU: bytes of one rank / its kernel time
        tr = committed_json("traffic.json") if args.workload == "C2" else None  # the committed capture is a C2 launch
        step_launches = int(dev_acc["step_launches"])
        cand_acs_per_s = step_acs * w.list_size * args.steps / (kernel_ms * 1e-3)  # one GPU, kernel time
        alu_peak = ISSUE_PEAK_THREAD_INSTR * sm_max_mhz * 1e6 / 2.0  # 2 thread-instructions per candidate-ACS
        line = {
            "metric": "reads decoded/s", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "batch": {"reads_per_step_per_gpu": B, "slots": slots, "orientations": len(batches),
                      "l2_note": "state ping-pong per step (8 GB at 64 reads) >> 126 MB L2: no flush needed between steps",
                      "sharding": "nanopore_dna_storage_b200.sharding.plan over the global read set"},
            "acs_updates_per_s": step_acs * args.steps * world / wall,
            "state_updates_per_s": step_updates * args.steps * world / wall,
            "device_ms_per_step": total_ms / args.steps, "kernel_ms_per_step": kernel_ms / args.steps,
            "gpu_launches": int(dev_acc["launches"] + host_acc["launches"]), "step_kernel_launches": step_launches,
            "exact_fallbacks": int(dev_acc["exact_fallbacks"]), "top1_correct_frac": ncorrect_all / n_reads_all,
            "chain_tables": {"keys_per_read_max": int(dev_acc["chain_keys_max"]),
                             "keys_per_read_mean": dev_acc["chain_inserts"] / max(1, len(reads) * args.steps),
                             "slots_per_read": int(dev_acc["chain_table_slots"])},
            "parity": {"reads_checked_against_reference_fixtures": fixture_all,
                       "fixtures": "tests/golden/reference_lists_r2.json (unmodified reference, same reads)"},
            "e2e": {"value": n_total / wall_e2e, "unit": "reads/s",
                    "h2d_bytes_per_step": int(host_acc["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": int(host_acc["d2h_bytes"]) // args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "hbm_busy_frac": (achieved / peak) * (tr["dram_bytes_per_launch"] / tr["algorithmic_bytes_same_launch"]) if tr else None,
                         "hbm_busy_note": "frac x (measured DRAM traffic / algorithmic bytes of the committed capture): how busy the "
                                          "launch keeps HBM; the gap to frac is re-read and partial-write traffic, not useful work",
                         "kernel": ("lva_step_l1" if w.list_size == 1 else "lva_step_fast<LP>" if w.list_size <= 16 else "lva_step_wide<LP>") +
                                   " (every trellis step launch)",
                         "algorithmic_bytes_per_step": step_bytes,
                         "avg_launch_us": 1e3 * kernel_ms / max(step_launches, 1),
                         "traffic": tr.get("dram_bytes_per_launch") if tr else None,
                         "traffic_launch": tr.get("kernel") if tr else None,
                         "traffic_algorithmic_same_launch": tr.get("algorithmic_bytes_same_launch") if tr else None,
                         "traffic_note": tr.get("note") if tr else None,
                         "alu": {"cand_acs_per_s": cand_acs_per_s, "peak": alu_peak, "frac": cand_acs_per_s / alu_peak,
                                 "note": "SURVEY.md 8(d) issue-slot ceiling: 2 thread-instructions per candidate-ACS"}},
            "clocks": clocks,
        }
        pipes = committed_json("pipes.json")
        if pipes and args.workload == "C2":
            line["roofline"]["issue_committed_capture"] = pipes
        if world == 1 and not args.no_cpu_baseline:
            # bounded live CPU sample: reads of THIS batch through the unmodified reference (latency mode keeps
            # it to ~15 s per read), lists compared byte for byte with what the GPU just produced
            from oracle import binding as ob
            sample_reads = reads[:max(1, args.cpu_reads)]
            with tempfile.TemporaryDirectory() as tmp:
                if ob.have_ref():
                    dt, out = ref_cli_decode_many(w, sample_reads, 1, cores, tmp)
                    kind = "reference"
                else:
                    dt, out = oracle_port_decode_many(w, sample_reads, cores)
                    kind = "port"
            for r in sample_reads:
                assert out[r.index] == lists[r.index], f"GPU list of read {r.index} differs from the CPU {kind}"
            line["cpu_baseline"] = {"value": len(sample_reads) / dt, "unit": "reads/s", "cores": cores, "kind": kind,
                                    "sample": f"reads {[r.index for r in sample_reads]} of the GPU batch, one process -t {cores}, "
                                              f"{dt:.1f} s; lists byte-equal to the GPU output"}
        print(json.dumps(line))
    for d in decs.values():
        d.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
