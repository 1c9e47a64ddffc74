#!/usr/bin/env python
"""bench.py -- reads decoded/s of the list-Viterbi hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU decoder

A "step" is one pass of the hot path over one batch of synthetic reads of the workload
(default C2 = m=11, r=5/6, msg_len=180, L=8, --max-deviation 20: BASELINE.json configs[1]).
`value`  : reads/s with the posteriors already resident in HBM (lva_decode_batch_device).
`e2e`    : reads/s through the reference-facing C ABI call with HOST buffers (lva_decode_batch):
           H2D of the posteriors and D2H of the candidate lists are inside the timed region.
Multi-GPU: one process per GPU (torchrun), every rank decodes its own reads (weak scaling, no
data-path collective); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nanopore_dna_storage_b200 import synth  # noqa: E402


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
    ap.add_argument("--cpu-reads", type=int, default=1, help="reads in the CPU baseline sample")
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


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def ncu_pipes():
    """Issue / pipe utilisation of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "pipes.json")) as f:
            return json.load(f)
    except Exception:
        return None


def cpu_reference_run(w: synth.Workload, reads, threads: int):
    """Time the reference's own decoder (oracle/_ref, unmodified sources compiled in place) on a
    bounded sample; falls back to the oracle port if the compiled reference is absent."""
    from oracle import binding as ob
    kind = "reference" if ob.have_ref() else "port"
    t0 = time.perf_counter()
    for r in reads:
        if kind == "reference":
            st, _ = ob.ref_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, r.rc,
                                  num_thr=threads)
        else:
            st, _, _ = ob.oracle_decode(r.post, w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, r.rc,
                                        num_thr=threads)
        assert st == 0
    dt = time.perf_counter() - t0
    return kind, len(reads) / dt, dt


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = synth.workloads()[args.workload]
    cfg = {"workload": f"{w.name} max_deviation={w.max_deviation} (BASELINE.json configs[1])"
           if args.workload == "C2" else w.name,
           "mem_conv": w.mem_conv, "rate": w.rate, "msg_len": w.msg_len, "list_size": w.list_size,
           "max_deviation": w.max_deviation,
           "channel": {"sub": w.sub, "del": w.dele, "ins": w.ins}, "sigma": w.sigma, "betas": list(w.betas)}
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        reads = synth.make_reads(w, max(1, args.cpu_reads), start=10_000_000)
        for _ in range(min(args.warmup, 1)):
            pass  # a warm-up decode would cost as much as a timed one; the CPU path has no cold state
        t_all, n_all = 0.0, 0
        kind = "port"
        for _ in range(args.steps):
            kind, rps, dt = cpu_reference_run(w, reads, threads)
            t_all += dt
            n_all += len(reads)
        value = n_all / t_all
        sample = f"{len(reads)} read(s) per step, -t {threads} (one process, OpenMP over trellis rows)"
        print(json.dumps({"impl": "reference", "metric": "reads decoded/s", "value": value, "unit": "reads/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": kind,
                                           "sample": sample},
                          "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    from nanopore_dna_storage_b200.decoder import ListDecoder

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dec = ListDecoder(w.mem_conv, w.rate, w.msg_len, w.list_size, w.max_deviation, device=local_rank,
                      max_slots=args.slots)
    B = args.reads_per_step or dec.slots
    reads = synth.make_reads(w, B, start=rank * 1_000_000)
    offsets = np.zeros(B + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([r.post.shape[0] for r in reads])
    flat = np.concatenate([r.post for r in reads], axis=0)
    host_pinned = torch.empty(flat.shape, dtype=torch.float32).pin_memory()
    host_pinned.copy_(torch.from_numpy(flat))
    host_np = host_pinned.numpy()
    dev_post = host_pinned.cuda(non_blocking=False)
    work = [dec.work(int(r.post.shape[0])) for r in reads]
    step_bytes = sum(x["state_bytes"] for x in work)
    step_acs = sum(x["edge_acs"] for x in work)
    step_updates = sum(x["state_updates"] for x in work)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run_device():
        return dec.decode_packed(None, offsets, device_ptr=dev_post.data_ptr())

    def run_host():
        return dec.decode_packed(host_np, offsets)

    for _ in range(args.warmup):
        out_dev = run_device()
    # ---- timed: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    kernel_ms = total_ms = 0.0
    launches = step_launches = fallbacks = 0
    for _ in range(args.steps):
        out_dev = run_device()
        tm = dec.timing()
        kernel_ms += tm["kernel_ms"]; total_ms += tm["total_ms"]
        launches += tm["launches"]; step_launches += tm["step_launches"]; fallbacks += tm["exact_fallbacks"]
    barrier()
    wall = time.perf_counter() - t0
    # ---- timed: host buffers through the C ABI (e2e) ----
    for _ in range(min(args.warmup, 1)):
        run_host()
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        out_host = run_host()
        tm = dec.timing()
        h2d += tm["h2d_bytes"]; d2h += tm["d2h_bytes"]
    barrier()
    wall_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(out_dev[0], out_host[0]) and np.array_equal(out_dev[2], out_host[2])

    times = torch.tensor([wall, wall_e2e, kernel_ms, total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, wall_e2e, kernel_ms, total_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        n_total = B * args.steps * world
        value = n_total / wall
        peak, peak_src = measured_peak_hbm()
        achieved = step_bytes * args.steps / (kernel_ms * 1e-3) / 1e9  # per GPU: bytes of one rank / its kernel time
        tr = ncu_traffic()
        ncorrect = sum(1 for i, r in enumerate(reads)
                       if out_dev[2][i] > 0 and np.array_equal(out_dev[0][i, 0], r.msg))
        line = {
            "metric": "reads decoded/s", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(cfg, reads_per_step_per_gpu=B, slots=dec.slots,
                                                l2_note="state ping-pong per step >> 126 MB L2 (no reuse across steps)"),
            "acs_updates_per_s": step_acs * args.steps * world / wall,
            "state_updates_per_s": step_updates * args.steps * world / wall,
            "device_ms_per_step": total_ms / args.steps, "kernel_ms_per_step": kernel_ms / args.steps,
            "gpu_launches": int(launches), "step_kernel_launches": int(step_launches),
            "exact_fallbacks": int(fallbacks), "top1_correct_frac": ncorrect / B,
            "e2e": {"value": n_total / wall_e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "kernel": "lva_step_fast<LP,NWD> (all step launches)",
                         "algorithmic_bytes_per_step": step_bytes,
                         "avg_launch_us": 1e3 * kernel_ms / max(step_launches, 1),
                         "traffic": tr.get("dram_bytes_per_launch") if tr else None,
                         "traffic_note": (f"ncu capture of {tr['kernel']}: {tr['dram_bytes_per_launch']} B DRAM vs "
                                          f"{tr['algorithmic_bytes_same_launch']} B algorithmic for that launch") if tr else None},
            "clocks": clocks,
        }
        pipes = ncu_pipes()
        if pipes and args.workload == "C2":
            line["roofline"]["issue"] = {k: pipes[k] for k in ("issue_active_pct", "alu_pipe_pct", "fma_pipe_pct",
                                                               "lsu_pipe_pct", "source") if k in pipes}
        if world == 1 and not args.no_cpu_baseline:
            cpu_reads = synth.make_reads(w, max(1, args.cpu_reads), start=10_000_000)
            kind, rps, dt = cpu_reference_run(w, cpu_reads, threads)
            line["cpu_baseline"] = {"value": rps, "unit": "reads/s", "cores": threads, "kind": kind,
                                    "sample": f"{len(cpu_reads)} read(s) of the same workload, one process -t {threads}, {dt:.1f} s"}
        print(json.dumps(line))
    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
