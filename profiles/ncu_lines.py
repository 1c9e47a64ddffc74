#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; ncu_lines.py src.csv <kernel substring> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
want = sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr = None
agg, tot_inst, tot_thr, tot_samp = {}, 0, 0, 0
i, cur, active, started = 0, None, False, False
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        fn = rows[i + 1][1]
        active = want in fn
        if started and not active:
            break
        started |= active
        cur = r[1].split("/")[-1]
        hdr = rows[i + 2]
        i += 3
        continue
    if active and r and r[0].isdigit() and hdr:
        n = len(hdr)
        tail = r[-(n - 2):]
        src = ",".join(r[1:len(r) - (n - 2)])
        try:
            inst, thr, samp = int(tail[5]), int(tail[6]), int(tail[2])
        except ValueError:
            i += 1
            continue
        k = (cur, int(r[0]))
        a = agg.get(k, [src, 0, 0, 0])
        a[1] += inst; a[2] += thr; a[3] += samp
        agg[k] = a
        tot_inst += inst; tot_thr += thr; tot_samp += samp
    i += 1
print(f"kernel~{want}: warp-inst {tot_inst}  thread-inst {tot_thr}  avg active threads {tot_thr / max(tot_inst, 1):.2f}  samples {tot_samp}")
for (f, l), (src, inst, thr, samp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f"{f:16s}{l:5d} inst {100 * inst / tot_inst:5.1f}%  thr/inst {thr / max(inst, 1):5.1f}  samples {100 * samp / max(tot_samp, 1):5.1f}% | {src.strip()[:100]}")
