#!/usr/bin/env python
"""Like ncu_lines.py but ranked by warp-stall samples. usage: ncu_samples.py src.csv <kernel substring> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
want = sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
agg, tot, i, cur, active, started, hdr = {}, 0, 0, None, False, False, None
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
        tot += samp
    i += 1
print("total samples", tot)
for (f, l), (src, inst, thr, samp) in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top_n]:
    print(f"{f:16s}{l:5d} samples {100 * samp / tot:5.1f}% inst {inst:9d} thr/inst {thr / max(inst, 1):5.1f} | {src.strip()[:95]}")
