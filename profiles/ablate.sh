#!/bin/bash
# developer A/B runs on one B200: same C2 batch, one knob per line (LVA_DBG switches: see lva_device.cuh;
# bit 1 = data movement only, bit 4 = no stores, bits 8.. = L2 prefetch distance in CTAs -- all but 0 give WRONG results)
for dbg in 0 2 18 $((592*256)); do
  echo "== LVA_DBG=$dbg"
  LVA_DBG=$dbg timeout 120 python profiles/quick.py C2 64 0 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('reads_per_s','kernel_ms','alg_GBps','exact_fallbacks','slots')})"
done
