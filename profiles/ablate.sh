#!/bin/bash
# developer A/B runs on one B200: same C2 batch, one knob per line (LVA_DBG ablations give WRONG results on purpose)
for cfg in "2 0" "6 0" "10 0" "14 0" "18 0" "30 0"; do
  set -- $cfg
  echo "== LVA_DBG=$1 slots=$2"
  LVA_DBG=$1 timeout 120 python profiles/quick.py C2 64 $2 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('reads_per_s','kernel_ms','alg_GBps','exact_fallbacks','slots')})"
done
for d in 2 14; do
LVA_DBG=$d timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:lva_step_fast -s 300 -c 2 --csv python profiles/quick.py C2 8 8 1 2>&1 | grep -E "lva_step_fast" | cut -d, -f 5,12- 
done
