for wl in C5 C3 C4r2 C4r7; do
  echo "== $wl"
  python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(json.dumps({k:d[k] for k in ('value','ms_per_step','batch','kernel_ms_per_step','exact_fallbacks','top1_correct_frac','parity')}))
print(json.dumps({'e2e':d['e2e'],'roofline_frac':d['roofline']['frac'],'achieved':d['roofline']['achieved'],'alu':d['roofline']['alu']['frac']}))"
done
