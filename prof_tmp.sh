python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','kernel_ms_per_step','exact_fallbacks','step_kernel_launches']}, d['roofline']['frac'], d['e2e'])"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lva_step_fast -s 900 -c 3 -o gpurun_out/prof_r01f python bench.py --steps 1 --warmup 0 --reads-per-step 8 --slots 8 --no-cpu-baseline > gpurun_out/b2.log 2>&1
