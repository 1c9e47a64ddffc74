mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 0 --reads-per-step 8 --slots 8 --no-cpu-baseline > gpurun_out/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lva_step_fast -s 900 -c 3 -o gpurun_out/prof_r01_final python bench.py --steps 1 --warmup 0 --reads-per-step 8 --slots 8 --no-cpu-baseline > gpurun_out/b2.log 2>&1
python bench.py --workload C1 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C1', d['value'], d['config']['slots'])"
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01.json; cat gpurun_out/bench_r01.json
