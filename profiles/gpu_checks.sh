python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size or survivor" 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', d['value'], d['config']['slots'], d['e2e']['value'], d['roofline']['frac'])"
python bench.py --workload C3 --steps 1 --warmup 0 --no-cpu-baseline --reads-per-step 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3', d['value'], d['config']['slots'], d['e2e']['value'], d['roofline']['frac'], d['top1_correct_frac'])"
