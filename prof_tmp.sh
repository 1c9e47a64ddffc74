mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lva_step_fast -s 900 -c 3 -o gpurun_out/prof_h python bench.py --steps 1 --warmup 0 --reads-per-step 8 --slots 8 --no-cpu-baseline > gpurun_out/b2.log 2>&1
