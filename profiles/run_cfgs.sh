#!/bin/bash
# scratch: bench lines of the other configs (not part of the product)
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --slots 128 > gpurun_out/c2_s128.log 2>&1
for wl in "$@"; do python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/cfg_$wl.log 2>&1; done
tail -c 600 gpurun_out/c2_s128.log | cut -c1-120
for wl in "$@"; do echo == $wl; tail -c 1500 gpurun_out/cfg_$wl.log | cut -c1-400; done
