#!/bin/bash
# Profiling recipe of this repo (run on the GPU box through gpurun; numbers printed under ncu are never bench values).
#   bash profiles/capture.sh r01b
tag=${1:-r01}
out=gpurun_out
export KB_BENCH_NO_RANDBENCH=1
# 1. launch list of the bench command: per-launch durations (cold cache, serialised) -> kernel SHARES of a step
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_$tag.log 2>&1
# 2. one full capture of the hot kernels in steady state: third timed step (pack, match, resolve) + the EM kernel.
#    Matching launches before it: 3 warm-up steps x 3 kernels + the warm-up EM + 2 timed steps x 3 = 16.
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|match_kernel|resolve_kernel|em_kernel" -s 16 -c 4 \
    -f -o $out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_$tag.log 2>&1
tail -3 $out/ncu_$tag.log
