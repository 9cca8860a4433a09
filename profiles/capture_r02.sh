#!/bin/bash
# Profiling recipe, round 2 (run on the GPU box through gpurun; numbers printed under ncu are never bench values).
#   bash profiles/capture_r02.sh r02
tag=${1:-r02}
out=gpurun_out
export KB_BENCH_NO_RANDBENCH=1 KB_BENCH_NO_CLI=1
# 1. launch list of the bench command: per-launch durations (cold cache, serialised) -> kernel SHARES of a step
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_$tag.log 2>&1
# 2. one full capture of the hot kernels in steady state.  bench.py --steps 3 --warmup 3 launches, of the kernels named
#    below: 3 warm-up steps x (pack, match, resolve) + the warm-up EM = 10, the untimed job 3 x 3 + EM = 10, then the
#    first timed job; its third step is launches 26-28 and its EM launch 29.
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|match_kernel|resolve_kernel|em_kernel|em_single_kernel" -s 26 -c 4 \
    -f -o $out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_$tag.log 2>&1
tail -3 $out/ncu_$tag.log
