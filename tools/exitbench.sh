#!/bin/bash
# process wall clock vs in-process time for a few shapes of a short CUDA process
nvcc -O2 -o /tmp/exitbench tools/exitbench.cu 2>/dev/null || exit 1
for args in "0 0 0 0" "34 0 0 0" "34 1 0 0" "34 1 1300 0" "34 1 1300 1" "0 0 1300 0" "2 1 256 0"; do
  for rep in 1 2; do
    S=$(date +%s%N); /tmp/exitbench $args; E=$(date +%s%N)
    echo "args=[$args] process_wall_s=$(python3 -c "print(($E-$S)/1e9)")"
  done
done
