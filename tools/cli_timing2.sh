#!/bin/bash
for t in 64 64 2; do
  echo "== -t $t"
  KB_EM_TRACE=1 KB_CLI_TIMING=1 kallisto_b200/kallisto_b200 quant -i bench_data/g62000.kidx -o /dev/shm/o1 --plaintext -t $t /dev/shm/c_1.fq /dev/shm/c_2.fq 2>&1 | grep -E "timing|em-trace"
done
