#!/bin/bash
# Experiment: which load instruction should the k-mer table probe use?  Variants of libkallisto_b200.so are
# built with -DKB_PROBE_LD=n (kernels_align.cu: ld256_probe) into variants/; each is timed with bench.py
# (CUDA events, no profiler) and then profiled for L2/DRAM traffic of match_kernel with ncu.
out=gpurun_out/ldsweep; mkdir -p $out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/base.json 2> $out/base.err   # builds the workload cache
for v in 0 1 2 3 4 5 6; do
  lib=variants/libkb_ld$v.so; [ $v = 0 ] && lib=kallisto_b200/libkallisto_b200.so
  for fetch in "" 32 128; do
    [ -n "$fetch" ] && [ $v != 0 ] && [ $v != 1 ] && continue
    tag=ld${v}_f${fetch:-def}
    KB_LIB_PATH=$lib KB_L2_FETCH=$fetch python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $out/$tag.json 2> $out/$tag.err
    KB_LIB_PATH=$lib KB_L2_FETCH=$fetch ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:match_kernel -s 3 -c 1 --csv --log-file $out/$tag.ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
    python - "$tag" "$out" <<'PY'
import json, sys, csv
tag, out = sys.argv[1:3]
d = json.loads(open(f"{out}/{tag}.json").read().strip().splitlines()[-1])
m = {}
try:
    for r in csv.DictReader(l for l in open(f"{out}/{tag}.ncu.csv") if l.startswith('"')):
        m[r["Metric Name"]] = r["Metric Value"] + " " + r["Metric Unit"]
except Exception as e:
    m = {"err": repr(e)}
print(tag, "match ms %.3f" % d["roofline"]["ms_per_launch"], "value %.1fM" % (d["value"] / 1e6), m, flush=True)
PY
  done
done
