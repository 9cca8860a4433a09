#!/bin/bash
# Experiment: match_kernel residency (blocks of 256 lanes per SM, -DKB_MATCH_MIN_BLOCKS) x L2 fetch granularity.
out=gpurun_out/occsweep; mkdir -p $out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/base.json 2> $out/base.err
for mb in 3 4 5 6; do
  lib=variants/libkb_mb$mb.so; [ $mb = 5 ] && lib=kallisto_b200/libkallisto_b200.so
  for fetch in "" 32; do
    tag=mb${mb}_f${fetch:-def}
    KB_LIB_PATH=$lib KB_L2_FETCH=$fetch python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/$tag.json 2> $out/$tag.err
    python - "$tag" "$out" <<'PY'
import json, sys
tag, out = sys.argv[1:3]
try:
    d = json.loads(open(f"{out}/{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(tag, "match ms %.3f resolve %.3f frac %.3f value %.1fM e2e %.1fM align_ms %.1f" % (r["ms_per_launch"], r["resolve_ms_per_launch"], r["frac"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["config"]["align_ms"]), flush=True)
except Exception as e:
    print(tag, "failed", repr(e), open(f"{out}/{tag}.err").read()[-400:])
PY
  done
done
