python - <<'PY'
import numpy as np, sys, os
sys.path.insert(0, '.')
import benchdata
rng = np.random.default_rng(1)
reads = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (1000000, 100), dtype=np.uint8)]
for i in range(8):
    benchdata.write_fastq_fast('/dev/shm/fx.fq', reads, 1, append=i > 0)
print(os.path.getsize('/dev/shm/fx.fq') / 1e6, "MB")
PY
for w in 4000000 16000000 64000000; do for t in 1 4 8 16 32; do echo -n "window/thread $w: "; KB_FASTX_DEBUG=1 KB_FASTX_WINDOW=$w tools/fxbench /dev/shm/fx.fq $t 2>&1 | tr '\n' ' '; echo; done; done
python -m pytest tests/test_gpu_cli.py -x -q -k "parallel or unequal" 2>&1 | tail -5
