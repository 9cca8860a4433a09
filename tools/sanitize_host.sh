#!/bin/bash
# Host-side code under sanitizers (run in the build container; no GPU needed).  What it covers:
#   1. csrc/index_v13.cpp       ASan+UBSan on damaged index files (byte flips, truncation, absurd length fields)
#   2. csrc/fast_inflate.hpp    ASan+UBSan on valid streams of every kind and on damaged ones; output checked against zlib
#   3. csrc/fastx.hpp           ASan+UBSan on garbage FASTA/FASTQ (plain and gzip, sequential and parallel reader)
#   4. csrc/cli_main.cpp        TSan and ASan+LSan of the whole command line against tests/stub/stub_abi.cpp
# Exit code 0 = nothing reported.
set -e
cd "$(dirname "$0")/.."
W=$(mktemp -d)
trap 'rm -rf "$W"' EXIT
SAN="-O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++17"
fail=0

echo "== 1. index parser"
cat > $W/ix.cpp <<'CPP'
#include "index_v13.hpp"
#include <cstdio>
int main(int argc, char** argv) {
  for (int i = 1; i < argc; ++i) {
    try { kb::FlatIndex f; kb::load_index_v13(argv[i], f, i % 2 == 0, 2); } catch (const std::exception&) {}
  }
  puts("done");
}
CPP
g++ $SAN -Ikallisto_b200/csrc -o $W/ix $W/ix.cpp kallisto_b200/csrc/index_v13.cpp -lpthread
python - "$W" <<'PY'
import sys, numpy as np
w = sys.argv[1]; rng = np.random.default_rng(9)
for name in ("config1", "manyecs"):
    data = open('tests/golden/%s/transcripts.kidx' % name, 'rb').read()
    for i in range(60):
        b = bytearray(data)
        if i % 3 == 0:
            for _ in range(int(rng.integers(1, 4))): b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif i % 3 == 1:
            b = b[:int(rng.integers(0, len(b)))]
        else:
            p = int(rng.integers(0, len(b) - 8)); b[p:p + 8] = int(rng.integers(0, 2 ** 63)).to_bytes(8, 'little')
        open('%s/%s_%03d.kidx' % (w, name, i), 'wb').write(bytes(b))
PY
$W/ix $W/*.kidx > /dev/null 2> $W/ix.err || { echo "index parser: sanitizer report"; head -20 $W/ix.err; fail=1; }

echo "== 2. gzip decoder"
cat > $W/gz.cpp <<'CPP'
#include "fast_inflate.hpp"
#include <cstdio>
int main(int, char** argv) {
  try {
    kb::FastGz g(argv[1]);
    const char* p; size_t n, tot = 0; uLong c = crc32(0, 0, 0);
    while (g.next_chunk(p, n)) { tot += n; c = crc32(c, (const Bytef*)p, (uInt)n); }
    printf("OK %zu %08lx\n", tot, c);
  } catch (const std::exception& e) { printf("ERR\n"); return 2; }
}
CPP
g++ $SAN -Ikallisto_b200/csrc -o $W/gz $W/gz.cpp -lz
python - "$W" <<'PY' || fail=1
import sys, os, random, subprocess, zlib
sys.path.insert(0, '.')
from tests.test_gz_host import corpus
w = sys.argv[1]; rnd = random.Random(1); bad = 0
cases = corpus()
for name, (blob, data) in cases.items():
    p = os.path.join(w, 'c.gz'); open(p, 'wb').write(blob)
    r = subprocess.run([os.path.join(w, 'gz'), p], capture_output=True, text=True)
    if r.returncode != 0 or r.stdout.split()[1:] != [str(len(data)), '%08x' % zlib.crc32(data)]:
        bad += 1; print('gzip decoder differs from zlib / sanitizer report:', name, r.stderr[-300:])
blobs = [cases[k][0] for k in ('fastq_l1', 'fastq_l9', 'fastq_fixed', 'rand_l6', 'multi', 'len258_l9')]
for i in range(1500):
    b = bytearray(rnd.choice(blobs)); m = i % 4
    if m == 0:
        for _ in range(rnd.randrange(1, 6)): b[rnd.randrange(len(b))] = rnd.randrange(256)
    elif m == 1: b = b[:rnd.randrange(len(b))]
    elif m == 2:
        q = rnd.randrange(len(b)); b[q:q] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 9)))
    else: b[rnd.randrange(10, len(b))] ^= 1 << rnd.randrange(8)
    p = os.path.join(w, 'd.gz'); open(p, 'wb').write(bytes(b))
    r = subprocess.run([os.path.join(w, 'gz'), p], capture_output=True, text=True)
    if r.returncode not in (0, 2): bad += 1; print('gzip decoder: sanitizer report on damaged input', i, r.stderr[-300:])
sys.exit(1 if bad else 0)
PY

echo "== 3. FASTA/FASTQ readers"
g++ $SAN -Ikallisto_b200/csrc -o $W/fx tools/fxbench.cpp -lz -lpthread
python - "$W" <<'PY'
import sys, random, gzip
w = sys.argv[1]; rnd = random.Random(11)
good = b"".join(b"@r%d\nACGTNACGT%s\n+\nIIIIIIIII%s\n" % (i, b"A" * (i % 7), b"I" * (i % 7)) for i in range(3000))
cases = {"rand": bytes(rnd.randrange(256) for _ in range(200000)), "ats": b"@" * 50000, "plus": b"@x\n" + b"+\n" * 30000, "nl": b"\n" * 40000,
         "cut1": good[:len(good) // 2 + 3], "cut2": good[:-5], "nul": good.replace(b"A", b"\0", 500), "fa": b"".join(b">t%d\nACGT\nAC\n" % i for i in range(5000))}
for i in range(12):
    b = bytearray(good)
    for _ in range(200): b[rnd.randrange(len(b))] = rnd.randrange(256)
    cases["flip%d" % i] = bytes(b)
for k, v in cases.items():
    open('%s/g_%s.fq' % (w, k), 'wb').write(v); open('%s/g_%s.fq.gz' % (w, k), 'wb').write(gzip.compress(v, 1))
PY
for f in $W/g_*; do for t in 1 4; do KB_FASTX_WINDOW=300 $W/fx $f $t > $W/fx.out 2>&1 || { echo "reader: sanitizer report on $f ($t threads)"; tail -5 $W/fx.out; fail=1; }; done; done

echo "== 4. command line against the stub (TSan, then ASan+LSan)"
mkdir -p $W/data
python - "$W" <<'PY'
import sys, gzip
for m in (1, 2): open('%s/data/r%d.fq' % (sys.argv[1], m), 'wb').write(gzip.open('tests/golden/synth_small/reads_%d.fastq.gz' % m).read())
PY
for san in thread address,undefined; do
  d=$W/cli_${san%%,*}; mkdir -p $d
  g++ -O1 -g -fsanitize=$san -std=c++17 -fPIC -shared -Iinclude -o $d/libkallisto_b200.so tests/stub/stub_abi.cpp
  g++ -O1 -g -fsanitize=$san -std=c++17 -Iinclude -Ikallisto_b200/csrc -o $d/cli kallisto_b200/csrc/cli_main.cpp -L$d -lkallisto_b200 -Wl,-rpath,$d -lz -lpthread
  KB_CLI_CLEANUP=1 KB_FASTX_WINDOW=30000 KB_CLI_BATCH_READS=700,1100 $d/cli quant -i tests/golden/synth_small/transcripts.kidx -o $d/o1 --plaintext -t 8 $W/data/r1.fq $W/data/r2.fq > /dev/null 2> $d/e1 || true
  KB_CLI_CLEANUP=1 KB_CLI_BATCH_READS=512,4096 $d/cli quant -i tests/golden/synth_small/transcripts.kidx -o $d/o2 --plaintext -t 8 tests/golden/synth_small/reads_1.fastq.gz tests/golden/synth_small/reads_2.fastq.gz > /dev/null 2> $d/e2 || true
  KB_CLI_BATCH_READS=300,470 $d/cli bus -i tests/golden/config1/transcripts.kidx -o $d/o3 -x 10xv2 -t 4 tests/golden/bus10x/sc_reads_1.fastq.gz tests/golden/bus10x/sc_reads_2.fastq.gz > /dev/null 2> $d/e3 || true
  python -c "import gzip,sys; a,b=[gzip.open('tests/golden/bus10x/sc_reads_%d.fastq.gz'%m,'rb').read().split(b'\\n') for m in (1,2)]; open(sys.argv[1],'wb').write(b''.join(b'\\n'.join(x[4*i:4*i+4])+b'\\n' for i in range(len(a)//4) for x in (a,b)))" $d/il.fq
  KB_CLI_CLEANUP=1 KB_CLI_BATCH_READS=333 $d/cli bus -i tests/golden/config1/transcripts.kidx -o $d/o3b -x 10xv2 -t 4 --inleaved $d/il.fq > /dev/null 2>> $d/e3 || true
  cmp -s $d/o3/output.bus $d/o3b/output.bus || { echo "interleaved input gave other records under -fsanitize=$san"; fail=1; }
  # sample-per-file bus run (file-set switching, flens.txt / index.saved / matrix.cells writers) and the HDF5 emitter
  KB_CLI_CLEANUP=1 KB_CLI_BATCH_READS=700,1100 $d/cli bus -i tests/golden/synth_small/transcripts.kidx -o $d/o4 -x bulk --paired -t 4 $W/data/r1.fq $W/data/r2.fq tests/golden/synth_small/reads_1.fastq.gz tests/golden/synth_small/reads_2.fastq.gz > /dev/null 2> $d/e4 || true
  KB_CLI_CLEANUP=1 $d/cli quant -i tests/golden/synth_small/transcripts.kidx -o $d/o5 -b 40 -t 4 $W/data/r1.fq $W/data/r2.fq > /dev/null 2> $d/e5 || true
  # ... and the reader behind h5dump, on the good file and on damaged copies of it
  $d/cli h5dump -o $d/o6 $d/o5/abundance.h5 > /dev/null 2> $d/e6 || true
  cmp -s $d/o6/abundance.tsv $d/o5/abundance.tsv || { echo "h5dump does not give abundance.tsv back under -fsanitize=$san"; fail=1; }
  python - "$d" <<'PY'
import random, sys
d = sys.argv[1]; rnd = random.Random(3); good = open(d + '/o5/abundance.h5', 'rb').read()
for i in range(40):
    b = bytearray(good[:rnd.randrange(1, len(good))] if i % 4 == 0 else good)
    for _ in range(rnd.choice((1, 2, 16))): b[rnd.randrange(len(b))] = rnd.randrange(256)
    open('%s/bad_%d.h5' % (d, i), 'wb').write(bytes(b))
PY
  for f in $d/bad_*.h5; do $d/cli h5dump -o $d/o7 $f > /dev/null 2>> $d/e6 || true; done
  if grep -q -E "Sanitizer|runtime error" $d/e6; then echo "h5dump under -fsanitize=$san: report"; grep -h -E "Sanitizer|runtime error" $d/e6 | head -5; fail=1; fi
  # bootstrap text files are written by several threads
  KB_CLI_CLEANUP=1 $d/cli quant -i tests/golden/synth_small/transcripts.kidx -o $d/o8 --plaintext -b 9 -t 4 $W/data/r1.fq $W/data/r2.fq > /dev/null 2>> $d/e5 || true
  [ -s $d/o8/bs_abundance_8.tsv ] || { echo "bootstrap text files missing under -fsanitize=$san"; fail=1; }
  [ -s $d/o4/index.saved ] && [ -s $d/o5/abundance.h5 ] || { echo "bus -x bulk / abundance.h5 outputs missing under -fsanitize=$san"; fail=1; }
  if grep -q -E "Sanitizer|runtime error" $d/e1 $d/e2 $d/e3 $d/e4 $d/e5; then echo "command line under -fsanitize=$san: report"; grep -h -E "Sanitizer|runtime error" $d/e1 $d/e2 $d/e3 $d/e4 $d/e5 | head; fail=1; fi
  cmp -s $d/o1/abundance.tsv $d/o2/abundance.tsv || { echo "plain and gzip input gave different digests"; fail=1; }
done
[ $fail = 0 ] && echo "clean"
exit $fail
