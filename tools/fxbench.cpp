// Ingest microbenchmark: reads/s of the command-line front end's FASTQ reader (csrc/fastx.hpp) filling 1 M-read batches.
//   g++ -O2 -std=c++17 -Ikallisto_b200/csrc -o tools/fxbench tools/fxbench.cpp -lz -lpthread ; tools/fxbench FILE THREADS
#include "fastx.hpp"
#include <chrono>
#include <cstdio>
int main(int argc, char** argv) {
  const int threads = atoi(argv[2]);
  kb::ReadBatch b;
  const size_t max_reads = 1 << 20, max_bases = max_reads * 160 + kb::FastxFile::kMaxRead;
  std::vector<char> bases(max_bases + 64);
  std::vector<uint32_t> off(max_reads + 1);
  b.bases = bases.data(); b.off = off.data(); b.cap_bases = max_bases; b.cap_reads = max_reads;
  auto t0 = std::chrono::steady_clock::now();
  kb::FastxReader f(argv[1], threads);
  size_t n = 0, nb = 0;
  for (;;) { b.clear(); if (!f.fill(b, max_reads)) break; n += b.n; nb += b.n_bases(); }
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("threads %d: %zu reads %zu bases %.3f s  %.1f M reads/s\n", threads, n, nb, dt, n / dt / 1e6);
}
