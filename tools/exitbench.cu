// Where does a short CUDA process spend its wall clock outside its own work?  Prints the in-process time at which it
// calls _exit(); the caller compares with the process wall clock.  Usage: exitbench <GiB device> <touch 0|1> <MiB pinned> <free 0|1>
#include <cuda_runtime.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  const double t0 = now();
  const size_t gib = argc > 1 ? atoll(argv[1]) : 0;
  const int touch = argc > 2 ? atoi(argv[2]) : 0;
  const size_t pin = argc > 3 ? atoll(argv[3]) : 0;
  const int do_free = argc > 4 ? atoi(argv[4]) : 0;
  cudaFree(0);
  const double t1 = now();
  void* d = nullptr;
  if (gib) cudaMalloc(&d, gib << 30);
  if (gib && touch) { cudaMemset(d, 0xFF, gib << 30); cudaDeviceSynchronize(); }
  const double t2 = now();
  void* h = nullptr;
  if (pin) cudaMallocHost(&h, pin << 20);
  const double t3 = now();
  if (do_free) { if (d) cudaFree(d); if (h) cudaFreeHost(h); }
  const double t4 = now();
  printf("{\"ctx_s\": %.3f, \"dev_alloc_s\": %.3f, \"pin_s\": %.3f, \"free_s\": %.3f, \"exit_called_at_s\": %.3f}\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0);
  fflush(stdout);
  _exit(0);
}
