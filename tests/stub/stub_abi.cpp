// TEST INFRASTRUCTURE: a stand-in for libkallisto_b200.so that implements the C ABI entry points the command line
// uses WITHOUT any device work, so that the host side of the CLI (option handling, parser threads, parallel and gzip
// readers, lock-step batch hand-over, writers) can run on a CPU-only box, also under ThreadSanitizer.  It digests
// every fragment it is given in order (FNV-1a over the bases of all mates) and reports the digest through the
// "estimated counts" of targets 0 and 1, which end up in abundance.tsv.  tests/test_cli_host_pipeline.py builds
// csrc/cli_main.cpp against it.  Never linked into the product.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kallisto_b200.h"

struct kb_index { int dummy; };
struct kb_quant {
  uint64_t h = 1469598103934665603ULL, n = 0, bases = 0;
  int nfiles = 0;
  uint64_t sample = 0;      // kb_bus_begin_sample
};
static std::string g_err;
static void mix(kb_quant* q, const char* p, uint32_t len) {
  for (uint32_t i = 0; i < len; ++i) { q->h ^= (unsigned char)p[i]; q->h *= 1099511628211ULL; }
  q->h ^= 0xFF;
  q->h *= 1099511628211ULL;
  q->bases += len;
}

extern "C" {
const char* kb_last_error(void) { return g_err.c_str(); }
const char* kb_version(void) { return "stub"; }
int kb_index_load(const char*, int, int, int, kb_index** out) { *out = new kb_index(); return KB_OK; }
void kb_index_free(kb_index* ix) { delete ix; }
int kb_index_get_info(const kb_index*, kb_index_info* info) {
  memset(info, 0, sizeof(*info));
  info->k = 31;
  info->n_targets = 3;
  info->n_kmers = 1;
  return KB_OK;
}
const char* kb_index_target_name(const kb_index*, uint32_t i) { static const char* n[3] = {"digest_lo", "digest_hi", "reads"}; return n[i % 3]; }
int kb_index_target_lens(const kb_index*, uint32_t* lens) { lens[0] = lens[1] = lens[2] = 1000; return KB_OK; }
int kb_quant_create(kb_index*, const kb_quant_opts*, kb_quant** out) { *out = new kb_quant(); return KB_OK; }
void kb_quant_free(kb_quant* q) { delete q; }
void* kb_host_alloc(size_t bytes) { return malloc(bytes); }
void kb_host_free(void* p) { free(p); }
int kb_quant_enable_timing(kb_quant*, int) { return KB_OK; }
int kb_quant_get_timings(kb_quant*, kb_kernel_timings* t) { memset(t, 0, sizeof(*t)); return KB_OK; }
int kb_pseudoalign_batch(kb_quant* q, const char* bases, const uint32_t* off, uint32_t n_reads, uint32_t, int32_t*) {
  for (uint32_t i = 0; i < n_reads; ++i) mix(q, bases + off[i], off[i + 1] - off[i]);
  q->n += n_reads;
  return KB_OK;
}
int kb_pseudoalign_batch_pe(kb_quant* q, const char* b1, const uint32_t* o1, const char* b2, const uint32_t* o2, uint32_t n_pairs,
                            uint32_t, int32_t*) {
  for (uint32_t i = 0; i < n_pairs; ++i) {
    mix(q, b1 + o1[i], o1[i + 1] - o1[i]);
    mix(q, b2 + o2[i], o2[i + 1] - o2[i]);
  }
  q->n += n_pairs;
  return KB_OK;
}
int kb_quant_get_flens(kb_quant* q, uint32_t* f) { memset(f, 0, 1000 * sizeof(uint32_t)); f[200] = 10; f[1] = (uint32_t)q->sample; return KB_OK; }
int kb_em_run(kb_quant* q, double, double, double* est, double* eff, int32_t* rounds, double*) {
  est[0] = (double)(q->h & 0xFFFFu);
  est[1] = (double)((q->h >> 16) & 0xFFFFu);
  est[2] = (double)((q->h >> 32) & 0xFFFFu);
  eff[0] = eff[1] = eff[2] = 800.0;
  if (rounds) *rounds = 1;
  return KB_OK;
}
int kb_quant_finalize(kb_quant* q, kb_run_stats* st) {
  memset(st, 0, sizeof(*st));
  st->n_processed = q->n;
  st->n_pseudoaligned = q->n;
  st->n_unique = q->n;
  st->n_ecs = 1;
  st->n_ec_entries = 1;
  return KB_OK;
}
int kb_quant_ec_table(kb_quant*, uint64_t* off, uint32_t* tids, uint32_t* counts, int32_t*) {
  off[0] = 0; off[1] = 1; tids[0] = 0;
  if (counts) counts[0] = 1;
  return KB_OK;
}
int kb_bootstrap_run(kb_quant*, double, double, uint64_t, int32_t, double*, uint32_t*, int32_t*) { return KB_OK; }
int kb_counts_to_tpm(const double* est, const double* eff, uint32_t n, double* tpm) {
  double tot = 0;
  for (uint32_t i = 0; i < n; ++i) { tpm[i] = est[i] / eff[i]; tot += tpm[i]; }
  for (uint32_t i = 0; i < n; ++i) tpm[i] = tot > 0 ? tpm[i] / tot * 1e6 : 0.0;
  return KB_OK;
}
// multi-device entry points: the stub has one "device"; --devices is covered by the GPU tests
struct kb_comm { int unused; };
int kb_comm_create_all(const int*, int n, kb_comm** out) { for (int i = 0; i < n; ++i) out[i] = new kb_comm(); return KB_OK; }
void kb_comm_free(kb_comm* c) { delete c; }
int kb_quant_merge_nccl(kb_quant*, kb_comm*, uint64_t, uint64_t*) { return KB_OK; }
int kb_quant_set_frag_base(kb_quant*, uint64_t) { return KB_OK; }
int kb_quant_merge_local(kb_quant*, kb_quant* const*, int32_t, uint64_t*) { return KB_OK; }
int kb_tcc_run(kb_index*, uint32_t, const uint64_t*, const uint32_t*, uint32_t, const uint64_t*, const uint32_t*, const uint32_t*, const double*, int32_t,
               double*, int32_t*) { return KB_OK; }
int kb_eff_lens(const kb_index*, const uint32_t*, double, double, double*, double*, double*) { return KB_OK; }
int kb_bus_create(kb_index*, const kb_bus_opts* o, kb_quant** out) { *out = new kb_quant(); (*out)->nfiles = o->nfiles; return KB_OK; }
int kb_bus_batch(kb_quant* q, const char* const* bases, const uint32_t* const* offs, uint32_t n_sets, kb_bus_record* rec, uint32_t* n_rec) {
  // one record per read set: barcode = running digest, so that output.bus depends on content AND order
  for (uint32_t i = 0; i < n_sets; ++i) {
    for (int f = 0; f < q->nfiles; ++f) mix(q, bases[f] + offs[f][i], offs[f][i + 1] - offs[f][i]);
    memset(&rec[i], 0, sizeof(rec[i]));
    rec[i].barcode = q->h;
    rec[i].umi = q->sample;
    rec[i].count = 1;
  }
  q->n += n_sets;
  *n_rec = n_sets;
  return KB_OK;
}
int kb_bus_begin_sample(kb_quant* q, uint64_t barcode) { q->sample = barcode; return KB_OK; }
int kb_bus_lengths(kb_quant*, uint32_t* bc, uint32_t* umi) { memset(bc, 0, 33 * 4); memset(umi, 0, 33 * 4); bc[16] = 1; umi[10] = 1; return KB_OK; }
}
