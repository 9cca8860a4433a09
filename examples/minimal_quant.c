/* Smallest caller of the C ABI (plain C99, no CUDA headers): load an index, push two batches of paired reads from
 * host memory, run the EM, print the five most abundant targets.  This is the call sequence INTEGRATION.md maps onto
 * MasterProcessor::processReads / EMAlgorithm::run; tests/test_cabi_host.py compiles and links it.
 *
 *   gcc -std=c99 -Iinclude examples/minimal_quant.c -Lkallisto_b200 -lkallisto_b200 -Wl,-rpath,$PWD/kallisto_b200 -o minimal_quant
 *   ./minimal_quant index.kidx
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kallisto_b200.h"

#define CHECK(call)                                                      \
  do {                                                                   \
    if ((call) != KB_OK) {                                               \
      fprintf(stderr, "%s\n", kb_last_error());                          \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s index.kidx\n", argv[0]);
    return 2;
  }
  kb_index* ix = NULL;
  CHECK(kb_index_load(argv[1], /*device*/ 0, /*load_positions*/ 0, /*threads*/ 4, &ix));
  kb_index_info info;
  CHECK(kb_index_get_info(ix, &info));

  kb_quant_opts qo;
  memset(&qo, 0, sizeof(qo));
  qo.paired = 1;
  qo.collect_fld = 1;
  kb_quant* q = NULL;
  CHECK(kb_quant_create(ix, &qo, &q));

  /* two mates of one fragment, interleaved like SR->fetchSequences hands them over (offsets delimit the reads) */
  const char bases[] = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT" "TTTTACGTACGTACGTACGTACGTACGTACGTACGTAAAA";
  const uint32_t off[3] = {0, 40, 80};
  for (int batch = 0; batch < 2; ++batch) CHECK(kb_pseudoalign_batch(q, bases, off, 2, 0, NULL));

  double* est = (double*)malloc(sizeof(double) * info.n_targets);
  double* eff = (double*)malloc(sizeof(double) * info.n_targets);
  int32_t rounds = 0;
  CHECK(kb_em_run(q, /*fragment length mean*/ 0.0, /*sd*/ 0.0, est, eff, &rounds, NULL));
  kb_run_stats st;
  CHECK(kb_quant_finalize(q, &st));
  printf("%llu fragments, %llu pseudoaligned, %llu equivalence classes, EM %d rounds\n", (unsigned long long)st.n_processed,
         (unsigned long long)st.n_pseudoaligned, (unsigned long long)st.n_ecs, (int)rounds);
  for (int k = 0; k < 5; ++k) {
    uint32_t best = 0;
    for (uint32_t t = 1; t < info.n_targets; ++t)
      if (est[t] > est[best]) best = t;
    printf("%s\t%g\n", kb_index_target_name(ix, best), est[best]);
    est[best] = -1.0;
  }
  free(est);
  free(eff);
  kb_quant_free(q);
  kb_index_free(ix);
  return 0;
}
