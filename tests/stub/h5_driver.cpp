// TEST INFRASTRUCTURE: drives csrc/h5_writer.hpp directly (tests/test_h5_host.py) with sizes the stub command line does not
// reach: n targets with names of growing length, B bootstrap vectors (more symbol table nodes than one default B-tree node holds).
#include <cstdlib>
#include <string>
#include <vector>

#include "h5_writer.hpp"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int n = atoi(argv[2]), B = atoi(argv[3]);
  kb::H5Writer w;
  std::vector<double> est(n);
  std::vector<int32_t> len(n);
  std::vector<std::string> ids(n);
  for (int i = 0; i < n; ++i) {
    est[i] = i * 0.5 + 1.0 / (i + 1);
    len[i] = 200 + 7 * i;
    ids[i] = "ENST" + std::to_string(i) + std::string((size_t)(i % 40), 'x') + "|gene";
  }
  w.add_f64(0, "est_counts", est.data(), est.size());
  const int aux = w.group("aux");
  w.add_str(aux, "ids", ids);
  w.add_i32(aux, "lengths", len.data(), len.size());
  const int32_t nb = B, np = 123456, iv = 13;
  w.add_i32(aux, "num_bootstrap", &nb, 1);
  w.add_i32(aux, "num_processed", &np, 1);
  w.add_i32(aux, "index_version", &iv, 1);
  std::vector<double> eff(n);
  for (int i = 0; i < n; ++i) eff[i] = len[i] - 150.25;
  w.add_f64(aux, "eff_lengths", eff.data(), eff.size());
  w.add_str(aux, "kallisto_version", {"0.51.1"});
  w.add_str(aux, "call", {"h5_driver " + std::string(argv[2]) + " " + argv[3]});
  w.add_str(aux, "start_time", {"Thu Jan  1 00:00:00 1970"});
  if (B > 0) {
    const int bs = w.group("bootstrap");
    for (int b = 0; b < B; ++b) {
      std::vector<double> v(n);
      for (int i = 0; i < n; ++i) v[i] = est[i] * (b + 1);
      w.add_f64(bs, "bs" + std::to_string(b), v.data(), v.size());
    }
  }
  return w.write(argv[1]) ? 0 : 1;
}
