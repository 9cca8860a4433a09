// kallisto index (format v13) -> flat, GPU-friendly arrays.
//
// The on-disk layout is the drop-in contract with the reference
// (written by KmerIndex::write, src/KmerIndex.cpp:570-661,1170-1224; read by
// KmerIndex::load, src/KmerIndex.cpp:1330-1559).  Nothing of Bifrost's in-memory
// machinery (minimizer index, BBHash MPHF, tiny vectors) is reproduced: we only
// need the unitig sequences (GRAPH section, ext/bifrost/src/IO.tcc:1635-1738) and
// the per-unitig mosaic equivalence classes (src/Node.hpp:51-72,
// src/BlockArray.hpp:418-471, src/SparseVector.tcc:396-422).  The MPHF blob and
// Bifrost's INDEX/meta section are length-prefixed and skipped.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace kb {

struct FlatIndex {
  int k = 0;
  int g = 0;
  bool graphless = false;   // index.saved of `kallisto bus`: targets only, no k-mers (usable by quant-tcc alone)
  uint32_t n_long = 0, n_short = 0, n_abund = 0;   // unitig kinds, in global-id order
  uint32_t n_unitigs() const { return n_long + n_short + n_abund; }

  // Long unitigs: Bifrost 2-bit packing (base i -> byte i>>2, bits (i&3)*2, A0 C1 G2 T3,
  // ext/bifrost/src/CompressedSequence.cpp:311-314), each unitig byte-aligned.
  std::vector<uint8_t>  useq;
  std::vector<uint64_t> useq_byteoff;   // n_long + 1
  std::vector<uint32_t> ulen;           // per unitig, bases (short/abundant: k)
  std::vector<uint64_t> skmer;          // n_short + n_abund canonical k-mers, right-aligned 2k bits
  uint64_t n_kmers = 0;                 // sum(len - k + 1): the "[index] number of k-mers" line

  // Mosaic EC blocks per unitig, sorted by lb (BlockArray).
  std::vector<uint64_t> blk_off;        // n_unitigs + 1
  std::vector<uint32_t> blk_lb, blk_ub, blk_ec;
  // Per (block, member-of-its-EC) strand byte, aligned with the EC's sorted tids:
  // 1 = unitig-forward is transcript-sense, 0 = antisense, 2 = both
  // (SparseVector::operator[], src/SparseVector.tcc:363-389).
  std::vector<uint64_t> blk_strand_off; // n_blocks + 1
  std::vector<uint8_t>  strand;

  // Per (block, member) sorted position lists (only when load_positions): values are
  // pos | antisense<<31 exactly as stored (src/KmerIndex.cpp:1023,1066).  CSR over the
  // same (block, member) slots as `strand`.
  bool has_positions = false;
  std::vector<uint64_t> pos_off;        // strand.size() + 1
  std::vector<uint32_t> pos_val;

  // Per (block, member), when load_positions: the read-independent part of KmerIndex::findPosition
  // (src/KmerIndex.cpp:2188-2292), 4 words each, same slots as `strand`:
  //   [0] smallest stored position of the transcript in this block (pos | antisense << 31)
  //   [1] case I   padding   (lb of the first block of the run of blocks holding tr that ends here, if pos == 0)
  //   [2] case III left_one  (ub of the nearest earlier block that does not hold tr, else 0)
  //   [3] case II/IV right_one - left_one over the blocks not holding tr after the first that does
  std::vector<uint32_t> fp_info;

  // Transcript sets de-duplicated by content (SparseVector::operator== compares only the
  // Roaring of transcript ids, src/SparseVector.tcc:391-394).
  std::vector<uint64_t> ec_off;         // n_ec + 1
  std::vector<uint32_t> ec_tid;
  uint32_t n_ec() const { return (uint32_t)(ec_off.size() - 1); }

  std::vector<uint32_t>    target_len;
  std::vector<std::string> target_name;
  std::vector<uint32_t>    onlist;      // sorted transcript ids on the on-list
  uint64_t dlist_n = 0;
  std::vector<uint64_t> dlist;   // D-list: canonical k-mers (right-aligned 2k bits); [0] is the dummy, the only one that is in the graph
  uint32_t num_targets() const { return (uint32_t)target_len.size(); }
};

// Throws std::runtime_error with a message on any format problem.
void load_index_v13(const std::string& path, FlatIndex& out, bool load_positions, int threads = 1);

// Decoders shared with the tests.
void decode_roaring_native(const uint8_t* p, size_t n, std::vector<uint32_t>& out);
void decode_roaring_portable(const uint8_t* p, size_t n, std::vector<uint32_t>& out);

inline uint64_t kmer_revcomp(uint64_t x, int k) {
  // 2-bit complement = bitwise not; reverse the 2-bit groups.
  x = ~x;
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = __builtin_bswap64(x);
  return x >> (64 - 2 * k);
}

}  // namespace kb
