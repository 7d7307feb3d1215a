// C++ host test of include/apgk_adapters.hpp: the reference-named classes over the C ABI.
// Expected values are the hand-computed vectors of tests/golden (ACGT K=2, GATTACA K=4) plus
// internal consistency checks; it prints "ADAPTERS OK" and exits 0 on success.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "apgk_adapters.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  using namespace apgk_ref;
  {  // ACGT, K=2  ->  AC:2  CG:1 ; spectrum [0,1,1]
    vecbasevector r; r.push_back("ACGT");
    std::vector<kmer_count<1>> R;
    SortKmers<2>(r, R);
    REQUIRE(R.size() == 2 && R[0].kmer[0] == 1 && R[0].count == 2 && R[1].kmer[0] == 6 && R[1].count == 1);
    KmerSpectrum s(2); s.FromReads(r);
    REQUIRE(s.size() == 3 && s[1] == 1 && s[2] == 1 && s.NumInstances() == 3 && s.NumDistinct() == 2);
    // record form: AC at +1 (forward) and at -3 (GT read backwards), the palindrome CG at +2
    std::vector<kmer_record<1>> Q;
    SortKmers<2>(r, Q);
    REQUIRE(Q.size() == 3);
    REQUIRE(Q[0].kmer[0] == 1 && Q[0].read_id == 0 && Q[0].pos == 1);
    REQUIRE(Q[1].kmer[0] == 1 && Q[1].read_id == 0 && Q[1].pos == -3);
    REQUIRE(Q[2].kmer[0] == 6 && Q[2].read_id == 0 && Q[2].pos == 2);
  }
  {  // GATTACA + its reverse complement, K=4 -> four k-mers, each seen twice
    vecbasevector r; r.push_back("GATTACA"); r.push_back("TGTAATC");
    KmerParcelsBuilder b(4, r, 8);
    b.Build();
    REQUIRE(b.NumKmersDistinct() == 4 && b.NumKmerInstances() == 8);
    uint64_t k[4]; uint32_t c[4];
    b.Records(0, 4, k, c);
    const uint64_t want[4] = {13, 60, 176, 196};
    for (int i = 0; i < 4; i++) REQUIRE(k[i] == want[i] && c[i] == 2);
    KmerSpectrum s = b.Spectrum();
    REQUIRE(s.size() == 3 && s[2] == 4);
    // batches: every k-mer once in each read, on opposite strands, at mirrored positions (p+1) + (4-p) = 5
    std::vector<uint64_t> off; std::vector<uint32_t> ids; std::vector<int32_t> pos;
    b.Batches(0, 4, off, ids, pos);
    REQUIRE(off.size() == 5 && ids.size() == 8 && pos.size() == 8);
    for (int i = 0; i < 4; i++) {
      REQUIRE(off[i + 1] - off[i] == 2);
      const uint64_t o = off[i] - off[0];
      REQUIRE(ids[o] == 0 && ids[o + 1] == 1);
      REQUIRE((pos[o] > 0) != (pos[o + 1] > 0));
      REQUIRE(std::abs(pos[o]) + std::abs(pos[o + 1]) == 5);
    }
  }
  {  // frequency table: random reads, K=24; every window's frequency >= 1, and equals Freq() of that window
    std::mt19937_64 g(7);
    vecbasevector r;
    std::string genome;
    for (int i = 0; i < 5000; i++) genome += "ACGT"[g() & 3];
    for (int i = 0; i < 400; i++) r.push_back(genome.substr(g() % 4900, 100));
    KmerFreqTable t(24, r);
    std::vector<uint32_t> f = t.ReadFreqs();
    REQUIRE(f.size() == 400 * 100);
    uint64_t valid = 0;
    for (size_t i = 0; i < f.size(); i++) {
      const bool inside = (i % 100) <= 100 - 24;
      REQUIRE((f[i] != 0xFFFFFFFFu) == inside);
      if (inside) { REQUIRE(f[i] >= 1); valid++; }
    }
    REQUIRE(valid == 400 * 77);
    // query one k-mer by value: first window of read 0
    std::string w = genome.substr(0, 24);
    (void)w;
    std::vector<kmer_count<1>> R;
    SortKmers<24>(r, R);
    uint64_t tot = 0;
    for (auto& x : R) tot += x.count;
    REQUIRE(tot == valid);
    for (size_t i = 1; i < R.size(); i++) REQUIRE(R[i - 1].kmer[0] < R[i].kmer[0]);
    REQUIRE(t.Freq(R[R.size() / 2].kmer) == R[R.size() / 2].count);
  }
  {  // K=40: two-word k-mers through the template
    vecbasevector r;
    std::mt19937_64 g(9);
    std::string s;
    for (int i = 0; i < 300; i++) s += "ACGT"[g() & 3];
    r.push_back(s); r.push_back(s);
    std::vector<kmer_count<2>> R;
    SortKmers<40>(r, R);
    REQUIRE(R.size() == 261);
    for (auto& x : R) REQUIRE(x.count == 2);
  }
  std::printf("ADAPTERS OK\n");
  return 0;
}
