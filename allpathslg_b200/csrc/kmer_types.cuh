// kmer_types.cuh -- k-mer key representation and bit helpers shared by every kernel.
//
// A k-mer of K bases is the 2K-bit unsigned integer with its FIRST base in the
// most significant position (A=0 C=1 G=2 T=3), stored right-aligned in
// W = ceil(2K/64) 64-bit words, w[0] most significant.  This is the layout
// SURVEY.md section 8 defines for the reference's kmer records (unverified
// against the reference source, which was not available: "parity unpinned").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define APGK_HD __host__ __device__ __forceinline__

namespace apgk {

template <int W>
struct Key {
  uint64_t w[W];
};

template <int W>
APGK_HD bool key_less(const Key<W>& a, const Key<W>& b) {
#pragma unroll
  for (int i = 0; i < W; i++) {
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i];
  }
  return false;
}
template <int W>
APGK_HD bool key_eq(const Key<W>& a, const Key<W>& b) {
  bool e = true;
#pragma unroll
  for (int i = 0; i < W; i++) e = e && (a.w[i] == b.w[i]);
  return e;
}

APGK_HD uint32_t lowmask32(int len) { return len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u); }
APGK_HD uint64_t lowmask64(int len) { return len >= 64 ? ~0ull : ((1ull << len) - 1ull); }

// bits [pos, pos+len) of the 64W-bit value (pos counted from the LSB of the
// whole value), 0 <= len <= 32.  Bits beyond the top word read as 0.
template <int W>
APGK_HD uint32_t key_bits(const Key<W>& k, int pos, int len) {
  const int q = pos >> 6;  // 64-bit chunk index counted from the least significant word
  const int sh = pos & 63;
  uint64_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < W; i++) {
    if (W - 1 - i == q) lo = k.w[i];
    if (W - 1 - i == q + 1) hi = k.w[i];
  }
  uint64_t v = lo >> sh;
  if (sh) v |= hi << (64 - sh);
  return (uint32_t)v & lowmask32(len);
}

// splitmix64 finaliser: owner hash for the multi-GPU shuffle and the synthetic generator.
APGK_HD uint64_t sm64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <int W>
APGK_HD uint64_t key_hash(const Key<W>& k) {
  uint64_t h = 0;
#pragma unroll
  for (int i = 0; i < W; i++) h = sm64(h ^ k.w[i]);
  return h;
}
// owner rank of a canonical k-mer among n_ranks (multiply-shift range reduction of the hash)
template <int W>
APGK_HD uint32_t key_owner(const Key<W>& k, uint32_t n_ranks) {
  const uint32_t h = (uint32_t)(key_hash(k) >> 32);
#ifdef __CUDA_ARCH__
  // Explicit mul.hi: with the 64-bit product form, ptxas 12.9 (sm_100a) folded the shift into a
  // wrong shared-memory address in a loop remainder of k_scatter_reads (illegal address at run time).
  return __umulhi(h, n_ranks);
#else
  return (uint32_t)(((uint64_t)h * (uint64_t)n_ranks) >> 32);
#endif
}

APGK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31));
#endif
}
APGK_HD uint64_t brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return __brevll(x);
#else
  x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  return (x >> 32) | (x << 32);
#endif
}
APGK_HD uint64_t swap_pairs(uint64_t x) {
  return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// How a 2K-bit key is cut into partition digits.  TB = max(2K, P) "virtual"
// bits: keys shorter than the prefix are left-padded (pad = TB - 2K, W == 1 only).
struct KeyGeom {
  int K;        // bases per k-mer
  int W;        // 64-bit words per key
  int TB;       // virtual key bits
  int pad;      // TB - 2K
  int D0, D1;   // level-0 / level-1 digit bits
  int REM;      // TB - D0 - D1 : bits left for the in-shared-memory sort
  int topbits;  // significant bits of w[0]
};

template <int W>
APGK_HD uint32_t digit_of(const Key<W>& k, int pos, int len, int pad) {
  if (W == 1) {
    uint64_t v = k.w[0] << pad;
    return (uint32_t)(v >> pos) & lowmask32(len);
  }
  return key_bits(k, pos, len);
}

}  // namespace apgk
