// extract.cuh -- canonical k-mer extraction from the 2-bit packed base stream.
//
// Replaces the inner loop of the reference's k-mer builders (the per-base
// "fw = fw<<2|b ; rc = rc>>2|(3-b)<<2(K-1) ; canonical = min(fw,rc)" roll that
// SURVEY.md section 8(a) attributes to kmers/SortKmers, kmers/KmerParcels and
// kmers/naif_kmer; no file:line exists -- the reference tree was empty).
//
// B200 formulation: the base stream is little-endian 2-bit (base q at bits
// [2q,2q+2)).  For a window starting at base p the K bases read as a
// little-endian integer ARE the reverse strand with the last base most
// significant, so   revcomp(window) = ~(stream >> 2p) & mask   -- one funnel
// shift per 32 bits, no roll, no warm-up.  The forward k-mer is initialised
// once per thread with a bit reversal (BREV) and then rolled two bits per
// position.  Each thread owns 16 consecutive window starts (one 32-bit word of
// bases), so every shift amount inside the unrolled loop is an immediate.
#pragma once
#include "kmer_types.cuh"

namespace apgk {

constexpr int POS_PER_THREAD = 16;

// Registers a thread needs for its 16 windows: 2W+1 stream words and the word
// holding bases [p+K, p+K+16) that feed the forward roll.
template <int W>
struct Window16 {
  uint32_t w[2 * W + 1];
  uint32_t v0;
};

// p must be a multiple of 16.  bases32 must be readable (zero padded) for
// 2W+3 words past the last base.
template <int W>
APGK_HD void load_window16(const uint32_t* __restrict__ bases32, uint64_t p, int K, Window16<W>& win) {
  const uint64_t wi = p >> 4;
  const int nw = (K + 15 + 15) >> 4;  // words that can hold needed bits: ceil((K+15)/16) <= 2W+1
#pragma unroll
  for (int m = 0; m < 2 * W + 1; m++) win.w[m] = (m < nw) ? bases32[wi + m] : 0u;
  const uint64_t vi = wi + (uint64_t)(K >> 4);
  const uint32_t ks = 2u * (uint32_t)(K & 15);
  win.v0 = funnel_r(bases32[vi], bases32[vi + 1], ks);
}

// Same for ANY window start p (not only multiples of 16): the words are funnel-shifted so that base p
// sits at bit 0, after which extractN below runs unchanged.  Used when a thread owns fewer than 16
// window starts (multi-word k-mers: 16 keys per thread do not fit the register file).
// bases32 must be readable (zero padded) for 2W+4 words past the last base.
template <int W>
APGK_HD void load_window_at(const uint32_t* __restrict__ bases32, uint64_t p, int K, Window16<W>& win) {
  const uint64_t wi = p >> 4;
  const uint32_t sh = 2u * (uint32_t)(p & 15);
  uint32_t raw[2 * W + 2];
#pragma unroll
  for (int m = 0; m < 2 * W + 2; m++) raw[m] = bases32[wi + m];
#pragma unroll
  for (int m = 0; m < 2 * W + 1; m++) win.w[m] = funnel_r(raw[m], raw[m + 1], sh);
  const uint64_t q = p + (uint64_t)K;
  const uint64_t vi = q >> 4;
  const uint32_t ks = 2u * (uint32_t)(q & 15);
  win.v0 = funnel_r(bases32[vi], bases32[vi + 1], ks);
}

// Calls f(j, canonical, canonical_is_reverse) for j = 0..NPOS-1 (window start p+j).
// Validity of each window is the caller's business (see window_valid_mask16).
template <int W, int NPOS, typename F>
APGK_HD void extractN(const Window16<W>& win, int K, F&& f) {
  const int topbits = 2 * K - 64 * (W - 1);
  const uint64_t topmask = lowmask64(topbits);
  // ---- forward k-mer of window 0: reverse the 2-bit groups of the LE window
  Key<W> fw;
  {
    uint64_t rm[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
      uint64_t x = (uint64_t)win.w[2 * i] | ((uint64_t)win.w[2 * i + 1] << 32);  // LE chunk i (i=0 least significant)
      rm[i] = swap_pairs(brev64(x));  // becomes word i of the reversed value, most significant first
    }
    const int s = 64 * W - 2 * K;  // 0..62
#pragma unroll
    for (int i = 0; i < W; i++) {
      uint64_t v = rm[i] >> s;
      if (i > 0 && s > 0) v |= rm[i - 1] << (64 - s);
      fw.w[i] = v;
    }
  }
#pragma unroll
  for (int j = 0; j < NPOS; j++) {
    // ---- reverse complement of window j straight from the stream
    Key<W> rc;
#pragma unroll
    for (int q = 0; q < W; q++) {
      uint32_t lo = funnel_r(win.w[2 * q], win.w[2 * q + 1], 2 * j);
      uint32_t hi = funnel_r(win.w[2 * q + 1], win.w[2 * q + 2], 2 * j);
      rc.w[W - 1 - q] = ~((uint64_t)lo | ((uint64_t)hi << 32));
    }
    rc.w[0] &= topmask;
    const bool use_rc = key_less(rc, fw);
    Key<W> c;
#pragma unroll
    for (int i = 0; i < W; i++) c.w[i] = use_rc ? rc.w[i] : fw.w[i];
    f(j, c, use_rc);
    // ---- roll the forward k-mer to window j+1
    if (j + 1 < NPOS) {
      const uint64_t b = (win.v0 >> (2 * j)) & 3u;
#pragma unroll
      for (int i = 0; i < W - 1; i++) fw.w[i] = (fw.w[i] << 2) | (fw.w[i + 1] >> 62);
      fw.w[W - 1] = (fw.w[W - 1] << 2) | b;
      fw.w[0] &= topmask;
    }
  }
}

template <int W, typename F>
APGK_HD void extract16(const Window16<W>& win, int K, F&& f) { extractN<W, POS_PER_THREAD>(win, K, static_cast<F&&>(f)); }

// Level-0 digits without building the k-mers.  The top D bits of canonical = min(fw, rc) are
// min(top D bits of fw, top D bits of rc): if the tops differ they decide the comparison, if they are
// equal both candidates share them.  The top bits of fw are the window's FIRST nb = ceil(D/2) bases,
// those of rc the complement of its LAST nb bases -- which, the stream being little-endian, is just
// ~(stream >> 2(p+K-nb)) & mask.  ~12 instructions per position instead of ~40 for the full k-mers.
// Requires nb <= 8 and K >= nb; one-word and multi-word k-mers alike.  f(j, digit) for j = 0..15.
template <typename F>
APGK_HD void top_digits16(const uint32_t* __restrict__ bases32, uint64_t p, int K, int D, F&& f) {
  const int nb = (D + 1) >> 1;
  const uint32_t mask = lowmask32(2 * nb);
  const uint64_t wi = p >> 4;
  // forward side: bases p .. p+15+nb-1 (at most 24 of them) live in two words
  const uint32_t f0 = bases32[wi], f1 = bases32[wi + 1];
  const uint64_t wf = ((uint64_t)f0 | ((uint64_t)f1 << 32)) >> (2 * nb);  // base (p + nb + j) at bits [2j, 2j+2)
  uint32_t x = f0 & mask;                                                  // first nb bases, little-endian
  x = (uint32_t)(swap_pairs(brev64((uint64_t)x)) >> (64 - 2 * nb));        // -> first base most significant
  // reverse side: the stream from base p + K - nb on, word-aligned in registers
  const uint64_t q = p + (uint64_t)(K - nb);
  const uint64_t ri = q >> 4;
  const uint32_t rs = 2u * (uint32_t)(q & 15);
  const uint32_t r0 = bases32[ri], r1 = bases32[ri + 1], r2 = bases32[ri + 2];
  const uint32_t u0 = funnel_r(r0, r1, rs), u1 = funnel_r(r1, r2, rs);
  const int drop = 2 * nb - D;  // 1 when D is odd
#pragma unroll
  for (int j = 0; j < POS_PER_THREAD; j++) {
    const uint32_t rc_top = ~funnel_r(u0, u1, 2 * j) & mask;
    const uint32_t d = (x < rc_top ? x : rc_top) >> drop;
    f(j, d);
    if (j + 1 < POS_PER_THREAD) x = ((x << 2) | ((uint32_t)(wf >> (2 * j)) & 3u)) & mask;
  }
}

// Bit j of the result is set iff window [p+j, p+j+K) lies inside one read:
// p+j+K <= total_bases and no read starts strictly inside the window.
// starts32: 1 bit per base (bit q of the bitmap = "a read starts at base q"),
// zero padded past the end.
APGK_HD uint32_t window_valid_mask16(const uint32_t* __restrict__ starts32, uint64_t p, int K, uint64_t total_bases) {
  if (p + (uint64_t)K > total_bases) return 0u;
  uint32_t ok = 0xFFFFu;
  const uint64_t room = total_bases - (uint64_t)K - p;  // largest valid j
  if (room < 15) ok = (2u << (uint32_t)room) - 1u;
  if (K < 2) return ok;
  const uint64_t lo = p + 1, hi = p + (uint64_t)K + 14;  // candidate interior positions (inclusive)
  uint32_t inv = 0;
  for (uint64_t wi = lo >> 5; wi <= (hi >> 5); wi++) {
    uint32_t bits = starts32[wi];
    if (wi == (lo >> 5)) bits &= 0xFFFFFFFFu << (uint32_t)(lo & 31);
    if (wi == (hi >> 5)) bits &= 0xFFFFFFFFu >> (31u - (uint32_t)(hi & 31));
    while (bits) {
#ifdef __CUDA_ARCH__
      const int b = __ffs((int)bits) - 1;
#else
      const int b = __builtin_ctz(bits);
#endif
      bits &= bits - 1;
      const int s = (int)((wi << 5) + (uint64_t)b - p);  // 1 .. K+14 : start strictly inside windows j with s-K < j < s
      const int j0 = s - K + 1 > 0 ? s - K + 1 : 0;
      const int j1 = s - 1 < 15 ? s - 1 : 15;
      if (j0 <= j1) inv |= ((2u << j1) - 1u) & ~((1u << j0) - 1u);
    }
  }
  return ok & ~inv;
}

}  // namespace apgk
