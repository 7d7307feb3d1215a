// partition.cuh -- deterministic most-significant-digit partition passes.
//
// This is the "k-mer parcel builder": the reference's KmerParcelsBuilder /
// SortKmers passes cut k-mer space into parcels and build each one separately
// (names from BASELINE.json; no file:line available).
// Here a parcel is a bucket of the key's leading digits and a pass is three
// kernels with no global atomics and no spin-waits:
//
//   k_hist_*    one CTA per CHUNK of tiles: shared-memory histogram of the digit
//               -> chunksum[chunk][bin]
//   k_segscan   per segment, exclusive scan over its chunks and over its bins
//               -> chunksum becomes "first output offset of this chunk's keys of bin d"
//   k_scatter_* one CTA per chunk, tile after tile: the tile's keys stay in registers,
//               take their rank inside (tile, bin) from ONE shared-memory atomic each
//               (order inside a bin is irrelevant for an MSD pass), a block scan turns
//               the counts into stage slots, the keys are placed in bin order in a
//               shared-memory stage and leave as coalesced runs.
//
// HBM traffic per key and pass: 1 read (hist) + 1 read + 1 write (scatter); the only
// table is one row of 4-byte counters per chunk (< 1 % of the key bytes).  Output
// placement is a pure function of the input, so results are reproducible run to run.
#pragma once
#include "extract.cuh"
#include <cstdio>

// Debug build (-DAPGK_CHECKS): bounds checks that report and skip instead of faulting.
#ifdef APGK_CHECKS
#define APGK_CHECK(cond, ...) do { if (!(cond)) { printf(__VA_ARGS__); } } while (0)
#define APGK_OK_OR_SKIP(cond) (cond)
#else
#define APGK_CHECK(cond, ...) do { } while (0)
#define APGK_OK_OR_SKIP(cond) true
#endif

namespace apgk {

// Shared memory through 32-bit shared-window addresses and explicit ld/st/atom.shared: with generic
// pointers into the dynamic shared array ptxas re-derived the array bases (S2UR SR_CgaCtaId, ULEA, ...)
// for every key -- ~10 of ~30 instructions per key and phase in the scatter kernels' first SASS.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
  return r;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ uint64_t lds_u64(uint32_t addr) {
  uint64_t r;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
template <int W>
__device__ __forceinline__ void sts_key(uint32_t addr, const Key<W>& k) {
#pragma unroll
  for (int i = 0; i < W; i++) sts_u64(addr + 8 * i, k.w[i]);
}
template <int W>
__device__ __forceinline__ Key<W> lds_key(uint32_t addr) {
  Key<W> k;
#pragma unroll
  for (int i = 0; i < W; i++) k.w[i] = lds_u64(addr + 8 * i);
  return k;
}

// Predicated forms: a per-key `if` around an atomic or a store compiles to BSSY/BRA/BSYNC per key, and
// at read ends nearly every warp has both kinds of lanes, so the branch never skips anything.
__device__ __forceinline__ uint32_t atoms_add_if(uint32_t addr, uint32_t v, uint32_t pred) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.u32 %0, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
               : "=r"(r) : "r"(addr), "r"(v), "r"(pred) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u64_if(uint32_t addr, uint64_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u64 [%0], %1;\n\t}" ::"r"(addr), "l"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void sts_u32_if(uint32_t addr, uint32_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void stg_u64_if(void* ptr, uint64_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u64 [%0], %1;\n\t}" ::"l"(ptr), "l"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void stg_u32_if(void* ptr, uint32_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}" ::"l"(ptr), "r"(v), "r"(pred) : "memory");
}
// L2 eviction policies for the partition passes: the key stream is read once (evict_first), the
// partially written sectors at the ends of each bin's run should stay until the next tile completes them.
#ifndef APGK_LD_HINT
#define APGK_LD_HINT 1  // measured: level-1 scatter 24.6 -> 18.7 ms; store hints made no difference
#endif
#ifndef APGK_SK_EARLY
#define APGK_SK_EARLY 1
#endif
#ifndef APGK_ST_HINT
#define APGK_ST_HINT 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t ldg_u64_pol(const void* ptr, uint64_t pol) {
  uint64_t r;
  asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(ptr), "l"(pol) : "memory");
  return r;
}
__device__ __forceinline__ void stg_u64_if_pol(void* ptr, uint64_t v, uint32_t pred, uint64_t pol) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.L2::cache_hint.u64 [%0], %1, %3;\n\t}" ::"l"(ptr), "l"(v), "r"(pred), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_u32_if_pol(void* ptr, uint32_t v, uint32_t pred, uint64_t pol) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.L2::cache_hint.u32 [%0], %1, %3;\n\t}" ::"l"(ptr), "r"(v), "r"(pred), "l"(pol) : "memory");
}
template <int W>
__device__ __forceinline__ void stg_if(Key<W>* ptr, const Key<W>& k, uint32_t pred, uint64_t pol) {
#pragma unroll
  for (int i = 0; i < W; i++) {
    if (APGK_ST_HINT) stg_u64_if_pol(&ptr->w[i], k.w[i], pred, pol);
    else stg_u64_if(&ptr->w[i], k.w[i], pred);
  }
}
__device__ __forceinline__ void stg_if(uint32_t* ptr, const uint32_t& k, uint32_t pred, uint64_t pol) {
  if (APGK_ST_HINT) stg_u32_if_pol(ptr, k, pred, pol);
  else stg_u32_if(ptr, k, pred);
}
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t* ptr, uint64_t) { return *ptr; }
template <int W>
__device__ __forceinline__ Key<W> ldg_stream(const Key<W>* ptr, uint64_t pol) {
  if constexpr (APGK_LD_HINT == 0) return *ptr;
  else {
    Key<W> k;
#pragma unroll
    for (int i = 0; i < W; i++) k.w[i] = ldg_u64_pol(&ptr->w[i], pol);
    return k;
  }
}

__device__ __forceinline__ void reds_add_if(uint32_t addr, uint32_t v, uint32_t pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t r;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(r) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return r;
}
// returns `val` (as if the slot already held it) when pred == 0
__device__ __forceinline__ uint32_t atoms_cas_if(uint32_t addr, uint32_t cmp, uint32_t val, uint32_t pred) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\tmov.u32 %0, %3;\n\t@p atom.shared.cas.b32 %0, [%1], %2, %3;\n\t}"
               : "=r"(r) : "r"(addr), "r"(cmp), "r"(val), "r"(pred) : "memory");
  return r;
}

template <typename T> struct SmemElem;
template <int W> struct SmemElem<Key<W>> {
  __device__ __forceinline__ static void st(uint32_t a, const Key<W>& k) { sts_key<W>(a, k); }
  __device__ __forceinline__ static void st_if(uint32_t a, const Key<W>& k, uint32_t pred) {
#pragma unroll
    for (int i = 0; i < W; i++) sts_u64_if(a + 8 * i, k.w[i], pred);
  }
  __device__ __forceinline__ static Key<W> ld(uint32_t a) { return lds_key<W>(a); }
};
template <> struct SmemElem<uint32_t> {
  __device__ __forceinline__ static void st(uint32_t a, const uint32_t& k) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(k) : "memory");
  }
  __device__ __forceinline__ static void st_if(uint32_t a, const uint32_t& k, uint32_t pred) { sts_u32_if(a, k, pred); }
  __device__ __forceinline__ static uint32_t ld(uint32_t a) { return lds_u32(a); }
};

constexpr int DIGIT_BITS = 0;   // digit = key bits [pos, pos+len)
constexpr int DIGIT_OWNER = 1;  // digit = owner rank of the canonical k-mer (multi-GPU shuffle)

// Host-side description of a digit; the kernels take it as a DigitFn<MODE>.
struct DigitSpec {
  int mode;
  int pos, len, pad;
  uint32_t n_ranks;
};

// Device functor.  For one-word keys the digit is two uniform shifts and a mask
// (the generic pad/mode/lowmask code cost ~25 of 59 instructions per position in the first profile).
template <int MODE>
struct DigitFn {
  int kind;          // W == 1: left shift applied after the right shift (tiny K: digit reaches below bit 0), else 0
  int sh;            // W == 1: right shift of the one-word key
  uint32_t mask;
  int rsh;           // 32-bit elements: (e >> rsh) & mask
  int pos, len;      // W > 1: key_bits(k, pos, len)
  uint32_t n_ranks;  // MODE == DIGIT_OWNER
  uint32_t flo, fwidth;  // FILTER kernels keep only digits d with d - flo < fwidth (one k-mer-space round)
  template <int W>
  __device__ __forceinline__ uint32_t operator()(const Key<W>& k) const {
    if constexpr (MODE == DIGIT_OWNER) return key_owner(k, n_ranks);
    else if constexpr (W == 1) {
      // branch-free: low word of a 64-bit right shift (one SHF.R.U64 once the amount is known < 64),
      // then the tiny-K left shift (0 otherwise) and the mask
      const uint32_t t = (uint32_t)(k.w[0] >> (sh & 63));
      return (t << (kind & 31)) & mask;
    } else return key_bits(k, pos, len);
  }
  __device__ __forceinline__ uint32_t operator()(uint32_t e) const { return (e >> rsh) & mask; }
};
template <int MODE>
inline DigitFn<MODE> make_digit_fn(const DigitSpec& ds) {
  DigitFn<MODE> f;
  const int eff = ds.pos - ds.pad;  // position of the digit inside the real (unpadded) one-word key
  if (eff >= 0) { f.kind = 0; f.sh = eff; }
  else { f.kind = -eff; f.sh = 0; }
  f.mask = lowmask32(ds.len);
  f.rsh = ds.pos;
  f.pos = ds.pos; f.len = ds.len; f.n_ranks = ds.n_ranks;
  f.flo = 0; f.fwidth = 0xFFFFFFFFu;
  return f;
}

// ---------------------------------------------------------------- block scan
// Exclusive scan of one value per thread over a block of NT threads (NT multiple of 32).
// warp_scratch: 33 words of shared memory.  Returns the exclusive prefix; total via scratch[32].
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  __syncthreads();  // scratch may still be in use by a previous call
  if (lane == 31) warp_scratch[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = lane < NT / 32 ? warp_scratch[lane] : 0u, it = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) it += u;
    }
    warp_scratch[lane] = it - t;
    if (lane == 31) warp_scratch[32] = it;
  }
  __syncthreads();
  return incl - v + warp_scratch[wid];
}

// One-barrier variant: every warp re-scans the warp totals itself.  The caller guarantees the
// scratch words are not being read by an earlier use (a barrier since then).  total_out = block total.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan1(uint32_t v, uint32_t* warp_scratch, uint32_t& total_out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_scratch[wid] = incl;
  __syncthreads();
  uint32_t t = lane < NT / 32 ? warp_scratch[lane] : 0u, it = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, it, o);
    if (lane >= o) it += u;
  }
  total_out = __shfl_sync(0xffffffffu, it, 31);
  const uint32_t wbase = __shfl_sync(0xffffffffu, it - t, wid);
  return incl - v + wbase;
}

// ---------------------------------------------------------------- level plan
struct LevelPlan {
  int bins;
  int n_segments;
  int chunk_tiles;
  uint32_t n_tiles, n_chunks;
  uint32_t tile_elems;         // elements (or base positions) per tile
  const uint32_t* seg_tile0;   // [n_segments+1]
  const uint32_t* seg_chunk0;  // [n_segments+1]
  const uint64_t* seg_start;   // [n_segments+1] element offset of each segment
};

// largest s with arr[s] <= x  (arr non-decreasing, arr[0] <= x < arr[n])
__device__ __forceinline__ int seg_of(const uint32_t* __restrict__ arr, int n, uint32_t x) {
  int lo = 0, hi = n;  // invariant arr[lo] <= x < arr[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (arr[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// chunk c of segment s covers tiles [seg_tile0[s] + (c - seg_chunk0[s]) * CT, ...).
__device__ __forceinline__ void chunk_tiles(const LevelPlan& lp, uint32_t c, int& s, uint32_t& t0, uint32_t& t1) {
  s = seg_of(lp.seg_chunk0, lp.n_segments, c);
  t0 = lp.seg_tile0[s] + (c - lp.seg_chunk0[s]) * (uint32_t)lp.chunk_tiles;
  t1 = t0 + (uint32_t)lp.chunk_tiles;
  if (t1 > lp.seg_tile0[s + 1]) t1 = lp.seg_tile0[s + 1];
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------- hist from packed reads
struct ReadStore {
  const uint32_t* bases32;   // 16 bases per word, zero padded
  const uint32_t* starts32;  // 1 bit per base: a read starts here
  uint64_t total_bases;
  int K;
};

// One CTA per CHUNK of tiles: shared-memory histogram over the whole chunk -> one row of chunksum.
// top_bits > 0 selects the cheap level-0 formulation (top_digits16): the digit is the top `top_bits`
// bits of the canonical k-mer and nothing else of the k-mer is needed.
template <int W, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_hist_reads(ReadStore rs, DigitFn<MODE> dg, LevelPlan lp, int top_bits,
                                                   uint32_t* __restrict__ chunksum, uint32_t chunk0 = 0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = (uint32_t*)smem_raw;
  const int bins = lp.bins;
  for (int i = threadIdx.x; i < bins; i += NT) hist[i] = 0;
  __syncthreads();
  int s; uint32_t t0, t1;
  const uint32_t chunk = blockIdx.x + chunk0;  // a launch may cover a slice of the chunks (streamed ingest)
  chunk_tiles(lp, chunk, s, t0, t1);
  const uint32_t hist_a = smem_u32(hist);
  for (uint32_t t = t0; t < t1; t++) {
    const uint64_t p = ((uint64_t)t * NT + threadIdx.x) * POS_PER_THREAD;
    if (p < rs.total_bases) {
      const uint32_t valid = window_valid_mask16(rs.starts32, p, rs.K, rs.total_bases);
      if (valid) {
        if (MODE == DIGIT_BITS && top_bits > 0) {
          top_digits16(rs.bases32, p, rs.K, top_bits, [&](int j, uint32_t d) {
            if ((valid >> j) & 1u) reds_add(hist_a + 4 * d, 1u);
          });
        } else {
          Window16<W> win;
          load_window16<W>(rs.bases32, p, rs.K, win);
          extract16<W>(win, rs.K, [&](int j, const Key<W>& c, bool) {
            if ((valid >> j) & 1u) reds_add(hist_a + 4 * dg(c), 1u);
          });
        }
      }
    }
  }
  __syncthreads();
  uint32_t* row = chunksum + (size_t)chunk * bins;
  for (int i = threadIdx.x; i < bins; i += NT) row[i] = hist[i];
}

// ---------------------------------------------------------------- scatter: one CTA per CHUNK of tiles
// Shared layout: stage[tile_elems] | G[bins] u64 | Gabs[bins] u64 | cnt[2][bins] u32 | scratch[40]
//   Gabs[d] absolute output index of the chunk's next key of bin d (advanced tile by tile)
//   G[d]    Gabs[d] minus the first stage slot of bin d in the current tile:  out index = G[d] + stage slot
//   cnt[t&1][d]  phase 1: keys of bin d in tile t (each key learns its rank from the atomic);
//                after the scan: first stage slot of bin d.  The other buffer is cleared for tile t+1.
// Ranking is two-phase with the tile's keys held in registers, so no per-tile table exists in HBM.
template <typename Elem>
__device__ __forceinline__ void scatter_smem_carve(unsigned char* raw, uint32_t tile_elems, int bins, Elem*& stage,
                                                   uint32_t*& cnt, unsigned long long*& G, unsigned long long*& Gabs,
                                                   uint32_t*& scratch) {
  stage = (Elem*)raw;
  size_t off = ((size_t)tile_elems * sizeof(Elem) + 15) & ~(size_t)15;
  G = (unsigned long long*)(raw + off);
  off += (size_t)bins * 8;
  Gabs = (unsigned long long*)(raw + off);
  off += (size_t)bins * 8;
  cnt = (uint32_t*)(raw + off);
  off += (size_t)bins * 8;
  scratch = (uint32_t*)(raw + off);
}
template <typename Elem>
inline size_t scatter_smem_bytes(uint32_t tile_elems, int bins) {
  return (((size_t)tile_elems * sizeof(Elem) + 15) & ~(size_t)15) + (size_t)bins * 24 + 40 * 4;
}

// Thread-contiguous ownership of bins: thread t owns bins [t*per, (t+1)*per).
// chunk start: Gabs[d] = first output index of bin d for this chunk, both counter buffers = 0.
template <int NT>
__device__ __forceinline__ void chunk_begin(const uint32_t* __restrict__ chunk_row, const uint64_t* __restrict__ bstart64,
                                            uint64_t seg_base, int bins, unsigned long long* Gabs, uint32_t* cnt) {
  const int per = (bins + NT - 1) / NT;
  for (int j = 0; j < per; j++) {
    const int d = threadIdx.x * per + j;
    if (d < bins) {
      unsigned long long g = seg_base + chunk_row[d];
      if (bstart64) g += bstart64[d];
      Gabs[d] = g;
      cnt[d] = 0;
      cnt[bins + d] = 0;
    }
  }
}
// After phase 1 (+ a barrier): cur[] holds the tile's counts.  Turns them into stage slots, sets G for
// the write-out, advances Gabs past the tile, clears the other counter buffer for the next tile, and
// returns the tile's key count.  One barrier inside, one at the end.
template <int NT>
__device__ __forceinline__ uint32_t tile_scan(int bins, unsigned long long* G, unsigned long long* Gabs, uint32_t* cur,
                                              uint32_t* nxt, uint32_t* scratch) {
  const int per = (bins + NT - 1) / NT;
  const int d0 = threadIdx.x * per;
  uint32_t sum = 0;
  for (int j = 0; j < per; j++) {
    const int d = d0 + j;
    if (d < bins) sum += cur[d];
  }
  uint32_t total;
  uint32_t excl = block_excl_scan1<NT>(sum, scratch, total);
  for (int j = 0; j < per; j++) {
    const int d = d0 + j;
    if (d < bins) {
      const uint32_t c = cur[d];
      const unsigned long long ga = Gabs[d];
      cur[d] = excl;       // first stage slot of bin d
      G[d] = ga - excl;    // out = G[d] + slot
      Gabs[d] = ga + c;
      nxt[d] = 0;
      excl += c;
    }
  }
  __syncthreads();
  return total;
}

// Software-pipelined over tiles like k_scatter_keys below: the extraction and ranking of tile t+1
// (ALU + shared atomics) is interleaved, position by position, with the write-out of tile t
// (LDS -> LDS -> STG), so the one resident CTA per SM overlaps its compute with its stores.
// NPOS window starts per thread (16 for one-word k-mers; 8 / 4 for two / three words so that the keys
// of a tile still fit the registers of a 1024-thread CTA -- with 16 the W=2 kernel spilled 200 bytes/thread).
template <int W, int NT, int MODE, bool FILTER, int NPOS = POS_PER_THREAD>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) k_scatter_reads(ReadStore rs, DigitFn<MODE> dg, LevelPlan lp,
                                                      const uint32_t* __restrict__ chunkpref,
                                                      const uint64_t* __restrict__ bstart64, Key<W>* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Key<W>* stage; uint32_t* cnt2; unsigned long long* G; unsigned long long* Gabs; uint32_t* scratch;
  const int bins = lp.bins;
  scatter_smem_carve<Key<W>>(smem_raw, NT * NPOS, bins, stage, cnt2, G, Gabs, scratch);
  int s; uint32_t t0, t1;
  chunk_tiles(lp, blockIdx.x, s, t0, t1);
  chunk_begin<NT>(chunkpref + (size_t)blockIdx.x * bins, bstart64, 0ull, bins, Gabs, cnt2);
  constexpr uint32_t ES = (uint32_t)sizeof(Key<W>);
  const uint32_t stage_a = smem_u32(stage), G_a = smem_u32(G), cnt2_a = smem_u32(cnt2);
  const uint64_t pol_st = APGK_ST_HINT == 1 ? l2_policy_evict_last() : (APGK_ST_HINT == 2 ? l2_policy_evict_first() : 0ull);
  Key<W> key[NPOS];
  uint32_t rk[NPOS / 2];  // two 16-bit ranks per register
  uint32_t valid = 0;               // bit j: window j of the tile in key[] is a k-mer of this round
  // tile t: 16 windows per thread -> key[], valid; each k-mer takes its rank inside (tile, bin) from one
  // shared atomic.  between(j) runs after window j (the write-out of the previous tile hooks in here).
  auto extract_rank = [&](uint32_t t, uint32_t cnt_a, auto&& between) {
    const uint64_t p = ((uint64_t)t * NT + threadIdx.x) * NPOS;
    const bool inside = p < rs.total_bases;
    valid = inside ? window_valid_mask16(rs.starts32, p, rs.K, rs.total_bases) : 0u;
    if (NPOS < 16) valid &= (1u << NPOS) - 1u;
#pragma unroll
    for (int j = 0; j < NPOS / 2; j++) rk[j] = 0;
    Window16<W> win;
    if (NPOS == 16) load_window16<W>(rs.bases32, inside ? p : 0ull, rs.K, win);
    else load_window_at<W>(rs.bases32, inside ? p : 0ull, rs.K, win);
    extractN<W, NPOS>(win, rs.K, [&](int j, const Key<W>& c, bool) {
      key[j] = c;
      const uint32_t d = dg(c);
      if (FILTER) valid &= ~((uint32_t)((d - dg.flo) >= dg.fwidth) << j);  // belongs to another round
      const uint32_t r = atoms_add_if(cnt_a + 4 * d, 1u, valid & (1u << j));
      rk[j >> 1] |= r << ((j & 1) * 16);
      between(j);
    });
  };
  __syncthreads();  // chunk_begin
  extract_rank(t0, cnt2_a, [](int) {});
  for (uint32_t t = t0; t < t1; t++) {
    const uint32_t cur = (t - t0) & 1u;
    uint32_t* cnt = cnt2 + cur * bins;
    uint32_t* cnt_next = cnt2 + (cur ^ 1u) * bins;
    const uint32_t cnt_a = cnt2_a + cur * 4u * (uint32_t)bins, cnt_next_a = cnt2_a + (cur ^ 1u) * 4u * (uint32_t)bins;
    __syncthreads();  // ranks of tile t complete; write-out of tile t-1 done with stage and G
    const uint32_t tile_n = tile_scan<NT>(bins, G, Gabs, cnt, cnt_next, scratch);
    // ---- place the keys in bin order in the stage
    // (the shared-memory asm statements keep their program order: batch the loads by hand so that
    // eight LDS are in flight instead of one LDS -> STS chain at a time)
    constexpr int PB = NPOS >= 8 ? 8 : NPOS;
#pragma unroll
    for (int j0 = 0; j0 < NPOS; j0 += PB) {
      uint32_t first[PB];
#pragma unroll
      for (int j = j0; j < j0 + PB; j++) first[j - j0] = lds_u32(cnt_a + 4 * dg(key[j]));
#pragma unroll
      for (int j = j0; j < j0 + PB; j++) {
        const uint32_t slot = first[j - j0] + ((rk[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
        SmemElem<Key<W>>::st_if(stage_a + slot * ES, key[j], valid & (1u << j));
      }
    }
    __syncthreads();
    // ---- coalesced runs of tile t out, interleaved with the windows of tile t+1.
    // Slots past tile_n hold stale keys: reading them and their (in-range) bins is harmless.
    Key<W> kq[4];
    unsigned long long gq[4];
    auto write_out = [&](int j) {  // items j-3 .. j every fourth call: 4 LDS, 4 LDS, 4 STG
      if ((j & 3) == 1) {
#pragma unroll
        for (int u = 0; u < 4; u++) kq[u] = lds_key<W>(stage_a + ((uint32_t)(j - 1 + u) * NT + threadIdx.x) * ES);
      } else if ((j & 3) == 2) {
#pragma unroll
        for (int u = 0; u < 4; u++) gq[u] = lds_u64(G_a + 8 * dg(kq[u]));
      } else if ((j & 3) == 3) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t i = (uint32_t)(j - 3 + u) * NT + threadIdx.x;
          stg_if(out + gq[u] + i, kq[u], (uint32_t)(i < tile_n), pol_st);
        }
      }
    };
    if (t + 1 < t1) {
      if (t + 2 < t1) {  // the bases / start bits of tile t+2 toward L1
        const uint64_t pn = ((uint64_t)(t + 2) * NT + threadIdx.x) * NPOS;
        if (pn < rs.total_bases) {
          if ((threadIdx.x % (128 / NPOS)) == 0) prefetch_l1(rs.bases32 + (pn >> 4));   // threads sharing a 32-byte sector
          if ((threadIdx.x % (256 / NPOS)) == 0) prefetch_l1(rs.starts32 + (pn >> 5));
        }
      }
      extract_rank(t + 1, cnt_next_a, write_out);
    } else {
#pragma unroll
      for (int j = 0; j < NPOS; j++) write_out(j);
    }
  }
}

// ---------------------------------------------------------------- hist / scatter from a key array
// Element conversion on the way out of a pass.
template <typename Out, typename In>
struct ElemCvt;
template <typename T>
struct ElemCvt<T, T> {
  __device__ __forceinline__ static T cvt(const T& e, int, int) { return e; }
};
template <>
struct ElemCvt<uint32_t, Key<1>> {  // keep only the REM low bits of the virtual key
  __device__ __forceinline__ static uint32_t cvt(const Key<1>& e, int pad, int rem) {
    return (uint32_t)(e.w[0] << pad) & lowmask32(rem);
  }
};

// elements a thread holds in registers per tile
template <typename Elem> struct TileItems { static constexpr int N = sizeof(Elem) <= 8 ? 16 : (sizeof(Elem) <= 16 ? 8 : 5); };

// With d2 > 0 the histogram is taken over the WIDE digit dgw = (level digit << d2) | next d2 key bits: the chunk
// row is its fold over the low d2 bits, and the wide counts are added to sub[(segment * bins << d2) + wide digit]
// -- the sub-bucket sizes the sharded exchange needs (table.cuh k_gather_split), for one more shared-memory
// reduction per chunk instead of a second pass over the partition buffer.
template <typename Elem, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_hist_keys(const Elem* __restrict__ src, LevelPlan lp, DigitFn<MODE> dg,
                                                  uint32_t* __restrict__ chunksum, DigitFn<MODE> dgw = DigitFn<MODE>{}, int d2 = 0,
                                                  uint32_t* __restrict__ sub = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = (uint32_t*)smem_raw;
  const int bins = lp.bins;
  const int hbins = bins << d2;
  for (int i = threadIdx.x; i < hbins; i += NT) hist[i] = 0;
  __syncthreads();
  if (blockIdx.x >= lp.seg_chunk0[lp.n_segments]) return;  // the grid is a host-side bound of the device-made plan
  int s; uint32_t t0, t1;
  chunk_tiles(lp, blockIdx.x, s, t0, t1);
  const uint64_t seg_lo = lp.seg_start[s], seg_hi = lp.seg_start[s + 1];
  const uint64_t e0 = seg_lo + (uint64_t)(t0 - lp.seg_tile0[s]) * lp.tile_elems;
  uint64_t e1 = seg_lo + (uint64_t)(t1 - lp.seg_tile0[s]) * lp.tile_elems;
  if (e1 > seg_hi) e1 = seg_hi;
  constexpr int U = 4;
  const uint32_t hist_a = smem_u32(hist);
  for (uint64_t base = e0; base < e1; base += (uint64_t)NT * U) {
    Elem r[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = base + (uint64_t)u * NT + threadIdx.x;
      if (i < e1) r[u] = src[i];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = base + (uint64_t)u * NT + threadIdx.x;
      if (i < e1) reds_add(hist_a + 4 * (d2 ? dgw(r[u]) : dg(r[u])), 1u);
    }
  }
  __syncthreads();
  uint32_t* row = chunksum + (size_t)blockIdx.x * bins;
  if (d2 == 0) {
    for (int i = threadIdx.x; i < bins; i += NT) row[i] = hist[i];
  } else {
    uint32_t* gsub = sub + ((size_t)s * bins << d2);
    for (int i = threadIdx.x; i < bins; i += NT) {
      uint32_t t = 0;
      for (int j = 0; j < (1 << d2); j++) {
        const uint32_t v = hist[(i << d2) | j];
        if (v) atomicAdd(&gsub[(i << d2) | j], v);
        t += v;
      }
      row[i] = t;
    }
  }
}

// Software-pipelined over tiles: the loads and the ranking atomics of tile t+1 are issued in the same
// barrier-free block as the write-out of tile t (LDS -> LDS -> STG chains), so one resident CTA per SM
// still has independent global loads, shared atomics and global stores in flight together.  The first
// version ran the phases back to back and stalled on long_scoreboard / lg_throttle / barrier in turn
// (profiles/r01_ncu_v10_summary.txt: 25% issue utilisation).
template <typename ElemIn, typename ElemOut, int NT, int MODE, bool FILTER>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) k_scatter_keys(const ElemIn* __restrict__ src, LevelPlan lp, DigitFn<MODE> dg,
                                                     const uint32_t* __restrict__ chunkpref,
                                                     const uint64_t* __restrict__ bstart64, int out_pad, int out_rem,
                                                     ElemOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemIn* stage; uint32_t* cnt2; unsigned long long* G; unsigned long long* Gabs; uint32_t* scratch;
  const int bins = lp.bins;
  scatter_smem_carve<ElemIn>(smem_raw, lp.tile_elems, bins, stage, cnt2, G, Gabs, scratch);
  if (blockIdx.x >= lp.seg_chunk0[lp.n_segments]) return;  // the grid is a host-side bound of the device-made plan
  int s; uint32_t t0, t1;
  chunk_tiles(lp, blockIdx.x, s, t0, t1);
  const uint64_t seg_lo = lp.seg_start[s], seg_hi = lp.seg_start[s + 1];
  chunk_begin<NT>(chunkpref + (size_t)blockIdx.x * bins, bstart64, bstart64 ? 0ull : seg_lo, bins, Gabs, cnt2);
  constexpr int ITEMS = TileItems<ElemIn>::N;  // lp.tile_elems == NT * ITEMS
  constexpr uint32_t ES = (uint32_t)sizeof(ElemIn);
  const uint32_t stage_a = smem_u32(stage), G_a = smem_u32(G), cnt2_a = smem_u32(cnt2);
  const uint64_t chunk_e0 = seg_lo + (uint64_t)(t0 - lp.seg_tile0[s]) * lp.tile_elems;
  const uint64_t pol_ld = APGK_LD_HINT ? l2_policy_evict_first() : 0ull;
  const uint64_t pol_st = APGK_ST_HINT == 1 ? l2_policy_evict_last() : (APGK_ST_HINT == 2 ? l2_policy_evict_first() : 0ull);
  ElemIn r[ITEMS];
  uint32_t rk[(ITEMS + 1) / 2];
  uint32_t live = 0;  // bit u: item u of the tile in r[] exists and belongs to this round
  // tile t's keys -> r[], live
  auto load_tile = [&](uint32_t t) {
    const uint64_t e0 = chunk_e0 + (uint64_t)(t - t0) * lp.tile_elems;
    live = 0;
#pragma unroll
    for (int u = 0; u < ITEMS; u++) {
      const uint64_t i = e0 + (uint64_t)u * NT + threadIdx.x;
      if (i < seg_hi) { r[u] = ldg_stream(src + i, pol_ld); live |= 1u << u; }
      else r[u] = ElemIn{};
    }
  };
  // each live key of r[] takes its rank inside (tile, bin) from one shared atomic
  auto rank_tile = [&](uint32_t cnt_a) {
#pragma unroll
    for (int u = 0; u < (ITEMS + 1) / 2; u++) rk[u] = 0;
#pragma unroll
    for (int u = 0; u < ITEMS; u++) {
      const uint32_t d = dg(r[u]);
      if (FILTER) live &= ~((uint32_t)((d - dg.flo) >= dg.fwidth) << u);
      const uint32_t rr = atoms_add_if(cnt_a + 4 * d, 1u, live & (1u << u));
      rk[u >> 1] |= rr << ((u & 1) * 16);
    }
  };
  __syncthreads();  // chunk_begin
  load_tile(t0);
  rank_tile(cnt2_a);
  for (uint32_t t = t0; t < t1; t++) {
    const uint32_t cur = (t - t0) & 1u;
    uint32_t* cnt = cnt2 + cur * bins;
    uint32_t* cnt_next = cnt2 + (cur ^ 1u) * bins;
    const uint32_t cnt_a = cnt2_a + cur * 4u * (uint32_t)bins, cnt_next_a = cnt2_a + (cur ^ 1u) * 4u * (uint32_t)bins;
    __syncthreads();  // ranks of tile t complete; write-out of tile t-1 done with stage and G
    const uint32_t tile_n = tile_scan<NT>(bins, G, Gabs, cnt, cnt_next, scratch);
    // ---- place the keys in bin order in the stage
    {
      constexpr int PB = ITEMS >= 8 ? 8 : ITEMS;  // loads in flight per batch (asm statements keep program order)
#pragma unroll
      for (int u0 = 0; u0 < ITEMS; u0 += PB) {
        uint32_t first[PB];
#pragma unroll
        for (int u = u0; u < u0 + PB && u < ITEMS; u++) first[u - u0] = lds_u32(cnt_a + 4 * dg(r[u]));
#pragma unroll
        for (int u = u0; u < u0 + PB && u < ITEMS; u++) {
          const uint32_t slot = first[u - u0] + ((rk[u >> 1] >> ((u & 1) * 16)) & 0xFFFFu);
          SmemElem<ElemIn>::st_if(stage_a + slot * ES, r[u], live & (1u << u));
        }
      }
    }
    __syncthreads();
    // ---- one block: loads of tile t+1 | coalesced runs of tile t out | ranks of tile t+1
    const bool more = t + 1 < t1;
#if APGK_SK_EARLY
    if (more) load_tile(t + 1);
#else
    if (more) {  // one 128-byte line of tile t+1 per thread toward L2
      const uint64_t i = chunk_e0 + (uint64_t)(t + 1 - t0) * lp.tile_elems + (uint64_t)threadIdx.x * (128 / ES);
      if (i < seg_hi && threadIdx.x * (128 / ES) < lp.tile_elems) prefetch_l2(src + i);
    }
#endif
    // slots past tile_n hold stale keys: reading them and their (in-range) bins is harmless
    {
      constexpr int WB = ITEMS >= 4 ? 4 : ITEMS;
#pragma unroll
      for (int u0 = 0; u0 < ITEMS; u0 += WB) {
        ElemIn e[WB];
        unsigned long long g[WB];
#pragma unroll
        for (int u = u0; u < u0 + WB && u < ITEMS; u++) e[u - u0] = SmemElem<ElemIn>::ld(stage_a + ((uint32_t)u * NT + threadIdx.x) * ES);
#pragma unroll
        for (int u = u0; u < u0 + WB && u < ITEMS; u++) g[u - u0] = lds_u64(G_a + 8 * dg(e[u - u0]));
#pragma unroll
        for (int u = u0; u < u0 + WB && u < ITEMS; u++) {
          const uint32_t i = (uint32_t)u * NT + threadIdx.x;
          stg_if(out + g[u - u0] + i, ElemCvt<ElemOut, ElemIn>::cvt(e[u - u0], out_pad, out_rem), (uint32_t)(i < tile_n), pol_st);
        }
      }
    }
#if !APGK_SK_EARLY
    if (more) load_tile(t + 1);
#endif
    if (more) rank_tile(cnt_next_a);
  }
}

// ---------------------------------------------------------------- scatter with TMA bulk write-out
// Same pass as k_scatter_keys, but a bin's run leaves the stage as ONE cp.async.bulk.global.shared::cta (the 1-D bulk
// copy of the TMA unit, SASS UBLKCP) issued by the thread that owns the bin, instead of element by element through
// registers (LDS key -> LDS offset -> STG: 2 of the 5 shared-memory instructions and the one global store per key;
// ncu named the shared-memory instruction queue -- mio_throttle -- as this kernel's top stall).  Bulk copies move
// 16-byte granules between 16-byte aligned addresses, so
//   * the stage holds OUTPUT elements (converted while they are placed), and bin d's region starts at a slot
//     congruent to the run's global element index modulo U = 16 / gcd(16, element bytes);
//   * the at most U - 1 elements of a run's last, incomplete granule are CARRIED (in registers of the thread that owns
//     the bin) into the front of the bin's region of the next tile, so every copy starts and ends on a granule; only the
//     chunk's very first granule per bin (it may belong half to the previous chunk) and its very last one go out as
//     scalar stores.  (A first version wrote every run's head and tail elements with scalar stores: 6 000 four-byte
//     stores per tile to 1 024 different lines doubled the L2 write transactions -- 25.9 ms against 18.6.)
// The copies are asynchronous: the thread issues them, ranks the next tile's keys while the TMA unit drains the stage,
// and only then waits (cp.async.bulk.wait_group.read) before the stage is written again.  Needs bins <= NT (thread d
// owns bin d and keeps the bin's output cursor, its share of the first granule and its carry in registers).
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename Elem> struct BulkUnit {   // elements per 16-byte alignment period
  static constexpr uint32_t ES = (uint32_t)sizeof(Elem);
  static constexpr uint32_t U = ES % 16 == 0 ? 1u : (ES % 8 == 0 ? 2u : 4u);
};
template <typename ElemOut>
__host__ __device__ inline uint32_t scatter_bulk_stage_slots(uint32_t tile_elems, int bins) {
  return tile_elems + 2u * BulkUnit<ElemOut>::U * (uint32_t)bins;   // every bin's region is padded at both ends
}
template <typename ElemOut>
inline size_t scatter_bulk_smem_bytes(uint32_t tile_elems, int bins) {
  return (((size_t)scatter_bulk_stage_slots<ElemOut>(tile_elems, bins) * sizeof(ElemOut) + 15) & ~(size_t)15) + (size_t)bins * 24 + 40 * 4;
}

template <typename ElemIn, typename ElemOut, int NT, int MODE, bool FILTER>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) k_scatter_keys_bulk(const ElemIn* __restrict__ src, LevelPlan lp, DigitFn<MODE> dg,
                                                          const uint32_t* __restrict__ chunkpref,
                                                          const uint64_t* __restrict__ bstart64, int out_pad, int out_rem,
                                                          ElemOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr uint32_t ESO = (uint32_t)sizeof(ElemOut);
  constexpr uint32_t U = BulkUnit<ElemOut>::U;
  constexpr int NCAR = U > 1 ? (int)U - 1 : 1;
  const int bins = lp.bins;   // <= NT (host)
  ElemOut* stage; uint32_t* cnt2; unsigned long long* G; unsigned long long* Gabs; uint32_t* scratch;
  scatter_smem_carve<ElemOut>(smem_raw, scatter_bulk_stage_slots<ElemOut>(lp.tile_elems, bins), bins, stage, cnt2, G, Gabs, scratch);
  if (blockIdx.x >= lp.seg_chunk0[lp.n_segments]) return;  // the grid is a host-side bound of the device-made plan
  int s; uint32_t t0, t1;
  chunk_tiles(lp, blockIdx.x, s, t0, t1);
  const uint64_t seg_lo = lp.seg_start[s], seg_hi = lp.seg_start[s + 1];
  chunk_begin<NT>(chunkpref + (size_t)blockIdx.x * bins, bstart64, bstart64 ? 0ull : seg_lo, bins, Gabs, cnt2);
  constexpr int ITEMS = TileItems<ElemIn>::N;  // lp.tile_elems == NT * ITEMS
  const uint32_t stage_a = smem_u32(stage), cnt2_a = smem_u32(cnt2);
  const uint64_t chunk_e0 = seg_lo + (uint64_t)(t0 - lp.seg_tile0[s]) * lp.tile_elems;
  const uint64_t pol_ld = APGK_LD_HINT ? l2_policy_evict_first() : 0ull;
  const bool owner = (int)threadIdx.x < bins;   // thread d owns bin d
  ElemIn r[ITEMS];
  uint32_t rk[(ITEMS + 1) / 2];
  uint32_t live = 0;  // bit u: item u of the tile in r[] exists and belongs to this round
  auto load_tile = [&](uint32_t t) {
    const uint64_t e0 = chunk_e0 + (uint64_t)(t - t0) * lp.tile_elems;
    live = 0;
#pragma unroll
    for (int u = 0; u < ITEMS; u++) {
      const uint64_t i = e0 + (uint64_t)u * NT + threadIdx.x;
      if (i < seg_hi) { r[u] = ldg_stream(src + i, pol_ld); live |= 1u << u; }
      else r[u] = ElemIn{};
    }
  };
  auto rank_tile = [&](uint32_t cnt_a) {
#pragma unroll
    for (int u = 0; u < (ITEMS + 1) / 2; u++) rk[u] = 0;
#pragma unroll
    for (int u = 0; u < ITEMS; u++) {
      const uint32_t d = dg(r[u]);
      if (FILTER) live &= ~((uint32_t)((d - dg.flo) >= dg.fwidth) << u);
      const uint32_t rr = atoms_add_if(cnt_a + 4 * d, 1u, live & (1u << u));
      rk[u >> 1] |= rr << ((u & 1) * 16);
    }
  };
  __syncthreads();  // chunk_begin
  // this thread's bin: where its next element goes, how many elements of the chunk's first (shared) granule are
  // still to come, and the elements of the last incomplete granule
  unsigned long long ga = owner ? Gabs[threadIdx.x] : 0ull;
  uint32_t hrem = owner ? (U - ((uint32_t)ga & (U - 1))) & (U - 1) : 0u;
  ElemOut car[NCAR];
#pragma unroll
  for (int i = 0; i < NCAR; i++) car[i] = ElemOut{};
  load_tile(t0);
  rank_tile(cnt2_a);
  for (uint32_t t = t0; t < t1; t++) {
    const uint32_t cur = (t - t0) & 1u;
    uint32_t* cnt = cnt2 + cur * bins;
    uint32_t* cnt_next = cnt2 + (cur ^ 1u) * bins;
    const uint32_t cnt_a = cnt2_a + cur * 4u * (uint32_t)bins, cnt_next_a = cnt2_a + (cur ^ 1u) * 4u * (uint32_t)bins;
    __syncthreads();  // ranks of tile t complete; every thread has waited for its bulk copies of tile t-1
    // ---- stage regions: bin d's starts at a multiple of U, its new elements follow ga % U slots later
    const uint32_t c = owner ? cnt[threadIdx.x] : 0u;
    const uint32_t ph = (uint32_t)ga & (U - 1);
    uint32_t total;
    const uint32_t S = block_excl_scan1<NT>(c ? (ph + c + (U - 1)) & ~(U - 1) : 0u, scratch, total);
    if (owner) {
      cnt[threadIdx.x] = S + ph;          // first stage slot of the bin's new elements
      cnt_next[threadIdx.x] = 0;
      if (c && hrem == 0) {               // the carried head of the granule
#pragma unroll
        for (int i = 0; i < NCAR; i++)
          if ((uint32_t)i < ph) SmemElem<ElemOut>::st(stage_a + (S + i) * ESO, car[i]);
      }
    }
    __syncthreads();
    // ---- place the converted elements in bin order in the stage
    {
      constexpr int PB = ITEMS >= 8 ? 8 : ITEMS;  // loads in flight per batch (asm statements keep program order)
#pragma unroll
      for (int u0 = 0; u0 < ITEMS; u0 += PB) {
        uint32_t first[PB];
#pragma unroll
        for (int u = u0; u < u0 + PB && u < ITEMS; u++) first[u - u0] = lds_u32(cnt_a + 4 * dg(r[u]));
#pragma unroll
        for (int u = u0; u < u0 + PB && u < ITEMS; u++) {
          const uint32_t slot = first[u - u0] + ((rk[u >> 1] >> ((u & 1) * 16)) & 0xFFFFu);
          SmemElem<ElemOut>::st_if(stage_a + slot * ESO, ElemCvt<ElemOut, ElemIn>::cvt(r[u], out_pad, out_rem), live & (1u << u));
        }
      }
    }
    fence_async_smem();   // the stage was written through the generic proxy; the bulk copies read it through the async proxy
    __syncthreads();
    // ---- the runs leave
    if (c) {
      uint32_t start, m;          // region slots [start, start + m) go to out[gidx ...]
      unsigned long long gidx;
      if (hrem) {                 // the chunk's first granule of this bin: its leading elements are not ours to rewrite
        const uint32_t x = hrem < c ? hrem : c;
        for (uint32_t i = 0; i < x; i++) out[ga + i] = SmemElem<ElemOut>::ld(stage_a + (S + ph + i) * ESO);
        hrem -= x;
        start = ph + x; gidx = ga + x; m = c - x;
      } else { start = 0; gidx = ga - ph; m = ph + c; }
      const uint32_t body = m & ~(U - 1);
      if (body) bulk_s2g(out + gidx, stage_a + (S + start) * ESO, body * ESO);
      if (hrem == 0) {            // the incomplete last granule travels on in registers
#pragma unroll
        for (int i = 0; i < NCAR; i++)
          if ((uint32_t)i < m - body) car[i] = SmemElem<ElemOut>::ld(stage_a + (S + start + body + i) * ESO);
      }
      ga += c;
    }
    bulk_commit();
    // ---- meanwhile: the next tile's keys and ranks
    if (t + 1 < t1) {
      load_tile(t + 1);
      rank_tile(cnt_next_a);
    }
    bulk_wait_read0();   // this thread's copies have read the stage
  }
  // ---- the chunk's last, incomplete granule of every bin
  if (owner && hrem == 0) {
    const uint32_t kc = (uint32_t)ga & (U - 1);
#pragma unroll
    for (int i = 0; i < NCAR; i++)
      if ((uint32_t)i < kc) out[ga - kc + i] = car[i];
  }
}

// ---------------------------------------------------------------- plans made on the device
// Exclusive scan of one 64-bit value per thread over the block (NT multiple of 32); total via scratch[32].
template <int NT>
__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long* scratch /* 33 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  __syncthreads();
  if (lane == 31) scratch[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned long long t = lane < NT / 32 ? scratch[lane] : 0ull, it = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) it += u;
    }
    scratch[lane] = it - t;
    if (lane == 31) scratch[32] = it;
  }
  __syncthreads();
  return incl - v + scratch[wid];
}

// What the host used to do between the two levels, with the level-0 totals it had to fetch first: the plan of
// a pass over the segments [s_lo, s_hi) of `tot` (segments outside the range count as empty), made by ONE CTA.
//   seg_start[s]  element offset of segment s relative to the start of segment s_lo   [n_seg + 1]
//   seg_tile0[s]  / seg_chunk0[s]   first tile / chunk of segment s                      [n_seg + 1]  (may be null)
// info[0] = elements in the range, info[1] = sum of ALL n_seg totals, info[2] |= 1 if a segment holds 2^32 or more.
constexpr int PLAN_NT = 1024;
__global__ void __launch_bounds__(PLAN_NT) k_plan_range(const unsigned long long* __restrict__ tot, int n_seg, int s_lo, int s_hi,
                                                         uint32_t tile_elems, uint32_t chunk_tiles,
                                                         unsigned long long* __restrict__ seg_start, uint32_t* __restrict__ seg_tile0,
                                                         uint32_t* __restrict__ seg_chunk0, unsigned long long* __restrict__ info) {
  __shared__ unsigned long long scratch[34];
  __shared__ unsigned long long carry[4];
  if (threadIdx.x == 0) { carry[0] = carry[1] = carry[2] = carry[3] = 0; }
  __syncthreads();
  unsigned int bad = 0;
  for (int s0 = 0; s0 < n_seg; s0 += PLAN_NT) {
    const int s = s0 + (int)threadIdx.x;
    const unsigned long long all = s < n_seg ? tot[s] : 0ull;
    const unsigned long long n = (s >= s_lo && s < s_hi) ? all : 0ull;
    if (all >= (1ull << 32)) bad = 1;
    const unsigned long long tiles = (n + tile_elems - 1) / tile_elems;
    const unsigned long long chunks = (tiles + chunk_tiles - 1) / chunk_tiles;
    const unsigned long long e_n = block_excl_scan_u64<PLAN_NT>(n, scratch);
    const unsigned long long t_n = scratch[32];
    const unsigned long long e_t = block_excl_scan_u64<PLAN_NT>(tiles, scratch);
    const unsigned long long t_t = scratch[32];
    const unsigned long long e_c = block_excl_scan_u64<PLAN_NT>(chunks, scratch);
    const unsigned long long t_c = scratch[32];
    const unsigned long long e_a = block_excl_scan_u64<PLAN_NT>(all, scratch);
    const unsigned long long t_a = scratch[32];
    (void)e_a;
    if (s < n_seg) {
      seg_start[s] = carry[0] + e_n;
      if (seg_tile0) seg_tile0[s] = (uint32_t)(carry[1] + e_t);
      if (seg_chunk0) seg_chunk0[s] = (uint32_t)(carry[2] + e_c);
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry[0] += t_n; carry[1] += t_t; carry[2] += t_c; carry[3] += t_a; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    seg_start[n_seg] = carry[0];
    if (seg_tile0) seg_tile0[n_seg] = (uint32_t)carry[1];
    if (seg_chunk0) seg_chunk0[n_seg] = (uint32_t)carry[2];
    if (info) { info[0] = carry[0]; info[1] = carry[3]; }
  }
  if (bad && info) atomicOr(&info[2], 1ull);
}

// ---------------------------------------------------------------- column scan over tiles
// One CTA per segment: chunk sums -> exclusive chunk prefixes (in place); bucket totals -> segtot64;
// exclusive scan of the totals -> bstart32 (offset of the bucket inside its segment) and, optionally,
// absolute bucket offsets.  With fold != 0 the bucket's start inside the segment is added to every
// chunk prefix, so a scatter CTA finds "offset of my chunk's first key of bin d inside the segment".
template <int NT>
__global__ void __launch_bounds__(NT) k_segscan(LevelPlan lp, uint32_t* __restrict__ chunksum,
                                                unsigned long long* __restrict__ segtot64,
                                                uint32_t* __restrict__ bstart32,
                                                unsigned long long* __restrict__ bofs /* may be null */, int fold) {
  __shared__ uint32_t scratch[34];
  __shared__ unsigned long long carry_s;
  const int s = blockIdx.x;
  const int bins = lp.bins;
  const uint32_t c0 = lp.seg_chunk0[s], c1 = lp.seg_chunk0[s + 1];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int d0 = 0; d0 < bins; d0 += NT) {
    const int d = d0 + threadIdx.x;
    unsigned long long run = 0;
    if (d < bins) {
      for (uint32_t c = c0; c < c1; c++) run += chunksum[(size_t)c * bins + d];
      segtot64[(size_t)s * bins + d] = run;
    }
    // exclusive scan of this slab's totals (32-bit is enough inside a segment < 2^32; host checks)
    uint32_t excl = block_excl_scan<NT>((uint32_t)run, scratch);
    unsigned long long carry = carry_s;
    if (d < bins) {
      const uint32_t bs = (uint32_t)(carry + excl);
      bstart32[(size_t)s * bins + d] = bs;
      if (bofs) bofs[(size_t)s * bins + d] = lp.seg_start[s] + carry + excl;
      uint32_t pre = fold ? bs : 0u;
      for (uint32_t c = c0; c < c1; c++) {
        const uint32_t v = chunksum[(size_t)c * bins + d];
        chunksum[(size_t)c * bins + d] = pre;
        pre += v;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + scratch[32];
    __syncthreads();
  }
}

}  // namespace apgk
