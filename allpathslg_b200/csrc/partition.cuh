// partition.cuh -- deterministic most-significant-digit partition passes.
//
// This is the "k-mer parcel builder": the reference's KmerParcelsBuilder /
// SortKmers passes cut k-mer space into parcels and build each one separately
// (names from BASELINE.json; no file:line available).
// Here a parcel is a bucket of the key's leading digits and a pass is three
// kernels with no global atomics and no spin-waits:
//
//   k_hist_*    per tile, a shared-memory histogram of the digit -> cnt16[tile][bin]
//   k_colsum / k_segscan / k_colapply
//               column-wise exclusive scan over tiles, per segment
//               -> base32[tile][bin] (offset of the tile's first key of that bin)
//   k_scatter_* per tile: keys are ranked into a shared-memory stage with one
//               shared-memory atomic each (ranking order inside a bin is
//               irrelevant for an MSD pass), then written out as coalesced runs.
//
// HBM traffic per key and pass: 1 read (hist) + 1 read + 1 write (scatter) plus
// (2+4)*2 bytes of table per tile-bin.  Output placement is a pure function of
// the input, so results are reproducible run to run.
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

constexpr int DIGIT_BITS = 0;   // digit = key bits [pos, pos+len)
constexpr int DIGIT_OWNER = 1;  // digit = owner rank of the canonical k-mer (multi-GPU shuffle)

struct DigitSpec {
  int mode;
  int pos, len, pad;
  uint32_t n_ranks;
  uint32_t lo, hi;  // keep only digits in [lo, hi) (k-mer space rounds); others are dropped
};

template <int W>
__device__ __forceinline__ uint32_t spec_digit(const DigitSpec& ds, const Key<W>& k) {
  if (ds.mode == DIGIT_OWNER) return key_owner(k, ds.n_ranks);
  return digit_of(k, ds.pos, ds.len, ds.pad);
}
__device__ __forceinline__ uint32_t spec_digit(const DigitSpec& ds, uint32_t e) {
  return (e >> ds.pos) & lowmask32(ds.len);
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

// Exclusive scan, in place, of arr[0..n) in shared memory by the whole block.
// Returns the total.  Every thread must call.
template <int NT>
__device__ __forceinline__ uint32_t block_scan_array(uint32_t* arr, int n, uint32_t* warp_scratch) {
  const int per = (n + NT - 1) / NT;
  const int b0 = threadIdx.x * per;
  uint32_t sum = 0;
  for (int j = 0; j < per; j++) {
    int b = b0 + j;
    if (b < n) sum += arr[b];
  }
  uint32_t excl = block_excl_scan<NT>(sum, warp_scratch);
  for (int j = 0; j < per; j++) {
    int b = b0 + j;
    if (b < n) {
      uint32_t v = arr[b];
      arr[b] = excl;
      excl += v;
    }
  }
  uint32_t total = warp_scratch[32];
  __syncthreads();
  return total;
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

// ---------------------------------------------------------------- hist / scatter from packed reads
struct ReadStore {
  const uint32_t* bases32;   // 16 bases per word, zero padded
  const uint32_t* starts32;  // 1 bit per base: a read starts here
  uint64_t total_bases;
  int K;
};

template <int W, int NT>
__global__ void __launch_bounds__(NT) k_hist_reads(ReadStore rs, DigitSpec ds, int bins, uint16_t* __restrict__ cnt16) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = (uint32_t*)smem_raw;
  for (int i = threadIdx.x; i < bins; i += NT) hist[i] = 0;
  __syncthreads();
  const uint64_t p = ((uint64_t)blockIdx.x * NT + threadIdx.x) * POS_PER_THREAD;
  if (p < rs.total_bases) {
    const uint32_t valid = window_valid_mask16(rs.starts32, p, rs.K, rs.total_bases);
    if (valid) {
      Window16<W> win;
      load_window16<W>(rs.bases32, p, rs.K, win);
      extract16<W>(win, rs.K, [&](int j, const Key<W>& c, bool) {
        if ((valid >> j) & 1u) {
          uint32_t d = spec_digit(ds, c);
          if (d >= ds.lo && d < ds.hi) atomicAdd(&hist[d], 1u);
        }
      });
    }
  }
  __syncthreads();
  uint16_t* row = cnt16 + (size_t)blockIdx.x * bins;
  for (int i = threadIdx.x; i < bins; i += NT) row[i] = (uint16_t)hist[i];
}

// Shared layout of a scatter CTA: stage[tile_elems] | cursor[bins] u32 | gbase[bins] u64 | scratch[33]
template <typename Elem>
__device__ __forceinline__ void scatter_smem_carve(unsigned char* raw, uint32_t tile_elems, int bins, Elem*& stage,
                                                   uint32_t*& cursor, unsigned long long*& gbase, uint32_t*& scratch) {
  stage = (Elem*)raw;
  size_t off = ((size_t)tile_elems * sizeof(Elem) + 15) & ~(size_t)15;
  gbase = (unsigned long long*)(raw + off);
  off += (size_t)bins * 8;
  cursor = (uint32_t*)(raw + off);
  off += (size_t)bins * 4;
  scratch = (uint32_t*)(raw + off);
}
template <typename Elem>
inline size_t scatter_smem_bytes(uint32_t tile_elems, int bins) {
  return (((size_t)tile_elems * sizeof(Elem) + 15) & ~(size_t)15) + (size_t)bins * 12 + 34 * 4;
}

// Prologue shared by both scatter kernels: cursor[] <- exclusive scan of the tile's
// counts; gbase[d] <- absolute output index of stage slot 0 if it belonged to bin d.
template <int NT>
__device__ __forceinline__ uint32_t scatter_prologue(const uint16_t* __restrict__ cnt_row,
                                                     const uint32_t* __restrict__ base_row,
                                                     const uint64_t* __restrict__ bstart64, uint64_t seg_base, int bins,
                                                     uint32_t* cursor, unsigned long long* gbase, uint32_t* scratch) {
  for (int i = threadIdx.x; i < bins; i += NT) cursor[i] = cnt_row[i];
  __syncthreads();
  uint32_t total = block_scan_array<NT>(cursor, bins, scratch);
  for (int i = threadIdx.x; i < bins; i += NT) {
    unsigned long long g = seg_base + base_row[i] - cursor[i];
    if (bstart64) g += bstart64[i];
    gbase[i] = g;
  }
  __syncthreads();
  return total;
}

template <int W, int NT>
__global__ void __launch_bounds__(NT) k_scatter_reads(ReadStore rs, DigitSpec ds, int bins,
                                                      const uint16_t* __restrict__ cnt16,
                                                      const uint32_t* __restrict__ base32,
                                                      const uint64_t* __restrict__ bstart64, Key<W>* __restrict__ out,
                                                      unsigned long long out_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Key<W>* stage; uint32_t* cursor; unsigned long long* gbase; uint32_t* scratch;
  scatter_smem_carve<Key<W>>(smem_raw, NT * POS_PER_THREAD, bins, stage, cursor, gbase, scratch);
  const size_t row = (size_t)blockIdx.x * bins;
  const uint32_t tile_n = scatter_prologue<NT>(cnt16 + row, base32 + row, bstart64, 0ull, bins, cursor, gbase, scratch);
  if (tile_n == 0) return;
  const uint64_t p = ((uint64_t)blockIdx.x * NT + threadIdx.x) * POS_PER_THREAD;
  if (p < rs.total_bases) {
    const uint32_t valid = window_valid_mask16(rs.starts32, p, rs.K, rs.total_bases);
    if (valid) {
      Window16<W> win;
      load_window16<W>(rs.bases32, p, rs.K, win);
      extract16<W>(win, rs.K, [&](int j, const Key<W>& c, bool) {
        if ((valid >> j) & 1u) {
          uint32_t d = spec_digit(ds, c);
          if (d >= ds.lo && d < ds.hi) {
            uint32_t pos = atomicAdd(&cursor[d], 1u);
            APGK_CHECK(pos < (uint32_t)(NT * POS_PER_THREAD) && d < (uint32_t)bins,
                       "scatter_reads: tile %u tid %d pos %u d %u bins %d tile_n %u\n", blockIdx.x, threadIdx.x, pos, d, bins, tile_n);
            if (APGK_OK_OR_SKIP(pos < (uint32_t)(NT * POS_PER_THREAD))) stage[pos] = c;
          }
        }
      });
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < tile_n; i += NT) {
    Key<W> k = stage[i];
    uint32_t d = spec_digit(ds, k);
    APGK_CHECK(d < (uint32_t)bins && gbase[d < (uint32_t)bins ? d : 0] + i < out_cap,
               "scatter_reads out: tile %u i %u d %u gbase %llu cap %llu\n", blockIdx.x, i, d, gbase[d < (uint32_t)bins ? d : 0], out_cap);
    if (APGK_OK_OR_SKIP(d < (uint32_t)bins && gbase[d] + i < out_cap)) out[gbase[d] + i] = k;
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

template <typename Elem, int NT>
__global__ void __launch_bounds__(NT) k_hist_keys(const Elem* __restrict__ src, LevelPlan lp, DigitSpec ds,
                                                  uint16_t* __restrict__ cnt16) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = (uint32_t*)smem_raw;
  const int bins = lp.bins;
  for (int i = threadIdx.x; i < bins; i += NT) hist[i] = 0;
  __syncthreads();
  const uint32_t tile = blockIdx.x;
  const int s = seg_of(lp.seg_tile0, lp.n_segments, tile);
  const uint64_t seg_lo = lp.seg_start[s], seg_hi = lp.seg_start[s + 1];
  const uint64_t e0 = seg_lo + (uint64_t)(tile - lp.seg_tile0[s]) * lp.tile_elems;
  const uint64_t e1 = e0 + lp.tile_elems < seg_hi ? e0 + lp.tile_elems : seg_hi;
  for (uint64_t i = e0 + threadIdx.x; i < e1; i += NT) {
    Elem e = src[i];
    uint32_t d = spec_digit(ds, e);
    if (d >= ds.lo && d < ds.hi) atomicAdd(&hist[d], 1u);
  }
  __syncthreads();
  uint16_t* row = cnt16 + (size_t)tile * bins;
  for (int i = threadIdx.x; i < bins; i += NT) row[i] = (uint16_t)hist[i];
}

template <typename ElemIn, typename ElemOut, int NT>
__global__ void __launch_bounds__(NT) k_scatter_keys(const ElemIn* __restrict__ src, LevelPlan lp, DigitSpec ds,
                                                     const uint16_t* __restrict__ cnt16,
                                                     const uint32_t* __restrict__ base32,
                                                     const uint64_t* __restrict__ bstart64, int out_pad, int out_rem,
                                                     ElemOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemIn* stage; uint32_t* cursor; unsigned long long* gbase; uint32_t* scratch;
  const int bins = lp.bins;
  scatter_smem_carve<ElemIn>(smem_raw, lp.tile_elems, bins, stage, cursor, gbase, scratch);
  const uint32_t tile = blockIdx.x;
  const int s = seg_of(lp.seg_tile0, lp.n_segments, tile);
  const uint64_t seg_lo = lp.seg_start[s], seg_hi = lp.seg_start[s + 1];
  const size_t row = (size_t)tile * bins;
  const uint32_t tile_n =
      scatter_prologue<NT>(cnt16 + row, base32 + row, bstart64, bstart64 ? 0ull : seg_lo, bins, cursor, gbase, scratch);
  if (tile_n == 0) return;
  const uint64_t e0 = seg_lo + (uint64_t)(tile - lp.seg_tile0[s]) * lp.tile_elems;
  const uint64_t e1 = e0 + lp.tile_elems < seg_hi ? e0 + lp.tile_elems : seg_hi;
  for (uint64_t i = e0 + threadIdx.x; i < e1; i += NT) {
    ElemIn e = src[i];
    uint32_t d = spec_digit(ds, e);
    if (d >= ds.lo && d < ds.hi) {
      uint32_t pos = atomicAdd(&cursor[d], 1u);
      stage[pos] = e;
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < tile_n; i += NT) {
    ElemIn e = stage[i];
    uint32_t d = spec_digit(ds, e);
    out[gbase[d] + i] = ElemCvt<ElemOut, ElemIn>::cvt(e, out_pad, out_rem);
  }
}

// ---------------------------------------------------------------- column scan over tiles
// chunk c of segment s covers tiles [seg_tile0[s] + (c - seg_chunk0[s]) * CT, ...).
__device__ __forceinline__ void chunk_tiles(const LevelPlan& lp, uint32_t c, int& s, uint32_t& t0, uint32_t& t1) {
  s = seg_of(lp.seg_chunk0, lp.n_segments, c);
  t0 = lp.seg_tile0[s] + (c - lp.seg_chunk0[s]) * (uint32_t)lp.chunk_tiles;
  t1 = t0 + (uint32_t)lp.chunk_tiles;
  if (t1 > lp.seg_tile0[s + 1]) t1 = lp.seg_tile0[s + 1];
}

template <int NT>
__global__ void __launch_bounds__(NT) k_colsum(LevelPlan lp, const uint16_t* __restrict__ cnt16,
                                               uint32_t* __restrict__ chunksum) {
  int s; uint32_t t0, t1;
  chunk_tiles(lp, blockIdx.x, s, t0, t1);
  const int bins = lp.bins;
  for (int d = threadIdx.x; d < bins; d += NT) {
    uint32_t sum = 0;
    for (uint32_t t = t0; t < t1; t++) sum += cnt16[(size_t)t * bins + d];
    chunksum[(size_t)blockIdx.x * bins + d] = sum;
  }
}

// One CTA per segment: chunk sums -> exclusive chunk prefixes (in place); bucket
// totals -> segtot64; exclusive scan of the totals -> bstart32 (offset of the
// bucket inside its segment) and, optionally, absolute bucket offsets.
template <int NT>
__global__ void __launch_bounds__(NT) k_segscan(LevelPlan lp, uint32_t* __restrict__ chunksum,
                                                unsigned long long* __restrict__ segtot64,
                                                uint32_t* __restrict__ bstart32,
                                                unsigned long long* __restrict__ bofs /* may be null */) {
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
      for (uint32_t c = c0; c < c1; c++) {
        uint32_t v = chunksum[(size_t)c * bins + d];
        chunksum[(size_t)c * bins + d] = (uint32_t)run;
        run += v;
      }
      segtot64[(size_t)s * bins + d] = run;
    }
    // exclusive scan of this slab's totals (32-bit is enough inside a segment < 2^32; host checks)
    uint32_t excl = block_excl_scan<NT>((uint32_t)run, scratch);
    unsigned long long carry = carry_s;
    if (d < bins) {
      bstart32[(size_t)s * bins + d] = (uint32_t)(carry + excl);
      if (bofs) bofs[(size_t)s * bins + d] = lp.seg_start[s] + carry + excl;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + scratch[32];
    __syncthreads();
  }
}

template <int NT>
__global__ void __launch_bounds__(NT) k_colapply(LevelPlan lp, const uint16_t* __restrict__ cnt16,
                                                 const uint32_t* __restrict__ chunksum,
                                                 const uint32_t* __restrict__ bstart32, int fold,
                                                 uint32_t* __restrict__ base32) {
  int s; uint32_t t0, t1;
  chunk_tiles(lp, blockIdx.x, s, t0, t1);
  const int bins = lp.bins;
  for (int d = threadIdx.x; d < bins; d += NT) {
    uint32_t run = chunksum[(size_t)blockIdx.x * bins + d];
    if (fold) run += bstart32[(size_t)s * bins + d];
    for (uint32_t t = t0; t < t1; t++) {
      uint32_t v = cnt16[(size_t)t * bins + d];
      base32[(size_t)t * bins + d] = run;
      run += v;
    }
  }
}

}  // namespace apgk
