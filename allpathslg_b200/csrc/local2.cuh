// local2.cuh -- the fast per-bucket "sort and count" kernel.
//
// k_local (local.cuh) pays ~45 block-wide barriers per bucket and leaves seven of
// its eight warps idle while the distinct keys are radix sorted; measured on B200 it
// ran at 23 Gkeys/s.  k_local2 keeps block-wide work to four short phases and gives
// every warp its own piece of the bucket:
//
//   phase 1  histogram of the next `sb` key bits (<= 64 sub-buckets)            [block]
//   phase 2  split the bucket into shared memory by sub-bucket, one shared
//            atomic per key (order inside a sub-bucket is irrelevant)            [block]
//   phase 3  each warp takes whole sub-buckets (~100-200 keys): groups identical
//            keys in a warp-private open-addressed table (one CAS/ADD per key),
//            remembers which lanes claimed new slots, splits the distinct keys by
//            6 more bits and ranks them inside those tiny groups by comparison  [warp, __syncwarp only]
//   phase 4  prefix sum of the sub-buckets' record counts, records written out in
//            ascending key order, spectrum updated                               [block]
//
// Buckets whose sub-buckets overflow the warp tables (low-complexity sequence) are
// appended to a deferred list and handled by k_local, bucket by bucket.
#pragma once
#include "local.cuh"

namespace apgk {

constexpr int L2_NT = 256;
constexpr int L2_NW = L2_NT / 32;
constexpr int L2_WCAP = 512;              // keys per sub-bucket a warp can take
constexpr int L2_WSLOTS = 512 + 128 + 16; // slots of the warp table
constexpr int L2_SUBMAX = 64;             // sub-buckets per bucket
constexpr int L2_SORTBINS = 64;
constexpr int L2_MAXBIN = 48;             // largest comparison-ranked group before we defer
constexpr int L2_TARGET_SUB = 160;        // aimed keys per sub-bucket

template <typename Elem>
struct Local2Smem {
  Elem* buf;          // [LM] bucket keys grouped by sub-bucket
  uint16_t* ord;      // [LM] element index of the j-th smallest distinct key of each sub-bucket
  uint16_t* ocnt;     // [LM] its multiplicity
  uint32_t* wtab;     // [NW][WSLOTS] idx | count << 16
  uint16_t* wrep;     // [NW][WCAP] slots claimed, in claim order
  uint16_t* wrep2;    // [NW][WCAP] the same, grouped by sort bin
  uint32_t* wbins;    // [NW][2][SORTBINS] counts / cursors
  uint32_t* shist;    // [SUBMAX] histogram, then running cursor
  uint32_t* sstart;   // [SUBMAX+1]
  uint32_t* snd;      // [SUBMAX] records per sub-bucket
  uint32_t* roff;     // [SUBMAX+1] exclusive prefix of snd
  uint32_t* spec;     // [SPEC_SMEM]
  uint32_t* misc;     // [8]
  __device__ __forceinline__ void carve(unsigned char* raw, int LM) {
    size_t off = 0;
    buf = (Elem*)raw; off += ((size_t)LM * sizeof(Elem) + 15) & ~(size_t)15;
    wtab = (uint32_t*)(raw + off); off += (size_t)L2_NW * L2_WSLOTS * 4;
    wbins = (uint32_t*)(raw + off); off += (size_t)L2_NW * 2 * L2_SORTBINS * 4;
    shist = (uint32_t*)(raw + off); off += L2_SUBMAX * 4;
    sstart = (uint32_t*)(raw + off); off += (L2_SUBMAX + 1) * 4;
    snd = (uint32_t*)(raw + off); off += L2_SUBMAX * 4;
    roff = (uint32_t*)(raw + off); off += (L2_SUBMAX + 1) * 4;
    spec = (uint32_t*)(raw + off); off += (size_t)SPEC_SMEM * 4;
    misc = (uint32_t*)(raw + off); off += 8 * 4;
    ord = (uint16_t*)(raw + off); off += (size_t)LM * 2;
    ocnt = (uint16_t*)(raw + off); off += (size_t)LM * 2;
    wrep = (uint16_t*)(raw + off); off += (size_t)L2_NW * L2_WCAP * 2;
    wrep2 = (uint16_t*)(raw + off);
  }
  static size_t bytes(int LM) {
    return (((size_t)LM * sizeof(Elem) + 15) & ~(size_t)15) + (size_t)L2_NW * L2_WSLOTS * 4 +
           (size_t)L2_NW * 2 * L2_SORTBINS * 4 + (4 * L2_SUBMAX + 2) * 4 + (size_t)SPEC_SMEM * 4 + 8 * 4 +
           (size_t)LM * 4 + (size_t)L2_NW * L2_WCAP * 4 + 16;
  }
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

template <typename Elem>
__device__ __forceinline__ bool elem_less(const Elem& a, const Elem& b);
template <>
__device__ __forceinline__ bool elem_less<uint32_t>(const uint32_t& a, const uint32_t& b) { return a < b; }
template <>
__device__ __forceinline__ bool elem_less<Key<1>>(const Key<1>& a, const Key<1>& b) { return key_less(a, b); }
template <>
__device__ __forceinline__ bool elem_less<Key<2>>(const Key<2>& a, const Key<2>& b) { return key_less(a, b); }
template <>
__device__ __forceinline__ bool elem_less<Key<3>>(const Key<3>& a, const Key<3>& b) { return key_less(a, b); }

// One warp: group + sort the ns keys buf[base .. base+ns).  Writes ord/ocnt[base + r] for the
// r-th smallest distinct key and returns the number of distinct keys, or 0xFFFFFFFF if a
// comparison group got too large (caller defers the bucket).  low_bits: key bits below the
// sub-bucket digit (they still distinguish keys inside the sub-bucket).
template <typename Elem>
__device__ __forceinline__ uint32_t warp_sort_count(Local2Smem<Elem>& sm, uint32_t base, uint32_t ns, int low_bits) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t* tab = sm.wtab + wid * L2_WSLOTS;
  volatile uint32_t* vtab = tab;
  uint16_t* rep = sm.wrep + wid * L2_WCAP;
  uint16_t* rep2 = sm.wrep2 + wid * L2_WCAP;
  uint32_t* bcnt = sm.wbins + wid * 2 * L2_SORTBINS;
  uint32_t* bcur = bcnt + L2_SORTBINS;
  const uint32_t nslots = ns + (ns >> 2) + 8;
  for (uint32_t i = lane; i < nslots; i += 32) tab[i] = SLOT_EMPTY;
  __syncwarp();
  // ---- group identical keys; lanes that claim a fresh slot append it to rep[]
  uint32_t nd = 0;
  for (uint32_t i0 = 0; i0 < ns; i0 += 32) {
    const uint32_t i = i0 + lane;
    int claimed = -1;
    if (i < ns) {
      const Elem e = sm.buf[base + i];
      uint32_t slot = __umulhi(ElemOps<Elem>::hash(e), nslots);
      const uint32_t mine = (base + i) | (1u << 16);
      while (true) {
        uint32_t cur = vtab[slot];
        if (cur == SLOT_EMPTY) {
          cur = atomicCAS(&tab[slot], SLOT_EMPTY, mine);
          if (cur == SLOT_EMPTY) { claimed = (int)slot; break; }
        }
        if (ElemOps<Elem>::eq(sm.buf[cur & 0xFFFFu], e)) { atomicAdd(&tab[slot], 1u << 16); break; }
        if (++slot == nslots) slot = 0;
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, claimed >= 0);
    if (claimed >= 0) rep[nd + __popc(m & lt_mask)] = (uint16_t)claimed;
    nd += __popc(m);
  }
  __syncwarp();
  if (nd == 1) {
    if (lane == 0) {
      const uint32_t t = tab[rep[0]];
      sm.ord[base] = (uint16_t)(t & 0xFFFFu);
      sm.ocnt[base] = (uint16_t)(t >> 16);
    }
    __syncwarp();
    return 1;
  }
  // ---- split the nd distinct keys by the next sbits bits
  const int sbits = low_bits < 6 ? low_bits : 6;
  const int shift2 = low_bits - sbits;
  const int nbin = 1 << sbits;
  for (int i = lane; i < 2 * L2_SORTBINS; i += 32) bcnt[i] = 0;
  __syncwarp();
  for (uint32_t j = lane; j < nd; j += 32) {
    const Elem e = sm.buf[tab[rep[j]] & 0xFFFFu];
    atomicAdd(&bcnt[ElemOps<Elem>::bits(e, shift2, sbits)], 1u);
  }
  __syncwarp();
  {
    const uint32_t c0 = bcnt[lane], c1 = bcnt[lane + 32];  // bins beyond nbin are zero
    const uint32_t i0 = warp_incl_scan(c0, lane);
    const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31);
    const uint32_t i1 = warp_incl_scan(c1, lane);
    bcur[lane] = i0 - c0;
    bcur[lane + 32] = t0 + i1 - c1;
    uint32_t mx = c0 > c1 ? c0 : c1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      uint32_t u = __shfl_xor_sync(0xffffffffu, mx, o);
      mx = u > mx ? u : mx;
    }
    if (mx > L2_MAXBIN) return 0xFFFFFFFFu;  // warp-uniform
  }
  (void)nbin;
  __syncwarp();
  for (uint32_t j = lane; j < nd; j += 32) {
    const uint16_t slot = rep[j];
    const Elem e = sm.buf[tab[slot] & 0xFFFFu];
    const uint32_t pos = atomicAdd(&bcur[ElemOps<Elem>::bits(e, shift2, sbits)], 1u);
    rep2[pos] = slot;
  }
  __syncwarp();
  // ---- rank inside each group by comparison (groups hold ~1 key)
  for (uint32_t j = lane; j < nd; j += 32) {
    const uint32_t t = tab[rep2[j]];
    const uint32_t idx = t & 0xFFFFu;
    const Elem e = sm.buf[idx];
    const uint32_t d = ElemOps<Elem>::bits(e, shift2, sbits);
    const uint32_t be = bcur[d], bs = be - bcnt[d];
    uint32_t r = 0;
    for (uint32_t q = bs; q < be; q++) {
      if (q != j) r += elem_less<Elem>(sm.buf[tab[rep2[q]] & 0xFFFFu], e) ? 1u : 0u;
    }
    sm.ord[base + bs + r] = (uint16_t)idx;
    sm.ocnt[base + bs + r] = (uint16_t)(t >> 16);
  }
  __syncwarp();
  return nd;
}

__device__ __forceinline__ int ilog2_ceil(uint32_t x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }

template <typename Elem, int W>
__global__ void __launch_bounds__(L2_NT) k_local2(const Elem* __restrict__ src, BucketTable bt, int rem_bits, EmitCtx<W> ec,
                                                  uint32_t* __restrict__ nd_out, uint32_t* __restrict__ deferred,
                                                  uint32_t deferred_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Local2Smem<Elem> sm;
  sm.carve(smem_raw, (int)bt.local_max);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < SPEC_SMEM; i += L2_NT) sm.spec[i] = 0;
  __syncthreads();
  for (uint32_t b = bt.b0 + blockIdx.x; b < bt.nb; b += gridDim.x) {
    const unsigned long long n64 = bt.bsize[b];
    if (n64 == 0) {
      if (tid == 0) nd_out[b] = 0;
      continue;
    }
    if (n64 > bt.local_max) continue;  // k_big's job
    const uint32_t n = (uint32_t)n64;
    const unsigned long long o = bt.bofs[b];
    const Elem* s = src + o;
    int sb = ilog2_ceil((n + L2_TARGET_SUB - 1) / L2_TARGET_SUB);
    if (sb > 6) sb = 6;
    if (sb > rem_bits) sb = rem_bits;
    const int S = 1 << sb;
    const int sub_shift = rem_bits - sb;
    // ---- phase 1: sub-bucket histogram
    if (tid < L2_SUBMAX) sm.shist[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += L2_NT) atomicAdd(&sm.shist[ElemOps<Elem>::bits(s[i], sub_shift, sb)], 1u);
    __syncthreads();
    if (wid == 0) {
      const uint32_t c0 = sm.shist[lane], c1 = sm.shist[lane + 32];
      const uint32_t i0 = warp_incl_scan(c0, lane);
      const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31);
      const uint32_t i1 = warp_incl_scan(c1, lane);
      sm.sstart[lane] = i0 - c0;
      sm.sstart[lane + 32] = t0 + i1 - c1;
      if (lane == 31) sm.sstart[64] = t0 + i1;
      uint32_t mx = c0 > c1 ? c0 : c1;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        uint32_t u = __shfl_xor_sync(0xffffffffu, mx, off);
        mx = u > mx ? u : mx;
      }
      if (lane == 0) sm.misc[0] = mx > L2_WCAP ? 1u : 0u;
      sm.shist[lane] = i0 - c0;  // becomes the scatter cursor
      sm.shist[lane + 32] = t0 + i1 - c1;
    }
    __syncthreads();
    bool defer = sm.misc[0] != 0;
    if (!defer) {
      // ---- phase 2: split into shared memory
      for (uint32_t i = tid; i < n; i += L2_NT) {
        const Elem e = s[i];
        const uint32_t pos = atomicAdd(&sm.shist[ElemOps<Elem>::bits(e, sub_shift, sb)], 1u);
        sm.buf[pos] = e;
      }
      __syncthreads();
      // ---- phase 3: per-warp group + sort
      for (int q = wid; q < S; q += L2_NW) {
        const uint32_t base = sm.sstart[q], ns = sm.sstart[q + 1] - base;
        uint32_t r = 0;
        if (ns) r = warp_sort_count<Elem>(sm, base, ns, sub_shift);
        if (lane == 0) {
          if (r == 0xFFFFFFFFu) { sm.misc[0] = 1u; r = 0; }
          sm.snd[q] = r;
        }
      }
      __syncthreads();
      defer = sm.misc[0] != 0;
    }
    if (defer) {
      if (tid == 0) {
        const uint32_t i = atomicAdd(&deferred[0], 1u);
        if (i < deferred_cap) deferred[1 + i] = b;
      }
      __syncthreads();
      continue;
    }
    // ---- phase 4: emit in ascending key order
    if (wid == 0) {
      const uint32_t c0 = lane < S ? sm.snd[lane] : 0u, c1 = lane + 32 < S ? sm.snd[lane + 32] : 0u;
      const uint32_t i0 = warp_incl_scan(c0, lane);
      const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31);
      const uint32_t i1 = warp_incl_scan(c1, lane);
      sm.roff[lane] = i0 - c0;
      sm.roff[lane + 32] = t0 + i1 - c1;
      if (lane == 31) {
        sm.roff[64] = t0 + i1;
        nd_out[b] = t0 + i1;
      }
    }
    __syncthreads();
    uint32_t* cnt_dst = ec.tmp_cnt + o;
    for (int q = wid; q < S; q += L2_NW) {
      const uint32_t base = sm.sstart[q], nd = sm.snd[q], ro = sm.roff[q];
      for (uint32_t j = lane; j < nd; j += 32) {
        const uint32_t c = sm.ocnt[base + j];
        if (ec.want_table) {
          ec.tmp_keys[o + ro + j] = rebuild_key<W>(sm.buf[sm.ord[base + j]], (uint64_t)b, ec.rem_bits, ec.pad);
          cnt_dst[ro + j] = c;
        }
        if (c < SPEC_SMEM) atomicAdd(&sm.spec[c], 1u);
        else spec_add_global(ec.spec_dense, ec.spec_ovf, ec.spec_ovf_cap, c);
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = tid; i < SPEC_SMEM; i += L2_NT) {
    const uint32_t v = sm.spec[i];
    if (v) atomicAdd(&ec.spec_dense[i], (unsigned long long)v);
  }
}

}  // namespace apgk
