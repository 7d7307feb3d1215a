// local.cuh -- per-bucket "sort and count" in shared memory.
//
// This is the reference's sorter/counter (the sort + run-length scan of
// SortKmers / a KmerParcel, and the (k-mer, frequency) records of naif_kmer's
// KmerKmerFreq -- SURVEY.md section 2.2; no file:line available) for one
// bucket of keys that share their leading P = D0+D1 bits:
//
//   1. the bucket is staged in shared memory;
//   2. an open-addressed shared-memory table groups identical keys with one
//      CAS/ADD per key (slot word = first-occurrence index | count << 16);
//   3. the DISTINCT keys (typically n / coverage of them) are LSD radix sorted
//      in shared memory, 8-bit digits, stable ranking by warp ballots;
//   4. (key, count) records are emitted in ascending key order and the
//      spectrum histogram is updated.
//
// Buckets larger than LOCAL_MAX take k_big: one CTA walks the bucket depth
// first with 8-bit MSD splits through a scratch buffer until the pieces fit.
#pragma once
#include "partition.cuh"

namespace apgk {

constexpr int LOCAL_NT = 256;
constexpr int LOCAL_NW = LOCAL_NT / 32;
constexpr int SPEC_SMEM = 1024;        // per-CTA shared spectrum bins
constexpr uint32_t SPEC_DENSE = 65536; // dense device spectrum bins; larger counts go to the overflow list
constexpr uint32_t SLOT_EMPTY = 0xFFFFFFFFu;

template <typename Elem>
struct ElemOps;
template <>
struct ElemOps<uint32_t> {
  __device__ __forceinline__ static uint32_t hash(uint32_t e) { return e * 0x9E3779B1u; }
  __device__ __forceinline__ static bool eq(uint32_t a, uint32_t b) { return a == b; }
  __device__ __forceinline__ static uint32_t bits(uint32_t e, int pos, int len) { return (e >> pos) & lowmask32(len); }
};
template <int W>
struct ElemOps<Key<W>> {
  __device__ __forceinline__ static uint32_t hash(const Key<W>& e) {
    uint64_t h = e.w[W - 1] * 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < W - 1; i++) h = (h ^ e.w[i]) * 0xBF58476D1CE4E5B9ull;
    return (uint32_t)(h >> 32);
  }
  __device__ __forceinline__ static bool eq(const Key<W>& a, const Key<W>& b) { return key_eq(a, b); }
  __device__ __forceinline__ static uint32_t bits(const Key<W>& e, int pos, int len) { return key_bits(e, pos, len); }
};

// How records leave a bucket.
template <int W>
struct EmitCtx {
  int want_table;
  int rem_bits, pad;            // to rebuild a full key from (bucket prefix, remainder)
  Key<W>* tmp_keys;             // temp keys, indexed by element offset (the dead level-0 buffer)
  uint32_t* tmp_cnt;            // temp counts, same indexing (own buffer: passes may re-read their input)
  unsigned long long* spec_dense;  // [SPEC_DENSE] device spectrum
  unsigned long long* spec_ovf;    // overflow list: [0] = count, then values
  uint32_t spec_ovf_cap;
};

template <int W>
__device__ __forceinline__ Key<W> rebuild_key(uint32_t rem, uint64_t prefix, int rem_bits, int pad) {
  Key<W> k;
#pragma unroll
  for (int i = 0; i < W; i++) k.w[i] = 0;
  k.w[W - 1] = ((prefix << rem_bits) | rem) >> pad;
  return k;
}
template <int W>
__device__ __forceinline__ Key<W> rebuild_key(const Key<W>& e, uint64_t, int, int) { return e; }

__device__ __forceinline__ void spec_add_global(unsigned long long* spec_dense, unsigned long long* spec_ovf,
                                                uint32_t ovf_cap, unsigned long long f) {
  if (f < SPEC_DENSE) atomicAdd(&spec_dense[f], 1ull);
  else {
    unsigned long long i = atomicAdd(&spec_ovf[0], 1ull);
    if (i < ovf_cap) spec_ovf[1 + i] = f;
  }
}

// Shared memory of a local CTA (LM = LOCAL_MAX elements):
//   elems[LM] | reps0[LM] u32 | tab[LM + LM/2 + 64] u32 (reps1 aliases tab) | cnt[LOCAL_NW][256] u32 |
//   spec[SPEC_SMEM] u32 | scratch[34] u32 | misc[8] u32
template <typename Elem>
struct LocalSmem {
  Elem* elems; uint32_t* reps0; uint32_t* tab; uint32_t* cnt; uint32_t* spec; uint32_t* scratch; uint32_t* misc;
  __device__ __forceinline__ void carve(unsigned char* raw, int LM) {
    elems = (Elem*)raw;
    size_t off = ((size_t)LM * sizeof(Elem) + 15) & ~(size_t)15;
    reps0 = (uint32_t*)(raw + off); off += (size_t)LM * 4;
    tab = (uint32_t*)(raw + off); off += ((size_t)LM + LM / 2 + 64) * 4;
    cnt = (uint32_t*)(raw + off); off += (size_t)LOCAL_NW * 256 * 4;
    spec = (uint32_t*)(raw + off); off += (size_t)SPEC_SMEM * 4;
    scratch = (uint32_t*)(raw + off); off += 34 * 4;
    misc = (uint32_t*)(raw + off);
  }
  static size_t bytes(int LM) {
    return (((size_t)LM * sizeof(Elem) + 15) & ~(size_t)15) + (size_t)LM * 4 + ((size_t)LM + LM / 2 + 64) * 4 +
           (size_t)LOCAL_NW * 256 * 4 + (size_t)SPEC_SMEM * 4 + 34 * 4 + 8 * 4;
  }
};

// Sort-and-count one bucket of n <= LOCAL_MAX elements.  All LOCAL_NT threads
// call.  Emits nd records at tmp index out_pos.. (keys) and cnt_dst[0..nd).
// sort_bits: number of low element bits that can differ inside the bucket.
// Returns nd (same value in every thread).
template <typename Elem, int W>
__device__ uint32_t local_bucket(LocalSmem<Elem>& sm, const Elem* __restrict__ src, uint32_t n, int sort_bits,
                                 uint64_t prefix, const EmitCtx<W>& ec, uint64_t out_pos, uint32_t* cnt_dst) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t nslots = n + n / 2 + 1;
  // ---- 1. stage + clear table
  for (uint32_t i = tid; i < n; i += LOCAL_NT) sm.elems[i] = src[i];
  for (uint32_t i = tid; i < nslots; i += LOCAL_NT) sm.tab[i] = SLOT_EMPTY;
  __syncthreads();
  // ---- 2. group identical keys
  volatile uint32_t* vtab = sm.tab;
  for (uint32_t i = tid; i < n; i += LOCAL_NT) {
    const Elem e = sm.elems[i];
    uint32_t slot = __umulhi(ElemOps<Elem>::hash(e), nslots);
    const uint32_t mine = i | (1u << 16);
    while (true) {
      uint32_t cur = vtab[slot];
      if (cur == SLOT_EMPTY) {
        cur = atomicCAS(&sm.tab[slot], SLOT_EMPTY, mine);
        if (cur == SLOT_EMPTY) break;
      }
      if (ElemOps<Elem>::eq(sm.elems[cur & 0xFFFFu], e)) {
        atomicAdd(&sm.tab[slot], 1u << 16);
        break;
      }
      if (++slot == nslots) slot = 0;
    }
  }
  __syncthreads();
  // ---- 3. compact occupied slots -> reps0
  const uint32_t per = (nslots + LOCAL_NT - 1) / LOCAL_NT;
  const uint32_t s0 = tid * per, s1 = min(s0 + per, nslots);
  uint32_t occ = 0;
  for (uint32_t s = s0; s < s1; s++) occ += (sm.tab[s] != SLOT_EMPTY);
  uint32_t wpos = block_excl_scan<LOCAL_NT>(occ, sm.scratch);
  const uint32_t nd = sm.scratch[32];
  for (uint32_t s = s0; s < s1; s++) {
    uint32_t v = sm.tab[s];
    if (v != SLOT_EMPTY) sm.reps0[wpos++] = v;
  }
  __syncthreads();
  // ---- 4. LSD radix sort of the nd representatives (stable ballot ranking); reps1 aliases the dead table
  uint32_t* rin = sm.reps0;
  uint32_t* rout = sm.tab;
  if (nd > 1 && sort_bits > 0) {
    const int npass = (sort_bits + 7) / 8;
    const int dl = (sort_bits + npass - 1) / npass;
    int wu = (int)((nd + 127) / 128);
    if (wu > LOCAL_NW) wu = LOCAL_NW;
    uint32_t chunk = (nd + wu - 1) / wu;
    chunk = (chunk + 31) & ~31u;
    for (int pass = 0; pass < npass; pass++) {
      const int pos = pass * dl;
      const int len = (pos + dl <= sort_bits) ? dl : sort_bits - pos;
      const int nb = 1 << len;
      for (int i = tid; i < wu * 256; i += LOCAL_NT) sm.cnt[i] = 0;
      __syncthreads();
      const uint32_t lo = wid * chunk, hi = min(lo + chunk, nd);
      if (wid < wu) {
        for (uint32_t i = lo + lane; i < hi; i += 32) {
          uint32_t d = ElemOps<Elem>::bits(sm.elems[rin[i] & 0xFFFFu], pos, len);
          atomicAdd(&sm.cnt[wid * 256 + d], 1u);
        }
      }
      __syncthreads();
      // per-bin totals over warps -> exclusive (bin-major, warp-minor) bases
      {
        uint32_t tot = 0;
        if (tid < nb) {
          for (int w = 0; w < wu; w++) {
            uint32_t v = sm.cnt[w * 256 + tid];
            sm.cnt[w * 256 + tid] = tot;
            tot += v;
          }
        }
        uint32_t base = block_excl_scan<LOCAL_NT>(tot, sm.scratch);
        if (tid < nb) {
          for (int w = 0; w < wu; w++) sm.cnt[w * 256 + tid] += base;
        }
      }
      __syncthreads();
      if (wid < wu) {
        for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
          const uint32_t i = i0 + lane;
          const bool act = i < hi;
          const uint32_t e = act ? rin[i] : 0u;
          const uint32_t d = act ? ElemOps<Elem>::bits(sm.elems[e & 0xFFFFu], pos, len) : 0u;
          uint32_t peers = __ballot_sync(0xffffffffu, act);
#pragma unroll
          for (int b = 0; b < 8; b++) {
            if (b < len) {
              const uint32_t v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
              peers &= ((d >> b) & 1u) ? v : ~v;
            }
          }
          const uint32_t before = __popc(peers & ((1u << lane) - 1u));
          uint32_t base = 0;
          if (act) base = sm.cnt[wid * 256 + d];
          __syncwarp();
          if (act && before == 0) sm.cnt[wid * 256 + d] = base + __popc(peers);
          __syncwarp();
          if (act) rout[base + before] = e;
        }
      }
      __syncthreads();
      uint32_t* t = rin; rin = rout; rout = t;
    }
  }
  // ---- 5. emit in ascending key order
  for (uint32_t j = tid; j < nd; j += LOCAL_NT) {
    const uint32_t e = rin[j];
    const uint32_t c = e >> 16;
    if (ec.want_table) {
      ec.tmp_keys[out_pos + j] = rebuild_key<W>(sm.elems[e & 0xFFFFu], prefix, ec.rem_bits, ec.pad);
      cnt_dst[j] = c;
    }
    if (c < SPEC_SMEM) atomicAdd(&sm.spec[c], 1u);
    else spec_add_global(ec.spec_dense, ec.spec_ovf, ec.spec_ovf_cap, c);
  }
  __syncthreads();
  return nd;
}

struct BucketTable {
  const unsigned long long* bofs;   // [nb] absolute element offset of each bucket
  const unsigned long long* bsize;  // [nb]
  uint32_t nb;                      // buckets [b0, nb) hold data (a shard owns a sub-range of the table)
  uint32_t local_max;
  uint32_t b0 = 0;
};

template <typename Elem>
__device__ __forceinline__ void spec_flush(LocalSmem<Elem>& sm, unsigned long long* spec_dense) {
  __syncthreads();
  for (int i = threadIdx.x; i < SPEC_SMEM; i += LOCAL_NT) {
    uint32_t v = sm.spec[i];
    if (v) atomicAdd(&spec_dense[i], (unsigned long long)v);
  }
}

// Persistent kernel over all level-1 buckets that fit in shared memory.
// With list == nullptr it walks all buckets; otherwise the bucket ids list[1 .. 1 + list[0]) (the
// deferred list k_local2 leaves behind).
template <typename Elem, int W>
__global__ void __launch_bounds__(LOCAL_NT) k_local(const Elem* __restrict__ src, BucketTable bt, int sort_bits,
                                                    EmitCtx<W> ec, uint32_t* __restrict__ nd_out,
                                                    const uint32_t* __restrict__ list, uint32_t list_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LocalSmem<Elem> sm;
  sm.carve(smem_raw, (int)bt.local_max);
  for (int i = threadIdx.x; i < SPEC_SMEM; i += LOCAL_NT) sm.spec[i] = 0;
  __syncthreads();
  const uint32_t n_iter = list ? (list[0] < list_cap ? list[0] : list_cap) : bt.nb - bt.b0;
  for (uint32_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
    const uint32_t b = list ? list[1 + it] : bt.b0 + it;
    const unsigned long long n = bt.bsize[b];
    if (n == 0) {
      if (threadIdx.x == 0) nd_out[b] = 0;
      continue;
    }
    if (n > bt.local_max) continue;  // k_big's job
    const unsigned long long o = bt.bofs[b];
    uint32_t* cnt_dst = ec.tmp_cnt + o;
    uint32_t nd = local_bucket<Elem, W>(sm, src + o, (uint32_t)n, sort_bits, (uint64_t)b, ec, o, cnt_dst);
    if (threadIdx.x == 0) nd_out[b] = nd;
  }
  spec_flush(sm, ec.spec_dense);
}

// ---------------------------------------------------------------- oversize buckets
struct BigParams {
  const uint32_t* big_list;       // bucket ids with size > local_max
  const uint32_t* n_big;          // device scalar
  unsigned int* ticket;           // device scalar, zeroed
  unsigned long long* scratch_cursor;  // device scalar, zeroed
  void* scratch;                  // sum(big sizes) elements
  unsigned long long* stacks;     // gridDim.x * BIG_STACK * 2 words
};
constexpr int BIG_STACK = 8192;  // >= 256 * ceil(max sort bits / 8) = 256 * 24

// node = {rel offset (40 bits) | buf (1 bit) << 40 | bits left (8 bits) << 48 ,  n}
__device__ __forceinline__ unsigned long long big_pack(uint64_t rel, int buf, int bp) {
  return rel | ((unsigned long long)buf << 40) | ((unsigned long long)bp << 48);
}

template <typename Elem, int W>
__global__ void __launch_bounds__(LOCAL_NT) k_big(Elem* __restrict__ src, BucketTable bt, int sort_bits, EmitCtx<W> ec,
                                                  uint32_t* __restrict__ nd_out, BigParams bp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LocalSmem<Elem> sm;
  sm.carve(smem_raw, (int)bt.local_max);
  for (int i = threadIdx.x; i < SPEC_SMEM; i += LOCAL_NT) sm.spec[i] = 0;
  __syncthreads();
  unsigned long long* stack = bp.stacks + (size_t)blockIdx.x * BIG_STACK * 2;
  uint32_t* hist = sm.cnt;        // 256 bins
  uint32_t* cursor = sm.cnt + 256;
  const uint32_t nbig = *bp.n_big;
  const int tid = threadIdx.x;
  while (true) {
    if (tid == 0) sm.misc[0] = atomicAdd(bp.ticket, 1u);
    __syncthreads();
    const uint32_t bi = sm.misc[0];
    __syncthreads();
    if (bi >= nbig) break;
    const uint32_t b = bp.big_list[bi];
    const unsigned long long n = bt.bsize[b], o = bt.bofs[b];
    if (tid == 0) {
      unsigned long long so = atomicAdd(bp.scratch_cursor, n);
      sm.misc[2] = (uint32_t)so; sm.misc[3] = (uint32_t)(so >> 32);
      stack[0] = big_pack(0, 0, sort_bits); stack[1] = n;
      sm.misc[1] = 1;  // stack size
    }
    __syncthreads();
    const unsigned long long so = (unsigned long long)sm.misc[2] | ((unsigned long long)sm.misc[3] << 32);
    Elem* bufs[2] = {src + o, (Elem*)bp.scratch + so};
    uint32_t* cnt_base = ec.tmp_cnt + o;
    unsigned long long run_nd = 0;
    while (true) {
      const uint32_t sp = sm.misc[1];
      if (sp == 0) break;
      const unsigned long long node = stack[(sp - 1) * 2], m = stack[(sp - 1) * 2 + 1];
      __syncthreads();
      if (tid == 0) sm.misc[1] = sp - 1;
      const uint64_t rel = node & ((1ull << 40) - 1);
      const int buf = (int)((node >> 40) & 1);
      const int bits_left = (int)(node >> 48);
      const Elem* s = bufs[buf] + rel;
      if (m <= bt.local_max) {
        __syncthreads();
        uint32_t nd = local_bucket<Elem, W>(sm, s, (uint32_t)m, bits_left, (uint64_t)b, ec, o + run_nd,
                                            cnt_base + run_nd);
        run_nd += nd;
        continue;
      }
      if (bits_left == 0) {  // m identical keys
        if (tid == 0) {
          if (ec.want_table) {
            ec.tmp_keys[o + run_nd] = rebuild_key<W>(s[0], (uint64_t)b, ec.rem_bits, ec.pad);
            cnt_base[run_nd] = m > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)m;
          }
          spec_add_global(ec.spec_dense, ec.spec_ovf, ec.spec_ovf_cap, m);
        }
        run_nd += 1;
        __syncthreads();
        continue;
      }
      const int db = bits_left < 8 ? bits_left : 8;
      const int sh = bits_left - db;
      hist[tid] = 0;  // LOCAL_NT == 256
      __syncthreads();
      for (unsigned long long i = tid; i < m; i += LOCAL_NT) atomicAdd(&hist[ElemOps<Elem>::bits(s[i], sh, db)], 1u);
      __syncthreads();
      const uint32_t d_first = ElemOps<Elem>::bits(s[0], sh, db);
      if (hist[d_first] == m) {  // one bin holds everything: no data movement, just consume the digit
        __syncthreads();
        if (tid == 0) {
          stack[(sp - 1) * 2] = big_pack(rel, buf, sh);
          stack[(sp - 1) * 2 + 1] = m;
          sm.misc[1] = sp;
        }
        __syncthreads();
        continue;
      }
      const uint32_t my_cnt = hist[tid];
      const uint32_t my_start = block_excl_scan<LOCAL_NT>(my_cnt, sm.scratch);
      cursor[tid] = my_start;
      __syncthreads();
      Elem* dst = bufs[buf ^ 1] + rel;
      for (unsigned long long i = tid; i < m; i += LOCAL_NT) {
        const Elem e = s[i];
        uint32_t pos = atomicAdd(&cursor[ElemOps<Elem>::bits(e, sh, db)], 1u);
        dst[pos] = e;
      }
      __syncthreads();
      // children, pushed in descending digit order so the smallest digit is popped first
      hist[tid] = my_cnt;      // hist is intact, cursor now holds ends; keep starts in registers
      cursor[tid] = my_start;
      __syncthreads();
      if (tid == 0) {
        uint32_t top = sp - 1;
        for (int d = 255; d >= 0; d--) {
          if (hist[d]) {
            stack[top * 2] = big_pack(rel + cursor[d], buf ^ 1, sh);
            stack[top * 2 + 1] = hist[d];
            top++;
          }
        }
        sm.misc[1] = top;
      }
      __threadfence_block();
      __syncthreads();
    }
    if (tid == 0) nd_out[b] = (uint32_t)run_nd;
    __syncthreads();
  }
  spec_flush(sm, ec.spec_dense);
}

}  // namespace apgk
