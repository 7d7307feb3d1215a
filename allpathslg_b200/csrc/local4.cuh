// local4.cuh -- sort-and-count for FULL-KEY buckets (multi-word k-mers, and one-word k-mers whose
// remainder below the prefix exceeds 32 bits) with the row-hash scheme of local3.cuh.
//
// A key of 16 or 24 bytes cannot be claimed with one CAS, so a table slot holds a 32-bit TAG (a hash of
// the whole key), a count and the index of one REPRESENTATIVE instance inside the bucket:
//
//   pass A  every instance of the range CASes its tag into its row (row = monotone in the 32 key bits
//           below the range prefix, scrambled start inside the row, probing wraps inside the row);
//           the thread that claims a fresh slot records its instance index as the representative
//   pass B  every instance finds the slot carrying its tag again (plain loads) and compares its FULL
//           key with the representative's (read back through L2): equal -> one RED on the count;
//           different -> two keys of this range share a tag, and the range is redone with another
//           hash seed (nothing has been emitted yet)
//   emit    as in k_local3: occupancy bitmap -> dense list -> one thread per distinct key ranks it
//           against the <= 16 entries of its row by full-key comparison and writes (k-mer, count)
//
// The table costs 12 bytes per slot whatever the key width.  A bucket is a stack of key RANGES
// "(nbits, prefix): the keys whose nbits bits below the bucket prefix equal the prefix's"; a range
// whose row overflows is split on the next bit, down to a single key value if need be, so oversize,
// skewed and low-complexity buckets (poly-A, tandem repeats) take more passes instead of another
// kernel.  (It replaces k_local2 + k_local + k_big for these element types: 308 + 1 ms at K=48 and
// 335 + 265 ms at K=96 on the 24 M x 250 bp workload, profiles/r01_ksweep.jsonl.)
#pragma once
#include "local3.cuh"

namespace apgk {

// diagnostics: [0] passes, [1] passes that overflowed a row, [2] passes redone for a tag collision, [3] buckets
__device__ unsigned long long g_l4_dbg[8];

constexpr int L4_ROW = 16;
constexpr int L4_STACK = 200;  // > REM + 1 pending ranges (REM <= 170)

struct Local4Smem {
  uint32_t* tag; uint32_t* cnt; uint32_t* rep; uint32_t* bitmap; uint16_t* list; uint32_t* spec; uint32_t* wsum; uint32_t* misc;
  uint32_t* stack;
  static __host__ __device__ size_t slots(int LM) { return (((size_t)LM + LM / 4 + 1 + L3_SLACK) + 127) & ~(size_t)127; }
  __device__ __forceinline__ void carve(unsigned char* raw, int LM, int W) {
    const size_t ns = slots(LM);
    tag = (uint32_t*)raw;
    cnt = tag + ns;
    rep = cnt + ns;
    bitmap = rep + ns;
    spec = bitmap + ns / 32 + 4;
    wsum = spec + SPEC_SMEM;
    misc = wsum + 40;
    stack = misc + 8;
    list = (uint16_t*)(stack + (size_t)L4_STACK * (1 + 2 * W));
  }
  static size_t bytes(int LM, int W) {
    return (slots(LM) * 3 + slots(LM) / 32 + 4 + SPEC_SMEM + 40 + 8 + (size_t)L4_STACK * (1 + 2 * W)) * 4 +
           ((size_t)LM + LM / 4 + 128) * 2 + 16;
  }
};

// all bits at positions >= sh of (a ^ b) are zero
template <int W>
__device__ __forceinline__ bool key_top_equal(const Key<W>& a, const Key<W>& b, int sh) {
  const int q = sh >> 6, r = sh & 63;  // word q from the least significant end holds bit sh
  bool eq = true;
#pragma unroll
  for (int i = 0; i < W; i++) {
    const int j = W - 1 - i;  // index from the least significant word
    const uint64_t x = a.w[i] ^ b.w[i];
    if (j > q) eq = eq && (x == 0);
    else if (j == q) eq = eq && ((x >> r) == 0);
  }
  return eq;
}

template <int NT, int W>
__global__ void __launch_bounds__(NT, 1536 / NT) k_local4(const Key<W>* __restrict__ src, BucketTable bt, int rem_bits, EmitCtx<W> ec,
                                                          uint32_t* __restrict__ nd_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Local4Smem sm;
  const int LM = (int)bt.local_max;
  sm.carve(smem_raw, LM, W);
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr int SE = 1 + 2 * W;  // stack entry: nbits, prefix words (low, high halves)
  for (int i = tid; i < SPEC_SMEM; i += NT) sm.spec[i] = 0;
  const uint32_t tag_a = smem_u32(sm.tag), cnt_a = smem_u32(sm.cnt);
  const uint32_t m_cap = (uint32_t)LM + (uint32_t)LM / 4 + 1;
  volatile uint32_t* vmisc = sm.misc;
  for (uint32_t b = bt.b0 + blockIdx.x; b < bt.nb; b += gridDim.x) {
    const unsigned long long n64 = bt.bsize[b];
    if (n64 == 0) {
      if (tid == 0) nd_out[b] = 0;
      continue;
    }
    const unsigned long long o = bt.bofs[b];
    const Key<W>* s = src + o;
    const uint32_t n = (uint32_t)n64;
    const uint32_t m_want = n < (uint32_t)LM ? n + (n >> 2) + 1 : m_cap;
    const uint32_t nrows = (m_want + L4_ROW - 1) / L4_ROW;
    const uint32_t ns = nrows * L4_ROW;
    const uint32_t ns4 = (ns + 3) >> 2;
    const uint32_t nwords = (ns + 31) >> 5;
    const bool big = n > (uint32_t)LM;
    uint32_t run_nd = 0;
    __syncthreads();  // previous bucket fully emitted
    if (tid == 0) {   // the whole bucket: nbits = 0, prefix irrelevant
      sm.stack[0] = 0;
      for (int i = 0; i < 2 * W; i++) sm.stack[1 + i] = 0;
      sm.misc[1] = 1;  // stack size
      sm.misc[4] = 0;  // hash seed
    }
    while (true) {
      __syncthreads();
      const uint32_t sp = sm.misc[1];
      if (sp == 0) break;
      const uint32_t seed = sm.misc[4];
      const int nbits = (int)sm.stack[(sp - 1) * SE];
      Key<W> pref;
#pragma unroll
      for (int i = 0; i < W; i++)
        pref.w[i] = (uint64_t)sm.stack[(sp - 1) * SE + 1 + 2 * i] | ((uint64_t)sm.stack[(sp - 1) * SE + 2 + 2 * i] << 32);
      const int below = rem_bits - nbits;  // key bits below the range prefix
      // ---- clear
      {
        uint4* t4 = reinterpret_cast<uint4*>(sm.tag);
        uint4* c4 = reinterpret_cast<uint4*>(sm.cnt);
        const uint4 e4 = make_uint4(SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY, SLOT_EMPTY), z4 = make_uint4(0, 0, 0, 0);
        for (uint32_t i = tid; i < ns4; i += NT) { t4[i] = e4; c4[i] = z4; }
        for (uint32_t i = tid; i < nwords + 2; i += NT) sm.bitmap[i] = 0;
        if (tid == 0) { sm.misc[0] = 0; sm.misc[3] = 0; }  // row-full flag, tag-collision flag
      }
      __syncthreads();
      if (tid == 0) sm.misc[1] = sp - 1;  // pop
      // tag, row and in-range test of one key
      auto locate = [&](const Key<W>& k, uint32_t& tag, uint32_t& row) -> bool {
        if (nbits > 0 && !key_top_equal<W>(k, pref, below)) return false;
        uint32_t t = (uint32_t)(sm64(key_hash(k) + seed) >> 32);
        tag = t == SLOT_EMPTY ? 0xFFFFFFFEu : t;
        const uint32_t xr = below >= 32 ? key_bits(k, below - 32, 32) : (below > 0 ? key_bits(k, 0, below) << (32 - below) : 0u);
        row = __umulhi(xr, nrows) * L4_ROW;
        return true;
      };
      // ---- pass A: claim slots by tag
      for (uint32_t i0 = 0; i0 < n; i0 += NT) {
        const uint32_t i = i0 + tid;
        if (i < n) {
          const Key<W> k = s[i];
          uint32_t tag, row;
          if (locate(k, tag, row)) {
            uint32_t h = tag >> 28;
            bool placed = false;
#pragma unroll 1
            for (int pr = 0; pr < L4_ROW; pr++, h = (h + 1) & (L4_ROW - 1)) {
              const uint32_t slot = row + h;
              const uint32_t cur = atoms_cas(tag_a + 4 * slot, SLOT_EMPTY, tag);
              if (cur == SLOT_EMPTY) {
                sm.rep[slot] = i;
                atomicOr(&sm.bitmap[slot >> 5], 1u << (slot & 31));
              }
              placed = (cur == SLOT_EMPTY) | (cur == tag);
              if (placed) break;
            }
            if (!placed) sm.misc[0] = 1u;  // row full: split this range
          }
        }
        __syncwarp();
        if (big && __any_sync(0xffffffffu, vmisc[0] != 0)) break;
      }
      __syncthreads();
      const bool failed = sm.misc[0] != 0;
      // ---- pass B: count, verifying the full key against the slot's representative
      if (!failed) {
        for (uint32_t i0 = 0; i0 < n; i0 += NT) {
          const uint32_t i = i0 + tid;
          if (i < n) {
            const Key<W> k = s[i];
            uint32_t tag, row;
            if (locate(k, tag, row)) {
              uint32_t h = tag >> 28, slot = row + h;
#pragma unroll 1
              for (int pr = 0; pr < L4_ROW; pr++, h = (h + 1) & (L4_ROW - 1)) {
                slot = row + h;
                if (lds_u32(tag_a + 4 * slot) == tag) break;
              }
              if (key_eq(k, s[sm.rep[slot]])) reds_add(cnt_a + 4 * slot, 1u);
              else sm.misc[3] = 1u;  // another key with this tag
            }
          }
          __syncwarp();
        }
      }
      __syncthreads();
      if (tid == 0) {
        atomicAdd(&g_l4_dbg[0], 1ull);
        if (failed) atomicAdd(&g_l4_dbg[1], 1ull);
        else if (sm.misc[3] != 0) atomicAdd(&g_l4_dbg[2], 1ull);
      }
      if (failed || sm.misc[3] != 0) {
        if (tid == 0) {
          uint32_t top = sm.misc[1];
          if (!failed) {
            // tag collision: same range again with another seed
            sm.misc[4] = seed + 1;
            top++;  // the entry is still in place above the popped top
          } else if (below > 0 && top + 2 <= (uint32_t)L4_STACK) {
            // split on the next bit: upper half (bit set) is popped second, lower half first
            const int bitpos = below - 1;  // position of the new prefix bit inside the key
            Key<W> up = pref;
            up.w[W - 1 - (bitpos >> 6)] |= 1ull << (bitpos & 63);
            Key<W> lo = pref;
            if (nbits == 0) {  // first split: the prefix starts as this bucket's own leading bits (taken from any key)
              // bits >= rem_bits are equal for every key of the bucket, and key_top_equal ignores nothing above:
              // copy them from the first key, clear everything below
              Key<W> k0 = s[0];
#pragma unroll
              for (int i = 0; i < W; i++) {
                const int j = W - 1 - i;
                const int q = rem_bits >> 6, r = rem_bits & 63;
                uint64_t v = k0.w[i];
                if (j < q) v = 0;
                else if (j == q) v = r ? (v >> r) << r : v;
                lo.w[i] = v;
              }
              up = lo;
              up.w[W - 1 - (bitpos >> 6)] |= 1ull << (bitpos & 63);
            }
            uint32_t* e = sm.stack + top * SE;
            e[0] = (uint32_t)(nbits + 1);
#pragma unroll
            for (int i = 0; i < W; i++) { e[1 + 2 * i] = (uint32_t)up.w[i]; e[2 + 2 * i] = (uint32_t)(up.w[i] >> 32); }
            top++;
            e = sm.stack + top * SE;
            e[0] = (uint32_t)(nbits + 1);
#pragma unroll
            for (int i = 0; i < W; i++) { e[1 + 2 * i] = (uint32_t)lo.w[i]; e[2 + 2 * i] = (uint32_t)(lo.w[i] >> 32); }
            top++;
          } else {
            sm.misc[2] = 1u;  // cannot happen: a single key value always fits; surfaces as a count mismatch
          }
          sm.misc[1] = top;
        }
        continue;
      }
      // ---- dense list of occupied slots, in slot order
      uint32_t nd_total;
      {
        constexpr int WPT = 2;  // bitmap words per thread; nwords <= NT * WPT
        uint32_t wd[WPT], c = 0;
#pragma unroll
        for (int u = 0; u < WPT; u++) {
          const uint32_t wi = tid * WPT + u;
          wd[u] = wi < nwords ? sm.bitmap[wi] : 0u;
          c += __popc(wd[u]);
        }
        uint32_t base = block_excl_scan1<NT>(c, sm.wsum, nd_total);
#pragma unroll
        for (int u = 0; u < WPT; u++) {
          uint32_t w = wd[u];
          const uint32_t s_base = (tid * WPT + u) << 5;
          while (w) {
            const int bit = __ffs((int)w) - 1;
            w &= w - 1;
            sm.list[base++] = (uint16_t)(s_base + bit);
          }
        }
      }
      __syncthreads();
      // ---- one thread per distinct key: rank inside its row by full-key comparison, emit
      for (uint32_t j0 = 0; j0 < nd_total; j0 += NT) {
        const uint32_t j = j0 + tid;
        uint32_t f = 0;
        if (j < nd_total) {
          const uint32_t q = sm.list[j];
          const Key<W> k = s[sm.rep[q]];
          const uint32_t rbase = q & ~(uint32_t)(L4_ROW - 1);
          uint32_t rowbits = (sm.bitmap[q >> 5] >> (rbase & 31)) & 0xFFFFu;
          const uint32_t before = __popc(rowbits & ((1u << (q & (L4_ROW - 1))) - 1u));
          uint32_t smaller = 0;
          rowbits &= ~(1u << (q & (L4_ROW - 1)));
          while (rowbits) {
            const int bit = __ffs((int)rowbits) - 1;
            rowbits &= rowbits - 1;
            smaller += key_less(s[sm.rep[rbase + bit]], k) ? 1u : 0u;
          }
          const unsigned long long pos = o + run_nd + (j - before + smaller);
          f = sm.cnt[q];
          if (ec.want_table) {
            ec.tmp_keys[pos] = k;
            ec.tmp_cnt[pos] = f;
          }
        }
        const uint32_t ones = __ballot_sync(0xffffffffu, f == 1u);
        if (lane == 0 && ones) atomicAdd(&sm.spec[1], (uint32_t)__popc(ones));
        if (f > 1u) {
          if (f < SPEC_SMEM) atomicAdd(&sm.spec[f], 1u);
          else spec_add_global(ec.spec_dense, ec.spec_ovf, ec.spec_ovf_cap, f);
        }
      }
      run_nd += nd_total;
    }
    if (tid == 0) { nd_out[b] = run_nd; atomicAdd(&g_l4_dbg[3], 1ull); }
  }
  __syncthreads();
  for (int i = tid; i < SPEC_SMEM; i += NT) {
    const uint32_t v = sm.spec[i];
    if (v) atomicAdd(&ec.spec_dense[i], (unsigned long long)v);
  }
}

}  // namespace apgk
