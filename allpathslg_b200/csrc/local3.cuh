// local3.cuh -- sort-and-count for 32-bit remainders by ORDER-PRESERVING hashing.
//
// Keys of one bucket share their leading bits, so the remainder r (< 2^REM) of a
// random-looking genome is spread evenly.  The home slot of a key is therefore
// chosen MONOTONE in the key:   home(r) = floor(r * M / 2^REM),  M ~ 1.25 n.
// Insertion is plain linear probing without wrap-around (one CAS + one ADD per
// key instance groups identical keys), and because the hash is monotone every
// maximal run of occupied slots ("cluster") holds exactly the keys whose homes
// fall inside it: clusters are already in ascending key order, and only the
// handful of entries inside a cluster have to be ranked against each other.
//
//   1. clear the table                                   [block]
//   2. insert the bucket's keys straight from HBM        [block]   1 LDS + (CAS) + 1 RED per key
//   3. per-thread occupied-slot counts, one block scan   [block]
//   4. every thread walks its slots: rank inside the cluster by comparison,
//      write (k-mer, count) at its final ascending position, update the spectrum
//
// Against k_local2 (split + warp tables + bin sort, 8.9 warp-instructions per key
// measured) this needs ~1: there is no separate sort at all.
//
// A bucket is handed to the general kernel (deferred list) when probing leaves the
// table or a cluster grows beyond L3_MAXWALK (low-complexity sequence).
#pragma once
#include "local2.cuh"

namespace apgk {

constexpr int L3_SLACK = 96;     // slots past the last home slot (no wrap-around)
constexpr int L3_MAXWALK = 64;   // longest cluster walk before the bucket is deferred

// Needs 1 <= REM <= 31 and LM <= 32767 (position and count share a word).
// shared memory: key[NS] u32 | cnt[NS] u32 | spec[SPEC_SMEM] | wsum[33] | misc[8]      NS = LM + LM/4 + SLACK
struct Local3Smem {
  uint32_t* key; uint32_t* cnt; uint32_t* spec; uint32_t* wsum; uint32_t* misc;
  static __host__ __device__ size_t slots(int LM) { return (size_t)LM + LM / 4 + L3_SLACK + 32; }
  __device__ __forceinline__ void carve(unsigned char* raw, int LM) {
    const size_t ns = slots(LM);
    key = (uint32_t*)raw;
    cnt = key + ns;
    spec = cnt + ns;
    wsum = spec + SPEC_SMEM;
    misc = wsum + 40;
  }
  static size_t bytes(int LM) { return (slots(LM) * 2 + SPEC_SMEM + 40 + 8) * 4; }
};

template <int NT, int W>
__global__ void __launch_bounds__(NT) k_local3(const uint32_t* src, BucketTable bt, int rem_bits, EmitCtx<W> ec,
                                               uint32_t* __restrict__ nd_out, uint32_t* __restrict__ deferred,
                                               uint32_t deferred_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Local3Smem sm;
  sm.carve(smem_raw, (int)bt.local_max);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NWARP = NT / 32;
  for (int i = tid; i < SPEC_SMEM; i += NT) sm.spec[i] = 0;
  volatile uint32_t* vkey = sm.key;
  const int up = 32 - rem_bits;  // remainder left-aligned in 32 bits
  for (uint32_t b = blockIdx.x; b < bt.nb; b += gridDim.x) {
    const unsigned long long n64 = bt.bsize[b];
    if (n64 == 0) {
      if (tid == 0) nd_out[b] = 0;
      continue;
    }
    if (n64 > bt.local_max) continue;  // k_big's job
    const uint32_t n = (uint32_t)n64;
    const unsigned long long o = bt.bofs[b];
    const uint32_t* s = src + o;
    const uint32_t m_home = n + (n >> 2) + 1;          // homes lie in [0, m_home)
    const uint32_t ns = m_home + L3_SLACK;             // probing may run into the slack
    // ---- 1. clear
    __syncthreads();  // previous bucket fully emitted
    for (uint32_t i = tid; i < ns; i += NT) { sm.key[i] = SLOT_EMPTY; sm.cnt[i] = 0; }
    if (tid == 0) sm.misc[0] = 0;
    __syncthreads();
    // ---- 2. insert
    for (uint32_t i = tid; i < n; i += NT) {
      const uint32_t k = s[i];
      uint32_t slot = __umulhi(k << up, m_home);
      while (true) {
        uint32_t cur = vkey[slot];
        if (cur == SLOT_EMPTY) cur = atomicCAS(&sm.key[slot], SLOT_EMPTY, k);
        if (cur == SLOT_EMPTY || cur == k) { atomicAdd(&sm.cnt[slot], 1u); break; }
        if (++slot >= ns) { sm.misc[0] = 1u; break; }
      }
    }
    __syncthreads();
    // ---- 3. occupied slots before each thread's chunk
    const uint32_t per = (ns + NT - 1) / NT;
    const uint32_t s0 = tid * per, s1 = min(s0 + per, ns);
    uint32_t occ = 0;
    for (uint32_t q = s0; q < s1; q++) occ += (sm.key[q] != SLOT_EMPTY);
    const uint32_t incl = warp_incl_scan(occ, lane);
    if (lane == 31) sm.wsum[wid] = incl;
    __syncthreads();
    uint32_t before = incl - occ;
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < NWARP; w++) {
      const uint32_t v = sm.wsum[w];
      if (w < wid) before += v;
      total += v;
    }
    // ---- 4. rank inside clusters; the final position is parked next to the count
    //         (count < 2^16 | position << 16 | 1 << 31) until the whole bucket is known to be good
    bool bad = false;
    if (sm.misc[0] == 0) {
      for (uint32_t q = s0; q < s1; q++) {
        const uint32_t k = sm.key[q];
        if (k == SLOT_EMPTY) continue;
        uint32_t left = 0, smaller = 0;
        for (int l = (int)q - 1; l >= 0; l--) {
          const uint32_t kl = sm.key[l];
          if (kl == SLOT_EMPTY) break;
          left++;
          smaller += kl < k;
          if (left > L3_MAXWALK) { bad = true; break; }
        }
        uint32_t right = 0;
        for (uint32_t r = q + 1; r < ns; r++) {
          const uint32_t kr = sm.key[r];
          if (kr == SLOT_EMPTY) break;
          smaller += kr < k;
          if (++right > L3_MAXWALK) { bad = true; break; }
        }
        const uint32_t pos = before - left + smaller;
        sm.cnt[q] = sm.cnt[q] | (pos << 16) | 0x80000000u;
        before++;
      }
    }
    if (bad) sm.misc[0] = 1u;
    __syncthreads();
    if (sm.misc[0] != 0) {  // hand the bucket to the general kernel; nothing has been written or counted
      if (tid == 0) {
        const uint32_t i = atomicAdd(&deferred[0], 1u);
        if (i < deferred_cap) deferred[1 + i] = b;
      }
      continue;
    }
    // ---- 5. emit in ascending key order (counts reuse the bucket's own, fully consumed, input range)
    uint32_t* cnt_dst = const_cast<uint32_t*>(src) + o;
    for (uint32_t q = s0; q < s1; q++) {
      const uint32_t v = sm.cnt[q];
      if (v & 0x80000000u) {
        const uint32_t f = v & 0xFFFFu, pos = (v >> 16) & 0x7FFFu;
        if (ec.want_table) {
          ec.tmp_keys[o + pos] = rebuild_key<W>(sm.key[q], (uint64_t)b, ec.rem_bits, ec.pad);
          cnt_dst[pos] = f;
        }
        if (f < SPEC_SMEM) atomicAdd(&sm.spec[f], 1u);
        else spec_add_global(ec.spec_dense, ec.spec_ovf, ec.spec_ovf_cap, f);
      }
    }
    if (tid == 0) nd_out[b] = total;
  }
  __syncthreads();
  for (int i = tid; i < SPEC_SMEM; i += NT) {
    const uint32_t v = sm.spec[i];
    if (v) atomicAdd(&ec.spec_dense[i], (unsigned long long)v);
  }
}

}  // namespace apgk
