// Micro-benchmarks that size the design choices of the k-mer pipeline on B200.
// Not product code: it measures the primitives DESIGN.md reasons about
// (HBM copy, shared-memory atomics, warp match, global cursor atomics,
// bucketed scatter writes).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -lineinfo -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ---------------------------------------------------------------- copy
__global__ void k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void k_read(const uint4* __restrict__ in, uint32_t* sink, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += stride) acc += in[i].x;
  if (acc == 0x12345678u) *sink = acc;
}

// ---------------------------------------------------------------- smem atomics
// mode 0: RED (no return) ; 1: ATOM with return used ; 2: 64-bit CAS insert + add
template <int MODE>
__global__ void k_atoms(uint32_t* out, int bins_mask, int iters, unsigned long long* cyc) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i <= bins_mask; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t s = (uint32_t)mix64(blockIdx.x * 1024ull + threadIdx.x);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    s = s * 1664525u + 1013904223u;
    uint32_t d = (s >> 12) & bins_mask;
    if (MODE == 0) atomicAdd(&sm[d], 1u);
    else acc += atomicAdd(&sm[d], 1u);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (MODE) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  else if (threadIdx.x <= bins_mask) out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

// baseline: same loop without the atomic (pure LCG + store) to subtract ALU cost
__global__ void k_atoms_base(uint32_t* out, int bins_mask, int iters, unsigned long long* cyc) {
  uint32_t s = (uint32_t)mix64(blockIdx.x * 1024ull + threadIdx.x);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    s = s * 1664525u + 1013904223u;
    acc += (s >> 12) & bins_mask;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---------------------------------------------------------------- warp match
// mode 0: hardware match.any ; 1: ballot loop over BITS
template <int MODE, int BITS>
__global__ void k_match(uint32_t* out, int iters, unsigned long long* cyc) {
  uint64_t s = mix64(blockIdx.x * 1024ull + threadIdx.x);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t d = (uint32_t)(s >> 40) & ((1u << BITS) - 1);
    uint32_t m;
    if (MODE == 0) m = __match_any_sync(0xffffffffu, d);
    else {
      m = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < BITS; b++) {
        uint32_t v = __ballot_sync(0xffffffffu, (d >> b) & 1);
        m &= ((d >> b) & 1) ? v : ~v;
      }
    }
    acc += __popc(m);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---------------------------------------------------------------- global atomics with return
__global__ void k_gatom(uint32_t* ctr, int n_ctr_mask, int iters, uint32_t* out) {
  uint64_t s = mix64(blockIdx.x * 1024ull + threadIdx.x);
  uint32_t acc = 0;
  for (int i = 0; i < iters; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    acc += atomicAdd(&ctr[(uint32_t)(s >> 40) & n_ctr_mask], 1u);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---------------------------------------------------------------- bucketed scatter (unstable MSD partition tile)
// Full prototype of one partition pass: TILE keys -> smem rank via ATOMS ->
// cursor reservation -> staged coalesced write.  keys: 64-bit.
template <int BITS, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
k_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
          unsigned long long* cursors, size_t n, int shift) {
  constexpr int BINS = 1 << BITS;
  constexpr int TILE = THREADS * ITEMS;
  extern __shared__ __align__(16) unsigned char smraw[];
  uint64_t* stage = (uint64_t*)smraw;                   // TILE
  uint32_t* cnt = (uint32_t*)(stage + TILE);            // BINS   (count -> local start)
  unsigned long long* gofs = (unsigned long long*)(cnt + BINS);  // BINS
  __shared__ uint32_t warp_tot[32];
  size_t tile0 = (size_t)blockIdx.x * TILE;
  for (int i = threadIdx.x; i < BINS; i += THREADS) cnt[i] = 0;
  __syncthreads();
  uint64_t key[ITEMS]; uint32_t rk[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    size_t idx = tile0 + (size_t)i * THREADS + threadIdx.x;
    key[i] = idx < n ? in[idx] : ~0ull;
  }
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    size_t idx = tile0 + (size_t)i * THREADS + threadIdx.x;
    uint32_t d = (uint32_t)(key[i] >> shift) & (BINS - 1);
    rk[i] = idx < n ? atomicAdd(&cnt[d], 1u) : 0;
  }
  __syncthreads();
  // exclusive scan of cnt[BINS] with THREADS threads (BINS/THREADS per thread)
  constexpr int PER = (BINS + THREADS - 1) / THREADS;
  uint32_t loc[PER]; uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < PER; j++) { int b = threadIdx.x * PER + j; loc[j] = b < BINS ? cnt[b] : 0; sum += loc[j]; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += v; }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v = threadIdx.x < THREADS / 32 ? warp_tot[threadIdx.x] : 0, iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, iv, o); if (threadIdx.x >= o) iv += u; }
    warp_tot[threadIdx.x] = iv - v;
  }
  __syncthreads();
  uint32_t excl = incl - sum + warp_tot[threadIdx.x >> 5];
#pragma unroll
  for (int j = 0; j < PER; j++) {
    int b = threadIdx.x * PER + j;
    if (b < BINS) {
      cnt[b] = excl;
      if (loc[j]) gofs[b] = atomicAdd(&cursors[b], (unsigned long long)loc[j]) - excl;
      excl += loc[j];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    size_t idx = tile0 + (size_t)i * THREADS + threadIdx.x;
    uint32_t d = (uint32_t)(key[i] >> shift) & (BINS - 1);
    if (idx < n) stage[cnt[d] + rk[i]] = key[i];
  }
  __syncthreads();
  int tile_n = (int)((n - tile0) < (size_t)TILE ? (n - tile0) : TILE);
  for (int j = threadIdx.x; j < tile_n; j += THREADS) {
    uint64_t k = stage[j];
    uint32_t d = (uint32_t)(k >> shift) & (BINS - 1);
    out[gofs[d] + j] = k;
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

template <int BITS, int THREADS, int ITEMS>
void run_scatter(const uint64_t* in, uint64_t* out, size_t n, unsigned long long* cursors, unsigned long long* h_base) {
  constexpr int BINS = 1 << BITS, TILE = THREADS * ITEMS;
  size_t smem = (size_t)TILE * 8 + BINS * 4 + BINS * 8;
  CK(cudaFuncSetAttribute(k_scatter<BITS, THREADS, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int grid = (int)((n + TILE - 1) / TILE);
  float best = 1e9;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemcpy(cursors, h_base, BINS * 8, cudaMemcpyHostToDevice));
    CK(cudaEventRecord(e0));
    k_scatter<BITS, THREADS, ITEMS><<<grid, THREADS, smem>>>(in, out, cursors, n, 50 - BITS);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms = time_ms(e0, e1); if (ms < best) best = ms;
  }
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scatter<BITS, THREADS, ITEMS>, THREADS, smem));
  printf("scatter bits=%2d threads=%4d items=%2d tile=%6d smem=%6zu occ=%d : %.3f ms  %.1f Gkeys/s  %.1f GB/s(r+w)\n",
         BITS, THREADS, ITEMS, TILE, smem, occ, best, n / best * 1e-6, n * 16.0 / best * 1e-6);
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("device %s sms=%d clock=%d kHz smem/blk optin=%zu l2=%d MB\n", p.name, p.multiProcessorCount, clk,
         p.sharedMemPerBlockOptin, p.l2CacheSize >> 20);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int SMS = p.multiProcessorCount;

  // ---- copy / read bandwidth
  {
    size_t bytes = 4ull << 30; size_t n = bytes / 16;
    uint4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    uint32_t* sink; CK(cudaMalloc(&sink, 4));
    for (int bpsm : {4, 8, 16}) for (int th : {256, 512}) {
      float best = 1e9, bestr = 1e9;
      for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0)); k_copy<<<SMS * bpsm, th>>>(a, b, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
        CK(cudaEventRecord(e0)); k_read<<<SMS * bpsm, th>>>(a, sink, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        ms = time_ms(e0, e1); if (ms < bestr) bestr = ms;
      }
      printf("copy  blocks/sm=%2d threads=%d : %.3f ms %.1f GB/s (r+w) | read %.3f ms %.1f GB/s\n", bpsm, th, best,
             2.0 * bytes / best * 1e-6, bestr, bytes / bestr * 1e-6);
    }
    CK(cudaGetLastError());
    CK(cudaFree(a)); CK(cudaFree(b)); CK(cudaFree(sink));
  }

  // ---- smem atomics
  {
    uint32_t* out; CK(cudaMalloc(&out, SMS * 8 * 1024 * 4));
    unsigned long long* cyc; CK(cudaMallocManaged(&cyc, SMS * 8 * 8));
    const int iters = 4096;
    for (int th : {256, 1024}) for (int bins : {256, 1024, 4096, 8192}) {
      double r[3];
      for (int mode = 0; mode < 3; mode++) {
        int blocks = SMS;  // one block per SM so cycles/SM is clean
        if (mode == 0) k_atoms<0><<<blocks, th, bins * 4>>>(out, bins - 1, iters, cyc);
        else if (mode == 1) k_atoms<1><<<blocks, th, bins * 4>>>(out, bins - 1, iters, cyc);
        else k_atoms_base<<<blocks, th>>>(out, bins - 1, iters, cyc);
        CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
        double mean = 0; for (int i = 0; i < blocks; i++) mean += cyc[i]; mean /= blocks;
        r[mode] = mean / ((double)iters * th);  // cycles per lane-op per SM
      }
      printf("atoms threads=%4d bins=%5d : RED %.3f cyc/lane  ATOM(ret) %.3f cyc/lane  (alu-only loop %.3f)\n", th, bins, r[0], r[1], r[2]);
    }
    // ---- match
    for (int th : {256, 1024}) {
      double r[6]; int k = 0;
      int blocks = SMS;
#define RUNM(MODE, BITS) k_match<MODE, BITS><<<blocks, th>>>(out, iters, cyc); CK(cudaDeviceSynchronize()); CK(cudaGetLastError()); \
      { double mean = 0; for (int i = 0; i < blocks; i++) mean += cyc[i]; mean /= blocks; r[k++] = mean / ((double)iters * th / 32); }
      RUNM(0, 4) RUNM(0, 8) RUNM(0, 10) RUNM(1, 4) RUNM(1, 8) RUNM(1, 10)
      printf("match threads=%4d : hw match.any 4b %.1f  8b %.1f  10b %.1f cyc/warp-instr/SM | ballot-loop 4b %.1f 8b %.1f 10b %.1f\n",
             th, r[0], r[1], r[2], r[3], r[4], r[5]);
    }
    CK(cudaFree(out)); CK(cudaFree(cyc));
  }

  // ---- global atomics with return
  {
    uint32_t *ctr, *out; CK(cudaMalloc(&ctr, 4 << 20)); CK(cudaMemset(ctr, 0, 4 << 20)); CK(cudaMalloc(&out, SMS * 8 * 256 * 4));
    for (int mask : {1023, (1 << 20) - 1}) {
      int iters = 2048;
      k_gatom<<<SMS * 8, 256>>>(ctr, mask, iters, out);
      CK(cudaEventRecord(e0)); k_gatom<<<SMS * 8, 256>>>(ctr, mask, iters, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms = time_ms(e0, e1);
      printf("global atomicAdd(ret) over %d counters: %.2f Gatom/s\n", mask + 1, (double)SMS * 8 * 256 * iters / ms * 1e-6);
    }
    CK(cudaFree(ctr)); CK(cudaFree(out));
  }

  // ---- scatter prototype: 1 G keys of 50 random bits
  {
    size_t n = (argc > 1) ? strtoull(argv[1], 0, 10) : (1ull << 29);
    uint64_t *in, *out; CK(cudaMalloc(&in, n * 8)); CK(cudaMalloc(&out, n * 8));
    uint64_t* h = (uint64_t*)malloc(n * 8);
    for (size_t i = 0; i < n; i++) h[i] = mix64(i) >> 14;
    CK(cudaMemcpy(in, h, n * 8, cudaMemcpyHostToDevice));
    unsigned long long* cursors; CK(cudaMalloc(&cursors, 4096 * 8));
    for (int bits : {8, 10, 11, 12}) {
      int bins = 1 << bits;
      unsigned long long* hb = (unsigned long long*)calloc(bins + 1, 8);
      for (size_t i = 0; i < n; i++) hb[(h[i] >> (50 - bits)) + 1]++;
      for (int b = 0; b < bins; b++) hb[b + 1] += hb[b];
      if (bits == 8) { run_scatter<8, 256, 16>(in, out, n, cursors, hb); run_scatter<8, 512, 16>(in, out, n, cursors, hb); }
      if (bits == 10) { run_scatter<10, 256, 16>(in, out, n, cursors, hb); run_scatter<10, 512, 16>(in, out, n, cursors, hb);
                        run_scatter<10, 512, 32>(in, out, n, cursors, hb); run_scatter<10, 1024, 16>(in, out, n, cursors, hb);
                        run_scatter<10, 256, 32>(in, out, n, cursors, hb); }
      if (bits == 11) { run_scatter<11, 512, 16>(in, out, n, cursors, hb); run_scatter<11, 1024, 16>(in, out, n, cursors, hb); }
      if (bits == 12) { run_scatter<12, 512, 32>(in, out, n, cursors, hb); run_scatter<12, 1024, 16>(in, out, n, cursors, hb); }
      // verify last run of this bits: every output key lies in its bucket range
      uint64_t* ho = (uint64_t*)malloc(n * 8);
      CK(cudaMemcpy(ho, out, n * 8, cudaMemcpyDeviceToHost));
      size_t bad = 0; uint64_t xin = 0, xout = 0;
      for (int b = 0; b < bins; b++) for (size_t i = hb[b]; i < hb[b + 1]; i++) if ((ho[i] >> (50 - bits)) != (uint64_t)b) bad++;
      for (size_t i = 0; i < n; i++) { xin += mix64(h[i]); xout += mix64(ho[i]); }
      printf("  verify bits=%d: misplaced=%zu checksum %s\n", bits, bad, xin == xout ? "ok" : "MISMATCH");
      free(ho); free(hb);
    }
    free(h); CK(cudaFree(in)); CK(cudaFree(out)); CK(cudaFree(cursors));
  }
  return 0;
}
