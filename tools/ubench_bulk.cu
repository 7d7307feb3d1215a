// Micro-benchmark: how fast can one SM push many SMALL runs from shared memory to scattered global
// destinations -- the write-out phase of the partition passes (partition.cuh) -- with
//   (a) per-element LDS -> STG through registers (what the scatter kernels did in round 1), and
//   (b) one cp.async.bulk.global.shared::cta (TMA 1-D bulk copy, SASS UBLKCP) per run, issued by many threads.
// Not product code.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench_bulk tools/ubench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Each CTA owns `bins` output streams of `stream_bytes` each inside out + blockIdx.x * bins * stream_bytes.
// Per tile every bin receives one run of run_bytes (tile = bins * run_bytes <= stage).
template <int MODE>   // 0: LDS/STG 8-byte elements, 1: one bulk copy per run (thread d -> bin d), 2: bulk, runs issued by lane 0 of each warp in a loop
__global__ void __launch_bounds__(1024, 1) k_runs(unsigned char* __restrict__ out, int bins, uint32_t run_bytes, int tiles,
                                                  size_t stream_bytes) {
  extern __shared__ __align__(128) unsigned char stage[];
  const uint32_t tile_bytes = (uint32_t)bins * run_bytes;
  for (uint32_t i = threadIdx.x; i < tile_bytes / 8; i += blockDim.x) ((uint64_t*)stage)[i] = i * 0x9E3779B97F4A7C15ull;
  __syncthreads();
  unsigned char* base = out + (size_t)blockIdx.x * bins * stream_bytes;
  const uint32_t stage_a = smem_u32(stage);
  for (int t = 0; t < tiles; t++) {
    const size_t toff = (size_t)t * run_bytes;
    if (MODE == 0) {
      const uint32_t per_run = run_bytes / 8;   // 8-byte elements per run
      for (uint32_t i = threadIdx.x; i < tile_bytes / 8; i += blockDim.x) {
        const uint32_t d = i / per_run, j = i % per_run;
        const uint64_t v = ((const uint64_t*)stage)[i];
        *(uint64_t*)(base + (size_t)d * stream_bytes + toff + (size_t)j * 8) = v;
      }
      __syncthreads();
    } else if (MODE == 1) {
      fence_async_smem();
      __syncthreads();
      for (int d = threadIdx.x; d < bins; d += blockDim.x)
        bulk_s2g(base + (size_t)d * stream_bytes + toff, stage_a + (uint32_t)d * run_bytes, run_bytes);
      bulk_commit();
      bulk_wait_read0();
      __syncthreads();
    } else {
      fence_async_smem();
      __syncthreads();
      if ((threadIdx.x & 31) == 0) {
        for (int d = threadIdx.x >> 5; d < bins; d += blockDim.x >> 5)
          bulk_s2g(base + (size_t)d * stream_bytes + toff, stage_a + (uint32_t)d * run_bytes, run_bytes);
        bulk_commit();
        bulk_wait_read0();
      }
      __syncthreads();
    }
  }
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int SMS = p.multiProcessorCount;
  printf("device %s sms=%d\n", p.name, SMS);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int bins = 1024;
  const int tiles = 256;
  for (uint32_t run_bytes : {16u, 32u, 64u, 128u}) {
    const size_t stream_bytes = (size_t)tiles * run_bytes;
    const size_t total = (size_t)SMS * bins * stream_bytes;
    unsigned char* out; CK(cudaMalloc(&out, total));
    const size_t smem = (size_t)bins * run_bytes;
    auto run = [&](int mode) {
      float best = 1e9;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(out, 0, total));
        CK(cudaEventRecord(e0));
        if (mode == 0) { CK(cudaFuncSetAttribute(k_runs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_runs<0><<<SMS, 1024, smem>>>(out, bins, run_bytes, tiles, stream_bytes); }
        if (mode == 1) { CK(cudaFuncSetAttribute(k_runs<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_runs<1><<<SMS, 1024, smem>>>(out, bins, run_bytes, tiles, stream_bytes); }
        if (mode == 2) { CK(cudaFuncSetAttribute(k_runs<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_runs<2><<<SMS, 1024, smem>>>(out, bins, run_bytes, tiles, stream_bytes); }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
      }
      return best;
    };
    const float m0 = run(0), m1 = run(1), m2 = run(2);
    // verify mode 1 wrote what mode 0 writes: compare a checksum of the first stream region
    const double gb = (double)total * 1e-9;
    printf("run=%4u B x %d bins x %d tiles/SM: LDS/STG %.3f ms %.0f GB/s | bulk/thread %.3f ms %.0f GB/s (%.1f cyc/copy/SM) | bulk/warp-leader %.3f ms %.0f GB/s\n",
           run_bytes, bins, tiles, m0, gb / m0 * 1e3, m1, gb / m1 * 1e3, (double)m1 * 1e-3 * 1.965e9 / ((double)bins * tiles), m2, gb / m2 * 1e3);
    CK(cudaFree(out));
  }
  // correctness spot check of the bulk path: distinct pattern per byte offset
  {
    const uint32_t run_bytes = 64; const size_t stream_bytes = (size_t)tiles * run_bytes;
    const size_t total = (size_t)bins * stream_bytes;
    unsigned char *o0, *o1; CK(cudaMalloc(&o0, total)); CK(cudaMalloc(&o1, total));
    CK(cudaFuncSetAttribute(k_runs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bins * run_bytes)));
    CK(cudaFuncSetAttribute(k_runs<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bins * run_bytes)));
    k_runs<0><<<1, 1024, bins * run_bytes>>>(o0, bins, run_bytes, tiles, stream_bytes);
    k_runs<1><<<1, 1024, bins * run_bytes>>>(o1, bins, run_bytes, tiles, stream_bytes);
    CK(cudaDeviceSynchronize());
    unsigned char* h0 = (unsigned char*)malloc(total); unsigned char* h1 = (unsigned char*)malloc(total);
    CK(cudaMemcpy(h0, o0, total, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h1, o1, total, cudaMemcpyDeviceToHost));
    size_t bad = 0; for (size_t i = 0; i < total; i++) bad += h0[i] != h1[i];
    printf("bulk vs LDS/STG output: %zu differing bytes of %zu\n", bad, total);
  }
  return 0;
}
