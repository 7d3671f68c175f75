// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128 / 2-CTA M=256) on B200, operands
// resident in shared memory (no TMA in the loop).  Answers: how many SM cycles does one
// 128xNx16 (1-CTA) or 256xNx16 (2-CTA pair) MMA occupy, i.e. what fraction of the nominal
// 4096 MAC/clk/SM a single-CTA tcgen05 kernel can reach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../text2human_b200/csrc/t2h_ptx.cuh"

using namespace t2h;

__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// out[blockIdx] = cycles for `reps` batches of (4 * nacc) MMAs
template <int N>
__global__ void __launch_bounds__(128, 1) bench_1cta(long long* out, int reps, int nacc, int kst, int rnd) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < (16384 * 2 + N * 128) / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // two fp16 values in [-2,2): sign|exp in {13..15}|mantissa random
    uint32_t v = (h & 0x83FF83FFu) | 0x34003400u | ((h >> 4) & 0x08000800u);
    reinterpret_cast<uint32_t*>(smem)[i] = rnd ? v : 0u;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_s, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_s;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 32768;
    const uint32_t idesc = umma_idesc_f16(128, N);
    uint32_t par = 0;
    // warm-up
    for (int j = 0; j < 4; ++j) umma_f16(tm, umma_desc_k128(a + j * 32), umma_desc_k128(b + j * 32), idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, par);
    par ^= 1;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int acc = 0; acc < nacc; ++acc)
        for (int j = 0; j < kst; ++j)
          umma_f16(tm + acc * N, umma_desc_k128(a + acc * 16384 + (j & 3) * 32), umma_desc_k128(b + (j & 3) * 32), idesc, j > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, par);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
bench_2cta(long long* out, int reps, int nacc, int kst, int rnd) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < (16384 * 2 + N * 64) / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    uint32_t v = (h & 0x83FF83FFu) | 0x34003400u | ((h >> 4) & 0x08000800u);
    reinterpret_cast<uint32_t*>(smem)[i] = rnd ? v : 0u;
  }
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_s;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 32768;
    // M = 256 across the pair, each CTA holds 128 rows of A and N/2 rows of B
    const uint32_t idesc = umma_idesc_f16(256, N);
    uint32_t par = 0;
    for (int j = 0; j < 4; ++j) umma_f16_2cta(tm, umma_desc_k128(a + j * 32), umma_desc_k128(b + j * 32), idesc, 1);
    umma_commit_2cta(&bar);
    mbar_wait(&bar, par);
    par ^= 1;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int acc = 0; acc < nacc; ++acc)
        for (int j = 0; j < kst; ++j)
          umma_f16_2cta(tm + acc * N, umma_desc_k128(a + acc * 16384 + (j & 3) * 32), umma_desc_k128(b + (j & 3) * 32), idesc, j > 0);
    }
    umma_commit_2cta(&bar);
    mbar_wait(&bar, par);
    long long t1 = clock64();
    out[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

template <typename K>
static void run(const char* name, K kern, int grid, int n, int m_total, int reps, int nacc, int kst, bool pair,
                int rnd) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  cudaMemset(d, 0, sizeof(long long) * grid);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<grid, 128, 200 * 1024>>>(d, reps / 10, nacc, kst, rnd);  // warm
  cudaEventRecord(e0);
  kern<<<grid, 128, 200 * 1024>>>(d, reps, nacc, kst, rnd);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s FAILED: %s\n", name, cudaGetErrorString(e));
    cudaFree(d);
    return;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[512];
  int cnt = pair ? grid / 2 : grid;
  cudaMemcpy(h, d, sizeof(long long) * cnt, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < cnt; ++i) mx = h[i] > mx ? h[i] : mx;
  const double mmas = (double)reps * nacc * kst;
  const double cyc = (double)mx / mmas;
  const double macs = (double)m_total * n * 16;
  const double tflops = 2.0 * macs * mmas * cnt / (ms * 1e-3) / 1e12;
  printf("%-22s %s grid=%3d %7.1f cyc/MMA %7.1f MAC/clk/SM  %8.3f ms  %7.1f TFLOP/s  (eff clk %.2f GHz)\n", name,
         rnd ? "rand" : "zero", grid, cyc, macs / cyc / (pair ? 2 : 1), ms, tflops, mx / (ms * 1e-3) / 1e9);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int reps = 20000;
  for (int rnd : {0, 1}) {
    const int grid = 148;
    run("1cta M128 N128 2acc", bench_1cta<128>, grid, 128, 128, reps, 2, 4, false, rnd);
    run("1cta M128 N256 1acc", bench_1cta<256>, grid, 256, 128, reps, 1, 4, false, rnd);
    run("1cta M128 N256 2acc", bench_1cta<256>, grid, 256, 128, reps, 2, 4, false, rnd);
    run("2cta M256 N128 2acc", bench_2cta<128>, grid, 128, 256, reps, 2, 4, true, rnd);
    run("2cta M256 N256 1acc", bench_2cta<256>, grid, 256, 256, reps, 1, 4, true, rnd);
    run("2cta M256 N256 2acc", bench_2cta<256>, grid, 256, 256, reps, 2, 4, true, rnd);
  }
  return 0;
}
