// Probe: shared-memory descriptor encoding for MN-major (transposed) tcgen05.mma operands with the 128-byte
// swizzle, as TMA would deposit a [K rows][64 MN elements] box.  Tries (LBO, SBO, K-step) candidates for an
// MN-major A (M=128, two 64-wide boxes) against a known-good K-major B, then an MN-major B (N=256, four
// boxes) against a K-major A, with small-integer data (exact in fp16/fp32) and prints which candidates
// reproduce D = A.B^T.  build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn_probe umma_mn_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../text2human_b200/csrc/t2h_ptx.cuh"
using namespace t2h;

constexpr int K = 64;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// mode 0: A MN-major [K][128] (2 boxes), B K-major [64][K]; D 128x64
// mode 1: A K-major [128][K], B MN-major [K][256] (4 boxes); D 128x256
__global__ void __launch_bounds__(128, 1) probe(const __half* Ag, const __half* Bg, float* D, int mode, uint32_t lbo,
                                                uint32_t sbo, uint32_t kstep, int major_bit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 16 KB
  uint8_t* sB = smem + 16384;    // up to 32 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int M = 128, N = mode == 0 ? 64 : 256;
  // emulate TMA SWIZZLE_128B deposits
  if (mode == 0) {
    for (int i = threadIdx.x; i < K * M; i += blockDim.x) {  // Ag[k][m]
      const int k = i / M, m = i % M, box = m / 64, mm = m % 64, j = mm / 8;
      *reinterpret_cast<__half*>(sA + box * (K * 128) + k * 128 + ((j ^ (k & 7)) << 4) + (mm % 8) * 2) = Ag[i];
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {  // Bg[n][k]
      const int n = i / K, k = i % K, j = k / 8;
      *reinterpret_cast<__half*>(sB + n * 128 + ((j ^ (n & 7)) << 4) + (k % 8) * 2) = Bg[i];
    }
  } else {
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) {  // Ag[m][k]
      const int m = i / K, k = i % K, j = k / 8;
      *reinterpret_cast<__half*>(sA + m * 128 + ((j ^ (m & 7)) << 4) + (k % 8) * 2) = Ag[i];
    }
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {  // Bg[k][n]
      const int k = i / N, n = i % N, box = n / 64, nn = n % 64, j = nn / 8;
      *reinterpret_cast<__half*>(sB + box * (K * 128) + k * 128 + ((j ^ (k & 7)) << 4) + (nn % 8) * 2) = Bg[i];
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_s, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_s;
  if (threadIdx.x == 0) {
    uint32_t idesc = umma_idesc_f16(M, N) | (1u << major_bit);
    for (int s = 0; s < 4; ++s) {
      uint64_t da, db;
      if (mode == 0) {
        da = make_desc(smem_u32(sA) + s * kstep, lbo, sbo);
        db = umma_desc_k128(smem_u32(sB) + s * 32);
      } else {
        da = umma_desc_k128(smem_u32(sA) + s * 32);
        db = make_desc(smem_u32(sB) + s * kstep, lbo, sbo);
      }
      umma_f16(tm, da, db, idesc, s ? 1 : 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tm + (uint32_t(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * N + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

int main() {
  const int M = 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    const int N = mode == 0 ? 64 : 256;
    const int na = M * K, nb = N * K;
    __half *hA = (__half*)malloc(na * 2), *hB = (__half*)malloc(nb * 2);
    float* ref = (float*)calloc(M * N, 4);
    float* fa = (float*)malloc(na * 4);
    float* fb = (float*)malloc(nb * 4);
    srand(7 + mode);
    for (int i = 0; i < na; ++i) { fa[i] = (float)(rand() % 7 - 3); hA[i] = __float2half(fa[i]); }
    for (int i = 0; i < nb; ++i) { fb[i] = (float)(rand() % 5 - 2); hB[i] = __float2half(fb[i]); }
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float s = 0;
        for (int k = 0; k < K; ++k) {
          const float a = mode == 0 ? fa[k * M + m] : fa[m * K + k];
          const float b = mode == 0 ? fb[n * K + k] : fb[k * N + n];
          s += a * b;
        }
        ref[m * N + n] = s;
      }
    __half *dA, *dB;
    float* dD;
    cudaMalloc(&dA, na * 2); cudaMalloc(&dB, nb * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, na * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, nb * 2, cudaMemcpyHostToDevice);
    float* out = (float*)malloc(M * N * 4);
    const uint32_t cand[] = {16, 128, 1024, 2048, 4096, 8192, 0};
    const uint32_t ksteps[] = {2048, 32, 256, 1024};
    const int major_bit = mode == 0 ? 15 : 16;
    int found = 0;
    for (uint32_t lbo : cand)
      for (uint32_t sbo : cand)
        for (uint32_t ks : ksteps) {
          cudaMemset(dD, 0, M * N * 4);
          probe<<<1, 128, 64 * 1024>>>(dA, dB, dD, mode, lbo, sbo, ks, major_bit);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("mode %d lbo %u sbo %u kstep %u: CUDA error %s\n", mode, lbo, sbo, ks, cudaGetErrorString(e));
            return 1;
          }
          cudaMemcpy(out, dD, M * N * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int i = 0; i < M * N; ++i) bad += (out[i] != ref[i]);
          if (bad == 0) {
            printf("MATCH mode %d (%s MN-major): lbo %u sbo %u kstep %u major_bit %d\n", mode, mode == 0 ? "A" : "B", lbo,
                   sbo, ks, major_bit);
            ++found;
          } else if (bad < M * N / 2) {
            printf("partial mode %d: lbo %u sbo %u kstep %u -> %d / %d wrong\n", mode, lbo, sbo, ks, bad, M * N);
          }
        }
    printf("mode %d: %d matching candidates\n", mode, found);
  }
  return 0;
}
