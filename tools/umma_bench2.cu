// Follow-up micro-benchmark: which ingredient of the real tap-GEMM slows tcgen05.mma (N=256) from
// 128 to ~210 cycles?  Flags: 1 = rotate the N operand over 36 KB-strided slab slots with +2 KB row
// shifts, 2 = rotate the M operand over 16 KB weight slots, 4 = four warps poll an mbarrier,
// 8 = four warps stream tcgen05.ld from the other accumulator stage, 16 = fence after every 4 MMAs,
// 32 = commit to an mbarrier after every 4 MMAs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../text2human_b200/csrc/t2h_ptx.cuh"
using namespace t2h;

__global__ void __launch_bounds__(256, 1) k(long long* out, int reps, int flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar, never, ring[4], tfull[2], tempty[2];
  __shared__ uint32_t tmem_s;
  __shared__ volatile int done;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x83FF83FFu) | 0x34003400u | ((h >> 4) & 0x08000800u);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&never, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ring[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
    done = 0;
  }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    const uint32_t base = smem_u32(smem);
    const uint32_t wring = base + 3 * 36864;  // weight slots after three slab slots
    const uint32_t idesc = umma_idesc_f16(128, 256);
    long long t0 = clock64();
    int sa = 0, sb = 0, dy = 0, rs = 0;
    if (flags & 64) {
      const int tiles = reps / 18;
      int as = 0, ap = 0;
      for (int t = 0; t < tiles; ++t) {
        mbar_wait(&tempty[as], ap ^ 1);
        tc_fence_after();
        for (int it = 0; it < 18; ++it)
          for (int j = 0; j < 4; ++j)
            umma_f16(tm + as * 256, umma_desc_k128(wring + j * 32), umma_desc_k128(base + j * 32), idesc, (it | j) ? 1 : 0);
        umma_commit(&tfull[as]);
        if (++as == 2) { as = 0; ap ^= 1; }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[blockIdx.x] = (clock64() - t0) * (long long)reps / (tiles * 18);
      done = 1;
    } else {
    for (int r = 0; r < reps; ++r) {
      const uint32_t slab = (flags & 1) ? base + sa * 36864 + dy * 2048 : base;
      const uint32_t w = (flags & 2) ? wring + sb * 16384 : wring;
      if (flags & 16) tc_fence_after();
      for (int j = 0; j < 4; ++j)
        umma_f16(tm + (r & 1) * 0, umma_desc_k128(w + j * 32), umma_desc_k128(slab + j * 32), idesc, 1);
      if (flags & 32) { umma_commit(&ring[rs]); rs = (rs + 1) & 3; }
      if (++dy == 3) { dy = 0; if (++sa == 3) sa = 0; }
      if (++sb == 3) sb = 0;
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
    done = 1;
    }
  } else if (warp >= 4) {
    if (flags & 64) {
      const int tiles = reps / 18;
      int as = 0, ap = 0;
      uint32_t r[32]; uint32_t acc = 0;
      const int nld = (flags & 128) ? 0 : 8;
      for (int t = 0; t < tiles; ++t) {
        mbar_wait(&tfull[as], ap);
        tc_fence_after();
        for (int k = 0; k < nld; ++k) {
          tmem_ld_32x32(tm + (uint32_t((warp & 3) * 32) << 16) + as * 256 + k * 32, r);
          tmem_ld_wait();
          acc += r[0] & 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        if (++as == 2) { as = 0; ap ^= 1; }
      }
      if (acc == 0xffffffffu) out[0] = acc;
    } else if (flags & 4) {
      while (!done) { mbar_try_wait(&never, 0); }
    } else if (flags & 8) {
      uint32_t r[32]; uint32_t acc = 0;
      while (!done) {
        tmem_ld_32x32(tm + (uint32_t((warp & 3) * 32) << 16) + 256 + (acc & 7) * 32, r);
        tmem_ld_wait();
        acc += r[0] & 1;
      }
      if (acc == 0xffffffffu) out[0] = acc;
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int reps = 18 * 300;
  for (int flags : {0, 64, 64 + 128, 8}) {
    k<<<148, 256, 220 * 1024>>>(d, reps, flags);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("flags %d FAILED %s\n", flags, cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("flags %2d: %.1f cyc/MMA\n", flags, (double)mx / (reps * 4.0));
  }
  return 0;
}
