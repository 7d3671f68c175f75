// What does one node of a dependent kernel chain cost on this machine when the kernels do (almost) nothing?
// A CUDA graph of N kernels in a stream-ordered chain, replayed; kernel "big" = the tap-GEMM's launch shape
// (256 threads, ~223 KB dynamic shared memory, optionally the TMEM allocation + barrier set-up of its prologue),
// kernel "small" = an elementwise pass's shape (512 x 128 threads, no shared memory).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/launch_floor tools/launch_floor.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../text2human_b200/csrc/t2h_ptx.cuh"
using namespace t2h;

struct Pad {
  char bytes[1400];  // the tap-GEMM passes ~1.3 KB of parameters (4 tensor maps + its geometry struct)
};

template <bool PROLOGUE>
__global__ void __launch_bounds__(256, 1) big(float* out, const __grid_constant__ Pad pad) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[32];
  __shared__ uint32_t tmem_s;
  if (PROLOGUE) {
    if (threadIdx.x == 32) {
      for (int i = 0; i < 32; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
    }
    if (threadIdx.x >= 64 && threadIdx.x < 96) {
      tmem_alloc(&tmem_s, 512);
      tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  pdl_wait();
  if (threadIdx.x == 0) out[blockIdx.x] = smem_raw[0] + pad.bytes[5];
  if (PROLOGUE) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x >= 64 && threadIdx.x < 96) {
      tc_fence_after();
      tmem_dealloc(tmem_s, 512);
    }
  }
}

__global__ void small(float* out) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) out[blockIdx.x] += 1.f;
}

static cudaError_t launch(void (*k)(float*, Pad), int grid, int smem, bool pdl, cudaStream_t st, float* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  Pad pad = {};
  return cudaLaunchKernelEx(&cfg, k, out, pad);
}
static cudaError_t launch_small(int grid, bool pdl, cudaStream_t st, float* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, small, out);
}

int main() {
  const int kSmem = 227 * 1024 - 4096;
  cudaFuncSetAttribute(big<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(big<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  float* out;
  cudaMalloc(&out, 4096 * 4);
  cudaMemset(out, 0, 4096 * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  const int N = 168;
  struct Case {
    const char* name;
    int pattern;  // 0: all big, 1: big,big,big,small repeating, 2: all small, 3: big with small smem
    bool prologue, pdl, carve;
    int grid;
  } cases[] = {
      {"big (223 KB smem), no prologue, PDL, grid 128", 0, false, true, false, 128},
      {"big (223 KB smem), no prologue, no PDL, grid 128", 0, false, false, false, 128},
      {"big + TMEM alloc/barrier prologue, PDL, grid 128", 0, true, true, false, 128},
      {"big + prologue, no PDL, grid 128", 0, true, false, false, 128},
      {"big + prologue, PDL, grid 148", 0, true, true, false, 148},
      {"big + prologue, PDL, grid 32", 0, true, true, false, 32},
      {"3 big : 1 small, PDL", 1, true, true, false, 128},
      {"3 big : 1 small, PDL, small prefers max shared carve-out", 1, true, true, true, 128},
      {"3 big : 1 small, no PDL", 1, true, false, false, 128},
      {"all small (512 x 128), PDL", 2, false, true, false, 128},
      {"all small, no PDL", 2, false, false, false, 128},
      {"256 thr, 8 KB smem, no prologue, PDL, grid 128", 3, false, true, false, 128},
  };
  for (const Case& c : cases) {
    cudaFuncSetAttribute(small, cudaFuncAttributePreferredSharedMemoryCarveout, c.carve ? 100 : -1);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < N; ++i) {
      cudaError_t e;
      if (c.pattern == 2 || (c.pattern == 1 && (i & 3) == 3))
        e = launch_small(512, c.pdl, st, out);
      else
        e = launch(c.prologue ? big<true> : big<false>, c.grid, c.pattern == 3 ? 8192 : kSmem, c.pdl, st, out);
      if (e != cudaSuccess) {
        printf("launch failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
    }
    cudaStreamEndCapture(st, &g);
    if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
      printf("instantiate failed\n");
      return 1;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("%-62s %7.2f us per node%s\n", c.name, ms * 1e3 / (reps * N), err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
  }
  return 0;
}
