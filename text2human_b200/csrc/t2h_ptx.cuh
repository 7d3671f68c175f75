// Inline-PTX wrappers for the sm_100a primitives the kernels use: mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// UMMA shared-memory + instruction descriptors.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace t2h {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// --------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const void* tmap, uint64_t* bar, void* smem, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Cluster multicast: the box is written to the same shared-memory offset of every CTA in `cta_mask`, and the
// mbarrier at the same offset in each of them receives the complete_tx.
__device__ __forceinline__ void tma_load_4d_mc(const void* tmap, uint64_t* bar, void* smem, int c0, int c1, int c2,
                                               int c3, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// L2 prefetch of a 4-D box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* tmap, uint64_t* bar, void* smem, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2)
      : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute
// may start while its predecessor is still running; `pdl_wait` blocks until every prerequisite grid has
// completed and its memory is visible, `pdl_launch_dependents` lets the NEXT kernel's CTAs start their
// prologue (barrier init, TMEM allocation, descriptor prefetch) early.  Both are no-ops without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// TMA store smem -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* smem, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA reduce-add smem -> global (fp32 elements are added to memory; same completion as a store)
__device__ __forceinline__ void tma_reduce_add_4d(const void* tmap, const void* smem, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until at most N of this thread's committed store groups still read shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 operands, fp32 accumulate, single CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread retire
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// the same arrival delivered to the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane
// (taddr.lane + i), columns taddr.col .. +31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// --------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor for a K-major tile stored as rows of 128
// bytes (64 fp16) with the 128-byte swizzle TMA applies: 8-row groups are
// 1024 bytes apart (SBO), LBO unused for swizzled K-major layouts.
//   bits [0,14)  start address >> 4        bits [16,30) LBO >> 4
//   bits [32,46) SBO >> 4                  bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same descriptor split for tight issue loops: a constant high word and a low word that advances
// by (bytes >> 4) — +2 per UMMA_K=16 step of fp16 inside the 128-byte swizzled row.
constexpr uint32_t kUmmaDescHiK128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16);
}
__device__ __forceinline__ uint64_t umma_desc_join(uint32_t lo) {
  return (static_cast<uint64_t>(kUmmaDescHiK128) << 32) | lo;
}
// MN-major ("transposed") operand, as TMA deposits [64 k rows][64 MN elements] boxes with the 128-byte
// swizzle, one 8 KB box per 64 rows/columns of the operand: same high word (SBO = 1024 B between groups of
// 8 k rows), LBO = 8192 B (box to box along MN) in the low word, and a K=16 step advances the start address
// by 16 k rows = 2048 B.  The instruction descriptor carries the major-ness (bit 15 for A, 16 for B).
// Encoding verified on hardware by tools/umma_mn_probe.cu (profiles/r01_umma_mn_major_probe.log).
constexpr uint32_t kUmmaStepK = 2;     // descriptor low-word advance per K=16 step, K-major
constexpr uint32_t kUmmaStepMN = 128;  // ... MN-major (2048 B >> 4)
__device__ __forceinline__ uint32_t umma_desc_lo_mn(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFF) >> 4) | ((8192u >> 4) << 16);
}
// `ksteps` (<= 4) K=16 steps of D (+)= A*B; acc is 0 only for the very first MMA of an accumulator.
__device__ __forceinline__ void umma_ksteps(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                            int ksteps, uint32_t& acc, uint32_t a_step = kUmmaStepK,
                                            uint32_t b_step = kUmmaStepK) {
  if (ksteps == 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      umma_f16(d_tmem, umma_desc_join(a_lo + a_step * j), umma_desc_join(b_lo + b_step * j), idesc, acc);
      acc = 1;
    }
  } else {
    for (int j = 0; j < ksteps; ++j) {
      umma_f16(d_tmem, umma_desc_join(a_lo + a_step * j), umma_desc_join(b_lo + b_step * j), idesc, acc);
      acc = 1;
    }
  }
}

// Instruction descriptor, kind::f16: fp16 A/B (format 0), fp32 accumulate
// (c_format 1, bits [4,6)), both K-major, N>>3 at bits [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// one lane of a fully converged warp (warp-uniform control flow keeps the issue loop on the uniform
// datapath; only the tcgen05 instruction itself is predicated on the elected lane)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------ misc helpers
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU(), transformer_arch.py:85), branch-free: erfc(|z|) = t P(t) e^{-z^2},
// t = 1 / (1 + 0.3275911 |z|) (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 on erf), evaluated as
// x >= 0: 0.5 x (2 - erfc|z|),  x < 0: 0.5 x erfc|z| -- no 1 + erf cancellation in the negative tail.  Against the
// fp64 function over [-8, 8] the absolute error is 4.2e-7, the same as the fp32 formula with an exact erff (4.5e-7,
// set by rounding at |x| ~ 6), at ~14 instructions instead of erff's two divergent ~50-instruction branches
// (the tap-GEMM epilogue of the transformer's fc1 applies it to 256 columns per thread).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(z * z) * 1.4426950408889634f));
  const float c = p * t * e;  // erfc(|z|)
  return 0.5f * x * (x >= 0.f ? 2.0f - c : c);
}
// derivative of the activation that follows a norm, as a function of the pre-activation u (act: 0 none,
// 1 swish = u*sigmoid(u), 2 LeakyReLU(0.2)); shared by t2h_norm_bwd's kernels and the conv epilogue that fuses its
// first pass
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == 1) {
    const float s = 1.0f / (1.0f + __expf(-u));
    return s * (1.0f + u * (1.0f - s));
  }
  if (act == 2) return u > 0.f ? 1.0f : 0.2f;
  return 1.0f;
}
// the same with an approximate reciprocal (2 MUFU + 7 FP32 instructions; relative error ~1e-6): for the sums the
// conv epilogue accumulates, where the precise division's slow path would sit on every output element
__device__ __forceinline__ float act_grad_fast(float u, int act) {
  if (act == 1) {
    float s;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + __expf(-u)));
    return s * fmaf(u, 1.0f - s, 1.0f);
  }
  if (act == 2) return u > 0.f ? 1.0f : 0.2f;
  return 1.0f;
}
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

}  // namespace t2h
