// t2h_attn_fwd: softmax(q k^T * scale) v of the index-prediction transformer's multi-head attention as ONE
// kernel per layer (reference: models/archs/transformer_arch.py:37-71, CausalSelfAttention.forward with
// causal=False) -- replaces the q k^T tap-GEMM, the softmax kernel and the p v tap-GEMM and their [B, nh, T, T]
// round trips through HBM.  Included by gemm_tc.cu (shares its tensor-map helper).
//
// One CTA per (sequence, head, block of 128 queries); head_dim = 64, T = 128 .. 512 keys (a multiple of 128):
//
//   warp 0     TMA producer   Q block and the K rows of the head as 64-row boxes of the fused q|k|v projection
//                             (fp16 hi / lo planes), then -- once the score MMAs have retired -- the V rows into the
//                             same shared-memory region
//   warp 1     MMA issuer     S = Q K^T into tensor memory (up to 512 fp32 columns = the whole key range; 3 products
//                             hi*lo + hi*hi + lo*hi in parity mode), then O += P_j V_j per 64-key chunk (V consumed
//                             MN-major, i.e. token-major as the projection wrote it; O reuses S's first 64 columns,
//                             which have been drained by then)
//   warp 2     TMEM allocator
//   warps 4-7  softmax        one query row per thread (TMEM lane = row, so max / sum need no shuffles): pass 1 row
//                             max, pass 2 e = 2^((s - max) * scale * log2 e), split into fp16 hi / lo and stored as
//                             the 128B-swizzled K-major A operand of the P V product (2-slot ring); finally
//                             O / sum(e) -> fp16 planes, heads side by side
#pragma once

namespace t2h {

struct AttnDev {
  int tokens, heads, qblocks, nchunks;  // nchunks = tokens / 64
  int terms;                            // 1: single fp16 plane, 2: hi + lo planes (3 tensor-core products)
  int q_col, k_col, v_col;
  float kfac;  // scale * log2(e)
  int debug;   // T2H_DEBUG (bit 16: timeline record, see g_t2h_dbg)
  __half* out;
  long long out_plane, ld_out;
};

constexpr int kAttnThreads = 256;
constexpr int kAttnQPlane = 128 * 128;   // 128 query rows x 64 fp16
constexpr int kAttnKVPlane = 512 * 128;  // up to 512 key (value) rows x 64 fp16
constexpr int kAttnPPlane = 128 * 128;   // one P chunk: 128 rows x 64 keys
constexpr int kAttnSmem = 2 * kAttnQPlane + 2 * kAttnKVPlane + 2 * 2 * kAttnPPlane + 1024;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ AttnDev P) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* qs = smem;                   // [plane][128 rows][128 B]
  uint8_t* kv = qs + 2 * kAttnQPlane;   // [plane][512 rows][128 B]: K, later V
  uint8_t* ps = kv + 2 * kAttnKVPlane;  // [slot][plane][128 rows][128 B]

  __shared__ __align__(8) uint64_t kv_full[8];  // 64-row chunk c: phase 0 = K (chunk 0 also carries Q), phase 1 = V
  __shared__ __align__(8) uint64_t s_full, o_full;
  __shared__ __align__(8) uint64_t p_full[2], p_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  __shared__ unsigned long long* trace_s;
  if (warp == 0 && lane == 0) {
    trace_s = nullptr;
    if ((P.debug & 16) && blockIdx.x == 0) {
      trace_s = g_t2h_dbg + (size_t)(atomicAdd(&g_t2h_dbg_n, 1u) % kTraceRecords) * 8;
      trace_s[0] = gtime_ns();
      trace_s[6] = gridDim.x;
      trace_s[7] = (1ull << 32) | (unsigned long long)P.nchunks;
    }
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int c = 0; c < 8; ++c) mbar_init(&kv_full[c], 1);
    mbar_init(&s_full, 1);
    mbar_init(&o_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&p_full[s], 128);
      mbar_init(&p_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();
  unsigned long long* const trace = trace_s;
  if (trace && threadIdx.x == 0) trace[1] = gtime_ns();

  const int qb = (int)blockIdx.x % P.qblocks;
  const int hb = (int)blockIdx.x / P.qblocks;
  const int head = hb % P.heads;
  const int seq = hb / P.heads;
  const int row0 = seq * P.tokens;    // first row of this sequence in the [rows][ld] projection
  const int qrow0 = row0 + qb * 128;  // first query row of this CTA
  const int NC = P.nchunks;
  const int T2 = P.terms;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int c = 0; c < NC; ++c) {
        mbar_expect_tx(&kv_full[c], (uint32_t)(8192 * T2 + (c == 0 ? kAttnQPlane * T2 : 0)));
        if (c == 0) {
          for (int pl = 0; pl < T2; ++pl)
            for (int hbox = 0; hbox < 2; ++hbox)
              tma_load_3d(&tmX, &kv_full[0], qs + pl * kAttnQPlane + hbox * 8192, P.q_col + head * 64,
                          qrow0 + 64 * hbox, pl);
        }
        for (int pl = 0; pl < T2; ++pl)
          tma_load_3d(&tmX, &kv_full[c], kv + pl * kAttnKVPlane + c * 8192, P.k_col + head * 64, row0 + 64 * c, pl);
      }
      mbar_wait(&s_full, 0);  // every score MMA has retired: K is dead, its rows take V
      for (int c = 0; c < NC; ++c) {
        mbar_expect_tx(&kv_full[c], (uint32_t)(8192 * T2));
        for (int pl = 0; pl < T2; ++pl)
          tma_load_3d(&tmX, &kv_full[c], kv + pl * kAttnKVPlane + c * 8192, P.v_col + head * 64, row0 + 64 * c, pl);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    const uint32_t q_lo = umma_desc_lo(smem_u32(qs));
    const uint32_t k_lo = umma_desc_lo(smem_u32(kv));
    const uint32_t p_lo = umma_desc_lo(smem_u32(ps));
    const uint32_t v_lo = umma_desc_lo_mn(smem_u32(kv));
    constexpr uint32_t QPL = kAttnQPlane >> 4, KVPL = kAttnKVPlane >> 4, PPL = kAttnPPlane >> 4;
    // S[:, 256 hf ..] = Q K^T, N = 256 keys per MMA (the tail half may be 128)
    const int nhalf = (NC + 3) >> 2;
    for (int hf = 0; hf < nhalf; ++hf) {
      const int nck = min(4, NC - 4 * hf);
      for (int c = 4 * hf; c < 4 * hf + nck; ++c) mbar_wait(&kv_full[c], 0);
      tc_fence_after();
      if (trace && lane == 0 && hf == 0) trace[2] = gtime_ns();
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_f16(128, 64 * nck);
        const uint32_t d = tmem_base + hf * 256;
        const uint32_t kb = k_lo + hf * (32768u >> 4);
        uint32_t acc = 0;
        if (T2 == 2) {
          umma_ksteps(d, q_lo, kb + KVPL, idesc, 4, acc);  // hi * lo
          umma_ksteps(d, q_lo, kb, idesc, 4, acc);         // hi * hi
          umma_ksteps(d, q_lo + QPL, kb, idesc, 4, acc);   // lo * hi
        } else {
          umma_ksteps(d, q_lo, kb, idesc, 4, acc);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&s_full);
    __syncwarp();
    // O (TMEM columns 0..63, drained of S by the time P_0 is complete) += P_j V_j
    const uint32_t idesc_pv = umma_idesc_f16(128, 64) | (1u << 16);  // B = V is MN-major
    uint32_t acc_o = 0;
    for (int j = 0; j < NC; ++j) {
      const int s = j & 1, k = j >> 1;
      mbar_wait(&kv_full[j], 1);
      mbar_wait(&p_full[s], k & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pa = p_lo + s * (2 * PPL);
        const uint32_t vb = v_lo + j * (8192u >> 4);
        if (T2 == 2) {
          umma_ksteps(tmem_base, pa, vb + KVPL, idesc_pv, 4, acc_o, kUmmaStepK, kUmmaStepMN);  // hi * lo
          umma_ksteps(tmem_base, pa, vb, idesc_pv, 4, acc_o, kUmmaStepK, kUmmaStepMN);         // hi * hi
          umma_ksteps(tmem_base, pa + PPL, vb, idesc_pv, 4, acc_o, kUmmaStepK, kUmmaStepMN);   // lo * hi
        } else {
          umma_ksteps(tmem_base, pa, vb, idesc_pv, 4, acc_o, kUmmaStepK, kUmmaStepMN);
        }
        umma_commit(&p_empty[s]);
      }
      acc_o = 1;
      __syncwarp();
    }
    if (elect_one()) umma_commit(&o_full);
    __syncwarp();
    if (trace && lane == 0) trace[3] = gtime_ns();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output, one query row per thread
    const int q = warp & 3;  // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    mbar_wait(&s_full, 0);
    tc_fence_after();
    float m = -INFINITY;
    for (int c = 0; c < 2 * NC; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
    }
    // (s - m) * kfac as one fma: the rounding of m * kfac is common to the row and cancels in e / sum(e)
    const float mk = m * P.kfac;
    float sum = 0.f;
    for (int j = 0; j < NC; ++j) {
      const int s = j & 1, k = j >> 1;
      mbar_wait(&p_empty[s], (k & 1) ^ 1);
      uint8_t* phi = ps + s * (2 * kAttnPPlane);
      uint8_t* plo = phi + kAttnPPlane;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32(trow + j * 64 + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __align__(16) __half hi[8];
          __align__(16) __half lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float ev = ex2_approx(fmaf(__uint_as_float(r[8 * g + e]), P.kfac, -mk));
            sum += ev;
            split_f16(ev, hi[e], lo[e]);
          }
          *reinterpret_cast<uint4*>(phi + swz(row, half * 4 + g)) = *reinterpret_cast<uint4*>(hi);
          if (T2 == 2) *reinterpret_cast<uint4*>(plo + swz(row, half * 4 + g)) = *reinterpret_cast<uint4*>(lo);
        }
      }
      tc_fence_before();         // the TMEM reads above precede the MMAs that overwrite columns 0..63
      fence_proxy_async_smem();  // P chunk visible to the tensor core's shared-memory reads
      mbar_arrive(&p_full[s]);
    }
    mbar_wait(&o_full, 0);
    tc_fence_after();
    if (trace && threadIdx.x == 128) trace[4] = gtime_ns();
    const float inv = 1.0f / sum;
    __half* ohi = P.out + (long long)(qrow0 + row) * P.ld_out + head * 64;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld_32x32(trow + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_f16(__uint_as_float(r[8 * g + e]) * inv, hi[e], lo[e]);
        *reinterpret_cast<uint4*>(ohi + half * 32 + g * 8) = *reinterpret_cast<uint4*>(hi);
        if (T2 == 2) *reinterpret_cast<uint4*>(ohi + P.out_plane + half * 32 + g * 8) = *reinterpret_cast<uint4*>(lo);
      }
    }
    if (trace && threadIdx.x == 128) trace[5] = gtime_ns();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace t2h

extern "C" int t2h_attn_fwd(const void* qkv, int terms, int64_t plane, int64_t ld, int64_t rows, int q_col, int k_col,
                            int v_col, int batch, int tokens, int heads, int head_dim, float scale, void* out,
                            int64_t out_plane, int64_t ld_out, t2h_stream_t stream) {
  using namespace t2h;
  T2H_CHECK_ARG(qkv && out && batch > 0 && heads > 0, "attn_fwd: bad args");
  T2H_CHECK_ARG(terms == 1 || terms == 2, "attn_fwd: terms=%d", terms);
  T2H_CHECK_ARG(head_dim == 64, "attn_fwd: head_dim=%d (only 64 is built; use the q k^T / softmax / p v launches)", head_dim);
  T2H_CHECK_ARG(tokens >= 128 && tokens <= 512 && tokens % 128 == 0,
                "attn_fwd: tokens=%d (need a multiple of 128 in 128..512: the whole key range lives in tensor memory)", tokens);
  T2H_CHECK_ARG(rows >= (int64_t)batch * tokens, "attn_fwd: rows=%lld < batch*tokens", (long long)rows);
  T2H_CHECK_ARG(q_col >= 0 && k_col >= 0 && v_col >= 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 &&
                    q_col + heads * 64 <= ld && k_col + heads * 64 <= ld && v_col + heads * 64 <= ld,
                "attn_fwd: q/k/v columns outside the projection (ld=%lld)", (long long)ld);
  T2H_CHECK_ARG(ld_out % 8 == 0 && out_plane % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                    ld_out >= heads * 64,
                "attn_fwd: output planes need 16-byte aligned rows");
  CUtensorMap tmX;
  const uint64_t dims[3] = {(uint64_t)ld, (uint64_t)rows, (uint64_t)terms};
  const uint64_t strides[3] = {1, (uint64_t)ld, (uint64_t)(terms == 2 ? plane : rows * ld)};
  const uint32_t box[3] = {64, 64, 1};
  int rc = make_tmap(&tmX, qkv, 2, 3, dims, strides, box, "attn_fwd qkv");
  if (rc != T2H_OK) return rc;
  static bool configured[64] = {};
  int dev = 0;
  T2H_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    T2H_CUDA(cudaFuncSetAttribute(attn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured[dev & 63] = true;
  }
  AttnDev P;
  memset(&P, 0, sizeof(P));
  P.tokens = tokens;
  P.heads = heads;
  P.qblocks = tokens / 128;
  P.nchunks = tokens / 64;
  P.terms = terms;
  P.q_col = q_col;
  P.k_col = k_col;
  P.v_col = v_col;
  P.kfac = scale * 1.4426950408889634f;
  P.out = reinterpret_cast<__half*>(out);
  P.out_plane = out_plane;
  P.ld_out = ld_out;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("T2H_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    P.debug = dbg;
  }
  const int grid = batch * heads * P.qblocks;
  T2H_CUDA(launch_pdl(attn_fused_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmem, as_stream(stream), 1, tmX, P));
  T2H_LAUNCH_OK();
  return T2H_OK;
}
