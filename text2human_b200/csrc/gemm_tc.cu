// t2h_tapgemm: persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//
//   warp 0   TMA producer   cp.async.bulk.tensor (4-D activation boxes with
//                           zero OOB fill = conv padding; 3-D weight boxes)
//   warp 1   MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16),
//                           accumulators live in tensor memory, double-buffered
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      tcgen05.ld -> alpha/bias/GELU/residual -> global
//
// The contraction loop runs over (tap, 64-channel chunk, product term).  A tap
// is a spatial shift of the activation box (3x3 conv = 9 taps, 1x1/Linear/bmm
// = 1 tap); a product term selects which fp16 planes feed the MMA so that
// hi*hi + hi*lo + lo*hi reproduces an fp32 product on the fp16 tensor pipe.
//
// Replaces cuDNN/cuBLAS calls made by the reference modules — see include/t2h.h.
#include <cuda.h>

#include "t2h_internal.h"
#include "t2h_ptx.cuh"

namespace t2h {

struct TapGemmDev {
  int n_img, H, W;
  int TW, TH;  // spatial box of one 128-row block
  int tiles_w, tiles_h, n_tiles_n, total_tiles;
  int n_out, C, kchunks, nterms;
  int a_term_imgs, a_bcast, b_term_g, b_batched, b_batched_h;
  int ntaps;
  int tap_dy[T2H_MAX_TAPS], tap_dx[T2H_MAX_TAPS], tap_img_off[T2H_MAX_TAPS];
  void* d;
  int d_mode, d_terms, vec_ok;
  long long d_plane, d_sn, d_sh, d_sw, d_sc;
  const float* bias;
  int bias_mode, act;
  float alpha;
  const float* residual;
  double* gn_stats;
  int gn_cpg;
};

constexpr int kBK = 64;              // fp16 elements per 128-byte swizzled row
constexpr int kABlockBytes = 128 * 128;  // one 128-row A block of a stage
constexpr int kThreads = 256;

template <int BN, int MBLK>
struct Cfg {
  static constexpr int kStageBytes = MBLK * kABlockBytes + BN * 128;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccCols = MBLK * BN;
  static constexpr int kTmemCols = (2 * kAccCols) < 32 ? 32 : (2 * kAccCols);
  static constexpr int kChunk = BN < 32 ? BN : 32;  // columns per tcgen05.ld
  // >= 120 KB so that at most one CTA (and one TMEM allocation) lives on an SM
  static constexpr int kSmemBytes =
      (kStages * kStageBytes + 1024) < 120 * 1024 ? 120 * 1024 : (kStages * kStageBytes + 1024);
  static_assert(kTmemCols <= 512, "accumulators exceed tensor memory");
  static_assert((kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be a power of two");
};

struct TileCoord {
  int img, h0, w0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const TapGemmDev& P, int tile, int mblk, int bn) {
  TileCoord t;
  int n_tile = tile % P.n_tiles_n;
  int m_tile = tile / P.n_tiles_n;
  int per_img = P.tiles_h * P.tiles_w;
  t.img = m_tile / per_img;
  int rem = m_tile - t.img * per_img;
  int ty = rem / P.tiles_w;
  int tx = rem - ty * P.tiles_w;
  t.h0 = ty * P.TH * mblk;
  t.w0 = tx * P.TW;
  t.n0 = n_tile * bn;
  return t;
}

template <int BN, int MBLK>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ TapGemmDev P) {
  using C = Cfg<BN, MBLK>;
  constexpr int STAGES = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kiters = P.ntaps * P.kchunks * P.nterms;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(P, tile, MBLK, BN);
        const int a_img = P.a_bcast ? 0 : t.img;
        const int b_g2 = P.b_batched ? t.img : 0;
        const int b_g = P.b_batched_h ? t.h0 : 0;
        for (int tap = 0; tap < P.ntaps; ++tap) {
          const int dy = P.tap_dy[tap], dx = P.tap_dx[tap], ioff = P.tap_img_off[tap];
          for (int ch = 0; ch < P.kchunks; ++ch) {
            for (int term = 0; term < P.nterms; ++term) {
              const int ta = (term == 2) ? P.a_term_imgs : 0;  // lo plane of A
              const int tb = (term == 1) ? P.b_term_g : 0;     // lo plane of B
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], C::kStageBytes);
              uint8_t* sa = smem + stage * C::kStageBytes;
#pragma unroll
              for (int mb = 0; mb < MBLK; ++mb)
                tma_load_4d(&tmA, &full_bar[stage], sa + mb * kABlockBytes, ch * kBK, t.w0 + dx,
                            t.h0 + mb * P.TH + dy, a_img + ioff + ta);
              tma_load_4d(&tmB, &full_bar[stage], sa + MBLK * kABlockBytes, ch * kBK, t.n0,
                          b_g + tap + tb, b_g2);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_f16(128, BN);
      const int last_steps = (P.C - (P.kchunks - 1) * kBK + 15) / 16;  // UMMA_K=16 steps
      int stage = 0, phase = 0, as = 0, ap = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], ap ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * C::kAccCols;
        for (int it = 0; it < kiters; ++it) {
          const int ch = (it / P.nterms) % P.kchunks;
          const int ksteps = (ch == P.kchunks - 1) ? last_steps : 4;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t b_addr = a_addr + MBLK * kABlockBytes;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t bdesc = umma_desc_k128(b_addr + j * 32);
#pragma unroll
            for (int mb = 0; mb < MBLK; ++mb) {
              const uint64_t adesc = umma_desc_k128(a_addr + mb * kABlockBytes + j * 32);
              umma_f16(d_base + mb * BN, adesc, bdesc, IDESC, (it > 0 || j > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
        if (++as == 2) {
          as = 0;
          ap ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    constexpr int CH = C::kChunk;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int th = row / P.TW, tw = row - th * P.TW;
    int as = 0, ap = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(P, tile, MBLK, BN);
      mbar_wait(&tfull_bar[as], ap);
      tc_fence_after();
#pragma unroll 1
      for (int mb = 0; mb < MBLK; ++mb) {
        const int h = t.h0 + mb * P.TH + th;
        const int w = t.w0 + tw;
        const bool row_ok = (h < P.H) && (w < P.W);
        const long long off = (long long)t.img * P.d_sn + (long long)h * P.d_sh + (long long)w * P.d_sw;
        const float row_bias = (P.bias_mode == T2H_BIAS_ROW && row_ok) ? P.bias[h * P.W + w] : 0.f;
#pragma unroll 1
        for (int cc = 0; cc < BN / CH; ++cc) {
          const int col0 = t.n0 + cc * CH;
          if (col0 >= P.n_out) break;  // warp-uniform
          uint32_t r[32];
          const uint32_t taddr =
              tmem_base + (uint32_t(q * 32) << 16) + as * C::kAccCols + mb * BN + cc * CH;
          if constexpr (CH == 32)
            tmem_ld_32x32(taddr, r);
          else
            tmem_ld_32x16(taddr, r);
          tmem_ld_wait();
          if (!row_ok) continue;
          float v[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]) * P.alpha + row_bias;
          if (P.bias_mode == T2H_BIAS_COL) {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (col0 + i < P.n_out) v[i] += __ldg(P.bias + col0 + i);
          }
          if (P.act == T2H_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = gelu_erf(v[i]);
          }
          if (P.d_mode == T2H_OUT_F32) {
            float* dst = reinterpret_cast<float*>(P.d);
            if (P.vec_ok) {
#pragma unroll
              for (int i = 0; i < CH; i += 4) {
                if (col0 + i < P.n_out) {
                  float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                  if (P.residual) {
                    const float4 rr = *reinterpret_cast<const float4*>(P.residual + off + col0 + i);
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                  }
                  *reinterpret_cast<float4*>(dst + off + col0 + i) = o;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < CH; ++i) {
                if (col0 + i < P.n_out) {
                  const long long o = off + (long long)(col0 + i) * P.d_sc;
                  float val = v[i];
                  if (P.residual) val += P.residual[o];
                  dst[o] = val;
                }
              }
            }
          } else {
            __half* dst = reinterpret_cast<__half*>(P.d);
            if (P.vec_ok) {
#pragma unroll
              for (int i = 0; i < CH; i += 8) {
                if (col0 + i < P.n_out) {
                  __align__(16) __half hi[8];
                  __align__(16) __half lo[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) split_f16(v[i + e], hi[e], lo[e]);
                  *reinterpret_cast<uint4*>(dst + off + col0 + i) = *reinterpret_cast<uint4*>(hi);
                  if (P.d_terms == 2)
                    *reinterpret_cast<uint4*>(dst + P.d_plane + off + col0 + i) =
                        *reinterpret_cast<uint4*>(lo);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < CH; ++i) {
                if (col0 + i < P.n_out) {
                  const long long o = off + (long long)(col0 + i) * P.d_sc;
                  __half hi, lo;
                  split_f16(v[i], hi, lo);
                  dst[o] = hi;
                  if (P.d_terms == 2) dst[P.d_plane + o] = lo;
                }
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        ap ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

static int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_elems, const uint32_t* box, const char* what) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) return fail(T2H_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_elems[i] * 2;  // bytes
      if (gstr[i - 1] % 16 != 0)
        return fail(T2H_EINVAL, "%s: stride of dim %d (%llu bytes) is not a multiple of 16", what, i,
                    (unsigned long long)gstr[i - 1]);
    }
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0)
    return fail(T2H_EINVAL, "%s: base pointer not 16-byte aligned", what);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(T2H_ECUDA,
                "%s: cuTensorMapEncodeTiled failed (%d) dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
                what, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return T2H_OK;
}

template <int BN, int MBLK>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TapGemmDev& P,
                  cudaStream_t stream) {
  using C = Cfg<BN, MBLK>;
  static bool configured = false;
  if (!configured) {
    T2H_CUDA(cudaFuncSetAttribute(tapgemm_kernel<BN, MBLK>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  tapgemm_kernel<BN, MBLK><<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, P);
  T2H_LAUNCH_OK();
  return T2H_OK;
}

}  // namespace t2h

using namespace t2h;

extern "C" int t2h_tapgemm(const t2h_tapgemm_params* p, t2h_stream_t stream) {
  T2H_CHECK_ARG(p && p->a && p->b && p->d, "tapgemm: null operand");
  T2H_CHECK_ARG(p->n_img > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->n_out > 0,
                "tapgemm: empty problem (n_img=%d H=%d W=%d C=%d n_out=%d)", p->n_img, p->H, p->W, p->C,
                p->n_out);
  T2H_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= T2H_MAX_TAPS, "tapgemm: ntaps=%d", p->ntaps);
  T2H_CHECK_ARG(p->nterms == 1 || p->nterms == 3, "tapgemm: nterms must be 1 or 3 (got %d)", p->nterms);
  T2H_CHECK_ARG(p->nterms == 1 || (p->a_terms == 2 && p->b_terms == 2),
                "tapgemm: nterms=3 needs hi/lo planes on both operands");
  T2H_CHECK_ARG(p->d_mode == T2H_OUT_F32 || p->d_mode == T2H_OUT_PLANES, "tapgemm: d_mode=%d", p->d_mode);
  T2H_CHECK_ARG(p->d_mode == T2H_OUT_F32 || p->d_terms == 1 || p->d_terms == 2, "tapgemm: d_terms=%d",
                p->d_terms);
  T2H_CHECK_ARG(p->d_mode == T2H_OUT_F32 || p->residual == nullptr,
                "tapgemm: residual add needs fp32 output");
  T2H_CHECK_ARG(p->bias_mode == T2H_BIAS_NONE || p->bias != nullptr, "tapgemm: bias_mode without bias");
  T2H_CHECK_ARG(!p->b_batched_h || p->H == 1 || p->tile_rows, "tapgemm: b_batched_h needs row tiles");
  T2H_CHECK_ARG(p->gn_stats == nullptr, "tapgemm: fused GroupNorm statistics not implemented yet");

  TapGemmDev P;
  P.n_img = p->n_img; P.H = p->H; P.W = p->W;
  P.n_out = p->n_out; P.C = p->C; P.kchunks = (p->C + kBK - 1) / kBK; P.nterms = p->nterms;
  P.a_term_imgs = p->a_term_imgs; P.a_bcast = p->a_bcast;
  P.b_term_g = p->b_term_g; P.b_batched = p->b_batched; P.b_batched_h = p->b_batched_h;
  P.ntaps = p->ntaps;
  for (int i = 0; i < T2H_MAX_TAPS; ++i) {
    P.tap_dy[i] = i < p->ntaps ? p->tap_dy[i] : 0;
    P.tap_dx[i] = i < p->ntaps ? p->tap_dx[i] : 0;
    P.tap_img_off[i] = i < p->ntaps ? p->tap_img_off[i] : 0;
  }
  P.d = p->d; P.d_mode = p->d_mode; P.d_terms = p->d_terms; P.d_plane = p->d_plane;
  P.d_sn = p->d_sn; P.d_sh = p->d_sh; P.d_sw = p->d_sw; P.d_sc = p->d_sc;
  P.bias = p->bias; P.bias_mode = p->bias_mode; P.act = p->act; P.alpha = p->alpha;
  P.residual = p->residual; P.gn_stats = p->gn_stats; P.gn_cpg = p->gn_cpg;

  // ---- tile shape
  // one 128-row block = TH x TW output positions of one image
  int TW, TH;
  if (p->H == 1 || p->tile_rows) {
    TW = 128; TH = 1;
  } else {
    TW = 16;
    while (TW > p->W && TW > 1) TW >>= 1;  // W < 16: narrower, taller boxes
    TH = 128 / TW;
  }
  int BN = 16;
  while (BN < p->n_out && BN < 256) BN <<= 1;
  int MBLK = 1;
  if (BN == 128 && p->H >= 2 * TH && !p->tile_rows) {
    long long tiles1 = (long long)p->n_img * ceil_div(p->H, TH) * ceil_div(p->W, TW);
    if (tiles1 >= 4LL * num_sms()) MBLK = 2;
  }
  P.TW = TW; P.TH = TH;
  P.tiles_w = ceil_div(p->W, TW);
  P.tiles_h = ceil_div(p->H, TH * MBLK);
  P.n_tiles_n = ceil_div(p->n_out, BN);
  long long total = (long long)p->n_img * P.tiles_w * P.tiles_h * P.n_tiles_n;
  T2H_CHECK_ARG(total < (1LL << 31), "tapgemm: too many tiles");
  P.total_tiles = (int)total;

  // ---- vectorised epilogue eligibility
  const int vw = (p->d_mode == T2H_OUT_F32) ? 4 : 8;
  bool vec = p->d_sc == 1 && p->n_out % vw == 0 && p->d_sw % vw == 0 && p->d_sh % vw == 0 &&
             p->d_sn % vw == 0 && (reinterpret_cast<uintptr_t>(p->d) % 16 == 0);
  if (p->d_mode == T2H_OUT_PLANES && p->d_terms == 2) vec = vec && (p->d_plane % 8 == 0);
  if (p->residual) vec = vec && (reinterpret_cast<uintptr_t>(p->residual) % 16 == 0);
  P.vec_ok = vec ? 1 : 0;

  // ---- tensor maps
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->a_W, (uint64_t)p->a_H, (uint64_t)p->a_imgs};
    uint64_t str[4] = {1, (uint64_t)p->a_sw, (uint64_t)p->a_sh, (uint64_t)p->a_sn};
    uint32_t box[4] = {(uint32_t)kBK, (uint32_t)TW, (uint32_t)TH, 1};
    int rc = make_tmap(&tmA, p->a, 4, dims, str, box, "tapgemm A");
    if (rc) return rc;
  }
  {
    const int g2 = p->b_groups2 > 0 ? p->b_groups2 : 1;
    uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->n_out, (uint64_t)p->b_groups, (uint64_t)g2};
    uint64_t str[4] = {1, (uint64_t)p->b_sn, (uint64_t)p->b_sg,
                       (uint64_t)(g2 > 1 ? p->b_sg2 : p->b_sg)};
    uint32_t box[4] = {(uint32_t)kBK, (uint32_t)BN, 1, 1};
    int rc = make_tmap(&tmB, p->b, 4, dims, str, box, "tapgemm B");
    if (rc) return rc;
  }

  cudaStream_t s = as_stream(stream);
  if (MBLK == 2) return launch<128, 2>(tmA, tmB, P, s);
  switch (BN) {
    case 16: return launch<16, 1>(tmA, tmB, P, s);
    case 32: return launch<32, 1>(tmA, tmB, P, s);
    case 64: return launch<64, 1>(tmA, tmB, P, s);
    case 128: return launch<128, 1>(tmA, tmB, P, s);
    default: return launch<256, 1>(tmA, tmB, P, s);
  }
}
