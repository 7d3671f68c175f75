// t2h_tapgemm: persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//
//   warp 0   A producer     cp.async.bulk.tensor 4-D activation "slabs": the rows a tile needs
//                           for all vertical taps of one horizontal shift (zero OOB fill =
//                           conv padding), so a 3x3 conv loads activations 3x, not 9x
//   warp 3   B producer     4-D weight boxes, one per tap (own ring: no head-of-line blocking)
//   warp 1   MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16) on shifted views
//                           of the slab; accumulators live in tensor memory, double-buffered
//   warp 2   TMEM allocator
//   warps 4-11 epilogue     (8-11 only in the plane / plain-fp32 modes, taking the upper 32 columns of each
//                           64-column unit) tcgen05.ld -> alpha/bias/GELU/residual/GroupNorm
//                           partial sums -> swizzled smem staging -> TMA store
//                           (residual tiles arrive by TMA load into smem)
//
// The contraction loop runs over (tap group, 64-channel chunk, tap, product term).  A tap is a
// spatial shift of the activation box (3x3 conv = 9 taps, 1x1/Linear/bmm = 1 tap); taps with the
// same horizontal shift form a group and share one slab.  A product term selects which fp16
// planes feed the MMA so that hi*lo + hi*hi + lo*hi reproduces an fp32 product on the fp16 tensor
// pipe; the hi/lo slabs and weight tiles are each loaded once per (group, chunk[, tap]).
//
// Replaces cuDNN/cuBLAS calls made by the reference modules — see include/t2h.h.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "t2h_internal.h"
#include "t2h_ptx.cuh"

namespace t2h {

enum { EPI_DIRECT = 0, EPI_TMA_F32 = 1, EPI_TMA_PLANES = 2 };

struct TapGemmDev {
  int n_img, H, W;
  int TW, TH;  // spatial box of one 128-row block
  int tiles_w, tiles_h, n_tiles_n, total_tiles;
  int n_out, C, kchunks, nterms;
  int pair;        // 1: launched as clusters of two CTAs that own vertically adjacent row tiles and share every
                   // weight tile: each CTA loads every other B tile and TMA-multicasts it to both
  int a_mn, b_mn;  // operand stored contraction-major (rows / columns contiguous): MN-major UMMA descriptors
  int ksplit, kper, total_work;  // split-K: work item = (tile, k-slice); total_work = total_tiles * ksplit
  int a_term_imgs, a_bcast, b_term_g, b_batched, b_batched_h;
  // taps grouped by (dx, img_off): one activation slab per group
  int ngroups, slab_rows;
  int a_slots, b_slots;  // ring depths (A slabs / B tiles)
  int debug;             // T2H_DEBUG bits (profiling experiments only): 1 skip epilogue work,
                         // 2 skip MMA issue, 4 skip TMA loads
  int g_dx[T2H_MAX_TAPS], g_ioff[T2H_MAX_TAPS], g_dy0[T2H_MAX_TAPS], g_ntaps[T2H_MAX_TAPS];
  int g_dyrel[T2H_MAX_TAPS][3], g_btap[T2H_MAX_TAPS][3];
  void* d;
  int d_mode, d_terms, epi_mode, d_term_imgs;
  long long d_plane, d_sn, d_sh, d_sw, d_sc;
  const float* bias;
  long long bias_sn;  // BIAS_COL: elements between the bias vectors of consecutive images (0 = shared)
  int bias_mode, act;
  float alpha;
  const float* residual;
  double* gn_stats;
  int gn_cpg, gn_groups;
  // conv weight-gradient mode (t2h_conv_wgrad): the contraction runs over 64-pixel patches (wg_PW x wg_PH) of the
  // images, both operands are NHWC planes read MN-major, the output "image" index is the tap whose (dy, dx,
  // img_off) shifts the X patch; accum: the epilogue reduce-adds into D even without split-K
  int wg, accum;
  int wg_pair, wg_ntaps;  // wg_pair: one 128 x 256 tile holds TWO taps side by side (Cin <= 128: N = 256 MMAs run at
                          // the full tensor rate, N = 128 at half of it); image index = tap pair
  // fused GroupNorm(+swish) activation producer (tapgemm_swap_kernel<MBLK, true>): fp32 NHWC source + statistics
  const float* ax;
  long long ax_sn, ax_sh, ax_sw;
  const double* ag_stats;
  const float *ag_gamma, *ag_beta;
  float ag_eps;
  int ag_swish, ag_groups, ag_hw, ag_H, ag_W;
  int partials;  // split-K without reduction: k-slice s of a tile is stored to image slot t.img + s of D
  // norm-backward sums in the swapped kernel's epilogue (see t2h_tapgemm_params.nb_sums): `residual` is x
  double* nb_sums;
  const double* nb_stats;
  const float *nb_gamma, *nb_beta;
  float nb_eps;
  int nb_act, nb_groups;
  int n_major;  // tile index = column tile * row tiles + row tile (set with nb_sums; default is row-major)
  int wg_PW, wg_PH, wg_pw, wg_ppi;
  int wg_dy[T2H_MAX_TAPS], wg_dx[T2H_MAX_TAPS], wg_ioff[T2H_MAX_TAPS];
};

// profiling trace (T2H_DEBUG bit 16): CTA 0 of every tap-GEMM / attention launch appends one record of 8 words --
// globaltimer (ns) at {kernel entry, griddepcontrol.wait passed, first operand tile landed, last MMA issued,
// accumulator complete (epilogue woke), CTA done}, then {total work items, contraction chunks per item | kind << 32}
constexpr int kTraceRecords = 2048;
__device__ unsigned long long g_t2h_dbg[kTraceRecords * 8];
__device__ unsigned int g_t2h_dbg_n;
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kBK = 64;                      // fp16 elements per 128-byte swizzled row
constexpr int kABlockBytes = 128 * 128;      // one 128-row A block of a stage
constexpr int kThreads = 384;  // warps 0-3 roles below, warps 4-7 and 8-11 two epilogue groups
constexpr int kEpiBufBytes = 128 * 128;      // one staging tile: 128 rows x 128 bytes
constexpr int kEpiBytes = 4 * kEpiBufBytes;  // 2 output + 2 residual staging tiles

constexpr int kMaxSlots = 8;
constexpr int kDynSmem = 227 * 1024 - 4096;  // opt-in limit minus ~4 KB of static shared memory
constexpr int kRingBytes = kDynSmem - 1024 - kEpiBytes;  // A ring + B ring

template <int BN, int MBLK>
struct Cfg {
  // A ring: slabs of (MBLK*TH + 2) x TW positions x 128 bytes (TW <= 16 whenever taps share a slab)
  static constexpr int kASlot = MBLK * kABlockBytes + 4096;
  static constexpr int kBSlot = BN * 128;
  static constexpr int kAccCols = MBLK * BN;
  static constexpr int kTmemCols = (2 * kAccCols) < 32 ? 32 : (2 * kAccCols);
  static constexpr int kChunk = BN < 32 ? BN : 32;  // columns per tcgen05.ld in the direct epilogue
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
  if (P.n_major) {  // column tile is the slow index: a contiguous tile range stays on one (column block, image)
    const int mt = P.total_tiles / P.n_tiles_n;
    n_tile = tile / mt;
    m_tile = tile - n_tile * mt;
  }
  if (P.pair) {
    // `tile` enumerates (row-tile pair, column tile); this CTA takes row 2*pair + rank.  An odd row count
    // leaves the last pair's second CTA with an out-of-range row tile (zero-filled loads, clipped stores).
    m_tile = 2 * m_tile + (int)(blockIdx.x & 1);
    if (m_tile >= P.n_img * P.tiles_h * P.tiles_w) {  // rows beyond the matrix: every access is out of range
      t.img = 0;
      t.h0 = 0;
      t.w0 = P.tiles_w * P.TW;
      t.n0 = n_tile * bn;
      return t;
    }
  }
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

// byte offset of 16-byte chunk j of row r inside a 128B-swizzled [128 rows][128 bytes] tile
__device__ __forceinline__ int swz(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }

// Per-warp reduction of 2*NG per-thread partial sums (NG groups x {sum, sumsq}) over the 32 rows
// a warp holds, using a halving butterfly (2*NG-1 + log shuffles), then shared-memory atomics.
template <int NG>
__device__ __forceinline__ void gn_accumulate(const float (&v)[32], bool row_ok, int ncols_valid,
                                              float* gs, int lane) {
  constexpr int W = 32 / NG;
  constexpr int NV = 2 * NG;
  float vals[NV];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int c = g * W + i;
      const float x = (row_ok && c < ncols_valid) ? v[c] : 0.f;
      s += x;
      ss += x * x;
    }
    vals[2 * g] = s;
    vals[2 * g + 1] = ss;
  }
  int idx = 0;
  int step = 0;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, ++step) {
    const int off = 1 << step;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? vals[i] : vals[i + n / 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      const float keep = upper ? vals[i + n / 2] : vals[i];
      vals[i] = keep + recv;
    }
    idx |= (upper ? 1 : 0) * (n / 2);
  }
#pragma unroll
  for (int off = NV; off < 32; off <<= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
  if (lane < NV) atomicAdd(&gs[idx], vals[0]);
}

template <int BN, int MBLK>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ TapGemmDev P) {
  using C = Cfg<BN, MBLK>;
  pdl_launch_dependents();  // the next kernel may start its prologue once every CTA of this one is running
  const int NA = P.a_slots, NB = P.b_slots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + NA * C::kASlot;
  uint8_t* out_buf = b_ring + NB * C::kBSlot;     // 2 tiles
  uint8_t* res_buf = out_buf + 2 * kEpiBufBytes;  // 2 tiles

  __shared__ __align__(8) uint64_t a_full[kMaxSlots];
  __shared__ __align__(8) uint64_t a_empty[kMaxSlots];
  __shared__ __align__(8) uint64_t b_full[kMaxSlots];
  __shared__ __align__(8) uint64_t b_empty[kMaxSlots];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t res_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float gsum[2][2 * 128];  // GroupNorm partial sums of the current / previous tile
  __shared__ __align__(16) float sbias[BN];  // column bias of the current tile (plane epilogue)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Two epilogue warp groups (warps 4-7: columns 0..31 of every 64-column unit, warps 8-11: columns 32..63) in the
  // modes whose epilogue is pure per-element work; the residual / GroupNorm-statistics / strided paths keep one.
  bool two_groups = false;
  if constexpr (BN >= 64) {
    two_groups = P.epi_mode == EPI_TMA_PLANES ||
                 (P.epi_mode == EPI_TMA_F32 && !P.residual && !P.gn_stats && P.act == T2H_ACT_NONE && !P.wg_pair &&
                  P.bias_mode != T2H_BIAS_ROW && !(P.debug & 1));
  }
  const int n_epi = two_groups ? 256 : 128;  // threads at the epilogue's named barriers
  __shared__ unsigned long long* trace_s;  // this launch's trace record (CTA 0, T2H_DEBUG bit 16), else null

  if (warp == 0 && lane == 0) {
    trace_s = nullptr;
    if ((P.debug & 16) && blockIdx.x == 0) {
      trace_s = g_t2h_dbg + (size_t)(atomicAdd(&g_t2h_dbg_n, 1u) % kTraceRecords) * 8;
      trace_s[0] = gtime_ns();
      trace_s[6] = (unsigned long long)P.total_work;
      trace_s[7] = (unsigned long long)min(P.kper, P.kchunks) * P.ngroups;
    }
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (P.epi_mode != EPI_DIRECT) tma_prefetch_desc(&tmD);
    if (P.epi_mode == EPI_TMA_F32 && P.residual) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], P.pair ? 2 : 1);  // pair: both CTAs' MMA warps release a (multicast) weight slot
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], two_groups ? 8 : 4);
      mbar_init(&res_bar[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, C::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * 2 * 128; i += kThreads) (&gsum[0][0])[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (P.pair) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  // everything above touched only shared / tensor memory and kernel parameters; from here on the previous
  // kernel's results are read (and buffers it may still be reading are overwritten)
  pdl_wait();
  unsigned long long* const trace = trace_s;
  if (trace && threadIdx.x == 0) trace[1] = gtime_ns();

  const int a_planes = (P.nterms == 3) ? 2 : 1;  // slabs per (group, chunk): hi [, lo]
  const int slab_bytes = P.slab_rows * P.TW * 128;
  // pair mode: the two CTAs of a cluster walk the same sequence of (row-tile pair, column tile) work items
  const int work0 = P.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_stride = P.pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t pair_rank = P.pair ? (blockIdx.x & 1u) : 0u;

  if (warp == 0) {
    // ---------------------------------------------- A producer (activation slabs)
    if (lane == 0) {
      int sa = 0, pa = 0;
      for (int work = work0; work < P.total_work; work += work_stride) {
        const int tile = work % P.total_tiles;
        const int ch0 = (work / P.total_tiles) * P.kper;
        const int ch1 = min(P.kchunks, ch0 + P.kper);
        const TileCoord t = decode_tile(P, tile, MBLK, BN);
        const int a_img = P.a_bcast ? 0 : t.img;
        for (int g = 0; g < P.ngroups; ++g) {
          for (int ch = ch0; ch < ch1; ++ch) {
            for (int pl = 0; pl < a_planes; ++pl) {  // hi, then lo
              mbar_wait(&a_empty[sa], pa ^ 1);
              if (P.debug & 4) {
                mbar_arrive(&a_full[sa]);
              } else {
                if (P.a_mn) {
                  // [k][row] storage: two boxes of 64 rows x 64 k per 128-row block
                  mbar_expect_tx(&a_full[sa], kABlockBytes);
                  int c1 = ch * kBK, c2 = t.h0, c3 = a_img + pl * P.a_term_imgs;
                  if (P.wg) {  // chunk = one 64-pixel patch of dY: (channel, x, y, image)
                    const int n = ch / P.wg_ppi, r = ch - n * P.wg_ppi, py = r / P.wg_pw, px = r - py * P.wg_pw;
                    c1 = px * P.wg_PW; c2 = py * P.wg_PH; c3 = n + pl * P.a_term_imgs;
                  }
#pragma unroll
                  for (int hbox = 0; hbox < 2; ++hbox)
                    tma_load_4d(&tmA, &a_full[sa], a_ring + sa * C::kASlot + hbox * 8192, t.w0 + 64 * hbox,
                                c1, c2, c3);
                } else {
                  mbar_expect_tx(&a_full[sa], slab_bytes);
                  tma_load_4d(&tmA, &a_full[sa], a_ring + sa * C::kASlot, ch * kBK, t.w0 + P.g_dx[g],
                              t.h0 + P.g_dy0[g], a_img + P.g_ioff[g] + pl * P.a_term_imgs);
                }
              }
              if (++sa == NA) {
                sa = 0;
                pa ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ---------------------------------------------- B producer (weight tiles)
    if (lane == 0) {
      int sb = 0, pb = 0;
      uint32_t bj = 0;  // running index of weight-tile loads (pair mode: even ones are rank 0's, odd ones rank 1's)
      const int b_planes = a_planes;
      for (int work = work0; work < P.total_work; work += work_stride) {
        const int tile = work % P.total_tiles;
        const int ch0 = (work / P.total_tiles) * P.kper;
        const int ch1 = min(P.kchunks, ch0 + P.kper);
        const TileCoord t = decode_tile(P, tile, MBLK, BN);
        const int b_g2 = P.b_batched ? t.img : 0;
        const int b_g = P.b_batched_h ? t.h0 : 0;
        for (int g = 0; g < P.ngroups; ++g) {
          for (int ch = ch0; ch < ch1; ++ch) {
            for (int tp = 0; tp < P.g_ntaps[g]; ++tp) {
              for (int pl = b_planes - 1; pl >= 0; --pl) {  // lo first, then hi (consumption order)
                mbar_wait(&b_empty[sb], pb ^ 1);
                if (P.debug & 4) {
                  mbar_arrive(&b_full[sb]);
                } else {
                  mbar_expect_tx(&b_full[sb], C::kBSlot);
                  if (P.pair) {
                    // weight tiles alternate between the two CTAs; the issuer multicasts to both
                    if ((bj++ & 1u) == pair_rank) {
                      if (P.b_mn) {
                        for (int q = 0; q < BN / 64; ++q)
                          tma_load_4d_mc(&tmB, &b_full[sb], b_ring + sb * C::kBSlot + q * 8192, t.n0 + 64 * q,
                                         ch * kBK, b_g + P.g_btap[g][tp] + pl * P.b_term_g, b_g2, 3);
                      } else {
                        tma_load_4d_mc(&tmB, &b_full[sb], b_ring + sb * C::kBSlot, ch * kBK, t.n0,
                                       b_g + P.g_btap[g][tp] + pl * P.b_term_g, b_g2, 3);
                      }
                    }
                  } else if (P.b_mn) {
                    // [k][column] storage: one box of 64 columns x 64 k per 64 output columns
                    int c1 = ch * kBK, c2 = b_g + P.g_btap[g][tp] + pl * P.b_term_g, c3 = b_g2;
                    if (P.wg && P.wg_pair) {
                      // two taps per tile: 64-column boxes 0,1 hold tap 2*img, boxes 2,3 tap 2*img + 1 (the last
                      // pair of an odd tap count re-reads the last tap; its columns are never stored)
                      const int n = ch / P.wg_ppi, r = ch - n * P.wg_ppi, py = r / P.wg_pw, px = r - py * P.wg_pw;
                      for (int q = 0; q < BN / 64; ++q) {
                        int tap = 2 * t.img + (q >> 1);
                        if (tap >= P.wg_ntaps) tap = P.wg_ntaps - 1;
                        tma_load_4d(&tmB, &b_full[sb], b_ring + sb * C::kBSlot + q * 8192, 64 * (q & 1),
                                    px * P.wg_PW + P.wg_dx[tap], py * P.wg_PH + P.wg_dy[tap],
                                    n + P.wg_ioff[tap] + pl * P.b_term_g);
                      }
                    } else {
                    if (P.wg) {  // the same patch of X, shifted by this output tile's tap
                      const int n = ch / P.wg_ppi, r = ch - n * P.wg_ppi, py = r / P.wg_pw, px = r - py * P.wg_pw;
                      c1 = px * P.wg_PW + P.wg_dx[t.img]; c2 = py * P.wg_PH + P.wg_dy[t.img];
                      c3 = n + P.wg_ioff[t.img] + pl * P.b_term_g;
                    }
                    for (int q = 0; q < BN / 64; ++q)
                      tma_load_4d(&tmB, &b_full[sb], b_ring + sb * C::kBSlot + q * 8192, t.n0 + 64 * q, c1, c2, c3);
                    }
                  } else {
                    tma_load_4d(&tmB, &b_full[sb], b_ring + sb * C::kBSlot, ch * kBK, t.n0,
                                b_g + P.g_btap[g][tp] + pl * P.b_term_g, b_g2);
                  }
                }
                if (++sb == NB) {
                  sb = 0;
                  pb ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer
    // One thread; per MMA it only adds to the two descriptor low words (see umma_ksteps), which keeps
    // the issue stream well under the ~120 cycles an MMA occupies the tensor pipe.
    {  // the whole warp runs the loop; one elected lane issues the tcgen05 instructions
      const uint32_t IDESC = umma_idesc_f16(128, BN) | (P.a_mn ? (1u << 15) : 0u) | (P.b_mn ? (1u << 16) : 0u);
      const uint32_t a_step = P.a_mn ? kUmmaStepMN : kUmmaStepK, b_step = P.b_mn ? kUmmaStepMN : kUmmaStepK;
      const int last_steps = (P.C - (P.kchunks - 1) * kBK + 15) / 16;  // UMMA_K=16 steps
      const uint32_t row16 = (uint32_t)(P.TW * 128) >> 4;  // one slab image row, in 16-byte units
      const uint32_t a_lo0 = P.a_mn ? umma_desc_lo_mn(smem_u32(a_ring)) : umma_desc_lo(smem_u32(a_ring));
      const uint32_t b_lo0 = P.b_mn ? umma_desc_lo_mn(smem_u32(b_ring)) : umma_desc_lo(smem_u32(b_ring));
      constexpr uint32_t A16 = C::kASlot >> 4, B16 = C::kBSlot >> 4, BLK16 = kABlockBytes >> 4;
      const bool dbg_nomma = (P.debug & 2) != 0;
      const int ngroups = P.ngroups, kchunks = P.kchunks;
      int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, ap = 0;
      for (int work = work0; work < P.total_work; work += work_stride) {
        const int tile = work % P.total_tiles;
        const int ch0 = (work / P.total_tiles) * P.kper;
        const int ch1 = min(P.kchunks, ch0 + P.kper);
        mbar_wait(&tempty_bar[as], ap ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * C::kAccCols;
        uint32_t acc[MBLK];
#pragma unroll
        for (int mb = 0; mb < MBLK; ++mb) acc[mb] = 0;
        // D[mb] (+)= A_view(mb, dy) * B_tile
        auto batch = [&](uint32_t a_lo, uint32_t b_lo, int ksteps) {
          if (dbg_nomma) return;
#pragma unroll
          for (int mb = 0; mb < MBLK; ++mb)
            { if (elect_one()) umma_ksteps(d_base + mb * BN, a_lo + mb * BLK16, b_lo, IDESC, ksteps, acc[mb], a_step, b_step); acc[mb] = 1; __syncwarp(); }
        };
        for (int g = 0; g < ngroups; ++g) {
          const int nt = P.g_ntaps[g];
          const uint32_t dy0 = P.g_dyrel[g][0] * row16, dy1 = P.g_dyrel[g][1] * row16,
                         dy2 = P.g_dyrel[g][2] * row16;
          for (int ch = ch0; ch < ch1; ++ch) {
            const int ksteps = (ch == kchunks - 1) ? last_steps : 4;
            // slab slots of this (group, chunk)
            const int sa_hi = sa, pa_hi = pa;
            if (++sa == NA) { sa = 0; pa ^= 1; }
            const int sa_lo = sa, pa_lo = pa;
            if (a_planes == 2) {
              if (++sa == NA) { sa = 0; pa ^= 1; }
            }
            mbar_wait(&a_full[sa_hi], pa_hi);
            tc_fence_after();
            if (trace && lane == 0 && work == work0 && g == 0 && ch == ch0) trace[2] = gtime_ns();
            const uint32_t ahi = a_lo0 + sa_hi * A16;
            const uint32_t alo = a_lo0 + sa_lo * A16;
            for (int tp = 0; tp < nt; ++tp) {
              const uint32_t dy = tp == 0 ? dy0 : (tp == 1 ? dy1 : dy2);
              if (a_planes == 1) {
                mbar_wait(&b_full[sb], pb);
                tc_fence_after();
                batch(ahi + dy, b_lo0 + sb * B16, ksteps);
                { if (elect_one()) { if (P.pair) umma_commit_mc(&b_empty[sb], 3); else umma_commit(&b_empty[sb]); } __syncwarp(); }
                if (++sb == NB) { sb = 0; pb ^= 1; }
              } else {
                mbar_wait(&b_full[sb], pb);  // B lo
                tc_fence_after();
                batch(ahi + dy, b_lo0 + sb * B16, ksteps);  // hi*lo
                { if (elect_one()) { if (P.pair) umma_commit_mc(&b_empty[sb], 3); else umma_commit(&b_empty[sb]); } __syncwarp(); }
                if (++sb == NB) { sb = 0; pb ^= 1; }
                mbar_wait(&b_full[sb], pb);  // B hi
                tc_fence_after();
                const uint32_t bhi = b_lo0 + sb * B16;
                batch(ahi + dy, bhi, ksteps);  // hi*hi
                if (tp == nt - 1) { if (elect_one()) umma_commit(&a_empty[sa_hi]); __syncwarp(); }  // hi slab done
                if (tp == 0) {
                  mbar_wait(&a_full[sa_lo], pa_lo);
                  tc_fence_after();
                }
                batch(alo + dy, bhi, ksteps);  // lo*hi
                { if (elect_one()) { if (P.pair) umma_commit_mc(&b_empty[sb], 3); else umma_commit(&b_empty[sb]); } __syncwarp(); }
                if (++sb == NB) { sb = 0; pb ^= 1; }
              }
            }
            { if (elect_one()) umma_commit(a_planes == 1 ? &a_empty[sa_hi] : &a_empty[sa_lo]); __syncwarp(); }
          }
        }
        { if (elect_one()) umma_commit(&tfull_bar[as]); __syncwarp(); }  // accumulator complete
        if (trace && lane == 0 && work == work0) trace[3] = gtime_ns();
        if (++as == 2) {
          as = 0;
          ap ^= 1;
        }
      }
    }
  } else if (warp >= 4 && (warp < 8 || two_groups)) {
    // ------------------------------------------------------------ epilogue
    const int grp = (warp - 4) >> 2;  // 0: warps 4-7, 1: warps 8-11
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int th = row / P.TW, tw = row - th * P.TW;
    const bool elected = (threadIdx.x == 128);
    int as = 0, ap = 0;
    int buf = 0;  // staging buffer toggle (persists across tiles)
    uint32_t res_par[2] = {0, 0};
    int tile_par = 0;  // gsum buffer of this tile

    for (int work = work0; work < P.total_work; work += work_stride) {
        const int tile = work % P.total_tiles;
        const int ch0 = (work / P.total_tiles) * P.kper;
        const int ch1 = min(P.kchunks, ch0 + P.kper);
      const TileCoord t = decode_tile(P, tile, MBLK, BN);

      const bool plain_f32 = two_groups && P.epi_mode == EPI_TMA_F32;
      if (plain_f32) {
        // ---- fp32 output with nothing but alpha / column bias in the epilogue (split-K slices, weight gradients,
        // plain projections): 64-column units = two 32-column staging tiles, two pairs of them, so that a unit is
        // converted while the previous unit's TMA stores still read theirs -- half the barrier rounds of the
        // general path below
        const int cols_left = P.n_out - t.n0;
        const int nuc = (cols_left >= BN) ? BN / 64 : (cols_left + 63) / 64;
        const int nunits = MBLK * nuc;
        const bool add_bias = P.bias_mode == T2H_BIAS_COL && ch0 == 0;  // split-K: the first k-slice carries the bias
        if (add_bias) {
          for (int i = threadIdx.x - 128; i < BN; i += 256)
            sbias[i] = (t.n0 + i < P.n_out) ? __ldg(P.bias + t.img * P.bias_sn + t.n0 + i) : 0.f;
        }
        mbar_wait(&tfull_bar[as], ap);
        tc_fence_after();
        if (trace && elected && work == work0) trace[4] = gtime_ns();
        for (int u = 0; u < nunits; ++u) {
          const int mb = u / nuc, cc = u - mb * nuc;
          const int col0 = t.n0 + cc * 64;
          uint8_t* const o0 = buf ? res_buf : out_buf;
          named_bar_sync(1, n_epi);  // this pair is free (elected waited for the stores issued two units ago)
          {
            const int half = grp;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * C::kAccCols + mb * BN + cc * 64 + half * 32, r);
            tmem_ld_wait();
            if (u == nunits - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[as]);
            }
            uint8_t* const ob = o0 + half * kEpiBufBytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(__uint_as_float(r[4 * j]) * P.alpha, __uint_as_float(r[4 * j + 1]) * P.alpha,
                                     __uint_as_float(r[4 * j + 2]) * P.alpha, __uint_as_float(r[4 * j + 3]) * P.alpha);
              if (add_bias) {
                const float4 b4 = *reinterpret_cast<const float4*>(sbias + cc * 64 + half * 32 + 4 * j);
                o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
              }
              *reinterpret_cast<float4*>(ob + swz(row, j)) = o;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(2, n_epi);
          if (elected) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c0 = col0 + half * 32;
              if (c0 >= P.n_out) break;
              const uint8_t* ob = o0 + half * kEpiBufBytes;
              if (P.partials)  // deterministic split-K: every k-slice owns a slab; a fixed-order pass sums them
                tma_store_4d(&tmD, ob, c0, t.w0, t.h0 + mb * P.TH, t.img + work / P.total_tiles);
              else if (P.ksplit > 1 || P.accum)
                tma_reduce_add_4d(&tmD, ob, c0, t.w0, t.h0 + mb * P.TH, t.img);
              else
                tma_store_4d(&tmD, ob, c0, t.w0, t.h0 + mb * P.TH, t.img);
            }
            tma_store_commit();
            tma_store_wait_read<1>();  // the other pair is free again
          }
          buf ^= 1;
        }
      } else if (P.epi_mode == EPI_TMA_F32) {
        // ---- fp32 NHWC output: 32-column units through swizzled smem + TMA store
        const int cols_left = P.n_out - t.n0;
        int nuc = (cols_left >= BN) ? BN / 32 : (cols_left + 31) / 32;
        if (P.wg_pair && 2 * t.img + 1 >= P.wg_ntaps) nuc = BN / 64;  // odd tap count: the last tile's upper half is unused
        const int nunits = MBLK * nuc;
        const bool has_res = P.residual != nullptr;
        auto issue_res = [&](int u, int b) {
          const int mb = u / nuc, cc = u - mb * nuc;
          mbar_expect_tx(&res_bar[b], kEpiBufBytes);
          tma_load_4d(&tmR, &res_bar[b], res_buf + b * kEpiBufBytes, t.n0 + cc * 32, t.w0,
                      t.h0 + mb * P.TH, t.img);
        };
        if (has_res && elected) {
          issue_res(0, buf);
          if (nunits > 1) issue_res(1, buf ^ 1);
        }
        mbar_wait(&tfull_bar[as], ap);
        tc_fence_after();
        if (trace && elected && work == work0) trace[4] = gtime_ns();
        for (int u = 0; u < nunits; ++u) {
          const int mb = u / nuc, cc = u - mb * nuc;
          const int col0 = t.n0 + cc * 32;
          const int h = t.h0 + mb * P.TH + th;
          const int w = t.w0 + tw;
          const bool row_ok = (h < P.H) && (w < P.W);
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * C::kAccCols + mb * BN + cc * 32, r);
          tmem_ld_wait();
          if (u == nunits - 1) {
            // every TMEM read of this accumulator is done: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
          }
          if (P.debug & 1) {
            if (has_res) {
              mbar_wait(&res_bar[buf], res_par[buf]);
              res_par[buf] ^= 1;
              named_bar_sync(2, 128);
              if (elected && u + 2 < nunits) issue_res(u + 2, buf);
            }
            buf ^= 1;
            continue;
          }
          float v[32];
          const float row_bias =
              (P.bias_mode == T2H_BIAS_ROW && row_ok) ? __ldg(P.bias + h * P.W + w) : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * P.alpha + row_bias;
          if (P.bias_mode == T2H_BIAS_COL && ch0 == 0) {  // split-K: the first k-slice carries the bias
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (col0 + i < P.n_out) {  // n_out % 4 == 0 in this mode
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.bias + t.img * P.bias_sn + col0 + i));
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            }
          }
          if (P.act == T2H_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          } else if (P.act == T2H_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (P.act == T2H_ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
          }
          if (has_res) {
            mbar_wait(&res_bar[buf], res_par[buf]);
            res_par[buf] ^= 1;
            const uint8_t* rb = res_buf + buf * kEpiBufBytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 rr = *reinterpret_cast<const float4*>(rb + swz(row, j));
              v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
            }
          }
          if (P.gn_stats) {
            float* gs = &gsum[tile_par][0];
            const int ncv = P.n_out - col0;
            switch (P.gn_cpg) {
              case 2: gn_accumulate<16>(v, row_ok, ncv, gs + 2 * (cc * 16), lane); break;
              case 4: gn_accumulate<8>(v, row_ok, ncv, gs + 2 * (cc * 8), lane); break;
              case 8: gn_accumulate<4>(v, row_ok, ncv, gs + 2 * (cc * 4), lane); break;
              case 16: gn_accumulate<2>(v, row_ok, ncv, gs + 2 * (cc * 2), lane); break;
              default: gn_accumulate<1>(v, row_ok, ncv, gs + 2 * ((cc * 32) / P.gn_cpg), lane); break;
            }
          }
          named_bar_sync(1, 128);  // out_buf[buf] is free (elected waited for its previous store)
          uint8_t* ob = out_buf + buf * kEpiBufBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(ob + swz(row, j)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          named_bar_sync(2, 128);  // staging tile complete; residual tile fully consumed
          if (elected) {
            if (P.wg_pair)   // unit cc of a two-tap tile: columns (cc & 3) * 32 of tap 2*img + (cc >> 2)
              tma_reduce_add_4d(&tmD, ob, (cc & (BN / 64 - 1)) * 32, t.w0, t.h0 + mb * P.TH, 2 * t.img + cc / (BN / 64));
            else if (P.partials)  // deterministic split-K: every k-slice owns a slab; a fixed-order pass sums them
              tma_store_4d(&tmD, ob, col0, t.w0, t.h0 + mb * P.TH, t.img + work / P.total_tiles);
            else if (P.ksplit > 1 || P.accum)
              tma_reduce_add_4d(&tmD, ob, col0, t.w0, t.h0 + mb * P.TH, t.img);  // partial sum of a k-slice
            else
              tma_store_4d(&tmD, ob, col0, t.w0, t.h0 + mb * P.TH, t.img);
            tma_store_commit();
            if (has_res && u + 2 < nunits) issue_res(u + 2, buf);
            tma_store_wait_read<1>();  // the other staging tile is free again
          }
          buf ^= 1;
        }
        if (P.gn_stats) {
          // all shared atomics of this tile happened before the last named barrier
          const int et = threadIdx.x - 128;
          const int ngt = (BN + P.gn_cpg - 1) / P.gn_cpg;  // groups this tile can touch
          for (int sl = et; sl < 2 * ngt; sl += 128) {
            const int g = t.n0 / P.gn_cpg + (sl >> 1);
            if (g < P.gn_groups)
              atomicAdd(&P.gn_stats[((long long)t.img * P.gn_groups + g) * 2 + (sl & 1)],
                        (double)gsum[tile_par][sl]);
            gsum[tile_par][sl] = 0.f;
          }
          tile_par ^= 1;
        }
      } else if (P.epi_mode == EPI_TMA_PLANES) {
        // ---- fp16 plane output: 64-column units, hi and lo tiles staged side by side
        const int cols_left = P.n_out - t.n0;
        const int nuc = (cols_left >= BN) ? BN / 64 : (cols_left + 63) / 64;
        const int nunits = MBLK * nuc;
        // the tile's column bias goes to shared memory while the contraction is still running (its L2 round trips
        // used to sit between every tcgen05.ld and the stores); visible after the first named barrier below
        // (single buffer: whoever gets here has passed the previous tile's last named barrier, which every thread
        // reaches only after its last read of the previous bias)
        float* const sb = sbias;
        if (P.bias_mode == T2H_BIAS_COL) {
          for (int i = threadIdx.x - 128; i < BN; i += n_epi)
            sb[i] = (t.n0 + i < P.n_out) ? __ldg(P.bias + t.img * P.bias_sn + t.n0 + i) : 0.f;
        }
        mbar_wait(&tfull_bar[as], ap);
        tc_fence_after();
        if (trace && elected && work == work0) trace[4] = gtime_ns();
        for (int u = 0; u < nunits; ++u) {
          const int mb = u / nuc, cc = u - mb * nuc;
          const int col0 = t.n0 + cc * 64;
          const int h = t.h0 + mb * P.TH + th;
          const int w = t.w0 + tw;
          const bool row_ok = (h < P.H) && (w < P.W);
          const float row_bias =
              (P.bias_mode == T2H_BIAS_ROW && row_ok) ? __ldg(P.bias + h * P.W + w) : 0.f;
          // two pairs of staging tiles (the residual tiles are unused in this mode): unit u + 1 is converted while
          // unit u's TMA stores still read theirs
          uint8_t* ohi = buf ? res_buf : out_buf;
          uint8_t* olo = ohi + kEpiBufBytes;
          named_bar_sync(1, n_epi);  // this pair is free (elected waited for the stores issued two units ago)
#pragma unroll
          for (int half = two_groups ? grp : 0; half < (two_groups ? grp + 1 : 2); ++half) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * C::kAccCols + mb * BN +
                                   cc * 64 + half * 32;
            tmem_ld_32x32(taddr, r);
            tmem_ld_wait();
            if (u == nunits - 1 && (two_groups || half == 1)) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[as]);
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * P.alpha + row_bias;
            if (P.bias_mode == T2H_BIAS_COL) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sb + cc * 64 + half * 32 + i);  // broadcast read
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            }
            if (P.act == T2H_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            } else if (P.act == T2H_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (P.act == T2H_ACT_LRELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __align__(16) __half hi[8];
              __align__(16) __half lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) split_f16(v[8 * j + e], hi[e], lo[e]);
              *reinterpret_cast<uint4*>(ohi + swz(row, half * 4 + j)) = *reinterpret_cast<uint4*>(hi);
              if (P.d_terms == 2)
                *reinterpret_cast<uint4*>(olo + swz(row, half * 4 + j)) = *reinterpret_cast<uint4*>(lo);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(2, n_epi);
          if (elected) {
            tma_store_4d(&tmD, ohi, col0, t.w0, t.h0 + mb * P.TH, t.img);
            if (P.d_terms == 2)
              tma_store_4d(&tmD, olo, col0, t.w0, t.h0 + mb * P.TH, t.img + P.d_term_imgs);
            tma_store_commit();
            tma_store_wait_read<1>();  // the other pair of staging tiles is free again
          }
          buf ^= 1;
        }
      } else {
        // ---- direct path: strided / tiny outputs (NCHW conv_out, n_out < 32, unaligned)
        constexpr int CH = C::kChunk;
        mbar_wait(&tfull_bar[as], ap);
        tc_fence_after();
        if (trace && elected && work == work0) trace[4] = gtime_ns();
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
                if (col0 + i < P.n_out) v[i] += __ldg(P.bias + t.img * P.bias_sn + col0 + i);
            }
            if (P.act == T2H_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] = gelu_erf(v[i]);
            } else if (P.act == T2H_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (P.act == T2H_ACT_LRELU) {
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
            }
            if (P.d_mode == T2H_OUT_F32) {
              float* dst = reinterpret_cast<float*>(P.d);
#pragma unroll
              for (int i = 0; i < CH; ++i) {
                if (col0 + i < P.n_out) {
                  const long long o = off + (long long)(col0 + i) * P.d_sc;
                  float val = v[i];
                  if (P.residual) val += P.residual[o];
                  dst[o] = val;
                }
              }
            } else {
              __half* dst = reinterpret_cast<__half*>(P.d);
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
      if (++as == 2) {
        as = 0;
        ap ^= 1;
      }
    }
    if (elected) tma_store_wait_read<0>();
    if (trace && elected) trace[5] = gtime_ns();
  }

  tc_fence_before();
  __syncthreads();
  if (P.pair) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace t2h

#include "gemm_tc_swap.cuh"

namespace t2h {

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

// elem_bytes: 2 (fp16) or 4 (fp32); strides in elements
static int make_tmap(CUtensorMap* tm, const void* base, int elem_bytes, int rank, const uint64_t* dims,
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
      gstr[i - 1] = strides_elems[i] * elem_bytes;
      if (gstr[i - 1] % 16 != 0)
        return fail(T2H_EINVAL, "%s: stride of dim %d (%llu bytes) is not a multiple of 16", what, i,
                    (unsigned long long)gstr[i - 1]);
    }
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0)
    return fail(T2H_EINVAL, "%s: base pointer not 16-byte aligned", what);
  CUresult r = enc(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(T2H_ECUDA,
                "%s: cuTensorMapEncodeTiled failed (%d) dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
                what, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return T2H_OK;
}

template <int BN, int MBLK>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                  const CUtensorMap& tmR, const TapGemmDev& P, cudaStream_t stream) {
  using C = Cfg<BN, MBLK>;
  static bool configured[64] = {};
  int dev = 0;
  T2H_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {  // per device: the attribute lives in the device's copy of the function
    T2H_CUDA(cudaFuncSetAttribute(tapgemm_kernel<BN, MBLK>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
    configured[dev & 63] = true;
  }
  // ring depths: enough bytes in flight to cover the TMA round trip at the tile's consumption rate.
  // With the 3-product split both (hi, lo) slabs of a (group, chunk) are live at once.
  TapGemmDev Q = P;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("T2H_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    Q.debug = dbg;
  }
  Q.a_slots = (MBLK == 2 || BN == 256) ? 3 : 4;
  int nb = (kRingBytes - Q.a_slots * C::kASlot) / C::kBSlot;
  Q.b_slots = nb > kMaxSlots ? kMaxSlots : nb;
  if (Q.b_slots < 3) return fail(T2H_EINVAL, "tapgemm: shared-memory rings do not fit");
  int grid = P.total_work < num_sms() ? P.total_work : num_sms();
  if (P.pair) {
    grid = 2 * P.total_work;
    if (grid > (num_sms() & ~1)) grid = num_sms() & ~1;
  }
  // the full dynamic allocation also keeps it to one CTA (one TMEM allocation) per SM
  T2H_CUDA(launch_pdl(tapgemm_kernel<BN, MBLK>, dim3(grid), dim3(kThreads), kDynSmem, stream, P.pair ? 2 : 1, tmA,
                      tmB, tmD, tmR, Q));
  return T2H_OK;
}

template <int MBLK, bool FUSE>
static int launch_swap(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                       const CUtensorMap& tmR, const TapGemmDev& P, cudaStream_t stream) {
  using C = Cfg<128, MBLK>;
  static bool configured[64] = {};
  int dev = 0;
  T2H_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    T2H_CUDA(cudaFuncSetAttribute(tapgemm_swap_kernel<MBLK, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kDynSmem));
    configured[dev & 63] = true;
  }
  TapGemmDev Q = P;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("T2H_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    Q.debug = dbg;
  }
  Q.a_slots = (MBLK == 2) ? 3 : 4;
  int nb = (kRingBytes - Q.a_slots * C::kASlot) / C::kBSlot;
  Q.b_slots = nb > kMaxSlots ? kMaxSlots : nb;
  int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  T2H_CUDA(launch_pdl(tapgemm_swap_kernel<MBLK, FUSE>, dim3(grid), dim3(kSwapThreads), kDynSmem, stream, 1, tmA, tmB,
                      tmD, tmR, Q));
  return T2H_OK;
}

}  // namespace t2h

using namespace t2h;

extern "C" int t2h_debug_read(long long* out, int n) {
  // out[0] = records written so far (the ring keeps the last kTraceRecords), then up to (n - 1) / 8 records
  T2H_CHECK_ARG(out && n >= 1, "debug_read: bad args");
  unsigned int cnt = 0;
  T2H_CUDA(cudaDeviceSynchronize());
  T2H_CUDA(cudaMemcpyFromSymbol(&cnt, g_t2h_dbg_n, sizeof(cnt)));
  out[0] = cnt;
  int words = n - 1;
  if (words > kTraceRecords * 8) words = kTraceRecords * 8;
  if (words > 0) T2H_CUDA(cudaMemcpyFromSymbol(out + 1, g_t2h_dbg, sizeof(long long) * words));
  return T2H_OK;
}

extern "C" int t2h_tapgemm(const t2h_tapgemm_params* p, t2h_stream_t stream) {
  T2H_CHECK_ARG(p && (p->a || p->a_f32) && p->b && p->d, "tapgemm: null operand");
  T2H_CHECK_ARG(p->n_img > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->n_out > 0,
                "tapgemm: empty problem (n_img=%d H=%d W=%d C=%d n_out=%d)", p->n_img, p->H, p->W, p->C,
                p->n_out);
  T2H_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= T2H_MAX_TAPS, "tapgemm: ntaps=%d", p->ntaps);
  T2H_CHECK_ARG(p->nterms == 1 || p->nterms == 3, "tapgemm: nterms must be 1 or 3 (got %d)", p->nterms);
  T2H_CHECK_ARG(p->nterms == 1 || ((p->a_terms == 2 || p->a_f32) && p->b_terms == 2),
                "tapgemm: nterms=3 needs hi/lo planes on both operands");
  T2H_CHECK_ARG(p->d_mode == T2H_OUT_F32 || p->d_mode == T2H_OUT_PLANES, "tapgemm: d_mode=%d", p->d_mode);
  T2H_CHECK_ARG(p->d_mode == T2H_OUT_F32 || p->d_terms == 1 || p->d_terms == 2, "tapgemm: d_terms=%d",
                p->d_terms);
  T2H_CHECK_ARG(p->d_mode == T2H_OUT_F32 || p->residual == nullptr,
                "tapgemm: residual add needs fp32 output");
  T2H_CHECK_ARG(p->bias_mode == T2H_BIAS_NONE || p->bias != nullptr, "tapgemm: bias_mode without bias");
  T2H_CHECK_ARG(!p->b_batched_h || p->H == 1 || p->tile_rows, "tapgemm: b_batched_h needs row tiles");

  TapGemmDev P;
  memset(&P, 0, sizeof(P));
  P.n_img = p->n_img; P.H = p->H; P.W = p->W;
  P.n_out = p->n_out; P.C = p->C; P.kchunks = (p->C + kBK - 1) / kBK; P.nterms = p->nterms;
  P.a_term_imgs = p->a_term_imgs; P.a_bcast = p->a_bcast;
  P.a_mn = p->a_mn ? 1 : 0; P.b_mn = p->b_mn ? 1 : 0;
  P.b_term_g = p->b_term_g; P.b_batched = p->b_batched; P.b_batched_h = p->b_batched_h;
  P.d = p->d; P.d_mode = p->d_mode; P.d_terms = p->d_terms; P.d_plane = p->d_plane;
  P.d_sn = p->d_sn; P.d_sh = p->d_sh; P.d_sw = p->d_sw; P.d_sc = p->d_sc;
  P.bias = p->bias; P.bias_sn = p->bias_mode == T2H_BIAS_COL ? p->bias_sn : 0; P.bias_mode = p->bias_mode; P.act = p->act; P.alpha = p->alpha;
  P.residual = p->residual; P.gn_stats = p->gn_stats; P.gn_cpg = p->gn_cpg;
  P.nb_sums = p->nb_sums; P.nb_stats = p->nb_stats; P.nb_gamma = p->nb_gamma; P.nb_beta = p->nb_beta;
  P.nb_eps = p->nb_eps; P.nb_act = p->nb_act; P.nb_groups = p->nb_groups;
  P.n_major = p->nb_sums ? 1 : 0;
  P.gn_groups = p->gn_cpg > 0 ? p->n_out / p->gn_cpg : 0;
  P.d_term_imgs = 0;

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
  // Swapped-operand kernel (gemm_tc_swap.cuh): spatial convs with Cout % 128 == 0 and a plain fp32
  // NHWC destination.  The pixel tile is the UMMA N operand (256 wide = full tensor rate).
  const bool rows_mode = (p->H == 1 || p->tile_rows);
  bool swap = !rows_mode && TW <= 16 && p->n_out % 128 == 0 && p->d_mode == T2H_OUT_F32 && p->d_sc == 1 &&
              p->bias_mode != T2H_BIAS_ROW && !p->a_bcast && !p->b_batched && !p->b_batched_h &&
              p->d_sw % 4 == 0 && p->d_sh % 4 == 0 && p->d_sh > 0 && (p->n_img == 1 || (p->d_sn % 4 == 0 && p->d_sn > 0)) &&
              reinterpret_cast<uintptr_t>(p->d) % 16 == 0 &&
              (!p->residual || reinterpret_cast<uintptr_t>(p->residual) % 16 == 0) &&
              (!p->gn_stats || (p->gn_cpg >= 1 && (p->gn_cpg & (p->gn_cpg - 1)) == 0 && p->n_out % p->gn_cpg == 0));
  // ... and with a direct (strided) epilogue for small-Cout convs writing NCHW (conv_out): the weight
  // tile is zero-padded to 128 rows by TMA, the 256-pixel N operand still issues at full rate.
  const bool swap_direct = !swap && !rows_mode && TW <= 16 && p->n_out <= 128 && p->d_mode == T2H_OUT_F32 &&
                           p->d_sc != 1 && p->bias_mode != T2H_BIAS_ROW && !p->a_bcast && !p->b_batched &&
                           !p->b_batched_h && !p->residual && !p->gn_stats && p->act == T2H_ACT_NONE;
  if (swap_direct) swap = true;
  {
    static int no_swap = -1;
    if (no_swap < 0) {
      const char* e = getenv("T2H_NO_SWAP");
      no_swap = e ? atoi(e) : 0;
    }
    if (no_swap) swap = false;
  }
  if (p->a_f32) {
    // fused GroupNorm(+swish) producer: only in the swapped kernel, whole 64-channel chunks, table of <= 256 channels
    bool ok = swap && p->C % 64 == 0 && p->C <= 256 && p->a_gn_stats && p->a_gn_gamma && p->a_gn_beta &&
              p->a_gn_groups > 0 && p->C % p->a_gn_groups == 0 && !p->a_bcast &&
              reinterpret_cast<uintptr_t>(p->a_f32) % 16 == 0 && p->a_sw % 4 == 0 && p->a_sh % 4 == 0 && p->a_sn % 4 == 0;
    for (int i = 0; i < p->ntaps; ++i) ok = ok && p->tap_img_off[i] == 0;
    T2H_CHECK_ARG(ok, "tapgemm: the fused GroupNorm producer needs a swapped-kernel conv (Cout %% 128 == 0 or a strided "
                      "small-Cout output), C %% 64 == 0, C <= 256 and aligned fp32 NHWC input (C=%d n_out=%d)", p->C, p->n_out);
    P.ax = p->a_f32; P.ax_sn = p->a_sn; P.ax_sh = p->a_sh; P.ax_sw = p->a_sw;
    P.ag_stats = p->a_gn_stats; P.ag_gamma = p->a_gn_gamma; P.ag_beta = p->a_gn_beta; P.ag_eps = p->a_gn_eps;
    P.ag_swish = p->a_gn_swish; P.ag_groups = p->a_gn_groups; P.ag_hw = p->a_H * p->a_W;
    P.ag_H = p->a_H; P.ag_W = p->a_W;
  }
  if (p->a_mn || p->b_mn) {
    T2H_CHECK_ARG(rows_mode && p->ntaps == 1 && p->tap_dx[0] == 0 && p->tap_dy[0] == 0,
                  "tapgemm: a_mn / b_mn operands need a row GEMM (H == 1 or tile_rows) with one tap");
    swap = false;
  }
  int BN = 16;
  while (BN < p->n_out && BN < 256) BN <<= 1;
  if (p->b_mn && BN < 64) BN = 64;  // an MN-major B tile is made of whole 64-column boxes
  int MBLK = 1;
  if (swap) {
    BN = 128;
    const long long tiles2 =
        (long long)p->n_img * ceil_div(p->H, 2 * TH) * ceil_div(p->W, TW) * ceil_div(p->n_out, 128);
    MBLK = (p->H >= 2 * TH && tiles2 >= 48) ? 2 : 1;
  } else if (BN == 128 && p->H >= 2 * TH && !p->tile_rows) {
    long long tiles1 = (long long)p->n_img * ceil_div(p->H, TH) * ceil_div(p->W, TW);
    if (tiles1 >= 4LL * num_sms()) MBLK = 2;
  }
  P.TW = TW; P.TH = TH;
  // ---- tap groups: taps with the same (dx, img_off) share one activation slab when a vertical
  // shift of the slab view stays 1024-byte aligned (TW*128 bytes per image row, TW >= 8)
  {
    const bool can_share = (TW >= 8) && (TW <= 16);
    P.ngroups = 0;
    int gmin[T2H_MAX_TAPS], gmax[T2H_MAX_TAPS];
    int gt_dy[T2H_MAX_TAPS][3];
    for (int i = 0; i < p->ntaps; ++i) {
      int g = -1;
      if (can_share)
        for (int k = 0; k < P.ngroups; ++k)
          if (P.g_dx[k] == p->tap_dx[i] && P.g_ioff[k] == p->tap_img_off[i] && P.g_ntaps[k] < 3) {
            const int lo = p->tap_dy[i] < gmin[k] ? p->tap_dy[i] : gmin[k];
            const int hi = p->tap_dy[i] > gmax[k] ? p->tap_dy[i] : gmax[k];
            if (hi - lo <= 2) { g = k; gmin[k] = lo; gmax[k] = hi; break; }
          }
      if (g < 0) {
        g = P.ngroups++;
        P.g_dx[g] = p->tap_dx[i]; P.g_ioff[g] = p->tap_img_off[i]; P.g_ntaps[g] = 0;
        gmin[g] = gmax[g] = p->tap_dy[i];
      }
      gt_dy[g][P.g_ntaps[g]] = p->tap_dy[i];
      P.g_btap[g][P.g_ntaps[g]] = p->use_tap_w ? p->tap_w[i] : i;
      P.g_ntaps[g]++;
    }
    int extra = 0;
    for (int g = 0; g < P.ngroups; ++g) {
      P.g_dy0[g] = gmin[g];
      for (int k = 0; k < P.g_ntaps[g]; ++k) P.g_dyrel[g][k] = gt_dy[g][k] - gmin[g];
      if (gmax[g] - gmin[g] > extra) extra = gmax[g] - gmin[g];
    }
    for (int g = P.ngroups; g < T2H_MAX_TAPS; ++g) {
      P.g_dx[g] = P.g_ioff[g] = P.g_dy0[g] = P.g_ntaps[g] = 0;
    }
    P.slab_rows = MBLK * TH + extra;
  }
  P.tiles_w = ceil_div(p->W, TW);
  P.tiles_h = ceil_div(p->H, TH * MBLK);
  P.n_tiles_n = ceil_div(p->n_out, BN);
  long long total = (long long)p->n_img * P.tiles_w * P.tiles_h * P.n_tiles_n;
  T2H_CHECK_ARG(total < (1LL << 31), "tapgemm: too many tiles");
  P.total_tiles = (int)total;

  // ---- epilogue mode
  const int esz = (p->d_mode == T2H_OUT_F32) ? 4 : 2;
  const int align_el = 16 / esz;  // elements per 16 bytes
  bool tma_ok = p->d_sc == 1 && p->n_out % (p->d_mode == T2H_OUT_F32 ? 4 : 8) == 0 &&
                p->d_sw % align_el == 0 && (reinterpret_cast<uintptr_t>(p->d) % 16 == 0) &&
                BN >= (p->d_mode == T2H_OUT_F32 ? 32 : 64);
  // strides of the dims the output domain really uses must be TMA-encodable
  if (p->H > 1) tma_ok = tma_ok && p->d_sh % align_el == 0 && p->d_sh > 0;
  if (p->n_img > 1) tma_ok = tma_ok && p->d_sn % align_el == 0 && p->d_sn > 0;
  if (p->d_mode == T2H_OUT_PLANES && p->d_terms == 2) {
    // the lo plane is addressed as extra images: its distance must be a whole number of d_sn
    const long long sn = (p->n_img > 1) ? p->d_sn : p->d_plane;
    tma_ok = tma_ok && sn > 0 && p->d_plane % sn == 0 && p->d_plane % align_el == 0;
  }
  if (p->residual) tma_ok = tma_ok && (reinterpret_cast<uintptr_t>(p->residual) % 16 == 0);
  if (p->bias_mode == T2H_BIAS_COL)
    tma_ok = tma_ok && (reinterpret_cast<uintptr_t>(p->bias) % 16 == 0) && p->bias_sn % 4 == 0;
  P.epi_mode = !tma_ok ? EPI_DIRECT : (p->d_mode == T2H_OUT_F32 ? EPI_TMA_F32 : EPI_TMA_PLANES);
  if (swap) P.epi_mode = swap_direct ? EPI_DIRECT : EPI_TMA_F32;
  if (p->nb_sums) {
    T2H_CHECK_ARG(swap && !swap_direct && p->residual && p->nb_stats && p->nb_gamma && p->nb_beta && !p->gn_stats &&
                      !p->a_f32 && p->act == T2H_ACT_NONE && p->nb_act >= 0 && p->nb_act <= 2 && p->nb_groups > 0 &&
                      p->n_out % p->nb_groups == 0 && p->k_split <= 1 && !p->accumulate && !p->k_partials,
                  "tapgemm: nb_sums needs a swapped-kernel conv (n_out %% 128 == 0, fp32 NHWC output), x in `residual`, "
                  "statistics / gamma / beta, and no other epilogue work");
  }
  // ---- split-K: k-slices of one tile go to different CTAs and are reduce-added into a zeroed output
  P.ksplit = 1;
  P.kper = P.kchunks;
  if (p->k_partials > 1) {
    T2H_CHECK_ARG(!swap && P.epi_mode == EPI_TMA_F32 && !p->residual && p->bias_mode == T2H_BIAS_NONE &&
                      p->act == T2H_ACT_NONE && !p->gn_stats && p->n_img == 1 && p->k_split <= 1 &&
                      p->d_slab > 0 && p->d_slab % 4 == 0,
                  "tapgemm: k_partials needs a plain single-image fp32 GEMM output and an aligned slab stride");
    int ks = p->k_partials < P.kchunks ? p->k_partials : P.kchunks;
    P.kper = ceil_div(P.kchunks, ks);
    P.ksplit = ceil_div(P.kchunks, P.kper);
    P.partials = 1;
  } else if (p->k_split > 1 && !swap && P.epi_mode == EPI_TMA_F32 && !p->residual && p->bias_mode != T2H_BIAS_ROW &&
      p->act == T2H_ACT_NONE && !p->gn_stats) {
    int ks = p->k_split < P.kchunks ? p->k_split : P.kchunks;
    P.kper = ceil_div(P.kchunks, ks);
    P.ksplit = ceil_div(P.kchunks, P.kper);
  } else {
    T2H_CHECK_ARG(p->k_split <= 1, "tapgemm: k_split needs an aligned fp32 output and no row bias/act/residual");
  }
  T2H_CHECK_ARG((long long)P.total_tiles * P.ksplit < (1LL << 31), "tapgemm: too many work items");
  P.total_work = P.total_tiles * P.ksplit;
  if (p->accumulate) {
    T2H_CHECK_ARG(!swap && P.epi_mode == EPI_TMA_F32 && !p->residual && p->bias_mode != T2H_BIAS_ROW &&
                      p->act == T2H_ACT_NONE && !p->gn_stats,
                  "tapgemm: accumulate needs an aligned fp32 output and no row bias/act/residual");
    P.accum = 1;
  }
  // ---- CTA pairs: plain row GEMMs with at least two row tiles run as clusters of two CTAs on vertically
  // adjacent row tiles that share each weight tile through TMA multicast (-1/3 of the L2->SM operand bytes)
  P.pair = 0;
  {
    static int pair_on = -1;
    if (pair_on < 0) {
      const char* e = getenv("T2H_PAIR");
      pair_on = e ? atoi(e) : 0;
    }
    if (pair_on && !swap && rows_mode && !p->tile_rows && p->n_img == 1 && p->H == 1 && P.ksplit == 1 &&
        !p->a_bcast && !p->b_batched && !p->b_batched_h && P.tiles_w >= 2 && P.total_tiles >= 32) {
      P.pair = 1;
      P.total_tiles = ceil_div(P.tiles_w, 2) * P.n_tiles_n;
      P.total_work = P.total_tiles;
    }
  }
  if (p->gn_stats && !swap) {
    T2H_CHECK_ARG(P.epi_mode == EPI_TMA_F32, "tapgemm: gn_stats needs an aligned fp32 NHWC output");
    T2H_CHECK_ARG(p->gn_cpg >= 2 && (p->gn_cpg & (p->gn_cpg - 1)) == 0 && p->n_out % p->gn_cpg == 0,
                  "tapgemm: gn_cpg=%d must be a power of two >= 2 dividing n_out", p->gn_cpg);
    T2H_CHECK_ARG(BN % p->gn_cpg == 0 || p->n_out <= BN, "tapgemm: group straddles column tiles");
  }

  // ---- tensor maps
  CUtensorMap tmA, tmB, tmD, tmR;
  if (!p->a_f32) {
    uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->a_W, (uint64_t)p->a_H, (uint64_t)p->a_imgs};
    uint64_t str[4] = {1, (uint64_t)p->a_sw, (uint64_t)p->a_sh, (uint64_t)p->a_sn};
    uint32_t box[4] = {(uint32_t)kBK, (uint32_t)TW, (uint32_t)P.slab_rows, 1};
    if (p->a_mn) {  // (row, k, h, img): rows contiguous, a_sw = distance between consecutive k
      dims[0] = (uint64_t)p->a_W; dims[1] = (uint64_t)p->C;
      box[0] = 64; box[1] = (uint32_t)kBK; box[2] = 1;
    }
    int rc = make_tmap(&tmA, p->a, 2, 4, dims, str, box, "tapgemm A");
    if (rc) return rc;
  }
  {
    const int g2 = p->b_groups2 > 0 ? p->b_groups2 : 1;
    uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->n_out, (uint64_t)p->b_groups, (uint64_t)g2};
    uint64_t str[4] = {1, (uint64_t)p->b_sn, (uint64_t)p->b_sg,
                       (uint64_t)(g2 > 1 ? p->b_sg2 : p->b_sg)};
    uint32_t box[4] = {(uint32_t)kBK, (uint32_t)BN, 1, 1};
    if (p->b_mn) {  // (n, k, g, g2): output columns contiguous, b_sn = distance between consecutive k
      dims[0] = (uint64_t)p->n_out; dims[1] = (uint64_t)p->C;
      box[0] = 64; box[1] = (uint32_t)kBK;
    }
    int rc = make_tmap(&tmB, p->b, 2, 4, dims, str, box, "tapgemm B");
    if (rc) return rc;
  }
  if (p->a_f32) tmA = tmB;  // unused by the fused-producer kernel
  tmD = tmA;
  tmR = tmA;
  if (P.epi_mode != EPI_DIRECT) {
    // dims the output domain does not use get a harmless, 16-byte-aligned stride
    const uint64_t sw = (uint64_t)p->d_sw;
    const uint64_t sh = (p->H > 1) ? (uint64_t)p->d_sh : sw * (uint64_t)p->W;
    uint64_t sn = (p->n_img > 1) ? (uint64_t)p->d_sn : sh * (uint64_t)p->H;
    uint64_t imgs = (uint64_t)p->n_img;
    if (P.partials) {  // one slab per k-slice, addressed through the image dim
      sn = (uint64_t)p->d_slab;
      imgs = (uint64_t)P.ksplit;
    }
    if (P.epi_mode == EPI_TMA_PLANES && p->d_terms == 2) {
      if (p->n_img == 1) sn = (uint64_t)p->d_plane;
      P.d_term_imgs = (int)(p->d_plane / (long long)sn);
      imgs = (uint64_t)P.d_term_imgs + (uint64_t)p->n_img;
    }
    uint64_t dims[4] = {(uint64_t)p->n_out, (uint64_t)p->W, (uint64_t)p->H, imgs};
    uint64_t str[4] = {1, sw, sh, sn};
    uint32_t box[4] = {(uint32_t)(esz == 4 ? 32 : 64), (uint32_t)TW, (uint32_t)TH, 1};
    if (swap) box[2] = (uint32_t)(32 / TW);  // one warp's 32 pixels
    int rc = make_tmap(&tmD, p->d, esz, 4, dims, str, box, "tapgemm D");
    if (rc) return rc;
    if (p->residual) {
      uint64_t rdims[4] = {(uint64_t)p->n_out, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->n_img};
      rc = make_tmap(&tmR, p->residual, 4, 4, rdims, str, box, "tapgemm residual");
      if (rc) return rc;
    }
  }

  cudaStream_t s = as_stream(stream);
  if (swap && p->a_f32)
    return MBLK == 2 ? launch_swap<2, true>(tmA, tmB, tmD, tmR, P, s) : launch_swap<1, true>(tmA, tmB, tmD, tmR, P, s);
  if (swap) return MBLK == 2 ? launch_swap<2, false>(tmA, tmB, tmD, tmR, P, s) : launch_swap<1, false>(tmA, tmB, tmD, tmR, P, s);
  if (MBLK == 2) return launch<128, 2>(tmA, tmB, tmD, tmR, P, s);
  switch (BN) {
    case 16: return launch<16, 1>(tmA, tmB, tmD, tmR, P, s);
    case 32: return launch<32, 1>(tmA, tmB, tmD, tmR, P, s);
    case 64: return launch<64, 1>(tmA, tmB, tmD, tmR, P, s);
    case 128: return launch<128, 1>(tmA, tmB, tmD, tmR, P, s);
    default: return launch<256, 1>(tmA, tmB, tmD, tmR, P, s);
  }
}

// ---------------------------------------------------------------------------------------------------------
// t2h_conv_wgrad: dW[tap][co][ci] += alpha * sum_{n,h,w} dY[n,h,w,co] * X[n + ioff(tap), h + dy(tap), w + dx(tap), ci]
// on the same tcgen05 kernel: both operands are NHWC fp16 planes consumed MN-major (the contraction index -- the
// pixel -- is the outer dimension of both), the K loop walks 64-pixel patches (one TMA box {64 ch, PW, PH, 1}
// per operand and patch; the tap is a coordinate shift of the X box, zero-filled outside the image = the conv's
// padding), one output tile per (tap, 128 couts, BN cins), patches split over the SMs and TMA-reduce-added.
// ---------------------------------------------------------------------------------------------------------
extern "C" int t2h_conv_wgrad(const t2h_conv_wgrad_params* p, t2h_stream_t stream) {
  T2H_CHECK_ARG(p && p->dy && p->x && p->dw, "conv_wgrad: null operand");
  T2H_CHECK_ARG(p->n_img > 0 && p->H > 0 && p->W > 0 && p->cout > 0 && p->cin > 0, "conv_wgrad: empty problem");
  T2H_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= T2H_MAX_TAPS, "conv_wgrad: ntaps=%d", p->ntaps);
  T2H_CHECK_ARG(p->nterms == 1 || (p->nterms == 3 && p->dy_terms == 2 && p->x_terms == 2),
                "conv_wgrad: nterms=%d needs hi/lo planes on both operands", p->nterms);
  T2H_CHECK_ARG(p->cin % 4 == 0 && p->dw_ld % 4 == 0 && p->dw_tap_stride % 4 == 0 &&
                    reinterpret_cast<uintptr_t>(p->dw) % 16 == 0,
                "conv_wgrad: dW needs cin %% 4 == 0 and 16-byte aligned rows (cin=%d ld=%lld)", p->cin,
                (long long)p->dw_ld);
  TapGemmDev P;
  memset(&P, 0, sizeof(P));
  int PW = 16;
  while (PW > 1 && PW / 2 >= p->W) PW >>= 1;  // narrow images: taller patches
  const int PH = 64 / PW;
  P.wg = 1; P.accum = 1;
  P.wg_PW = PW; P.wg_PH = PH;
  P.wg_pw = ceil_div(p->W, PW);
  P.wg_ppi = P.wg_pw * ceil_div(p->H, PH);
  for (int i = 0; i < p->ntaps; ++i) {
    P.wg_dy[i] = p->tap_dy[i]; P.wg_dx[i] = p->tap_dx[i]; P.wg_ioff[i] = p->tap_img_off[i];
  }
  const long long chunks = (long long)p->n_img * P.wg_ppi;
  T2H_CHECK_ARG(chunks < (1LL << 24), "conv_wgrad: too many pixel patches");
  P.n_img = p->ntaps; P.H = 1; P.W = p->cout;
  P.n_out = p->cin; P.kchunks = (int)chunks; P.C = P.kchunks * kBK; P.nterms = p->nterms;
  P.a_term_imgs = p->dy_term_imgs; P.b_term_g = p->x_term_imgs;
  P.a_mn = 1; P.b_mn = 1;
  P.d = p->dw; P.d_mode = T2H_OUT_F32; P.alpha = p->alpha;
  P.bias_mode = T2H_BIAS_NONE; P.act = T2H_ACT_NONE;
  P.TW = 128; P.TH = 1;
  P.ngroups = 1; P.g_ntaps[0] = 1; P.slab_rows = 1;
  int BN = p->cin <= 64 ? 64 : (p->cin <= 128 ? 128 : 256);
  P.tiles_w = ceil_div(p->cout, 128); P.tiles_h = 1;
  P.n_tiles_n = ceil_div(p->cin, BN);
  P.wg_ntaps = p->ntaps;
  {
    static int no_pair = -1;
    if (no_pair < 0) {
      const char* e = getenv("T2H_WGRAD_NO_PAIR");
      no_pair = e ? atoi(e) : 0;
    }
    // 65..128 input channels: two taps share one 128 x 256 tile, so the MMAs issue with N = 256 (full rate)
    if (!no_pair && p->cin > 64 && p->cin <= 128 && p->ntaps >= 2) {
      P.wg_pair = 1;
      BN = 256;
      P.n_tiles_n = 1;
      P.n_img = ceil_div(p->ntaps, 2);
      P.n_out = 256;
    }
  }
  P.total_tiles = P.n_img * P.tiles_w * P.n_tiles_n;
  P.epi_mode = EPI_TMA_F32;
  int ks = p->k_split > 0 ? p->k_split : num_sms() / P.total_tiles;
  if (ks < 1) ks = 1;
  if (ks > P.kchunks) ks = P.kchunks;
  P.kper = ceil_div(P.kchunks, ks);
  P.ksplit = ceil_div(P.kchunks, P.kper);
  T2H_CHECK_ARG((long long)P.total_tiles * P.ksplit < (1LL << 31), "conv_wgrad: too many work items");
  P.total_work = P.total_tiles * P.ksplit;

  CUtensorMap tmA, tmB, tmD, tmR;
  {
    uint64_t dims[4] = {(uint64_t)p->cout, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->dy_imgs};
    uint64_t str[4] = {1, (uint64_t)p->dy_sw, (uint64_t)p->dy_sh, (uint64_t)p->dy_sn};
    uint32_t box[4] = {64, (uint32_t)PW, (uint32_t)PH, 1};
    int rc = make_tmap(&tmA, p->dy, 2, 4, dims, str, box, "conv_wgrad dY");
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p->cin, (uint64_t)p->x_W, (uint64_t)p->x_H, (uint64_t)p->x_imgs};
    uint64_t str[4] = {1, (uint64_t)p->x_sw, (uint64_t)p->x_sh, (uint64_t)p->x_sn};
    uint32_t box[4] = {64, (uint32_t)PW, (uint32_t)PH, 1};
    int rc = make_tmap(&tmB, p->x, 2, 4, dims, str, box, "conv_wgrad X");
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p->cin, (uint64_t)p->cout, 1, (uint64_t)p->ntaps};
    uint64_t str[4] = {1, (uint64_t)p->dw_ld, (uint64_t)p->dw_ld * (uint64_t)p->cout, (uint64_t)p->dw_tap_stride};
    uint32_t box[4] = {32, 128, 1, 1};
    int rc = make_tmap(&tmD, p->dw, 4, 4, dims, str, box, "conv_wgrad dW");
    if (rc) return rc;
  }
  tmR = tmA;
  cudaStream_t s = as_stream(stream);
  switch (BN) {
    case 64: return launch<64, 1>(tmA, tmB, tmD, tmR, P, s);
    case 128: return launch<128, 1>(tmA, tmB, tmD, tmR, P, s);
    default: return launch<256, 1>(tmA, tmB, tmD, tmR, P, s);
  }
}

#include "attn_fused.cuh"
