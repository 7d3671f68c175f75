// Transposed ("swapped-operand") variant of the tap-GEMM for spatial convolutions whose output
// channel count is a multiple of 128.
//
// tcgen05.mma (kind::f16, M=128) occupies the tensor pipe ~120 cycles for any N <= 256
// (tools/umma_bench.cu, profiles/r01_umma_issue_rate.txt): only N=256 instructions reach the
// nominal 4096 MAC/clk/SM.  A 128-channel conv tiled as [pixels x couts] can only issue N=128.
// Here the roles are swapped: the weight tile is the M=128 operand and the activation slab view
// (256 consecutive slab rows = 256 output pixels) is the N=256 operand, so D^T[cout, pixel]
// accumulates at full rate for every Cout that is a multiple of 128.
//
// Epilogue: TMEM lane = output channel, TMEM column = pixel.  Each of the 4 epilogue warps owns 32
// channels = one 128-byte row segment of the NHWC output, works independently (no block barriers):
// tcgen05.ld 32 pixels -> +bias (per lane) -> +residual (its own TMA-loaded 4 KB tile) -> GroupNorm
// partial sums (two registers per thread) -> conflict-free transposed st.shared into a swizzled
// [32 pixels][32 channels] tile -> its own TMA store.
#pragma once

namespace t2h {

constexpr int kWarpTile = 32 * 128;  // 32 pixels x 32 fp32 channels

constexpr int kSwapThreads = 384;  // 4 control warps + 8 epilogue warps (two per TMEM lane quadrant)

// FUSE = true: the activation operand is NOT read as fp16 planes by TMA.  The kernel takes the fp32 NHWC tensor the
// previous conv wrote plus its GroupNorm statistics, and four producer warps (8-11, taken from the epilogue's eight)
// build the swizzled hi / lo slabs the MMAs read themselves: 128-bit loads -> (x - mean) * rstd * gamma + beta ->
// swish -> fp16 split -> st.shared in the 128-byte-swizzle image a TMA box would have produced.  This is
// Normalize() + nonlinearity() (vqgan_arch.py:510-517) folded into the consuming conv: the gn_apply pass and its
// 8 bytes per element of HBM traffic disappear.  Same arithmetic, same order as gn_apply_kernel -> identical slabs.
template <int MBLK, bool FUSE>
__global__ void __launch_bounds__(kSwapThreads, 1)
tapgemm_swap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
                    const __grid_constant__ TapGemmDev P) {
  using C = Cfg<128, MBLK>;
  constexpr int NPIX = MBLK * 128;  // pixels per tile = UMMA N
  constexpr int EW = FUSE ? 4 : 8;   // epilogue warps
  pdl_launch_dependents();  // the next kernel may start its prologue once every CTA of this one is running
  const int NA = P.a_slots, NB = P.b_slots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + NA * C::kASlot;
  uint8_t* out_buf = b_ring + NB * C::kBSlot;     // 8 warps x 4 KB
  uint8_t* res_buf = out_buf + 2 * kEpiBufBytes;  // 8 warps x 4 KB

  __shared__ __align__(8) uint64_t a_full[kMaxSlots];
  __shared__ __align__(8) uint64_t a_empty[kMaxSlots];
  __shared__ __align__(8) uint64_t b_full[kMaxSlots];
  __shared__ __align__(8) uint64_t b_empty[kMaxSlots];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t res_bar[8];
  __shared__ uint32_t tmem_base_s;
  __shared__ float gn_tab[FUSE ? 512 : 1];  // FUSE: per-channel scale | shift of the current image (C <= 256)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if (!FUSE) tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (P.epi_mode != EPI_DIRECT) tma_prefetch_desc(&tmD);
    if (P.residual) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EW);
    }
    for (int s = 0; s < 8; ++s) mbar_init(&res_bar[s], 1);
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
  // everything above touched only shared / tensor memory and kernel parameters; from here on the previous
  // kernel's results are read (and buffers it may still be reading are overwritten)
  pdl_wait();
  // tile schedule: round-robin over the CTAs, or -- when the epilogue accumulates per-image sums (nb_sums) -- one
  // contiguous range per CTA, so that a CTA stays inside one image and flushes its sums once or twice per launch
  // instead of once per tile (round-robin put ~1000 double atomics per launch on every (image, channel) address)
  const bool contig = P.nb_sums != nullptr;
  const int tile_first = contig ? (int)((long long)blockIdx.x * P.total_tiles / gridDim.x) : (int)blockIdx.x;
  const int tile_end = contig ? (int)((long long)(blockIdx.x + 1) * P.total_tiles / gridDim.x) : P.total_tiles;
  const int tile_step = contig ? 1 : (int)gridDim.x;

  const int a_planes = (P.nterms == 3) ? 2 : 1;
  const int slab_bytes = P.slab_rows * P.TW * 128;

  if (FUSE && warp >= 8) {
    // ---------------------------------------------- activation slab producers: GroupNorm + swish + fp16 split
    const int pt = threadIdx.x - 256;  // 0..127
    float* s_scale = gn_tab;
    float* s_shift = gn_tab + 256;
    const int tw_shift = 31 - __clz(P.TW);
    const int R = P.slab_rows * P.TW;  // pixels of one slab
    const int units = R * 8;           // 8 channels (one 16-byte smem chunk) each
    const int cpg = P.C / P.ag_groups;
    const double cnt = (double)P.ag_hw * cpg;
    int cur_img = -1;
    int sa = 0, pa = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_step) {
      const TileCoord t = decode_tile(P, tile, MBLK, 128);
      if (t.img != cur_img) {
        named_bar_sync(3, 128);  // nobody still reads the previous image's table
        for (int c = pt; c < P.C; c += 128) {
          const int g = c / cpg;
          const double su = P.ag_stats[((long long)t.img * P.ag_groups + g) * 2 + 0];
          const double sq = P.ag_stats[((long long)t.img * P.ag_groups + g) * 2 + 1];
          const double mean = su / cnt;
          double var = sq / cnt - mean * mean;
          if (var < 0) var = 0;
          const float rstd = (float)(1.0 / sqrt(var + (double)P.ag_eps));
          const float ga = P.ag_gamma[c] * rstd;
          s_scale[c] = ga;
          s_shift[c] = P.ag_beta[c] - (float)mean * ga;
        }
        named_bar_sync(3, 128);
        cur_img = t.img;
      }
      const float* ximg = P.ax + (long long)t.img * P.ax_sn;
      for (int g = 0; g < P.ngroups; ++g) {
        const int h_base = t.h0 + P.g_dy0[g], w_base = t.w0 + P.g_dx[g];
        for (int ch = 0; ch < P.kchunks; ++ch) {
          const int s_hi = sa;
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (++sa == NA) { sa = 0; pa ^= 1; }
          int s_lo = -1;
          if (a_planes == 2) {
            s_lo = sa;
            mbar_wait(&a_empty[sa], pa ^ 1);
            if (++sa == NA) { sa = 0; pa ^= 1; }
          }
          uint8_t* hi_base = a_ring + s_hi * C::kASlot;
          uint8_t* lo_base = a_ring + (s_lo >= 0 ? s_lo : s_hi) * C::kASlot;
          const int c0 = ch * kBK;
#pragma unroll 4
          for (int u = pt; u < units; u += 128) {
            const int r = u >> 3, j = u & 7;
            const int h = h_base + (r >> tw_shift), w = w_base + (r & (P.TW - 1));
            const bool ok = (h >= 0) && (h < P.ag_H) && (w >= 0) && (w < P.ag_W);
            const float* src = ximg + (long long)h * P.ax_sh + (long long)w * P.ax_sw + c0 + j * 8;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (ok) {
              v0 = __ldg(reinterpret_cast<const float4*>(src));
              v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            }
            const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            __align__(16) __half hh[8];
            __align__(16) __half ll[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float y = 0.f;
              if (ok) {   // conv zero padding applies to the NORMALISED activation: outside pixels stay 0
                y = f[e] * s_scale[c0 + j * 8 + e] + s_shift[c0 + j * 8 + e];
                if (P.ag_swish) y = y / (1.0f + __expf(-y));
              }
              split_f16(y, hh[e], ll[e]);
            }
            *reinterpret_cast<uint4*>(hi_base + swz(r, j)) = *reinterpret_cast<const uint4*>(hh);
            if (s_lo >= 0) *reinterpret_cast<uint4*>(lo_base + swz(r, j)) = *reinterpret_cast<const uint4*>(ll);
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core's async proxy
          named_bar_sync(3, 128);
          if (pt == 0) {
            mbar_arrive(&a_full[s_hi]);
            if (s_lo >= 0) mbar_arrive(&a_full[s_lo]);
          }
        }
      }
    }
  } else if (!FUSE && warp == 0) {
    // ---------------------------------------------- activation slab producer
    if (lane == 0 && !(P.debug & 8)) {
      int sa = 0, pa = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        const TileCoord t = decode_tile(P, tile, MBLK, 128);
        for (int g = 0; g < P.ngroups; ++g)
          for (int ch = 0; ch < P.kchunks; ++ch)
            for (int pl = 0; pl < a_planes; ++pl) {
              mbar_wait(&a_empty[sa], pa ^ 1);
              if (P.debug & 4) {
                mbar_arrive(&a_full[sa]);
              } else {
                mbar_expect_tx(&a_full[sa], slab_bytes);
                tma_load_4d(&tmA, &a_full[sa], a_ring + sa * C::kASlot, ch * kBK, t.w0 + P.g_dx[g],
                            t.h0 + P.g_dy0[g], t.img + P.g_ioff[g] + pl * P.a_term_imgs);
              }
              if (++sa == NA) {
                sa = 0;
                pa ^= 1;
              }
            }
      }
    }
  } else if (warp == 3) {
    // ---------------------------------------------- weight tile producer (128 couts x 64 k)
    if (lane == 0 && !(P.debug & 8)) {
      int sb = 0, pb = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        const TileCoord t = decode_tile(P, tile, MBLK, 128);
        for (int g = 0; g < P.ngroups; ++g)
          for (int ch = 0; ch < P.kchunks; ++ch)
            for (int tp = 0; tp < P.g_ntaps[g]; ++tp)
              for (int pl = a_planes - 1; pl >= 0; --pl) {  // lo first, then hi
                mbar_wait(&b_empty[sb], pb ^ 1);
                if (P.debug & 4) {
                  mbar_arrive(&b_full[sb]);
                } else {
                  mbar_expect_tx(&b_full[sb], C::kBSlot);
                  tma_load_4d(&tmB, &b_full[sb], b_ring + sb * C::kBSlot, ch * kBK, t.n0,
                              P.g_btap[g][tp] + pl * P.b_term_g, 0);
                }
                if (++sb == NB) {
                  sb = 0;
                  pb ^= 1;
                }
              }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------- MMA issuer: D^T[cout, pixel] += W * slab^T
    // One thread; everything it needs per MMA is two 32-bit adds (descriptor low words), so the
    // issue stream stays far below the 128 cycles an M128 x N256 x K16 MMA occupies the pipe.
    {  // the whole warp runs the loop; one elected lane issues the tcgen05 instructions
      constexpr uint32_t IDESC = umma_idesc_f16(128, NPIX);
      const int last_steps = (P.C - (P.kchunks - 1) * kBK + 15) / 16;
      const uint32_t row16 = (uint32_t)(P.TW * 128) >> 4;  // one slab image row, in 16-byte units
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_ring));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(b_ring));
      constexpr uint32_t A16 = C::kASlot >> 4, B16 = C::kBSlot >> 4;
      const bool dbg_nobar = (P.debug & 8) != 0, dbg_nomma = (P.debug & 2) != 0;
      const int ngroups = P.ngroups, kchunks = P.kchunks;
      int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, ap = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        mbar_wait(&tempty_bar[as], ap ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * C::kAccCols;
        uint32_t acc = 0;
        for (int g = 0; g < ngroups; ++g) {
          const int nt = P.g_ntaps[g];
          const uint32_t dy0 = P.g_dyrel[g][0] * row16, dy1 = P.g_dyrel[g][1] * row16,
                         dy2 = P.g_dyrel[g][2] * row16;
          for (int ch = 0; ch < kchunks; ++ch) {
            const int ksteps = (ch == kchunks - 1) ? last_steps : 4;
            const int sa_hi = sa, pa_hi = pa;
            if (++sa == NA) { sa = 0; pa ^= 1; }
            const int sa_lo = sa, pa_lo = pa;
            if (a_planes == 2) {
              if (++sa == NA) { sa = 0; pa ^= 1; }
            }
            if (!dbg_nobar) mbar_wait(&a_full[sa_hi], pa_hi);
            tc_fence_after();
            const uint32_t xhi = a_lo0 + sa_hi * A16;
            const uint32_t xlo = a_lo0 + sa_lo * A16;
            for (int tp = 0; tp < nt; ++tp) {
              const uint32_t dy = tp == 0 ? dy0 : (tp == 1 ? dy1 : dy2);
              if (a_planes == 1) {
                if (!dbg_nobar) mbar_wait(&b_full[sb], pb);
                tc_fence_after();
                if (!dbg_nomma) { if (elect_one()) umma_ksteps(d_base, b_lo0 + sb * B16, xhi + dy, IDESC, ksteps, acc); acc = 1; __syncwarp(); }
                if (!dbg_nobar) { if (elect_one()) umma_commit(&b_empty[sb]); __syncwarp(); }
                if (++sb == NB) { sb = 0; pb ^= 1; }
              } else {
                if (!dbg_nobar) mbar_wait(&b_full[sb], pb);  // w_lo
                tc_fence_after();
                if (!dbg_nomma) { if (elect_one()) umma_ksteps(d_base, b_lo0 + sb * B16, xhi + dy, IDESC, ksteps, acc); acc = 1; __syncwarp(); }  // x_hi*w_lo
                if (!dbg_nobar) { if (elect_one()) umma_commit(&b_empty[sb]); __syncwarp(); }
                if (++sb == NB) { sb = 0; pb ^= 1; }
                if (!dbg_nobar) mbar_wait(&b_full[sb], pb);  // w_hi
                tc_fence_after();
                const uint32_t whi = b_lo0 + sb * B16;
                if (!dbg_nomma) { if (elect_one()) umma_ksteps(d_base, whi, xhi + dy, IDESC, ksteps, acc); acc = 1; __syncwarp(); }  // x_hi*w_hi
                if (tp == nt - 1 && !dbg_nobar) { if (elect_one()) umma_commit(&a_empty[sa_hi]); __syncwarp(); }
                if (tp == 0 && !dbg_nobar) {
                  mbar_wait(&a_full[sa_lo], pa_lo);
                  tc_fence_after();
                }
                if (!dbg_nomma) { if (elect_one()) umma_ksteps(d_base, whi, xlo + dy, IDESC, ksteps, acc); acc = 1; __syncwarp(); }  // x_lo*w_hi
                if (!dbg_nobar) { if (elect_one()) umma_commit(&b_empty[sb]); __syncwarp(); }
                if (++sb == NB) { sb = 0; pb ^= 1; }
              }
            }
            if (!dbg_nobar) { if (elect_one()) umma_commit(a_planes == 1 ? &a_empty[sa_hi] : &a_empty[sa_lo]); __syncwarp(); }
          }
        }
        { if (elect_one()) umma_commit(&tfull_bar[as]); __syncwarp(); }
        if (P.debug & 128) mbar_wait(&tfull_bar[as], ap);  // experiment: serialise tiles (wait for this tile's MMAs)
        if (++as == 2) {
          as = 0;
          ap ^= 1;
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + EW) {
    // ---------------------------------------------- per-warp transposed epilogue
    // Two warps per TMEM lane quadrant take alternate 32-pixel chunks, so one warp's TMEM/TMA
    // latencies hide behind the other's arithmetic.  Each warp owns one 4 KB output tile and one
    // 4 KB residual tile in shared memory.
    const int q = warp & 3;
    const int e = warp - 4;          // 0..7
    const int half = e >> 2;         // which of the quadrant's two warps
    uint8_t* my_out = out_buf + e * kWarpTile;
    uint8_t* my_res = res_buf + e * kWarpTile;
    const bool has_res = P.residual != nullptr;
    const int tw_shift = 31 - __clz(P.TW);      // TW is a power of two <= 32
    const int rows_per_chunk = 32 >> tw_shift;  // image rows covered by 32 pixels
    constexpr int NCH = NPIX / 32;
    int as = 0, ap = 0;
    uint32_t res_par = 0;
    const int cpg = P.gn_cpg;
    const int red = cpg < 32 ? cpg : 32;  // lanes sharing a GroupNorm group inside this warp
    const bool nb = P.nb_sums != nullptr;  // norm-backward sums: the "residual" tile is x and is not added
    // (plain locals and an explicit flush at both sites: a by-reference lambda put the running sums in local memory)
    float nb_mean = 0.f, nb_rstd = 0.f, nb_ga = 0.f, nb_be = 0.f, nb_s1 = 0.f, nb_s2 = 0.f;
    int nb_img = -1, nb_c0 = -1;

    for (int tile = tile_first; tile < tile_end; tile += tile_step) {
      const TileCoord t = decode_tile(P, tile, MBLK, 128);
      const int c0 = t.n0 + q * 32;  // this warp's first output channel
      const float bias_c =
          (P.bias_mode == T2H_BIAS_COL && c0 + lane < P.n_out) ? __ldg(P.bias + t.img * P.bias_sn + c0 + lane) : 0.f;
      auto issue_res = [&](int k) {
        mbar_expect_tx(&res_bar[e], kWarpTile);
        tma_load_4d(&tmR, &res_bar[e], my_res, c0, t.w0, t.h0 + k * rows_per_chunk, t.img);
      };
      if (has_res && lane == 0) issue_res(half);
      constexpr int KSTEP = EW / 4;  // chunks are dealt round-robin to the quadrant's warps
      if (nb && lane == 0 && tile + tile_step < tile_end) {
        // the NEXT tile's x boxes go to L2 now, while its MMAs run: the epilogue's own loads then cost an L2 hit
        // instead of a DRAM round trip under load, four times per tile on the critical path
        const TileCoord tn = decode_tile(P, tile + tile_step, MBLK, 128);
        for (int k = half; k < NPIX / 32; k += KSTEP)
          tma_prefetch_4d(&tmR, tn.n0 + q * 32, tn.w0, tn.h0 + k * rows_per_chunk, tn.img);
      }
      // interior tiles need no per-pixel validity test for the GroupNorm sums
      const bool interior = (t.h0 + MBLK * P.TH <= P.H) && (t.w0 + P.TW <= P.W);
      float gs = 0.f, gss = 0.f;
      // norm-backward: this lane's channel constants of image t.img (mean, rstd from the forward statistics); the
      // sums run on across consecutive tiles of the same (image, channel block) and are flushed when that changes
      if (nb && (t.img != nb_img || c0 != nb_c0)) {
        if (nb_img >= 0 && nb_c0 + lane < P.n_out) {
          double* dst = P.nb_sums + ((long long)nb_img * P.n_out + nb_c0 + lane) * 2;
          atomicAdd(dst, (double)nb_s1);
          atomicAdd(dst + 1, (double)nb_s2);
        }
        nb_s1 = 0.f;
        nb_s2 = 0.f;
        nb_img = t.img;
        nb_c0 = c0;
        const int c = min(c0 + lane, P.n_out - 1);
        const int ncpg = P.n_out / P.nb_groups;
        const double cnt = (double)P.H * P.W * ncpg;
        const double* st = P.nb_stats + ((long long)t.img * P.nb_groups + c / ncpg) * 2;
        const double m = st[0] / cnt;
        double var = st[1] / cnt - m * m;
        if (var < 0) var = 0;
        nb_mean = (float)m;
        nb_rstd = (float)(1.0 / sqrt(var + (double)P.nb_eps));
        nb_ga = __ldg(P.nb_gamma + c);
        nb_be = __ldg(P.nb_beta + c);
      }
      mbar_wait(&tfull_bar[as], ap);
      tc_fence_after();
#pragma unroll 1
      for (int k = half; k < NCH; k += KSTEP) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * C::kAccCols + k * 32, r);
        tmem_ld_wait();
        if (k + KSTEP >= NCH) {
          // this warp's last TMEM read of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        if (P.debug & 1) {
          if (has_res) {
            mbar_wait(&res_bar[e], res_par);
            res_par ^= 1;
            __syncwarp();
            if (lane == 0 && k + KSTEP < NCH) issue_res(k + KSTEP);
          }
          continue;
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * P.alpha + bias_c;
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
          mbar_wait(&res_bar[e], res_par);
          res_par ^= 1;
          if (nb) {
            // v = dL/d act(norm(x)); accumulate pass 1 of the norm backward: sum du, sum du*xhat over the pixels.
            // Pixels outside the image are zeroed up front (their x tile is zero-filled), the activation is chosen
            // outside the element loop, two partial sums break the dependency chains.
            if (!interior) {
              const int hrow0 = t.h0 + k * rows_per_chunk;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (!((hrow0 + (i >> tw_shift) < P.H) && (t.w0 + (i & (P.TW - 1)) < P.W))) v[i] = 0.f;
            }
            float s1b = 0.f, s2b = 0.f;
            if (P.nb_act == 1) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float x0 = *reinterpret_cast<const float*>(my_res + swz(i, lane >> 2) + ((lane & 3) << 2));
                const float x1 = *reinterpret_cast<const float*>(my_res + swz(i + 1, lane >> 2) + ((lane & 3) << 2));
                const float xh0 = (x0 - nb_mean) * nb_rstd, xh1 = (x1 - nb_mean) * nb_rstd;
                const float du0 = v[i] * act_grad_fast(fmaf(xh0, nb_ga, nb_be), 1);
                const float du1 = v[i + 1] * act_grad_fast(fmaf(xh1, nb_ga, nb_be), 1);
                nb_s1 += du0; s1b += du1;
                nb_s2 = fmaf(du0, xh0, nb_s2); s2b = fmaf(du1, xh1, s2b);
              }
            } else {
              const int a = P.nb_act;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float x0 = *reinterpret_cast<const float*>(my_res + swz(i, lane >> 2) + ((lane & 3) << 2));
                const float x1 = *reinterpret_cast<const float*>(my_res + swz(i + 1, lane >> 2) + ((lane & 3) << 2));
                const float xh0 = (x0 - nb_mean) * nb_rstd, xh1 = (x1 - nb_mean) * nb_rstd;
                const float du0 = v[i] * ((a == 2 && fmaf(xh0, nb_ga, nb_be) <= 0.f) ? 0.2f : 1.0f);
                const float du1 = v[i + 1] * ((a == 2 && fmaf(xh1, nb_ga, nb_be) <= 0.f) ? 0.2f : 1.0f);
                nb_s1 += du0; s1b += du1;
                nb_s2 = fmaf(du0, xh0, nb_s2); s2b = fmaf(du1, xh1, s2b);
              }
            }
            nb_s1 += s1b;
            nb_s2 += s2b;
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)  // row = pixel i, word = this lane's channel: conflict-free
              v[i] += *reinterpret_cast<const float*>(my_res + swz(i, lane >> 2) + ((lane & 3) << 2));
          }
          __syncwarp();
          if (lane == 0 && k + KSTEP < NCH) issue_res(k + KSTEP);  // overlaps the rest of this chunk
        }
        if (P.gn_stats) {
          if (interior) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              gs += v[i];
              gss = fmaf(v[i], v[i], gss);
            }
          } else {
            const int hrow0 = t.h0 + k * rows_per_chunk;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const bool ok = (hrow0 + (i >> tw_shift) < P.H) && (t.w0 + (i & (P.TW - 1)) < P.W);
              const float x = ok ? v[i] : 0.f;
              gs += x;
              gss = fmaf(x, x, gss);
            }
          }
        }
        if (P.epi_mode == EPI_DIRECT) {
          // strided destination (NCHW conv_out, Cout < 128): the few valid channels store their pixels
          const int c = c0 + lane;
          if (c < P.n_out) {
            float* dst = reinterpret_cast<float*>(P.d) + (long long)t.img * P.d_sn + (long long)c * P.d_sc;
            const int hrow0 = t.h0 + k * rows_per_chunk;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int h = hrow0 + (i >> tw_shift), w = t.w0 + (i & (P.TW - 1));
              if (h < P.H && w < P.W) dst[(long long)h * P.d_sh + (long long)w * P.d_sw] = v[i];
            }
          }
          continue;
        }
        if (lane == 0) tma_store_wait_read<0>();  // my_out's previous store has drained
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          *reinterpret_cast<float*>(my_out + swz(i, lane >> 2) + ((lane & 3) << 2)) = v[i];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmD, my_out, c0, t.w0, t.h0 + k * rows_per_chunk, t.img);
          tma_store_commit();
        }
      }
      if (P.gn_stats) {
        for (int off = 1; off < red; off <<= 1) {
          gs += __shfl_xor_sync(0xffffffffu, gs, off);
          gss += __shfl_xor_sync(0xffffffffu, gss, off);
        }
        if ((lane & (red - 1)) == 0) {
          const int g = (c0 + lane) / cpg;
          double* dst = P.gn_stats + ((long long)t.img * P.gn_groups + g) * 2;
          atomicAdd(dst, (double)gs);
          atomicAdd(dst + 1, (double)gss);
        }
      }
      if (++as == 2) {
        as = 0;
        ap ^= 1;
      }
    }
    if (nb && nb_img >= 0 && nb_c0 + lane < P.n_out) {
      double* dst = P.nb_sums + ((long long)nb_img * P.n_out + nb_c0 + lane) * 2;
      atomicAdd(dst, (double)nb_s1);
      atomicAdd(dst + 1, (double)nb_s2);
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace t2h
