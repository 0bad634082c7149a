// Two-CTA (cta_group::2) variant of the per-tap implicit-GEMM convolution (conv_fwd.cu) for the
// layers the halo kernels do not take (Cout > 128 or W < 128: the deep levels).  The two CTAs of a
// cluster own two neighbouring 128-pixel tiles for the same N tile and share the weight tile: each
// streams its own activation tile and HALF of the weight tile per k-step (16 + <=16 KB instead of
// 16 + <=32 KB), one tcgen05.mma.cta_group::2 (M = 256) feeds both accumulators.  At batch 4 the
// deep layers are L2 -> shared-memory bound in the one-CTA kernel.  bf16, 3x3, 64-channel chunks.
// Barrier protocol as in conv_halo2.cu.
#include "launch.cuh"
#include <cstdlib>
#include "conv.h"
#include "conv_epilogue.cuh"
#include "ptx.cuh"
#include "cluster2.cuh"

namespace ub2 {

static constexpr int kF2MaxStages = 8;
static constexpr int kF2EpiWarps = 8;
static constexpr int kF2Threads = 64 + 32 * kF2EpiWarps;
static constexpr int kF2ATileBytes = 128 * 64 * 2;

struct Fwd2SmemHeader {
  uint64_t full[kF2MaxStages];
  uint64_t empty[kF2MaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// TAPS: 1 or 9 (unrolled in the producer).  ACC: see conv_fwd.cu.
template <int TAPS, bool ACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kF2Threads, 1)
conv_fwd2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const ConvFwdParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  const int stage_bytes = p.stage_bytes;
  const int stages = p.stages;
  Fwd2SmemHeader* hdr = reinterpret_cast<Fwd2SmemHeader*>(tiles + stages * stage_bytes);
  float* s_stats = reinterpret_cast<float*>(hdr + 1);  // [4 quarters][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int C0 = p.C0;
  const int Ctot = p.C0 + p.C1;
  const int kchunks = Ctot / 64;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int m_pairs = (tiles_m + 1) / 2;
  const int n_tiles = p.n_tiles;
  const int total_items = m_pairs * n_tiles;   // cluster work items: (pixel-tile pair, N tile)
  const int BN = p.BN;
  const int half = BN / 2;
  const int bn_cols = (BN + 31) & ~31;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.C1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&hdr->full[i], 1);    // leader: armed with both CTAs' bytes
      mbar_init(&hdr->empty[i], 1);   // multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->tmem_full[i], 1);
      mbar_init(&hdr->tmem_empty[i], 2 * kF2EpiWarps);   // leader: epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(&hdr->tmem_base, p.tmem_cols);
  if (warp >= 2 && p.stats != nullptr) {
    for (int i = threadIdx.x - 64; i < 4 * 2 * p.Cout; i += 32 * kF2EpiWarps) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();   // the prologue above touched no global memory; everything below may (launch.cuh)

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t tx_bytes = 128u * 64u * 2u + static_cast<uint32_t>(half) * 64u * 2u;
    const int tw = p.tiles_w, th = p.tiles_h, BW = p.BW, BH = p.BH, BI = p.BI;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < total_items; t += n_clusters) {
      const int nt = t % n_tiles;
      const int mt = 2 * (t / n_tiles) + static_cast<int>(rank);   // == tiles_m for a phantom tile: out of bounds
      const int w0 = (mt % tw) * BW;
      const int h0 = ((mt / tw) % th) * BH;
      const int i0 = (mt / (tw * th)) * BI;
      const int n0 = nt * BN + static_cast<int>(rank) * half;
#pragma unroll
      for (int tap = 0; tap < TAPS; ++tap) {
        const int dr = (TAPS == 9) ? tap / 3 - 1 : 0;
        const int ds = (TAPS == 9) ? tap % 3 - 1 : 0;
        for (int c = 0; c < Ctot; c += 64) {
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = tiles + stage * stage_bytes;
            const uint32_t bar = mapa_u32(smem_u32(&hdr->full[stage]), 0);
            if (leader) mbar_expect_tx(&hdr->full[stage], 2 * tx_bytes);
            if (c < C0)
              tma2_load_4d(sa, &tmA0, bar, c, w0 + ds, h0 + dr, i0);
            else
              tma2_load_4d(sa, &tmA1, bar, c - C0, w0 + ds, h0 + dr, i0);
            tma2_load_2d(sa + kF2ATileBytes, &tmB, bar, tap * Ctot + c, n0);
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------- MMA issuer (leader CTA only)
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem_u32(tiles) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = a_lo0 + (kF2ATileBytes >> 4);
      const uint32_t stage_inc = static_cast<uint32_t>(stage_bytes) >> 4;
      const int ksteps = TAPS * kchunks;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < total_items; t += n_clusters, ++it) {
        const int as = it & 1;
        mbar_wait(&hdr->tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * bn_cols;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&hdr->full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + stage * stage_inc;
            const uint32_t b_lo = b_lo0 + stage * stage_inc;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k);
              const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
              umma2_bf16(d_tmem, da, db, idesc, (ks | k) != 0);
            }
            umma2_commit(&hdr->empty[stage]);
            if (ks == ksteps - 1) umma2_commit(&hdr->tmem_full[as]);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs, own tile)
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int w_l = row % p.BW;
    const int h_l = (row / p.BW) % p.BH;
    const int i_l = row / (p.BW * p.BH);
    const int nchunks = bn_cols / 32;
    const bool want_stats = p.stats != nullptr;
    float* my_stats = s_stats + q * 2 * p.Cout;
    float acc_s[ACC ? 32 : 1], acc_q[ACC ? 32 : 1];
    if (ACC) {
#pragma unroll
      for (int i = 0; i < (ACC ? 32 : 1); ++i) acc_s[i] = acc_q[i] = 0.f;
    }
    int it = 0;
    for (int t = cluster_id; t < total_items; t += n_clusters, ++it) {
      const int nt = t % n_tiles;
      const int mt = 2 * (t / n_tiles) + static_cast<int>(rank);
      const int w = (mt % p.tiles_w) * p.BW + w_l;
      const int h = ((mt / p.tiles_w) % p.tiles_h) * p.BH + h_l;
      const int n = (mt / (p.tiles_w * p.tiles_h)) * p.BI + i_l;
      const int n0 = nt * BN;
      const bool valid = (w < p.W) && (h < p.H) && (n < p.N);
      const size_t pix = (static_cast<size_t>(valid ? n : 0) * p.H + h) * p.W + w;
      const int as = it & 1;
      mbar_wait(&hdr->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      for (int j = grp; j < nchunks; j += 2) {
        epi_chunk<ACC>(p, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * bn_cols + j * 32,
                       n0 + j * 32, n0 + BN, valid, pix, lane, want_stats, my_stats, acc_s, acc_q);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&hdr->tmem_empty[as]), 0));
    }
    if (want_stats)
      epi_finish<ACC, 32 * kF2EpiWarps>(p, s_stats, my_stats, lane, grp, nchunks, threadIdx.x - 64, acc_s, acc_q);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

static int pow2_floor2(int x) {
  int r = 1;
  while (r * 2 <= x) r *= 2;
  return r;
}

// returns 1 if the shape is not eligible (caller continues with the one-CTA kernel)
int conv_fwd2_launch(const ConvFwdArgs& a, cudaStream_t stream) {
  static const int enabled = [] { const char* e = getenv("UB2_FWD2"); return e ? atoi(e) : 1; }();
  const int Ctot = a.C0 + a.C1;
  // 1x1 convolutions (one k-step per tile) stay on the one-CTA kernel: measured slower in pairs
  if (!enabled || a.tf32 || a.taps != 9 || a.bn_override > 0 || a.grid_override > 0) return 1;
  if (Ctot % 64 != 0 || a.C0 % 64 != 0 || a.Cout % 32 != 0) return 1;
  if (a.ld_in0 % 8 != 0 || (a.C1 > 0 && a.ld_in1 % 8 != 0) || a.ld0 % 8 != 0) return UB2_ERR_ALIGN;
  if (a.out1 != nullptr && (a.split % 8 != 0 || a.ld1 % 8 != 0)) return UB2_ERR_ALIGN;

  ConvFwdParams p{};
  p.N = a.N; p.H = a.H; p.W = a.W; p.C0 = a.C0; p.C1 = a.C1; p.Cout = a.Cout; p.taps = a.taps; p.kc = 64;
  p.BW = pow2_floor2(a.W < 128 ? a.W : 128);
  p.BH = pow2_floor2(a.H < 128 / p.BW ? a.H : 128 / p.BW);
  p.BI = 128 / (p.BW * p.BH);
  p.tiles_w = (a.W + p.BW - 1) / p.BW;
  p.tiles_h = (a.H + p.BH - 1) / p.BH;
  p.tiles_n = (a.N + p.BI - 1) / p.BI;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  if (m_tiles < 2) return 1;
  const int split = (a.out1 != nullptr) ? a.split : (1 << 30);

  struct Mc4 { int v[4]; int& operator[](int i) { return v[i]; } };
  static PerDevice<Mc4> max_clusters_pd;
  Mc4& max_clusters = max_clusters_pd.ref();
  int BN = a.Cout <= 256 ? a.Cout : 256;
  // few pixel tiles: halve the N tile so that the cluster grid covers the chip
  const int sms = num_sms();
  while (BN > 64 && BN % 64 == 0 && a.Cout % (BN / 2) == 0 && (split == (1 << 30) || split % (BN / 2) == 0) &&
         ((m_tiles + 1) / 2) * ((a.Cout + BN - 1) / BN) * 2 <= sms / 2)
    BN /= 2;
  if (a.Cout % BN != 0 || BN % 32 != 0) return 1;
  if (split != (1 << 30) && split % BN != 0) return 1;   // an N tile must not straddle the two outputs
  p.BN = BN;
  p.n_tiles = a.Cout / BN;
  const int bn_cols = (BN + 31) & ~31;
  int tmem_cols = 32;
  while (tmem_cols < 2 * bn_cols) tmem_cols *= 2;
  p.tmem_cols = tmem_cols;
  const int b_bytes = (((BN / 2) * 128) + 1023) & ~1023;   // this CTA's half of the weight tile
  p.stage_bytes = kF2ATileBytes + b_bytes;
  const int stats_bytes = a.stats ? 4 * 2 * a.Cout * 4 : 0;
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(Fwd2SmemHeader)) - stats_bytes;
  int stages = budget / p.stage_bytes;
  if (stages > kF2MaxStages) stages = kF2MaxStages;
  if (stages < 2) return 1;
  p.stages = stages;
  p.out0 = reinterpret_cast<__nv_bfloat16*>(a.out0); p.ld0 = a.ld0;
  p.out1 = reinterpret_cast<__nv_bfloat16*>(a.out1); p.ld1 = a.ld1; p.split = split;
  p.accumulate = a.accumulate;
  p.scale = a.scale; p.shift = a.shift; p.relu = a.relu;
  p.stats = a.stats;
  {
    static const int wide_env = [] { const char* e = getenv("UB2_WIDE_STORE"); return e ? atoi(e) : 1; }();
    p.wide_store = wide_env && conv_wide_store_ok(a.out0, a.ld0, a.out1, a.ld1, a.split, a.Cout);
  }

  const bool acc = a.stats != nullptr && bn_cols <= 64 && p.n_tiles == 1;
  const int variant = (a.taps == 9 ? 0 : 2) + (acc ? 1 : 0);
  if (max_clusters[variant] == 0) {
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (sms / 2));
    cfg.blockDim = dim3(kF2Threads);
    cfg.dynamicSmemBytes = 227 * 1024 - 2048;
    int n = 0;
    cudaError_t e;
    switch (variant) {
      case 0: e = cudaFuncSetAttribute(conv_fwd2_kernel<9, false>, attr, 227 * 1024);
              if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, conv_fwd2_kernel<9, false>, &cfg); break;
      case 1: e = cudaFuncSetAttribute(conv_fwd2_kernel<9, true>, attr, 227 * 1024);
              if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, conv_fwd2_kernel<9, true>, &cfg); break;
      case 2: e = cudaFuncSetAttribute(conv_fwd2_kernel<1, false>, attr, 227 * 1024);
              if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, conv_fwd2_kernel<1, false>, &cfg); break;
      default: e = cudaFuncSetAttribute(conv_fwd2_kernel<1, true>, attr, 227 * 1024);
               if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, conv_fwd2_kernel<1, true>, &cfg); break;
    }
    if (e != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      return 1;
    }
    max_clusters[variant] = n;
  }

  CUtensorMap tmA0, tmA1, tmB;
  const uint32_t boxA[4] = {64u, static_cast<uint32_t>(p.BW), static_cast<uint32_t>(p.BH), static_cast<uint32_t>(p.BI)};
  int rc = make_tmap_nhwc(&tmA0, a.in0, a.N, a.H, a.W, a.C0, a.ld_in0, boxA, 128);
  if (rc) return rc;
  if (a.C1 > 0) {
    rc = make_tmap_nhwc(&tmA1, a.in1, a.N, a.H, a.W, a.C1, a.ld_in1, boxA, 128);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  rc = make_tmap_2d(&tmB, a.wgt, static_cast<uint64_t>(a.taps) * Ctot, a.Cout, static_cast<uint64_t>(a.taps) * Ctot, 64,
                    BN / 2, 128);
  if (rc) return rc;

  const int total_items = ((m_tiles + 1) / 2) * p.n_tiles;
  const int resident = cap_clusters(max_clusters[variant]);
  const int clusters = resident < total_items ? resident : total_items;
  const int grid = 2 * clusters;
  if (a.stats && grid > a.stats_rows) return UB2_ERR_WORKSPACE;
  const size_t smem = 1024 + static_cast<size_t>(stages) * p.stage_bytes + sizeof(Fwd2SmemHeader) + stats_bytes;
  switch (variant) {
    case 0: note_variant(2); launch(conv_fwd2_kernel<9, false>, grid, kF2Threads, smem, stream, tmA0, tmA1, tmB, p); break;
    case 1: note_variant(2); launch(conv_fwd2_kernel<9, true>, grid, kF2Threads, smem, stream, tmA0, tmA1, tmB, p); break;
    case 2: note_variant(2); launch(conv_fwd2_kernel<1, false>, grid, kF2Threads, smem, stream, tmA0, tmA1, tmB, p); break;
    default: note_variant(2); launch(conv_fwd2_kernel<1, true>, grid, kF2Threads, smem, stream, tmA0, tmA1, tmB, p); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (a.grid_used) *a.grid_used = grid;
  return 0;
}

}  // namespace ub2
