// Two-CTA (cta_group::2) variant of the halo-resident weight-gradient kernel (conv_wgrad_halo.cu):
// the two CTAs of a cluster own two neighbouring 64-channel chunks of the same (tap group, pixel
// range) and share dy — each streams only HALF of dy's output channels, one tcgen05.mma.cta_group::2
// (M = 256) issued by the leader feeds both CTAs' accumulators.  Per SM the MMA then reads its A
// atoms and half of B from shared memory, and the dy traffic from L2 halves.  Barrier protocol as in
// conv_halo2.cu.  Needs an even number of channel chunks (C0 + C1 a multiple of 128).
//
// Original description (one-CTA kernel):
// Halo-resident weight-gradient kernel for the wide 3x3 layers (W a multiple of 128, Cout <= 128).
//
//   dW[tap*Ctot + c, co] = sum_pixels a[p + shift(tap), c] * dy[p, co]
//
// The per-tap wgrad kernel (conv_wgrad.cu) loads one tap-shifted activation box per 64 im2col rows
// and stage; at Cout <= 128 that is 3-5x more L2->smem traffic than the tensor pipe can hide.  Here
// a work item owns one 64-channel chunk (x one tap group x one pixel range): per block of R image
// rows x 128 columns ONE TMA box {64 ch, 130, R+2} brings the halo block, and all taps are shifted
// MN-major descriptors into it (the swizzle depends on absolute smem address bits only:
// tools/probe/umma_shift.cu).  One MMA covers two taps: M = 128 = two 64-channel atoms whose
// distance (the descriptor's leading byte offset) is the shift between the two taps.  dy streams
// through a small ring in 64-pixel pieces and is reused by every tap.  Accumulators (up to five
// 128 x N tiles) stay in TMEM for the whole pixel range; fp32 partial tiles per split are folded
// by wgrad_reduce_kernel in a fixed order.
#include "launch.cuh"
#include <cstdlib>
#include "conv.h"
#include "ptx.cuh"
#include "cluster2.cuh"

namespace ub2 {

static constexpr int kG2Threads = 192;
static constexpr int kG2MaxBStages = 12;
static constexpr int kG2RW = 130;

struct WgHalo2Params {
  int N, H, W, C0, C1, Cout, R;
  int bn_cols, kchunks, tapgroups, splits, per_split, blocks, segs_w, blocks_h;
  int a_bytes, b_stage_bytes, b_stages, tmem_cols;
  int half;   // dy channels streamed by each CTA (Cout / 2: 64 -> 128-byte rows, 32 -> 64-byte rows)
  // Single channel chunk (C0 + C1 == 64): the pair splits the TAPS instead.  Both CTAs issue the
  // same three tap pairs (virtual taps 0..5); the peer's halo block is loaded two image rows lower,
  // so its virtual taps 0,1,2 are the real taps 6,7,8 (its other three tiles are discarded):
  // three MMA pairs per k-step instead of five, and each CTA streams half of dy.
  int tapsplit;
  float* partial;
};

struct WgHalo2Header {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kG2MaxBStages], b_empty[kG2MaxBStages];
  uint64_t tmem_full, tmem_empty;
  uint32_t tmem_base, pad;
};

// offset (in 16-byte units) of tap t's window inside the halo block: ((1+dr)*130 + 1+ds) pixels
__device__ __forceinline__ int tap_off2(int t) { return ((t / 3) * kG2RW + (t % 3)) * 8; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
conv_wgrad_halo2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmDY, const WgHalo2Params p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* sA = sbase;
  uint8_t* sB = sbase + 2 * p.a_bytes;
  WgHalo2Header* hdr = reinterpret_cast<WgHalo2Header*>(sB + p.b_stages * p.b_stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.R;
  const int Ctot = p.C0 + p.C1;
  const int Mtot = 9 * Ctot;

  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.C1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->a_full[i], 1);    // leader: armed with both CTAs' bytes
      mbar_init(&hdr->a_empty[i], 1);   // multicast commit
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&hdr->b_full[i], 1);
      mbar_init(&hdr->b_empty[i], 1);
    }
    mbar_init(&hdr->tmem_full, 1);
    mbar_init(&hdr->tmem_empty, 2 * 4);  // leader: the epilogue warps of both CTAs
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(&hdr->tmem_base, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();   // the prologue above touched no global memory; everything below may (launch.cuh)

  // pair item -> (two neighbouring channel chunks, tap group, split); this CTA owns chunk 2k + rank
  const int pair_groups = p.tapsplit ? 1 : (p.kchunks / 2) * p.tapgroups;
  const int items = pair_groups * p.splits;   // (shadows the one-CTA item count above)
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int tap_shift = p.tapsplit ? 6 * static_cast<int>(rank) : 0;   // real tap = virtual tap + tap_shift
  const int row_shift = p.tapsplit ? 2 * static_cast<int>(rank) : 0;   // image rows the halo block starts lower
  auto decode_item = [&](int it, int& c, int& t_begin, int& t_end, int& sp) {
    sp = it / pair_groups;
    const int g = it % pair_groups;
    if (p.tapsplit) {
      c = 0;
      t_begin = 0;
      t_end = 6;
      return;
    }
    c = (2 * (g % (p.kchunks / 2)) + static_cast<int>(rank)) * 64;
    const int tg = g / (p.kchunks / 2);
    if (p.tapgroups == 1) { t_begin = 0; t_end = 9; }
    else if (tg == 0) { t_begin = 0; t_end = 6; }
    else { t_begin = 6; t_end = 9; }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const uint32_t a_tx = 64u * kG2RW * (R + 2) * 2u;
    const uint32_t b_tx = static_cast<uint32_t>(p.half) * 64u * 2u;   // 64 pixels x half the channels
    int abuf = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int it = cluster_id; it < items; it += n_clusters) {
      int c, t_begin, t_end, sp;
      decode_item(it, c, t_begin, t_end, sp);
      const int blk_begin = sp * p.per_split;
      const int blk_end = min(blk_begin + p.per_split, p.blocks);
      int seg = blk_begin % p.segs_w;
      int hb = (blk_begin / p.segs_w) % p.blocks_h;
      int n = blk_begin / (p.segs_w * p.blocks_h);
      for (int blk = blk_begin; blk < blk_end; ++blk) {
        const int w0 = seg * 128, h0 = hb * R;
        mbar_wait(&hdr->a_empty[abuf], aphase ^ 1);
        if (elect_one()) {
          const uint32_t bar = mapa_u32(smem_u32(&hdr->a_full[abuf]), 0);
          if (leader) mbar_expect_tx(&hdr->a_full[abuf], 2 * a_tx);
          if (c < p.C0)
            tma2_load_4d(sA + abuf * p.a_bytes, &tmA0, bar, c, w0 - 1, h0 - 1 + row_shift, n);
          else
            tma2_load_4d(sA + abuf * p.a_bytes, &tmA1, bar, c - p.C0, w0 - 1, h0 - 1 + row_shift, n);
        }
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        for (int kc = 0; kc < 2 * R; ++kc) {  // 64-pixel pieces: row kc/2, half kc%2
          mbar_wait(&hdr->b_empty[bs], bphase ^ 1);
          if (elect_one()) {
            const uint32_t bar = mapa_u32(smem_u32(&hdr->b_full[bs]), 0);
            if (leader) mbar_expect_tx(&hdr->b_full[bs], 2 * b_tx);
            tma2_load_4d(sB + bs * p.b_stage_bytes, &tmDY, bar, static_cast<int>(rank) * p.half,
                         w0 + (kc & 1) * 64, h0 + (kc >> 1), n);
          }
          if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        }
        if (++seg == p.segs_w) {
          seg = 0;
          if (++hb == p.blocks_h) { hb = 0; ++n; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(256, p.Cout, 1, 1);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      // B: this CTA's half of dy, MN-major rows of `half` channels (128-byte or 64-byte swizzle)
      const uint32_t b_pitch = static_cast<uint32_t>(p.half) * 2u;
      const uint32_t b_hi = ((8u * b_pitch) >> 4) | (1u << 14) | ((b_pitch == 128u ? 2u : 4u) << 29);
      const uint32_t b_kadv = (16u * b_pitch) >> 4;
      const uint32_t a_base = (smem_u32(sA) & 0x3FFFFu) >> 4;
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);
      const uint32_t a_inc = static_cast<uint32_t>(p.a_bytes) >> 4;
      const uint32_t b_inc = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
      int abuf = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      int n_it = 0;
      for (int it = cluster_id; it < items; it += n_clusters, ++n_it) {
        int c, t_begin, t_end, sp;
        decode_item(it, c, t_begin, t_end, sp);
        const int blk_begin = sp * p.per_split;
        const int blk_end = min(blk_begin + p.per_split, p.blocks);
        const int npairs = (t_end - t_begin + 1) / 2;
        mbar_wait(&hdr->tmem_empty, (n_it & 1) ^ 1);
        tc_fence_after();
        for (int blk = blk_begin; blk < blk_end; ++blk) {
          mbar_wait(&hdr->a_full[abuf], aphase);
          tc_fence_after();
          const uint32_t a_buf = a_base + abuf * a_inc;
          for (int kc = 0; kc < 2 * R; ++kc) {
            mbar_wait(&hdr->b_full[bs], bphase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b_lo = b_lo0 + bs * b_inc;
              // first pixel of this 64-pixel piece inside the block (16-byte units)
              const uint32_t kpix = static_cast<uint32_t>(((kc >> 1) * kG2RW + (kc & 1) * 64) * 8);
              for (int pr = 0; pr < npairs; ++pr) {
                const int t0 = t_begin + 2 * pr;
                const int off0 = tap_off2(t0);
                const int lbo = (t0 + 1 < t_end) ? tap_off2(t0 + 1) - off0 : 8;  // single tap: any valid atom
                const uint32_t a_lo = ((a_buf + kpix + off0) & 0x3FFFu) | (static_cast<uint32_t>(lbo) << 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k * 128);
                  const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k * b_kadv);
                  umma2_bf16(tmem_base + pr * p.bn_cols, da, db, idesc, (blk > blk_begin) || (kc > 0) || (k > 0));
                }
              }
              umma2_commit(&hdr->b_empty[bs]);
              if (kc == 2 * R - 1) {
                umma2_commit(&hdr->a_empty[abuf]);
                if (blk == blk_end - 1) umma2_commit(&hdr->tmem_full);
              }
            }
            __syncwarp();
            if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
          }
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;
    const int row = q * 32 + lane;  // TMEM lane: rows 0..63 = first tap of the pair, 64..127 = second
    int n_it = 0;
    for (int it = cluster_id; it < items; it += n_clusters, ++n_it) {
      int c, t_begin, t_end, sp;
      decode_item(it, c, t_begin, t_end, sp);
      const int npairs = (t_end - t_begin + 1) / 2;
      mbar_wait(&hdr->tmem_full, n_it & 1);
      tc_fence_after();
      for (int pr = 0; pr < npairs; ++pr) {
        const int vtap = t_begin + 2 * pr + (row >> 6);
        const int tap = vtap + tap_shift;
        const bool rvalid = vtap < t_end && tap < 9;
        const int m = tap * Ctot + c + (row & 63);
        float* dst = p.partial + (static_cast<size_t>(sp) * Mtot + m) * p.Cout;
        for (int j = 0; j < p.bn_cols / 32; ++j) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + pr * p.bn_cols + j * 32, raw);
          tmem_ld_wait();
          if (rvalid) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int col = j * 32 + g * 4;
              if (col < p.Cout) {
                float4 o;
                o.x = __uint_as_float(raw[g * 4 + 0]);
                o.y = __uint_as_float(raw[g * 4 + 1]);
                o.z = __uint_as_float(raw[g * 4 + 2]);
                o.w = __uint_as_float(raw[g * 4 + 3]);
                *reinterpret_cast<float4*>(dst + col) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&hdr->tmem_empty), 0));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be signalling this CTA's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

// returns 1 if the shape is not eligible (caller falls back to the per-tap kernel)
int conv_wgrad_halo2_launch(const ConvWgradArgs& a, cudaStream_t stream) {
  static const int enabled = [] { const char* e = getenv("UB2_WGRAD2"); return e ? atoi(e) : 1; }();
  const int Ctot = a.C0 + a.C1;
  // tap split (Ctot == 64) only pays at Cout = 128: both CTAs load the same halo block, and at
  // Cout = 64 that doubled activation traffic outweighs the saved MMAs (measured: 0.68 -> 0.75 ms)
  if (!enabled || a.taps != 9 || a.W % 128 != 0 || a.C0 % 64 != 0) return 1;
  if (Ctot % 128 != 0 && !(Ctot == 64 && a.Cout == 128)) return 1;
  if ((a.Cout != 64 && a.Cout != 128) || a.splits_override > 0) return 1;
  if (a.ld_in0 % 8 != 0 || (a.C1 > 0 && a.ld_in1 % 8 != 0) || a.ld_dy % 8 != 0) return UB2_ERR_ALIGN;

  WgHalo2Params p{};
  p.N = a.N; p.H = a.H; p.W = a.W; p.C0 = a.C0; p.C1 = a.C1; p.Cout = a.Cout;
  p.bn_cols = (a.Cout + 31) & ~31;
  p.kchunks = Ctot / 64;
  p.tapsplit = Ctot == 64;
  p.tapgroups = (p.tapsplit || 5 * p.bn_cols <= 512) ? 1 : 2;   // five / three 128-row tiles of bn_cols columns in TMEM
  const int tiles = (p.tapsplit || p.tapgroups == 2) ? 3 : 5;
  int tmem_cols = 32;
  while (tmem_cols < tiles * p.bn_cols) tmem_cols *= 2;
  p.tmem_cols = tmem_cols;
  p.half = a.Cout / 2;
  p.b_stage_bytes = p.half * 2 * 64;   // 64 pixels x half the channels: 8 KB or 4 KB
  const int budget = wgrad_smem_budget() - 1024 - static_cast<int>(sizeof(WgHalo2Header));
  int R = a.H < 2 ? a.H : 2;
  p.a_bytes = ((64 * kG2RW * (R + 2) * 2) + 1023) & ~1023;
  int b_stages = (budget - 2 * p.a_bytes) / p.b_stage_bytes;
  if (b_stages < 2) return 1;
  if (b_stages > kG2MaxBStages) b_stages = kG2MaxBStages;
  p.R = R;
  p.b_stages = b_stages;
  p.segs_w = a.W / 128;
  p.blocks_h = (a.H + R - 1) / R;
  p.blocks = a.N * p.blocks_h * p.segs_w;
  // resident CTA pairs (see conv_halo2.cu)
  static PerDevice<int> max_clusters_pd;
  int& max_clusters = max_clusters_pd.ref();
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    cfg.blockDim = dim3(kG2Threads);
    cfg.dynamicSmemBytes = 227 * 1024 - 2048;
    int n = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, conv_wgrad_halo2_kernel, &cfg);
    if (e != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      return 1;
    }
    max_clusters = n;
  }
  const int groups = p.tapsplit ? 1 : (p.kchunks / 2) * p.tapgroups;   // cluster work items per split
  const int resident = cap_clusters(max_clusters);
  int splits = resident / groups;
  if (splits < 1) splits = 1;
  if (splits > p.blocks) splits = p.blocks;
  if (splits > a.max_splits) splits = a.max_splits;
  p.per_split = (p.blocks + splits - 1) / splits;
  splits = (p.blocks + p.per_split - 1) / p.per_split;
  p.splits = splits;
  p.partial = a.partial;

  CUtensorMap tmA0, tmA1, tmDY;
  const uint32_t boxA[4] = {64u, static_cast<uint32_t>(kG2RW), static_cast<uint32_t>(R + 2), 1u};
  const uint32_t boxB[4] = {static_cast<uint32_t>(p.half), 64u, 1u, 1u};
  int rc = make_tmap_nhwc(&tmA0, a.in0, a.N, a.H, a.W, a.C0, a.ld_in0, boxA, 128);
  if (rc) return rc;
  if (a.C1 > 0) {
    rc = make_tmap_nhwc(&tmA1, a.in1, a.N, a.H, a.W, a.C1, a.ld_in1, boxA, 128);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  rc = make_tmap_nhwc(&tmDY, a.dy, a.N, a.H, a.W, a.Cout, a.ld_dy, boxB, p.half * 2);
  if (rc) return rc;

  const int items = groups * splits;
  const int grid = 2 * (items < resident ? items : resident);
  const size_t smem = 1024 + 2 * static_cast<size_t>(p.a_bytes) + static_cast<size_t>(b_stages) * p.b_stage_bytes +
                      sizeof(WgHalo2Header);
  note_variant(14);
  launch_co(conv_wgrad_halo2_kernel, grid, kG2Threads, smem, stream, tmA0, tmA1, tmDY, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (a.splits_used) *a.splits_used = splits;
  return 0;
}

}  // namespace ub2
