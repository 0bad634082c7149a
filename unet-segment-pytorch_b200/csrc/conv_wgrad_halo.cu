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
#include "conv.h"
#include "ptx.cuh"

namespace ub2 {

static constexpr int kGThreads = 192;
static constexpr int kGMaxBStages = 8;
static constexpr int kGRW = 130;

struct WgHaloParams {
  int N, H, W, C0, C1, Cout, R;
  int bn_cols, kchunks, tapgroups, splits, per_split, blocks, segs_w, blocks_h;
  int a_bytes, b_stage_bytes, b_stages, b_subs, tmem_cols;
  float* partial;
};

struct WgHaloHeader {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kGMaxBStages], b_empty[kGMaxBStages];
  uint64_t tmem_full, tmem_empty;
  uint32_t tmem_base, pad;
};

// offset (in 16-byte units) of tap t's window inside the halo block: ((1+dr)*130 + 1+ds) pixels
__device__ __forceinline__ int tap_off(int t) { return ((t / 3) * kGRW + (t % 3)) * 8; }

__global__ void __launch_bounds__(kGThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmDY, const WgHaloParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* sA = sbase;
  uint8_t* sB = sbase + 2 * p.a_bytes;
  WgHaloHeader* hdr = reinterpret_cast<WgHaloHeader*>(sB + p.b_stages * p.b_stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.R;
  const int Ctot = p.C0 + p.C1;
  const int Mtot = 9 * Ctot;
  const int per_item_groups = p.kchunks * p.tapgroups;
  const int items = per_item_groups * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.C1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->a_full[i], 1);
      mbar_init(&hdr->a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&hdr->b_full[i], 1);
      mbar_init(&hdr->b_empty[i], 1);
    }
    mbar_init(&hdr->tmem_full, 1);
    mbar_init(&hdr->tmem_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();   // the prologue above touched no global memory; everything below may (launch.cuh)

  // item -> (channel chunk, tap group, split); the items of one split are neighbours (share dy in L2)
  auto decode_item = [&](int it, int& c, int& t_begin, int& t_end, int& sp) {
    sp = it / per_item_groups;
    const int g = it % per_item_groups;
    c = (g % p.kchunks) * 64;
    const int tg = g / p.kchunks;
    if (p.tapgroups == 1) { t_begin = 0; t_end = 9; }
    else if (tg == 0) { t_begin = 0; t_end = 6; }
    else { t_begin = 6; t_end = 9; }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const uint32_t a_tx = 64u * kGRW * (R + 2) * 2u;
    const uint32_t b_tx = static_cast<uint32_t>(p.b_subs) * 64u * 128u;
    int abuf = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
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
          mbar_expect_tx(&hdr->a_full[abuf], a_tx);
          if (c < p.C0)
            tma_load_4d(sA + abuf * p.a_bytes, &tmA0, &hdr->a_full[abuf], c, w0 - 1, h0 - 1, n);
          else
            tma_load_4d(sA + abuf * p.a_bytes, &tmA1, &hdr->a_full[abuf], c - p.C0, w0 - 1, h0 - 1, n);
        }
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        for (int kc = 0; kc < 2 * R; ++kc) {  // 64-pixel pieces: row kc/2, half kc%2
          mbar_wait(&hdr->b_empty[bs], bphase ^ 1);
          if (lane == 0) mbar_expect_tx(&hdr->b_full[bs], b_tx);
          __syncwarp();
          if (lane < p.b_subs)
            tma_load_4d(sB + bs * p.b_stage_bytes + lane * 8192, &tmDY, &hdr->b_full[bs], lane * 64,
                        w0 + (kc & 1) * 64, h0 + (kc >> 1), n);
          if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        }
        if (++seg == p.segs_w) {
          seg = 0;
          if (++hb == p.blocks_h) { hb = 0; ++n; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, p.bn_cols >= p.Cout ? p.Cout : p.bn_cols, 1, 1);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_base = (smem_u32(sA) & 0x3FFFFu) >> 4;
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);
    const uint32_t a_inc = static_cast<uint32_t>(p.a_bytes) >> 4;
    const uint32_t b_inc = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
    int abuf = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    int n_it = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_it) {
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
            const uint32_t kpix = static_cast<uint32_t>(((kc >> 1) * kGRW + (kc & 1) * 64) * 8);
            for (int pr = 0; pr < npairs; ++pr) {
              const int t0 = t_begin + 2 * pr;
              const int off0 = tap_off(t0);
              const int lbo = (t0 + 1 < t_end) ? tap_off(t0 + 1) - off0 : 8;  // single tap: any valid atom
              const uint32_t a_lo = ((a_buf + kpix + off0) & 0x3FFFu) | (static_cast<uint32_t>(lbo) << 16);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k * 128);
                const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k * 128);
                umma_bf16(tmem_base + pr * p.bn_cols, da, db, idesc, (blk > blk_begin) || (kc > 0) || (k > 0));
              }
            }
            umma_commit(&hdr->b_empty[bs]);
            if (kc == 2 * R - 1) {
              umma_commit(&hdr->a_empty[abuf]);
              if (blk == blk_end - 1) umma_commit(&hdr->tmem_full);
            }
          }
          __syncwarp();
          if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
        }
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;
    const int row = q * 32 + lane;  // TMEM lane: rows 0..63 = first tap of the pair, 64..127 = second
    int n_it = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_it) {
      int c, t_begin, t_end, sp;
      decode_item(it, c, t_begin, t_end, sp);
      const int npairs = (t_end - t_begin + 1) / 2;
      mbar_wait(&hdr->tmem_full, n_it & 1);
      tc_fence_after();
      for (int pr = 0; pr < npairs; ++pr) {
        const int tap = t_begin + 2 * pr + (row >> 6);
        const bool rvalid = tap < t_end;
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
      if (lane == 0) mbar_arrive(&hdr->tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// returns 1 if the shape is not eligible (caller falls back to the per-tap kernel)
int conv_wgrad_halo2_launch(const ConvWgradArgs& a, cudaStream_t stream);   // conv_wgrad_halo2.cu

int conv_wgrad_halo_launch(const ConvWgradArgs& a, cudaStream_t stream) {
  {
    const int rc2 = conv_wgrad_halo2_launch(a, stream);
    if (rc2 != 1) return rc2;
  }
  const int Ctot = a.C0 + a.C1;
  if (a.taps != 9 || a.W % 128 != 0 || a.C0 % 64 != 0 || Ctot % 64 != 0) return 1;
  if (a.Cout % 16 != 0 || a.Cout > 128 || a.splits_override > 0) return 1;
  if (a.ld_in0 % 8 != 0 || (a.C1 > 0 && a.ld_in1 % 8 != 0) || a.ld_dy % 8 != 0) return UB2_ERR_ALIGN;

  WgHaloParams p{};
  p.N = a.N; p.H = a.H; p.W = a.W; p.C0 = a.C0; p.C1 = a.C1; p.Cout = a.Cout;
  p.bn_cols = (a.Cout + 31) & ~31;
  p.kchunks = Ctot / 64;
  p.tapgroups = (5 * p.bn_cols <= 512) ? 1 : 2;   // five / three 128-row tiles of bn_cols columns in TMEM
  const int tiles = p.tapgroups == 1 ? 5 : 3;
  int tmem_cols = 32;
  while (tmem_cols < tiles * p.bn_cols) tmem_cols *= 2;
  p.tmem_cols = tmem_cols;
  p.b_subs = (a.Cout + 63) / 64;
  p.b_stage_bytes = p.b_subs * 8192;
  const int budget = wgrad_smem_budget() - 1024 - static_cast<int>(sizeof(WgHaloHeader));
  int R = a.H < 2 ? a.H : 2;
  p.a_bytes = ((64 * kGRW * (R + 2) * 2) + 1023) & ~1023;
  int b_stages = (budget - 2 * p.a_bytes) / p.b_stage_bytes;
  if (b_stages < 2) return 1;
  if (b_stages > kGMaxBStages) b_stages = kGMaxBStages;
  p.R = R;
  p.b_stages = b_stages;
  p.segs_w = a.W / 128;
  p.blocks_h = (a.H + R - 1) / R;
  p.blocks = a.N * p.blocks_h * p.segs_w;
  const int groups = p.kchunks * p.tapgroups;
  int splits = num_sms() / groups;
  if (splits < 1) splits = 1;
  if (splits > p.blocks) splits = p.blocks;
  if (splits > a.max_splits) splits = a.max_splits;
  p.per_split = (p.blocks + splits - 1) / splits;
  splits = (p.blocks + p.per_split - 1) / p.per_split;
  p.splits = splits;
  p.partial = a.partial;

  CUtensorMap tmA0, tmA1, tmDY;
  const uint32_t boxA[4] = {64u, static_cast<uint32_t>(kGRW), static_cast<uint32_t>(R + 2), 1u};
  const uint32_t boxB[4] = {64u, 64u, 1u, 1u};
  int rc = make_tmap_nhwc(&tmA0, a.in0, a.N, a.H, a.W, a.C0, a.ld_in0, boxA, 128);
  if (rc) return rc;
  if (a.C1 > 0) {
    rc = make_tmap_nhwc(&tmA1, a.in1, a.N, a.H, a.W, a.C1, a.ld_in1, boxA, 128);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  rc = make_tmap_nhwc(&tmDY, a.dy, a.N, a.H, a.W, a.Cout, a.ld_dy, boxB, 128);
  if (rc) return rc;

  const int items = groups * splits;
  const int grid = items < num_sms() ? items : num_sms();
  const size_t smem = 1024 + 2 * static_cast<size_t>(p.a_bytes) + static_cast<size_t>(b_stages) * p.b_stage_bytes +
                      sizeof(WgHaloHeader);
  static PerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.ref();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  note_variant(13);
  launch_co(conv_wgrad_halo_kernel, grid, kGThreads, smem, stream, tmA0, tmA1, tmDY, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (a.splits_used) *a.splits_used = splits;
  return 0;
}

}  // namespace ub2
