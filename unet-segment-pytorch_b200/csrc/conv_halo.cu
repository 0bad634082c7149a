// Halo-resident 3x3 implicit-GEMM convolution for the wide layers (W a multiple of 128).
//
// The per-tap kernel (conv_fwd.cu) re-fetches every activation tile once per filter tap: at
// Cout <= 128 its 16 KB A tile per 4 MMAs exceeds what L2 can feed.  Here a work item is R output
// rows x 128 columns of one image: per 64-channel chunk ONE TMA box {64 ch, 130, R+2} lands the
// whole halo block in shared memory, and the nine taps are nine *shifted descriptors* into it —
// the 128B swizzle is a function of absolute shared-memory address bits, so a K-major operand may
// start at any 128-byte row (probed on B200: tools/probe/umma_shift.cu).  Row r of the item and
// tap (dr,ds) read rows [(r+1+dr)*130 + 1+ds, +128) of the block.  Activation traffic per output
// tile drops from 9 x 16 KB to (R+2)/R x 16.6 KB, and each weight tile (tap, chunk) is reused by
// the R accumulators, which all live in TMEM (2 x R x N columns, double buffered across items).
//
// Warp roles, barriers and epilogue are those of conv_fwd.cu.
#include "launch.cuh"
#include <cstdlib>
#include "conv.h"
#include "conv_epilogue.cuh"
#include "ptx.cuh"

namespace ub2 {

static constexpr int kHEpiWarps = 8;
static constexpr int kHThreads = 64 + 32 * kHEpiWarps;
static constexpr int kHMaxBStages = 8;
static constexpr int kRW = 130;  // 128 output columns + 2 halo columns

struct HaloSmemHeader {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kHMaxBStages], b_empty[kHMaxBStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base, pad;
};

static int g_conv_mode = 0;
int g_conv_mode_wgrad = 0;
void conv_set_mode(int mode) {
  g_conv_mode = mode;
  g_conv_mode_wgrad = mode;
}

template <bool ACC>
__global__ void __launch_bounds__(kHThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const ConvFwdParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* sA = sbase;                       // 2 x a_bytes
  uint8_t* sB = sbase + 2 * p.a_bytes;       // b_stages x b_stage_bytes
  HaloSmemHeader* hdr = reinterpret_cast<HaloSmemHeader*>(sB + p.b_stages * p.b_stage_bytes);
  float* s_stats = reinterpret_cast<float*>(hdr + 1);  // [4 quarters][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.R;
  const int C0 = p.C0;
  const int Ctot = p.C0 + p.C1;
  const int kchunks = Ctot / 64;
  const int BN = p.BN;
  const int bn_cols = (BN + 31) & ~31;
  const int n_tiles = p.n_tiles;
  const int items = p.N * p.blocks_h * p.segs_w * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.C1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->a_full[i], 1);
      mbar_init(&hdr->a_empty[i], 1);
      mbar_init(&hdr->tmem_full[i], 1);
      mbar_init(&hdr->tmem_empty[i], kHEpiWarps);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&hdr->b_full[i], 1);
      mbar_init(&hdr->b_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, p.tmem_cols);
  if (warp >= 2 && p.stats != nullptr) {
    for (int i = threadIdx.x - 64; i < 4 * 2 * p.Cout; i += 32 * kHEpiWarps) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();   // the prologue above touched no global memory; everything below may (launch.cuh)

  // item -> (n tile, column segment, row block, image); n tile fastest: neighbours share the A block
  auto decode = [&](int t, int& nt, int& w0, int& h0, int& n) {
    nt = t % n_tiles;
    int r = t / n_tiles;
    w0 = (r % p.segs_w) * 128;
    r /= p.segs_w;
    h0 = (r % p.blocks_h) * R;
    n = r / p.blocks_h;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const uint32_t a_tx = 64u * kRW * (R + 2) * 2u;
    const uint32_t b_tx = static_cast<uint32_t>(BN) * 128u;
    int abuf = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int t = blockIdx.x; t < items; t += gridDim.x) {
      int nt, w0, h0, n;
      decode(t, nt, w0, h0, n);
      const int n0 = nt * BN;
      for (int c = 0; c < Ctot; c += 64) {
        mbar_wait(&hdr->a_empty[abuf], aphase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&hdr->a_full[abuf], a_tx);
          if (c < C0)
            tma_load_4d(sA + abuf * p.a_bytes, &tmA0, &hdr->a_full[abuf], c, w0 - 1, h0 - 1, n);
          else
            tma_load_4d(sA + abuf * p.a_bytes, &tmA1, &hdr->a_full[abuf], c - C0, w0 - 1, h0 - 1, n);
        }
        if (++abuf == 2) {
          abuf = 0;
          aphase ^= 1;
        }
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&hdr->b_empty[bs], bphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&hdr->b_full[bs], b_tx);
            tma_load_2d(sB + bs * p.b_stage_bytes, &tmB, &hdr->b_full[bs], tap * Ctot + c, n0);
          }
          if (++bs == p.b_stages) {
            bs = 0;
            bphase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, 128B swizzle
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_inc = static_cast<uint32_t>(p.a_bytes) >> 4;
    const uint32_t b_inc = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
    int abuf = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < items; t += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&hdr->tmem_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + as * R * bn_cols;
      for (int kcidx = 0; kcidx < kchunks; ++kcidx) {
        mbar_wait(&hdr->a_full[abuf], aphase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + abuf * a_inc;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&hdr->b_full[bs], bphase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = b_lo0 + bs * b_inc;
            const int dr = tap / 3 - 1, ds = tap % 3 - 1;
            // 16-byte units: one block row = 130 * 8, one pixel = 8
            uint32_t a_row = a_lo + static_cast<uint32_t>(((1 + dr) * kRW + 1 + ds) * 8);
            for (int r = 0; r < R; ++r, a_row += kRW * 8) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_row + 2 * k);
                const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
                umma_bf16(d0 + r * bn_cols, da, db, idesc, (kcidx | tap | k) != 0);
              }
            }
            umma_commit(&hdr->b_empty[bs]);
            if (tap == 8) {
              umma_commit(&hdr->a_empty[abuf]);
              if (kcidx == kchunks - 1) umma_commit(&hdr->tmem_full[as]);
            }
          }
          __syncwarp();
          if (++bs == p.b_stages) {
            bs = 0;
            bphase ^= 1;
          }
        }
        if (++abuf == 2) {
          abuf = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;  // output column within the segment
    const int nchunks = bn_cols / 32;
    const bool want_stats = p.stats != nullptr;
    float* my_stats = s_stats + q * 2 * p.Cout;
    float acc_s[ACC ? 32 : 1], acc_q[ACC ? 32 : 1];
    if (ACC) {
#pragma unroll
      for (int i = 0; i < (ACC ? 32 : 1); ++i) acc_s[i] = acc_q[i] = 0.f;
    }
    int it = 0;
    for (int t = blockIdx.x; t < items; t += gridDim.x, ++it) {
      int nt, w0, h0, n;
      decode(t, nt, w0, h0, n);
      const int n0 = nt * BN;
      const int as = it & 1;
      mbar_wait(&hdr->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      for (int r = 0; r < R; ++r) {
        const int h = h0 + r;
        const bool valid = h < p.H;
        const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w0 + row;
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (as * R + r) * bn_cols;
        for (int j = grp; j < nchunks; j += 2)
          epi_chunk<ACC>(p, tcol + j * 32, n0 + j * 32, n0 + BN, valid, pix, lane, want_stats, my_stats,
                         acc_s, acc_q);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->tmem_empty[as]);
    }
    if (want_stats)
      epi_finish<ACC, 32 * kHEpiWarps>(p, s_stats, my_stats, lane, grp, nchunks, threadIdx.x - 64, acc_s, acc_q);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int conv_halo2_launch(const ConvFwdArgs& a, cudaStream_t stream);   // conv_halo2.cu: cta_group::2 variant

int conv_halo_launch(const ConvFwdArgs& a, cudaStream_t stream) {
  if (g_conv_mode != 1) {
    const int rc2 = conv_halo2_launch(a, stream);
    if (rc2 != 1) return rc2;
  }
  const int Ctot = a.C0 + a.C1;
  if (g_conv_mode == 1 || a.taps != 9 || a.W % 128 != 0 || a.C0 % 64 != 0 || Ctot % 64 != 0) return 1;
  // Cout > 128: the per-tap kernel with a 256-wide N tile is already tensor-bound (measured)
  if (a.Cout % 16 != 0 || a.Cout > 128 || a.bn_override > 0) return 1;
  const int BN = a.Cout;
  if (a.out1 != nullptr && (a.split % 8 != 0 || a.ld1 % 8 != 0)) return UB2_ERR_ALIGN;
  if (a.ld_in0 % 8 != 0 || (a.C1 > 0 && a.ld_in1 % 8 != 0) || a.ld0 % 8 != 0) return UB2_ERR_ALIGN;
  const int bn_cols = (BN + 31) & ~31;

  ConvFwdParams p{};
  p.N = a.N; p.H = a.H; p.W = a.W; p.C0 = a.C0; p.C1 = a.C1; p.Cout = a.Cout; p.taps = 9; p.kc = 64;
  p.BN = BN;
  p.n_tiles = (a.Cout + BN - 1) / BN;
  const int stats_bytes = a.stats ? 4 * 2 * a.Cout * 4 : 0;
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(HaloSmemHeader)) - stats_bytes;
  p.b_stage_bytes = BN * 128;  // BN % 16 == 0 and kc == 64: a multiple of 1024 only if BN % 8 == 0
  p.b_stage_bytes = (p.b_stage_bytes + 1023) & ~1023;
  int R = 256 / bn_cols;  // two accumulator sets of R tiles in 512 TMEM columns
  if (R > 3) R = 3;       // R = 4 leaves room for only 3 weight stages: the B ring starves (measured)
  static const int r_env = [] { const char* e = getenv("UB2_HALO_R"); return e ? atoi(e) : 0; }();
  if (r_env > 0 && r_env < R) R = r_env;   // tuning experiments
  if (R > a.H) R = a.H;
  int b_stages = 0;
  for (; R >= 1; --R) {
    p.a_bytes = ((64 * kRW * (R + 2) * 2) + 1023) & ~1023;
    b_stages = (budget - 2 * p.a_bytes) / p.b_stage_bytes;
    if (b_stages >= 3) break;
  }
  if (R < 1) return 1;
  if (b_stages > kHMaxBStages) b_stages = kHMaxBStages;
  p.R = R;
  p.b_stages = b_stages;
  p.segs_w = a.W / 128;
  p.blocks_h = (a.H + R - 1) / R;
  int tmem_cols = 32;
  while (tmem_cols < 2 * R * bn_cols) tmem_cols *= 2;
  p.tmem_cols = tmem_cols;
  const int split = (a.out1 != nullptr) ? a.split : (1 << 30);
  p.out0 = reinterpret_cast<__nv_bfloat16*>(a.out0); p.ld0 = a.ld0;
  p.out1 = reinterpret_cast<__nv_bfloat16*>(a.out1); p.ld1 = a.ld1; p.split = split;
  p.accumulate = a.accumulate;
  p.scale = a.scale; p.shift = a.shift; p.relu = a.relu;
  p.stats = a.stats;
  {
    static const int wide_env = [] { const char* e = getenv("UB2_WIDE_STORE"); return e ? atoi(e) : 1; }();
    p.wide_store = wide_env && conv_wide_store_ok(a.out0, a.ld0, a.out1, a.ld1, a.split, a.Cout);
  }

  CUtensorMap tmA0, tmA1, tmB;
  const uint32_t boxA[4] = {64u, static_cast<uint32_t>(kRW), static_cast<uint32_t>(R + 2), 1u};
  int rc = make_tmap_nhwc(&tmA0, a.in0, a.N, a.H, a.W, a.C0, a.ld_in0, boxA, 128);
  if (rc) return rc;
  if (a.C1 > 0) {
    rc = make_tmap_nhwc(&tmA1, a.in1, a.N, a.H, a.W, a.C1, a.ld_in1, boxA, 128);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  rc = make_tmap_2d(&tmB, a.wgt, 9ull * Ctot, a.Cout, 9ull * Ctot, 64, BN, 128);
  if (rc) return rc;

  const int items = a.N * p.blocks_h * p.segs_w * p.n_tiles;
  int grid = num_sms();
  if (a.grid_override > 0) grid = a.grid_override;
  if (grid > items) grid = items;
  if (a.stats && grid > a.stats_rows) return UB2_ERR_WORKSPACE;
  const size_t smem = 1024 + 2 * static_cast<size_t>(p.a_bytes) + static_cast<size_t>(b_stages) * p.b_stage_bytes +
                      sizeof(HaloSmemHeader) + stats_bytes;
  static PerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.ref();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const bool acc = a.stats != nullptr && bn_cols <= 64 && p.n_tiles == 1;
  note_variant(3);
  if (acc) launch(conv_halo_kernel<true>, grid, kHThreads, smem, stream, tmA0, tmA1, tmB, p);
  else launch(conv_halo_kernel<false>, grid, kHThreads, smem, stream, tmA0, tmA1, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (a.grid_used) *a.grid_used = grid;
  return 0;
}

}  // namespace ub2
