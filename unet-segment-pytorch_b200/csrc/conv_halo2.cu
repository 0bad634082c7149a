// Two-CTA (cta_group::2) variant of the halo-resident 3x3 convolution (conv_halo.cu).
//
// At Cout <= 128 a single-CTA tcgen05.mma is bound by shared-memory reads: a 128 x N x 16 MMA reads
// 4 KB of A and N*32 B of B for 8*N tensor cycles (N = 64: 6 KB / 32 clk, a 67 % ceiling; N = 128:
// exactly the 128 B/clk the SM has, so the TMA fills push it over).  Here two CTAs of a cluster
// (one TPC) work on two neighbouring work items in lock step: each keeps its own halo block (A) and
// only HALF of every weight tile (B rows [r*N/2, (r+1)*N/2)); one tcgen05.mma.cta_group::2 issued by
// the leader computes both CTAs' 128 x N accumulators (M = 256), each SM reading its A and half of
// B: 5 KB (N = 64) / 6 KB (N = 128) per MMA and half the weight-tile TMA traffic.
//
// Protocol (barrier objects live at the same offsets in both CTAs):
//   a_full / b_full   : in the LEADER, count 1: the leader's producer arms it with BOTH CTAs' byte
//                       count, the peer only issues its TMA (cp.async.bulk.tensor ... .cta_group::2
//                       with the leader's barrier address) — no remote arrive, hence no cluster-scope
//                       release fence in the producer loop (that fence made the first version
//                       producer bound: ncu stall "membar")
//   a_empty / b_empty : per CTA, arrived by the leader's tcgen05.commit.multicast (mask 0b11)
//   tmem_full         : per CTA, same multicast commit
//   tmem_empty        : in the LEADER, count 2 x epilogue warps (the peer's warps arrive remotely)
// Work items are paired (2p, 2p+1); an odd tail item gets a phantom partner whose TMA boxes lie
// outside the tensor (zero fill) and whose epilogue stores nothing.
#include "launch.cuh"
#include <cstdlib>
#include "conv.h"
#include "conv_epilogue.cuh"
#include "ptx.cuh"
#include "cluster2.cuh"

namespace ub2 {

static constexpr int kH2EpiWarps = 8;
static constexpr int kH2Threads = 64 + 32 * kH2EpiWarps;
static constexpr int kH2MaxBStages = 12;
static constexpr int kRW2 = 130;

struct Halo2SmemHeader {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kH2MaxBStages], b_empty[kH2MaxBStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base, pad;
};

template <bool ACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kH2Threads, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const ConvFwdParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* sA = sbase;                       // 2 x a_bytes
  uint8_t* sB = sbase + 2 * p.a_bytes;       // b_stages x b_stage_bytes (half weight tiles)
  Halo2SmemHeader* hdr = reinterpret_cast<Halo2SmemHeader*>(sB + p.b_stages * p.b_stage_bytes);
  float* s_stats = reinterpret_cast<float*>(hdr + 1);  // [4 quarters][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int R = p.R;
  const int C0 = p.C0;
  const int Ctot = p.C0 + p.C1;
  const int kchunks = Ctot / 64;
  const int BN = p.BN;                 // whole Cout (n_tiles == 1)
  const int half = BN / 2;
  const int bn_cols = (BN + 31) & ~31;
  const int items = p.N * p.blocks_h * p.segs_w;
  const int pairs = (items + 1) / 2;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.C1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->a_full[i], 1);
      mbar_init(&hdr->a_empty[i], 1);
      mbar_init(&hdr->tmem_full[i], 1);
      mbar_init(&hdr->tmem_empty[i], 2 * kH2EpiWarps);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&hdr->b_full[i], 1);
      mbar_init(&hdr->b_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(&hdr->tmem_base, p.tmem_cols);
  if (warp >= 2 && p.stats != nullptr) {
    for (int i = threadIdx.x - 64; i < 4 * 2 * p.Cout; i += 32 * kH2EpiWarps) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();   // the prologue above touched no global memory; everything below may (launch.cuh)

  // item of this CTA in pair `pr`: (column segment, row block, image); a phantom item (odd tail)
  // gets n = N: every TMA box is out of bounds (zero fill) and nothing is stored
  auto decode = [&](int pr, int& w0, int& h0, int& n) {
    const int t = 2 * pr + static_cast<int>(rank);
    int r = t;
    w0 = (r % p.segs_w) * 128;
    r /= p.segs_w;
    h0 = (r % p.blocks_h) * R;
    n = (t < items) ? r / p.blocks_h : p.N;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t a_tx = 64u * kRW2 * (R + 2) * 2u;
    const uint32_t b_tx = static_cast<uint32_t>(half) * 128u;
    int abuf = 0, bs = 0;
    uint32_t aphase = 0, bphase = 0;
    for (int pr = cluster_id; pr < pairs; pr += n_clusters) {
      int w0, h0, n;
      decode(pr, w0, h0, n);
      for (int c = 0; c < Ctot; c += 64) {
        mbar_wait(&hdr->a_empty[abuf], aphase ^ 1);
        if (elect_one()) {
          const uint32_t bar = mapa_u32(smem_u32(&hdr->a_full[abuf]), 0);
          if (leader) mbar_expect_tx(&hdr->a_full[abuf], 2 * a_tx);
          if (c < C0)
            tma2_load_4d(sA + abuf * p.a_bytes, &tmA0, bar, c, w0 - 1, h0 - 1, n);
          else
            tma2_load_4d(sA + abuf * p.a_bytes, &tmA1, bar, c - C0, w0 - 1, h0 - 1, n);
        }
        if (++abuf == 2) {
          abuf = 0;
          aphase ^= 1;
        }
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&hdr->b_empty[bs], bphase ^ 1);
          if (elect_one()) {
            const uint32_t bar = mapa_u32(smem_u32(&hdr->b_full[bs]), 0);
            if (leader) mbar_expect_tx(&hdr->b_full[bs], 2 * b_tx);
            tma2_load_2d(sB + bs * p.b_stage_bytes, &tmB, bar, tap * Ctot + c, static_cast<int>(rank) * half);
          }
          if (++bs == p.b_stages) {
            bs = 0;
            bphase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------- MMA issuer (leader CTA only)
      const uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, 128B swizzle
      const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_inc = static_cast<uint32_t>(p.a_bytes) >> 4;
      const uint32_t b_inc = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
      int abuf = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      int it = 0;
      for (int pr = cluster_id; pr < pairs; pr += n_clusters, ++it) {
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
              uint32_t a_row = a_lo + static_cast<uint32_t>(((1 + dr) * kRW2 + 1 + ds) * 8);
              for (int r = 0; r < R; ++r, a_row += kRW2 * 8) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_row + 2 * k);
                  const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
                  umma2_bf16(d0 + r * bn_cols, da, db, idesc, (kcidx | tap | k) != 0);
                }
              }
              umma2_commit(&hdr->b_empty[bs]);
              if (tap == 8) {
                umma2_commit(&hdr->a_empty[abuf]);
                if (kcidx == kchunks - 1) umma2_commit(&hdr->tmem_full[as]);
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
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs, own rows)
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
    for (int pr = cluster_id; pr < pairs; pr += n_clusters, ++it) {
      int w0, h0, n;
      decode(pr, w0, h0, n);
      const int as = it & 1;
      mbar_wait(&hdr->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      for (int r = 0; r < R; ++r) {
        const int h = h0 + r;
        const bool valid = h < p.H && n < p.N;
        const size_t pix = (static_cast<size_t>(n < p.N ? n : 0) * p.H + h) * p.W + w0 + row;
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (as * R + r) * bn_cols;
        for (int j = grp; j < nchunks; j += 2)
          epi_chunk<ACC>(p, tcol + j * 32, j * 32, BN, valid, pix, lane, want_stats, my_stats, acc_s, acc_q);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&hdr->tmem_empty[as]), 0));
    }
    if (want_stats)
      epi_finish<ACC, 32 * kH2EpiWarps>(p, s_stats, my_stats, lane, grp, nchunks, threadIdx.x - 64, acc_s, acc_q);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading / signalling this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

// returns 1 if the shape is not eligible (caller falls back to the one-CTA halo kernel)
int conv_halo2_launch(const ConvFwdArgs& a, cudaStream_t stream) {
  static const int enabled = [] { const char* e = getenv("UB2_HALO2"); return e ? atoi(e) : 1; }();
  const int Ctot = a.C0 + a.C1;
  if (!enabled || a.taps != 9 || a.W % 128 != 0 || a.C0 % 64 != 0 || Ctot % 64 != 0) return 1;
  if (a.Cout % 32 != 0 || a.Cout > 128 || a.bn_override > 0 || a.grid_override > 0) return 1;
  const int BN = a.Cout;
  if (a.out1 != nullptr && (a.split % 8 != 0 || a.ld1 % 8 != 0)) return UB2_ERR_ALIGN;
  if (a.ld_in0 % 8 != 0 || (a.C1 > 0 && a.ld_in1 % 8 != 0) || a.ld0 % 8 != 0) return UB2_ERR_ALIGN;
  const int bn_cols = (BN + 31) & ~31;

  ConvFwdParams p{};
  p.N = a.N; p.H = a.H; p.W = a.W; p.C0 = a.C0; p.C1 = a.C1; p.Cout = a.Cout; p.taps = 9; p.kc = 64;
  p.BN = BN;
  p.n_tiles = 1;
  const int stats_bytes = a.stats ? 4 * 2 * a.Cout * 4 : 0;
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(Halo2SmemHeader)) - stats_bytes;
  p.b_stage_bytes = ((BN / 2) * 128 + 1023) & ~1023;   // half a weight tile per CTA
  int R = 256 / bn_cols;  // two accumulator sets of R tiles in 512 TMEM columns
  if (R > 4) R = 4;
  static const int r_env = [] { const char* e = getenv("UB2_HALO2_R"); return e ? atoi(e) : 0; }();
  if (r_env > 0 && r_env < R) R = r_env;   // tuning experiments
  if (R > a.H) R = a.H;
  int b_stages = 0;
  for (; R >= 1; --R) {
    p.a_bytes = ((64 * kRW2 * (R + 2) * 2) + 1023) & ~1023;
    b_stages = (budget - 2 * p.a_bytes) / p.b_stage_bytes;
    if (b_stages >= 5) break;
  }
  if (R < 1) return 1;
  if (b_stages > kH2MaxBStages) b_stages = kH2MaxBStages;
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

  const int items = a.N * p.blocks_h * p.segs_w;
  const int pairs = (items + 1) / 2;

  CUtensorMap tmA0, tmA1, tmB;
  const uint32_t boxA[4] = {64u, static_cast<uint32_t>(kRW2), static_cast<uint32_t>(R + 2), 1u};
  int rc = make_tmap_nhwc(&tmA0, a.in0, a.N, a.H, a.W, a.C0, a.ld_in0, boxA, 128);
  if (rc) return rc;
  if (a.C1 > 0) {
    rc = make_tmap_nhwc(&tmA1, a.in1, a.N, a.H, a.W, a.C1, a.ld_in1, boxA, 128);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  rc = make_tmap_2d(&tmB, a.wgt, 9ull * Ctot, a.Cout, 9ull * Ctot, 64, BN / 2, 128);
  if (rc) return rc;

  const size_t smem = 1024 + 2 * static_cast<size_t>(p.a_bytes) + static_cast<size_t>(b_stages) * p.b_stage_bytes +
                      sizeof(Halo2SmemHeader) + stats_bytes;
  static PerDevice<bool> attr_set_pd;
  bool& attr_set = attr_set_pd.ref();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  static const int verbose = [] { const char* v = getenv("UB2_VERBOSE"); return v ? atoi(v) : 0; }();
  const bool acc = a.stats != nullptr && bn_cols <= 64;
  // Persistent grid = the number of CTA pairs that can be resident at once: a pair needs both SMs
  // of one TPC, and not every TPC of the chip has two (floor-sweeping), so this is < SMs / 2.
  struct Mc2 { int v[2]; int& operator[](int i) { return v[i]; } };
  static PerDevice<Mc2> max_clusters_pd;
  Mc2& max_clusters = max_clusters_pd.ref();
  if (max_clusters[acc] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    cfg.blockDim = dim3(kH2Threads);
    cfg.dynamicSmemBytes = 227 * 1024 - 2048;   // the largest request any shape makes
    int n = 0;
    cudaError_t e = acc ? cudaOccupancyMaxActiveClusters(&n, conv_halo2_kernel<true>, &cfg)
                        : cudaOccupancyMaxActiveClusters(&n, conv_halo2_kernel<false>, &cfg);
    if (e != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      return 1;   // no clusters here: one-CTA kernel
    }
    static const int cap_env = [] { const char* v = getenv("UB2_HALO2_CLUSTERS"); return v ? atoi(v) : 0; }();
    if (cap_env > 0 && cap_env < n) n = cap_env;
    max_clusters[acc] = n;
  }
  int clusters = cap_clusters(max_clusters[acc]);
  if (verbose) fprintf(stderr, "ub2: conv_halo2 max resident clusters %d, pairs %d, R %d, b_stages %d\n", clusters, pairs, R, b_stages);
  if (clusters > pairs) clusters = pairs;
  const int grid = 2 * clusters;
  if (a.stats && grid > a.stats_rows) return UB2_ERR_WORKSPACE;
  note_variant(4);
  if (acc) launch(conv_halo2_kernel<true>, grid, kH2Threads, smem, stream, tmA0, tmA1, tmB, p);
  else launch(conv_halo2_kernel<false>, grid, kH2Threads, smem, stream, tmA0, tmA1, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (a.grid_used) *a.grid_used = grid;
  return 0;
}

}  // namespace ub2
