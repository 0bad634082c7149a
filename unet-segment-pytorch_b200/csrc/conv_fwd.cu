// NHWC bf16 implicit-GEMM convolution (3x3 pad 1, or 1x1) on tcgen05 tensor cores.
//
//   out[n,h,w,co] = sum_{tap,c} in[n, h+r-1, w+s-1, c] * wgt[co, tap, c]
//
// GEMM view: M = pixels (tiles of 128 = BW x BH x BI box of one TMA load),
// N = Cout (tile BN <= 256), K = taps * (C0 + C1) walked as (tap, channel chunk).
// The channel axis may span two source tensors (skip, upsampled): the
// reference's torch.cat([x2, x1], dim=1) (layers.py:105, :254) is never
// materialised.  The same kernel is the dgrad pass when given the
// tap-flipped / transposed weight pack and the output split over two tensors.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
// (and TMEM owner), warps 2..5 = epilogue.  smem ring of `stages` (A,B) tiles,
// two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Epilogue: optional per-channel affine + ReLU (BN folded for
// inference), optional accumulate into the destination, bf16 store, and
// per-channel sum / sum-of-squares of the rounded outputs (BatchNorm batch
// statistics, layers.py:33) reduced with a register butterfly.
#include "conv.h"
#include "ptx.cuh"

namespace ub2 {

static constexpr int kMaxStages = 8;
static constexpr int kThreads = 192;
static constexpr int kATileBytes = 128 * 64 * 2;  // 16 KB slot

struct FwdSmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// Reduce 32 columns across the 32 lanes of a warp: on return lane l holds the
// column-l total in v[0].  31 shuffles instead of 32*5.
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int k = 0; k < m; ++k) {
      float keep = up ? v[k + m] : v[k];
      float send = up ? v[k] : v[k + m];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const ConvFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-align the tile ring (128B swizzle atoms repeat every 1024 B).
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  const int stage_bytes = p.stage_bytes;
  FwdSmemHeader* hdr = reinterpret_cast<FwdSmemHeader*>(tiles + p.stages * stage_bytes);
  float* s_stats = reinterpret_cast<float*>(hdr + 1);  // [4 warps][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Ctot = p.C0 + p.C1;
  const int kchunks = Ctot / p.kc;
  const int ksteps = p.taps * kchunks;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = tiles_m * p.n_tiles;
  const int bn_cols = (p.BN + 31) & ~31;  // TMEM columns per accumulator buffer

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.C1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&hdr->full[i], 1);
      mbar_init(&hdr->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->tmem_full[i], 1);
      mbar_init(&hdr->tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, p.tmem_cols);
  if (warp >= 2 && p.stats != nullptr) {
    for (int i = threadIdx.x - 64; i < 4 * 2 * p.Cout; i += 128) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx_bytes = 128u * p.kc * 2u + static_cast<uint32_t>(p.BN) * p.kc * 2u;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles;
        const int mt = t / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.BW;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.BH;
        const int i0 = (mt / (p.tiles_w * p.tiles_h)) * p.BI;
        const int n0 = nt * p.BN;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kcidx = 0; kcidx < kchunks; ++kcidx) {
            const int c = kcidx * p.kc;
            mbar_wait(&hdr->empty[stage], phase ^ 1);
            uint8_t* sa = tiles + stage * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            mbar_expect_tx(&hdr->full[stage], tx_bytes);
            if (c < p.C0)
              tma_load_4d(sa, &tmA0, &hdr->full[stage], c, w0 + ds, h0 + dr, i0);
            else
              tma_load_4d(sa, &tmA1, &hdr->full[stage], c - p.C0, w0 + ds, h0 + dr, i0);
            tma_load_2d(sb, &tmB, &hdr->full[stage], tap * Ctot + c, n0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
      const uint32_t sbo = 16u * p.kc;  // 8 rows of kc bf16
      const uint32_t ltype = (p.kc == 64) ? 2u : (p.kc == 32) ? 4u : 6u;
      const int kinner = p.kc / 16;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(&hdr->tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * bn_cols;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&hdr->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * stage_bytes);
          const uint32_t sb = sa + kATileBytes;
          for (int k = 0; k < kinner; ++k) {
            const uint64_t da = make_smem_desc(sa + k * 32, 16, sbo, ltype);
            const uint64_t db = make_smem_desc(sb + k * 32, 16, sbo, ltype);
            umma_bf16(d_tmem, da, db, idesc, (ks | k) != 0);
          }
          umma_commit(&hdr->empty[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&hdr->tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int ew = warp - 2;
    const int row = q * 32 + lane;
    const int w_l = row % p.BW;
    const int h_l = (row / p.BW) % p.BH;
    const int i_l = row / (p.BW * p.BH);
    const int nchunks = bn_cols / 32;
    float* my_stats = s_stats + ew * 2 * p.Cout;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int nt = t % p.n_tiles;
      const int mt = t / p.n_tiles;
      const int w = (mt % p.tiles_w) * p.BW + w_l;
      const int h = ((mt / p.tiles_w) % p.tiles_h) * p.BH + h_l;
      const int n = (mt / (p.tiles_w * p.tiles_h)) * p.BI + i_l;
      const int n0 = nt * p.BN;
      const bool valid = (w < p.W) && (h < p.H) && (n < p.N);
      const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
      const int as = it & 1;
      mbar_wait(&hdr->tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      for (int j = 0; j < nchunks; ++j) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * bn_cols + j * 32, raw);
        tmem_ld_wait();
        const int cbase = n0 + j * 32;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        // destination of each 8-channel vector (split output for the dgrad of a virtual concat)
        __nv_bfloat16* dst[4];
        bool dvalid[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = cbase + g * 8;
          dvalid[g] = valid && (c < p.Cout) && (c < n0 + p.BN);
          if (c < p.split)
            dst[g] = p.out0 + pix * p.ld0 + c;
          else
            dst[g] = p.out1 + pix * p.ld1 + (c - p.split);
        }
        if (p.scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = min(cbase + i, p.Cout - 1);
            v[i] = fmaf(v[i], __ldg(p.scale + c), __ldg(p.shift + c));
          }
        }
        if (p.accumulate) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (dvalid[g]) {
              const uint4 o = *reinterpret_cast<const uint4*>(dst[g]);
              v[g * 8 + 0] += bf16_lo(o.x);
              v[g * 8 + 1] += bf16_hi(o.x);
              v[g * 8 + 2] += bf16_lo(o.y);
              v[g * 8 + 3] += bf16_hi(o.y);
              v[g * 8 + 4] += bf16_lo(o.z);
              v[g * 8 + 5] += bf16_hi(o.z);
              v[g * 8 + 6] += bf16_lo(o.w);
              v[g * 8 + 7] += bf16_hi(o.w);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
          o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
          o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
          o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
          if (dvalid[g]) *reinterpret_cast<uint4*>(dst[g]) = o;
          if (p.stats != nullptr) {
            // statistics of the values as stored (bf16-rounded), zero for masked pixels
            const float m = dvalid[g] ? 1.f : 0.f;
            v[g * 8 + 0] = m * bf16_lo(o.x);
            v[g * 8 + 1] = m * bf16_hi(o.x);
            v[g * 8 + 2] = m * bf16_lo(o.y);
            v[g * 8 + 3] = m * bf16_hi(o.y);
            v[g * 8 + 4] = m * bf16_lo(o.z);
            v[g * 8 + 5] = m * bf16_hi(o.z);
            v[g * 8 + 6] = m * bf16_lo(o.w);
            v[g * 8 + 7] = m * bf16_hi(o.w);
          }
        }
        if (p.stats != nullptr) {
          float sq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
          const float s1 = butterfly32(v, lane);
          const float s2 = butterfly32(sq, lane);
          const int c = cbase + lane;
          if (c < p.Cout) {
            my_stats[c] += s1;
            my_stats[p.Cout + c] += s2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->tmem_empty[as]);
    }
    if (p.stats != nullptr) {
      // combine the four epilogue warps, one double pair per channel per CTA
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = threadIdx.x - 64; c < p.Cout; c += 128) {
        double s1 = 0.0, s2 = 0.0;
        for (int e = 0; e < 4; ++e) {
          s1 += static_cast<double>(s_stats[e * 2 * p.Cout + c]);
          s2 += static_cast<double>(s_stats[e * 2 * p.Cout + p.Cout + c]);
        }
        p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 0) * p.Cout + c] = s1;
        p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 1) * p.Cout + c] = s2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------
// host side

static int pow2_floor(int x) {
  int r = 1;
  while (r * 2 <= x) r *= 2;
  return r;
}

int conv_fwd_launch(const ConvFwdArgs& a, cudaStream_t stream) {
  const int Ctot = a.C0 + a.C1;
  if (a.taps != 1 && a.taps != 9) return UB2_ERR_SHAPE;
  if (a.N <= 0 || a.H <= 0 || a.W <= 0 || a.C0 <= 0 || a.C1 < 0) return UB2_ERR_SHAPE;
  if (Ctot % 16 != 0 || a.C0 % 16 != 0 || a.Cout % 16 != 0) return UB2_ERR_SHAPE;
  if (a.ld_in0 % 8 != 0 || (a.C1 > 0 && a.ld_in1 % 8 != 0) || a.ld0 % 8 != 0) return UB2_ERR_ALIGN;
  int kc = 64;
  while (a.C0 % kc != 0 || Ctot % kc != 0) kc /= 2;

  ConvFwdParams p{};
  p.N = a.N; p.H = a.H; p.W = a.W;
  p.C0 = a.C0; p.C1 = a.C1; p.Cout = a.Cout;
  p.taps = a.taps; p.kc = kc;
  p.BW = pow2_floor(a.W < 128 ? a.W : 128);
  p.BH = pow2_floor(a.H < 128 / p.BW ? a.H : 128 / p.BW);
  p.BI = 128 / (p.BW * p.BH);
  p.tiles_w = (a.W + p.BW - 1) / p.BW;
  p.tiles_h = (a.H + p.BH - 1) / p.BH;
  p.tiles_n = (a.N + p.BI - 1) / p.BI;
  // N tile: whole Cout when it fits one accumulator buffer; must divide the split point.
  int BN = a.Cout <= 256 ? a.Cout : 256;
  const int split = (a.out1 != nullptr) ? a.split : (1 << 30);
  if (a.out1 != nullptr) {
    if (a.split % 8 != 0 || a.ld1 % 8 != 0) return UB2_ERR_ALIGN;
  }
  if (a.bn_override > 0) BN = a.bn_override;
  if (BN % 16 != 0 || BN > 256) return UB2_ERR_SHAPE;
  p.BN = BN;
  p.n_tiles = (a.Cout + BN - 1) / BN;
  const int bn_cols = (BN + 31) & ~31;
  int tmem_cols = 32;
  while (tmem_cols < 2 * bn_cols) tmem_cols *= 2;
  p.tmem_cols = tmem_cols;
  const int b_bytes = ((BN * kc * 2) + 1023) & ~1023;
  p.stage_bytes = kATileBytes + b_bytes;
  const int stats_bytes = a.stats ? 4 * 2 * a.Cout * 4 : 0;
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(FwdSmemHeader)) - stats_bytes;
  int stages = budget / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return UB2_ERR_SHAPE;
  p.stages = stages;
  p.out0 = reinterpret_cast<__nv_bfloat16*>(a.out0); p.ld0 = a.ld0;
  p.out1 = reinterpret_cast<__nv_bfloat16*>(a.out1); p.ld1 = a.ld1; p.split = split;
  p.accumulate = a.accumulate;
  p.scale = a.scale; p.shift = a.shift; p.relu = a.relu;
  p.stats = a.stats;

  CUtensorMap tmA0, tmA1, tmB;
  const uint32_t boxA[4] = {static_cast<uint32_t>(kc), static_cast<uint32_t>(p.BW),
                            static_cast<uint32_t>(p.BH), static_cast<uint32_t>(p.BI)};
  int rc = make_tmap_nhwc(&tmA0, a.in0, a.N, a.H, a.W, a.C0, a.ld_in0, boxA, kc * 2);
  if (rc) return rc;
  if (a.C1 > 0) {
    rc = make_tmap_nhwc(&tmA1, a.in1, a.N, a.H, a.W, a.C1, a.ld_in1, boxA, kc * 2);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  rc = make_tmap_2d(&tmB, a.wgt, static_cast<uint64_t>(a.taps) * Ctot, a.Cout,
                    static_cast<uint64_t>(a.taps) * Ctot, kc, BN, kc * 2);
  if (rc) return rc;

  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  int grid = num_sms();
  if (a.grid_override > 0) grid = a.grid_override;
  if (grid > total_tiles) grid = total_tiles;
  if (a.stats && grid > a.stats_rows) return UB2_ERR_WORKSPACE;
  const size_t smem = 1024 + static_cast<size_t>(stages) * p.stage_bytes + sizeof(FwdSmemHeader) +
                      stats_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  conv_fwd_kernel<<<grid, kThreads, smem, stream>>>(tmA0, tmA1, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  if (a.grid_used) *a.grid_used = grid;
  return 0;
}

}  // namespace ub2
