// Shared epilogue of the tcgen05 convolution kernels: one 32-column chunk of a 128-row
// accumulator tile: TMEM -> registers -> (affine, accumulate, ReLU) -> bf16 store, plus the
// BatchNorm statistics of the stored values.
#pragma once
#include "conv.h"
#include "ptx.cuh"

namespace ub2 {

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

// taddr: TMEM address of the chunk (lane quarter + column); cbase: first output channel of the
// chunk; n_hi: end of this N tile; pix: output pixel of this thread's row (valid if `valid`).
template <bool ACC, bool TF32 = false>
__device__ __forceinline__ void epi_chunk(const ConvFwdParams& p, uint32_t taddr, int cbase, int n_hi,
                                          bool valid, size_t pix, int lane, bool want_stats,
                                          float* my_stats, float (&acc_s)[ACC ? 32 : 1],
                                          float (&acc_q)[ACC ? 32 : 1]) {
  uint32_t raw[32];
  tmem_ld32(taddr, raw);
  tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
  // destination of each 8-channel vector (split output for the dgrad of a virtual concat)
  __nv_bfloat16* dst[4];
  bool dvalid[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = cbase + g * 8;
    dvalid[g] = valid && (c < p.Cout) && (c < n_hi);
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
  if (TF32) {
    // fp32 activations (evaluation in TF32 mode): values rounded to TF32 where they are produced
    float* dstf = reinterpret_cast<float*>(p.out0) + pix * p.ld0 + cbase;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (valid && cbase + g * 4 < p.Cout && cbase + g * 4 < n_hi) {
        float4 o;
        if (p.tf32 == 2) {   // training (3xTF32): the accumulator is fp32-accurate, keep it
          o.x = v[g * 4 + 0]; o.y = v[g * 4 + 1]; o.z = v[g * 4 + 2]; o.w = v[g * 4 + 3];
        } else {
          o.x = tf32_round(v[g * 4 + 0]);
          o.y = tf32_round(v[g * 4 + 1]);
          o.z = tf32_round(v[g * 4 + 2]);
          o.w = tf32_round(v[g * 4 + 3]);
        }
        *reinterpret_cast<float4*>(dstf + g * 4) = o;
      }
    }
    return;
  }
  uint4 held = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 o;
    o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
    o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
    o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
    o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
    // 16-channel pairs go out as one 256-bit store: a warp's rows are 128+ bytes apart, so a
    // 16-byte store per lane fills half a 32-byte sector and every sector is written twice
    if (p.wide_store) {
      if ((g & 1) == 0) {
        held = o;
      } else if (dvalid[g - 1] && dvalid[g] && dst[g] == dst[g - 1] + 8) {
        st_global_256(dst[g - 1], held, o);
      } else {
        if (dvalid[g - 1]) *reinterpret_cast<uint4*>(dst[g - 1]) = held;
        if (dvalid[g]) *reinterpret_cast<uint4*>(dst[g]) = o;
      }
    } else if (dvalid[g]) {
      *reinterpret_cast<uint4*>(dst[g]) = o;
    }
    if (want_stats) {
      // statistics of the values as stored (bf16-rounded): BatchNorm then normalises
      // exactly the tensor it measured; masked pixels / channels count as 0
      if (dvalid[g]) {
        v[g * 8 + 0] = bf16_lo(o.x);
        v[g * 8 + 1] = bf16_hi(o.x);
        v[g * 8 + 2] = bf16_lo(o.y);
        v[g * 8 + 3] = bf16_hi(o.y);
        v[g * 8 + 4] = bf16_lo(o.z);
        v[g * 8 + 5] = bf16_hi(o.z);
        v[g * 8 + 6] = bf16_lo(o.w);
        v[g * 8 + 7] = bf16_hi(o.w);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[g * 8 + i] = 0.f;
      }
    }
  }
  if (want_stats) {
    if (ACC) {
#pragma unroll
      for (int i = 0; i < (ACC ? 32 : 1); ++i) {
        acc_s[i] += v[i];
        acc_q[i] = fmaf(v[i], v[i], acc_q[i]);
      }
    } else {
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
}

// End of kernel: fold the running sums (ACC) and the four lane quarters into one fp64 row.
template <bool ACC, int EPI_THREADS>
__device__ __forceinline__ void epi_finish(const ConvFwdParams& p, float* s_stats, float* my_stats,
                                           int lane, int grp, int nchunks, int epi_tid,
                                           float (&acc_s)[ACC ? 32 : 1], float (&acc_q)[ACC ? 32 : 1]) {
  if (ACC) {
    // one cross-lane reduction for the whole CTA (this warp owns chunk `grp`)
    float a[32], b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      a[i] = acc_s[ACC ? i : 0];
      b[i] = acc_q[ACC ? i : 0];
    }
    const float s1 = butterfly32(a, lane);
    const float s2 = butterfly32(b, lane);
    const int c = grp * 32 + lane;
    if (grp < nchunks && c < p.Cout) {
      my_stats[c] = s1;
      my_stats[p.Cout + c] = s2;
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
  for (int c = epi_tid; c < p.Cout; c += EPI_THREADS) {
    double s1 = 0.0, s2 = 0.0;
    for (int e = 0; e < 4; ++e) {
      s1 += static_cast<double>(s_stats[e * 2 * p.Cout + c]);
      s2 += static_cast<double>(s_stats[e * 2 * p.Cout + p.Cout + c]);
    }
    p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 0) * p.Cout + c] = s1;
    p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 1) * p.Cout + c] = s2;
  }
}

}  // namespace ub2
