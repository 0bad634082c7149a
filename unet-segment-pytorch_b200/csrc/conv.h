// Internal (C++) interface between the C-ABI layer (capi.cu) and the tcgen05
// convolution kernels.  Not part of the public ABI: see include/unetb200.h.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define UB2_ERR_SHAPE (-1)      /* unsupported shape */
#define UB2_ERR_ALIGN (-2)      /* pointer / stride alignment */
#define UB2_ERR_WORKSPACE (-3)  /* workspace too small */
#define UB2_ERR_DRIVER (-4)     /* driver entry point / tensor-map encode failed */
#define UB2_ERR_ARCH (-5)       /* not an sm_100 device */

namespace ub2 {

// ---- forward / dgrad implicit GEMM -------------------------------------------------------
struct ConvFwdArgs {
  const void* in0;  // (N,H,W,C0) bf16, channel stride ld_in0
  const void* in1;  // optional second source (N,H,W,C1) — virtual concat — or null
  const void* wgt;  // (Cout, taps, C0+C1) bf16
  void* out0;       // (N,H,W,*) bf16, channels [0,split) of the result, channel stride ld0
  void* out1;       // optional: channels [split,Cout) go here (index c-split), stride ld1
  const float* scale;  // optional per-channel affine applied to the accumulator
  const float* shift;
  double* stats;    // optional (stats_rows,2,Cout) per-CTA sum / sum of squares
  int* grid_used;   // optional out: number of stats rows written
  int N, H, W, C0, C1, Cout, taps;
  int ld_in0, ld_in1, ld0, ld1, split;
  int accumulate, relu;
  int stats_rows;
  int bn_override, grid_override;  // tuning / tests; 0 = automatic
  int tf32;  // fp32 NHWC activations, fp32 (Cout,taps,Cin) weights, fp32 output: kind::tf32; 1 = output rounded to
             // TF32 (it feeds the next single-pass convolution: eval), 2 = output kept fp32 (3xTF32 training)
};

struct ConvFwdParams {
  int N, H, W, C0, C1, Cout, taps, kc;
  int BW, BH, BI, tiles_w, tiles_h, tiles_n;
  int BN, n_tiles, stages, stage_bytes, tmem_cols;
  __nv_bfloat16* out0;
  __nv_bfloat16* out1;
  int ld0, ld1, split, accumulate, relu;
  int wide_store;  // 256-bit epilogue stores: pointers 32-byte aligned, ld / split multiples of 16
  int tf32;        // fp32 operands / fp32 output (out0 is a float*)
  const float* scale;
  const float* shift;
  double* stats;
  // halo-resident variant (conv_halo.cu): R output rows x 128 columns per work item
  int R, segs_w, blocks_h, a_bytes, b_stage_bytes, b_stages;
};

int conv_fwd_launch(const ConvFwdArgs& a, cudaStream_t stream);
// Halo-resident 3x3 kernel; returns 1 if the shape is not eligible (caller falls back).
int conv_halo_launch(const ConvFwdArgs& a, cudaStream_t stream);
void conv_set_mode(int mode);  // 0 = automatic, 1 = never use the halo kernel (A/B testing)

// ---- wgrad implicit GEMM -----------------------------------------------------------------
struct ConvWgradArgs {
  const void* in0;  // forward input source 0 (N,H,W,C0) bf16
  const void* in1;  // forward input source 1 or null
  const void* dy;   // (N,H,W,Cout) bf16 gradient of the raw conv output
  float* partial;   // (splits, taps*(C0+C1), Cout) fp32 workspace
  int* splits_used; // out
  int N, H, W, C0, C1, Cout, taps;
  int ld_in0, ld_in1, ld_dy;
  int max_splits;   // rows available in `partial`
  int splits_override;
};

struct ConvWgradParams {
  int N, H, W, C0, C1, Cout, taps, mc;
  int BW, BH, BI, chunks_w, chunks_h, chunks_n;
  int BN, n_tiles, m_tiles, splits, stages, stage_bytes, tmem_cols;
  int a_sub_bytes, b_sub_bytes, a_bytes;
  float* partial;
};

int conv_wgrad_launch(const ConvWgradArgs& a, cudaStream_t stream);

// ---- tensor maps -------------------------------------------------------------------------
// NHWC bf16 activation viewed as a 4-D tensor {C, W, H, N}; box = {bc, bw, bh, bi}.
int make_tmap_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int ld,
                   const uint32_t box[4], int swizzle_bytes, int elem_bytes = 2);
// Row-major bf16 matrix {cols (inner), rows}; box = {bc, br}.
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld,
                 uint32_t bc, uint32_t br, int swizzle_bytes, int elem_bytes = 2);
int num_sms();
// Resident CTA pairs a cluster kernel may use: the occupancy figure, capped by the SMs this process may fill
// (UB2_RESERVE_SMS leaves SMs to co-running kernels of another stream — the NCCL all-reduce of a multi-GPU step).
inline int cap_clusters(int occupancy_clusters) { const int c = num_sms() / 2; return occupancy_clusters < c ? occupancy_clusters : c; }
// Shared-memory budget of a weight-gradient CTA (bytes).  The trainer runs weight gradients on a side stream next to
// the bandwidth-bound BatchNorm passes of the following layer; those only find room on an SM if the resident
// weight-gradient CTA leaves some shared memory (and it does leave 3/4 of the registers), and only under the common
// 164 KB carveout of launch.cuh's launch_co().  Costs the deepest pipelines one or two stages (down1.3: 68 -> 72 us
// alone, the others unchanged; 131 KB would double the 512^2 layers).  UB2_WGRAD_SMEM_KB.
inline int wgrad_smem_budget() {
  static const int kb = [] { const char* e = getenv("UB2_WGRAD_SMEM_KB"); int v = e ? atoi(e) : 163; return v < 64 ? 64 : (v > 227 ? 227 : v); }();
  return kb * 1024;
}
// Which kernel the dispatcher picked for the calling thread's last convolution launch (tests assert it):
// forward / dgrad 1 = conv_fwd (one CTA, per tap), 2 = conv_fwd2 (CTA pair, per tap), 3 = conv_halo (one CTA,
// halo resident), 4 = conv_halo2 (CTA pair, halo resident); weight gradient 11..14 likewise.
void note_variant(int code);
int last_variant();
// Launch state that is per device (function attributes and occupancy belong to a context, and two B200s
// of one box need not have the same number of complete TPCs): every `static` cache of the launchers is
// one slot per device ordinal.  Slots hold idempotent results, so a racing first use by two host
// threads (forward thread / autograd thread) writes the same value twice.
constexpr int kMaxDevices = 64;
int device_index();   // current CUDA device, clamped to [0, kMaxDevices)
template <typename T>
struct PerDevice {
  T v[kMaxDevices] = {};
  T& ref() { return v[device_index()]; }
};

inline int conv_wide_store_ok(const void* out0, int ld0, const void* out1, int ld1, int split, int Cout) {
  if ((reinterpret_cast<uintptr_t>(out0) & 31) != 0 || ld0 % 16 != 0 || Cout % 16 != 0) return 0;
  if (out1 != nullptr && ((reinterpret_cast<uintptr_t>(out1) & 31) != 0 || ld1 % 16 != 0 || split % 16 != 0)) return 0;
  return 1;
}

}  // namespace ub2
