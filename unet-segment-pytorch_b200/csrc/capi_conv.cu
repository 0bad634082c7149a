// C-ABI entry points for the tcgen05 convolution family (see include/unetb200.h).
#include "../../include/unetb200.h"
#include "conv.h"

using namespace ub2;

extern "C" {

int ub2_conv_fwd(const void* in0, int ld_in0, int C0, const void* in1, int ld_in1, int C1,
                 const void* wgt, void* out0, int ld0, void* out1, int ld1, int split, int N, int H,
                 int W, int Cout, int taps, const float* scale, const float* shift, int relu,
                 int accumulate, double* stats, int stats_rows, int* stats_rows_used,
                 int bn_override, int grid_override, void* stream) {
  ConvFwdArgs a{};
  a.in0 = in0; a.in1 = in1; a.wgt = wgt; a.out0 = out0; a.out1 = out1;
  a.scale = scale; a.shift = shift; a.stats = stats; a.grid_used = stats_rows_used;
  a.N = N; a.H = H; a.W = W; a.C0 = C0; a.C1 = C1; a.Cout = Cout; a.taps = taps;
  a.ld_in0 = ld_in0; a.ld_in1 = ld_in1; a.ld0 = ld0; a.ld1 = ld1; a.split = split;
  a.accumulate = accumulate; a.relu = relu; a.stats_rows = stats_rows;
  a.bn_override = bn_override; a.grid_override = grid_override;
  return conv_fwd_launch(a, static_cast<cudaStream_t>(stream));
}

int ub2_conv_fwd_tf32(const float* in0, int ld_in0, int C0, const float* in1, int ld_in1, int C1,
                      const float* wgt, float* out, int ld_out, int N, int H, int W, int Cout, int taps,
                      const float* scale, const float* shift, int relu, int exact_out, void* stream) {
  ConvFwdArgs a{};
  a.in0 = in0; a.in1 = in1; a.wgt = wgt; a.out0 = out;
  a.scale = scale; a.shift = shift; a.relu = relu;
  a.N = N; a.H = H; a.W = W; a.C0 = C0; a.C1 = C1; a.Cout = Cout; a.taps = taps;
  a.ld_in0 = ld_in0; a.ld_in1 = ld_in1; a.ld0 = ld_out;
  a.tf32 = exact_out ? 2 : 1;
  return conv_fwd_launch(a, static_cast<cudaStream_t>(stream));
}

int ub2_conv_wgrad(const void* in0, int ld_in0, int C0, const void* in1, int ld_in1, int C1,
                   const void* dy, int ld_dy, float* partial, int max_splits, int* splits_used,
                   int N, int H, int W, int Cout, int taps, int splits_override, void* stream) {
  ConvWgradArgs a{};
  a.in0 = in0; a.in1 = in1; a.dy = dy; a.partial = partial; a.splits_used = splits_used;
  a.N = N; a.H = H; a.W = W; a.C0 = C0; a.C1 = C1; a.Cout = Cout; a.taps = taps;
  a.ld_in0 = ld_in0; a.ld_in1 = ld_in1; a.ld_dy = ld_dy;
  a.max_splits = max_splits; a.splits_override = splits_override;
  return conv_wgrad_launch(a, static_cast<cudaStream_t>(stream));
}

int ub2_num_sms(void) { return num_sms(); }

int ub2_last_conv_variant(void) { return last_variant(); }

int ub2_set_conv_mode(int mode) {
  conv_set_mode(mode);
  return 0;
}

}  // extern "C"
