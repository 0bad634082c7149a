/* libunetb200 — C ABI of the B200 (sm_100a) Attention U-Net hot path.
 *
 * Boundary contract:
 *  - plain pointers and sizes only, no torch / C++ types; every entry point returns
 *    0 on success, a positive cudaError_t, or a negative UB2_ERR_* code; no exception
 *    crosses the boundary;
 *  - the library never allocates, frees or retains device memory: activations, grads,
 *    statistics and workspaces are owned by the caller (torch's allocator);
 *  - every call is asynchronous on the passed cudaStream_t (as void*), never
 *    synchronises the device, and is re-entrant (autograd calls backward entry points
 *    from its own thread);
 *  - activations are NHWC bf16 ("(N,H,W,C), channel stride ld"), parameters stay in the
 *    reference's fp32 OIHW layout and are packed to bf16 by ub2_pack_conv_weight.
 *
 * Each entry point cites the reference code (paths relative to the reference repo)
 * whose work it replaces.
 */
#ifndef UNETB200_H
#define UNETB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB2_OK 0
#define UB2_ERR_SHAPE (-1)
#define UB2_ERR_ALIGN (-2)
#define UB2_ERR_WORKSPACE (-3)
#define UB2_ERR_DRIVER (-4)
#define UB2_ERR_ARCH (-5)

int ub2_version(void);
int ub2_num_sms(void);

/* ---- convolution on tcgen05 tensor cores ------------------------------------------------
 * ub2_conv_fwd: nn.Conv2d(k=3,p=1,bias=False) of DoubleConv (unet/models/layers.py:32,35)
 * and the 1x1 projections of AttentionGate (layers.py:152,158) as an NHWC bf16 implicit
 * GEMM.  The input channel axis may span two tensors (in0: C0 channels, in1: C1 channels):
 * this is torch.cat([x2, x1], dim=1) of Up/AttentionUp (layers.py:105, :254) without the
 * copy.  wgt is (Cout, taps, C0+C1) bf16.  Optional epilogue: per-channel scale/shift
 * (BatchNorm folded, layers.py:33 in eval mode), ReLU (layers.py:34), accumulate into the
 * destination, split of the output channels over two tensors (the dgrad of a concat), and
 * per-CTA per-channel sum / sum-of-squares rows (stats: (stats_rows,2,Cout) doubles) for
 * train-mode BatchNorm statistics.  Called with the flipped/transposed weight pack it is
 * the data-gradient pass autograd runs for the same nn.Conv2d.
 */
int ub2_conv_fwd(const void* in0, int ld_in0, int C0, const void* in1, int ld_in1, int C1,
                 const void* wgt, void* out0, int ld0, void* out1, int ld1, int split, int N, int H,
                 int W, int Cout, int taps, const float* scale, const float* shift, int relu,
                 int accumulate, double* stats, int stats_rows, int* stats_rows_used,
                 int bn_override, int grid_override, void* stream);

/* ub2_conv_wgrad: the weight-gradient pass of the same nn.Conv2d (autograd of
 * layers.py:32,35,152,158).  Writes split-K partial tiles (splits, taps*(C0+C1), Cout) fp32;
 * ub2_wgrad_reduce folds them, in a fixed order, into the OIHW fp32 .grad tensor.
 */
int ub2_conv_wgrad(const void* in0, int ld_in0, int C0, const void* in1, int ld_in1, int C1,
                   const void* dy, int ld_dy, float* partial, int max_splits, int* splits_used,
                   int N, int H, int W, int Cout, int taps, int splits_override, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETB200_H */
