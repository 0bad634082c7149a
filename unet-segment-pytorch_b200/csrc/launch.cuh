// Kernel launch with programmatic dependent launch (PDL), and its device side.
//
// A training step is ~250 kernels of 28 us on average, replayed from a CUDA graph: between two kernels the
// GPU drains the first grid, launches the second and runs its prologue (barrier init, TMEM allocation,
// descriptor prefetch) — a few microseconds each, several percent of the step.  Every kernel of this
// library is therefore launched with cudaLaunchAttributeProgrammaticStreamSerialization and
//   * calls pdl_trigger() first thing: once all its CTAs have started, the NEXT kernel in the stream may be
//     launched and its CTAs become resident wherever an SM has room (the tail of this grid);
//   * calls pdl_wait() before its first access to global memory: it blocks until every prerequisite grid has
//     completed and flushed, so ordering and visibility are exactly those of a plain stream launch
//     (transitively: a kernel cannot finish before its predecessor has).
// Stream capture keeps the edge as a programmatic dependency in the graph.
//
// MEASURED (B200, same box, interleaved, profiles/r02_pdl_ab.md): no gain — batch-4 step 7.045 ms plain vs
// 7.085 ms with PDL, batch 32 49.2 vs 50.0 ms.  Inside a replayed graph the kernels are already back to
// back (the step time equals the sum of the kernel times), and early-resident dependents only take
// scheduling slots from the tail they wait for.  So the attribute is OFF by default (UB2_PDL=1 turns it
// on); without it griddepcontrol.wait / launch_dependents are no-ops.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

namespace ub2 {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("UB2_PDL");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

// Shared-memory carveout of the kernels that are meant to run side by side on one SM.  An SM runs kernels of two
// streams concurrently only if they agree on its L1 / shared-memory split; by default the driver picks the smallest
// carveout that fits each kernel, so a streaming kernel (no shared memory, all L1) never becomes resident next to a
// tensor-core kernel (227 KB carveout) even when registers and threads are free.  The trainer runs the weight-gradient
// GEMMs on a side stream next to the BatchNorm-backward passes (ops.SideStream): both sides are launched with
// launch_co(), which asks for the same split, 164 KB shared + 92 KB L1 (72 %).  Measured (tools/probe/overlap.py,
// B200, batch 4): with 196 KB or more of shared memory the BatchNorm passes lose 12 % (a 60 KB L1 cannot hold the
// loads a streaming kernel needs in flight), with 164 KB they run at full speed and the pair takes 81 us instead of
// 100 (256-channel level), 69 instead of 89 (512).  UB2_CO_CARVEOUT=<percent> overrides, -1 leaves the driver's choice.
inline int co_carveout_percent() {
  static const int pct = [] {
    const char* e = getenv("UB2_CO_CARVEOUT");
    const int v = e ? atoi(e) : 72;
    return v > 100 ? 100 : v;
  }();
  return pct;
}

inline void apply_carveout(const void* kernel, int pct) {
  if (pct < 0) return;
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  (void)cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (done.insert({dev, kernel}).second)
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors: cudaGetLastError() at the call site
}

// launch() for a kernel that shares SMs with a kernel of another stream (see co_carveout_percent()).
template <typename... KArgs, typename... Args>
inline void launch_co(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  apply_carveout(reinterpret_cast<const void*>(kernel), co_carveout_percent());
  launch(kernel, grid, block, smem, stream, std::forward<Args>(args)...);
}

}  // namespace ub2
