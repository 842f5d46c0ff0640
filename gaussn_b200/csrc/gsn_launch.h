// Internal launch interface between the C-ABI translation unit (gsn_api.cu) and the translation
// units that instantiate the log-likelihood kernel (gsn_kernels.cu, compiled once per slice of
// {launch shape} x {covariance kinds} so that the slices build in parallel).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include "gsn_chol_dmma.cuh"
#include "gsn_predict.cuh"
#include "gsn_onchip.cuh"

namespace gsn {

struct LaunchArgs {
  ProblemDev pd;
  const double* theta;
  long long B, ld;
  const double* hostmean;
  GatherOut out;
  int* info;
  unsigned flags;
  double* workspace;
  unsigned* counter;
  int grid;
  size_t smem_bytes;
  cudaStream_t stream;
  int cluster = 1;   // CTAs per theta (launch_logl_cluster_* only)
};

// Each returns cudaSuccess, a CUDA error, or cudaErrorInvalidValue if `kind` is not in its slice.
// a == nullptr only prepares the kernel (shared-memory attribute, image load) on the current device.
cudaError_t launch_logl_wide_a(int kind, const LaunchArgs* a);     // kinds 0..4
cudaError_t launch_logl_wide_b(int kind, const LaunchArgs* a);     // kinds 5..10
cudaError_t launch_logl_narrow_a(int kind, const LaunchArgs* a);
cudaError_t launch_logl_narrow_b(int kind, const LaunchArgs* a);
cudaError_t launch_logl_tiny_a(int kind, const LaunchArgs* a);
cudaError_t launch_logl_tiny_b(int kind, const LaunchArgs* a);

// One theta per thread-block cluster (gsn_cluster.cuh).  a != nullptr launches (a->grid CTAs in clusters of a->cluster);
// a == nullptr prepares the kernel and reports how many clusters of `cluster` CTAs can be resident at once.
cudaError_t launch_logl_cluster_a(int kind, const LaunchArgs* a, size_t smem, int cluster, int* max_clusters);   // kinds 0..4
cudaError_t launch_logl_cluster_b(int kind, const LaunchArgs* a, size_t smem, int cluster, int* max_clusters);   // kinds 5..10

// Factor resident in shared memory (gsn_onchip.cuh): band blocks of at most kOnchipMaxTiles tile rows.
struct OnchipLaunchArgs {
  ProblemDev pd;
  OnchipDev od;
  const double* theta;
  long long B, ld;
  const double* hostmean;
  GatherOut out;
  int* info;
  unsigned flags;
  unsigned* counter;
  int grid;
  int warps;          // warps per theta (= CTA size / 32): 1, 2, 4, 8 or 16
  size_t smem_bytes;
  cudaStream_t stream;
  bool reg = false;   // one warp per theta with the factor in registers (band blocks of at most kRegMaxTiles tile rows)
};
// a == nullptr prepares the kernels and reports their register counts: regs[0] the shared-memory shape, regs[1] the register shape.
cudaError_t launch_logl_onchip_a(int kind, const OnchipLaunchArgs* a, int* regs);   // kinds 0..4
cudaError_t launch_logl_onchip_b(int kind, const OnchipLaunchArgs* a, int* regs);   // kinds 5..10

struct PredictLaunchArgs {
  ProblemDev pd;
  PredictDev pa;
  const double* theta;
  long long B, ld;
  int* info;
  double* workspace;
  unsigned* counter;
  int grid;
  size_t smem_bytes;
  cudaStream_t stream;
  int cluster = 1;   // > 1: one theta per thread-block cluster of this many CTAs (grid counts CTAs)
};
cudaError_t launch_predict_a(int kind, const PredictLaunchArgs* a);   // kinds 0..4
cudaError_t launch_predict_b(int kind, const PredictLaunchArgs* a);   // kinds 5..10

}  // namespace gsn
