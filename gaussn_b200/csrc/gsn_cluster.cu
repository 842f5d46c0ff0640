// Instantiations of gsn::logl_cluster_kernel (one theta per thread-block cluster).  Compiled twice with
//   -DGSN_SLICE_NAME=... -DGSN_SLICE_LO=.. -DGSN_SLICE_HI=..      (see gaussn_b200/build.py)
#include <algorithm>
#include "gsn_launch.h"
#include "gsn_cluster.cuh"

namespace gsn {

template <int KIND>
static cudaError_t prepare_one() {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute(logl_cluster_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  return cudaSuccess;
}

static void fill_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute& attr, int n_ctas, int cluster, size_t smem, cudaStream_t stream) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)n_ctas, 1, 1);
  cfg.blockDim = dim3(CfgCluster::kWarps * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)cluster;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
}

template <int KIND>
static cudaError_t launch_one(const LaunchArgs& a) {
  cudaError_t e = prepare_one<KIND>();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr;
  fill_config(cfg, attr, a.grid, a.cluster, a.smem_bytes, a.stream);
  return cudaLaunchKernelEx(&cfg, logl_cluster_kernel<KIND>, a.pd, a.theta, a.B, a.ld, a.hostmean, a.out, a.info, a.flags,
                            a.workspace, a.counter);
}

template <int KIND>
static cudaError_t max_clusters_one(size_t smem, int cluster, int* n) {
  cudaError_t e = prepare_one<KIND>();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr;
  fill_config(cfg, attr, cluster, cluster, smem, nullptr);
  return cudaOccupancyMaxActiveClusters(n, logl_cluster_kernel<KIND>, &cfg);
}

template <int KIND>
static cudaError_t dispatch(int kind, const LaunchArgs* a, size_t smem, int cluster, int* n) {
  if constexpr (KIND > GSN_SLICE_HI) {
    return cudaErrorInvalidValue;
  } else {
    if (kind == KIND) return a ? launch_one<KIND>(*a) : max_clusters_one<KIND>(smem, cluster, n);
    return dispatch<KIND + 1>(kind, a, smem, cluster, n);
  }
}

// a != nullptr: launch (a->grid CTAs in clusters of a->cluster).
// a == nullptr: prepare the kernel and report in *max_clusters how many clusters of `cluster` CTAs can be resident at once
// with smem bytes per CTA.
cudaError_t GSN_SLICE_NAME(int kind, const LaunchArgs* a, size_t smem, int cluster, int* max_clusters) {
  return dispatch<GSN_SLICE_LO>(kind, a, smem, cluster, max_clusters);
}

}  // namespace gsn
