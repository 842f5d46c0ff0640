// Instantiations of gsn::predict_dmma_kernel.  Compiled twice with
//   -DGSN_SLICE_NAME=... -DGSN_SLICE_LO=.. -DGSN_SLICE_HI=..      (see gaussn_b200/build.py)
#include <algorithm>
#include "gsn_launch.h"
#include "gsn_predict.cuh"

namespace gsn {

template <int KIND>
static cudaError_t prepare_one() {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute(predict_dmma_kernel<KIND, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(predict_dmma_kernel<KIND, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  return cudaSuccess;
}

template <int KIND>
static cudaError_t launch_one(const PredictLaunchArgs& a) {
  cudaError_t e = prepare_one<KIND>();
  if (e != cudaSuccess) return e;
  if (a.cluster > 1) {   // one theta per thread-block cluster of a.cluster CTAs
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr;
    cfg.gridDim = dim3((unsigned)a.grid, 1, 1);
    cfg.blockDim = dim3(CfgWide::kWarps * 32, 1, 1);
    cfg.dynamicSmemBytes = a.smem_bytes;
    cfg.stream = a.stream;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)a.cluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, predict_dmma_kernel<KIND, 2>, a.pd, a.pa, a.theta, a.B, a.ld, a.info, a.workspace, a.counter);
  }
  predict_dmma_kernel<KIND, 1><<<a.grid, CfgWide::kWarps * 32, a.smem_bytes, a.stream>>>(a.pd, a.pa, a.theta, a.B, a.ld, a.info,
                                                                                        a.workspace, a.counter);
  return cudaGetLastError();
}

template <int KIND>
static cudaError_t dispatch(int kind, const PredictLaunchArgs* a) {
  if constexpr (KIND > GSN_SLICE_HI) {
    return cudaErrorInvalidValue;
  } else {
    if (kind == KIND) return a ? launch_one<KIND>(*a) : prepare_one<KIND>();
    return dispatch<KIND + 1>(kind, a);
  }
}

cudaError_t GSN_SLICE_NAME(int kind, const PredictLaunchArgs* a) { return dispatch<GSN_SLICE_LO>(kind, a); }

}  // namespace gsn
