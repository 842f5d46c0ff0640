// Instantiations of gsn::logl_onchip_kernel (factor resident in shared memory, small band blocks).  Compiled twice with
//   -DGSN_SLICE_NAME=... -DGSN_SLICE_LO=.. -DGSN_SLICE_HI=..      (see gaussn_b200/build.py)
#include "gsn_launch.h"
#include "gsn_onchip.cuh"

namespace gsn {

template <int KIND>
static cudaError_t prepare_one(int* regs) {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute(logl_onchip_kernel<KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (regs) {
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, logl_onchip_kernel<KIND, false>);
    if (e != cudaSuccess) return e;
    regs[0] = fa.numRegs;
    e = cudaFuncGetAttributes(&fa, logl_onchip_kernel<KIND, true>);
    if (e != cudaSuccess) return e;
    regs[1] = fa.numRegs;
  }
  return cudaSuccess;
}

template <int KIND>
static cudaError_t launch_one(const OnchipLaunchArgs& a) {
  cudaError_t e = prepare_one<KIND>(nullptr);
  if (e != cudaSuccess) return e;
  if (a.reg)
    logl_onchip_kernel<KIND, true><<<a.grid, 32, a.smem_bytes, a.stream>>>(a.pd, a.od, a.theta, a.B, a.ld, a.hostmean, a.out, a.info,
                                                                           a.flags, a.counter);
  else
    logl_onchip_kernel<KIND, false><<<a.grid, a.warps * 32, a.smem_bytes, a.stream>>>(a.pd, a.od, a.theta, a.B, a.ld, a.hostmean, a.out,
                                                                                      a.info, a.flags, a.counter);
  return cudaGetLastError();
}

template <int KIND>
static cudaError_t dispatch(int kind, const OnchipLaunchArgs* a, int* regs) {   // a == nullptr: prepare only
  if constexpr (KIND > GSN_SLICE_HI) {
    return cudaErrorInvalidValue;
  } else {
    if (kind == KIND) return a ? launch_one<KIND>(*a) : prepare_one<KIND>(regs);
    return dispatch<KIND + 1>(kind, a, regs);
  }
}

#if defined(GSN_ONCHIP_TRACE) && GSN_SLICE_LO == 0
extern "C" int gsn_debug_onchip_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_onchip_trace, sizeof(long long) * (size_t)n);
}
#endif

cudaError_t GSN_SLICE_NAME(int kind, const OnchipLaunchArgs* a, int* regs) { return dispatch<GSN_SLICE_LO>(kind, a, regs); }

}  // namespace gsn
