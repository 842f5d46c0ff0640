// Instantiations of gsn::logl_dmma_kernel.  Compiled four times with
//   -DGSN_SLICE_CFG={CfgWide|CfgNarrow} -DGSN_SLICE_NAME=... -DGSN_SLICE_LO=.. -DGSN_SLICE_HI=..
// (see gaussn_b200/build.py) so that the 22 kernel variants build in parallel.
#include "gsn_launch.h"

namespace gsn {

// Raises the kernel's dynamic shared-memory limit once per device; this also forces the (lazily loaded)
// kernel image onto the device, which gsn_problem_create uses to keep that cost out of the first batch call.
template <int KIND>
static cudaError_t prepare_one() {
  using Cfg = GSN_SLICE_CFG;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute(logl_dmma_kernel<KIND, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  return cudaSuccess;
}

template <int KIND>
static cudaError_t launch_one(const LaunchArgs& a) {
  using Cfg = GSN_SLICE_CFG;
  auto kern = logl_dmma_kernel<KIND, Cfg>;
  cudaError_t e = prepare_one<KIND>();
  if (e != cudaSuccess) return e;
  kern<<<a.grid, Cfg::kWarps * 32, a.smem_bytes, a.stream>>>(a.pd, a.theta, a.B, a.ld, a.hostmean, a.out, a.info, a.flags,
                                                            a.workspace, a.counter);
  return cudaGetLastError();
}

template <int KIND>
static cudaError_t dispatch(int kind, const LaunchArgs* a) {   // a == nullptr: prepare only
  if constexpr (KIND > GSN_SLICE_HI) {
    return cudaErrorInvalidValue;
  } else {
    if (kind == KIND) return a ? launch_one<KIND>(*a) : prepare_one<KIND>();
    return dispatch<KIND + 1>(kind, a);
  }
}

cudaError_t GSN_SLICE_NAME(int kind, const LaunchArgs* a) { return dispatch<GSN_SLICE_LO>(kind, a); }

#if defined(GSN_PHASE_TIMERS) && defined(GSN_PHASE_EXPORT)
extern "C" int gsn_debug_phase_cycles(unsigned long long* out, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
  }
  return (int)e;
}
#endif

}  // namespace gsn
