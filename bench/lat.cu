// Dependent-chain latencies and single-warp issue rates of the FP64 instructions the small-N kernel is made of, in SM
// cycles (clock64 around a chain of 256 dependent instructions, one warp per SM sub-partition unless stated).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/lat bench/lat.cu ; prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double2& d, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
}

constexpr int kIters = 256;

template <int MODE>
__global__ void chain(double* out, long long* cycles, double seed) {
  __shared__ __align__(16) double2 buf[64];
  double x = seed + threadIdx.x * 1e-3, y = 1.0000001, z = 0.999;
  double2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_double2(x + i, x - i);
  buf[threadIdx.x & 63] = make_double2(x, y);
  __syncthreads();
  unsigned idx = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; ++i) {
    if (MODE == 0) x = fma(x, y, z);
    if (MODE == 1) x = x * y;
    if (MODE == 2) x = x + y;
    if (MODE == 3) dmma(acc[0], y, z);
    if (MODE == 4) { dmma(acc[0], y, z); dmma(acc[1], y, z); }
    if (MODE == 5) { dmma(acc[0], y, z); dmma(acc[1], y, z); dmma(acc[2], y, z); dmma(acc[3], y, z); }
    if (MODE == 6) { for (int k = 0; k < 8; ++k) dmma(acc[k], y, z); }
    if (MODE == 7) x = rsqrt(x) + 1.5;
    if (MODE == 8) x = 1.0 / x + 0.5;
    if (MODE == 9) x = sqrt(x) + 1.5;
    if (MODE == 10) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (MODE == 11) { idx = (unsigned)__double2loint(buf[idx].x) & 31; }
    if (MODE == 12) x = __drcp_rn(x) + 0.5;
    if (MODE == 13) { double a = x, b = x + 1, c = x + 2, d = x + 3; a = fma(a, y, z); b = fma(b, y, z); c = fma(c, y, z); d = fma(d, y, z); x = a + b + c + d; }
  }
  const long long t1 = clock64();
  if (MODE == 11) { x = idx; }
  double s = x;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
double run(int threads, double* out, long long* cyc) {
  if (MODE == 11) {
    // pointer chase: buf[i].x = i + 1
  }
  chain<MODE><<<1, threads>>>(out, cyc, MODE == 11 ? 1.0 : 1.25);
  cudaDeviceSynchronize();
  chain<MODE><<<1, threads>>>(out, cyc, MODE == 11 ? 1.0 : 1.25);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  return (double)h / kIters;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&cyc, 8);
  printf("{\"unit\": \"SM cycles per loop iteration, one warp\"");
  printf(", \"dfma_dep\": %.1f", run<0>(32, out, cyc));
  printf(", \"dmul_dep\": %.1f", run<1>(32, out, cyc));
  printf(", \"dadd_dep\": %.1f", run<2>(32, out, cyc));
  printf(", \"dmma_dep_1acc\": %.1f", run<3>(32, out, cyc));
  printf(", \"dmma_2acc_per_iter\": %.1f", run<4>(32, out, cyc));
  printf(", \"dmma_4acc_per_iter\": %.1f", run<5>(32, out, cyc));
  printf(", \"dmma_8acc_per_iter\": %.1f", run<6>(32, out, cyc));
  printf(", \"dmma_8acc_per_iter_4warps\": %.1f", run<6>(128, out, cyc));
  printf(", \"dmma_8acc_per_iter_8warps\": %.1f", run<6>(256, out, cyc));
  printf(", \"rsqrt_plus_add_dep\": %.1f", run<7>(32, out, cyc));
  printf(", \"div_plus_add_dep\": %.1f", run<8>(32, out, cyc));
  printf(", \"sqrt_plus_add_dep\": %.1f", run<9>(32, out, cyc));
  printf(", \"drcp_plus_add_dep\": %.1f", run<12>(32, out, cyc));
  printf(", \"shfl_dep\": %.1f", run<10>(32, out, cyc));
  printf(", \"lds128_chase\": %.1f", run<11>(32, out, cyc));
  printf(", \"dfma_4indep_plus_3add\": %.1f", run<13>(32, out, cyc));
  printf(", \"dfma_dep_8warps\": %.1f", run<0>(256, out, cyc));
  printf("}\n");
  return 0;
}
