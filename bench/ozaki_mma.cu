// Rate side of the Ozaki spike: how fast are exact int8 x int8 -> int32 slice products on this GPU through the
// warp-level path (mma.sync.m16n8k32.s8, which compiles for sm_100a), next to the FP64 DMMA rate they would replace?
// (tcgen05.mma kind::i8 is the full-rate path on B200: nominal 4.5 dense POP/s.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/ozaki_mma bench/ozaki_mma.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) imma(int* out, int iters) {
  int acc[8][4];
  for (int i = 0; i < 8; ++i) for (int k = 0; k < 4; ++k) acc[i][k] = threadIdx.x + i + k;
  const unsigned a0 = 0x01020304u + threadIdx.x, a1 = 0x05060708u, a2 = 0x090a0b0cu, a3 = 0x0d0e0f10u, b0 = 0x01010101u + blockIdx.x, b1 = 0x02020202u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[i][0]), "+r"(acc[i][1]), "+r"(acc[i][2]), "+r"(acc[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
  for (int i = 0; i < 8; ++i) for (int k = 0; k < 4; ++k) s += acc[i][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_k(double* out, int iters) {
  double2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_double2(threadIdx.x * 1e-3 + i, 1.0);
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * blockIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[i].x), "+d"(acc[i].y) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 20000;
  void* out; cudaMalloc(&out, 8 * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms_i = 0, ms_d = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); imma<<<blocks, threads>>>((int*)out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_i, e0, e1);
    cudaEventRecord(e0); dmma_k<<<blocks, threads>>>((double*)out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_d, e0, e1);
  }
  const double warps = (double)blocks * threads / 32;
  const double tops = 2.0 * 16 * 8 * 32 * 8.0 * iters * warps / (ms_i * 1e-3) / 1e12;
  const double tf = 2.0 * 8 * 8 * 4 * 8.0 * iters * warps / (ms_d * 1e-3) / 1e12;
  printf("{\"int8_mma_sync_m16n8k32_tops\": %.1f, \"fp64_dmma_tflops\": %.2f, \"effective_fp64_tflops_at_45_slice_products\": %.2f, \"at_55\": %.2f, "
         "\"nominal_tcgen05_i8_tops\": 4500, \"effective_at_45_on_tcgen05_nominal\": %.1f}\n", tops, tf, tops / 45, tops / 55, 4500.0 / 45);
  return 0;
}
