// Latency of one 8x8 pivot-tile elimination (gsn::chol8_inv_warp) by a single warp on an otherwise idle SM, and with
// 1..3 other warps of the same SM sub-partition running back-to-back DMMAs (the contention the factorisation sees).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I. -o bench/chol8_lat bench/chol8_lat.cu
#include <cstdio>
#include "../gaussn_b200/csrc/gsn_chol_dmma.cuh"

__global__ void k(double* out, long long* cyc, int busy_warps, int busy_part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    // SPD tile: 2 I + 0.1 ones
    double2 a = make_double2((2 * q == r ? 2.0 : 0.0) + 0.1, (2 * q + 1 == r ? 2.0 : 0.0) + 0.1);
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < 64; ++it) {
      double2 nW; double mant = 1.0; int expo = 0, failrow = -1;
      gsn::chol8_inv_warp(a, lane, nW, mant, expo, failrow);
      acc += nW.x + mant;
      a.x += 1e-9 * nW.y;      // the next elimination depends on this one
    }
    const long long t1 = clock64();
    if (lane == 0) { *cyc = (t1 - t0) / 64; stop = 1; }
    out[lane] = acc;
  } else if ((warp & 3) == busy_part && (warp >> 2) >= 1 && (warp >> 2) <= busy_warps) {
    // warps 4, 8, 12 share sub-partition 0 with warp 0; warps 5, 9, 13 sit on sub-partition 1
    double2 d[4] = {{1, 1}, {1, 1}, {1, 1}, {1, 1}};
    while (!stop) {
      for (int i = 0; i < 16; ++i) { gsn::dmma(d[0], 1.0, 1e-9); gsn::dmma(d[1], 1.0, 1e-9); gsn::dmma(d[2], 1.0, 1e-9); gsn::dmma(d[3], 1.0, 1e-9); }
    }
    out[32 + threadIdx.x] = d[0].x + d[1].x + d[2].x + d[3].x;
  }
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  printf("{\"chol8_inv_warp_cycles\": {");
  for (int part = 0; part <= 1; ++part)
    for (int busy = part ? 3 : 0; busy <= 3; ++busy) {
      k<<<1, 512>>>(out, cyc, busy, part);
      cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%s\"%d_dmma_warps_on_%s_subpartition\": %lld", (busy || part) ? ", " : "", busy, part ? "another" : "the_same", h);
    }
  printf("}}\n");
  return 0;
}
