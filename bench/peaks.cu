// FP64 peak microbenchmarks for B200 (sm_100a).
//
// MEASURED_PEAKS.json holds HBM and bf16 peaks only; the GP log-likelihood path
// is FP64-compute bound, so its roofline denominator has to be measured here
// (SURVEY.md F11, section 8d "FP64 peak").  Three candidates are timed:
//   * a DFMA chain kernel (many independent accumulators per thread),
//   * a DMMA kernel (mma.sync m8n8k4 f64, SASS DMMA.8x8x4; the larger PTX f64
//     shapes all lower to this instruction on sm_100a),
//   * cuBLAS DGEMM 8192^3.
// plus the latencies/throughputs the Cholesky kernels are designed around:
// DMMA dependent-issue latency, DFMA+DMMA co-issue, exp/sqrt/div throughput,
// and shared-memory LDS.64 bandwidth.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo \
//              -o bench/peaks bench/peaks.cu -lcublas
// Run:    bench/peaks > gpurun_out/peaks.json
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- DFMA: ILP independent chains per thread ----
template <int ILP>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- DMMA: ILP independent accumulator tiles per warp ----
template <int ILP>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  double A = a + threadIdx.x * 1e-9, B = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], A, B);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- DMMA + DFMA interleaved (shared pipe?) ----
template <int NM, int NF>
__global__ void __launch_bounds__(1024) k_mix(double* out, int iters, double a, double b) {
  double c0[NM + 1], c1[NM + 1], f[NF + 1];
#pragma unroll
  for (int i = 0; i < NM; ++i) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
  double A = a + threadIdx.x * 1e-9, B = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NM; ++i) dmma884(c0[i], c1[i], A, B);
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- single-warp dependent latency (cycles per op) ----
__global__ void k_lat(long long* cyc, double* out, int iters, double a, double b) {
  double c0 = 1, c1 = 2, f = 3, q = 1.5, r = 2.0, e = 0.5;
  double A = a, B = b;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) dmma884(c0, c1, A, B);
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) f = fma(f, a, b);
  long long t2 = clock64();
  for (int it = 0; it < iters; ++it) q = sqrt(q + a);
  long long t3 = clock64();
  for (int it = 0; it < iters; ++it) r = b / (r + a);
  long long t4 = clock64();
  for (int it = 0; it < iters; ++it) e = exp(-e * a);
  long long t5 = clock64();
  for (int it = 0; it < iters; ++it) e = rsqrt(e + a);
  long long t6 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5;
  }
  out[threadIdx.x] = c0 + c1 + f + q + r + e;
}

// ---- throughput of exp / sqrt / div / rsqrt (elements per second) ----
template <int OP>
__global__ void __launch_bounds__(256) k_func(double* out, int iters, double a) {
  constexpr int ILP = 4;
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = 0.25 + 1e-3 * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) v[i] = exp(-v[i] * a);
      if (OP == 1) v[i] = sqrt(v[i] + a);
      if (OP == 2) v[i] = a / (v[i] + a);
      if (OP == 3) v[i] = rsqrt(v[i] + a);
      if (OP == 4) v[i] = sin(v[i] + a);
      if (OP == 5) v[i] = log(v[i] + a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- shared memory LDS.64 / LDS.128 bandwidth ----
template <int VEC>
__global__ void __launch_bounds__(1024) k_lds(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  int base = (threadIdx.x * VEC) & 8191;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int idx = (base + u * 1024 * VEC / 8 * 8) & 8191;
      if (VEC == 1) { s += sm[idx]; }
      else { double2 v = *reinterpret_cast<const double2*>(&sm[idx & ~1]); s += v.x + v.y; }
    }
    base = (base + 64) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  template <class F> double best_ms(F f, int reps = 5) {
    f(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
      CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, (double)ms);
    }
    return best;
  }
};

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 64 * sms));
  Timer T;
  std::string js = "{";
  char buf[512];
  snprintf(buf, sizeof buf, "\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", prop.name, sms, prop.clockRate);
  js += buf;

  // DFMA throughput: blocks/SM x threads sweep, ILP 8
  {
    double best = 0; int bt = 0, bb = 0;
    for (int threads : {256, 512, 1024}) for (int bps : {1, 2, 4}) {
      if (threads * bps > 2048) continue;
      int iters = 4096;
      double ms = T.best_ms([&] { k_dfma<8><<<sms * bps, threads>>>(out, iters, 1.0000001, 1e-9); });
      double tf = 2.0 * 8 * iters * (double)threads * bps * sms / (ms * 1e-3) / 1e12;
      if (tf > best) { best = tf; bt = threads; bb = bps; }
    }
    snprintf(buf, sizeof buf, ", \"dfma_tflops\": %.3f, \"dfma_cfg\": \"%dthr x %dblk/SM ilp8\"", best, bt, bb); js += buf;
  }
  // DMMA throughput sweep: warps/SM and ILP
  {
    double best = 0; std::string cfg;
    std::string table = ", \"dmma_sweep\": [";
    bool first = true;
    auto run = [&](int ilp, int threads, int bps) {
      int iters = 2048;
      double ms = 0;
      auto launch = [&] {
        switch (ilp) {
          case 1: k_dmma<1><<<sms * bps, threads>>>(out, iters, 1.0, 1e-9); break;
          case 2: k_dmma<2><<<sms * bps, threads>>>(out, iters, 1.0, 1e-9); break;
          case 4: k_dmma<4><<<sms * bps, threads>>>(out, iters, 1.0, 1e-9); break;
          case 8: k_dmma<8><<<sms * bps, threads>>>(out, iters, 1.0, 1e-9); break;
          case 16: k_dmma<16><<<sms * bps, threads>>>(out, iters, 1.0, 1e-9); break;
        }
      };
      ms = T.best_ms(launch);
      double warps = threads / 32.0 * bps * sms;
      double tf = 2.0 * 256 * ilp * iters * warps / (ms * 1e-3) / 1e12;
      snprintf(buf, sizeof buf, "%s{\"ilp\": %d, \"warps_per_sm\": %d, \"tflops\": %.3f}", first ? "" : ", ", ilp, threads / 32 * bps, tf);
      table += buf; first = false;
      if (tf > best) { best = tf; snprintf(buf, sizeof buf, "ilp%d %dwarps/SM", ilp, threads / 32 * bps); cfg = buf; }
    };
    for (int ilp : {1, 2, 4, 8, 16}) for (int w : {4, 8, 16, 32}) run(ilp, std::min(w * 32, 1024), 1);
    table += "]";
    snprintf(buf, sizeof buf, ", \"dmma_tflops\": %.3f, \"dmma_cfg\": \"%s\"", best, cfg.c_str()); js += buf;
    js += table;
  }
  // mixed DMMA + DFMA: does DFMA work ride for free next to DMMA?
  {
    int iters = 2048, threads = 512;
    double ms_m = T.best_ms([&] { k_mix<8, 0><<<sms, threads>>>(out, iters, 1.0, 1e-9); });
    double ms_f = T.best_ms([&] { k_mix<0, 16><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    double ms_x = T.best_ms([&] { k_mix<8, 16><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    snprintf(buf, sizeof buf, ", \"mix_ms\": {\"dmma8\": %.4f, \"dfma16\": %.4f, \"both\": %.4f}", ms_m, ms_f, ms_x); js += buf;
  }
  // latencies
  {
    long long* cyc; CK(cudaMalloc(&cyc, 8 * sizeof(long long)));
    int iters = 4096;
    k_lat<<<1, 32>>>(cyc, out, iters, 1e-9, 1.0); CK(cudaDeviceSynchronize());
    k_lat<<<1, 32>>>(cyc, out, iters, 1e-9, 1.0); CK(cudaDeviceSynchronize());
    long long h[6]; CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
    snprintf(buf, sizeof buf, ", \"latency_cycles\": {\"dmma\": %.1f, \"dfma\": %.1f, \"sqrt\": %.1f, \"div\": %.1f, \"exp\": %.1f, \"rsqrt\": %.1f}",
             h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters, h[5] / (double)iters);
    js += buf;
  }
  // function throughput (G elements / s, whole chip)
  {
    const char* names[6] = {"exp", "sqrt", "div", "rsqrt", "sin", "log"};
    js += ", \"func_gelem_per_s\": {";
    for (int op = 0; op < 6; ++op) {
      int iters = 512, threads = 256, bps = 8;
      auto launch = [&] {
        switch (op) {
          case 0: k_func<0><<<sms * bps, threads>>>(out, iters, 1.001); break;
          case 1: k_func<1><<<sms * bps, threads>>>(out, iters, 1.001); break;
          case 2: k_func<2><<<sms * bps, threads>>>(out, iters, 1.001); break;
          case 3: k_func<3><<<sms * bps, threads>>>(out, iters, 1.001); break;
          case 4: k_func<4><<<sms * bps, threads>>>(out, iters, 1.001); break;
          case 5: k_func<5><<<sms * bps, threads>>>(out, iters, 1.001); break;
        }
      };
      double ms = T.best_ms(launch);
      double ge = 4.0 * iters * threads * bps * sms / (ms * 1e-3) / 1e9;
      snprintf(buf, sizeof buf, "%s\"%s\": %.2f", op ? ", " : "", names[op], ge); js += buf;
    }
    js += "}";
  }
  // LDS bandwidth
  {
    int iters = 2048, threads = 1024;
    CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    double ms1 = T.best_ms([&] { k_lds<1><<<sms, threads, 65536>>>(out, iters); });
    double ms2 = T.best_ms([&] { k_lds<2><<<sms, threads, 65536>>>(out, iters); });
    double b1 = 8.0 * 8 * iters * threads * sms / (ms1 * 1e-3) / 1e12;
    double b2 = 16.0 * 8 * iters * threads * sms / (ms2 * 1e-3) / 1e12;
    snprintf(buf, sizeof buf, ", \"lds_tbs\": {\"lds64\": %.2f, \"lds128\": %.2f}", b1, b2); js += buf;
  }
  // cuBLAS DGEMM
  {
    cublasHandle_t h; cublasCreate(&h);
    js += ", \"dgemm_tflops\": {";
    bool first = true;
    for (int n : {2048, 4096, 8192}) {
      double *A, *B, *C; size_t bytes = sizeof(double) * (size_t)n * n;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
      CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
      double one = 1.0, zero = 0.0;
      double ms = T.best_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
      double tf = 2.0 * n * (double)n * n / (ms * 1e-3) / 1e12;
      snprintf(buf, sizeof buf, "%s\"%d\": %.3f", first ? "" : ", ", n, tf); js += buf; first = false;
      // sustained: back to back for ~2 s at n=8192
      if (n == 8192) {
        int reps = std::max(1, (int)(2000.0 / ms));
        double msr = T.best_ms([&] { for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); }, 1);
        snprintf(buf, sizeof buf, ", \"8192_sustained\": %.3f", 2.0 * n * (double)n * n * reps / (msr * 1e-3) / 1e12); js += buf;
      }
      cudaFree(A); cudaFree(B); cudaFree(C);
    }
    js += "}";
    cublasDestroy(h);
  }
  // sustained DMMA for ~2 s (power-capped clocks)
  {
    int iters = 1 << 16, threads = 512;
    double ms = T.best_ms([&] { for (int r = 0; r < 8; ++r) k_dmma<8><<<sms, threads>>>(out, iters, 1.0, 1e-9); }, 1);
    double tf = 8.0 * 2.0 * 256 * 8 * iters * (threads / 32.0) * sms / (ms * 1e-3) / 1e12;
    snprintf(buf, sizeof buf, ", \"dmma_tflops_sustained\": %.3f, \"dmma_sustained_ms\": %.1f", tf, ms); js += buf;
  }
  js += "}";
  printf("%s\n", js.c_str());
  return 0;
}
