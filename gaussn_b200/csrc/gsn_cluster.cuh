// One theta on several SMs: the log-likelihood kernel for latency-bound calls.
//
// The reference's samplers evaluate one parameter vector per call (gaussn/gausSN.py:321-323, :335), and a call with a
// handful of thetas leaves most of the GPU idle when each theta is confined to one CTA.  Here a thread-block cluster of
// 2, 4 or 8 four-warp CTAs (chosen per launch) works on ONE theta: the same work items, the same dependency-ordered list, the
// same factor workspace in global memory as gsn::logl_dmma_kernel -- only the control words (dataflow flags done[] / rdy[],
// the item counter, the failure index, the per-column partial sums) are those of the cluster's rank-0 CTA, which the other
// CTAs reach through distributed shared memory with cluster-scope acquire / release (Ctl<kCl> in gsn_chol_dmma.cuh).
// Every arithmetic operation of an item is independent of which warp executes it, so the result is bit-identical to the
// one-CTA shapes.  Each CTA keeps its own copy of the per-epoch vectors (computed redundantly: O(N) work).
//
// Per theta: cluster barrier; rank 0 finalises the previous theta, takes the next one and resets the control words;
// cluster barrier; all CTAs pull items until the list is exhausted.
#pragma once
#include <algorithm>
#include "gsn_chol_dmma.cuh"

namespace gsn {

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Shared memory of a cluster CTA.  (Padding the request above half an SM, to force one CTA per SM, was tried: a lone cluster
// is spread over SMs anyway, and the padding halves the number of clusters the GPU holds -- B = 100 at N = 1024: 3.0 ms
// as one CTA per theta, 2.1 ms as 100 two-CTA clusters sharing SMs.)
inline size_t logl_cluster_smem_bytes(int Npad, bool has_aux) { return logl_dmma_smem_bytes<CfgCluster>(Npad, has_aux); }

template <int KIND>
__global__ void __launch_bounds__(CfgCluster::kWarps * 32, 1)
logl_cluster_kernel(ProblemDev pd, const double* __restrict__ theta, long long B, long long ld,
                    const double* __restrict__ hostmean, GatherOut out, int* __restrict__ info,
                    unsigned flags, double* workspace, unsigned* counter) {
  using Cfg = CfgCluster;
  constexpr int kCl = 2;   // "control words are remote"; the actual cluster size is whatever the launch asked for
  constexpr int kWarps = Cfg::kWarps, kThreads = Cfg::kWarps * 32, kStages = Cfg::kStages;
  using CtaShared = CtaSharedT<Cfg::kWarps, Cfg::kChunks>;
  using C = Ctl<kCl>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaShared& sh = *reinterpret_cast<CtaShared*>(smem_raw);   // control words: only rank 0's copy is used (through Ctl)
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned crank = cluster_ctarank();
  const int Npad = pd.Npad, nblk = pd.nblk;
  const SlotGeom g(Npad);
  const int nct = g.nct;
  double* slot = workspace + (size_t)cluster_id_x() * g.total;
  unsigned char* sp = smem_raw + (sizeof(CtaShared) + 15) / 16 * 16;
  double2* ring = reinterpret_cast<double2*>(sp) + (size_t)(tid >> 5) * kStages * 8 * 32 + lane;
  sp += (size_t)kWarps * kStages * 8 * 512;
  double* vec = (Npad <= kSmemVecMaxNpad) ? reinterpret_cast<double*>(sp) : slot + g.vec_off;
  double* xs = vec;
  double* bs = vec + Npad;
  double* aux = kernel_has_aux<KIND>() ? vec + 2 * (size_t)Npad : vec;
  double* rs = slot + g.vec_off + 3 * (size_t)Npad;
  double* theta_s = sh.theta;   // each CTA's own copy

  for (int i = tid; i < 64; i += kThreads) sh.exptab[i] = kExp2Tab[i];

  long long prev = -1;   // theta whose items have all been issued and that still has to be reduced (rank 0, thread 0)
  for (;;) {
    cluster_sync_all();   // every item of the previous theta is stored; nobody reads its control words any more
    if (crank == 0) {
      if (tid == 0 && prev >= 0) {   // logL = -1/2 (N ln 2 pi + ln det + z^T z)      gausSN.py:151-158
        double zsum = 0.0, ldsum = 0.0;
        for (int jj = 0; jj < nblk; ++jj) { zsum += sh.zzj[jj]; ldsum += sh.ldj[jj]; }
        double ll = -0.5 * (pd.nlog2pi + ldsum + zsum);
        int inf = 0;
        if (sh.fail != 0) { ll = nan(""); inf = failed_row(pd, sh.fail); }
        else if (!isfinite(ll)) inf = -1;
        if ((flags & GSN_FLAG_NEG_INF) && !isfinite(ll)) ll = -INFINITY;
        for (int r = 0; r < out.n; ++r) out.ptr[r][out.offset + prev] = ll;
        if (info) info[prev] = inf;
      }
      __syncthreads();
      if (tid == 0) { sh.next_theta = (int)atomicAdd(counter, 1u); sh.fail = 0; sh.next_item = 0; }
      for (int c = tid; c <= nblk; c += kThreads) { sh.done[c] = 0; sh.rdy[c] = 0; sh.zzj[c] = 0.0; sh.ldj[c] = 0.0; }
    }
    cluster_sync_all();   // the control words of the next theta are in place
    const long long it = C::ld_relaxed(&sh.next_theta);
    if (it >= B) break;
    prev = it;

    for (int k = tid; k < pd.P; k += kThreads) {
      const int col = pd.theta_col[k];
      theta_s[k] = (col >= 0) ? theta[it * ld + col] : pd.theta_fixed[k];
    }
    __syncthreads();
    const double* kp = theta_s;
    const double* mp = theta_s + pd.nk;
    const double* lp = theta_s + pd.nk + pd.nm;
    const KernelConsts kc = kernel_consts<KIND>(kp);

    // per-epoch vectors, one copy per CTA (gausSN.py:139-142); see logl_dmma_kernel for the non-finite guard
    int bad = (tid == 0) && !(isfinite(kc.amp) && isfinite(kc.c1) && isfinite(kc.c2) && isfinite(kc.c3) && isfinite(kc.c4));
    for (int n = tid; n < Npad; n += kThreads) {
      double x_s = 0.0, b = 0.0, r = 0.0, ax = 0.0;
      const int u = pd.rowmap[n];
      if (u >= 0) {
        lens_eval(pd.lens_kind, lp, pd.img[n], pd.x[n], x_s, b);
        const double m = hostmean ? hostmean[it * pd.N + u] : mean_eval(pd.mean_kind, mp, x_s);
        r = b * m - pd.y[n];
        if (kernel_has_aux<KIND>()) ax = kernel_aux<KIND>(kc, x_s);
        bad |= !isfinite(x_s) || !isfinite(ax);
      }
      xs[n] = x_s; bs[n] = b; rs[n] = r;
      if (kernel_has_aux<KIND>()) aux[n] = ax;
    }
    if (__syncthreads_or(bad)) {   // every CTA of the cluster sees the same inputs and takes this branch together
      if (crank == 0 && tid == 0) {
        const double bad_ll = (flags & GSN_FLAG_NEG_INF) ? -INFINITY : nan("");
        for (int r = 0; r < out.n; ++r) out.ptr[r][out.offset + it] = bad_ll;
        if (info) info[it] = -1;
      }
      prev = -1;
      continue;
    }

    for (;;) {
      int idx = 0;
      if (lane == 0) idx = C::fetch_add(&sh.next_item, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx >= pd.n_items || C::ld_relaxed(&sh.fail) != 0) break;
      const unsigned item = pd.items[idx];
      const int kb0 = (int)(item >> 24), c = (int)((item >> 12) & 0xfffu), j = (int)(item & 0xfffu);
      if (c != j) {
        OffItem<KIND, kStages, kCl> it_;
        build_item<KIND, false>(pd, kc, sh.exptab, xs, bs, aux, c, j, lane, ring);
        it_.load_built(ring);
        if (!it_.gemm(sh, slot, nct, c, j, kb0, j, lane, ring)) break;
        if (!wait_flag<kCl>(&sh.rdy[j], 0, &sh.fail)) break;
        it_.solve(slot, g, c, j, lane, ring);
        __syncwarp();
        if (lane == 0) C::st_release(&sh.done[c], j + 1);
      } else {
        DiagItem<KIND, kStages, true, kCl> it_;
        build_item<KIND, false>(pd, kc, sh.exptab, xs, bs, aux, c, c, lane, ring);
        it_.load_built(ring, rs, c, lane);
        if (!it_.gemm(sh, slot, nct, c, kb0, c, lane, ring)) break;
        it_.factor(sh, slot, g, c, lane, 0);
      }
    }
  }
  cluster_sync_all();   // rank 0's shared memory must outlive the last remote read of its control words
}

}  // namespace gsn
