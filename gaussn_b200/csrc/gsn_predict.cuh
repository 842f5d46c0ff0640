// Batched GP prediction (conditional mean and covariance) on sm_100a, built from the work items of the
// log-likelihood kernel (gsn_chol_dmma.cuh).
//
// Replaces gaussn/gausSN.py:362-403 (GP.predict) as gaussn/utils.py:193-209 drives it: for every posterior sample
// theta, de-lens the observations of one band and condition the GP on them at the epochs x'.
//
// One factorisation does it all.  With V the conditioning epochs and U the prediction epochs, take the augmented matrix
//
//        A = [ C_VV   .   ]      C_VV = b b^T o K(x~, x~) + diag(yerr^2)      (the likelihood's covariance of these rows)
//            [ K_UV  K_UU ]      K_UV = K(x', x~) diag(b),  K_UU = K(x', x')
//
// and run the blocked left-looking Cholesky of the likelihood kernel over its first nVb block columns only:
//   block rows of V:  L (as in the likelihood), with the residual row z = L^-1 (b m(x~) - y) riding along;
//   block rows of U:  W = K_UV L^-T                                   (ordinary off-diagonal items: GEMM + solve);
//   trailing blocks:  S = W W^T - K_UU = -(conditional covariance)    ("trail" items: build + GEMM, no solve, no factor);
//   residual row against the U columns:  W z - m(x') = -(conditional mean).
// Scaling the rows of V by the magnifications instead of dividing the data by them, as utils.py:204 does, is the same
// linear algebra: C_VV = D (K + diag((yerr/b)^2)) D and K_UV = K(x', x~) D with D = diag(b), so D cancels in both moments.
//
// The reference solves with LU (np.linalg.solve), which also "works" when C_VV is not positive definite; here such a theta
// gives NaN moments and info = index of the failing pivot, the likelihood's convention.
#pragma once
#include "gsn_chol_dmma.cuh"
#include "gsn_cluster.cuh"

namespace gsn {

constexpr unsigned kTrailBit = 1u << 23;   // item word flag: trailing block (no solve / factor; result goes to the output)

struct PredictDev {
  int row0, nV;          // rows [row0, row0 + nV) of the problem are conditioned on
  int u0, nTot;          // prediction epochs occupy rows [u0, nTot) of the augmented matrix; u0 = nV rounded up to 32
  int Npad, nblk, nVb;   // padded order, its block count, block columns that are factorised (u0 / 32)
  int M;                 // nTot - u0
  const double* xprime;  // [M]
  const double* mean_v;  // [B, nV] host-evaluated mean at the shifted epochs, or null
  const double* mean_u;  // [B, M]  host-evaluated mean at x', or null
  const unsigned* items;
  int n_items;
  double* expectation;   // [B, M]
  double* variance;      // [B, M, M] or null
};

// Work-item order: the V rows exactly as make_item_order_rows sweeps them (groups of two), then the U rows in groups of
// two -- first their blocks in the factorised columns, then their trailing blocks.  Every item follows all it waits for.
inline int make_item_order_predict(int nVb, int nblk, bool want_cov, unsigned* out /* room for nblk (nblk + 1) / 2 */) {
  int n = 0;
  auto word = [&](int c, int j) { return ((unsigned)c << 12) | (unsigned)j; };
  const int group = 2;
  for (int g0 = 0; g0 < nVb; g0 += group) {
    const int g1 = g0 + group < nVb ? g0 + group : nVb;
    for (int j = 0; j < g0; ++j)
      for (int c = g0; c < g1; ++c) out[n++] = word(c, j);
    for (int c = g0; c < g1; ++c) {
      for (int j = g0; j < c; ++j) out[n++] = word(c, j);
      out[n++] = word(c, c);
    }
  }
  for (int g0 = nVb; g0 < nblk; g0 += group) {
    const int g1 = g0 + group < nblk ? g0 + group : nblk;
    for (int j = 0; j < nVb; ++j)
      for (int c = g0; c < g1; ++c) out[n++] = word(c, j);
    for (int c = g0; c < g1; ++c) {
      if (want_cov)
        for (int j = nVb; j < c; ++j) out[n++] = kTrailBit | word(c, j);
      out[n++] = kTrailBit | word(c, c);
    }
  }
  return n;
}

// Column-major variant (look-ahead on the diagonal blocks) for the cluster mode: the widest front of independent items.
inline int make_item_order_predict_cols(int nVb, int nblk, bool want_cov, unsigned* out) {
  int n = 0;
  auto word = [&](int c, int j) { return ((unsigned)c << 12) | (unsigned)j; };
  out[n++] = word(0, 0);
  for (int j = 0; j < nVb; ++j) {
    if (j + 1 < nblk) out[n++] = word(j + 1, j);
    if (j + 1 < nVb) out[n++] = word(j + 1, j + 1);
    for (int c = j + 2; c < nblk; ++c) out[n++] = word(c, j);
  }
  for (int c = nVb; c < nblk; ++c) {
    if (want_cov)
      for (int j = nVb; j < c; ++j) out[n++] = kTrailBit | word(c, j);
    out[n++] = kTrailBit | word(c, c);
  }
  return n;
}

__host__ __device__ inline size_t predict_slot_doubles(int Npad) { return SlotGeom(Npad).vec_off + (size_t)5 * Npad; }
inline size_t predict_smem_bytes(int Npad) {
  size_t b = (sizeof(CtaSharedT<CfgWide::kWarps, CfgWide::kChunks>) + 15) / 16 * 16;
  b += (size_t)CfgWide::kWarps * CfgWide::kStages * 8 * 512;
  if (Npad <= kSmemVecMaxNpad) b += (size_t)Npad * 5 * sizeof(double);
  return b;
}

inline size_t predict_cluster_smem_bytes(int Npad) { return predict_smem_bytes(Npad); }

// kCl == 1: one CTA per theta (batches of posterior samples).  kCl > 1: one theta per thread-block cluster, as in
// gsn_cluster.cuh -- for GP.predict called with a single parameter set on a long light curve.
template <int KIND, int kCl>
__global__ void __launch_bounds__(CfgWide::kWarps * 32, kCl == 1 ? CfgWide::kCtasPerSm : 1)
predict_dmma_kernel(ProblemDev pd, PredictDev pa, const double* __restrict__ theta, long long B, long long ld,
                    int* __restrict__ info, double* workspace, unsigned* counter) {
  using Cfg = CfgWide;
  using C = Ctl<kCl>;
  constexpr int kWarps = Cfg::kWarps, kThreads = Cfg::kWarps * 32, kStages = Cfg::kStages;
  using CtaShared = CtaSharedT<Cfg::kWarps, Cfg::kChunks>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtaShared& sh = *reinterpret_cast<CtaShared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Npad = pa.Npad, nblk = pa.nblk, M = pa.M;
  const SlotGeom g(Npad);
  const int nct = g.nct;
  const unsigned crank = kCl == 1 ? 0u : cluster_ctarank();
  double* slot = workspace + (size_t)(kCl == 1 ? blockIdx.x : cluster_id_x()) * predict_slot_doubles(Npad);
  unsigned char* sp = smem_raw + (sizeof(CtaShared) + 15) / 16 * 16;
  double2* ring = reinterpret_cast<double2*>(sp) + (size_t)warp * kStages * 8 * 32 + lane;
  sp += (size_t)kWarps * kStages * 8 * 512;
  double* vec = (Npad <= kSmemVecMaxNpad) ? reinterpret_cast<double*>(sp) : slot + g.vec_off;
  double* xs = vec;
  double* bs = vec + Npad;
  double* rs = vec + 2 * (size_t)Npad;
  double* ns = vec + 3 * (size_t)Npad;
  double* aux = vec + 4 * (size_t)Npad;
  const AugGeom ag{pa.nV, pa.u0, pa.nTot, ns};
  const int r_in = lane >> 2, cp = (lane & 3) * 2;

  for (int i = tid; i < 64; i += kThreads) sh.exptab[i] = kExp2Tab[i];

  // a theta whose factorisation failed (C_VV not positive definite, or NaN): the moments are undefined
  auto close_theta = [&](long long t, int failed) {
    if (failed != 0) {
      for (int i = tid; i < M; i += kThreads) pa.expectation[t * M + i] = nan("");
      if (pa.variance) for (long long i = tid; i < (long long)M * M; i += kThreads) pa.variance[t * (long long)M * M + i] = nan("");
    }
    if (tid == 0 && info) info[t] = failed;
  };

  long long prev = -1;   // cluster mode: theta whose items have all been issued and that rank 0 still has to close
  for (;;) {
    long long it;
    if constexpr (kCl == 1) {
      __syncthreads();
      if (tid == 0) sh.next_theta = (int)atomicAdd(counter, 1u);
      __syncthreads();
      it = sh.next_theta;
    } else {
      cluster_sync_all();   // every item of the previous theta is done
      if (crank == 0) {
        if (prev >= 0) close_theta(prev, sh.fail);
        __syncthreads();
        if (tid == 0) { sh.next_theta = (int)atomicAdd(counter, 1u); sh.fail = 0; sh.next_item = 0; }
        for (int c = tid; c <= nblk; c += kThreads) { sh.done[c] = 0; sh.rdy[c] = 0; sh.zzj[c] = 0.0; sh.ldj[c] = 0.0; }
      }
      cluster_sync_all();   // control words of the next theta are in place (rank 0's copy; see Ctl)
      it = C::ld_relaxed(&sh.next_theta);
    }
    if (it >= B) break;
    prev = it;
    double* expct = pa.expectation + it * M;
    double* var = pa.variance ? pa.variance + it * (long long)M * M : nullptr;

    for (int k = tid; k < pd.P; k += kThreads) {
      const int col = pd.theta_col[k];
      sh.theta[k] = (col >= 0) ? theta[it * ld + col] : pd.theta_fixed[k];
    }
    if constexpr (kCl == 1) {
      if (tid == 0) { sh.fail = 0; sh.next_item = 0; }
      for (int c = tid; c <= nblk; c += kThreads) { sh.done[c] = 0; sh.rdy[c] = 0; sh.zzj[c] = 0.0; sh.ldj[c] = 0.0; }
    }
    __syncthreads();
    const double* kp = sh.theta;
    const double* mp = sh.theta + pd.nk;
    const double* lp = sh.theta + pd.nk + pd.nm;
    const KernelConsts kc = kernel_consts<KIND>(kp);

    // ---- per-epoch vectors of the augmented system
    int bad = (tid == 0) && !(isfinite(kc.amp) && isfinite(kc.c1) && isfinite(kc.c2) && isfinite(kc.c3) && isfinite(kc.c4));
    for (int n = tid; n < Npad; n += kThreads) {
      double x_s = 0.0, b = 0.0, r = 0.0, noise = 0.0, ax = 0.0;
      if (n < pa.nV) {            // observed epoch: shifted, magnified (gausSN.py:139-142; utils.py:202-204)
        const int row = pd.introw[pa.row0 + n];
        lens_eval(pd.lens_kind, lp, pd.img[row], pd.x[row], x_s, b);
        const double m = pa.mean_v ? pa.mean_v[it * pa.nV + n] : mean_eval(pd.mean_kind, mp, x_s);
        r = b * m - pd.y[row];
        noise = pd.yerr2[row];
      } else if (n >= pa.u0 && n < pa.nTot) {   // prediction epoch: intrinsic frame, "observed" value 0 so that r = m(x')
        x_s = pa.xprime[n - pa.u0];
        b = 1.0;
        r = pa.mean_u ? pa.mean_u[it * M + (n - pa.u0)] : mean_eval(pd.mean_kind, mp, x_s);
      }
      if (kernel_has_aux<KIND>() && !ag.is_pad(n)) ax = kernel_aux<KIND>(kc, x_s);
      bad |= !isfinite(x_s) || !isfinite(ax);
      xs[n] = x_s; bs[n] = b; rs[n] = r; ns[n] = noise; aux[n] = ax;
    }
    if (__syncthreads_or(bad)) {   // in a cluster every CTA sees the same inputs and takes this branch together
      if (crank == 0) close_theta(it, -1);
      prev = -1;
      continue;
    }

    for (;;) {
      int idx = 0;
      if (lane == 0) idx = C::fetch_add(&sh.next_item, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx >= pa.n_items || C::ld_relaxed(&sh.fail) != 0) break;
      const unsigned item = pa.items[idx];
      const bool trail = (item & kTrailBit) != 0;
      const int c = (int)((item >> 12) & 0x7ffu), j = (int)(item & 0xfffu);
      if (!trail) {
        if (c != j) {
          OffItem<KIND, kStages, kCl> it_;
          build_item<KIND, false, true>(pd, kc, sh.exptab, xs, bs, aux, c, j, lane, ring, &ag);
          it_.load_built(ring);
          if (!it_.gemm(sh, slot, nct, c, j, 0, j, lane, ring)) break;
          if (!wait_flag<kCl>(&sh.rdy[j], 0, &sh.fail)) break;
          it_.solve(slot, g, c, j, lane, ring);
          __syncwarp();
          if (lane == 0) C::st_release(&sh.done[c], j + 1);
        } else {
          DiagItem<KIND, kStages, true, kCl> it_;
          build_item<KIND, false, true>(pd, kc, sh.exptab, xs, bs, aux, c, c, lane, ring, &ag);
          it_.load_built(ring, rs, c, lane);
          if (!it_.gemm(sh, slot, nct, c, 0, c, lane, ring)) break;
          it_.factor(sh, slot, g, c, lane, warp);
        }
      } else if (c != j) {
        // trailing off-diagonal block: -(conditional covariance)(c, j), mirrored into (j, c)
        OffItem<KIND, kStages, kCl> it_;
        build_item<KIND, false, true>(pd, kc, sh.exptab, xs, bs, aux, c, j, lane, ring, &ag);
        it_.load_built(ring);
        if (!it_.gemm(sh, slot, nct, c, j, 0, pa.nVb, lane, ring)) break;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int row = 32 * c + 8 * a + r_in - pa.u0, col = 32 * j + 8 * b + cp - pa.u0;
            if (row < M) {
              if (col < M) { var[(size_t)row * M + col] = -it_.S[a][b].x; var[(size_t)col * M + row] = -it_.S[a][b].x; }
              if (col + 1 < M) { var[(size_t)row * M + col + 1] = -it_.S[a][b].y; var[(size_t)(col + 1) * M + row] = -it_.S[a][b].y; }
            }
          }
      } else {
        // trailing diagonal block (lower triangle, mirrored) and this block's slice of the conditional mean
        DiagItem<KIND, kStages, true, kCl> it_;
        build_item<KIND, false, true>(pd, kc, sh.exptab, xs, bs, aux, c, c, lane, ring, &ag);
        it_.load_built(ring, rs, c, lane);
        if (!it_.gemm(sh, slot, nct, c, 0, pa.nVb, lane, ring)) break;
        if (var) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
              const double2 v = it_.S[a * (a + 1) / 2 + b];
              const int row = 32 * c + 8 * a + r_in - pa.u0, col = 32 * c + 8 * b + cp - pa.u0;
              if (row < M) {
                if (col <= row) { var[(size_t)row * M + col] = -v.x; var[(size_t)col * M + row] = -v.x; }
                if (col + 1 <= row) { var[(size_t)row * M + col + 1] = -v.y; var[(size_t)(col + 1) * M + row] = -v.y; }
              }
            }
        }
        if (r_in == 0) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int col = 32 * c + 8 * b + cp - pa.u0;
            if (col < M) expct[col] = -it_.Z[b].x;
            if (col + 1 < M) expct[col + 1] = -it_.Z[b].y;
          }
        }
      }
    }

    if constexpr (kCl == 1) {
      __syncthreads();
      close_theta(it, sh.fail);
    }
  }
  if constexpr (kCl > 1) cluster_sync_all();   // rank 0's shared memory must outlive the last remote read
}

}  // namespace gsn
