// Device-side plugin zoo: covariance functions, mean functions, lensing models.
//
// Closed-form restatements of the reference's plugin classes for sm_100a.  Every
// function names the reference lines whose arithmetic it follows (paths relative
// to the reference checkout).  All arithmetic is FP64 (the reference forces
// jax_enable_x64, gaussn/gausSN.py:23).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/gsn.h"

namespace gsn {

constexpr int kMaxKernelParams = 5;
constexpr int kMaxMeanParams = 8;
constexpr int kMaxImages = 16;
constexpr int kMaxLensPerImage = 5;
constexpr int kMaxTheta = kMaxKernelParams + kMaxMeanParams + kMaxLensPerImage * (kMaxImages - 1);

__host__ __device__ inline int kernel_nparams(int kind) {
  switch (kind) {
    case GSN_K_EXPSQUARED: case GSN_K_EXPONENTIAL: case GSN_K_MATERN32: case GSN_K_MATERN52: case GSN_K_OU: return 2;
    case GSN_K_CONSTANT: return 1;
    case GSN_K_DOTPRODUCT: return 0;
    case GSN_K_RATQUAD: case GSN_K_PERIODIC: return 3;
    case GSN_K_GIBBS: return 5;
    case GSN_K_JJ: return 4;
  }
  return -1;
}
__host__ __device__ inline int mean_nparams(int kind) {
  switch (kind) {
    case GSN_M_UNIFORM: return 1;
    case GSN_M_SIN: case GSN_M_GAUSSIAN: return 3;
    case GSN_M_EXPFUNCTION: return 2;
    case GSN_M_BAZIN2009: return 5;
    case GSN_M_KARPENKA2012: case GSN_M_VILLAR2019: return 6;
    case GSN_M_HOST: return 0;
  }
  return -1;
}
__host__ __device__ inline int lens_nparams_per_image(int kind) {
  switch (kind) {
    case GSN_L_NONE: return 0;
    case GSN_L_CONSTANT: return 2;
    case GSN_L_SIGMOID: case GSN_L_SINUSOIDAL: return 5;
  }
  return -1;
}

// ---------------------------------------------------------------------------
// Covariance functions.  KernelConsts holds what depends on theta only, computed
// once per theta; cov_elem() is the per-element work.  `aux` is a per-epoch
// quantity hoisted out of the N^2 loop (Gibbs: tau(x), JJ: the period p(x)).
// ---------------------------------------------------------------------------
struct KernelConsts {
  double amp;  // leading factor (A^2, A or c)
  double c1;   // kind-specific
  double c2;
  double c3;
  double c4;
};

template <int KIND>
__device__ __forceinline__ KernelConsts kernel_consts(const double* kp) {
  KernelConsts c{0.0, 0.0, 0.0, 0.0, 0.0};
  if (KIND == GSN_K_EXPSQUARED) {         // kernels.py:55   A**2 * exp(-d**2/(2*tau**2))
    c.amp = kp[0] * kp[0];
    c.c1 = 1.0 / (2.0 * (kp[1] * kp[1]));
  } else if (KIND == GSN_K_EXPONENTIAL) { // kernels.py:109-110   A**2 * exp(-sqrt(d**2)/tau)
    c.amp = kp[0] * kp[0];
    c.c1 = kp[1];
  } else if (KIND == GSN_K_CONSTANT) {    // kernels.py:157
    c.amp = kp[0];
  } else if (KIND == GSN_K_MATERN32 || KIND == GSN_K_MATERN52 || KIND == GSN_K_OU) {  // :250-251, :305-308, :498-499
    c.amp = kp[0];
    c.c1 = kp[1];
    c.c2 = 3.0 * (kp[1] * kp[1]);         // Matern52: 3*(l**2)
  } else if (KIND == GSN_K_RATQUAD) {     // kernels.py:366   A**2 * (1 + d**2/(2*alpha*tau**2))
    c.amp = kp[0] * kp[0];
    c.c1 = 2.0 * kp[2] * (kp[1] * kp[1]);
  } else if (KIND == GSN_K_GIBBS) {       // kernels.py:438-447
    c.amp = kp[0] * kp[0];
    c.c1 = kp[1];                          // lambda
    c.c2 = kp[2];                          // p
    c.c3 = kp[3];                          // mu
    c.c4 = kp[4];                          // sigma
  } else if (KIND == GSN_K_PERIODIC) {    // kernels.py:523-526
    c.amp = kp[0];
    c.c1 = kp[1];                          // l
    c.c2 = kp[2];                          // p
  } else if (KIND == GSN_K_JJ) {          // kernels.py:552-558
    c.amp = kp[0];
    c.c1 = kp[1];                          // l
    c.c2 = kp[2];                          // m
    c.c3 = kp[3];                          // b
  }
  return c;
}

template <int KIND>
__device__ __forceinline__ constexpr bool kernel_has_aux() { return KIND == GSN_K_GIBBS || KIND == GSN_K_JJ; }

// per-epoch auxiliary value (only meaningful when kernel_has_aux<KIND>())
template <int KIND>
__device__ __forceinline__ double kernel_aux(const KernelConsts& c, double x) {
  if (KIND == GSN_K_GIBBS) {
    // tau(x) = lambda * (1 - p * N(x | mu, sigma))      kernels.py:438-441
    const double sigma = c.c4;
    const double d = x - c.c3;
    const double normal = exp(-(d * d) / (2.0 * (sigma * sigma))) / (sigma * sqrt(2.0 * M_PI));
    return c.c1 * (1.0 - (c.c2 * normal));
  }
  if (KIND == GSN_K_JJ) return c.c2 * x + c.c3;  // _p(x), kernels.py:549-550
  return 0.0;
}

// k(x_i, x_j) as the reference's _covariance writes element [i, j].  For the JJ
// kernel the period comes from the column epoch j (kernels.py:555 broadcasting),
// so the value is not symmetric in (i, j).
template <int KIND>
__device__ __forceinline__ double cov_elem(const KernelConsts& c, double xi, double xj, double auxi, double auxj) {
  const double d = xi - xj;
  if (KIND == GSN_K_EXPSQUARED) {
    return c.amp * exp(-(d * d) * c.c1);
  } else if (KIND == GSN_K_EXPONENTIAL) {
    return c.amp * exp(-sqrt(d * d) / c.c1);
  } else if (KIND == GSN_K_CONSTANT) {
    return c.amp;
  } else if (KIND == GSN_K_DOTPRODUCT) {   // kernels.py:196
    return xi * xj;
  } else if (KIND == GSN_K_MATERN32) {
    const double s = sqrt(3.0 * (d * d)) / c.c1;
    return c.amp * ((1.0 + s) * exp(-s));
  } else if (KIND == GSN_K_MATERN52) {
    const double r2 = d * d;
    const double s = sqrt(5.0 * r2) / c.c1;
    const double a = 1.0 + s + (5.0 * r2 / c.c2);
    return c.amp * a * exp(-s);
  } else if (KIND == GSN_K_RATQUAD) {
    return c.amp * (1.0 + (d * d) / c.c1);
  } else if (KIND == GSN_K_GIBBS) {
    const double root_num = 2.0 * auxi * auxj;
    const double denom = (auxi * auxi) + (auxj * auxj);
    return c.amp * sqrt(root_num / denom) * exp(-(d * d) / denom);
  } else if (KIND == GSN_K_OU) {
    return c.amp * exp(-sqrt(d * d) / c.c1);
  } else if (KIND == GSN_K_PERIODIC) {
    const double s = sin(M_PI * (sqrt(d * d) / c.c2)) / c.c1;
    return c.amp * exp(-2.0 * (s * s));
  } else if (KIND == GSN_K_JJ) {
    const double s = sin(M_PI * (sqrt(d * d) / auxj)) / c.c1;
    return c.amp * exp(-2.0 * (s * s));
  }
  return 0.0;
}

// The value that enters the factorisation: (cov + cov^T)/2, the symmetrisation
// jnp.linalg.cholesky applies (gausSN.py:150).  Only JJ is asymmetric.
template <int KIND>
__device__ __forceinline__ double cov_elem_sym(const KernelConsts& c, double xi, double xj, double auxi, double auxj) {
  if (KIND == GSN_K_JJ) {
    return 0.5 * (cov_elem<KIND>(c, xi, xj, auxi, auxj) + cov_elem<KIND>(c, xj, xi, auxj, auxi));
  }
  return cov_elem<KIND>(c, xi, xj, auxi, auxj);
}

// ---------------------------------------------------------------------------
// Mean functions (meanfuncs.py), evaluated at the shifted epoch t.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double mean_eval(int kind, const double* mp, double t) {
  switch (kind) {
    case GSN_M_UNIFORM:       // meanfuncs.py:45
      return mp[0];
    case GSN_M_SIN:           // meanfuncs.py:172   A*sin(t*w + phi)
      return mp[0] * sin((t * mp[1]) + mp[2]);
    case GSN_M_GAUSSIAN: {    // meanfuncs.py:223-224
      const double sigma = mp[2];
      const double d = t - mp[1];
      const double e = -(d * d) / (2.0 * (sigma * sigma));
      return mp[0] * exp(e) / (sigma * sqrt(2.0 * M_PI));
    }
    case GSN_M_EXPFUNCTION:   // meanfuncs.py:272
      return mp[0] * exp(t * mp[1]);
    case GSN_M_BAZIN2009: {   // meanfuncs.py:334-336
      const double a = exp(-(t - mp[2]) / mp[3]);
      const double b = 1.0 + exp((t - mp[2]) / mp[4]);
      return mp[0] * (a / b) + mp[1];
    }
    case GSN_M_KARPENKA2012: {  // meanfuncs.py:357-362, :404-407
      const double A = exp(mp[0]);
      const double B = exp(mp[1]);
      const double dt1 = t - mp[2];
      const double a = 1.0 + (B * (dt1 * dt1));
      const double b = exp(-(t - mp[3]) / mp[4]);
      const double c = 1.0 + exp(-(t - mp[3]) / mp[5]);
      return A * a * (b / c);
    }
    case GSN_M_VILLAR2019: {  // meanfuncs.py:472-483
      const double A = mp[0], beta = mp[1], t1 = mp[2], t0 = mp[3], Tfall = mp[4], Trise = mp[5];
      const double denom = 1.0 + exp(-(t - t0) / Trise);
      double mu;
      if (t < t1) {
        mu = A + (beta * (t - t0));
      } else {
        const double constant = A + (beta * (t1 - t0));
        mu = constant * exp(-(t - t1) / Tfall);
      }
      return mu / denom;
    }
  }
  return 0.0;
}

// ---------------------------------------------------------------------------
// Lensing models (lensingmodels.py): shifted epoch and magnification of one
// epoch that belongs to image `img` (0 = reference image).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void lens_eval(int kind, const double* lp, int img, double x, double& xs, double& b) {
  if (kind == GSN_L_NONE) {          // lensingmodels.py:13-14
    xs = x; b = 1.0; return;
  }
  if (kind == GSN_L_CONSTANT) {      // lensingmodels.py:29-30, :64, :77
    double delta = 0.0, beta = 1.0;
    if (img > 0) { delta = lp[2 * (img - 1)]; beta = lp[2 * (img - 1) + 1]; }
    xs = x - delta; b = beta; return;
  }
  // five parameters per extra image; image 1 is (0, 1, 0, 0, 0): :136-140, :259-263
  double delta = 0.0, beta0 = 1.0, beta1 = 0.0, p3 = 0.0, p4 = 0.0;
  if (img > 0) {
    const double* q = lp + 5 * (img - 1);
    delta = q[0]; beta0 = q[1]; beta1 = q[2]; p3 = q[3]; p4 = q[4];
  }
  xs = x - delta;
  if (kind == GSN_L_SIGMOID) {       // lensingmodels.py:196-197   (p3, p4) = (r, t0)
    const double denom = 1.0 + exp(-p3 * (xs - p4));
    b = beta0 + (beta1 / denom);
  } else {                           // lensingmodels.py:319-320   (p3, p4) = (t0, T); T = 0 for image 1 gives NaN
    const double sinusoid = sin((xs - p3) / p4);
    b = beta0 + (beta1 * sinusoid);
  }
}

}  // namespace gsn
