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

// ---------------------------------------------------------------------------
// exp() for the N^2 covariance build.  libdevice's exp costs ~22 DFMA-equivalents
// per element (measured, profiles/r01_fp64_peaks.json), i.e. ~12% on top of the
// Cholesky at N = 512.  This one reduces with a 64-entry table of 2^(j/64) and a
// degree-5 polynomial: x = (64 k + j) ln2/64 + r, |r| <= ln2/128,
// exp(x) = 2^k * T[j] * (1 + p(r)); truncation error r^6/720 < 4e-17, total error
// about 1 ulp.  Results below 2^-1021 (x < -708) are flushed to zero.
// ---------------------------------------------------------------------------
static __constant__ double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// Polynomial and reduction constants live in constant memory so that the FP64
// instructions take them as constant-bank operands instead of rebuilding each
// 64-bit immediate with two moves per use.
static __constant__ double kExpC[8] = {
    0x1.71547652b82fep+6,    // 64 / ln 2
    6755399441055744.0,      // 1.5 * 2^52: round-to-nearest-integer trick
    -0x1.62e42fee00000p-7,   // -(ln 2 / 64), high part (32 significant bits)
    -0x1.a39ef35793c76p-39,  // -(ln 2 / 64), low part
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};

// Branch-free core: the caller guarantees x is not NaN (the hot kernel checks the
// kernel constants and the shifted epochs once per theta).  Results that would be
// subnormal (x < -708.4) are flushed to zero, overflow gives +inf.
// kBranch: the range check as a (rarely taken) branch instead of selects -- fewer live registers; the register shape of
// gsn_onchip.cuh, which interleaves the builds of a whole tile column anyway, is 1-7 % faster with it.
template <bool kBranch = false>
__device__ __forceinline__ double fast_exp(double x, const double* __restrict__ tab) {
#ifdef GSN_ACCURATE_EXP   // accuracy experiments (tools/arbiter_check.py): libdevice's exp, < 1 ulp
  return exp(x);
#endif
  const double t = fma(x, kExpC[0], kExpC[1]);
  const int n = __double2loint(t);
  const double nf = t - kExpC[1];
  double r = fma(nf, kExpC[2], x);
  r = fma(nf, kExpC[3], r);
  double p = fma(r, kExpC[4], kExpC[5]);
  p = fma(p, r, kExpC[6]);
  p = fma(p, r, kExpC[7]);
  p = fma(p * r, r, r);                       // r + r^2/2 + r^3/6 + r^4/24 + r^5/120
  const double T = tab[n & 63];
  const double v = fma(T, p, T);              // in [1, 2)
  const int k = n >> 6;
  int hi = __double2hiint(v) + (k << 20);
  int lo = __double2loint(v);
  // |x| >= 708 (or +-inf): underflow is flushed to zero, overflow (x > 709.782712893384 = 0x40862e42'fefa39ef) saturates to
  // +inf.  Selects on the integer halves, no branch: a branch here made every exponential of the covariance build its own
  // reconvergence region, so the eight independent exponentials of a tile row ran one after the other instead of
  // interleaved -- eight dependent chains of 13 FP64 operations in sequence.
  const int xh = __double2hiint(x);
  if constexpr (kBranch) {
    if ((xh & 0x7fffffff) >= 0x40862000) {
      if (xh < 0) return 0.0;
      if (x > 709.782712893384) return __hiloint2double(0x7ff00000, 0);
    }
    return __hiloint2double(hi, lo);
  }
  const bool under = xh < 0 && (xh & 0x7fffffff) >= 0x40862000;
  const bool over = xh > 0x40862e42 || (xh == 0x40862e42 && (unsigned)__double2loint(x) > 0xfefa39efu);
  hi = under ? 0 : over ? 0x7ff00000 : hi;
  lo = (under || over) ? 0 : lo;
  return __hiloint2double(hi, lo);
}

// Full-range wrapper for callers that cannot rule out NaN.
template <bool kBranch = false>
__device__ __forceinline__ double fast_exp_checked(double x, const double* __restrict__ tab) {
  return (x != x) ? x : fast_exp<kBranch>(x, tab);
}

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
// once per theta (reciprocals included, so that the N^2 loop has no division);
// cov_elem() is the per-element work.  `aux` is a per-epoch quantity hoisted out
// of the N^2 loop (Gibbs: tau(x), JJ: 1 / p(x)).
//
// Deviation from the reference's operation order, by design: the reference
// divides per element (e.g. -d^2 / (2 tau^2)); here the divisor's reciprocal is
// formed once per theta and multiplied in, and sqrt(c d^2) is evaluated as
// sqrt(c) |d|.  Each changes an exponent by at most ~2 ulp, i.e. a covariance
// entry by |exponent| * 2^-52 relative -- the same order as exp()'s own rounding.
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
  KernelConsts c{1.0, 0.0, 0.0, 0.0, 0.0};
  if (KIND == GSN_K_EXPSQUARED) {         // kernels.py:55   A**2 * exp(-d**2/(2*tau**2))
    c.amp = kp[0] * kp[0];
    c.c1 = 1.0 / (2.0 * (kp[1] * kp[1]));
  } else if (KIND == GSN_K_EXPONENTIAL) { // kernels.py:109-110   A**2 * exp(-sqrt(d**2)/tau)
    c.amp = kp[0] * kp[0];
    c.c1 = 1.0 / kp[1];
  } else if (KIND == GSN_K_CONSTANT) {    // kernels.py:157
    c.amp = kp[0];
  } else if (KIND == GSN_K_MATERN32) {    // kernels.py:250-251   A (1 + s) exp(-s), s = sqrt(3 d^2)/l
    c.amp = kp[0];
    c.c1 = sqrt(3.0) / kp[1];
  } else if (KIND == GSN_K_MATERN52) {    // kernels.py:305-308   s = sqrt(5 d^2)/l, + 5 d^2/(3 l^2)
    c.amp = kp[0];
    c.c1 = sqrt(5.0) / kp[1];
    c.c2 = 5.0 / (3.0 * (kp[1] * kp[1]));
  } else if (KIND == GSN_K_OU) {          // kernels.py:498-499   A exp(-sqrt(d^2)/l)
    c.amp = kp[0];
    c.c1 = 1.0 / kp[1];
  } else if (KIND == GSN_K_RATQUAD) {     // kernels.py:366   A**2 * (1 + d**2/(2*alpha*tau**2))
    c.amp = kp[0] * kp[0];
    c.c1 = 1.0 / (2.0 * kp[2] * (kp[1] * kp[1]));
  } else if (KIND == GSN_K_GIBBS) {       // kernels.py:438-447
    c.amp = kp[0] * kp[0];
    c.c1 = kp[1];                          // lambda
    c.c2 = kp[2];                          // p
    c.c3 = kp[3];                          // mu
    c.c4 = kp[4];                          // sigma
  } else if (KIND == GSN_K_PERIODIC) {    // kernels.py:523-526
    c.amp = kp[0];
    c.c1 = 1.0 / kp[1];                    // 1 / l
    c.c2 = 1.0 / kp[2];                    // 1 / p
  } else if (KIND == GSN_K_JJ) {          // kernels.py:552-558
    c.amp = kp[0];
    c.c1 = 1.0 / kp[1];                    // 1 / l
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
  if (KIND == GSN_K_JJ) return 1.0 / (c.c2 * x + c.c3);  // 1 / _p(x), kernels.py:549-550
  return 0.0;
}

// k(x_i, x_j) as the reference's _covariance writes element [i, j].  For the JJ
// kernel the period comes from the column epoch j (kernels.py:555 broadcasting),
// so the value is not symmetric in (i, j).  `tab` is the 2^(j/64) table of
// fast_exp (shared memory in the hot kernel).
// cov_unit is the covariance without its leading amplitude factor (c.amp), which the
// hot kernel folds into a per-column coefficient.  The arguments handed to fast_exp
// are finite-or-infinite but never NaN as long as the kernel constants and the
// epochs are finite, which the hot kernel checks once per theta.
template <int KIND, bool kBranch = false>
__device__ __forceinline__ double cov_unit(const KernelConsts& c, double xi, double xj, double auxi, double auxj,
                                           const double* __restrict__ tab) {
  const double d = xi - xj;
  if (KIND == GSN_K_EXPSQUARED) {
    return fast_exp<kBranch>(-(d * d) * c.c1, tab);
  } else if (KIND == GSN_K_EXPONENTIAL || KIND == GSN_K_OU) {
    return fast_exp<kBranch>(-fabs(d) * c.c1, tab);
  } else if (KIND == GSN_K_CONSTANT) {
    return 1.0;
  } else if (KIND == GSN_K_DOTPRODUCT) {   // kernels.py:196
    return xi * xj;
  } else if (KIND == GSN_K_MATERN32) {
    const double s = fabs(d) * c.c1;
    return (1.0 + s) * fast_exp<kBranch>(-s, tab);
  } else if (KIND == GSN_K_MATERN52) {
    const double s = fabs(d) * c.c1;
    const double a = 1.0 + s + ((d * d) * c.c2);
    return a * fast_exp<kBranch>(-s, tab);
  } else if (KIND == GSN_K_RATQUAD) {
    return 1.0 + (d * d) * c.c1;
  } else if (KIND == GSN_K_GIBBS) {
    const double root_num = 2.0 * auxi * auxj;
    const double inv = 1.0 / ((auxi * auxi) + (auxj * auxj));
    return sqrt(root_num * inv) * fast_exp_checked<kBranch>(-(d * d) * inv, tab);
  } else if (KIND == GSN_K_PERIODIC) {
    const double s = sin(M_PI * (fabs(d) * c.c2)) * c.c1;
    return fast_exp_checked<kBranch>(-2.0 * (s * s), tab);
  } else if (KIND == GSN_K_JJ) {
    const double s = sin(M_PI * (fabs(d) * auxj)) * c.c1;
    return fast_exp_checked<kBranch>(-2.0 * (s * s), tab);
  }
  return 0.0;
}

// (cov + cov^T)/2 without the amplitude: the symmetrisation jnp.linalg.cholesky applies
// (gausSN.py:150).  Only JJ is asymmetric.
template <int KIND, bool kBranch = false>
__device__ __forceinline__ double cov_unit_sym(const KernelConsts& c, double xi, double xj, double auxi, double auxj,
                                               const double* __restrict__ tab) {
  if (KIND == GSN_K_JJ) {
    return 0.5 * (cov_unit<KIND, kBranch>(c, xi, xj, auxi, auxj, tab) + cov_unit<KIND, kBranch>(c, xj, xi, auxj, auxi, tab));
  }
  return cov_unit<KIND, kBranch>(c, xi, xj, auxi, auxj, tab);
}

// Full covariance element (single-call API path; NaN-safe).
template <int KIND>
__device__ __forceinline__ double cov_elem(const KernelConsts& c, double xi, double xj, double auxi, double auxj,
                                           const double* __restrict__ tab) {
  const double u = cov_unit<KIND>(c, xi, xj, auxi, auxj, tab);
  const double d = xi - xj;
  // the branch-free exp core does not propagate NaN: restore it for arbitrary user input
  if (d != d || c.c1 != c.c1 || c.c2 != c.c2) return d + c.c1 + c.c2;
  return c.amp * u;
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
