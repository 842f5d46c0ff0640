/*
 * gsn.h -- C ABI of the B200-native batched GP log-likelihood engine (libgsn.so).
 *
 * This is the drop-in boundary for the one hot path of erinhay/GausSN: the GP
 * log-likelihood that the nested/MCMC sampler evaluates for many parameter
 * vectors theta.  The reference has no FFI of its own (it is pure Python on
 * JAX); each entry point below names the reference interface it stands in for
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; no torch/CUDA types in the signatures (a CUDA stream
 *     is passed as void*);
 *   - every function returns 0 on success and a negative gsn_status on misuse
 *     or CUDA failure; the message is available from gsn_last_error()
 *     (thread-local);
 *   - a numerical failure for one theta (covariance not positive definite, NaN
 *     in the inputs) is a VALUE, never an error code: logl[i] is NaN (or -inf
 *     with GSN_FLAG_NEG_INF, which is what GP.jointprobability returns,
 *     gaussn/gausSN.py:203-204) and info[i] != 0;
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with GSN_ERR_CUDA;
 *   - streams: a handle owns one set of scratch state (work counter, factor
 *     workspace, theta map), so its launches are serialised on the device: every
 *     batch call is ordered behind the handle's previous launch by an event,
 *     whichever streams the two were issued on (the *_host entry points use a
 *     stream of the handle's own).  gsn_problem_set_data / _set_theta_map wait for
 *     the last launch before they touch device state.  Calls on ONE handle must
 *     not be made from two host threads at once; different handles are independent.
 */
#ifndef GSN_H_
#define GSN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSN_ABI_VERSION 1

typedef struct gsn_problem gsn_problem;

typedef enum gsn_status {
  GSN_OK = 0,
  GSN_ERR_INVALID = -1,   /* bad argument (null pointer, size, unknown kind, unsorted indices) */
  GSN_ERR_CUDA = -2,      /* CUDA runtime failure (no device, out of memory, launch failure) */
  GSN_ERR_UNSUPPORTED = -3
} gsn_status;

/* Covariance functions: gaussn/kernels.py.  theta slots in constructor order. */
typedef enum gsn_kernel_kind {
  GSN_K_EXPSQUARED = 0,   /* kernels.py:4-56     [A, tau]                  A^2 exp(-d^2 / (2 tau^2)) */
  GSN_K_EXPONENTIAL = 1,  /* kernels.py:58-111   [A, tau]                  A^2 exp(-|d| / tau) */
  GSN_K_CONSTANT = 2,     /* kernels.py:113-158  [c]                       c */
  GSN_K_DOTPRODUCT = 3,   /* kernels.py:160-197  []                        x_i x_j */
  GSN_K_MATERN32 = 4,     /* kernels.py:199-252  [A, l]                    A (1 + s) exp(-s), s = sqrt(3 d^2)/l */
  GSN_K_MATERN52 = 5,     /* kernels.py:254-309  [A, l]                    A (1 + s + 5 d^2/(3 l^2)) exp(-s), s = sqrt(5 d^2)/l */
  GSN_K_RATQUAD = 6,      /* kernels.py:311-367  [A, tau, alpha]           A^2 (1 + d^2/(2 alpha tau^2))   (as written) */
  GSN_K_GIBBS = 7,        /* kernels.py:369-448  [A, lambda, p, mu, sigma] */
  GSN_K_OU = 8,           /* kernels.py:450-500  [A, l]                    A exp(-|d| / l) */
  GSN_K_PERIODIC = 9,     /* kernels.py:502-527  [A, l, p]                 A exp(-2 (sin(pi |d| / p) / l)^2) */
  GSN_K_JJ = 10,          /* kernels.py:529-559  [A, l, m, b]              period m x_j + b from the column epoch; symmetrised */
  GSN_K_COUNT = 11
} gsn_kernel_kind;

/* Mean functions: gaussn/meanfuncs.py. */
typedef enum gsn_mean_kind {
  GSN_M_UNIFORM = 0,      /* meanfuncs.py:4-45     [c] */
  GSN_M_SIN = 1,          /* meanfuncs.py:123-172  [A, w, phi] */
  GSN_M_GAUSSIAN = 2,     /* meanfuncs.py:174-224  [A, mu, sigma] */
  GSN_M_EXPFUNCTION = 3,  /* meanfuncs.py:226-273  [A, tau] */
  GSN_M_BAZIN2009 = 4,    /* meanfuncs.py:275-337  [A, beta, t0, Tfall, Trise] */
  GSN_M_KARPENKA2012 = 5, /* meanfuncs.py:339-408  [lnA, lnB, t1, t0, Tfall, Trise, (ignored ...)] */
  GSN_M_VILLAR2019 = 6,   /* meanfuncs.py:410-484  [A, beta, t1, t0, Tfall, Trise] */
  GSN_M_HOST = 7,         /* mean supplied by the caller per theta (sncosmoMean, meanfuncs.py:47-121): see gsn_logl_batch_hostmean */
  GSN_M_COUNT = 8
} gsn_mean_kind;

/* Lensing models: gaussn/lensingmodels.py.  Parameters are interleaved per extra
 * image (image 1 is the reference image with delta = 0, beta = 1). */
typedef enum gsn_lens_kind {
  GSN_L_NONE = 0,         /* lensingmodels.py:5-14     no params; no band mask (mask = 1) */
  GSN_L_CONSTANT = 1,     /* lensingmodels.py:16-118   [delta, beta] x (n_images - 1) */
  GSN_L_SIGMOID = 2,      /* lensingmodels.py:120-241  [delta, beta0, beta1, r, t0] x (n_images - 1) */
  GSN_L_SINUSOIDAL = 3,   /* lensingmodels.py:243-364  [delta, beta0, beta1, t0, T] x (n_images - 1) */
  GSN_L_COUNT = 4
} gsn_lens_kind;

/* flags for gsn_logl_batch* */
#define GSN_FLAG_NEG_INF 1u  /* map NaN / +-inf results to -inf (GP.jointprobability, gausSN.py:203-204) */

/* gsn_query selectors */
typedef enum gsn_query_what {
  GSN_Q_N = 0,               /* number of epochs */
  GSN_Q_NPARAMS = 1,         /* full theta width: kernel + mean + lensing */
  GSN_Q_NPARAMS_KERNEL = 2,
  GSN_Q_NPARAMS_MEAN = 3,
  GSN_Q_NPARAMS_LENS = 4,
  GSN_Q_WORKSPACE_BYTES = 5, /* device workspace owned by the handle so far (it grows with the largest batch evaluated) */
  GSN_Q_GRID = 6,            /* persistent CTAs launched per batch call */
  GSN_Q_DEVICE = 7,
  GSN_Q_LAUNCHES = 8,        /* kernels launched through this handle so far */
  GSN_Q_NPAD = 9,            /* padded matrix order used on the device */
  GSN_Q_PATH = 10,           /* launch shape of the handle's most recent batch call (gsn_path; 0 before the first) */
  GSN_Q_CLUSTER2_MAX = 11,   /* largest batch that runs one theta per cluster of 2 / 4 / 8 CTAs (0: size not used at this N) */
  GSN_Q_CLUSTER4_MAX = 12,
  GSN_Q_CLUSTER8_MAX = 13,
  GSN_Q_ONCHIP = 14,         /* 1 if the handle's batch calls run with the factor resident in shared memory (GSN_PATH_ONCHIP) */
  GSN_Q_ONCHIP_WARPS = 15,   /* warps that shared one theta in the most recent such launch */
  GSN_Q_ONCHIP_SMEM = 16,    /* bytes of shared memory one theta occupies in that shape */
  GSN_Q_ONCHIP_REG = 17      /* 1 if the most recent such launch kept the factor in registers (band blocks of at most 64 rows, one warp per theta) */
} gsn_query_what;

typedef enum gsn_path {
  GSN_PATH_DMMA_BLOCKED = 1,        /* four warps per theta (band blocks above 336 rows, or smaller batches): blocked left-looking Cholesky, FP64 DMMA updates, L in a global workspace */
  GSN_PATH_DMMA_BLOCKED_NARROW = 2, /* the same kernel, one warp per theta (largest band block <= 336 rows and B >= 512, or >= 1024 above 224 rows) */
  GSN_PATH_ONCHIP = 4,              /* every band block has at most 224 rows: factor resident in shared memory at 8-row granularity, 1-16 warps per theta, no workspace (any batch size) */
  GSN_PATH_DMMA_CLUSTER = 3         /* latency-bound calls (N >= 256, at most as many thetas as clusters fit on the GPU): a thread-block cluster of 2, 4 or 8 four-warp CTAs shares one theta */
} gsn_path;

/*
 * Create a problem: one data set (x, y, yerr, segment layout) and one plugin
 * triple.  Stands in for the state GP.optimize_parameters builds before it
 * starts a sampler: gaussn/gausSN.py:276-291 (arrays to the device,
 * _prepare_indices :74-101, lensingmodel.import_from_gp lensingmodels.py:102-118).
 *
 *   x, y, yerr  host arrays of N doubles; rows must already be sorted by
 *               (band, image) exactly as _prepare_indices assumes.  Copied.
 *   indices     host array of n_bands*n_images + 1 cumulative row counts,
 *               band-major / image-minor, indices[0] == 0, indices[last] == N.
 *   n_mean_params  len(meanfunc.params) (Karpenka2012 may carry extra, unused
 *               entries: meanfuncs.py:357-362); pass -1 for the class default.
 *               For GSN_M_HOST it is the number of theta slots reserved for the
 *               host-evaluated mean function (they are skipped on the device).
 */
int gsn_problem_create(gsn_problem** out, int device, int64_t N,
                       const double* x, const double* y, const double* yerr,
                       int32_t n_bands, int32_t n_images, const int64_t* indices,
                       int32_t kernel_kind, int32_t mean_kind, int32_t lens_kind,
                       int32_t n_mean_params);

int gsn_problem_destroy(gsn_problem* p);

/* Replace the observations of a handle in place (same N, same band / image structure): what the reference does
 * when its plotting loop builds a fresh GP for every posterior sample and band (gaussn/utils.py:193-204) or when
 * optimize_parameters is called again with new data (gausSN.py:276-280) -- without paying for a new handle.
 * Waits for the handle's launches in flight. */
int gsn_problem_set_data(gsn_problem* p, const double* x, const double* y, const double* yerr);

/*
 * fix_* handling of GP.jointprobability (gaussn/gausSN.py:173-194) without
 * host-side theta expansion.  col_of_param[k] is the column of the caller's
 * theta matrix that feeds full-theta slot k (order kernel | mean | lensing), or
 * -1 when slot k is held fixed at fixed_values[k].  Default: identity.
 */
int gsn_problem_set_theta_map(gsn_problem* p, const int32_t* col_of_param, const double* fixed_values);

/*
 * Batched GP.loglikelihood (gaussn/gausSN.py:135-160): logl[i] for each row of
 * theta.  DEVICE pointers, caller-owned; asynchronous on `cuda_stream`.
 *
 *   theta_dev  [B, ld] row-major doubles, ld >= number of mapped columns
 *   logl_dev   [B] doubles
 *   info_dev   [B] int32 or NULL: 0 ok; k > 0 the k-th pivot (1-based) was not
 *              positive; -1 the result is not finite for another reason
 */
int gsn_logl_batch(gsn_problem* p, const double* theta_dev, int64_t B, int64_t ld,
                   double* logl_dev, int32_t* info_dev, uint32_t flags, void* cuda_stream);

/* gsn_logl_batch with the theta-sharded all-gather fused into the kernel: every result is stored at index
 * out_offset + i of EACH of the n_out device buffers (this GPU's gathered buffer and the peer-mapped buffers of
 * the other GPUs of the box, reachable over NVLink), so that after a cross-rank barrier every rank holds the
 * logL of all shards without a separate collective.  out_ptrs is a host array of n_out (<= 16) device pointers.
 * Replaces the (absent) multi-device story of the reference; see gaussn_b200/sharding.py. */
int gsn_logl_batch_gather(gsn_problem* p, const double* theta_dev, int64_t B, int64_t ld,
                          double* const* out_ptrs, int32_t n_out, int64_t out_offset,
                          int32_t* info_dev, uint32_t flags, void* cuda_stream);

/* The same call with HOST buffers: pinned staging, H2D of theta and D2H of the
 * results inside, synchronous.  This is the end-to-end path bench.py times. */
int gsn_logl_batch_host(gsn_problem* p, const double* theta_host, int64_t B, int64_t ld,
                        double* logl_host, int32_t* info_host, uint32_t flags);

/* As gsn_logl_batch for a problem created with GSN_M_HOST (a mean function
 * that only exists on the host, e.g. sncosmoMean, meanfuncs.py:47-121):
 * mean_dev is [B, N] row-major and holds m(shifted_x) for each theta, evaluated
 * by the caller at the shifted epochs gsn_lens_batch returns; the device applies
 * the magnification, mean = b * m (gausSN.py:142), and everything after it. */
int gsn_logl_batch_hostmean(gsn_problem* p, const double* theta_dev, int64_t B, int64_t ld,
                            const double* mean_dev, double* logl_dev, int32_t* info_dev,
                            uint32_t flags, void* cuda_stream);

/* Shifted epochs and magnifications for each theta (what lensingmodel.lens
 * returns, lensingmodels.py:79-100), so that a host mean function can be
 * evaluated at the shifted epochs.  shifted_dev and b_dev are [B, N]. */
int gsn_lens_batch(gsn_problem* p, const double* theta_dev, int64_t B, int64_t ld,
                   double* shifted_dev, double* b_dev, void* cuda_stream);

/* GP.predict for a batch of theta (gaussn/gausSN.py:362-403), applied the way the reference's plotting code applies it
 * (gaussn/utils.py:193-209): per posterior sample, to the de-lensed observations of one band.
 *   rows [row0, row0 + n_rows) of the problem -- one band's rows, or all of them -- are the conditioning epochs: shifted
 *     by the sample's time delays, data and errors divided by its magnifications (utils.py:202-204);
 *   xprime_dev [M]: prediction epochs, in the frame of the reference image (delta = 0, magnification 1);
 *   expectation_dev [B, M]  = m(x') + K(x', x~) C^-1 (y / b - m(x~)),  C = K(x~, x~) + diag((yerr / b)^2)   gausSN.py:400
 *   variance_dev [B, M, M]  = K(x', x') - K(x', x~) C^-1 K(x~, x')     (NULL: expectation only)               gausSN.py:401
 *   mean_v_dev [B, n_rows], mean_u_dev [B, M]: the mean function at the shifted epochs / at x', evaluated by the caller,
 *     for GSN_M_HOST problems; NULL otherwise;
 *   info_dev [B] or NULL: 0 ok; k > 0: C is not positive definite at pivot k (the moments are NaN; the reference's LU
 *     solve would return numbers without meaning); -1: non-finite parameters.
 * No band mask is applied (GP.predict has none).  n_rows rounded up to 32, plus M, must not exceed 8192.  GSN_K_JJ is not
 * supported (the reference's K(x', x) needs len(x') == len(x), kernels.py:555).  theta follows the handle's theta map. */
int gsn_predict_batch(gsn_problem* p, const double* theta_dev, int64_t B, int64_t ld, int64_t row0, int64_t n_rows,
                      const double* xprime_dev, int64_t M, const double* mean_v_dev, const double* mean_u_dev,
                      double* expectation_dev, double* variance_dev, int32_t* info_dev, void* cuda_stream);

/* The same call with HOST buffers (copies inside, synchronous). */
int gsn_predict_batch_host(gsn_problem* p, const double* theta_host, int64_t B, int64_t ld, int64_t row0, int64_t n_rows,
                           const double* xprime_host, int64_t M, const double* mean_v_host, const double* mean_u_host,
                           double* expectation_host, double* variance_host, int32_t* info_host);

int64_t gsn_query(const gsn_problem* p, int32_t what);

/* Measurement aid, not part of the reference's surface: the FP64 tensor-pipe rate of `device` in TFLOP/s (back-to-back
 * independent DMMA.8x8x4 chains on every SM for about `seconds`), i.e. the roofline denominator bench.py reports against. */
int gsn_measure_fp64_peak(int device, double seconds, double* tflops_out);

/* Introspection, no device needed: the order in which the work items (32x32 blocks (c, j), c >= j, of the blocked
 * Cholesky) of one evaluation are handed to warps, as built by gsn_problem_create for a data set of N rows whose
 * band index per row is band_of_row (NULL = one band).  Item word: (first_col << 24) | (c << 12) | j.  Returns the
 * number of items (also when out is NULL or cap is too small), -1 on bad arguments.  Blocks that the band mask
 * (lensingmodels.py:39-51) makes identically zero are not listed. */
int64_t gsn_item_order(int64_t N, const int32_t* band_of_row, int32_t n_bands, int32_t masked, uint32_t* out, int64_t cap);

/* Introspection, no device needed: the tile-slot table of the on-chip shape (gsn_onchip.cuh) for a band block of
 * `tile_rows` 8-row tile rows.  The factor's 8x8 tiles live in recycled shared-memory slots: tile (i, k), k <= i, of the
 * lower triangle is held in slot out[i (i + 1) / 2 + k]; the tiles of row r are dead `lag` tile columns after column r
 * (lag 2 for teams of several warps, 1 for a one-warp team) and their slots are handed to the tiles born then.  Returns the
 * number of slots (also when out is NULL or cap is too small), -1 on bad arguments.  tests/test_host_logic.py replays the
 * schedule against this table: no tile may take a slot whose previous tenant can still be read. */
int64_t gsn_onchip_slot_table(int32_t tile_rows, int32_t lag, uint8_t* out, int64_t cap);

/* The same for gsn_predict_batch with n_rows conditioning epochs and M prediction epochs: items of the augmented matrix
 * (n_rows padded to 32, then the M epochs).  Bit 23 of a word marks a trailing block (both c and j among the prediction
 * rows: accumulated and written out, neither solved nor factored); c = (word >> 12) & 0x7ff, j = word & 0xfff. */
int64_t gsn_predict_item_order(int64_t n_rows, int64_t M, int32_t want_cov, uint32_t* out, int64_t cap);
/* want_cov bit 1 (value 2 or 3) selects the column-major order used when one theta runs on a thread-block cluster. */

/* ---- single-call plugin evaluations (HOST pointers, synchronous) ------------ */

/* kernel.covariance(x, x_prime, params): gaussn/kernels.py (each class's
 * _covariance).  K_out is [n, m] row-major.  x_prime == NULL means x_prime = x.
 * GSN_K_JJ is returned as written (asymmetric) and needs n == m. */
int gsn_covariance(int device, int32_t kernel_kind, const double* params, int32_t n_params,
                   const double* x, int64_t n, const double* x_prime, int64_t m, double* K_out);

/* meanfunc.mean(t, params): gaussn/meanfuncs.py. */
int gsn_mean(int device, int32_t mean_kind, const double* params, int32_t n_params,
             const double* t, int64_t n, double* mean_out);

/* lensingmodel.lens(x, params) after import_from_gp: gaussn/lensingmodels.py.
 * For GSN_L_NONE b_out is filled with 1. */
int gsn_lens(int device, int32_t lens_kind, const double* params, int32_t n_params,
             const double* x, int64_t N, int32_t n_bands, int32_t n_images, const int64_t* indices,
             double* shifted_out, double* b_out);

/* number of theta slots of a plugin (-1 for an unknown kind) */
int32_t gsn_kernel_nparams(int32_t kernel_kind);
int32_t gsn_mean_nparams(int32_t mean_kind);
int32_t gsn_lens_nparams(int32_t lens_kind, int32_t n_images);

int32_t gsn_abi_version(void);
int32_t gsn_device_count(void);   /* CUDA devices visible to this process (0 without a driver) */
const char* gsn_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GSN_H_ */
