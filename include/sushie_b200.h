/*
 * sushie_b200.h -- C ABI of the B200-native SuShiE IBSS engine.
 *
 * This is the drop-in boundary for the hot path of mancusolab/sushie: the whole
 * iterative-Bayesian-stepwise-selection loop that the reference runs as
 *     for o_iter in range(max_iter): _update_effects(...)        sushie/infer.py:498-548
 *     for o_iter in range(max_iter): _update_effects_ss(...)     sushie/infer_ss.py:348-395
 * (jitted bodies at sushie/infer.py:590-635 and sushie/infer_ss.py:437-478) is moved
 * behind these entry points and runs on the GPU with convergence decided on device.
 * The reference has no FFI of its own (it is pure Python on JAX), so the symbols below
 * are what a ctypes binding inside sushie/infer.py would load; INTEGRATION.md shows it.
 *
 * Conventions
 *   - plain C, no C++ or torch types; every pointer is a HOST pointer unless named d_*;
 *   - caller owns all host buffers; the library owns all device memory (inside sb_ctx /
 *     sb_batch); no host pointer is retained after a call returns, except that a batch
 *     created with sb_batch_create_* has already copied everything it needs;
 *   - every function returns SB_OK (0) or a negative SB_E* code; sb_last_error() gives
 *     the message; per-problem numerical trouble (NaN ELBO => the reference's
 *     "revert to previous iteration and stop") is reported through the result fields
 *     elbo_increase / status, not as a call failure;
 *   - one sb_ctx per (host thread, GPU); calls on one ctx must be serialised by the
 *     caller; different contexts are independent;
 *   - there is no CPU fallback: without a usable CUDA device sb_init fails.
 *
 * Array layouts are the reference's (C order):
 *   alpha[L][p], post_mean[L][p][k], post_mean_sq[L][p][k][k], weighted_sum_covar[L][k][k],
 *   kl[L], log_bf[L][p], effect_covar[L][k][k], resid_var[k], elbo[max_iter+1]
 * (sushie/infer.py:30-64 and 453-475).
 */
#ifndef SUSHIE_B200_H
#define SUSHIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_VERSION 110 /* 0.1.1 */
#define SB_MAX_ANCESTRIES 6
#define SB_MAX_COVAR_BASIS 8 /* intercept + 7 covariates */

/* return codes */
#define SB_OK 0
#define SB_EINVAL (-1)   /* bad argument / unsupported shape */
#define SB_ECUDA (-2)    /* CUDA runtime error */
#define SB_ENOMEM (-3)   /* device or host allocation failed */
#define SB_ENODEV (-4)   /* no usable CUDA device */

/* sb_result.status */
#define SB_STATUS_OK 0
#define SB_STATUS_NONFINITE 1 /* the ELBO became NaN (posterior precision not positive definite, non-finite input):
                                 the reference's "ELBO decreased" branch ran -- previous iteration returned,
                                 elbo_increase == 0 (infer.py:523-533) */

/* element types of genotype / LD inputs and of their device storage */
#define SB_F64 0
#define SB_F32 1
#define SB_I8 2 /* hard-call genotypes (int8). As sb_opts.storage (individual level, every problem x_dtype ==
                  SB_I8, entries in 0..127, p <= 8192): raw bytes stay resident, centring/scaling is folded into the vectors */

#define SB_B2 3 /* hard calls 0/1/2 as bit planes (sb_opts.storage only; individual level, every problem x_dtype ==
                  SB_I8, n_i <= 1024, p <= 8192): the single-effect products become table look-ups (no fp64 widening of
                  genotypes); the one-byte matrix stays resident for the purity kernel */

/* sb_opts.flags */
#define SB_FLAG_SINGLE_PHASE 1 /* keep one team size for the whole batch (no suspend / regroup phases) */

/* sb_ind_problem.standardize bits (device-side restatement of infer.py:326-333) */
#define SB_CENTER 1
#define SB_SCALE 2

typedef struct sb_ctx sb_ctx;
typedef struct sb_batch sb_batch;

/* One individual-level fine-mapping problem = the arguments _update_effects sees after
 * the host prologue of infer_sushie (sushie/infer.py:453-494). */
typedef struct {
  int32_t k;                 /* ancestries, 1..SB_MAX_ANCESTRIES */
  int32_t p;                 /* SNPs */
  int32_t L;                 /* single effects */
  const int32_t *n;          /* [k] samples per ancestry (no zero padding needed) */
  const void *const *X;      /* k pointers, row-major [n_i][p] of x_dtype */
  const double *const *y;    /* k pointers, [n_i], already centred (/scaled) */
  int32_t x_dtype;           /* SB_F64 | SB_F32 | SB_I8 */
  int32_t standardize;       /* 0, or SB_CENTER | SB_SCALE: centre/scale columns on device */
  const double *pi;          /* [p] prior inclusion probabilities, NULL => 1/p */
  const double *resid_var;   /* [k] initial residual variances */
  const double *effect_covar;/* [L][k][k] initial prior effect covariances */
  const double *adj_times;   /* [k][k] _PriorAdjustor.times (used when em_update == 0) */
  const double *adj_plus;    /* [k][k] _PriorAdjustor.plus  (used when em_update == 0) */
  int32_t em_update;         /* 1: _EMOptFunc (infer.py:128-140); 0: _NoopOptFunc (143-159) */
  /* Genotypes residualised on covariates (utils.regress_covar, infer.py:319-335) WITHOUT leaving hard-call storage
   * (sb_opts.storage == SB_B2 only): covar_q[i] = orthonormal basis of [1, covariates_i], row-major
   * [n_i][covar_q_cols[i]] (the Q of the reduced QR the reference's `ols` computes), at most SB_MAX_COVAR_BASIS
   * columns; the engine applies (I - Q Q^T) as a low-rank correction and centring is implied.  NULL: no covariates. */
  const double *const *covar_q;
  const int32_t *covar_q_cols;
  /* Row views (cross-validation folds, cli.py:323-375: every fold is a row subset of one shuffled matrix): when
   * x_rows is not NULL, X[i] points at a SOURCE matrix of x_src_rows[i] rows and this problem's genotypes are its rows
   * x_rows[i][0 .. n[i]) in that order.  Problems of one batch that pass the same X[i] pointer (a gene's full fit and
   * its folds) share ONE upload of the source; the rows are gathered on the device.  NULL: X[i] has n[i] rows. */
  const int32_t *const *x_rows;
  const int32_t *x_src_rows;
} sb_ind_problem;

/* One summary-level problem = the arguments of _update_effects_ss (infer_ss.py:437-446);
 * XtX = ns * LD and Xty = sqrt(ns) * sqrt(ns / (ns + z^2)) * z are formed on device
 * (infer_ss.py:340-344). */
typedef struct {
  int32_t k, p, L;
  const double *ns;          /* [k] sample sizes */
  const void *const *ld;     /* k pointers, row-major [p][p] of ld_dtype */
  const double *const *z;    /* k pointers, [p] z-scores */
  int32_t ld_dtype;          /* SB_F64 | SB_F32 */
  const double *pi;
  const double *resid_var;
  const double *effect_covar;
  const double *adj_times;
  const double *adj_plus;
  int32_t em_update;
} sb_ss_problem;

typedef struct {
  int32_t max_iter;          /* reference default 500 */
  double min_tol;            /* reference default 1e-4 (CLI 1e-3) */
  int32_t storage;           /* device storage of X / LD: SB_F64 (parity mode), SB_F32, SB_I8 or SB_B2 (see above) */
  int32_t team_size;         /* thread blocks cooperating on one locus in the first phase; 0 = choose */
  int32_t flags;             /* SB_FLAG_* */
  int32_t reserve_sms;       /* SMs the IBSS launches leave free (0 = use the whole device).  The persistent kernel
                                otherwise holds every SM for seconds; with several host threads feeding one GPU, the
                                small kernels of the other threads (column statistics of the next chunk, set
                                selection and purity of the previous one) would queue behind it */
  int32_t reserved[2];
} sb_opts;

/* Caller-allocated result of one problem.  Any pointer may be NULL to skip that field. */
typedef struct {
  double *alpha;
  double *post_mean;
  double *post_mean_sq;
  double *weighted_sum_covar;
  double *kl;
  double *log_bf;
  double *effect_covar;      /* final prior effect covariances [L][k][k] */
  double *resid_var;         /* final residual variances [k] */
  double *elbo;              /* [max_iter + 1], elbo[0] = -inf, entries past n_iter untouched */
  int32_t n_iter;            /* outer iterations executed */
  int32_t elbo_increase;     /* 0 => ELBO decreased, previous iteration returned */
  int32_t status;            /* SB_STATUS_* */
  int32_t reserved;
  /* in: NULL, or a permutation of 0..L-1: effect l of every array above is the engine's effect effect_order[l]
   * (the reference's _reorder_l, infer.py:811-832, applied while the download is copied out) */
  const int32_t *effect_order;
} sb_result;

/* Credible-set selection of one problem (reference make_cs, infer.py:899-935), computed on the device per effect l
 * (engine order, before any re-ordering): SNPs by descending alpha (ties: ascending SNP index), the running sum of
 * the sorted alphas accumulated left to right as pandas' cumsum does, and the number of SNPs of the set
 * (first position whose running sum reaches the threshold, inclusive).  Caller-allocated; any pointer may be NULL. */
typedef struct {
  int32_t *order;            /* [L][p] SNP indices, descending alpha */
  double *csum;              /* [L][p] running sum of alpha in that order */
  int32_t *sizes;            /* [L] SNPs in the credible set */
} sb_selection;

typedef struct {
  double device_ms;          /* CUDA-event time of the IBSS kernel(s) on the ctx stream */
  double prologue_ms;        /* device-side standardisation / XtX / Xty kernels */
  int64_t n_ser;             /* single-effect updates executed, summed over problems */
  int64_t n_iter;            /* outer iterations executed, summed over problems */
  double algorithmic_bytes;  /* n_ser-weighted one-traversal bytes */
  int32_t n_launches;        /* kernels launched by the last create+run */
  int32_t team_size;         /* team size actually used */
  int32_t n_teams;
  int32_t rows_per_band;
  int32_t n_stages;
  int32_t smem_bytes;
  int32_t n_run_launches;    /* IBSS kernel launches (phases of the team-size ladder) of the last sb_batch_run */
} sb_run_stats;

int sb_version(void);
int sb_device_count(void);
int sb_init(int device, sb_ctx **out);
int sb_destroy(sb_ctx *ctx);
const char *sb_last_error(const sb_ctx *ctx); /* ctx may be NULL: last init error */
/* Use an existing CUDA stream (a cudaStream_t cast to void*) instead of the ctx's own. */
int sb_set_stream(sb_ctx *ctx, void *cuda_stream);

/* One-shot entry points: upload, run to convergence, download.  These are what replaces the
 * reference's host loops (infer.py:496-548 / infer_ss.py:346-395). */
int sb_ibss_ind(sb_ctx *ctx, const sb_ind_problem *probs, int n_probs, const sb_opts *opts,
                sb_result *results, sb_run_stats *stats);
int sb_ibss_ss(sb_ctx *ctx, const sb_ss_problem *probs, int n_probs, const sb_opts *opts,
               sb_result *results, sb_run_stats *stats);

/* Device-resident batches: upload once, run (repeatedly), fetch. */
int sb_batch_create_ind(sb_ctx *ctx, const sb_ind_problem *probs, int n_probs,
                        const sb_opts *opts, sb_batch **out);
int sb_batch_create_ss(sb_ctx *ctx, const sb_ss_problem *probs, int n_probs,
                       const sb_opts *opts, sb_batch **out);
int sb_batch_run(sb_ctx *ctx, sb_batch *batch, sb_run_stats *stats);
int sb_batch_fetch(sb_ctx *ctx, sb_batch *batch, sb_result *results);
int sb_batch_destroy(sb_ctx *ctx, sb_batch *batch);
/* Device pointer to the stored (centred/scaled, padded) genotypes or LD of one problem:
 * row-major [n_i][ld_elems]; used by the purity kernels and by tests. */
int sb_batch_storage(sb_batch *batch, int prob, int ancestry, void **d_ptr, int64_t *ld_elems);

/* Credible-set purity on device (infer.py:936-965): for each of n_sets index lists
 * (idx + offsets[s] .. idx + offsets[s+1]) return min |corr| per ancestry, out[n_sets][k].
 * Individual level: corr = X_sel^T X_sel / n_k on the stored genotypes; summary level:
 * LD sub-matrix. */
int sb_batch_purity(sb_ctx *ctx, sb_batch *batch, int prob, const int32_t *idx,
                    const int32_t *offsets, int n_sets, double *out_min_abs);
/* Device-side credible-set selection for every problem of a batch that has run: one block per (problem, effect)
 * sorts alpha in shared memory and scans it; thresholds[n_probs] (make_cs's `threshold` of every problem),
 * out[n_probs]. */
int sb_batch_select(sb_ctx *ctx, sb_batch *batch, const double *thresholds, sb_selection *out);
/* The same for the sets of many problems of the batch in one launch (one call per chunk of loci instead of
 * one per locus): set s belongs to problem set_prob[s] and lists idx[offsets[s] .. offsets[s+1]);
 * out[n_sets][SB_MAX_ANCESTRIES], entries of ancestries a problem does not have are 0. */
int sb_batch_purity_many(sb_ctx *ctx, sb_batch *batch, int n_sets, const int32_t *set_prob,
                         const int32_t *idx, const int32_t *offsets, double *out_min_abs);

#ifdef __cplusplus
}
#endif
#endif /* SUSHIE_B200_H */
