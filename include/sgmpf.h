/*
 * sgmpf.h -- C-ABI of the B200 (sm_100a) buffered particle-filter gradient library.
 *
 * This is the drop-in boundary for ONE hot path of the reference `sgmcmc_ssm` package: the batched
 * time loop of the buffered particle filter / smoother behind
 *     sampler.noisy_gradient(kind='pf', ...) -> Helper.pf_gradient_estimate(...)
 *     -> particle_filters.buffered_smoother.buffered_pf_wrapper(...)
 * (reference files, relative to sgmcmc_ssm/:
 *     particle_filters/buffered_smoother.py:12-199   time loop, dispatch on `pf`, average_statistic
 *     particle_filters/pf.py:7-377                   pf step, nemeth / poyiadjis / paris / filter
 *     particle_filters/kernels.py:82-138             LatentGaussianKernel
 *     models/{svm,lgssm,garch}/kernels.py            proposal + reweight
 *     models/{svm,lgssm,garch}/helper.py             complete-data score functions)
 *
 * The reference is pure Python/numpy and has no FFI of its own; the binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md.  One call to sgm_pf_run() executes the whole t-loop for a
 * batch of independent work items (chain x sequence x subsequence) on the given CUDA stream.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all pointers are DEVICE pointers unless marked HOST
 *   - the library allocates nothing and keeps no global mutable state except a thread-local error
 *     string; every buffer (incl. workspace) is caller-owned
 *   - calls are asynchronous on `stream`; there is no hidden synchronisation; the per-item `status`
 *     flags are read by the caller after its own sync
 *   - return value: 0 on success, negative SGM_ERR_* otherwise (never throws, never aborts)
 */
#ifndef SGMPF_H_
#define SGMPF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGM_VERSION 200 /* 0.2.0 */

/* model: models/{lgssm,svm,garch} */
enum { SGM_MODEL_LGSSM = 0, SGM_MODEL_SVM = 1, SGM_MODEL_GARCH = 2 };
/* proposal kernel: Helper._get_kernel (svm/helper.py:56-65, garch/helper.py:48-57, lgssm/helper.py:1200-1214) */
enum { SGM_KERNEL_PRIOR = 0, SGM_KERNEL_OPTIMAL = 1 };
/* smoother: buffered_pf_wrapper `pf` string (buffered_smoother.py:156-199).
 * "poyiadjis_N" is SGM_PF_NEMETH with lambduh = 1.0 (buffered_smoother.py:175-180). */
enum { SGM_PF_NEMETH = 0, SGM_PF_POY_N2 = 1, SGM_PF_PARIS = 2, SGM_PF_FILTER = 3 };
/* arithmetic / storage type of the particle arrays */
enum { SGM_F32 = 0, SGM_F64 = 1 };
/* random numbers: device Philox4x32-7, or arrays recorded from the reference's numpy stream */
enum { SGM_RNG_PHILOX = 0, SGM_RNG_INJECTED = 1 };
/* resampling scheme.  MULTINOMIAL = reference semantics (pf.py:27-29; iid uniforms, child i gets
 * searchsorted(cdf, u_i, 'right')).  MULTINOMIAL_SORTED draws the *order statistics* of N iid
 * uniforms directly (same law up to a permutation of the children, which every estimator here is
 * invariant to) so ancestor indices are non-decreasing and the gather streams.  SYSTEMATIC and
 * STRATIFIED are extensions (not in the reference).  In INJECTED mode the uniforms are used as given;
 * pass SORTED only if they are ascending. */
enum { SGM_RESAMPLE_MULTINOMIAL = 0, SGM_RESAMPLE_MULTINOMIAL_SORTED = 1,
       SGM_RESAMPLE_SYSTEMATIC = 2, SGM_RESAMPLE_STRATIFIED = 3 };
/* additive statistic carried by the smoother */
/* SGM_STAT_PRED: row width of `grad` / horizon slots of `inj_pred`, and the largest num_steps_ahead (the reference's
 * default is 10, sgmcmc_sampler.py:61; its drivers use 5). */
#define SGM_PRED_SLOTS 16
#define SGM_PRED_MAX_STEPS 14

enum { SGM_STAT_SCORE = 0,     /* complete-data log-likelihood gradient (pf_gradient_estimate)   */
       SGM_STAT_SUFF = 1,      /* [x', x'^2, x x'] (lgssm/svm) or [x', x'^2, x'^4] (garch)        */
       SGM_STAT_NONE = 2,      /* log-likelihood only (pf_loglikelihood_estimate)                 */
       SGM_STAT_PRED = 3 };    /* k-step-ahead predictive log-likelihood, pf = FILTER only
                                * (pf_predictive_loglikelihood_estimate; pf.py:73-76 `logsumexp`)   */

/* O(N^2) smoother back end: AUTO = tensor cores (TF32 mma, FP32 accumulate) for SGM_F32, FP32/FP64 pipe
 * otherwise; FP32_PIPE forces the CUDA-core kernel (any dtype); TENSOR requires SGM_F32. */
enum { SGM_N2_AUTO = 0, SGM_N2_FP32_PIPE = 1, SGM_N2_TENSOR = 2 };

/* Device random variates of an SGM_F64 run (Philox only; ignored for SGM_F32 and INJECTED).  NATIVE: 52-bit uniforms,
 * f64 log / sincospi / sqrt transforms.  F32: the variates of the f32 path (32-bit uniforms, MUFU transforms), widened to
 * f64 -- every operation on the particle system (propagation, weights, CDF, statistics) still runs in f64; only the
 * resolution of the random inputs is 2^-24 instead of 2^-53 (a change of the sampling law far below Monte-Carlo error). */
enum { SGM_VARIATES_NATIVE = 0, SGM_VARIATES_F32 = 1 };

/* Kernel family of the O(N) smoothers.
 *   TILES   warp-tile kernels streaming the particle arrays through HBM / L2 (any N; the throughput path): one header + one step
 *           launch per time step, or -- same device functions, bit-identical results -- ONE launch for the whole time loop when
 *           the batch is small: an item's CTAs as a thread-block cluster with the cluster barrier between the steps
 *           (N <= 32768, all clusters resident), else a cooperative launch with a grid barrier (all CTAs resident: <= 148)
 *   STEPS   TILES restricted to one header + one step launch per time step (what the single launches are checked against)
 *   SMALL   one CTA per item, particle system resident in shared memory, whole time loop in one launch (N <= 2048)
 *   CLUSTER one thread-block cluster (2..8 CTAs) per item, particle system in distributed shared memory, whole time loop
 *           in one launch (256 < N <= 16384 and few enough items that all clusters are resident: items x CTAs <= 148)
 *   AUTO    SMALL while N <= 2048 and the batch is small (N <= 512 or items x N <= 1.2e6), else TILES.  (CLUSTER is never
 *           chosen automatically: measured slower than SMALL up to N = 2048 and than the cooperative TILES launch above.) */
enum { SGM_PATH_AUTO = 0, SGM_PATH_TILES = 1, SGM_PATH_SMALL = 2, SGM_PATH_CLUSTER = 3, SGM_PATH_STEPS = 4 };

/* error codes */
enum { SGM_OK = 0, SGM_ERR_INVALID = -1, SGM_ERR_UNSUPPORTED = -2, SGM_ERR_WORKSPACE = -3,
       SGM_ERR_CUDA = -4, SGM_ERR_DEVICE = -5 };

/* per-item status bits (device int32) */
enum { SGM_STATUS_NAN_WEIGHT = 1,   /* NaN / +inf log-weight: np.random.choice would raise ValueError */
       SGM_STATUS_ZERO_WEIGHT = 2,  /* all weights underflowed                                         */
       SGM_STATUS_AR_OVERFLOW = 4 };/* PaRIS, INJECTED mode: accept-reject hit max_accept_reject with more than
                                     * manual_sample_threshold entries left (pf.py:326: a commented-out diagnostic in the reference).
                                     * With device randoms the exact fallback is part of normal operation: no flag. */

#define SGM_THETA_STRIDE 12
/* theta layout (doubles, values exactly as the reference Parameters object computes them):
 *   SVM   : A, LQinv, Qinv, LRinv, Rinv                 [10] = Q, [11] = R (SGM_STAT_PRED only)
 *   LGSSM : A, LQinv, Qinv, C, LRinv, Rinv              [10] = Q, [11] = R (SGM_STAT_PRED only)
 *   GARCH : alpha, beta, gamma, mu, phi, lambduh, LRinv, Rinv, R                                    */

typedef struct sgm_pf_desc {
    int32_t struct_bytes;          /* sizeof(sgm_pf_desc), checked */
    int32_t model, kernel, pf, dtype, rng_mode, resample, stat_kind;
    int32_t n_items;               /* B: work items in this batch                                   */
    int32_t n_particles;           /* N                                                             */
    int32_t max_T;                 /* max over items of T_buf                                       */
    int32_t Ntilde;                /* PaRIS backward samples per particle (pf.py:185)               */
    int32_t accept_reject;         /* PaRIS: 1 = accept-reject (default), 0 = naive O(N^2)          */
    int32_t max_accept_reject;     /* <0: int(100*log10(N/10)) (pf.py:284-285) in INJECTED mode, the kernel's own
                                    * proposal budget (512) with device randoms (a cost knob, not part of the law) */
    int32_t manual_sample_threshold; /* <0: int(10*log10(N/10)) (pf.py:286-287); INJECTED mode only */
    int32_t item_id_base;          /* global index of item 0 (keeps Philox streams rank-invariant)  */
    int32_t n2_mode;               /* O(N^2) smoother: SGM_N2_AUTO / SGM_N2_FP32_PIPE / SGM_N2_TENSOR */
    int32_t pred_steps_ahead;      /* SGM_STAT_PRED: num_steps_ahead K (0..SGM_PRED_MAX_STEPS); K + 1 entries */
    int32_t pred_per_horizon;      /* SGM_STAT_PRED: 0 = the reference's log-sum over ALL horizons (pf.py:73-76),
                                    * 1 = one log-sum per horizon                                    */
    int32_t variates;              /* SGM_VARIATES_*: precision the device random variates are generated in        */
    int32_t path;                  /* SGM_PATH_*: kernel family for the O(N) smoothers (AUTO picks by N)            */
    int32_t reserved1;             /* must be 0                                                     */
    double lambduh;                /* Nemeth shrinkage (pf.py:140); 1.0 = Poyiadjis O(N)            */
    uint64_t seed, offset;         /* Philox key / call counter                                     */
    const uint64_t* offset_dev;    /* optional DEVICE pointer: when set the call counter is read from there by the
                                    * kernels and `offset` is ignored -- lets a captured CUDA graph of this call be
                                    * replayed with fresh random numbers                              */

    /* per-item inputs */
    const double* obs;             /* flat observations                                             */
    const int64_t* obs_off;        /* [B] start of the item's buffered window in obs                */
    const int32_t* T_buf;          /* [B] window length (left buffer + subsequence + right buffer)  */
    const int32_t* t1;             /* [B] relative subsequence start  (buffered_smoother.py:96)     */
    const int32_t* tL;             /* [B] relative subsequence end (exclusive)                      */
    const double* step_weights;    /* flat per-step weights, or NULL                                */
    const int64_t* wts_off;        /* [B] offset into step_weights, -1 = all ones; may be NULL      */
    const double* theta;           /* [B][SGM_THETA_STRIDE]                                         */
    const double* prior_mean;      /* [B]                                                           */
    const double* prior_var;       /* [B]                                                           */

    /* INJECTED randoms (float64, reference consumption order) */
    const double* inj_z0;          /* [B][N]          sample_x0 normals                             */
    const double* inj_u;           /* [B][max_T][N]   resampling uniforms                           */
    const double* inj_z;           /* [B][max_T][N]   proposal normals                              */
    const double* inj_extra;       /* flat PaRIS accept-reject / exact-sampling uniforms            */
    const int64_t* inj_extra_off;  /* [B][max_T]      start of (item, step)'s slice of inj_extra    */
    const double* inj_pred;        /* [B][max_T][SGM_PRED_SLOTS][N] SGM_STAT_PRED: normals of the predictive statistic
                                    * (svm/helper.py:379, garch/kernels.py:66), horizon-major        */

    /* outputs */
    double* grad;                  /* [B][8]  final weighted-average statistic (average_statistic) in slots 0..p-1; PaRIS with
                                    * device randoms also reports in slots 6, 7 the accept-reject proposals made and the entries
                                    * resolved by the exact sampler (sums over the item's time steps);
                                    * SGM_STAT_PRED: [B][SGM_PRED_SLOTS], entries 0..K               */
    double* loglik;                /* [B]     log-likelihood estimate over [t1, tL)                 */
    int32_t* status;               /* [B]                                                           */
    void* out_x;                   /* optional [B][N][n]  final particles (dtype)                   */
    void* out_lw;                  /* optional [B][N]     final log-weights (dtype)                 */
    void* out_stats;               /* optional [B][N][p]  final statistics (dtype)                  */
    int32_t* trace_anc;            /* optional [B][max_T][N]       ancestor indices per step        */
    void* trace_x;                 /* optional [B][max_T+1][N][n]  particles per step (dtype)       */
    void* trace_lw;                /* optional [B][max_T+1][N]     log-weights per step (dtype)     */
    int32_t* trace_J;              /* optional [B][max_T][N][Ntilde] PaRIS backward indices         */

    void* workspace;               /* >= sgm_pf_workspace_bytes(desc), 256-byte aligned             */
    uint64_t workspace_bytes;
    /* optional two-stream pipelining of the O(N) smoothers: the batch is split in two halves that alternate on
     * `stream` and `aux_stream`, so one half's (latency-bound) per-step header and launch gaps hide behind the
     * other half's step kernel.  All three handles are caller-owned; leave NULL to run on `stream` only. */
    void* aux_stream;              /* cudaStream_t                                                   */
    void* ev_aux_fork;             /* cudaEvent_t (timing disabled is fine)                          */
    void* ev_aux_join;             /* cudaEvent_t                                                    */
    void* ev_steps_begin;          /* optional cudaEvent_t recorded on `stream` before the first ...  */
    void* ev_steps_end;            /* ... and after the last step-kernel launch (bench roofline timing) */
} sgm_pf_desc;

/* library version (SGM_VERSION) */
int sgm_version(void);
/* 0 iff the current CUDA device has compute capability 10.x (B200); SGM_ERR_DEVICE otherwise */
int sgm_device_check(void);
/* thread-local description of the last error returned on this thread */
const char* sgm_last_error(void);
/* statistic width p and latent width n for (model, stat_kind); negative on error */
int sgm_stat_dim(int32_t model, int32_t stat_kind);
int sgm_state_dim(int32_t model);
/* bytes of workspace sgm_pf_run needs for this descriptor (0 on invalid descriptor) */
uint64_t sgm_pf_workspace_bytes(const sgm_pf_desc* d);
/* run the whole buffered t-loop for the batch on `stream` (a cudaStream_t passed as void*) */
int sgm_pf_run(const sgm_pf_desc* d, void* stream);
/* IMQ kernel Stein discrepancy of a trace (sgmcmc_ssm/trace_metric_functions.py:20-81): x, gradlogp are DEVICE
 * [num_points][dim] float64 (dim <= 8); writes ceil(num_points / 256) partial sums of sum_{i,j} k0(x_i, x_j) to the
 * DEVICE array `partial`; KSD = sqrt(sum(partial)) / num_points. */
int sgm_ksd_imq(const double* x, const double* gradlogp, int32_t num_points, int32_t dim, double c, double beta,
                double* partial, void* stream);
/* self-test hook for the table-driven functions of the f64 variate transforms (csrc/fastlog.cuh), DEVICE arrays:
 *   SGM_SELFTEST_LOG        y[i] = ln(x[i]) for n normal positive doubles                      (<= 4 ulp on (0, 1))
 *   SGM_SELFTEST_SINCOS2PI  (y[2i], y[2i+1]) = (sin, cos)(2 pi v), v = the 60 top bits of the 64-bit PATTERN of x[i]
 *                           read as a binary fraction (k = top 8 bits, f = next 52)            (|error| <= 1e-15)
 * so that the parity tests can pin them against the host's libm */
#define SGM_SELFTEST_LOG 0
#define SGM_SELFTEST_SINCOS2PI 1
int sgm_selftest_math(int32_t what, const double* x, double* y, int64_t n, void* stream);
/* number of kernel launches the last sgm_pf_run / sgm_sgld_run on this thread issued (for bench accounting) */
int64_t sgm_last_launch_count(void);

/* ---- device-resident SG-MCMC iterations -------------------------------------------------------------------------
 * K iterations of C independent chains without a host round trip.  Replaces, for the n = m = 1 models and kind='pf',
 * the reference's per-iteration Python (relative to sgmcmc_ssm/):
 *     sgmcmc_sampler.py:1969-2017  random_subsequence_and_weights        (window draw + importance weights)
 *     sgmcmc_sampler.py:259-288    _random_subsequence_and_buffers       (buffers)
 *     sgmcmc_sampler.py:390-464    _noisy_grad_loglikelihood / noisy_gradient (minibatch mean, + grad log-prior, / T)
 *     sgmcmc_sampler.py:1249-1283  SeqSGMCMCSampler._noisy_grad_loglikelihood (sequence choice, T / S rescale)
 *     sgmcmc_sampler.py:529-567    _get_sgmcmc_noise, sample_sgld;  :613-640 sample_sgrld;  :466-480 step_sgd
 *     sgmcmc_sampler.py:650-656    project_parameters  (model defaults: |A| <= 0.9999, Cholesky diagonals > 0, LGSSM C = 1)
 *     variables/covariance.py:229-243, variables/matrices.py:575-590, variables/garch_var.py:150-163   grad_logprior
 * Every iteration = sgld_prepare_kernel -> the sgm_pf_run launch sequence for the C x sequences x minibatch work
 * items -> sgld_update_kernel, all enqueued on `stream`; the Philox call counter lives in device memory and is advanced
 * on the device, so a captured CUDA graph of this call replays with fresh random numbers. */
enum { SGM_STEP_SGLD = 0, SGM_STEP_SGRLD = 1 /* LGSSM only (LGSSMPreconditioner) */, SGM_STEP_SGD = 2 };
enum { SGM_PARTITION_UNIFORM = 0, SGM_PARTITION_NAIVE = 1, SGM_PARTITION_STRICT = 2 };
#define SGM_PARAM_STRIDE 8
#define SGM_HYPER_STRIDE 16
/* parameter slots (var_dict order) and prior hyper-parameter slots per model:
 *   SVM   : [A, LQinv, LRinv]                          hyper [mean_A, var_col_A, df_Qinv, scale_Qinv, df_Rinv, scale_Rinv]
 *   LGSSM : [A, C, LQinv, LRinv]                       hyper [mean_A, var_col_A, mean_C, var_col_C, df_Qinv, scale_Qinv, df_Rinv, scale_Rinv]
 *   GARCH : [log_mu, logit_phi, logit_lambduh, LRinv]  hyper [scale_mu, shape_mu, alpha_phi, beta_phi, alpha_lambduh, beta_lambduh, df_Rinv, scale_Rinv] */

typedef struct sgm_sgld_desc {
    int32_t struct_bytes;          /* sizeof(sgm_sgld_desc), checked */
    int32_t method;                /* SGM_STEP_*                                                     */
    int32_t n_chains;              /* C chains advanced in lock-step                                 */
    int32_t minibatch;             /* M subsequences per (chain, sequence) and iteration             */
    int32_t n_seqs;                /* observation sequences (1 for a plain sampler)                  */
    int32_t num_sequences;         /* Seq samplers: sequences drawn per iteration without replacement (<= 8), -1 = all */
    int32_t subsequence_length;    /* S, -1 = whole sequence                                         */
    int32_t buffer_length;         /* B, -1 = whole sequence                                         */
    int32_t partition;             /* SGM_PARTITION_* (options['partition_style'])                    */
    int32_t n_iters;               /* K iterations in this call                                      */
    int32_t project;               /* 1 = project_parameters after every step                        */
    int32_t prior_x0;              /* 0 = prior_mean / prior_var per chain; 1 = GARCH: alpha / (1 - beta - gamma) of the
                                    * chain's CURRENT parameters (garch/helper.py:324-332, forward_message None) */
    int32_t trace_every;           /* > 0: parameters copied to `trace` every that many iterations   */
    int32_t trace_rows;            /* rows of `trace` (row r = after r * trace_every iterations; row 0 is the caller's) */
    int32_t max_seq_len;           /* longest sequence (bounds T_buf when S or B is -1)              */
    int32_t no_persistent;         /* 1 = always one launch sequence per iteration (default 0: a batch with one work item
                                    * per chain and N <= 2048 runs all K iterations inside ONE persistent kernel)     */
    double epsilon;                /* step size                                                      */
    double T_total;                /* sum of the sequence lengths (the 1 / T scaling of gradient and noise) */
    const double* obs;             /* DEVICE flat observations, sequence after sequence              */
    const int64_t* seq_off;        /* DEVICE [n_seqs + 1] offsets into obs                           */
    double* params;                /* DEVICE [C][SGM_PARAM_STRIDE] in / out                          */
    const double* hyper;           /* DEVICE [C][SGM_HYPER_STRIDE]                                   */
    const double* prior_mean;      /* DEVICE [C] (prior_x0 == 0)                                     */
    const double* prior_var;       /* DEVICE [C]                                                     */
    int32_t* chain_status;         /* DEVICE [C]: SGM_STATUS_* bits; a flagged chain keeps its last finite parameters */
    double* trace;                 /* DEVICE optional [trace_rows][C][SGM_PARAM_STRIDE]              */
    uint64_t* offset_dev;          /* DEVICE Philox call counter (read by every kernel, + 1 per iteration) -- required */
    int64_t* iter_dev;             /* DEVICE iterations completed so far (+ 1 per iteration) -- required */
    /* INJECTED parity mode (pf.rng_mode == SGM_RNG_INJECTED): the draws the reference takes from the numpy stream */
    const int32_t* inj_start;      /* [K][items] np.random.randint window starts (strict: block index) */
    const int32_t* inj_seq;        /* [K][C][num_sequences] sequence picks (num_sequences != -1)      */
    const double* inj_noise;       /* [K][C][SGM_PARAM_STRIDE] standard normals of _get_sgmcmc_noise / precondition_noise */
    void* workspace;               /* >= sgm_sgld_workspace_bytes(desc), 256-byte aligned            */
    uint64_t workspace_bytes;
    /* particle-filter template: model, kernel, pf, dtype, rng_mode, resample, variates, n_particles, lambduh, Ntilde, ...,
     * seed, item_id_base (global index of chain 0 x items per chain), aux stream / events.  stat_kind must be
     * SGM_STAT_SCORE.  n_items, max_T, the per-item arrays, outputs, workspace and offset_dev are filled in by the
     * library.  INJECTED: inj_z0 / inj_u / inj_z carry a leading [K] dimension ([K][items][N], [K][items][max_T][N]). */
    sgm_pf_desc pf;
} sgm_sgld_desc;

/* items per iteration (C x sequences x M) and the max_T the library uses for this descriptor; negative on error */
int32_t sgm_sgld_items(const sgm_sgld_desc* d);
int32_t sgm_sgld_max_steps(const sgm_sgld_desc* d);
uint64_t sgm_sgld_workspace_bytes(const sgm_sgld_desc* d);
/* enqueue n_iters iterations on `stream`; asynchronous, no hidden synchronisation */
int sgm_sgld_run(const sgm_sgld_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGMPF_H_ */
