/* cocos_b200.h -- C ABI of libcocos_b200.so: the B200 (sm_100a) implementation of
 * the Heston Monte-Carlo hot path of michaelnowotny/cocos.
 *
 * The reference has no FFI of its own: its seam is Python duck typing
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it
 * stands in for (paths relative to the reference tree).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, a cb_status code otherwise; the message
 *    is available from cb_last_error() (thread-local);
 *  - the caller owns every buffer; the library never frees caller memory;
 *  - "d_" pointers are device pointers on the context's device, "h_" pointers are
 *    host pointers; plain C types only (no torch, no C++ types);
 *  - a cb_ctx binds one device, one stream and a small workspace; calls on one
 *    context must come from one thread at a time, different contexts are
 *    independent;
 *  - there is no CPU fallback: without a CUDA device every compute call fails with
 *    CB_ERR_CUDA.
 */
#ifndef COCOS_B200_H
#define COCOS_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CB_API __attribute__((visibility("default")))
#else
#define CB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CB_OK = 0,
    CB_ERR_INVALID = 1,     /* bad argument (Python shim raises ValueError) */
    CB_ERR_CUDA = 2,        /* CUDA runtime failure (RuntimeError) */
    CB_ERR_NCCL = 3,        /* NCCL failure or libnccl not loadable (RuntimeError) */
    CB_ERR_UNSUPPORTED = 4
} cb_status;

typedef struct cb_ctx cb_ctx;       /* one device + one stream + workspace */
typedef struct cb_comm cb_comm;     /* one rank of a communicator: NCCL handle + NVLink peer mappings */

/* Heston model scalars: the arguments of simulate_heston_model
 * (examples/stochastic_volatility/heston_utilities.py:24-35). */
typedef struct {
    double T, mu, kappa, v_bar, sigma_v, rho, x0, v0;
} cb_heston_params;

/* Path-dependent payoffs evaluated inside the fused kernel (the reference only motivates
 * them, README.md:376-382).  kind 0: European (terminal price), 1: arithmetic-average Asian over the N-1 simulated
 * dates, 2: up-and-out, 3: down-and-out (knock-out monitored at the N-1 simulated dates, `barrier` = price level). */
typedef struct {
    int kind;
    int is_put;        /* 0: max(S-K,0), 1: max(K-S,0) */
    double barrier;
} cb_payoff_spec;

#define CB_MAX_STRIKES 64
#define CB_UNIQUE_ID_BYTES 128
#define CB_P2P_HANDLE_BYTES 64

/* ---- library / device management: cocos/device.py:95-205 ---------------- */
CB_API const char *cb_last_error(void);
CB_API const char *cb_version(void);
CB_API int cb_device_count(int *n);                                   /* ComputeDeviceManager.get_compute_devices */
CB_API int cb_device_name(int dev, char *buf, size_t buflen);         /* ComputeDevice.name */
CB_API int cb_device_attributes(int dev, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);
CB_API int cb_ctx_create(int dev, void *stream_or_null, cb_ctx **out); /* set_compute_device; NULL: own stream */
CB_API int cb_ctx_destroy(cb_ctx *ctx);
CB_API int cb_ctx_sync(cb_ctx *ctx);                                  /* cocos.device.sync (device.py:195-205) */
CB_API int cb_ctx_device(cb_ctx *ctx, int *dev);
CB_API void *cb_ctx_stream(cb_ctx *ctx);
/* CUDA-event stopwatch on the context's stream (used by bench.py) */
CB_API int cb_timer_start(cb_ctx *ctx);
CB_API int cb_timer_stop(cb_ctx *ctx, float *elapsed_ms);             /* synchronises the stop event */
/* tuning knobs of the fused kernels: threads per block, blocks per SM, pairs per thread (0 = default) */
CB_API int cb_ctx_set_launch(cb_ctx *ctx, int threads, int blocks_per_sm, int pairs_per_thread);
/* kernel-variant selector used while tuning (0 = default, 1 = no software pipelining, 2 = register-capped); every
 * variant computes the same stream and the same results */
CB_API int cb_ctx_set_variant(cb_ctx *ctx, int variant);

/* caller-side memory helpers for hosts that do not bring their own allocator: the storage behind the device array
 * handle that stands in for cocos.numerics.ndarray (cocos/numerics/_array.py:150-171).  cb_malloc / cb_free are
 * stream-ordered on the context's stream (the device's cudaMallocAsync pool, freed blocks are kept for reuse). */
CB_API int cb_malloc(cb_ctx *ctx, size_t bytes, void **d_ptr);
CB_API int cb_free(cb_ctx *ctx, void *d_ptr);
CB_API int cb_memcpy_h2d(cb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);   /* synchronous */
CB_API int cb_memcpy_d2h(cb_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);   /* synchronous */
CB_API int cb_memcpy_d2d(cb_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);   /* asynchronous, same device */
/* device-to-device across contexts (ComputeDevicePool replicas: cocos/scientific/kde.py:698-767 ships the KDE to
 * every worker); ordered on dst_ctx's stream after src_ctx's stream has drained */
CB_API int cb_memcpy_peer(cb_ctx *dst_ctx, void *d_dst, cb_ctx *src_ctx, const void *d_src, size_t bytes);
/* ndarray.astype (cocos/numerics/_array.py:221-228); kinds: 0 float32, 1 float64, 2 int32, 3 int64, 4 uint32 */
CB_API int cb_cast(cb_ctx *ctx, void *d_dst, int dst_kind, const void *d_src, int src_kind, uint64_t n);
/* exp (op 0) / log (op 1) of a float32 or float64 array: cn.exp(x_T) -> terminal prices, the step between the
 * simulation and the density estimate in the reference's workflow (notebooks/Getting Started.ipynb cells 65-83) */
CB_API int cb_unary(cb_ctx *ctx, void *d_dst, const void *d_src, uint64_t n, int op, int is_f64);
/* 2-D transpose of a row-major rows x cols array of 4- or 8-byte elements (ndarray.T / transpose) */
CB_API int cb_transpose(cb_ctx *ctx, void *d_dst, const void *d_src, uint64_t rows, uint64_t cols, int elem_bytes);

/* ---- random arrays: cocos/numerics/random.py:127-153 (rand, randn) -------
 * Element order and stream layout are documented in DESIGN.md; `subsequence`
 * separates launches drawn from one seed (the host generator hands them out). */
CB_API int cb_philox_words(cb_ctx *ctx, uint32_t *d_out, uint64_t n_blocks, uint64_t first_index,
                    uint32_t block, uint64_t seed, uint32_t subsequence);      /* raw 4-word blocks (KAT) */
CB_API int cb_rand_f32(cb_ctx *ctx, float *d_out, uint64_t n, uint64_t seed, uint32_t subsequence);
CB_API int cb_rand_f64(cb_ctx *ctx, double *d_out, uint64_t n, uint64_t seed, uint32_t subsequence);
CB_API int cb_randn_f32(cb_ctx *ctx, float *d_out, uint64_t n, uint64_t seed, uint32_t subsequence);
CB_API int cb_randn_f64(cb_ctx *ctx, double *d_out, uint64_t n, uint64_t seed, uint32_t subsequence);
/* randn_antithetic / rand_antithetic (random.py:190-245): `dims[ndim]` is the OUTPUT
 * shape (row-major), draws fill the first ceil(d/2) entries along `axis`, the
 * remaining floor(d/2) are the reflections (-z, or 1-u). */
CB_API int cb_randn_antithetic_f32(cb_ctx *ctx, float *d_out, const uint64_t *dims, int ndim, int axis,
                            uint64_t seed, uint32_t subsequence);
CB_API int cb_randn_antithetic_f64(cb_ctx *ctx, double *d_out, const uint64_t *dims, int ndim, int axis,
                            uint64_t seed, uint32_t subsequence);
CB_API int cb_rand_antithetic_f32(cb_ctx *ctx, float *d_out, const uint64_t *dims, int ndim, int axis,
                           uint64_t seed, uint32_t subsequence);
CB_API int cb_rand_antithetic_f64(cb_ctx *ctx, double *d_out, const uint64_t *dims, int ndim, int axis,
                           uint64_t seed, uint32_t subsequence);

/* ---- derived distributions: cocos/numerics/random.py:252-531 ---------------
 * kind: 0 uniform(a=low, b=high)   1 exponential(a=scale)        2 normal(a=loc, b=scale)
 *       3 lognormal(a=mean, b=sigma) 4 logistic(a=loc, b=scale)  5 wald(a=mean, b=scale)
 *       6 gamma(a=shape, b=scale) [chisquare(df) = gamma(df/2, 2)]   7 beta(a, b)
 * float32 output, as in the reference (its derived distributions all start from rand(n)/randn(n)). */
CB_API int cb_random_distribution_f32(cb_ctx *ctx, float *d_out, uint64_t n, int kind, double a, double b,
                                      uint64_t seed, uint32_t subsequence);
/* randint (random.py:252-287): integers in [low, high); int32 or int64 output */
CB_API int cb_randint(cb_ctx *ctx, void *d_out, int out_is_64, uint64_t n, int64_t low, int64_t high,
                      uint64_t seed, uint32_t subsequence);
/* choice (random.py:290-303): out[i] = src[index[i]] for 4- or 8-byte elements */
CB_API int cb_gather(cb_ctx *ctx, void *d_out, const void *d_src, const int32_t *d_index, uint64_t n, int elem_bytes);
/* multivariate_normal (random.py:523-531): rows of randn(rows, d) times cholesky(cov)^T plus mean;
 * h_chol is the LOWER Cholesky factor, row-major d x d (host), h_mean has d entries (host) */
CB_API int cb_multivariate_normal_f32(cb_ctx *ctx, float *d_out, uint64_t rows, int d, const float *h_chol,
                                      const float *h_mean, uint64_t seed, uint32_t subsequence);
CB_API int cb_multivariate_normal_f64(cb_ctx *ctx, double *d_out, uint64_t rows, int d, const double *h_chol,
                                      const double *h_mean, uint64_t seed, uint32_t subsequence);

/* ---- Gaussian KDE: cocos/scientific/kde.py:152-197 (evaluate), :1057-1170 (covariance, whitening) -------
 * points are row-major (n, d) float32 on the device.
 * cb_kde_moments_f32   h_out = [sum w, sum w x (d), sum w x x^T (d*d)] in double (weights NULL = all ones)
 * cb_kde_whiten_f32    d_out = d_in @ M for a host d x d row-major matrix M (the whitening, pre-scaled)
 * cb_kde_evaluate_f32  d_out[j] = scale * sum_i w_i * 2^(-|p_i - q_j|^2) for whitened, pre-scaled p (n,d), q (m,d) */
CB_API int cb_kde_moments_f32(cb_ctx *ctx, const float *d_points, const float *d_weights_or_null, uint64_t n, int d,
                              double *h_out);
CB_API int cb_kde_whiten_f32(cb_ctx *ctx, const float *d_in, uint64_t n, int d, const float *h_matrix, float *d_out);
CB_API int cb_kde_evaluate_f32(cb_ctx *ctx, const float *d_wpoints, const float *d_weights_or_null, uint64_t n, int d,
                               const float *d_wxi, uint64_t m, double scale, float *d_out);
/* integrate_box_1d (kde.py:869-902): sum_i w_i (Phi((high-x_i)/stdev) - Phi((low-x_i)/stdev)), d_points = the raw 1-D
 * dataset (n floats), weights NULL = 1 each (the caller divides by n); accumulated in double */
CB_API int cb_kde_integrate_box_1d_f32(cb_ctx *ctx, const float *d_points, const float *d_weights_or_null, uint64_t n,
                                       double low, double high, double stdev, double *h_out);
/* resample (kde.py:987-1025): d_out (d, size) row-major = dataset[:, index] + chol @ z with index drawn from the
 * weights (d_cumweights = inclusive cumulative sums in double, NULL = uniform) and z ~ N(0, I); h_chol = LOWER Cholesky factor
 * of the kernel covariance (host, row-major d x d); one Philox block per sample */
CB_API int cb_kde_resample_f32(cb_ctx *ctx, const float *d_points, const double *d_cumweights_or_null, uint64_t n, int d,
                               const float *h_chol, uint64_t size, uint64_t seed, uint32_t subsequence, float *d_out);

/* ---- Heston paths ---------------------------------------------------------
 * Parity kernels: simulate_heston_model (heston_utilities.py:58-96) with the
 * normals injected.  d_Z is [N-1][ceil(R/2)][2]; path i < ceil(R/2) uses +Z, path
 * i >= ceil(R/2) uses -Z[i-ceil(R/2)] (random.py:199-214).  Operation order and
 * roundings are the reference's; d_x, d_v receive the R terminal values. */
CB_API int cb_heston_terminal_injected_f32(cb_ctx *ctx, const float *d_Z, const cb_heston_params *p,
                                    uint64_t R, uint32_t N, float *d_x, float *d_v);
CB_API int cb_heston_terminal_injected_f64(cb_ctx *ctx, const double *d_Z, const cb_heston_params *p,
                                    uint64_t R, uint32_t N, double *d_x, double *d_v);
/* Same, taking HOST buffers: copies Z in, runs, copies x and v out (end-to-end form). */
CB_API int cb_heston_terminal_injected_host_f32(cb_ctx *ctx, const float *h_Z, const cb_heston_params *p,
                                         uint64_t R, uint32_t N, float *h_x, float *h_v);

/* Fused kernel: in-kernel Philox + Box-Muller + antithetic pair + Euler steps +
 * call payoff reduction = simulate_and_compute_option_price
 * (heston_utilities.py:122-164) without materialising anything.
 *   R, N          paths and time points of the whole simulation (N-1 Euler steps)
 *   pair_first, pair_count   the antithetic pairs [pair_first, pair_first+pair_count)
 *                 of the ceil(R/2) global pairs this call simulates (a shard);
 *                 pair p owns paths p and p+ceil(R/2); the last pair of an odd R
 *                 has no partner
 *   h_strikes[nK] strikes sharing the same paths (1 <= nK <= CB_MAX_STRIKES)
 *   comm          NULL, or a communicator: the 2*nK partial sums of all ranks are
 *                 added up (replaces ComputeDevicePool.map_reduce's host fold,
 *                 cocos/multi_processing/device_pool.py:246-251) -- inside the
 *                 kernel over NVLink peer memory when the communicator has peer
 *                 mappings (cb_comm_p2p_*) and the launch uses the default
 *                 geometry, else by one ncclAllReduce(sum, double) behind the
 *                 kernel on the same stream; every rank must make the same call
 *   d_x, d_v      NULL, or device arrays receiving terminal values of this shard:
 *                 path of pair p at [p-pair_first], partner at
 *                 [pair_count + p-pair_first] (reference layout when the shard is
 *                 the whole simulation)
 *   d_traj        NULL, or [N][2*pair_count] (same path layout per row): the whole
 *                 log-price trajectory, row 0 = x0 (path-materialising mode)
 * _f64 keeps the state (x, v), the updates and the payoff sums in double; its normals
 * and the diffusion square root are single-precision grade like the _f32 kernel's
 * (MUFU lg2/sin/cos, RSQ64H seed; bias against an exact-libm double model of the same
 * stream: -6.5e-9 in price units).  Bit-for-bit double arithmetic on
 * caller-supplied normals is cb_heston_terminal_injected_f64.
 * _launch is asynchronous; _finish synchronises the stream and copies the sums
 * (UNDISCOUNTED: sum of payoffs, sum of squared antithetic-unit averages) out. */
CB_API int cb_heston_price_launch_f32(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                               const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                               uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                               float *d_x, float *d_v, float *d_traj);
CB_API int cb_heston_price_launch_f64(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                               const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                               uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                               double *d_x, double *d_v, double *d_traj);
CB_API int cb_heston_price_finish(cb_ctx *ctx, int nK, double *h_sum, double *h_sumsq);
/* Same fused simulation with a path-dependent (or put) payoff; sums are collected with cb_heston_price_finish. */
CB_API int cb_heston_payoff_launch_f32(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                                       const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                       uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                       const cb_payoff_spec *spec);
CB_API int cb_heston_payoff_launch_f64(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                                       const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                       uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                       const cb_payoff_spec *spec);
/* Path-materialising mode with both state variables (the HBM-write-bound variant, 8 bytes per path-step): the
 * log price x_t of every time point goes to d_traj_x and the variance v_t to d_traj_v (NULL: log prices only, i.e.
 * cb_heston_price_launch's d_traj), each [N][2*pair_count], row t = time point t (row 0 = x0 / v0), columns
 * [first halves | partners] of the shard -- the state simulate_heston_model ping-pongs through
 * (heston_utilities.py:60-64, :85-91), kept instead of overwritten.  The stores are 2-D TMA tensor stores from a
 * shared-memory stage when pair_count*sizeof(real) is a multiple of 16 and the shard holds no unpaired path,
 * warp-coalesced direct stores otherwise.  d_x/d_v (terminal state) may be NULL; sums as in cb_heston_price_launch. */
CB_API int cb_heston_trajectory_launch_f32(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                                           const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                           uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                           float *d_x, float *d_v, float *d_traj_x, float *d_traj_v);
CB_API int cb_heston_trajectory_launch_f64(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                                           const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                           uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                           double *d_x, double *d_v, double *d_traj_x, double *d_traj_v);
/* The whole multi-GPU job of simulate_and_compute_option_price_gpu (heston_utilities.py:217-270) in one call:
 * shard g of G owns pairs [g*ceil(P/G), min((g+1)*ceil(P/G), P)), P = ceil(R/2); every context's stream gets its
 * kernel followed by the one ncclAllReduce (grouped); the reduced sums are read once from ctxs[0].  comms may be
 * NULL when G == 1; spec may be NULL (European call).  Results do not depend on G beyond f64 summation order. */
CB_API int cb_heston_price_multi_f32(int G, cb_ctx *const *ctxs, cb_comm *const *comms, const cb_heston_params *p,
                                     uint64_t R, uint32_t N, const double *h_strikes, int nK, uint64_t seed,
                                     uint32_t subsequence, const cb_payoff_spec *spec, double *h_sum, double *h_sumsq);
CB_API int cb_heston_price_multi_f64(int G, cb_ctx *const *ctxs, cb_comm *const *comms, const cb_heston_params *p,
                                     uint64_t R, uint32_t N, const double *h_strikes, int nK, uint64_t seed,
                                     uint32_t subsequence, const cb_payoff_spec *spec, double *h_sum, double *h_sumsq);
/* Opt-in scheme with the same law of terminal prices (and of antithetic pairs) as the Euler kernel above but half
 * the normals: the log price is sampled conditionally on the variance path (csrc/heston_conditional.cu).  Statistical
 * mode only (no trajectory); results are collected with cb_heston_price_finish. */
CB_API int cb_heston_price_conditional_launch_f32(cb_ctx *ctx, const cb_heston_params *p, uint64_t R, uint32_t N,
                                                  const double *h_strikes, int nK, uint64_t seed, uint32_t subsequence,
                                                  uint64_t pair_first, uint64_t pair_count, cb_comm *comm,
                                                  float *d_x, float *d_v);

/* compute_option_price_from_simulated_paths (heston_utilities.py:99-119) on a
 * device array of terminal log prices in the reference layout. */
CB_API int cb_call_payoff_sums_f32(cb_ctx *ctx, const float *d_x, uint64_t R, const double *h_strikes, int nK,
                            double *h_sum, double *h_sumsq);
CB_API int cb_call_payoff_sums_f64(cb_ctx *ctx, const double *d_x, uint64_t R, const double *h_strikes, int nK,
                            double *h_sum, double *h_sumsq);

/* ---- Monte-Carlo pi: estimate_pi (examples/monte_carlo_pi/
 * monte_carlo_pi_multi_gpu_example.py:18-33) fused: draws, test and count in one
 * kernel.  antithetic != 0 gives the rand_antithetic([n], 0) point layout. */
CB_API int cb_pi_count_launch(cb_ctx *ctx, uint64_t n, uint64_t seed, uint32_t subsequence, int antithetic,
                       uint64_t draw_first, uint64_t draw_count, cb_comm *comm);
CB_API int cb_pi_count_finish(cb_ctx *ctx, uint64_t *h_inside);
/* estimate_pi over a device pool in one call (draw shards as above, one allreduce of the count). */
CB_API int cb_pi_count_multi(int G, cb_ctx *const *ctxs, cb_comm *const *comms, uint64_t n, uint64_t seed,
                             uint32_t subsequence, int antithetic, uint64_t *h_inside);

/* NCCL is loaded lazily on the first communicator call: a copy already mapped into the process wins, then the
 * path given here (or $COCOS_B200_NCCL_LIBRARY), then the system libnccl.so.2.  Name the host framework's bundled
 * copy when that framework may be imported AFTER the first communicator is made (same SONAME, newer version). */
CB_API int cb_nccl_library(const char *path);
/* ---- collectives: stands in for ComputeDevicePool's process pool
 * (cocos/multi_processing/device_pool.py:58-253).  libnccl.so.2 is loaded on
 * first use. */
CB_API int cb_comm_unique_id(void *h_id /* CB_UNIQUE_ID_BYTES */);
CB_API int cb_comm_init_rank(cb_ctx *ctx, int nranks, int rank, const void *h_id, cb_comm **out);
CB_API int cb_comm_init_all(int n, cb_ctx *const *ctxs, cb_comm **out /* [n] */);   /* one process, n devices */
CB_API int cb_comm_destroy(cb_comm *comm);
/* Fused peer exchange: with it the cross-GPU sum of the fused Heston launches runs INSIDE the simulation kernel --
 * every rank's last block writes its totals into every rank's exchange buffer through NVLink peer mappings and folds
 * them in rank order -- and no collective follows the kernel.  cb_comm_init_all connects the buffers by itself when
 * every device pair has peer access.  One process per GPU: each rank exports its buffer as a CUDA IPC handle, the
 * host gathers the nranks handles (rank order), each rank imports them, and once EVERY rank has succeeded all call
 * cb_comm_p2p_enable(comm, 1); otherwise the ncclAllReduce stays in the path. */
CB_API int cb_comm_p2p_export(cb_ctx *ctx, cb_comm *comm, void *h_handle /* CB_P2P_HANDLE_BYTES */);
CB_API int cb_comm_p2p_import(cb_ctx *ctx, cb_comm *comm, const void *h_handles /* nranks * CB_P2P_HANDLE_BYTES */);
CB_API int cb_comm_p2p_enable(cb_comm *comm, int enable);
CB_API int cb_comm_p2p_active(const cb_comm *comm, int *active);
/* Tear-down across processes: every rank unmaps its peers' buffers, the host synchronises the ranks, then each
 * rank calls cb_comm_destroy (an exported buffer must not be freed while another process still maps it). */
CB_API int cb_comm_p2p_disconnect(cb_comm *comm);
CB_API int cb_allreduce_sum_f64(cb_ctx *ctx, cb_comm *comm, double *d_buf, size_t count);
CB_API int cb_group_start(void);
CB_API int cb_group_end(void);

/* ---- pipe-peak calibration (roofline denominators) ---------
 * kind: 0 FFMA, 1 MUFU.LG2, 2 MUFU.SQRT, 3 MUFU.SIN, 4 MUFU.EX2, 5 MUFU.RSQ, 6 DFMA,
 *       7 Philox4x32-10 calls, 8 HBM write (bytes), 9 LOP3 (ALU pipe),
 *       10 HBM copy (read+write bytes), 11-17 integer-multiply issue-cost experiments
 *       (IMAD.WIDE alone / with FFMA / with LOP3, IMAD.HI, IMAD lo),
 *       18-20 packed FFMA2 (fma.rn.f32x2) alone / with FFMA / with LOP3,
 *       21-38 pipe-model probes: mixes of half Philox rounds (IMAD.WIDE + XOR), three-register FFMA, LOP3 and
 *       FADD on independent chains (csrc/peaks.cu: peak_mix_kernel)
 * ops_per_s: thread-level operations per second (FFMA counts 1 op = 2 flop). */
CB_API int cb_peak_measure(cb_ctx *ctx, int kind, int repeats, double *ops_per_s, float *ms_best);

#ifdef __cplusplus
}
#endif
#endif /* COCOS_B200_H */
