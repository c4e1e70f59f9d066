// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cocos_b200.h"
#include "philox.cuh"

namespace cb {

constexpr int kMaxStrikes = CB_MAX_STRIKES;
constexpr int kMaxAcc = 2 * kMaxStrikes;          // sums then sums of squares
constexpr int kMaxGridBlocks = 148 * 32 * 2;      // workspace rows for per-block partials

// launch geometry chosen by the host for one fused launch
struct LaunchShape {
    int threads;            // per block
    int blocks;             // grid
    int pairs_per_thread;   // 1, 2 or 4 interleaved pairs per thread
    int variant;            // 0 = default kernel; >0 = experimental variants of the float single-strike kernel
};

// Launch-uniform constants of the fused kernels, precomputed in double on the host.
template <typename real>
struct FusedConsts {
    real drift_v;        // -0.5*dt                 x += drift_v*v + sqrt(v)*dB0   (mu*dt is added once, see x_shift)
    real decay;          // 1 - kappa*dt            v  = decay*v + pull + sqrt(v)*w
    real pull;           // kappa*v_bar*dt
    real w0, w1;         // sigma_v*rho, sigma_v*sqrt(1-rho^2)   w = w0*dB0 + w1*dB1
    real v0;
    real x0;
    real mu_dt;          // mu*dt (trajectory rows and the final shift use it)
    real floor_v;        // float32 eps, also for double state (heston_utilities.py:91)
    float neg2ln2_dt;    // -2 ln2 * dt: MUFU.LG2 output -> (r*sqrt(dt))^2
    uint32_t steps;      // N-1 Euler steps
    uint32_t subsequence;
    uint64_t pair_first; // shard
    uint64_t pair_count;
    uint64_t partnered;  // floor(R/2): pairs with global index < partnered own two paths
    int nK;
    // payoff (0 = terminal/European, 1 = arithmetic-average Asian, 2 = knock-out barrier, discretely monitored)
    int payoff_kind;
    real payoff_sign;    // +1 call, -1 put
    real barrier_sign;   // +1 up-and-out, -1 down-and-out
    real barrier_log;    // barrier_sign * log(barrier level)
    PhiloxKeys keys;
    real strikes[kMaxStrikes];
};

// Cross-GPU reduction fused into the simulation kernel (no collective library in the data path): the last block of
// every rank writes its totals into slot [epoch parity][rank] of EVERY rank's exchange buffer through NVLink peer
// mappings, raises its flag there (release, system scope), waits until all ranks' flags have arrived on its own
// device (acquire) and folds the slots in rank order -- every rank ends with the same bits in `result`.
// Two slot parities make the next launch's writes safe: a rank can only be one launch ahead of the slowest one.
// Every slot carries a tag (launch number and element count of the data it holds) that the folding rank verifies:
// ranks whose launch counters have drifted apart (a failed launch on one of them, mismatched calls) end with the
// poison value in `result` instead of a silently wrong sum.
constexpr int kMaxPeers = 8;
constexpr int kSlotStride = kMaxAcc + 2;   // kMaxAcc values + the tag (+ padding to 16 bytes), in 8-byte words
struct PeerExchange {
    int nranks;                       // < 2: no exchange
    int rank;
    unsigned long long *slots[kMaxPeers];   // rank p's buffer, mapped here: [2][kMaxPeers][kSlotStride]
    unsigned int *flags[kMaxPeers];   // rank p's flags, mapped here: [kMaxPeers], flags[r] = last epoch published by r
};
constexpr size_t kExchangeSlotBytes = sizeof(unsigned long long) * 2 * kMaxPeers * kSlotStride;
constexpr size_t kExchangeBytes = kExchangeSlotBytes + sizeof(unsigned int) * kMaxPeers;
// what `result` holds when a peer never showed up or its slot carried another launch's tag (quiet NaNs with
// recognisable payloads; as integers they are far above any possible count)
constexpr unsigned long long kExchangeTimeoutBits = 0x7FF8C0C05B200DEAull;
constexpr unsigned long long kExchangeMismatchBits = 0x7FF8C0C05B200BADull;

struct ReduceWorkspace {
    double *partials;    // [kMaxGridBlocks][kMaxAcc]
    unsigned int *ticket;
    double *result;      // [kMaxAcc]: sums[nK] then sumsq[nK]
    // fused cross-GPU step (XCHG kernels only): the communicator's mapping table, resident on this device -- eight bytes
    // of kernel parameters instead of the table itself, read by the last block only -- and this launch's number
    const PeerExchange *px;
    unsigned int epoch;  // launch number on the communicator (> 0), identical on every rank
    unsigned int sig;    // signature of the call (kind, seed, subsequence, sizes): ranks must agree on it, the slots' tags carry it
};

template <typename real>
cudaError_t launch_heston_fused(const FusedConsts<real> &c, const LaunchShape &shape, ReduceWorkspace ws,
                                real *d_x, real *d_v, cudaStream_t stream);

// true when launch_heston_fused has an instantiation with the fused cross-GPU exchange for this geometry
template <typename real>
bool fused_exchange_supported(const LaunchShape &shape, int payoff_kind);

// occupancy and register count of the instantiation launch_heston_fused would pick (xchg: with the fused exchange)
template <typename real>
cudaError_t fused_occupancy(int threads, int pairs_per_thread, bool multi, int payoff_kind, int variant, bool xchg,
                            int *blocks_per_sm, int *regs);

// path-materialising kernel (heston_trajectory.cu): every step of x (and v when d_traj_v != NULL) goes to HBM
template <typename real>
bool trajectory_bulk_eligible(const FusedConsts<real> &c, const real *d_traj_x, const real *d_traj_v);
template <typename real>
cudaError_t trajectory_occupancy(bool with_v, bool multi, bool bulk, int *threads, int *blocks_per_sm);
template <typename real>
cudaError_t launch_heston_trajectory(const FusedConsts<real> &c, int blocks, bool bulk, ReduceWorkspace ws,
                                     real *d_x, real *d_v, real *d_traj_x, real *d_traj_v, cudaStream_t stream);

// conditional-Gaussian scheme (heston_conditional.cu)
cudaError_t launch_heston_conditional(const FusedConsts<float> &c, float sigma_v, float rho, float dt, int threads,
                                      int blocks, ReduceWorkspace ws, float *d_x, float *d_v, cudaStream_t stream);
cudaError_t conditional_occupancy(int threads, bool multi, int *blocks_per_sm);

// parity kernels (reference operation order, injected normals)
struct InjectedConsts32 {
    float dt, root_dt, m0, m1, mu, kappa, v_bar, sigma_v, x0, v0;
};
struct InjectedConsts64 {
    double dt, root_dt, m0, m1, mu, kappa, v_bar, sigma_v, x0, v0;
};
cudaError_t launch_heston_injected_f32(const float *d_Z, const InjectedConsts32 &c, uint64_t R, uint32_t N,
                                       float *d_x, float *d_v, cudaStream_t stream);
cudaError_t launch_heston_injected_f64(const double *d_Z, const InjectedConsts64 &c, uint64_t R, uint32_t N,
                                       double *d_x, double *d_v, cudaStream_t stream);

// payoff reduction over a device array of terminal log prices
template <typename real>
cudaError_t launch_call_payoff_sums(const real *d_x, uint64_t R, const double *strikes, int nK,
                                    ReduceWorkspace ws, int sm_count, cudaStream_t stream);

// random arrays
cudaError_t launch_philox_words(uint32_t *d_out, uint64_t n_blocks, uint64_t first, uint32_t block,
                                const PhiloxKeys &k, uint32_t subseq, cudaStream_t s);
template <typename real, bool NORMAL>
cudaError_t launch_random_fill(real *d_out, uint64_t n, const PhiloxKeys &k, uint32_t subseq, cudaStream_t s);
template <typename real, bool NORMAL>
cudaError_t launch_random_antithetic(real *d_out, uint64_t outer, uint64_t d_axis, uint64_t inner,
                                     const PhiloxKeys &k, uint32_t subseq, cudaStream_t s);

// derived distributions (dist_kernels.cu)
cudaError_t launch_dist_fill(int kind, float *d_out, uint64_t n, float a, float b, const PhiloxKeys &k,
                             uint32_t subseq, cudaStream_t s);
cudaError_t launch_randint(void *d_out, int out_is_64, uint64_t n, float divisor, long long low, const PhiloxKeys &k,
                           uint32_t subseq, cudaStream_t s);
cudaError_t launch_gather(void *d_out, const void *d_src, const int *d_idx, uint64_t n, int elem_bytes, cudaStream_t s);
template <typename real>
cudaError_t launch_mvn_transform(real *d_zx, uint64_t rows, int d, const real *d_chol, const real *d_mean,
                                 cudaStream_t s);

// Gaussian KDE (kde_kernels.cu)
cudaError_t launch_kde_moments(const float *d_pts, const float *d_w, uint64_t n, int d, double *d_out, cudaStream_t s);
cudaError_t launch_kde_whiten(const float *d_in, uint64_t n, int d, const float *d_W, float *d_out, cudaStream_t s);
cudaError_t launch_kde_box_1d(const float *d_pts, const float *d_w, uint64_t n, double low, double high, double sd,
                              double *d_out, cudaStream_t s);
cudaError_t launch_kde_resample(const float *d_pts, const double *d_cumw, uint64_t n, int d, const float *d_chol,
                                uint64_t size, const PhiloxKeys &k, uint32_t subseq, float *d_out, cudaStream_t s);
int kde_plan_chunks(uint64_t n, uint64_t m, int sm_count, int *pj_out);
cudaError_t launch_kde_evaluate(const float *d_p, const float *d_w, uint64_t n, int d, const float *d_q, uint64_t m,
                                double scale, float *d_partial, int chunks, int pj, float *d_out, cudaStream_t s);

// Monte-Carlo pi
cudaError_t launch_pi_count(uint64_t n, int antithetic, uint64_t draw_first, uint64_t draw_count,
                            const PhiloxKeys &k, uint32_t subseq, unsigned long long *d_inside,
                            int resident_blocks, ReduceWorkspace ws, cudaStream_t s);
cudaError_t pi_occupancy(int antithetic, int *blocks_per_sm);

// storage plumbing of the device array handle (array_kernels.cu); kinds: 0 f32, 1 f64, 2 i32, 3 i64, 4 u32
cudaError_t launch_cast(void *dst, int dst_kind, const void *src, int src_kind, uint64_t n, cudaStream_t s);
cudaError_t launch_unary(void *dst, const void *src, uint64_t n, int op, int is64, cudaStream_t s);   // op 0 exp, 1 log
cudaError_t launch_transpose(void *dst, const void *src, uint64_t rows, uint64_t cols, int elem_bytes, cudaStream_t s);

// calibration
cudaError_t launch_peak_kernel(int kind, int sm_count, void *d_scratch, size_t scratch_bytes, double *work_done,
                               cudaStream_t s);

}  // namespace cb
