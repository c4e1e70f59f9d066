// Path-materialising Heston kernel (sm_100a): the fused simulation of heston_fused.cu with EVERY step written to
// HBM -- log prices, and optionally variances -- while the payoff sums are still reduced in the same launch.
//
//   simulate_heston_model (the x/v state of every step)   examples/stochastic_volatility/heston_utilities.py:58-96
//   "paths requested" mode of the north star; the reference itself only ever returns the terminal state (:93-96)
//
// Output layout (HBM): traj_x[N][2*pairs] (and traj_v likewise), row t = time point t, columns
// [first halves | antithetic partners] of the shard's pairs -- a block owns `blockDim.x` CONSECUTIVE pairs, so per
// step and half it produces one contiguous run of blockDim.x values.
//
// HBM-write bound: 4 B (x) or 8 B (x and v) per path-step and nothing read.  The stores must not sit on the issue
// port of a kernel that is already issue-bound, so they do not go through STG at all:
//   * every step each thread drops its values into a shared-memory stage with plain STS (conflict-free: consecutive
//     threads, consecutive words);
//   * after kStage steps ONE thread of the block hands the stage to the TMA unit: one 2-D tensor store
//     (cp.async.bulk.tensor.2d, SASS UTMASTG) per staged BOX -- [kStage rows][256 columns] of the trajectory matrix,
//     one box per array and half for a 256-thread block, two side by side for a 512-thread one -- i.e. 2 to 8
//     instructions per STAGE for the whole block instead of 2 (or 4) STG + address arithmetic per THREAD and step.
//     The issuing thread's serial work between the two halves of the block barrier is a dozen instructions, so the
//     other warps hardly wait for it.  One tensor map per array and half ([N][pair_count] with the row pitch of the
//     full matrix): columns beyond the shard's pairs and rows beyond the last step are clipped by the TMA unit, so
//     ragged tiles and the 1..5 left-over steps need no special case;
//   * two stages alternate; a stage is reused only after the bulk group that read it has finished reading
//     (cp.async.bulk.wait_group.read by the issuing thread, then the block barrier that every flush has anyway).
// Shards whose rows cannot be cut into 16-byte aligned segments (pair count not a multiple of 16 bytes, or the lone
// unpaired path of an odd R inside the shard) take the DIRECT variant: predicated, warp-coalesced STG per step.
#include <cstring>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "kernels.h"
#include "heston_step.cuh"
#include "payoff_reduce.cuh"

namespace cb {

#ifndef CB_TRAJ_STAGE
#define CB_TRAJ_STAGE 6
#endif
#ifndef CB_TRAJ_MINB
#define CB_TRAJ_MINB 1
#endif
constexpr int kStage = CB_TRAJ_STAGE;   // steps per stage: a multiple of three (one Philox block feeds three steps)
static_assert(kStage % 3 == 0, "a stage is a whole number of Philox blocks");
constexpr int kBox = 256;                       // columns of one TMA box (the hardware limit per dimension)
// Block width (fixed per instantiation: stage offsets are immediates).  The float (x, v) variant is HBM-bound and runs
// 512 threads wide -- each row segment it writes is 2 KiB instead of 1 KiB, two boxes side by side: 0.526-0.538 ms
// against 0.553-0.558 ms on the same box (DRAM page locality); everything else is issue-bound and fastest at 256.
template <typename real, bool WITH_V, bool BULK>
constexpr int kTrajThreadsOf = (WITH_V && BULK && sizeof(real) == 4) ? 512 : 256;

// tensor maps of one launch: [0] x first halves, [1] x partners, [2] v first halves, [3] v partners
struct alignas(64) TrajMaps {
    CUtensorMap m[4];
};

// smem box [kStage][kBox] -> rows row.., columns col.. of the mapped matrix (out-of-range parts are clipped)
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap *map, uint32_t col, uint32_t row, uint32_t ssrc) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col),
                 "r"(row), "r"(ssrc)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_reads() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Launch bounds, 256-thread variants: uncapped registers measured best -- x only: 71 registers / 3 resident blocks
// 0.403 ms against 0.427 ms at 48 registers / 5 blocks.  The 512-thread (x, v) variant is capped at 64 registers so
// that two blocks are resident (at 256 threads it was within 1% for every cap tried: 79 / 64 / 48 registers).
// WITH_V: also materialise the variance; BULK: TMA tensor stores from a shared-memory stage (else direct STG)
template <typename real, bool WITH_V, bool MULTI, bool BULK>
__global__ void __launch_bounds__((kTrajThreadsOf<real, WITH_V, BULK>),
                                  (kTrajThreadsOf<real, WITH_V, BULK> == 512 ? 2 : CB_TRAJ_MINB))
heston_trajectory_kernel(const __grid_constant__ FusedConsts<real> c, const __grid_constant__ TrajMaps maps,
                         ReduceWorkspace ws, real *__restrict__ x_out, real *__restrict__ v_out,
                         real *__restrict__ traj_x, real *__restrict__ traj_v) {
    constexpr int NARR = WITH_V ? 4 : 2;                  // staged arrays per step: x1, x2 (, v1, v2)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    real *stage = reinterpret_cast<real *>(smem_raw);     // [2][NARR][kSubs][kStage][kBox]: dense boxes
    constexpr int kTrajThreads = kTrajThreadsOf<real, WITH_V, BULK>;
    constexpr int kSubs = kTrajThreads / kBox;            // boxes side by side per block and array
    constexpr uint32_t B = kTrajThreads;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t tiles = (c.pair_count + B - 1) / B;
    const uint64_t rounds = (tiles + gridDim.x - 1) / gridDim.x;
    const uint64_t row_len = 2 * c.pair_count;

    __shared__ real s_strikes[kMaxStrikes];
    if (MULTI) stage_strikes(s_strikes, c.strikes, c.nK);
    PayoffAccumulator<real, MULTI> acc;
    const real x_shift = fma((real)c.steps, c.mu_dt, c.x0);
    const StepRegs<real> k(c);
    const uint32_t one_bits = 0x3F800000u | (c.steps >> 31);
    PhiloxKeys keys;
#pragma unroll
    for (int r = 0; r < 10; ++r) { keys.k0[r] = pin(c.keys.k0[r]); keys.k1[r] = pin(c.keys.k1[r]); }

    constexpr uint32_t stage_elems = kStage * NARR * B;   // one stage, in elements
    const uint32_t stage_smem = (uint32_t)__cvta_generic_to_shared(stage);
    // the two stages alternate for the whole launch (not per tile): a stage is rewritten only after the flush that
    // waited for its last reader
    uint32_t parity = 0;
    const uint32_t my_col = (threadIdx.x / kBox) * (kStage * kBox) + (threadIdx.x % kBox);
    real *mine = stage + my_col;                           // this thread's column in the current stage

    for (uint64_t round = 0; round < rounds; ++round) {
        const uint64_t tile = round * gridDim.x + blockIdx.x;
        if (tile >= tiles) break;                          // whole block: uniform, no barrier is skipped by a part of it
        const uint64_t tile_first = tile * B;
        const uint64_t local = tile_first + threadIdx.x;
        const bool active = local < c.pair_count;
        const uint64_t gp = c.pair_first + (active ? local : 0);     // idle threads replay pair 0, masked everywhere
        const bool paired = active && gp < c.partnered;
        PairStream rng;
        rng.init((uint32_t)gp, (uint32_t)(gp >> 32), c.subsequence, c.keys);
        real x1 = (real)0, x2 = (real)0, u1 = k.u0, u2 = k.u0;     // u = c1*v inside the time loop

        // row 0: the initial state (direct, coalesced; one row out of N)
        if (active) { traj_x[local] = c.x0; if (WITH_V) traj_v[local] = c.v0; }
        if (paired) { traj_x[c.pair_count + local] = c.x0; if (WITH_V) traj_v[c.pair_count + local] = c.v0; }

        real step_f = (real)0;
        uint32_t flushed = 0;                              // steps already handed to the copy engine (thread 0 needs it)
        real *row1 = traj_x + local, *row2 = traj_x + c.pair_count + local;          // DIRECT variant
        real *vrow1 = traj_v + local, *vrow2 = traj_v + c.pair_count + local;

        // one step's values -> stage slot s (BULK) or straight to HBM (DIRECT)
        auto emit = [&](int s) {
            step_f += (real)1;                                   // exact: steps < 2^24
            const real shift = fma(step_f, c.mu_dt, c.x0);       // the drift the loop keeps out of x1/x2
            const real y1 = x1 + shift, y2 = x2 + shift;
            real w1, w2;
            if constexpr (WITH_V) { w1 = plain_max(u1 * k.inv_c1, c.floor_v); w2 = plain_max(u2 * k.inv_c1, c.floor_v); }
            if constexpr (BULK) {
                mine[(0 * kSubs * kStage + s) * kBox] = y1;
                mine[(1 * kSubs * kStage + s) * kBox] = y2;
                if constexpr (WITH_V) { mine[(2 * kSubs * kStage + s) * kBox] = w1; mine[(3 * kSubs * kStage + s) * kBox] = w2; }
            } else {
                row1 += row_len;
                row2 += row_len;
                if (active) *row1 = y1;
                if (paired) *row2 = y2;
                if constexpr (WITH_V) {
                    vrow1 += row_len;
                    vrow2 += row_len;
                    if (active) *vrow1 = w1;
                    if (paired) *vrow2 = w2;
                }
            }
        };
        // hand the staged steps to the copy engine and switch stages (every thread of the block calls this); `n` of
        // the kStage rows are new -- the rows after them lie beyond the last time point and are clipped
        auto flush = [&](uint32_t n) {
            if constexpr (BULK) {
                fence_async_smem();                              // my STS are visible to the async proxy
                if (threadIdx.x == 0) bulk_wait_reads();         // the OTHER stage (issued one flush ago) has been read
                __syncthreads();
                if (threadIdx.x == 0) {
                    constexpr uint32_t box_bytes = kStage * kBox * (uint32_t)sizeof(real);
                    const uint32_t src0 = stage_smem + parity * stage_elems * (uint32_t)sizeof(real);
#pragma unroll
                    for (int a = 0; a < NARR; ++a)
#pragma unroll
                        for (int sub = 0; sub < kSubs; ++sub)
                            tensor_store_2d(&maps.m[a], (uint32_t)tile_first + sub * kBox, flushed + 1,
                                            src0 + (a * kSubs + sub) * box_bytes);
                    bulk_commit();
                }
                flushed += n;
                parity ^= 1u;
                mine = stage + parity * stage_elems + my_col;
            }
        };

        // Two Philox blocks = six Euler steps = one stage per trip (software-pipelined like the pricing kernel)
        const uint32_t trips = c.steps / kStage, rest = c.steps - kStage * trips;
        uint4 cur = rng.block(0u, keys);
        uint64_t m1b = 0;
        for (uint32_t t = 0; t < trips; ++t) {
#pragma unroll
            for (int h = 0; h < kStage / 3; ++h) {
                StepDraw d0, d1, d2;
                m1b += kPhiloxM1;
                const uint4 nxt = rng.block_from_product(m1b, keys);
                split_draws(cur, one_bits, d0, d1, d2);
                euler_pair(x1, u1, x2, u2, d0.t_radius, d0.t_angle, k);
                emit(3 * h + 0);
                euler_pair(x1, u1, x2, u2, d1.t_radius, d1.t_angle, k);
                emit(3 * h + 1);
                euler_pair(x1, u1, x2, u2, d2.t_radius, d2.t_angle, k);
                emit(3 * h + 2);
                cur = nxt;
            }
            flush(kStage);
        }
        if (rest) {                                         // 1..5 steps left: uniform across the block
            uint32_t done = 0;
            for (int h = 0; h < kStage / 3 && done < rest; ++h) {
                StepDraw d[3];
                m1b += kPhiloxM1;
                const uint4 nxt = rng.block_from_product(m1b, keys);
                split_draws(cur, one_bits, d[0], d[1], d[2]);
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (done < rest) {
                        euler_pair(x1, u1, x2, u2, d[j].t_radius, d[j].t_angle, k);
                        emit(3 * h + j);
                        ++done;
                    }
                cur = nxt;
            }
            flush(rest);
        }

        // ---- terminal values and payoffs (as the pricing kernel) ----
        const real xT1 = x1 + x_shift, xT2 = x2 + x_shift;
        if (x_out != nullptr) {
            const real vT1 = u1 == k.floor_u ? c.floor_v : u1 * k.inv_c1;
            const real vT2 = u2 == k.floor_u ? c.floor_v : u2 * k.inv_c1;
            if (active) { x_out[local] = xT1; v_out[local] = vT1; }
            if (paired) { x_out[c.pair_count + local] = xT2; v_out[c.pair_count + local] = vT2; }
        }
        acc.add(exp_real(xT1), exp_real(xT2), active, paired, c.strikes[0], s_strikes, lane, c.payoff_sign);
    }
    if constexpr (BULK) {
        if (threadIdx.x == 0) bulk_wait_all();             // the stage must outlive the copies that read it
    }
    if constexpr (kTrajThreads > kFusedMaxThreads) {       // wide block: the (now idle) stage holds the per-warp rows
        static_assert(BULK && sizeof(real) * 2 * kStage * NARR * kTrajThreads >= sizeof(double) * (kTrajThreads / 32) * kMaxAcc,
                      "the stage is large enough for the reduction rows");
        __syncthreads();                                   // after thread 0's wait: nobody overwrites a stage in flight
        acc.template publish<true>(c.nK, ws, reinterpret_cast<double (*)[kMaxAcc]>(smem_raw));
    } else {
        acc.template publish<true>(c.nK, ws);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------

template <typename real>
static const void *trajectory_fn(bool with_v, bool multi, bool bulk) {
#define CB_TRAJ(V, M, K) (const void *)heston_trajectory_kernel<real, V, M, K>
    if (with_v) {
        if (multi) return bulk ? CB_TRAJ(true, true, true) : CB_TRAJ(true, true, false);
        return bulk ? CB_TRAJ(true, false, true) : CB_TRAJ(true, false, false);
    }
    if (multi) return bulk ? CB_TRAJ(false, true, true) : CB_TRAJ(false, true, false);
    return bulk ? CB_TRAJ(false, false, true) : CB_TRAJ(false, false, false);
#undef CB_TRAJ
}

template <typename real>
static int trajectory_threads(bool with_v, bool bulk) {
    return with_v ? (bulk ? kTrajThreadsOf<real, true, true> : kTrajThreadsOf<real, true, false>)
                  : (bulk ? kTrajThreadsOf<real, false, true> : kTrajThreadsOf<real, false, false>);
}

template <typename real>
static size_t trajectory_smem_bytes(bool with_v, bool bulk) {
    return bulk ? (size_t)2 * kStage * (with_v ? 4 : 2) * trajectory_threads<real>(with_v, bulk) * sizeof(real) : 0;
}

// cuTensorMapEncodeTiled through the runtime (the library links no libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// one half ([rows][pair_count] columns starting at `base`, row pitch = the full matrix row) as a 2-D tensor with a
// [kStage][kBox] box
template <typename real>
static cudaError_t encode_half(CUtensorMap *map, real *base, uint64_t pair_count, uint64_t rows) {
    EncodeTiledFn encode = encode_tiled();
    if (encode == nullptr) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {pair_count, rows};
    const cuuint64_t strides[1] = {2 * pair_count * sizeof(real)};
    const cuuint32_t box[2] = {(cuuint32_t)kBox, (cuuint32_t)kStage};
    const cuuint32_t elem[2] = {1, 1};
    const CUtensorMapDataType type = sizeof(real) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const CUresult r = encode(map, type, 2, base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <typename real>
bool trajectory_bulk_eligible(const FusedConsts<real> &c, const real *d_traj_x, const real *d_traj_v) {
    const bool aligned = (reinterpret_cast<uintptr_t>(d_traj_x) % 16 == 0) &&
                         (d_traj_v == nullptr || reinterpret_cast<uintptr_t>(d_traj_v) % 16 == 0);
    const bool rows_cut = (c.pair_count * sizeof(real)) % 16 == 0;
    const bool all_partnered = c.pair_first + c.pair_count <= c.partnered;   // no lone path of an odd R in this shard
    const bool coordinates_fit = c.pair_count < (1ull << 31);               // tensor coordinates are 32-bit signed
    return aligned && rows_cut && all_partnered && coordinates_fit && c.pair_count > 0 && encode_tiled() != nullptr;
}

template <typename real>
cudaError_t trajectory_occupancy(bool with_v, bool multi, bool bulk, int *threads, int *blocks_per_sm) {
    const void *fn = trajectory_fn<real>(with_v, multi, bulk);
    const size_t smem = trajectory_smem_bytes<real>(with_v, bulk);
    *threads = trajectory_threads<real>(with_v, bulk);
    if (smem > 32 * 1024) {   // static shared memory (strikes, reduction scratch) comes on top of the stage
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, *threads, smem);
}

template <typename real>
cudaError_t launch_heston_trajectory(const FusedConsts<real> &c, int blocks, bool bulk, ReduceWorkspace ws,
                                     real *d_x, real *d_v, real *d_traj_x, real *d_traj_v, cudaStream_t stream) {
    const bool with_v = d_traj_v != nullptr;
    const void *fn = trajectory_fn<real>(with_v, c.nK > 1, bulk);
    const size_t smem = trajectory_smem_bytes<real>(with_v, bulk);
    if (smem > 32 * 1024) {   // static shared memory (strikes, reduction scratch) comes on top of the stage
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    TrajMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (bulk) {
        const uint64_t rows = (uint64_t)c.steps + 1;
        cudaError_t e = encode_half(&maps.m[0], d_traj_x, c.pair_count, rows);
        if (e == cudaSuccess) e = encode_half(&maps.m[1], d_traj_x + c.pair_count, c.pair_count, rows);
        if (e == cudaSuccess && with_v) e = encode_half(&maps.m[2], d_traj_v, c.pair_count, rows);
        if (e == cudaSuccess && with_v) e = encode_half(&maps.m[3], d_traj_v + c.pair_count, c.pair_count, rows);
        if (e != cudaSuccess) return e;
    }
    void *args[] = {(void *)&c,   (void *)&maps,     (void *)&ws,      (void *)&d_x,
                    (void *)&d_v, (void *)&d_traj_x, (void *)&d_traj_v};
    return cudaLaunchKernel(fn, dim3(blocks), dim3(trajectory_threads<real>(with_v, bulk)), args, smem, stream);
}

#define CB_INSTANTIATE(real)                                                                                       \
    template bool trajectory_bulk_eligible<real>(const FusedConsts<real> &, const real *, const real *);           \
    template cudaError_t trajectory_occupancy<real>(bool, bool, bool, int *, int *);                               \
    template cudaError_t launch_heston_trajectory<real>(const FusedConsts<real> &, int, bool, ReduceWorkspace,     \
                                                        real *, real *, real *, real *, cudaStream_t);
CB_INSTANTIATE(float)
CB_INSTANTIATE(double)

}  // namespace cb
