// Storage plumbing of the device array handle (cocos_b200/numerics/_array.py): element-type conversion and a 2-D
// transpose.  Stands in for what the reference gets from ArrayFire behind ndarray.astype / ndarray.T
// (cocos/numerics/_array.py:150-803); everything else an array could do is fused into the hot-path kernels.
// Both are plain HBM-bound passes: coalesced reads and writes, grid-stride loops, a 32x33 shared-memory tile for the
// transpose (no bank conflicts).
#include "kernels.h"

namespace cb {

template <typename dst_t, typename src_t>
__global__ void __launch_bounds__(256)
cast_kernel(dst_t *__restrict__ dst, const src_t *__restrict__ src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (dst_t)src[i];
}

template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(T *__restrict__ dst, const T *__restrict__ src, uint64_t rows, uint64_t cols) {
    __shared__ T tile[32][33];
    const uint64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8 threads
    for (uint64_t t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
        const uint64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
        for (uint32_t j = ty; j < 32; j += 8)
            if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = src[(r0 + j) * cols + c0 + tx];
        __syncthreads();
        for (uint32_t j = ty; j < 32; j += 8)
            if (c0 + j < cols && r0 + tx < rows) dst[(c0 + j) * rows + r0 + tx] = tile[tx][j];
        __syncthreads();
    }
}

// exp / log of a whole array: the one elementwise step between the simulation and the density estimate in the
// reference's workflow (terminal prices S_T = exp(x_T), notebooks/Getting Started.ipynb cells 65-83)
template <typename T, int OP>
__global__ void __launch_bounds__(256)
unary_kernel(T *__restrict__ dst, const T *__restrict__ src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = OP == 0 ? exp(src[i]) : log(src[i]);
}

static int grid_for(uint64_t items, int per_block) {
    uint64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    const uint64_t cap = 148ull * 16;
    return (int)(b > cap ? cap : b);
}

// kinds: 0 float32, 1 float64, 2 int32, 3 int64, 4 uint32
template <typename dst_t>
static cudaError_t cast_from(dst_t *dst, const void *src, int src_kind, uint64_t n, cudaStream_t s) {
    const int g = grid_for(n, 256);
    switch (src_kind) {
        case 0: cast_kernel<dst_t, float><<<g, 256, 0, s>>>(dst, (const float *)src, n); break;
        case 1: cast_kernel<dst_t, double><<<g, 256, 0, s>>>(dst, (const double *)src, n); break;
        case 2: cast_kernel<dst_t, int32_t><<<g, 256, 0, s>>>(dst, (const int32_t *)src, n); break;
        case 3: cast_kernel<dst_t, long long><<<g, 256, 0, s>>>(dst, (const long long *)src, n); break;
        case 4: cast_kernel<dst_t, uint32_t><<<g, 256, 0, s>>>(dst, (const uint32_t *)src, n); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_cast(void *dst, int dst_kind, const void *src, int src_kind, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    switch (dst_kind) {
        case 0: return cast_from<float>((float *)dst, src, src_kind, n, s);
        case 1: return cast_from<double>((double *)dst, src, src_kind, n, s);
        case 2: return cast_from<int32_t>((int32_t *)dst, src, src_kind, n, s);
        case 3: return cast_from<long long>((long long *)dst, src, src_kind, n, s);
        case 4: return cast_from<uint32_t>((uint32_t *)dst, src, src_kind, n, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_unary(void *dst, const void *src, uint64_t n, int op, int is64, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int g = grid_for(n, 256);
    if (op < 0 || op > 1) return cudaErrorInvalidValue;
    if (is64) {
        if (op == 0) unary_kernel<double, 0><<<g, 256, 0, s>>>((double *)dst, (const double *)src, n);
        else unary_kernel<double, 1><<<g, 256, 0, s>>>((double *)dst, (const double *)src, n);
    } else {
        if (op == 0) unary_kernel<float, 0><<<g, 256, 0, s>>>((float *)dst, (const float *)src, n);
        else unary_kernel<float, 1><<<g, 256, 0, s>>>((float *)dst, (const float *)src, n);
    }
    return cudaGetLastError();
}

cudaError_t launch_transpose(void *dst, const void *src, uint64_t rows, uint64_t cols, int elem_bytes, cudaStream_t s) {
    if (rows == 0 || cols == 0) return cudaSuccess;
    const int g = grid_for(((rows + 31) / 32) * ((cols + 31) / 32), 1);
    if (elem_bytes == 4) transpose_kernel<uint32_t><<<g, 256, 0, s>>>((uint32_t *)dst, (const uint32_t *)src, rows, cols);
    else if (elem_bytes == 8) transpose_kernel<unsigned long long><<<g, 256, 0, s>>>((unsigned long long *)dst, (const unsigned long long *)src, rows, cols);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace cb
