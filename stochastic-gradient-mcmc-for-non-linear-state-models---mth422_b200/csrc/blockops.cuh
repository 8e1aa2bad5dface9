// Warp-shuffle / block-level reductions and scans with a FIXED combination tree (deterministic).
// All block primitives assume blockDim.x == 32 * NW (default NT = 256) and are called by every thread of the block.
#pragma once
#include <cuda_runtime.h>

namespace sgm {

constexpr int NT = 256;            // threads per CTA
constexpr int NWARP = NT / 32;
constexpr unsigned FULL = 0xffffffffu;

template <class T> __device__ __forceinline__ T warp_incl_scan(T v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T n = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += n;
    }
    return v;
}
// the same scan when only lanes < NL hold data (zeros elsewhere): levels d >= NL add nothing to lanes < NL and are skipped,
// so lanes < NL get bit-identical results; lanes >= NL are NOT valid
template <int NL, class T> __device__ __forceinline__ T warp_incl_scan_low(T v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < NL; d <<= 1) {
        T n = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += n;
    }
    return v;
}
template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
// NaN handling: fmax drops NaNs, which is fine here -- a NaN (or +inf) log-weight still poisons the tile
// sum through exp(lw - m), and the degenerate-weight status flag is raised from the sum.
__device__ __forceinline__ float nan_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double nan_max(double a, double b) { return fmax(a, b); }
template <class T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = nan_max(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
// f32: sm_100a reduces a float max over the warp in ONE instruction (CREDUX.MAX.F32, NaNs dropped like fmaxf) instead of five
// dependent shuffle + FMNMX pairs.  Called by all 32 lanes, converged.
template <> __device__ __forceinline__ float warp_max<float>(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// Block max.  `sh` needs NW elements.  Result valid in all threads.
template <int NW = NWARP, class T> __device__ __forceinline__ T block_max(T v, T* sh) {
    v = warp_max(v);
    __syncthreads();                      // protect sh from a previous use
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = sh[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r = nan_max(r, sh[w]);
    return r;
}
// Block sum, fixed order (warp tree, then warps 0..7 left to right).
template <int NW = NWARP, class T> __device__ __forceinline__ T block_sum(T v, T* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = sh[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r += sh[w];
    return r;
}
// Block exclusive scan of one value per thread (thread order); returns exclusive prefix, sets total.
template <int NW = NWARP, class T> __device__ __forceinline__ T block_excl_scan(T v, T* sh, T& total) {
    T incl = warp_incl_scan(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = incl;
    __syncthreads();
    T base = T(0), tot = T(0);
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        T s = sh[k];
        if (k < w) base += s;
        tot += s;
    }
    total = tot;
    return base + incl - v;
}

}  // namespace sgm
