// Inverse-multiquadric kernel Stein discrepancy of a parameter trace (trace_metric_functions.py:20-81):
//   KSD = sqrt( sum_{i,j} k0(x_i, x_j) ) / K,   k = (c^2 + |x - y|^2)^(-beta),
//   k0 = g_i.g_j k + grad_x k . g_j + grad_y k . g_i + tr(grad_x grad_y k)
// an all-pairs O(K^2 d) double sum -- the reference evaluates it with (block^2, d) numpy temporaries.  Here: thread i
// keeps (x_i, g_i) in registers, tiles of the j side go through shared memory, f64 throughout (the terms cancel),
// one partial sum per CTA (fixed order), summed on the host.
#pragma once
#include "blockops.cuh"

namespace sgm {

constexpr int KSD_MAXD = 8;
constexpr int KSD_JT = 256;

__global__ void __launch_bounds__(NT) ksd_imq_kernel(const double* __restrict__ x, const double* __restrict__ g, int K, int d,
                                                     double c2, double beta, double* __restrict__ partial) {
    __shared__ double s_x[KSD_JT][KSD_MAXD];
    __shared__ double s_g[KSD_JT][KSD_MAXD];
    __shared__ double sh_d[NWARP];
    const int i = blockIdx.x * NT + threadIdx.x;
    double xi[KSD_MAXD], gi[KSD_MAXD];
    for (int q = 0; q < KSD_MAXD; ++q) { xi[q] = (i < K && q < d) ? x[(size_t)i * d + q] : 0.0; gi[q] = (i < K && q < d) ? g[(size_t)i * d + q] : 0.0; }
    double acc = 0.0;
    for (int j0 = 0; j0 < K; j0 += KSD_JT) {
        __syncthreads();
        for (int e = threadIdx.x; e < KSD_JT * KSD_MAXD; e += NT) {
            const int j = e / KSD_MAXD, q = e % KSD_MAXD;
            const bool ok = (j0 + j < K) && (q < d);
            s_x[j][q] = ok ? x[(size_t)(j0 + j) * d + q] : 0.0;
            s_g[j][q] = ok ? g[(size_t)(j0 + j) * d + q] : 0.0;
        }
        __syncthreads();
        const int jn = min(KSD_JT, K - j0);
        if (i < K) {
            for (int j = 0; j < jn; ++j) {
                double diff2 = 0.0, gg = 0.0, gid = 0.0, gjd = 0.0;       // diff = x_i - x_j
#pragma unroll
                for (int q = 0; q < KSD_MAXD; ++q) {
                    const double dq = xi[q] - s_x[j][q];
                    diff2 += dq * dq; gg += gi[q] * s_g[j][q]; gid += gi[q] * dq; gjd += s_g[j][q] * dq;
                }
                const double base = diff2 + c2;
                const double bb = pow(base, -beta);
                const double coeff = -2.0 * beta * bb / base;
                // trace_metric_functions.py:66-74 with (index0, index1) = (i, j)
                acc += gg * bb + (-gid) * coeff + gjd * coeff + (-(double)d + 2.0 * (beta + 1.0) * diff2 / base) * coeff;
            }
        }
    }
    const double tot = block_sum(acc, sh_d);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

}  // namespace sgm
