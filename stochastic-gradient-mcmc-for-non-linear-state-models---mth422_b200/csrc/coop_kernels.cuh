// Latency regime of the TILE kernels (N > 4096, few items): the whole time loop in ONE cooperative launch.
// The per-step path issues 2 * T + 2 tiny launches (header + step per time step; as a replayed CUDA graph still ~5 us per
// launch pair: one item, N = 2^16: 0.62 ms per 60-step gradient).  Here all CTAs of the batch are co-resident
// (cudaLaunchCooperativeKernel) and a grid-wide barrier (cooperative_groups grid.sync(): release / acquire at GPU scope incl.
// the L1 invalidation the double-buffered particle arrays need) replaces the kernel boundary; every CTA of an item builds
// its OWN copy of the item header in shared memory from the tile summaries (same code as pf_header_kernel, so the result is
// bit-identical everywhere), which removes the second barrier per step; CTA 0 of the item applies the header's side effects.
// The arithmetic is that of the per-step kernels (same init_body / header_body / step_body): results are bit-identical.
#pragma once
#include <cooperative_groups.h>
#include <utility>
#include "pf_kernels.cuh"

namespace sgm {

constexpr int COOP_MAX_Q = NT;            // tiles per item (N <= 65536: one tile per header thread, as in pf_header_kernel<..., 256>)

template <class R, class Model, bool SORTED, int FM, bool RAGGED>
__global__ void __launch_bounds__(NT, 2) pf_coop_kernel(KArgs a) {
    namespace cgr = cooperative_groups;
    cgr::grid_group grid = cgr::this_grid();
    __shared__ __align__(32) R s_cdf_all[NWARP][SORTED ? WIN_BYTES / sizeof(R) : WT];
    __shared__ double sh_d[NWARP];
    extern __shared__ double s_hdr[];                 // hdr_stride(Q) doubles
    const int b = a.b0 + blockIdx.y, g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    init_body<R, Model>(a, b, g, s_cdf_all[warp]);
    // the header's Gamma draws of the NEXT step do not depend on this step's weights: they are drawn between the arrive and
    // the wait of the grid barrier (same values as header_body would draw: counter-based), off the step's critical path
    const bool pre = a.Q >= 32 && (FM != FM_GENERIC || (SORTED && uses_spacings(a)));
    double gv[2] = {0.0, 0.0};
    if (pre) header_gammas<NT>(a, b, 0, gv);
    grid.sync();
    for (int t = 0; t < a.max_T; ++t) {
        header_body<R, Model, NT>(a, b, t, 0, sh_d, s_hdr, g == 0, pre ? gv : nullptr);
        __syncthreads();
        step_body<R, Model, SORTED, FM, RAGGED, WIN_BYTES>(a, b, t, g * NWARP + warp, lane, s_cdf_all[warp], s_hdr);
        auto token = grid.barrier_arrive();
        if (pre && t + 1 < a.max_T) header_gammas<NT>(a, b, t + 1, gv);
        grid.barrier_wait(std::move(token));
    }
    if (g == 0) header_body<R, Model, NT>(a, b, a.max_T, 1, sh_d);
}

}  // namespace sgm
