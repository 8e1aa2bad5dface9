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
#include <cstdio>
#include "pf_kernels.cuh"

namespace sgm {

constexpr int COOP_MAX_Q = NT;            // tiles per item (N <= 65536: one tile per header thread, as in pf_header_kernel<..., 256>)

#ifndef SGM_COOP_MIN_CTAS
#define SGM_COOP_MIN_CTAS 2
#endif
// CL = true: the CTAs of ONE item form a thread-block cluster (G <= 16 CTAs, i.e. N <= 32768) and the barrier between two time
// steps is the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) instead of the grid barrier's atomic +
// polling round trips through L2.  Items are independent clusters: no cooperative launch, no limit on the batch size.
template <class R, class Model, bool SORTED, int FM, bool RAGGED, bool CL = false>
__global__ void __launch_bounds__(NT, SGM_COOP_MIN_CTAS) pf_coop_kernel(KArgs a) {
    namespace cgr = cooperative_groups;
    __shared__ __align__(32) R s_cdf_all[NWARP][SORTED ? WIN_BYTES / sizeof(R) : WT];
    __shared__ double sh_d[NWARP];
    extern __shared__ double s_hdr[];                 // hdr_stride(Q) doubles
    const int b = a.b0 + blockIdx.y, g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    init_body<R, Model>(a, b, g, s_cdf_all[warp]);
    // Off the critical path of a step: T_buf is loaded once; the log-likelihood accumulator of the item lives in shared memory
    // (CTA 0 of the item); the header's Gamma draws of the NEXT step and their block scan do not depend on this step's
    // weights and are made between the arrive and the wait of the grid barrier (counter-based: the same values)
    __shared__ double s_acc[ACC_STRIDE];
    if (threadIdx.x < ACC_STRIDE) s_acc[threadIdx.x] = 0.0;
    HeaderCoop hc;
    hc.Tb = a.T_buf[b]; hc.scanned = false; hc.gv[0] = hc.gv[1] = hc.grun = hc.gtot = 0.0; hc.acc = s_acc;
    const bool pre = a.Q >= 32 && (FM != FM_GENERIC || (SORTED && uses_spacings(a)));
    if (pre) header_gamma_scan<NT>(a, b, 0, sh_d, hc);
    auto barrier_arrive = [&]() -> unsigned {
        if constexpr (CL) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); return 0u; }
        else return cgr::this_grid().barrier_arrive();
    };
    auto barrier_wait = [&](unsigned token) {
        if constexpr (CL) { (void)token; asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
        else cgr::this_grid().barrier_wait(std::move(token));
    };
    // (drawing the warp tile's own variates of the next step there as well was tried and is slower: 0.47 -> 0.50 ms at N = 2^16 --
    // 16 more live registers across the header push the kernel to 128 registers with spills, and the wait is not idle enough)
    barrier_wait(barrier_arrive());
#ifdef SGM_COOP_TIMING
    long long tc[5] = {0, 0, 0, 0, 0}, c0 = clock64(), c1;
#define SGM_TICK(i) { c1 = clock64(); tc[i] += c1 - c0; c0 = c1; }
#else
#define SGM_TICK(i)
#endif
    for (int t = 0; t < a.max_T; ++t) {
        header_body<R, Model, NT>(a, b, t, 0, sh_d, s_hdr, g == 0, &hc);
        __syncthreads();
        SGM_TICK(0)
        step_body<R, Model, SORTED, FM, RAGGED, WIN_BYTES>(a, b, t, g * NWARP + warp, lane, s_cdf_all[warp], s_hdr);
        SGM_TICK(1)
        const unsigned token = barrier_arrive();
        SGM_TICK(2)
        if (pre && t + 1 < a.max_T) header_gamma_scan<NT>(a, b, t + 1, sh_d, hc);
        SGM_TICK(3)
        barrier_wait(token);
        SGM_TICK(4)
    }
#ifdef SGM_COOP_TIMING
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0)
        printf("coop timing cta %d (cycles/step): header %lld  step %lld  arrive %lld  gamma %lld  wait %lld\n", (int)blockIdx.x,
               tc[0] / a.max_T, tc[1] / a.max_T, tc[2] / a.max_T, tc[3] / a.max_T, tc[4] / a.max_T);
#endif
    if (g == 0) header_body<R, Model, NT>(a, b, a.max_T, 1, sh_d, nullptr, true, &hc);
}

}  // namespace sgm
