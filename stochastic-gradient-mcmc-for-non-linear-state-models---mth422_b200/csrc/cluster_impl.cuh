// Host side of the thread-block-cluster kernels (cluster_kernels.cuh): shape / cluster-size choice and launches.
// Instantiated per arithmetic type in sgmpf_cl32.cu / sgmpf_cl64.cu (own translation units: they compile in parallel
// with the tile / shared-memory kernels).
#pragma once
#include "host_common.cuh"
#include "cluster_kernels.cuh"

namespace sgmhost {

struct ClusterPlan { int C, nl; };

// Smallest per-CTA share NL (= most SMs per item) such that a cluster of C <= 8 CTAs holds the item and all B clusters
// are resident at once (one wave: the regime where a gradient is latency-, not throughput-bound).
// Measured (SVM f32, one item, 60 steps, ms per gradient; cluster / one CTA / tile kernels):
//   N = 1000 (4 x 256): 0.177 / 0.163 / 0.71      N = 2048 (8 x 256): 0.218 / 0.338 / 0.71      N = 4096 (8 x 512): 0.476 / - / 0.65
//   N = 8192 (8 x 1024): 1.27 / - / 0.75 (0.55 as a CUDA graph)      N = 16384 (8 x 2048): 1.60 / - / 0.75
// The two random DSMEM reads per child (fine-CDF group, parent record) do not coalesce and cost ~10-16 cycles of the SM's
// DSMEM port EACH, i.e. time grows with the particles per CTA (with barrier.cluster's release / acquire -- MEMBAR.ALL.GPU +
// CCTL.IVALL in SASS -- replaced by relaxed arrive / wait the N = 1000 gradient is still 0.156 ms: the fences are ~12 %).  So
// AUTO takes the cluster kernel where the item is spread thin (256 particles per CTA) and one SM is not enough:
// 1024 < N <= 2048; from N = 4096 on the cooperative single-launch form of the tile kernels (coop_kernels.cuh) is faster
// (N = 4096, one item: 0.41 ms against 0.48 ms).  `forced` (SGM_PATH_CLUSTER) takes any plan.
inline bool cluster_plan(int N, int B, ClusterPlan& p, bool forced = true) {
    // path = auto no longer selects this kernel: after the sampled-CDF search of the shared-memory kernel one CTA (1024 threads x 2
    // particles) runs N = 2048 in 0.190 ms against 0.215 ms for the cluster of 8 x 256 (1-8 items; profiles/probe_r03s_latency.json)
    if (N <= 256 || !forced) return false;
    static const int NLS[4] = {256, 512, 1024, 2048};
    for (int k = 0; k < (forced ? 4 : 1); ++k) {
        int C = 2;
        while (C * NLS[k] < N) C *= 2;
        if (C <= 8 && (int64_t)B * C <= 148) { p.C = C; p.nl = NLS[k]; return true; }
    }
    return false;
}

template <class K, class... Args>
bool launch_cluster(K kern, int B, int C, int nth, size_t bytes, cudaStream_t stream, Args... args) {
    if (bytes > 227 * 1024) return false;
    if (bytes > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return false; }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(B * C); cfg.blockDim = dim3(nth); cfg.dynamicSmemBytes = bytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, args...) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}

template <class R, class Model, bool FAST>
bool launch_cluster_pf(const KArgs& a, const ClusterPlan& p, cudaStream_t s) {
    const int nth = p.nl == 2048 ? 1024 : p.nl;
    const size_t bytes = cluster_smem_bytes<R>(p.nl, p.C, nth, Model::NX, Model::NP);
    switch (p.nl) {
        case 256: return launch_cluster(pf_cluster_kernel<R, Model, 256, 1, FAST>, a.B, p.C, 256, bytes, s, a);
        case 512: return launch_cluster(pf_cluster_kernel<R, Model, 512, 1, FAST>, a.B, p.C, 512, bytes, s, a);
        case 1024: return launch_cluster(pf_cluster_kernel<R, Model, 1024, 1, FAST>, a.B, p.C, 1024, bytes, s, a);
        default: return launch_cluster(pf_cluster_kernel<R, Model, 1024, 2, FAST>, a.B, p.C, 1024, bytes, s, a);
    }
}
template <class R, class Model>
bool launch_cluster_model(const KArgs& a, const ClusterPlan& p, cudaStream_t s) {
    return small_fast_config(a) ? launch_cluster_pf<R, Model, true>(a, p, s) : launch_cluster_pf<R, Model, false>(a, p, s);
}

// true = launched (the whole time loop of every item of the batch in one cluster launch)
template <class R>
bool run_cluster(const sgm_pf_desc* d, const KArgs& a, cudaStream_t s) {
    ClusterPlan p;
    if (!cluster_plan(a.N, a.B, p, d->path == SGM_PATH_CLUSTER)) return false;
    switch (d->model) {
        case SGM_MODEL_SVM: return launch_cluster_model<R, SvmPrior>(a, p, s);
        case SGM_MODEL_LGSSM: return d->kernel == SGM_KERNEL_PRIOR ? launch_cluster_model<R, LgssmPrior>(a, p, s) : launch_cluster_model<R, LgssmOptimal>(a, p, s);
        default: return d->kernel == SGM_KERNEL_PRIOR ? launch_cluster_model<R, GarchPrior>(a, p, s) : launch_cluster_model<R, GarchOptimal>(a, p, s);
    }
}

// ---- persistent SG-MCMC kernel, one cluster per chain (production configuration only) ---------------------------------
template <class R, class Model, int NTH, int PPT>
__global__ void __launch_bounds__(NTH, 1) sgld_cluster_kernel(SgldArgs sa, KArgs a, int K) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int c = (int)(blockIdx.x / C);
    const uint64_t o0 = *sa.offset_dev;
    const int64_t it0 = *sa.iter_dev;
    a.offset_dev = nullptr;
    const uint32_t k1 = a.key.k1;
    for (int k = 0; k < K; ++k) {
        const uint64_t o = o0 + (uint64_t)k;
        if (rank == 0) sgld_prepare_item(sa, c, k, o, (int)threadIdx.x, (int)blockDim.x);
        cluster.sync();                                  // the item arrays (global memory) are visible to every CTA
        a.key.offset = (uint32_t)(o & 0xffffffffu);
        a.key.k1 = k1 ^ (uint32_t)(o >> 32);
        cluster_pf_item<R, Model, NTH, PPT, true>(a, c, small_smem);      // ends with a cluster barrier
        if (rank == 0 && threadIdx.x == 0) sgld_update_chain(sa, c, k, o, it0 + k);
    }
}

template <class R, class Model>
bool launch_sgld_cluster_model(const SgldArgs& sa, const KArgs& a, const ClusterPlan& p, int K, cudaStream_t s) {
    const int nth = p.nl == 2048 ? 1024 : p.nl;
    const size_t bytes = cluster_smem_bytes<R>(p.nl, p.C, nth, Model::NX, Model::NP);
    switch (p.nl) {
        case 256: return launch_cluster(sgld_cluster_kernel<R, Model, 256, 1>, a.B, p.C, 256, bytes, s, sa, a, K);
        case 512: return launch_cluster(sgld_cluster_kernel<R, Model, 512, 1>, a.B, p.C, 512, bytes, s, sa, a, K);
        case 1024: return launch_cluster(sgld_cluster_kernel<R, Model, 1024, 1>, a.B, p.C, 1024, bytes, s, sa, a, K);
        default: return launch_cluster(sgld_cluster_kernel<R, Model, 1024, 2>, a.B, p.C, 1024, bytes, s, sa, a, K);
    }
}

// 1 = launched, 0 = not eligible
template <class R>
int run_sgld_cluster(const sgm_pf_desc* d, const sgm::SgldArgs& sa, const KArgs& a, int K, cudaStream_t s) {
    ClusterPlan p;
    if (!small_fast_config(a) || !cluster_plan(a.N, a.B, p, d->path == SGM_PATH_CLUSTER)) return 0;
    bool ok;
    switch (d->model) {
        case SGM_MODEL_SVM: ok = launch_sgld_cluster_model<R, SvmPrior>(sa, a, p, K, s); break;
        case SGM_MODEL_LGSSM: ok = d->kernel == SGM_KERNEL_PRIOR ? launch_sgld_cluster_model<R, LgssmPrior>(sa, a, p, K, s)
                                                                  : launch_sgld_cluster_model<R, LgssmOptimal>(sa, a, p, K, s); break;
        default: ok = d->kernel == SGM_KERNEL_PRIOR ? launch_sgld_cluster_model<R, GarchPrior>(sa, a, p, K, s)
                                                     : launch_sgld_cluster_model<R, GarchOptimal>(sa, a, p, K, s);
    }
    return ok ? 1 : 0;
}

}  // namespace sgmhost
