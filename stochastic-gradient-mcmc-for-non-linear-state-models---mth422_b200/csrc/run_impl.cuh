// Launch orchestration of one sgm_pf_run call for arithmetic type R (template definitions; instantiated explicitly
// in sgmpf_f32.cu and sgmpf_f64.cu).
#pragma once
#include <type_traits>
#include "host_common.cuh"
#include "sgld_kernels.cuh"
#include "coop_kernels.cuh"
#include <stdlib.h>

namespace sgmhost {

// N <= 2048: the shared-memory-resident kernel (small_kernels.cuh), one CTA per item.  Thread count / particles per
// thread by N; for N in (512, 1024] a latency-bound batch (at most two CTAs per SM) spreads an item over 1024 threads,
// a throughput-bound one packs two particles per thread so that four items share an SM.
#ifndef SGM_SMALL_MAX_N
#define SGM_SMALL_MAX_N 2048
#endif
// ... while the batch is small enough: the shared-memory kernel is bound by instruction issue and shared-memory
// wavefronts (random parent gathers), the tile kernels by HBM -- measured crossover (SVM / LGSSM f32, 60 steps):
// N = 256: never; N = 1000: ~1650 items (4096 items: 3.7 vs 2.8 ms); N = 2048: ~480 items (9.4 vs 4.4 ms).
// N in (256, 1024]: 512 threads x 2 particles for every batch size.  A/B against 1024 threads x 1 particle (LGSSM f32, 60
// steps): N = 1000, 1 item 0.171 vs 0.182 ms, 296 items 0.29 vs 0.36 ms, persistent SGLD 5.5e3 vs 5.2e3 it/s (one Philox
// call and one Box-Muller pair serve both particles of a thread; half as many warps meet at the two barriers).
#ifndef SGM_SMALL_MID_NTH            // thread shape of the shared-memory kernel for 256 < N <= 1024 (A/B: scripts/build_variant.sh)
#define SGM_SMALL_MID_NTH 512
#define SGM_SMALL_MID_PPT 2
#endif
#ifndef SGM_SMALL_LATENCY_ITEMS
#define SGM_SMALL_LATENCY_ITEMS 0
#endif
#ifndef SGM_CLSYNC_MAX_CLUSTER       // largest cluster (CTAs of one item) of that launch: 8 is the portable limit, 16 the hardware's
#define SGM_CLSYNC_MAX_CLUSTER 16
#endif
#ifndef SGM_CLSYNC_MAX_CTAS          // batches up to this many CTAs take the cluster-synchronised single launch of the tile kernels
#define SGM_CLSYNC_MAX_CTAS 592
#endif
#ifndef SGM_SMALL_MAX_PARTICLES
#define SGM_SMALL_MAX_PARTICLES 1200000
#endif
// number of SMs of the current device (batches of at most that many items take the latency instantiations)
inline int sm_count() {
    static thread_local int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    }
    return n;
}
template <class R, class Model, int NTH, int PPT, bool FAST, bool LAT = false>
bool launch_small_shape(const KArgs& a, cudaStream_t stream) {
    const size_t bytes = small_smem_bytes<R>(NTH * PPT, Model::NX, Model::NP);
    if (bytes > 227 * 1024) return false;
    auto kern = pf_small_kernel<R, Model, NTH, PPT, FAST, LAT>;
    if (bytes > 48 * 1024) {
        static thread_local size_t granted = 0;            // per instantiation (and host thread): raise the limit once
        if (granted < bytes) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return false; }
            granted = 227 * 1024;
        }
    }
    kern<<<a.B, NTH, bytes, stream>>>(a);
    return true;
}
template <class R, class Model>
bool launch_small(const KArgs& a, cudaStream_t stream) {
    if (small_fast_config(a)) {
        const bool lat = a.B <= sm_count();            // at most one item per SM: the latency instantiation (no register cap)
        if (a.N <= 256) return lat ? launch_small_shape<R, Model, 256, 1, true, true>(a, stream) : launch_small_shape<R, Model, 256, 1, true>(a, stream);
#if SGM_SMALL_LATENCY_ITEMS > 0
        if (a.N <= 1024 && a.B <= SGM_SMALL_LATENCY_ITEMS) return launch_small_shape<R, Model, 1024, 1, true>(a, stream);
#endif
        if (a.N <= 1024) return lat ? launch_small_shape<R, Model, SGM_SMALL_MID_NTH, SGM_SMALL_MID_PPT, true, true>(a, stream)
                                    : launch_small_shape<R, Model, SGM_SMALL_MID_NTH, SGM_SMALL_MID_PPT, true>(a, stream);
        // (1024 < N <= 2048 at 512 threads x 4 particles without the register cap: SVM +2 %, LGSSM -3 % -- not worth a shape)
        return launch_small_shape<R, Model, 1024, 2, true>(a, stream);
    }
    // every other configuration (injected randoms, Nemeth shrinkage, filter, traces, ...): flags read at run time; at most one
    // item per SM and N <= 1024: 512 threads x 2 particles without the 64-register cap of a 1024-thread CTA
    if (a.N <= 1024 && a.B <= sm_count()) return launch_small_shape<R, Model, 512, 2, false, true>(a, stream);
    if (a.N <= 1024) return launch_small_shape<R, Model, 1024, 1, false>(a, stream);
    return launch_small_shape<R, Model, 1024, 2, false>(a, stream);
}

// kernel arguments of one call (workspace carved up, descriptor fields copied)
inline int make_kargs(const sgm_pf_desc* d, KArgs& a) {
    const Layout L = make_layout(d);
    if (!d->workspace || d->workspace_bytes < L.total) return fail(SGM_ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)d->workspace & 255) != 0) return fail(SGM_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    char* ws = reinterpret_cast<char*>(d->workspace);
    memset(&a, 0, sizeof(a));
    a.B = d->n_items; a.N = d->n_particles; a.G = (a.N + TILE - 1) / TILE; a.Q = (a.N + WT - 1) / WT; a.max_T = d->max_T;
    a.pf = d->pf; a.rng_mode = d->rng_mode; a.resample = d->resample; a.stat_kind = d->stat_kind;
    a.Ntilde = d->Ntilde; a.accept_reject = d->accept_reject;
    const double l10 = log10((double)a.N / 10.0);
    // pf.py:284-285.  With device randoms the cap is only a cost knob (both the accept-reject and the exact sampler
    // draw from the backward kernel exactly), so the default is the kernel's own budget (PARIS_CAP proposals).
    a.max_ar = d->max_accept_reject >= 0 ? d->max_accept_reject
                                         : (d->rng_mode == SGM_RNG_PHILOX ? PARIS_CAP : (int)(100.0 * l10));
    a.manual_thresh = d->manual_sample_threshold >= 0 ? d->manual_sample_threshold : (int)(10.0 * l10);  // pf.py:286-287
    if (a.max_ar < 0) a.max_ar = 0;
    a.variates32 = (d->dtype == SGM_F64 && d->variates == SGM_VARIATES_F32) ? 1 : 0;
    a.need_lw = (backward_pf(d->pf) || d->out_lw || d->trace_lw || d->stat_kind == SGM_STAT_PRED) ? 1 : 0;
    a.pred_K = d->pred_steps_ahead; a.pred_per_horizon = d->pred_per_horizon; a.inj_pred = d->inj_pred;
    a.lambduh = d->lambduh;
    a.key.k0 = (uint32_t)(d->seed & 0xffffffffu); a.key.k1 = (uint32_t)(d->seed >> 32);
    a.key.item = (uint32_t)d->item_id_base;
    a.offset_dev = d->offset_dev;
    if (!d->offset_dev) { a.key.offset = (uint32_t)(d->offset & 0xffffffffu); a.key.k1 ^= (uint32_t)(d->offset >> 32); }
    a.obs = d->obs; a.obs_off = d->obs_off; a.T_buf = d->T_buf; a.t1 = d->t1; a.tL = d->tL;
    a.step_weights = d->step_weights; a.wts_off = d->step_weights ? d->wts_off : nullptr; a.theta = d->theta;
    a.prior_mean = d->prior_mean; a.prior_var = d->prior_var;
    a.inj_z0 = d->inj_z0; a.inj_u = d->inj_u; a.inj_z = d->inj_z; a.inj_extra = d->inj_extra; a.inj_extra_off = d->inj_extra_off;
    for (int k = 0; k < 2; ++k) {
        a.rec[k] = ws + L.rec[k]; a.tail[k] = ws + L.tail[k]; a.fine[k] = ws + L.fine[k]; a.lw[k] = ws + L.lw[k];
        a.sub[k] = reinterpret_cast<double*>(ws + L.sub[k]);
        if (d->pf == SGM_PF_PARIS) a.Llist[k] = reinterpret_cast<int32_t*>(ws + L.Llist[k]);
    }
    a.acc = reinterpret_cast<double*>(ws + L.acc);
    a.hdr = reinterpret_cast<double*>(ws + L.hdr);
    a.thc = ws + L.thc;
    a.yw = ws + L.yw;
    if (d->pf == SGM_PF_PARIS) {
        a.Jidx = reinterpret_cast<int32_t*>(ws + L.Jidx);
        a.counters = reinterpret_cast<int32_t*>(ws + L.counters);
        a.pcdf = reinterpret_cast<double*>(ws + L.pcdf);
        a.pguide = reinterpret_cast<int32_t*>(ws + L.pguide);
        a.pkey = ws + L.pkey;
    }
    if (d->pf == SGM_PF_POY_N2) { a.n2part = ws + L.n2part; a.n2_tensor = n2_use_tensor(d->dtype, d->n2_mode) ? 1 : 0; }
    a.grad = d->grad; a.loglik = d->loglik; a.status = d->status;
    a.out_x = d->out_x; a.out_lw = d->out_lw; a.out_stats = d->out_stats;
    a.trace_anc = d->trace_anc; a.trace_x = d->trace_x; a.trace_lw = d->trace_lw; a.trace_J = d->trace_J;

    return SGM_OK;
}

// ---- cooperative single-launch form of the tile kernels (coop_kernels.cuh) ------------------------------------------
template <class K>
bool launch_coop_kernel(K kern, const KArgs& a, cudaStream_t stream) {
    const size_t dyn = hdr_stride(a.Q) * sizeof(double);
    static thread_local int per_sm = -1;                 // per instantiation: co-resident CTAs per SM at the largest header
    if (per_sm < 0) {
        int n = 0;
        const size_t dyn_max = hdr_stride(COOP_MAX_Q) * sizeof(double);       // + 32 KB static: above the 48 KB default limit
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, NT, dyn_max) != cudaSuccess) { cudaGetLastError(); n = 0; }
        per_sm = n;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((int64_t)a.G * a.B > (int64_t)per_sm * sms) return false;
    KArgs args = a;
    void* params[1] = {&args};
    if (cudaLaunchCooperativeKernel((const void*)kern, dim3(a.G, a.B), dim3(NT), params, dyn, stream) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}
// the same kernel with one thread-block cluster per item (G <= 8 CTAs) and the cluster barrier between the steps
template <class K>
bool launch_clsync_kernel(K kern, const KArgs& a, cudaStream_t stream) {
    const size_t dyn = hdr_stride(a.Q) * sizeof(double);
    static thread_local bool attr_set = false;           // per instantiation
    if (!attr_set) {
        const size_t dyn_max = hdr_stride(COOP_MAX_Q) * sizeof(double);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_max) != cudaSuccess) { cudaGetLastError(); return false; }
        // clusters of 9-16 CTAs (16384 < N <= 32768) are a non-portable size: allowed on B200, the occupancy query below decides
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(a.G, a.B); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = dyn; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // every cluster of the batch must be resident at once: a cluster that has to wait for a free slot waits for a WHOLE time loop
    // (measured: 64 items x 5 CTAs = 320 CTAs against 296 slots ran 0.40 ms instead of the per-step path's 0.32 ms)
    static thread_local int max_clusters[17] = {0};                              // per instantiation, by cluster size
    if (max_clusters[a.G] == 0) {
        int n = 0;
        cudaLaunchConfig_t q = cfg;
        q.gridDim = dim3(a.G, 4096); q.dynamicSmemBytes = hdr_stride(COOP_MAX_Q) * sizeof(double);
        if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) { cudaGetLastError(); n = 0; }
        max_clusters[a.G] = n > 0 ? n : -1;
    }
    if (a.B > max_clusters[a.G]) return false;
    if (cudaLaunchKernelEx(&cfg, kern, a) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}
template <class R, class Model>
bool launch_clsync(const sgm_pf_desc* d, const KArgs& a, int fm, cudaStream_t stream) {
    if (fm == FM_POY) return (a.N % WT == 0) ? launch_clsync_kernel(pf_coop_kernel<R, Model, true, FM_POY, false, true>, a, stream)
                                             : launch_clsync_kernel(pf_coop_kernel<R, Model, true, FM_POY, true, true>, a, stream);
    if (d->resample == SGM_RESAMPLE_MULTINOMIAL) return launch_clsync_kernel(pf_coop_kernel<R, Model, false, FM_GENERIC, false, true>, a, stream);
    return launch_clsync_kernel(pf_coop_kernel<R, Model, true, FM_GENERIC, false, true>, a, stream);
}
template <class R, class Model>
bool launch_coop(const sgm_pf_desc* d, const KArgs& a, int fm, cudaStream_t stream) {
    if (fm == FM_POY) return (a.N % WT == 0) ? launch_coop_kernel(pf_coop_kernel<R, Model, true, FM_POY, false>, a, stream)
                                             : launch_coop_kernel(pf_coop_kernel<R, Model, true, FM_POY, true>, a, stream);
    if (d->resample == SGM_RESAMPLE_MULTINOMIAL) return launch_coop_kernel(pf_coop_kernel<R, Model, false, FM_GENERIC, false>, a, stream);
    return launch_coop_kernel(pf_coop_kernel<R, Model, true, FM_GENERIC, false>, a, stream);
}

template <class R, class Model>
int run_impl(const sgm_pf_desc* d, cudaStream_t stream) {
    KArgs a;
    { const int rc = make_kargs(d, a); if (rc != SGM_OK) return rc; }
    const dim3 grid(a.G, a.B), block(NT);
    int64_t launches = 0;
    const bool pred = d->stat_kind == SGM_STAT_PRED;
    const bool small_ok = a.N <= SGM_SMALL_MAX_N && !backward_pf(d->pf) && !pred && d->path != SGM_PATH_TILES && d->path != SGM_PATH_STEPS &&
                          (a.N <= 512 || (int64_t)a.B * a.N <= SGM_SMALL_MAX_PARTICLES || d->path == SGM_PATH_SMALL);
    // few items (all clusters resident at once): a cluster of CTAs per item, particle system in distributed shared memory
    const bool cluster_ok = !backward_pf(d->pf) && !pred && (d->path == SGM_PATH_AUTO || d->path == SGM_PATH_CLUSTER) &&
                            (a.N > 1024 || d->path == SGM_PATH_CLUSTER);
    bool done_small = false;
    if (cluster_ok || small_ok) {
        if (d->ev_steps_begin) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_begin), stream);
        if (cluster_ok) done_small = run_cluster<R>(d, a, stream);
        if (!done_small && small_ok) done_small = launch_small<R, Model>(a, stream);
        if (done_small) ++launches;
        if (d->ev_steps_end) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_end), stream);
    }
    if (done_small) {
        // results written by the kernel itself (incl. the optional exports)
        set_launch_count(launches);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
        return SGM_OK;
    } else {
        // Two-stream pipelining (O(N) smoothers, big batches): the halves alternate on `stream` / `aux_stream`, so
        // the one-CTA-per-item header kernel and the launch gap of one half overlap the step kernel of the other.
        const bool piped = d->aux_stream && d->ev_aux_fork && d->ev_aux_join && !backward_pf(d->pf) && !pred &&
                           (int64_t)a.B * a.G >= 2 * 148 * 4 && a.B >= 2;
        const int nh = piped ? 2 : 1;
        // the production configurations run the step kernel with their run-time flags folded to constants (FM_*)
        const bool fast_cfg = d->rng_mode == SGM_RNG_PHILOX && d->resample == SGM_RESAMPLE_MULTINOMIAL_SORTED &&
                              d->stat_kind == SGM_STAT_SCORE && !a.need_lw && !d->trace_anc && !d->trace_x && !d->trace_lw;
        const int fm = !fast_cfg ? FM_GENERIC
                                 : (d->pf == SGM_PF_NEMETH ? (d->lambduh == 1.0 ? FM_POY : FM_SHRINK)
                                                           : (d->pf == SGM_PF_FILTER ? FM_FILTER : FM_GENERIC));
        cudaStream_t sh[2] = {stream, piped ? reinterpret_cast<cudaStream_t>(d->aux_stream) : stream};
        KArgs ah[2] = {a, a};
        int nb[2] = {piped ? a.B / 2 : a.B, piped ? a.B - a.B / 2 : 0};
        ah[1].b0 = nb[0];
        if (piped) {
            cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_aux_fork), stream);
            cudaStreamWaitEvent(sh[1], reinterpret_cast<cudaEvent_t>(d->ev_aux_fork), 0);
        }
        // one CTA per item; 1024 threads once an item has more than one tile per header thread
        auto launch_header = [&](int h, int t, int final_pass) {
            if (a.Q > NT) pf_header_kernel<R, Model, 1024><<<nb[h], 1024, 0, sh[h]>>>(ah[h], t, final_pass);
            else pf_header_kernel<R, Model><<<nb[h], block, 0, sh[h]>>>(ah[h], t, final_pass);
            ++launches;
        };
        auto launch_fast = [&](auto fmc, int h, int t) {
            constexpr int FMV = decltype(fmc)::value;
            constexpr int FW = StepShape<R, true>::WARPS;
            const dim3 gf((a.Q + FW - 1) / FW, nb[h]);
            if (a.N % WT == 0) pf_step_kernel<R, Model, true, FMV, false><<<gf, 32 * FW, 0, sh[h]>>>(ah[h], t);
            else pf_step_kernel<R, Model, true, FMV, true><<<gf, 32 * FW, 0, sh[h]>>>(ah[h], t);   // ragged last tile
        };
        auto launch_step = [&](int h, int t) {
            constexpr int SW = StepShape<R, false>::WARPS;
            ++launches;
            if (fm == FM_POY) return launch_fast(std::integral_constant<int, FM_POY>{}, h, t);
            if (fm == FM_SHRINK) return launch_fast(std::integral_constant<int, FM_SHRINK>{}, h, t);
            if (fm == FM_FILTER) return launch_fast(std::integral_constant<int, FM_FILTER>{}, h, t);
            const dim3 gh((a.Q + SW - 1) / SW, nb[h]), bs(32 * SW);
            if (d->resample == SGM_RESAMPLE_MULTINOMIAL) pf_step_kernel<R, Model, false><<<gh, bs, 0, sh[h]>>>(ah[h], t);
            else pf_step_kernel<R, Model, true><<<gh, bs, 0, sh[h]>>>(ah[h], t);
        };
        // few items (every CTA of the batch resident at once, at most one CTA per SM): the whole time loop in ONE
        // cooperative launch, grid barrier instead of the kernel boundary; same arithmetic, bit-identical results
        static const bool no_coop_env = getenv("SGM_NO_COOP") != nullptr;        // A/B switch for the benches
        const bool no_coop = no_coop_env || d->path == SGM_PATH_STEPS;
        // (Q <= 256: above that the per-step path builds its headers with 1024 threads, i.e. another summation order)
        const bool coop_ok = !no_coop && !piped && !backward_pf(d->pf) && !pred && a.Q <= NT && a.max_T > 0 &&
                             (int64_t)a.G * a.B <= 148;
        bool done_coop = false;
        // N <= 32768 (an item = at most 16 CTAs = one cluster; 9-16 is a non-portable size) and a batch of at most SGM_CLSYNC_MAX_CTAS CTAs: one launch
        // with the cluster barrier between the steps (no co-residency requirement, so also for more CTAs than SMs)
        static const bool no_clsync = getenv("SGM_NO_CLSYNC") != nullptr;         // A/B switch
        const bool cl_ok = !no_coop && !no_clsync && !piped && !backward_pf(d->pf) && !pred && a.Q <= NT && a.max_T > 0 && a.G <= SGM_CLSYNC_MAX_CLUSTER &&
                           (int64_t)a.G * a.B <= SGM_CLSYNC_MAX_CTAS;
        if (cl_ok) {
            if (d->ev_steps_begin) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_begin), stream);
            done_coop = launch_clsync<R, Model>(d, a, fm == FM_POY ? FM_POY : FM_GENERIC, stream);
            if (done_coop) { ++launches; if (d->ev_steps_end) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_end), stream); }
        }
        if (!done_coop && coop_ok) {
            if (d->ev_steps_begin) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_begin), stream);
            done_coop = launch_coop<R, Model>(d, a, fm == FM_POY ? FM_POY : FM_GENERIC, stream);
            if (done_coop) { ++launches; if (d->ev_steps_end) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_end), stream); }
        }
        if (!done_coop) {
        for (int h = 0; h < nh; ++h) { pf_init_kernel<R, Model><<<dim3(a.G, nb[h]), block, 0, sh[h]>>>(ah[h]); ++launches; }
        if (d->ev_steps_begin) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_begin), stream);
        for (int t = 0; t < a.max_T; ++t) {
            for (int h = 0; h < nh; ++h) {
                launch_header(h, t, 0);
                launch_step(h, t);
            }
            if (pred) { pf_pred_kernel<R, Model><<<a.B, block, 0, stream>>>(a, t); ++launches; }
            if (d->pf == SGM_PF_POY_N2) launches += launch_poyiadjis_n2<R, Model>(a, t, stream);
            else if (d->pf == SGM_PF_PARIS) launches += launch_paris<R, Model>(a, t, stream);
        }
        for (int h = 0; h < nh; ++h) launch_header(h, a.max_T, 1);
        if (piped) {
            cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_aux_join), sh[1]);
            cudaStreamWaitEvent(stream, reinterpret_cast<cudaEvent_t>(d->ev_aux_join), 0);
        }
        if (d->ev_steps_end) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_end), stream);
        }
    }
    if (d->out_x || d->out_lw || d->out_stats) {
        pf_export_kernel<R, Model><<<dim3((a.N + NT - 1) / NT, a.B), block, 0, stream>>>(a); ++launches;
    }
    set_launch_count(launches);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
    return SGM_OK;
}

// ---- persistent SG-MCMC kernel: K whole iterations of a chain inside ONE launch ---------------------------------
// One CTA per chain (one work item per chain and iteration: minibatch 1, one sequence, N <= 2048): thread 0 draws the
// window and rebuilds theta (sgld_prepare_item), the CTA runs the shared-memory particle filter (small_pf_item), thread
// 0 applies the SG-MCMC update (sgld_update_chain) -- no kernel boundary, no host, nothing but the chain's few scalars
// in global memory.  Iteration k uses Philox call offset (*offset_dev + k): exactly the numbers the launch-per-iteration
// path draws, so both paths give bit-identical chains.
template <class R, class Model, int NTH, int PPT, bool FAST, bool LAT = false>
__global__ void __launch_bounds__(NTH, small_min_ctas(NTH, LAT)) sgld_persistent_kernel(SgldArgs sa, KArgs a, int K) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int c = blockIdx.x;
    const uint64_t o0 = *sa.offset_dev;
    const int64_t it0 = *sa.iter_dev;
    const size_t B = (size_t)a.B, N = (size_t)a.N, T = (size_t)a.max_T;
    const double *z0 = a.inj_z0, *iu = a.inj_u, *iz = a.inj_z;
    a.offset_dev = nullptr;
    const uint32_t k1 = a.key.k1;
    for (int k = 0; k < K; ++k) {
        const uint64_t o = o0 + (uint64_t)k;
        sgld_prepare_item(sa, c, k, o, (int)threadIdx.x, (int)blockDim.x);      // all threads: the stores are shared out
        __syncthreads();
        a.key.offset = (uint32_t)(o & 0xffffffffu);
        a.key.k1 = k1 ^ (uint32_t)(o >> 32);
        if (z0) { a.inj_z0 = z0 + (size_t)k * B * N; a.inj_u = iu + (size_t)k * B * T * N; a.inj_z = iz + (size_t)k * B * T * N; }
        small_pf_item<R, Model, NTH, PPT, FAST>(a, c, small_smem);
        __syncthreads();
        if (threadIdx.x == 0) sgld_update_chain(sa, c, k, o, it0 + k);
        __syncthreads();
    }
}

template <class R, class Model, int NTH, int PPT, bool FAST, bool LAT = false>
bool launch_persistent_shape(const SgldArgs& sa, const KArgs& a, int K, cudaStream_t stream) {
    const size_t bytes = small_smem_bytes<R>(NTH * PPT, Model::NX, Model::NP);
    if (bytes > 227 * 1024) return false;
    auto kern = sgld_persistent_kernel<R, Model, NTH, PPT, FAST, LAT>;
    if (bytes > 48 * 1024) {
        static thread_local size_t granted = 0;
        if (granted < bytes) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return false; }
            granted = 227 * 1024;
        }
    }
    kern<<<a.B, NTH, bytes, stream>>>(sa, a, K);
    return true;
}
template <class R, class Model>
bool launch_persistent(const SgldArgs& sa, const KArgs& a, int K, cudaStream_t stream) {
    if (small_fast_config(a)) {
        const bool lat = a.B <= sm_count();
        if (a.N <= 256) return lat ? launch_persistent_shape<R, Model, 256, 1, true, true>(sa, a, K, stream) : launch_persistent_shape<R, Model, 256, 1, true>(sa, a, K, stream);
#if SGM_SMALL_LATENCY_ITEMS > 0
        if (a.N <= 1024 && a.B <= SGM_SMALL_LATENCY_ITEMS) return launch_persistent_shape<R, Model, 1024, 1, true>(sa, a, K, stream);
#endif
        if (a.N <= 1024) return lat ? launch_persistent_shape<R, Model, SGM_SMALL_MID_NTH, SGM_SMALL_MID_PPT, true, true>(sa, a, K, stream)
                                    : launch_persistent_shape<R, Model, SGM_SMALL_MID_NTH, SGM_SMALL_MID_PPT, true>(sa, a, K, stream);
        return launch_persistent_shape<R, Model, 1024, 2, true>(sa, a, K, stream);
    }
    if (a.N <= 1024 && a.B <= sm_count()) return launch_persistent_shape<R, Model, 512, 2, false, true>(sa, a, K, stream);
    if (a.N <= 1024) return launch_persistent_shape<R, Model, 1024, 1, false>(sa, a, K, stream);
    return launch_persistent_shape<R, Model, 1024, 2, false>(sa, a, K, stream);
}

// 1 = launched, 0 = not eligible (the caller falls back to one launch sequence per iteration), < 0 = error
template <class R>
int run_sgld_persistent(const sgm_pf_desc* d, const SgldArgs& sa, int K, cudaStream_t s) {
    if (sa.ipc != 1 || backward_pf(d->pf) || d->stat_kind != SGM_STAT_SCORE) return 0;
    KArgs a;
    { const int rc = make_kargs(d, a); if (rc != SGM_OK) return rc; }
    bool ok = false;
    // few chains: one cluster of CTAs per chain (N up to 8 x 2048); else one CTA per chain (N <= 2048)
    if (d->path == SGM_PATH_CLUSTER || (d->path == SGM_PATH_AUTO && a.N > 1024)) ok = run_sgld_cluster<R>(d, sa, a, K, s) == 1;
    if (!ok && (d->n_particles > SGM_SMALL_MAX_N || d->path == SGM_PATH_TILES || d->path == SGM_PATH_STEPS || d->path == SGM_PATH_CLUSTER)) return 0;
    if (!ok) switch (d->model) {
        case SGM_MODEL_SVM: ok = launch_persistent<R, SvmPrior>(sa, a, K, s); break;
        case SGM_MODEL_LGSSM: ok = d->kernel == SGM_KERNEL_PRIOR ? launch_persistent<R, LgssmPrior>(sa, a, K, s) : launch_persistent<R, LgssmOptimal>(sa, a, K, s); break;
        default: ok = d->kernel == SGM_KERNEL_PRIOR ? launch_persistent<R, GarchPrior>(sa, a, K, s) : launch_persistent<R, GarchOptimal>(sa, a, K, s);
    }
    if (!ok) return 0;
    sgld_advance_kernel<<<1, 1, 0, s>>>(sa, K);
    set_launch_count(2);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
    return 1;
}

template <class R>
int run_model(const sgm_pf_desc* d, cudaStream_t s) {
    switch (d->model) {
        case SGM_MODEL_SVM: return run_impl<R, SvmPrior>(d, s);
        case SGM_MODEL_LGSSM: return d->kernel == SGM_KERNEL_PRIOR ? run_impl<R, LgssmPrior>(d, s) : run_impl<R, LgssmOptimal>(d, s);
        default: return d->kernel == SGM_KERNEL_PRIOR ? run_impl<R, GarchPrior>(d, s) : run_impl<R, GarchOptimal>(d, s);
    }
}


}  // namespace sgmhost
