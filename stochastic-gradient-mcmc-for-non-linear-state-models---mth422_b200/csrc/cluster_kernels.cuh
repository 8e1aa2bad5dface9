// Latency regime, few items: ONE THREAD-BLOCK CLUSTER per item, particle system resident in the DISTRIBUTED SHARED
// MEMORY of its C CTAs (Hopper / Blackwell clusters: barrier.cluster + DSMEM loads and stores).  The single-CTA kernel
// of small_kernels.cuh is bound by the instruction issue of ONE SM (N = 1000: ~330 warp instructions per warp and time
// step on 32 warps); here the C SMs of a cluster share the particles of an item, so a time step is C times less work
// per SM plus two cluster barriers, and N up to 8 x 2048 fits without touching global memory inside the time loop.
//
// Ownership: CTA r of the cluster owns particles [r NL, (r + 1) NL), NL = NTH * PPT; records (statistics + state) and the
// fine CDF of those particles live in its shared memory.  One time step (same phases as small_kernels.cuh):
//   A  per warp: max / exp / inclusive scan of its particles' weights; the warp summary is STORED INTO EVERY CTA of the
//      cluster (DSMEM stores, lane j -> CTA j)                                            -> cluster barrier 1
//   B  every warp combines all C * NW warp summaries (local copy): M, total, its own offset / scale, S-bar; each particle's
//      CDF entry goes to the owner's `fine` array, every 8th entry also to the `coarse` array of EVERY CTA (the only
//      replicated data: N / 8 values)                                                     -> cluster barrier 2
//   C  per child: binary search of the local coarse array -> group of 8 parents; ONE remote 32-byte (f32) read of that
//      group's fine entries from its owner CTA finishes the search in registers; one remote 16-byte read gathers the parent
//      record; propose, reweight, statistic update, local store of the child
// Semantics are those of small_kernels.cuh / the reference (searchsorted(cdf, u, 'right'), iid multinomial uniforms).
#pragma once
#include <cooperative_groups.h>
#include "small_kernels.cuh"

namespace sgm {

namespace cg = cooperative_groups;
constexpr int CL_MAX_GW = 256;                    // warps per cluster (C * NTH / 32) the summary combine handles
constexpr int CL_JMAX = CL_MAX_GW / 32;

// ---- distributed shared memory, explicit PTX --------------------------------------------------------------------
// cooperative_groups' cluster.sync() compiles to MEMBAR.ALL.GPU + CCTL.IVALL + UCGABAR (a GPU-scope fence and an L1
// flush: ~1 us with remote stores in flight) and map_shared_rank() to generic LD.E / ST.E.  The time loop only exchanges
// SHARED memory inside the cluster, so it uses the cluster barrier with release / acquire at cluster scope and
// ld / st.shared::cluster on addresses mapped with `mapa`.
__device__ __forceinline__ void cluster_barrier() {
#ifdef SGM_CLUSTER_RELAXED      /* measurement only: no memory ordering, NOT a correct exchange */
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#else
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ uint32_t dsmem_map(const void* local, int rank) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void dsmem_st(uint32_t addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ float dsmem_ld(uint32_t addr, float) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ double dsmem_ld(uint32_t addr, double) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ Vec4T<float> dsmem_ld4(uint32_t addr, float) {
    Vec4T<float> v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ Vec4T<double> dsmem_ld4(uint32_t addr, double) {
    Vec4T<double> v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(addr + 16) : "memory");
    return v;
}

template <class R> __host__ __device__ inline size_t cluster_smem_bytes(int nl, int C, int nthreads, int nx, int np) {
    // coarse[C nl / 8] | fine[nl] | rec[2][nl] (16-byte records) | tail[2][KT][nl] | summaries[C NW][8]
    return sizeof(R) * ((size_t)C * nl / 8 + (size_t)nl * (size_t)(1 + 2 * 4 + 2 * (nx + np - 4)) + (size_t)C * (nthreads / 32) * SM_STRIDE);
}

template <class R, class Model, int NTH, int PPT, bool FAST>
__device__ void cluster_pf_item(const KArgs& a, int b, unsigned char* smem) {
    constexpr int NW = NTH / 32, NX = Model::NX, NP = Model::NP, W = NX + NP, KT = W - 4, NL = NTH * PPT;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int GW = C * NW, J = (GW + 31) / 32, NG = C * NL / 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gw = rank * NW + warp;
    const int N = a.N;
    const int Tb = a.T_buf[b], t1 = a.t1[b], tL = a.tL[b];
    const int nstat = FAST ? NP : stat_width<Model>(a.stat_kind);
    const bool carries = FAST || a.pf == SGM_PF_NEMETH, filter = !FAST && a.pf == SGM_PF_FILTER;
    const bool shrink = !FAST && carries && a.lambduh != 1.0;
    const bool injected = !FAST && a.rng_mode == SGM_RNG_INJECTED;
    const bool strat = !FAST && !injected && (a.resample == SGM_RESAMPLE_SYSTEMATIC || a.resample == SGM_RESAMPLE_STRATIFIED);
    const bool tracing = !FAST && (a.trace_anc || a.trace_x || a.trace_lw);
    const bool var32 = sizeof(R) == 8 && a.variates32 && !injected;
    R* const coarse = reinterpret_cast<R*>(smem);
    R* const fine = coarse + NG;
    Vec4T<R>* const rec0 = reinterpret_cast<Vec4T<R>*>(fine + NL);
    R* const tail0 = reinterpret_cast<R*>(rec0 + 2 * NL);
    R* const summ = tail0 + 2 * NL * KT;
    const typename Model::template Theta<R> th = Model::template load<R>(a.theta + (size_t)b * SGM_THETA_STRIDE);
    const RngKey key = item_key(a, b);
    const R NEG_INF = -Mth<R>::inf();
    const size_t item_off = (size_t)b * N;
    const double* obs = a.obs + a.obs_off[b];
    const double* wts = (a.wts_off && a.wts_off[b] >= 0) ? a.step_weights + a.wts_off[b] : nullptr;
    const R lam = (R)a.lambduh;
    const uint32_t gtid = (uint32_t)(rank * NTH + tid);            // random-number index: the thread's place in the item

    auto draw = [&](uint32_t step, R* u, R* z) {
        if (var32) {
            float uf[PPT], zf[PPT];
            small_draw<PPT>(key, gtid, step, uf, zf);
#pragma unroll
            for (int k = 0; k < PPT; ++k) { u[k] = (R)uf[k]; z[k] = (R)zf[k]; }
        } else {
            small_draw<PPT>(key, gtid, step, u, z);
        }
    };
    auto store_p = [&](int buf, int li, const R* r) {                // local records only
        Vec4T<R> v; v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3];
        rec0[buf * NL + li] = v;
#pragma unroll
        for (int q = 0; q < KT; ++q) tail0[(buf * KT + q) * NL + li] = r[4 + q];
    };
    auto load_remote = [&](int owner, int buf, int li, R* r) {       // DSMEM gather of a parent record
        const Vec4T<R> v = dsmem_ld4(dsmem_map(rec0 + buf * NL + li, owner), (R)0);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
#pragma unroll
        for (int q = 0; q < KT; ++q) r[4 + q] = dsmem_ld(dsmem_map(tail0 + (buf * KT + q) * NL + li, owner), (R)0);
    };

    R lw[PPT], sv[PPT][4];
    int par = 0;
    {   // ---- init (buffered_smoother.py:67-75) ----
        const R mean = (R)a.prior_mean[b], sd = (R)::sqrt(a.prior_var[b]);
        R u[PPT], z[PPT];
        if (!injected) draw(0xffffu, u, z);
        for (int j = tid; j < NG; j += NTH) coarse[j] = Mth<R>::inf();         // groups never written (>= N) stay +inf
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int li = tid * PPT + k, i = rank * NL + li;
            lw[k] = (i < N) ? (R)0 : NEG_INF;
            sv[k][0] = sv[k][1] = sv[k][2] = sv[k][3] = (R)0;
            fine[li] = Mth<R>::inf();
            if (i < N) {
                if (injected) z[k] = (R)a.inj_z0[item_off + i];
                R r[W];
#pragma unroll
                for (int q = 0; q < W; ++q) r[q] = (R)0;
                Model::init(mean, sd, z[k], r + NP);
                store_p(0, li, r);
                if (tracing) {
                    if (a.trace_x) for (int d = 0; d < NX; ++d)
                        reinterpret_cast<R*>(a.trace_x)[((size_t)b * (a.max_T + 1) * N + i) * NX + d] = r[NP + d];
                    if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[(size_t)b * (a.max_T + 1) * N + i] = (R)0;
                }
            }
        }
    }
    cluster.sync();                                   // every CTA's coarse array is initialised before remote stores land in it
    double loglik = 0.0, accf[4] = {0.0, 0.0, 0.0, 0.0};
    int status = 0;
    R wt_prev = (R)0;

    for (int t = 0;; ++t) {
        const bool final_pass = t >= Tb;
        const bool need_ws = nstat > 0 && (filter || shrink || (carries && final_pass));
        const bool in_sub = !final_pass && t >= t1 && t < tL;
        const R y = final_pass ? (R)0 : (R)obs[t];
        const R wt = in_sub ? (R)(wts ? wts[t - t1] : 1.0) : (R)0;
        // ---- A: warp summary, stored into every CTA of the cluster -------------------------------------
        R m = NEG_INF;
#pragma unroll
        for (int k = 0; k < PPT; ++k) m = nan_max(m, lw[k]);
        m = warp_max(m);
        const R msafe = (m == NEG_INF) ? (R)0 : m;
        R w[PPT], pre[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) { w[k] = Mth<R>::exp(lw[k] - msafe); pre[k] = w[k] + (k ? pre[k - 1] : (R)0); }
        const R incl = warp_incl_scan(pre[PPT - 1]);
        const R excl = incl - pre[PPT - 1];
#pragma unroll
        for (int k = 0; k < PPT; ++k) pre[k] += excl;
        const R s_w = __shfl_sync(FULL, incl, 31);
        R wsum[4] = {(R)0, (R)0, (R)0, (R)0};
        if (need_ws) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                R acc = (R)0;
#pragma unroll
                for (int k = 0; k < PPT; ++k) acc += sv[k][q] * w[k];
                wsum[q] = (q < nstat) ? warp_sum(acc) : (R)0;
            }
        }
        if (lane < C) {
            const uint32_t p = dsmem_map(summ + gw * SM_STRIDE, lane);
            dsmem_st(p, m); dsmem_st(p + (uint32_t)sizeof(R), s_w);
            if (need_ws) {
#pragma unroll
                for (int q = 0; q < 4; ++q) dsmem_st(p + (uint32_t)((2 + q) * sizeof(R)), wsum[q]);
            }
        }
        cluster_barrier();                                                    // cluster barrier 1
        // ---- B: combine the GW warp summaries (lane l: warps l J .. l J + J - 1) ------------------------------
        R mloc = NEG_INF;
#pragma unroll
        for (int jj = 0; jj < CL_JMAX; ++jj) { const int e = lane * J + jj; if (jj < J && e < GW) mloc = nan_max(mloc, summ[e * SM_STRIDE]); }
        const R M = warp_max(mloc);
        R run = (R)0, cand_off = (R)0, cand_sc = (R)0, ws_l[4] = {(R)0, (R)0, (R)0, (R)0};
        const int my_slot = gw % J;
#pragma unroll
        for (int jj = 0; jj < CL_JMAX; ++jj) {
            const int e = lane * J + jj;
            if (jj < J && e < GW) {
                const R me = summ[e * SM_STRIDE];
                const R el = (me != NEG_INF) ? Mth<R>::exp(me - M) : (R)0;
                if (jj == my_slot) { cand_off = run; cand_sc = el; }
                run += el * summ[e * SM_STRIDE + 1];
                if (need_ws) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) ws_l[q] += el * summ[e * SM_STRIDE + 2 + q];
                }
            }
        }
        const R lincl = warp_incl_scan(run);
        const R total = __shfl_sync(FULL, lincl, 31);
        const R off_me = __shfl_sync(FULL, (lincl - run) + cand_off, gw / J);
        const R sc_me = __shfl_sync(FULL, cand_sc, gw / J);
        R sbar[4] = {(R)0, (R)0, (R)0, (R)0};
        if (need_ws) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < nstat) sbar[q] = warp_sum(ws_l[q]) / total;
        }
        if (tid == 0) {
            if (!(total > (R)0) || !(total < Mth<R>::inf()) || !(M == M) || !(M > NEG_INF && M < Mth<R>::inf()))
                status |= (total == (R)0 || M == NEG_INF) ? SGM_STATUS_ZERO_WEIGHT : SGM_STATUS_NAN_WEIGHT;
            if (t > 0) {
                if (wt_prev != (R)0) loglik += (double)wt_prev * ((double)M + (double)Mth<R>::log(total / (R)N));
                if (filter) for (int j = 0; j < nstat; ++j) accf[j] += (double)sbar[j];
            }
            wt_prev = in_sub ? (wts ? (R)wts[t - t1] : (R)1) : (R)0;
        }
        if (final_pass) {
            if (tid == 0 && rank == 0) {
                a.loglik[b] = loglik;
                a.status[b] = status;
                for (int j = 0; j < 8; ++j) a.grad[(size_t)b * 8 + j] = 0.0;
                for (int j = 0; j < nstat; ++j) a.grad[(size_t)b * 8 + j] = filter ? accf[j] : (double)sbar[j];
            }
            break;
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int li = tid * PPT + k, i = rank * NL + li;
            if (i < N) {
                const R c = Mth<R>::fma(sc_me, pre[k], off_me);
                fine[li] = c;
                if ((i & 7) == 7) for (int r = 0; r < C; ++r) dsmem_st(dsmem_map(coarse + (i >> 3), r), c);
            }
        }
        cluster_barrier();                                                    // cluster barrier 2
        // ---- C: resample -> propagate -> reweight -> statistic update ----------------------------------
        const R hs = (carries || filter) ? wt : (R)0;
        const int stat_kind = (in_sub && hs != (R)0) ? (FAST ? (int)SGM_STAT_SCORE : a.stat_kind) : (int)SGM_STAT_NONE;
        R u[PPT], z[PPT];
        if (!injected) {
            draw((uint32_t)t, u, z);
            if (!FAST && a.resample == SGM_RESAMPLE_SYSTEMATIC) {
                R u4[4];
                if (var32) { float f4[4]; rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, f4); u4[0] = (R)f4[0]; }
                else rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, u4);
#pragma unroll
                for (int k = 0; k < PPT; ++k) u[k] = u4[0];
            }
        }
        const int g_last = (N - 1) >> 3;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int li = tid * PPT + k, i = rank * NL + li;
            if (injected) { u[k] = (i < N) ? (R)a.inj_u[((size_t)b * a.max_T + t) * N + i] : (R)0; z[k] = (i < N) ? (R)a.inj_z[((size_t)b * a.max_T + t) * N + i] : (R)0; }
            const R tg = strat ? (((R)i + u[k]) / (R)N) * total : u[k] * total;
            // group of 8 parents: number of coarse entries (= last CDF entry of every group) <= target
            int g = 0;
#pragma unroll 1
            for (int step = NG >> 1; step > 0; step >>= 1)
                if (coarse[g + step - 1] <= tg) g += step;
            g = min(g, g_last);
            const int owner = (g * 8) / NL, lo8 = (g * 8) % NL;
            const uint32_t rf = dsmem_map(fine + lo8, owner);
            const Vec4T<R> f0 = dsmem_ld4(rf, (R)0), f1 = dsmem_ld4(rf + (uint32_t)(4 * sizeof(R)), (R)0);
            int cnt = (f0.x <= tg) + (f0.y <= tg) + (f0.z <= tg) + (f0.w <= tg) + (f1.x <= tg) + (f1.y <= tg) + (f1.z <= tg) + (f1.w <= tg);
            int anc = min(g * 8 + cnt, N - 1);
            if (!(tg < total)) anc = N - 1;
            lw[k] = NEG_INF;
            if (i < N) {
                R ra[W], rn[W];
                load_remote(anc / NL, par, anc % NL, ra);
                Model::propagate(th, ra + NP, y, z[k], rn + NP);
                lw[k] = Model::log_weight(th, ra + NP, rn + NP, y);
                R h[4] = {(R)0, (R)0, (R)0, (R)0};
                if (stat_kind == SGM_STAT_SCORE) Model::score(th, ra + NP, rn + NP, y, h);
                else if (stat_kind == SGM_STAT_SUFF) Model::suff(ra + NP, rn + NP, h);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const R sa = (q < NP) ? ra[q] : (R)0;
                    if (!carries) sv[k][q] = h[q] * hs;
                    else if (!shrink) sv[k][q] = sa + h[q] * hs;
                    else sv[k][q] = lam * sa + ((R)((1.0 - a.lambduh) * (double)sbar[q]) + h[q] * hs);
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) rn[q] = carries ? sv[k][q] : (R)0;
                store_p(par ^ 1, li, rn);
                if (tracing) {
                    if (a.trace_anc) a.trace_anc[((size_t)b * a.max_T + t) * N + i] = anc;
                    if (a.trace_x) for (int d = 0; d < NX; ++d)
                        reinterpret_cast<R*>(a.trace_x)[(((size_t)b * (a.max_T + 1) + t + 1) * N + i) * NX + d] = rn[NP + d];
                    if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[((size_t)b * (a.max_T + 1) + t + 1) * N + i] = lw[k];
                }
            }
        }
        par ^= 1;
        // no barrier here: the next step's two cluster barriers order these stores before the next remote gathers, and every
        // remote read of the old records / CDF above precedes the reader's arrival at cluster barrier 1 of the next step
    }
    if (!FAST && (a.out_x || a.out_lw || a.out_stats)) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int li = tid * PPT + k, i = rank * NL + li;
            if (i < N) {
                R r[W];
                load_remote(rank, par, li, r);
                if (a.out_x) for (int d = 0; d < NX; ++d) reinterpret_cast<R*>(a.out_x)[(item_off + i) * NX + d] = r[NP + d];
                if (a.out_lw) reinterpret_cast<R*>(a.out_lw)[item_off + i] = lw[k];
                if (a.out_stats) for (int q = 0; q < NP; ++q) reinterpret_cast<R*>(a.out_stats)[(item_off + i) * NP + q] = r[q];
            }
        }
    }
    cluster.sync();        // no CTA may exit (and release its shared memory) while others still read it
}

template <class R, class Model, int NTH, int PPT, bool FAST>
__global__ void __launch_bounds__(NTH, 1) pf_cluster_kernel(KArgs a) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int C = (int)cg::this_cluster().num_blocks();
    cluster_pf_item<R, Model, NTH, PPT, FAST>(a, a.b0 + (int)(blockIdx.x / C), small_smem);
}

}  // namespace sgm
