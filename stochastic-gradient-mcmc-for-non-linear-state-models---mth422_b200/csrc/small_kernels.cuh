// Small particle counts (N <= 2048): the whole buffered time loop of an item in ONE CTA with the particle system
// resident in SHARED MEMORY -- no global-memory traffic inside the time loop at all (the per-step tile kernels stream
// 40-56 B per particle-step through HBM / L2; here a time step is two block barriers and shared-memory accesses).
// This is the SG-MCMC regime of the reference's own scripts (N = 1000, one subsequence per iteration), where a
// gradient is latency-bound: every thread owns PPT (1 or 2) particles instead of a warp owning 256, so the dependent
// instruction chain of a time step is ~1/8 as long as in the warp-tile kernels.
//
// One time step (reference pf.py:7-38 + :138-181 / :40-82, buffered_smoother.py:94-126):
//   A  per warp: max of the log-weights, w = exp(lw - m_warp), inclusive scan (registers + shuffles), weighted
//      statistic sums; lane 0 stores the warp summary                                             -> barrier 1
//   B  every warp combines the <= 32 warp summaries itself (shuffles): global max M, total, offsets, scales,
//      S-bar; writes the CDF entries of its particles to shared memory (c_i = off_w + sc_w * prefix_i: the same
//      hierarchical form as the tile kernels); thread 0 adds the log-likelihood increment        -> barrier 2
//   C  per particle: uniform -> branch-free binary search of the shared-memory CDF (np.random.choice ==
//      searchsorted(cdf, u, 'right')), gather the parent record (one 16-byte shared-memory load), propose, reweight,
//      update the statistic, store the child
// Resampling is the reference's plain multinomial (iid uniforms): parents are gathered from shared memory, so the
// sorted-target trick of the streaming kernels has nothing to gain here ('multinomial_sorted' has the same law and
// takes this path too); systematic / stratified targets are supported as in the tile kernels.
// FAST = the production configuration (device randoms, multinomial law, Poyiadjis O(N) with the model score, no traces /
// exports) with its run-time flags folded to constants; the generic instantiation reads every flag at run time.
#pragma once
#include "pf_kernels.cuh"

namespace sgm {

enum : uint32_t { STREAM_SMALL = 9 };
constexpr int SM_STRIDE = 8;          // R values per warp summary: m, s, ws[0..3]

template <class R> __host__ __device__ inline size_t small_smem_bytes(int npad, int nx, int np) {
    // cdf | sampled CDF levels (npad / 2) | rec[2] (16-byte records) | tail[2][KT] | warp summaries
    return sizeof(R) * ((size_t)npad * (size_t)(1 + 2 * 4 + 2 * (nx + np - 4)) + (size_t)npad / 2 + 32 * SM_STRIDE);
}

// u[PPT] uniforms and z[PPT] standard normals of thread `tid` at `step`
template <int PPT> __device__ __forceinline__ void small_draw(const RngKey& key, uint32_t tid, uint32_t step, float* u, float* z) {
#pragma unroll
    for (int h = 0; h < (PPT + 1) / 2; ++h) {
        const uint4 r = rng_raw(key, tid, step, STREAM_SMALL, (uint32_t)h);
        const float rad = sqrt_approx(__log2f(u01f(r.z)) * -1.3862943611198906f);
        float s, c;
        __sincosf(6.28318530717958647692f * u01f(r.w), &s, &c);
        u[2 * h] = u01f(r.x); z[2 * h] = rad * c;
        if (2 * h + 1 < PPT) { u[2 * h + 1] = u01f(r.y); z[2 * h + 1] = rad * s; }
    }
}
template <int PPT> __device__ __forceinline__ void small_draw(const RngKey& key, uint32_t tid, uint32_t step, double* u, double* z) {
#pragma unroll
    for (int h = 0; h < (PPT + 1) / 2; ++h) {
        const uint4 a = rng_raw(key, tid, step, STREAM_SMALL, (uint32_t)(2 * h)), b = rng_raw(key, tid, step, STREAM_SMALL, (uint32_t)(2 * h + 1));
        const double rad = sqrt(-2.0 * log(u01d(b.x, b.y)));
        double s, c;
        sincospi(2.0 * u01d(b.z, b.w), &s, &c);
        u[2 * h] = u01d(a.x, a.y); z[2 * h] = rad * c;
        if (2 * h + 1 < PPT) { u[2 * h + 1] = u01d(a.z, a.w); z[2 * h + 1] = rad * s; }
    }
}

// ---- 4-ary search on sampled copies of the CDF ---------------------------------------------------------------------
// ncu (round 2): half of the shared-memory kernel's stall samples sat on the 10 dependent LDS of the binary search.  The CDF
// is therefore kept with sampled copies  L1[j] = cdf[4 j + 3], L2[j] = cdf[16 j + 15], ...  down to a top level of 4 or 8
// entries; a search reads ONE 16-byte group per level (5 dependent loads for 1024 entries instead of 10) and counts the
// entries <= target among its first three (the fourth is the parent level's entry, already known to exceed the target).
// Entries beyond N are +inf on every level.  Same result as searchsorted(cdf, target, 'right').
__host__ __device__ constexpr int samp_off(int npad, int sz) {      // offset of the level with `sz` entries inside the sampled area
    int off = 0;
    for (int s = npad / 4; s > sz; s /= 4) off += s;
    return off;
}
__host__ __device__ constexpr int samp_top(int npad) {               // size of the top level: 4 or 8
    int s = npad / 4;
    while (s > 8) s /= 4;
    return s;
}
// Searches run on shared-memory BYTE addresses: with  ad = base + OFF(level) + 16 g  the group of the next level is at
// 4 ad + K + 16 #{entries <= target},  K a per-level constant; nvcc turns this into one LEA and three predicated adds.
__device__ __forceinline__ Vec4T<float> lds_group(uint32_t ad, float) {
    Vec4T<float> v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad));
    return v;
}
__device__ __forceinline__ Vec4T<double> lds_group(uint32_t ad, double) {
    Vec4T<double> v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ad));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(v.z), "=d"(v.w) : "r"(ad));
    return v;
}
template <class R, int NPAD> __host__ __device__ constexpr int samp_level_bytes(int sz) {      // byte offset of a level from the CDF
    return sz == NPAD ? 0 : (NPAD + samp_off(NPAD, sz)) * (int)sizeof(R);
}
template <class R, int PPT, int NPAD, int SZ>
__device__ __forceinline__ void samp_descend(uint32_t base, const R* tg, uint32_t* ad) {
    constexpr int GB = 4 * (int)sizeof(R);
    if constexpr (SZ == NPAD) {                        // the CDF itself: ad becomes sizeof(R) * searchsorted index
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const Vec4T<R> v = lds_group(ad[k], (R)0);
            uint32_t n = ad[k] - base;
            if (v.x <= tg[k]) n += (uint32_t)sizeof(R);
            if (v.y <= tg[k]) n += (uint32_t)sizeof(R);
            if (v.z <= tg[k]) n += (uint32_t)sizeof(R);
            ad[k] = n;
        }
    } else {
        const uint32_t K = (uint32_t)(samp_level_bytes<R, NPAD>(SZ * 4) - 4 * samp_level_bytes<R, NPAD>(SZ)) - 3u * base;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const Vec4T<R> v = lds_group(ad[k], (R)0);
            uint32_t n = 4u * ad[k] + K;
            if (v.x <= tg[k]) n += GB;
            if (v.y <= tg[k]) n += GB;
            if (v.z <= tg[k]) n += GB;
            ad[k] = n;
        }
        samp_descend<R, PPT, NPAD, SZ * 4>(base, tg, ad);
    }
}
template <class R, int PPT, int NPAD>
__device__ __forceinline__ void samp_search(const R* cdf, const R* tg, int* anc) {
    constexpr int TOP = samp_top(NPAD);
    constexpr int GB = 4 * (int)sizeof(R);
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(cdf);
    uint32_t ad[PPT];
    if constexpr (TOP == 4) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) ad[k] = base + samp_level_bytes<R, NPAD>(TOP);
        samp_descend<R, PPT, NPAD, TOP>(base, tg, ad);
    } else {                                           // 8 top entries: two groups, seven of the entries counted
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const Vec4T<R> v0 = lds_group(base + samp_level_bytes<R, NPAD>(TOP), (R)0);
            const Vec4T<R> v1 = lds_group(base + samp_level_bytes<R, NPAD>(TOP) + GB, (R)0);
            uint32_t n = base + samp_level_bytes<R, NPAD>(TOP * 4);
            if (v0.x <= tg[k]) n += GB;
            if (v0.y <= tg[k]) n += GB;
            if (v0.z <= tg[k]) n += GB;
            if (v0.w <= tg[k]) n += GB;
            if (v1.x <= tg[k]) n += GB;
            if (v1.y <= tg[k]) n += GB;
            if (v1.z <= tg[k]) n += GB;
            ad[k] = n;
        }
        samp_descend<R, PPT, NPAD, TOP * 4>(base, tg, ad);
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) anc[k] = (int)(ad[k] / (uint32_t)sizeof(R));
}
// store CDF entry i and its sampled copies
template <class R, int NPAD>
__device__ __forceinline__ void samp_store(R* cdf, R* samp, int i, R c) {
    cdf[i] = c;
    int sz = NPAD / 4, off = 0, sh = 2;
#pragma unroll
    for (; sz >= 4; sz /= 4, sh += 2) {
        if ((i & ((1 << sh) - 1)) == (1 << sh) - 1) samp[off + (i >> sh)] = c;
        off += sz;
        if (sz <= 8) break;
    }
}

// The whole time loop of item b by the calling CTA (NTH threads, every thread calls).  `smem` = small_smem_bytes<R>()
// bytes of 16-byte aligned shared memory.  Results go to a.grad / a.loglik / a.status (and the optional outputs).
template <class R, class Model, int NTH, int PPT, bool FAST>
__device__ void small_pf_item(const KArgs& a, int b, unsigned char* smem) {
    constexpr int NW = NTH / 32, NX = Model::NX, NP = Model::NP, W = NX + NP, KT = W - 4, NPAD = NTH * PPT;
    static_assert(NW <= 32, "at most 32 warps");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N;
    const int Tb = a.T_buf[b], t1 = a.t1[b], tL = a.tL[b];
    const int nstat = FAST ? NP : stat_width<Model>(a.stat_kind);
    const bool carries = FAST || a.pf == SGM_PF_NEMETH, filter = !FAST && a.pf == SGM_PF_FILTER;
    const bool shrink = !FAST && carries && a.lambduh != 1.0;
    const bool injected = !FAST && a.rng_mode == SGM_RNG_INJECTED;
    const bool strat = !FAST && !injected && (a.resample == SGM_RESAMPLE_SYSTEMATIC || a.resample == SGM_RESAMPLE_STRATIFIED);
    const bool tracing = !FAST && (a.trace_anc || a.trace_x || a.trace_lw);
    const bool var32 = sizeof(R) == 8 && a.variates32 && !injected;
    R* const cdf = reinterpret_cast<R*>(smem);
    R* const samp = cdf + NPAD;                        // sampled CDF levels (NPAD / 2 entries reserved)
    Vec4T<R>* const rec0 = reinterpret_cast<Vec4T<R>*>(samp + NPAD / 2);
    R* const tail0 = reinterpret_cast<R*>(rec0 + 2 * NPAD);
    R* const summ = tail0 + 2 * NPAD * KT;
    const typename Model::template Theta<R> th = Model::template load<R>(a.theta + (size_t)b * SGM_THETA_STRIDE);
    const RngKey key = item_key(a, b);
    const R NEG_INF = -Mth<R>::inf();
    const size_t item_off = (size_t)b * N;
    const double* obs = a.obs + a.obs_off[b];
    const double* wts = (a.wts_off && a.wts_off[b] >= 0) ? a.step_weights + a.wts_off[b] : nullptr;
    const R lam = (R)a.lambduh;

    auto draw = [&](uint32_t step, R* u, R* z) {
        if (var32) {
            float uf[PPT], zf[PPT];
            small_draw<PPT>(key, (uint32_t)tid, step, uf, zf);
#pragma unroll
            for (int k = 0; k < PPT; ++k) { u[k] = (R)uf[k]; z[k] = (R)zf[k]; }
        } else {
            small_draw<PPT>(key, (uint32_t)tid, step, u, z);
        }
    };
    // record of particle i: r[0..NP) statistics, r[NP..W) state -- first four components as one 16-byte vector
    auto load_p = [&](int buf, int i, R* r) {
        const Vec4T<R> v = rec0[buf * NPAD + i];
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
#pragma unroll
        for (int q = 0; q < KT; ++q) r[4 + q] = tail0[(buf * KT + q) * NPAD + i];
    };
    auto store_p = [&](int buf, int i, const R* r) {
        Vec4T<R> v; v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3];
        rec0[buf * NPAD + i] = v;
#pragma unroll
        for (int q = 0; q < KT; ++q) tail0[(buf * KT + q) * NPAD + i] = r[4 + q];
    };

    R lw[PPT], sv[PPT][4];
    int par = 0;
    {   // ---- init: x0 ~ N(prior_mean, prior_var), lw = 0, statistics = 0  (buffered_smoother.py:67-75) ----
        const R mean = (R)a.prior_mean[b], sd = (R)::sqrt(a.prior_var[b]);
        R u[PPT], z[PPT];
        if (!injected) draw(0xffffu, u, z);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid * PPT + k;
            lw[k] = (i < N) ? (R)0 : NEG_INF;
            sv[k][0] = sv[k][1] = sv[k][2] = sv[k][3] = (R)0;
            cdf[i] = Mth<R>::inf();                       // entries >= N stay +inf: the searches need no bound check
            if (i < NPAD / 2) samp[i] = Mth<R>::inf();
            if (i < N) {
                if (injected) z[k] = (R)a.inj_z0[item_off + i];
                R r[W];
#pragma unroll
                for (int q = 0; q < W; ++q) r[q] = (R)0;
                Model::init(mean, sd, z[k], r + NP);
                store_p(0, i, r);
                if (tracing) {
                    if (a.trace_x) for (int d = 0; d < NX; ++d)
                        reinterpret_cast<R*>(a.trace_x)[((size_t)b * (a.max_T + 1) * N + i) * NX + d] = r[NP + d];
                    if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[(size_t)b * (a.max_T + 1) * N + i] = (R)0;
                }
            }
        }
    }
    double loglik = 0.0, accf[4] = {0.0, 0.0, 0.0, 0.0};
    int status = 0;
    R wt_prev = (R)0;                                  // log-likelihood weight of the step that produced the current weights

    for (int t = 0;; ++t) {
        const bool final_pass = t >= Tb;
        const bool need_ws = nstat > 0 && (filter || shrink || (carries && final_pass));
        // this step's observation and statistic weight: issued first, consumed in phase C
        const bool in_sub = !final_pass && t >= t1 && t < tL;
        const R y = final_pass ? (R)0 : (R)obs[t];
        const R wt = in_sub ? (R)(wts ? wts[t - t1] : 1.0) : (R)0;
        // this step's random numbers do not depend on the weights: drawn here, in the same basic block as phase A, so that
        // the Philox rounds and the Box-Muller transforms interleave with the max / exp / scan chain below (a step of this
        // kernel is a chain of dependent latencies on 4 warps per scheduler, not a throughput problem)
        R u[PPT], z[PPT];
        if (!injected && !final_pass) {
            draw((uint32_t)t, u, z);
            if (!FAST && a.resample == SGM_RESAMPLE_SYSTEMATIC) {
                R u4[4];
                if (var32) { float f4[4]; rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, f4); u4[0] = (R)f4[0]; }
                else rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, u4);
#pragma unroll
                for (int k = 0; k < PPT; ++k) u[k] = u4[0];
            }
        }
        // ---- A: warp summary --------------------------------------------------------------------------
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
        R wsum[4] = {(R)0, (R)0, (R)0, (R)0};
        if (need_ws) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                R acc = (R)0;
#pragma unroll
                for (int k = 0; k < PPT; ++k) acc += sv[k][q] * w[k];          // w = 0 beyond N
                wsum[q] = (q < nstat) ? warp_sum(acc) : (R)0;
            }
        }
        if (lane == 31) {                                  // the lane that holds the warp's weight sum; m is warp-uniform
            R* p = summ + warp * SM_STRIDE;
            p[0] = m; p[1] = incl;
            if (need_ws) { p[2] = wsum[0]; p[3] = wsum[1]; p[4] = wsum[2]; p[5] = wsum[3]; }
        }
        __syncthreads();                                                      // barrier 1
        // ---- B: cross-warp combine (every warp, redundantly: fixed shuffle trees => identical results) ----
        const bool has = lane < NW;
        const R ml = has ? summ[lane * SM_STRIDE] : NEG_INF;
        const R M = warp_max(ml);
        const R el = (has && ml != NEG_INF) ? Mth<R>::exp(ml - M) : (R)0;
        const R loc = has ? el * summ[lane * SM_STRIDE + 1] : (R)0;
        const R lincl = warp_incl_scan_low<NW>(loc);       // valid in lanes < NW
        const R total = __shfl_sync(FULL, lincl, NW - 1);
        const R off_me = __shfl_sync(FULL, lincl - loc, warp);
        const R sc_me = (m != NEG_INF) ? Mth<R>::exp(m - M) : (R)0;       // = lane `warp`'s el (same operands)
        R sbar[4] = {(R)0, (R)0, (R)0, (R)0};
        if (need_ws) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < nstat) sbar[q] = warp_sum(has ? el * summ[lane * SM_STRIDE + 2 + q] : (R)0) / total;
        }
        if (tid == 0) {
            if (!(total > (R)0) || !(total < Mth<R>::inf()) || !(M == M) || !(M > NEG_INF && M < Mth<R>::inf()))
                status |= (total == (R)0 || M == NEG_INF) ? SGM_STATUS_ZERO_WEIGHT : SGM_STATUS_NAN_WEIGHT;
            if (t > 0) {
                // buffered_smoother.py:124-126, max-shifted; the logarithm in the arithmetic type of the run
                if (wt_prev != (R)0) loglik += (double)wt_prev * ((double)M + (double)Mth<R>::log(total / (R)N));
                if (filter) for (int j = 0; j < nstat; ++j) accf[j] += (double)sbar[j];      // pf.py:77-80
            }
            wt_prev = in_sub ? (wts ? (R)wts[t - t1] : (R)1) : (R)0;
        }
        if (final_pass) {
            if (tid == 0) {
                a.loglik[b] = loglik;
                a.status[b] = status;
                for (int j = 0; j < 8; ++j) a.grad[(size_t)b * 8 + j] = 0.0;
                for (int j = 0; j < nstat; ++j) a.grad[(size_t)b * 8 + j] = filter ? accf[j] : (double)sbar[j];
            }
            break;
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid * PPT + k;
            if (i < N) samp_store<R, NPAD>(cdf, samp, i, Mth<R>::fma(sc_me, pre[k], off_me));
        }
        __syncthreads();                                                      // barrier 2
        // ---- C: resample -> propagate -> reweight -> statistic update ----------------------------------
        const R hs = (carries || filter) ? wt : (R)0;
        const int stat_kind = (in_sub && hs != (R)0) ? (FAST ? (int)SGM_STAT_SCORE : a.stat_kind) : (int)SGM_STAT_NONE;
        const R cmax = cdf[N - 1];
        R tg[PPT];
        int ancs[PPT];
        bool over[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid * PPT + k;
            if (injected) { u[k] = (i < N) ? (R)a.inj_u[((size_t)b * a.max_T + t) * N + i] : (R)0; z[k] = (i < N) ? (R)a.inj_z[((size_t)b * a.max_T + t) * N + i] : (R)0; }
            tg[k] = strat ? (((R)i + u[k]) / (R)N) * total : u[k] * total;
            over[k] = !(tg[k] < cmax);                         // u * total rounded up to the total (or a NaN weight)
        }
        // searchsorted(cdf, target, 'right') = number of entries <= target (entries >= N are +inf)
        samp_search<R, PPT, NPAD>(cdf, tg, ancs);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid * PPT + k;
            const int anc = over[k] ? N - 1 : min(ancs[k], N - 1);
            lw[k] = NEG_INF;
            if (i < N) {
                R ra[W], rn[W];
                load_p(par, anc, ra);
                Model::propagate(th, ra + NP, y, z[k], rn + NP);
                lw[k] = Model::log_weight(th, ra + NP, rn + NP, y);
                R h[4] = {(R)0, (R)0, (R)0, (R)0};
                if (stat_kind == SGM_STAT_SCORE) Model::score(th, ra + NP, rn + NP, y, h);
                else if (stat_kind == SGM_STAT_SUFF) Model::suff(ra + NP, rn + NP, h);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // pf.py:175-179 (Nemeth / Poyiadjis O(N)), pf.py:70-71 (filter: nothing carried)
                    const R sa = (q < NP) ? ra[q] : (R)0;
                    if (!carries) sv[k][q] = h[q] * hs;
                    else if (!shrink) sv[k][q] = sa + h[q] * hs;
                    else sv[k][q] = lam * sa + ((R)((1.0 - a.lambduh) * (double)sbar[q]) + h[q] * hs);
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) rn[q] = carries ? sv[k][q] : (R)0;
                store_p(par ^ 1, i, rn);
                if (tracing) {
                    if (a.trace_anc) a.trace_anc[((size_t)b * a.max_T + t) * N + i] = anc;
                    if (a.trace_x) for (int d = 0; d < NX; ++d)
                        reinterpret_cast<R*>(a.trace_x)[(((size_t)b * (a.max_T + 1) + t + 1) * N + i) * NX + d] = rn[NP + d];
                    if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[((size_t)b * (a.max_T + 1) + t + 1) * N + i] = lw[k];
                }
            }
        }
        par ^= 1;
        // no barrier here: the next step's barriers 1 and 2 order these stores before the next gathers, and every
        // read of the old buffers / the CDF above precedes barrier 1 of the next step
    }
    // ---- optional export of the final particle system (each thread: its own particles) ----
    if (!FAST && (a.out_x || a.out_lw || a.out_stats)) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tid * PPT + k;
            if (i < N) {
                R r[W];
                load_p(par, i, r);
                if (a.out_x) for (int d = 0; d < NX; ++d) reinterpret_cast<R*>(a.out_x)[(item_off + i) * NX + d] = r[NP + d];
                if (a.out_lw) reinterpret_cast<R*>(a.out_lw)[item_off + i] = lw[k];
                if (a.out_stats) for (int q = 0; q < NP; ++q) reinterpret_cast<R*>(a.out_stats)[(item_off + i) * NP + q] = r[q];
            }
        }
    }
}

// host-side eligibility of the FAST instantiation
inline bool small_fast_config(const KArgs& a) {
    return a.rng_mode == SGM_RNG_PHILOX && a.pf == SGM_PF_NEMETH && a.lambduh == 1.0 && a.stat_kind == SGM_STAT_SCORE &&
           (a.resample == SGM_RESAMPLE_MULTINOMIAL || a.resample == SGM_RESAMPLE_MULTINOMIAL_SORTED) &&
           !a.trace_anc && !a.trace_x && !a.trace_lw && !a.out_x && !a.out_lw && !a.out_stats;
}

// LAT = false: compiled for 1024 resident threads per SM (64 registers), the throughput shape of mid-size batches.
// LAT = true : one CTA per SM (batches of at most one item per SM, where nothing else could be resident anyway): the register cap
//              goes, ptxas keeps ~110 registers and schedules the step's dependent chain with more loads / shuffles in flight --
//              measured 0.1235 -> 0.1130 ms per 60-step gradient at N = 1000 (same arithmetic, same results).
constexpr int small_min_ctas(int nth, bool lat) { return lat ? 1 : 1024 / nth; }
template <class R, class Model, int NTH, int PPT, bool FAST, bool LAT = false>
__global__ void __launch_bounds__(NTH, small_min_ctas(NTH, LAT)) pf_small_kernel(KArgs a) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    small_pf_item<R, Model, NTH, PPT, FAST>(a, a.b0 + blockIdx.x, small_smem);
}

}  // namespace sgm
