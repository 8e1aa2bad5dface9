// Kernels of the batched buffered particle filter / smoother (O(N) paths).
//
// Data layout in HBM (all caller workspace; B items, N particles, G = ceil(N / TILE) tiles per item):
//   rec [2][B][N][4]   R   first four components of the particle record  (stats..., then state)
//   tail[2][B][N][KT]  R   remaining KT = NX + NP - 4 components (SVM 0, LGSSM 1, GARCH 2)
//   fine[2][B][N]      R   tile-local inclusive scan of exp(lw - m_tile)   ("fine" CDF)
//   lw  [2][B][N]      R   log-weights (only written when a consumer needs them)
//   part[2][B][G][8]   f64 per-tile partials: m_tile, s_tile = sum exp(lw - m_tile), ws[0..3]
//   acc [B][8]         f64 running log-likelihood (+ filter statistic)
// Double-buffered on step parity: step t reads [t & 1] and writes [(t + 1) & 1].
//
// The global CDF of an item is never materialised: c_i = coarse[g] + fine_i * exp(m_g - M) where
// coarse[] (f64, G + 1 entries) is rebuilt in shared memory by every CTA from `part` (flash-style
// rescaling of tile-local sums).  A search is a binary search over coarse[] in shared memory followed
// by a binary search inside one tile of `fine`.
#pragma once
#include "blockops.cuh"
#include "models.cuh"
#include "rng.cuh"

namespace sgm {

constexpr int KPT = 8;                 // consecutive particles per thread
constexpr int TILE = NT * KPT;         // 2048 particles per CTA tile
constexpr int MAX_TILES = 1024;        // N <= 2^21
constexpr int PSTRIDE = 8;             // doubles per `part` entry
constexpr int ACC_STRIDE = 8;          // doubles per `acc` entry

struct KArgs {
    int B, N, G, max_T;
    int pf, rng_mode, resample, stat_kind, Ntilde, accept_reject, max_ar, manual_thresh;
    int need_lw;
    double lambduh;
    RngKey key;            // .item holds item_id_base
    const double* obs; const int64_t* obs_off; const int32_t* T_buf; const int32_t* t1; const int32_t* tL;
    const double* step_weights; const int64_t* wts_off; const double* theta;
    const double* prior_mean; const double* prior_var;
    const double* inj_z0; const double* inj_u; const double* inj_z; const double* inj_extra; const int64_t* inj_extra_off;
    void* rec[2]; void* tail[2]; void* fine[2]; void* lw[2]; double* part[2]; double* acc;
    int32_t* Jidx; int32_t* Llist[2]; int32_t* counters;
    double* grad; double* loglik; int32_t* status;
    void* out_x; void* out_lw; void* out_stats; int32_t* trace_anc; void* trace_x; void* trace_lw; int32_t* trace_J;
};

// ---- record access ---------------------------------------------------------------------------
template <class R> struct alignas(4 * sizeof(R)) Vec4T { R x, y, z, w; };

template <class R, int W> __device__ __forceinline__ void load_rec(const void* rec, const void* tail, size_t idx, R* r) {
    const Vec4T<R> v = reinterpret_cast<const Vec4T<R>*>(rec)[idx];
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    if (W == 5) r[4] = reinterpret_cast<const R*>(tail)[idx];
    if (W == 6) { const R* t = reinterpret_cast<const R*>(tail) + 2 * idx; r[4] = t[0]; r[5] = t[1]; }
}
template <class R, int W> __device__ __forceinline__ void store_rec(void* rec, void* tail, size_t idx, const R* r) {
    Vec4T<R> v; v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3];
    reinterpret_cast<Vec4T<R>*>(rec)[idx] = v;
    if (W == 5) reinterpret_cast<R*>(tail)[idx] = r[4];
    if (W == 6) { R* t = reinterpret_cast<R*>(tail) + 2 * idx; t[0] = r[4]; t[1] = r[5]; }
}

// ---- shared-memory CDF header built by every CTA --------------------------------------------------
struct CdfHeader {
    double coarse[MAX_TILES + 1];   // exclusive prefix of tile masses in units of exp(-M); [G] = total
    double einv[MAX_TILES];         // exp(M - m_g)
    double red[NWARP];
    double M, total;
    double sbar[4];
};

// Builds hdr from part[]; returns false if the weights are degenerate (status flagged by caller).
// Also returns (through hdr.sbar) sum_g e_g * ws_g[k] / total for k < nws.
__device__ __forceinline__ void build_cdf_header(const double* __restrict__ part, int G, int nws, CdfHeader& hdr) {
    const int tid = threadIdx.x;
    double m = -Mth<double>::inf();
    for (int g = tid; g < G; g += NT) m = nan_max(m, part[(size_t)g * PSTRIDE]);
    const double M = block_max(m, hdr.red);
    const int per = (G + NT - 1) / NT;
    const int g0 = tid * per;
    double loc = 0.0, ws[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < per; ++k) {
        const int g = g0 + k;
        if (g < G) {
            const double* p = part + (size_t)g * PSTRIDE;
            const double e = (p[0] == -Mth<double>::inf()) ? 0.0 : ::exp(p[0] - M);
            loc += e * p[1];
            for (int q = 0; q < nws; ++q) ws[q] += e * p[2 + q];
        }
    }
    double total;
    double run = block_excl_scan(loc, hdr.red, total);
    for (int k = 0; k < per; ++k) {
        const int g = g0 + k;
        if (g < G) {
            const double* p = part + (size_t)g * PSTRIDE;
            const double e = (p[0] == -Mth<double>::inf()) ? 0.0 : ::exp(p[0] - M);
            hdr.coarse[g] = run;
            hdr.einv[g] = 1.0 / e;
            run += e * p[1];
        }
    }
    for (int q = 0; q < nws; ++q) {
        const double s = block_sum(ws[q], hdr.red);
        if (tid == 0) hdr.sbar[q] = s / total;
    }
    if (tid == 0) { hdr.coarse[G] = total; hdr.M = M; hdr.total = total; }
    __syncthreads();
}

// searchsorted(cdf, u, side='right') on the hierarchical CDF: first index whose cumulative mass
// exceeds target = u * total.
template <class R>
__device__ __forceinline__ int search_cdf(double target, const CdfHeader& hdr, int G, const R* __restrict__ fine, int N) {
    if (!(target < hdr.total)) target = hdr.total * (1.0 - 1.2e-16);
    int lo = 0, hi = G;
    while (lo < hi) {                           // tile whose [coarse[g], coarse[g+1]) contains target
        const int mid = (lo + hi) >> 1;
        if (hdr.coarse[mid + 1] <= target) lo = mid + 1; else hi = mid;
    }
    const int g = min(lo, G - 1);
    const R r = (R)((target - hdr.coarse[g]) * hdr.einv[g]);
    const int base = g * TILE;
    const int len = min(TILE, N - base);
    const R* f = fine + base;
    lo = 0; hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (f[mid] <= r) lo = mid + 1; else hi = mid;
    }
    return base + min(lo, len - 1);
}

// ---- per-tile epilogue: tile max, tile-local scan of exp(lw - m), per-tile partials ---------------
template <class R, int W, int NP>
__device__ __forceinline__ void tile_epilogue(const R* lwn, int i0, int N, R* fine_out, double* part_out,
                                              const void* rec_new, const void* tail_new, size_t item_off,
                                              bool need_ws, int nws, R* sh_r, double* sh_d) {
    R m = -Mth<R>::inf();
#pragma unroll
    for (int c = 0; c < KPT; ++c) if (i0 + c < N) m = nan_max(m, lwn[c]);
    m = block_max(m, sh_r);
    R w[KPT], run = (R)0;
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        w[c] = (i0 + c < N) ? ((m == -Mth<R>::inf()) ? (R)0 : Mth<R>::exp(lwn[c] - m)) : (R)0;
        run += w[c];
    }
    R total;
    R pre = block_excl_scan(run, sh_r, total);
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        pre += w[c];
        if (i0 + c < N) fine_out[i0 + c] = pre;
    }
    double ws[4] = {0.0, 0.0, 0.0, 0.0};
    if (need_ws) {
#pragma unroll
        for (int c = 0; c < KPT; ++c) {
            if (i0 + c < N) {
                R r[W];
                load_rec<R, W>(rec_new, tail_new, item_off + i0 + c, r);
                for (int q = 0; q < nws; ++q) ws[q] += (double)(r[q] * w[c]);
            }
        }
        for (int q = 0; q < nws; ++q) ws[q] = block_sum(ws[q], sh_d);
    }
    if (threadIdx.x == 0) {
        part_out[0] = (double)m;
        part_out[1] = (double)total;
        for (int q = 0; q < 4; ++q) part_out[2 + q] = ws[q];
    }
}

template <class Model> __device__ __forceinline__ int stat_width(int stat_kind) {
    return stat_kind == SGM_STAT_SCORE ? Model::NP : (stat_kind == SGM_STAT_SUFF ? 3 : 0);
}

// ---- init: x0 ~ N(prior_mean, prior_var), lw = 0, stats = 0  (buffered_smoother.py:67-75) ----------
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_init_kernel(KArgs a) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x;
    const int N = a.N;
    const size_t item_off = (size_t)b * N;
    const int i0 = g * TILE + tid * KPT;
    RngKey key = a.key; key.item += (uint32_t)b;
    const R mean = (R)a.prior_mean[b], sd = (R)::sqrt(a.prior_var[b]);
    R lwn[KPT];
    R z[KPT];
    if (a.rng_mode == SGM_RNG_PHILOX) {
        rng_normal4(key, (uint32_t)(i0 >> 2), 0xffffu, z);
        rng_normal4(key, (uint32_t)(i0 >> 2) + 1u, 0xffffu, z + 4);
    } else {
#pragma unroll
        for (int c = 0; c < KPT; ++c) z[c] = (i0 + c < N) ? (R)a.inj_z0[item_off + i0 + c] : (R)0;
    }
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        lwn[c] = (R)0;
        const int i = i0 + c;
        if (i < N) {
            R r[W];
#pragma unroll
            for (int q = 0; q < W; ++q) r[q] = (R)0;
            Model::init(mean, sd, z[c], r + NP);
            store_rec<R, W>(a.rec[0], a.tail[0], item_off + i, r);
            if (a.need_lw) reinterpret_cast<R*>(a.lw[0])[item_off + i] = (R)0;
            if (a.trace_x) {
                R* tx = reinterpret_cast<R*>(a.trace_x) + ((size_t)b * (a.max_T + 1) * N + i) * NX;
                for (int q = 0; q < NX; ++q) tx[q] = r[NP + q];
            }
            if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[(size_t)b * (a.max_T + 1) * N + i] = (R)0;
        }
    }
    tile_epilogue<R, W, NP>(lwn, i0, N, reinterpret_cast<R*>(a.fine[0]) + item_off,
                            a.part[0] + ((size_t)b * a.G + g) * PSTRIDE, a.rec[0], a.tail[0], item_off,
                            false, 0, sh_r, sh_d);
    if (g == 0 && tid == 0) {
        for (int q = 0; q < ACC_STRIDE; ++q) a.acc[(size_t)b * ACC_STRIDE + q] = 0.0;
        a.status[b] = 0;
        if (a.counters) a.counters[b * 16] = 0;
    }
}

// Bookkeeping done once per item per step by (tile 0, thread 0): log-likelihood increment of the step
// that produced the current weights (buffered_smoother.py:124-126; here with the max shift, i.e.
// M + log(total / N)), filter statistic (pf.py:77-80), degeneracy flags.
__device__ __forceinline__ void item_bookkeeping(const KArgs& a, int b, int t_done, const CdfHeader& hdr, int nws) {
    if (!(hdr.total > 0.0) || !(hdr.total < Mth<double>::inf()) || !(hdr.M == hdr.M) || !(fabs(hdr.M) < Mth<double>::inf())) {
        a.status[b] |= (hdr.total == 0.0 || hdr.M == -Mth<double>::inf()) ? SGM_STATUS_ZERO_WEIGHT : SGM_STATUS_NAN_WEIGHT;
    }
    if (t_done < 0) return;
    double* acc = a.acc + (size_t)b * ACC_STRIDE;
    if (t_done >= a.t1[b] && t_done < a.tL[b]) {
        const double wt = (a.wts_off && a.wts_off[b] >= 0) ? a.step_weights[a.wts_off[b] + (t_done - a.t1[b])] : 1.0;
        acc[0] += wt * (hdr.M + ::log(hdr.total / (double)a.N));
    }
    if (a.pf == SGM_PF_FILTER) for (int q = 0; q < nws; ++q) acc[1 + q] += hdr.sbar[q];
}

// ---- one resample -> propagate -> reweight -> statistic-update step (pf.py:7-38, 138-181, 40-82) ---
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_step_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ CdfHeader hdr;
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
    __shared__ double sh_gam[2];
    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x;
    const int Tb = a.T_buf[b];
    if (t >= Tb) return;
    const int N = a.N, G = a.G, par = t & 1;
    const size_t item_off = (size_t)b * N;
    const int nws = stat_width<Model>(a.stat_kind);
    const bool carries = (a.pf == SGM_PF_NEMETH);            // stats follow the resampled genealogy here
    const bool shrink = carries && (a.lambduh != 1.0);
    const bool hdr_ws = shrink || (a.pf == SGM_PF_FILTER);

    build_cdf_header(a.part[par] + (size_t)b * G * PSTRIDE, G, hdr_ws ? nws : 0, hdr);
    if (g == 0 && tid == 0) item_bookkeeping(a, b, t - 1, hdr, nws);

    const R* fine_old = reinterpret_cast<const R*>(a.fine[par]) + item_off;
    const int i0 = g * TILE + tid * KPT;
    RngKey key = a.key; key.item += (uint32_t)b;
    const double total = hdr.total;

    // ---- uniforms for the resampling step -------------------------------------------------------
    double target[KPT];
    if (a.rng_mode == SGM_RNG_INJECTED) {
        const double* u = a.inj_u + ((size_t)b * a.max_T + t) * N;
#pragma unroll
        for (int c = 0; c < KPT; ++c) target[c] = (i0 + c < N) ? u[i0 + c] * total : 0.0;
    } else if (a.resample == SGM_RESAMPLE_MULTINOMIAL) {
        R u[KPT];
        rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, u);
        rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, u + 4);
#pragma unroll
        for (int c = 0; c < KPT; ++c) target[c] = (double)u[c] * total;
    } else if (a.resample == SGM_RESAMPLE_MULTINOMIAL_SORTED) {
        // Order statistics of N iid uniforms via exponential spacings: within a tile the normalised
        // partial sums of P Exp(1) draws are independent of their total, which is Gamma(P, 1); so the
        // tile totals are drawn directly (one Gamma per tile) and no cross-tile pass is needed.
        R e[KPT], run = (R)0;
        rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, e);
        rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, e + 4);
#pragma unroll
        for (int c = 0; c < KPT; ++c) { e[c] = (i0 + c < N) ? -Mth<R>::log(e[c]) : (R)0; run += e[c]; }
        R tile_sum;
        R pre = block_excl_scan(run, sh_r, tile_sum);
        // gamma prefix for this tile: sum of Gamma(P_g') for g' < g, and the grand total
        double gl = 0.0, gt = 0.0;
        for (int g2 = tid; g2 <= G; g2 += NT) {
            const double shape = (g2 == G) ? 1.0 : (double)min(TILE, N - g2 * TILE);
            const double gam = rng_gamma(key, (uint32_t)g2, (uint32_t)t, shape);
            gt += gam;
            if (g2 < g) gl += gam;
            if (g2 == g) sh_gam[0] = gam;
        }
        gl = block_sum(gl, sh_d);
        gt = block_sum(gt, sh_d);
        const double scale = total / gt, gmine = sh_gam[0] / (double)tile_sum;
#pragma unroll
        for (int c = 0; c < KPT; ++c) { pre += e[c]; target[c] = (gl + gmine * (double)pre) * scale; }
    } else {
        // systematic: u_i = (i + U) / N with one U per item-step; stratified: u_i = (i + U_i) / N
        R u[KPT];
        if (a.resample == SGM_RESAMPLE_SYSTEMATIC) {
            R u4[4];
            rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, u4);
#pragma unroll
            for (int c = 0; c < KPT; ++c) u[c] = u4[0];
        } else {
            rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, u);
            rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, u + 4);
        }
#pragma unroll
        for (int c = 0; c < KPT; ++c) target[c] = ((double)(i0 + c) + (double)u[c]) / (double)N * total;
    }

    // ---- proposal normals ------------------------------------------------------------------------
    R z[KPT];
    if (a.rng_mode == SGM_RNG_INJECTED) {
        const double* zz = a.inj_z + ((size_t)b * a.max_T + t) * N;
#pragma unroll
        for (int c = 0; c < KPT; ++c) z[c] = (i0 + c < N) ? (R)zz[i0 + c] : (R)0;
    } else {
        rng_normal4(key, (uint32_t)(i0 >> 2), (uint32_t)t, z);
        rng_normal4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, z + 4);
    }

    const typename Model::template Theta<R> th = Model::template load<R>(a.theta + (size_t)b * SGM_THETA_STRIDE);
    const R y = (R)a.obs[a.obs_off[b] + t];
    const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
    const R wt = in_sub ? ((a.wts_off && a.wts_off[b] >= 0) ? (R)a.step_weights[a.wts_off[b] + (t - a.t1[b])] : (R)1) : (R)0;
    const R lam = (R)a.lambduh;
    R sbar[4] = {(R)0, (R)0, (R)0, (R)0};
    if (shrink) for (int q = 0; q < nws; ++q) sbar[q] = (R)((1.0 - a.lambduh) * hdr.sbar[q]);

    void* rec_new = a.rec[par ^ 1];
    void* tail_new = a.tail[par ^ 1];
    R lwn[KPT];
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        const int i = i0 + c;
        lwn[c] = (R)0;
        if (i < N) {
            const int anc = search_cdf<R>(target[c], hdr, G, fine_old, N);
            R ra[W], rn[W];
            load_rec<R, W>(a.rec[par], a.tail[par], item_off + anc, ra);
            Model::propagate(th, ra + NP, y, z[c], rn + NP);
            lwn[c] = Model::log_weight(th, ra + NP, rn + NP, y);
            R h[4] = {(R)0, (R)0, (R)0, (R)0};
            if (in_sub) {
                if (a.stat_kind == SGM_STAT_SCORE) Model::score(th, ra + NP, rn + NP, y, h);
                else if (a.stat_kind == SGM_STAT_SUFF) Model::suff(ra + NP, rn + NP, h);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (carries) rn[q] = shrink ? (lam * ra[q] + sbar[q] + h[q] * wt) : (ra[q] + h[q] * wt);   // pf.py:175-179
                else if (a.pf == SGM_PF_FILTER) rn[q] = h[q] * wt;                                         // pf.py:70-71
                else rn[q] = (R)0;                                                                           // set by the backward kernel
            }
            store_rec<R, W>(rec_new, tail_new, item_off + i, rn);
            if (a.need_lw) reinterpret_cast<R*>(a.lw[par ^ 1])[item_off + i] = lwn[c];
            if (a.trace_anc) a.trace_anc[((size_t)b * a.max_T + t) * N + i] = anc;
            if (a.trace_x) {
                R* tx = reinterpret_cast<R*>(a.trace_x) + (((size_t)b * (a.max_T + 1) + t + 1) * N + i) * NX;
                for (int q = 0; q < NX; ++q) tx[q] = rn[NP + q];
            }
            if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[((size_t)b * (a.max_T + 1) + t + 1) * N + i] = lwn[c];
        }
    }
    const bool need_ws = (nws > 0) && (a.pf == SGM_PF_FILTER || shrink || (carries && t == Tb - 1));
    tile_epilogue<R, W, NP>(lwn, i0, N, reinterpret_cast<R*>(a.fine[par ^ 1]) + item_off,
                            a.part[par ^ 1] + ((size_t)b * G + g) * PSTRIDE, rec_new, tail_new, item_off,
                            need_ws, nws, sh_r, sh_d);
}

// ---- final: last log-likelihood term + average_statistic (buffered_smoother.py:151-154) -----------
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_final_kernel(KArgs a) {
    __shared__ CdfHeader hdr;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int Tb = a.T_buf[b], par = Tb & 1, G = a.G;
    const int nws = stat_width<Model>(a.stat_kind);
    build_cdf_header(a.part[par] + (size_t)b * G * PSTRIDE, G, nws, hdr);
    if (tid == 0) {
        item_bookkeeping(a, b, Tb - 1, hdr, nws);
        const double* acc = a.acc + (size_t)b * ACC_STRIDE;
        a.loglik[b] = acc[0];
        for (int q = 0; q < 8; ++q) a.grad[(size_t)b * 8 + q] = 0.0;
        for (int q = 0; q < nws; ++q) a.grad[(size_t)b * 8 + q] = (a.pf == SGM_PF_FILTER) ? acc[1 + q] : hdr.sbar[q];
    }
}

// ---- optional export of the final particle system (out['x_t'], ['log_weights'], ['statistics']) ---
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_export_kernel(KArgs a) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int b = blockIdx.y;
    const int i = blockIdx.x * NT + threadIdx.x;
    if (i >= a.N) return;
    const int par = a.T_buf[b] & 1;
    const size_t idx = (size_t)b * a.N + i;
    R r[W];
    load_rec<R, W>(a.rec[par], a.tail[par], idx, r);
    if (a.out_x) for (int q = 0; q < NX; ++q) reinterpret_cast<R*>(a.out_x)[idx * NX + q] = r[NP + q];
    if (a.out_stats) for (int q = 0; q < NP; ++q) reinterpret_cast<R*>(a.out_stats)[idx * NP + q] = r[q];
    if (a.out_lw) reinterpret_cast<R*>(a.out_lw)[idx] = reinterpret_cast<const R*>(a.lw[par])[idx];
}

}  // namespace sgm
