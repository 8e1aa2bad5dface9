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
constexpr int MAX_TILES = 512;         // N <= 2^20
constexpr int CAP = 4096;              // parents staged in shared memory per CTA (sorted resampling)
constexpr int PSTRIDE = 8;             // doubles per `part` entry
constexpr int ACC_STRIDE = 8;          // doubles per `acc` entry
constexpr int THC_BYTES = 128;         // per-item slot for the model's derived constants

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
    const double* gam;     // [B][max_T][G + 2] exclusive prefix of per-tile Gamma draws (sorted multinomial)
    char* thc;             // [B][THC_BYTES]   Model::Theta<R> (derived constants) written by the init kernel
    void* yw;              // [B][max_T][2] R  (y_t, statistic weight of step t: w_t inside [t1, tL), else 0)
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
    double e[MAX_TILES];            // exp(m_g - M)
    double einv[MAX_TILES];         // exp(M - m_g)
    double M, total;
    double sbar[4];
};

// Built by warp 0 with shuffles only; the CALLER issues the __syncthreads() that publishes it.  hdr.sbar[k] receives
// sum_g e_g * ws_g[k] / total for k < nws.  Fixed summation order -> deterministic.
__device__ __forceinline__ void build_cdf_header(const double* __restrict__ part, int G, int nws, CdfHeader& hdr) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double m = -Mth<double>::inf();
        for (int g = lane; g < G; g += 32) m = nan_max(m, part[(size_t)g * PSTRIDE]);
        const double M = warp_max(m);
        double carry = 0.0, ws[4] = {0.0, 0.0, 0.0, 0.0};
        for (int g0 = 0; g0 < G; g0 += 32) {
            const int g = g0 + lane;
            double e = 0.0, v = 0.0;
            if (g < G) {
                const double* p = part + (size_t)g * PSTRIDE;
                e = (p[0] == -Mth<double>::inf()) ? 0.0 : ::exp(p[0] - M);
                v = e * p[1];
                for (int q = 0; q < nws; ++q) ws[q] += e * p[2 + q];
            }
            const double incl = warp_incl_scan(v);
            if (g < G) { hdr.coarse[g] = carry + (incl - v); hdr.e[g] = e; hdr.einv[g] = 1.0 / e; }
            carry += __shfl_sync(FULL, incl, 31);
        }
        for (int q = 0; q < nws; ++q) {
            const double sq = warp_sum(ws[q]);
            if (lane == 0) hdr.sbar[q] = sq / carry;
        }
        if (lane == 0) { hdr.coarse[G] = carry; hdr.M = M; hdr.total = carry; }
    }
}

// tile g with coarse[g] <= target < coarse[g + 1]  (branch-free, CTA-uniform trip count)
__device__ __forceinline__ int coarse_search(double target, const CdfHeader& hdr, int G, int step0) {
    int pos = 0;
    for (int step = step0; step > 0; step >>= 1)
        if (pos + step <= G && hdr.coarse[pos + step] <= target) pos += step;
    return min(pos, G - 1);
}
__device__ __forceinline__ int pow2_floor(int x) { return 1 << (31 - __clz(max(x, 1))); }

// searchsorted(cdf, u, side='right') on the hierarchical CDF: first index whose cumulative mass
// exceeds target = u * total.  One thread, dependent loads (used off the hot path).
template <class R>
__device__ __forceinline__ int search_cdf(double target, const CdfHeader& hdr, int G, const R* __restrict__ fine, int N) {
    if (!(target < hdr.total)) target = hdr.total * (1.0 - 1.2e-16);
    const int g = coarse_search(target, hdr, G, pow2_floor(G));
    const R r = (R)((target - hdr.coarse[g]) * hdr.einv[g]);
    const int base = g * TILE;
    const int len = min(TILE, N - base);
    const R* f = fine + base;
    int pos = 0;
#pragma unroll
    for (int step = TILE / 2; step > 0; step >>= 1)
        if (pos + step <= len && f[pos + step - 1] <= r) pos += step;
    return base + min(pos, len - 1);
}

// Same search executed cooperatively by a full warp (all lanes pass the same target):
// 32-ary probing, three dependent loads instead of eleven.
template <class R>
__device__ __forceinline__ int warp_search_cdf(double target, const CdfHeader& hdr, int G, const R* __restrict__ fine, int N) {
    const int lane = threadIdx.x & 31;
    if (!(target < hdr.total)) target = hdr.total * (1.0 - 1.2e-16);
    const int g = coarse_search(target, hdr, G, pow2_floor(G));
    const R r = (R)((target - hdr.coarse[g]) * hdr.einv[g]);
    const int base = g * TILE;
    const int len = min(TILE, N - base);
    const R* f = fine + base;
    constexpr int SEG = TILE / 32;                                   // 64
    bool le = (lane * SEG < len) ? (f[min(len, (lane + 1) * SEG) - 1] <= r) : false;
    const int s1 = __popc(__ballot_sync(FULL, le)) * SEG;
    le = (s1 + lane * 2 < len) ? (f[min(s1 + lane * 2 + 1, len - 1)] <= r) : false;
    const int s2 = s1 + __popc(__ballot_sync(FULL, le)) * 2;
    int res = s2;
    if (s2 < len && f[s2] <= r) res = s2 + 1;
    return base + min(res, len - 1);
}

// ---- per-tile epilogue: tile max, tile-local scan of exp(lw - m), per-tile partials ---------------
template <class R, int W, int NP>
__device__ __forceinline__ void tile_epilogue(const R* lwn, int i0, int N, R* fine_out, double* part_out,
                                              const void* rec_new, const void* tail_new, size_t item_off,
                                              bool need_ws, int nws, R* sh_r, double* sh_d) {
    // callers set lwn[c] = -inf for children beyond N, so they carry zero weight below
    R m = -Mth<R>::inf();
#pragma unroll
    for (int c = 0; c < KPT; ++c) m = nan_max(m, lwn[c]);
    m = block_max(m, sh_r);
    // m == -inf (every weight of the tile zero): shift by 0 instead, exp(-inf) = 0 and a NaN log-weight
    // still poisons the tile sum so that the item gets flagged
    const R msafe = (m == -Mth<R>::inf()) ? (R)0 : m;
    R w[KPT], run = (R)0;
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        w[c] = Mth<R>::exp(lwn[c] - msafe);
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
        const int i = i0 + c;
        lwn[c] = (i < N) ? (R)0 : -Mth<R>::inf();
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
    if (g == 0) {
        // per-item derived constants and per-step (observation, statistic weight) pairs, read by every
        // later kernel with one vector load instead of pointer chasing + double-precision log / div
        const int Tb = a.T_buf[b];
        for (int t = tid; t < Tb; t += NT) {
            const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
            const double wt = in_sub ? ((a.wts_off && a.wts_off[b] >= 0) ? a.step_weights[a.wts_off[b] + (t - a.t1[b])] : 1.0) : 0.0;
            R* yw = reinterpret_cast<R*>(a.yw) + ((size_t)b * a.max_T + t) * 2;
            yw[0] = (R)a.obs[a.obs_off[b] + t];
            yw[1] = (R)wt;
        }
        if (tid == 0) {
            *reinterpret_cast<typename Model::template Theta<R>*>(a.thc + (size_t)b * THC_BYTES) =
                Model::template load<R>(a.theta + (size_t)b * SGM_THETA_STRIDE);
            for (int q = 0; q < ACC_STRIDE; ++q) a.acc[(size_t)b * ACC_STRIDE + q] = 0.0;
            a.status[b] = 0;
            if (a.counters) a.counters[b * 16] = 0;
        }
    }
}

template <class R, class Model>
__device__ __forceinline__ typename Model::template Theta<R> load_thc(const KArgs& a, int b) {
    return *reinterpret_cast<const typename Model::template Theta<R>*>(a.thc + (size_t)b * THC_BYTES);
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

// ---- gather parents -> propagate -> reweight -> statistic update -> store (pf.py:30-36, 168-179) -----
template <class R, class Model>
__device__ __forceinline__ void propagate_store(const KArgs& a, int b, int t, int par, int i0, size_t item_off,
                                                const int* anc, const R* z, const CdfHeader& hdr, int nws,
                                                bool carries, bool shrink, R* lwn) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int N = a.N;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const R* ywp = reinterpret_cast<const R*>(a.yw) + ((size_t)b * a.max_T + t) * 2;
    const R y = ywp[0], wt = ywp[1];
    const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
    const R lam = (R)a.lambduh;
    R sbar[4] = {(R)0, (R)0, (R)0, (R)0};
    if (shrink) for (int q = 0; q < nws; ++q) sbar[q] = (R)((1.0 - a.lambduh) * hdr.sbar[q]);
    const bool tracing = a.need_lw || a.trace_anc || a.trace_x || a.trace_lw;
    // new statistic = keep * parent statistic + sbar + h * hs   (one FMA chain for every smoother):
    //   Poyiadjis O(N)/Nemeth: keep = lambduh, hs = w_t (pf.py:175-179); filter: keep = 0, hs = w_t
    //   (pf.py:70-71); O(N^2)/PaRIS: keep = hs = 0 (the backward kernel writes the statistic)
    const R keep = carries ? (shrink ? lam : (R)1) : (R)0;
    const R hs = (carries || a.pf == SGM_PF_FILTER) ? wt : (R)0;
    const int stat_kind = (in_sub && hs != (R)0) ? a.stat_kind : SGM_STAT_NONE;
    const void* rec_old = a.rec[par];
    const void* tail_old = a.tail[par];
    void* rec_new = a.rec[par ^ 1];
    void* tail_new = a.tail[par ^ 1];
#pragma unroll
    for (int h0 = 0; h0 < KPT; h0 += 4) {
        R ra[4][W];
#pragma unroll
        for (int c = 0; c < 4; ++c)                       // four independent parent gathers in flight
            if (i0 + h0 + c < N) load_rec<R, W>(rec_old, tail_old, item_off + anc[h0 + c], ra[c]);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const int c = h0 + c4, i = i0 + c;
            lwn[c] = -Mth<R>::inf();
            if (i < N) {
                R rn[W];
                Model::propagate(th, ra[c4] + NP, y, z[c], rn + NP);
                lwn[c] = Model::log_weight(th, ra[c4] + NP, rn + NP, y);
                R h[4] = {(R)0, (R)0, (R)0, (R)0};
                if (stat_kind == SGM_STAT_SCORE) Model::score(th, ra[c4] + NP, rn + NP, y, h);
                else if (stat_kind == SGM_STAT_SUFF) Model::suff(ra[c4] + NP, rn + NP, h);
#pragma unroll
                for (int q = 0; q < NP; ++q) rn[q] = keep * ra[c4][q] + (sbar[q] + h[q] * hs);
                store_rec<R, W>(rec_new, tail_new, item_off + i, rn);
                if (tracing) {
                    if (a.need_lw) reinterpret_cast<R*>(a.lw[par ^ 1])[item_off + i] = lwn[c];
                    if (a.trace_anc) a.trace_anc[((size_t)b * a.max_T + t) * N + i] = anc[c];
                    if (a.trace_x) {
                        R* tx = reinterpret_cast<R*>(a.trace_x) + (((size_t)b * (a.max_T + 1) + t + 1) * N + i) * NX;
                        for (int q = 0; q < NX; ++q) tx[q] = rn[NP + q];
                    }
                    if (a.trace_lw) reinterpret_cast<R*>(a.trace_lw)[((size_t)b * (a.max_T + 1) + t + 1) * N + i] = lwn[c];
                }
            }
        }
    }
}

template <class R>
__device__ __forceinline__ void draw_normals(const KArgs& a, const RngKey& key, int b, int t, int i0, R* z) {
    if (a.rng_mode == SGM_RNG_INJECTED) {
        const double* zz = a.inj_z + ((size_t)b * a.max_T + t) * a.N;
#pragma unroll
        for (int c = 0; c < KPT; ++c) z[c] = (i0 + c < a.N) ? (R)zz[i0 + c] : (R)0;
    } else {
        rng_normal4(key, (uint32_t)(i0 >> 2), (uint32_t)t, z);
        rng_normal4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, z + 4);
    }
}

// ---- step kernel, iid resampling uniforms (reference semantics, pf.py:27-29) --------------------------
// Per-child coarse (shared) + fine (global) branch-free binary searches, the 8 children of a thread
// interleaved for memory-level parallelism.
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_step_kernel(KArgs a, int t) {
    constexpr int NP = Model::NP, W = Model::NX + NP;
    __shared__ CdfHeader hdr;
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
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
    __syncthreads();
    if (g == 0 && tid == 0) item_bookkeeping(a, b, t - 1, hdr, nws);

    const R* __restrict__ fine_old = reinterpret_cast<const R*>(a.fine[par]) + item_off;
    const int i0 = g * TILE + tid * KPT;
    RngKey key = a.key; key.item += (uint32_t)b;
    const double total = hdr.total;

    double target[KPT];
    if (a.rng_mode == SGM_RNG_INJECTED) {
        const double* u = a.inj_u + ((size_t)b * a.max_T + t) * N;
#pragma unroll
        for (int c = 0; c < KPT; ++c) target[c] = (i0 + c < N) ? u[i0 + c] * total : 0.0;
    } else {
        R u[KPT];
        rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, u);
        rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, u + 4);
#pragma unroll
        for (int c = 0; c < KPT; ++c) target[c] = (double)u[c] * total;
    }
    int anc[KPT];
    {
        int gsel[KPT], pos[KPT], len[KPT];
        R rr[KPT];
        const int step0 = pow2_floor(G);
#pragma unroll
        for (int c = 0; c < KPT; ++c) {
            if (!(target[c] < total)) target[c] = total * (1.0 - 1.2e-16);
            gsel[c] = coarse_search(target[c], hdr, G, step0);
            rr[c] = (R)((target[c] - hdr.coarse[gsel[c]]) * hdr.einv[gsel[c]]);
            len[c] = min(TILE, N - gsel[c] * TILE);
            pos[c] = 0;
        }
#pragma unroll 1
        for (int step = TILE / 2; step > 0; step >>= 1) {
#pragma unroll
            for (int c = 0; c < KPT; ++c) {
                const int idx = pos[c] + step;
                if (idx <= len[c] && fine_old[gsel[c] * TILE + idx - 1] <= rr[c]) pos[c] = idx;
            }
        }
#pragma unroll
        for (int c = 0; c < KPT; ++c) anc[c] = gsel[c] * TILE + min(pos[c], len[c] - 1);
    }
    R z[KPT], lwn[KPT];
    draw_normals<R>(a, key, b, t, i0, z);
    propagate_store<R, Model>(a, b, t, par, i0, item_off, anc, z, hdr, nws, carries, shrink, lwn);
    const bool need_ws = (nws > 0) && (a.pf == SGM_PF_FILTER || shrink || (carries && t == Tb - 1));
    tile_epilogue<R, W, NP>(lwn, i0, N, reinterpret_cast<R*>(a.fine[par ^ 1]) + item_off,
                            a.part[par ^ 1] + ((size_t)b * G + g) * PSTRIDE, a.rec[par ^ 1], a.tail[par ^ 1], item_off,
                            need_ws, nws, sh_r, sh_d);
}

// ---- step kernel, ascending resampling targets ---------------------------------------------------------
// (order-statistics multinomial / systematic / stratified, or INJECTED uniforms the caller declares
// sorted).  The CTA's 2048 children hit ONE contiguous parent range [lo, hi]: two warp-cooperative
// searches find it, the CDF of the range is staged in shared memory in global units, and every thread
// merges its 8 consecutive children against it, so parent records are gathered as a stream.
template <class R, class Model>
__global__ void __launch_bounds__(NT, (sizeof(R) == 4 ? 4 : 2)) pf_step_sorted_kernel(KArgs a, int t) {
    constexpr int NP = Model::NP, W = Model::NX + NP;
    __shared__ CdfHeader hdr;
    __shared__ R s_cdf[CAP + 16];
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
    __shared__ int s_range[2];
    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Tb = a.T_buf[b];
    if (t >= Tb) return;
    const int N = a.N, G = a.G, par = t & 1;
    const size_t item_off = (size_t)b * N;
    const int nws = stat_width<Model>(a.stat_kind);
    const bool carries = (a.pf == SGM_PF_NEMETH);
    const bool shrink = carries && (a.lambduh != 1.0);
    const bool hdr_ws = shrink || (a.pf == SGM_PF_FILTER);
    const R* __restrict__ fine_old = reinterpret_cast<const R*>(a.fine[par]) + item_off;
    const int i0 = g * TILE + tid * KPT;
    const int n_valid = min(TILE, N - g * TILE);
    RngKey key = a.key; key.item += (uint32_t)b;

    // warp 0 rebuilds the CDF header while the other warps draw their randoms
    build_cdf_header(a.part[par] + (size_t)b * G * PSTRIDE, G, hdr_ws ? nws : 0, hdr);

    // u[c]: tile-local position in (0, 1] of child c for the order-statistics sampler, else the uniform
    R u[KPT], z[KPT];
    double gl = 0.0, gw = 0.0, scale_over_total = 0.0;
    const bool spacings = (a.rng_mode == SGM_RNG_PHILOX) && (a.resample == SGM_RESAMPLE_MULTINOMIAL_SORTED);
    if (spacings) {
        // Order statistics of N iid uniforms via exponential spacings.  Within a tile the normalised
        // partial sums of P Exp(1) draws are independent of their total, which is Gamma(P, 1); the tile
        // totals are drawn directly (gamma_prefix_kernel), so no cross-tile scan is needed.
        const double* gam = a.gam + ((size_t)b * a.max_T + t) * (G + 2);
        gl = gam[g]; gw = gam[g + 1] - gl; scale_over_total = 1.0 / gam[G + 1];
        R run = (R)0;
        rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, u);
        rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, u + 4);
#pragma unroll
        for (int c = 0; c < KPT; ++c) { run += (i0 + c < N) ? -Mth<R>::log(u[c]) : (R)0; u[c] = run; }
        const R incl = warp_incl_scan(run);
        if (lane == 31) sh_r[warp] = incl;
        draw_normals<R>(a, key, b, t, i0, z);
        __syncthreads();                                   // header + warp totals
        R base = (R)0, tile_sum = (R)0;
#pragma unroll
        for (int k = 0; k < NWARP; ++k) { const R sk = sh_r[k]; if (k < warp) base += sk; tile_sum += sk; }
        base += incl - run;
        const R inv = Mth<R>::rcp(tile_sum);
#pragma unroll
        for (int c = 0; c < KPT; ++c) u[c] = (base + u[c]) * inv;
    } else {
        if (a.rng_mode == SGM_RNG_INJECTED) {
#pragma unroll
            for (int c = 0; c < KPT; ++c) u[c] = (R)0;            // targets come from inj_u (target_of)
        } else if (a.resample == SGM_RESAMPLE_SYSTEMATIC) {
            R u4[4];
            rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, u4);
#pragma unroll
            for (int c = 0; c < KPT; ++c) u[c] = u4[0];
        } else {
            rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, u);
            rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, u + 4);
        }
        draw_normals<R>(a, key, b, t, i0, z);
        __syncthreads();                                   // header
    }
    if (g == 0 && tid == 0) item_bookkeeping(a, b, t - 1, hdr, nws);
    const double total = hdr.total;

    // target of child c in global CDF units: tA + tB * u[c]  (f64; only used for the two range searches)
    double tA, tB;
    if (spacings) { tB = gw * (scale_over_total * total); tA = gl * (scale_over_total * total); }
    else { tB = total / (double)N; tA = (double)i0 * tB; }
    auto target_of = [&](int c) -> double {
        double tg;
        if (a.rng_mode == SGM_RNG_INJECTED) tg = (i0 + c < N) ? a.inj_u[((size_t)b * a.max_T + t) * N + i0 + c] * total : 0.0;
        else if (spacings) tg = tA + tB * (double)u[c];
        else tg = tA + tB * ((double)c + (double)u[c]);
        return (tg < total) ? tg : total * (1.0 - 1.2e-16);
    };

    // ---- parent range of this tile ---------------------------------------------------------------------
    {
        const int last = n_valid - 1, w_last = (last / KPT) >> 5, l_last = (last / KPT) & 31, c_last = last % KPT;
        if (warp == 0) {
            const int lo = warp_search_cdf<R>(__shfl_sync(FULL, target_of(0), 0), hdr, G, fine_old, N);
            if (lane == 0) s_range[0] = lo;
        }
        if (warp == w_last) {
            double tl = target_of(0);
#pragma unroll
            for (int c = 1; c < KPT; ++c) if (c == c_last) tl = target_of(c);
            const int hi = warp_search_cdf<R>(__shfl_sync(FULL, tl, l_last), hdr, G, fine_old, N);
            if (lane == 0) s_range[1] = hi;
        }
    }
    __syncthreads();
    const int lo = s_range[0], range = s_range[1] - lo + 1;
    int anc[KPT];
    if (range >= 1 && range <= CAP - 16) {
        // stage the CDF of [lo, hi] in global units relative to cbase (f32 is enough: the range spans a
        // few tiles at most), padded with +inf up to a power of two so the searches need no bound checks
        const int g_lo = lo / TILE;
        const double cbase = hdr.coarse[g_lo];
        for (int k = tid; k < range + 16; k += NT) {                 // 16 sentinels for the windowed probes
            R v = Mth<R>::inf();
            if (k < range) {
                const int p = lo + k, gp = p / TILE;
                v = (R)(hdr.coarse[gp] - cbase) + fine_old[p] * (R)hdr.e[gp];
            }
            s_cdf[k] = v;
        }
        __syncthreads();
        // relative targets in f32: rA + rB * u
        R rt[KPT];
        if (a.rng_mode == SGM_RNG_INJECTED) {
#pragma unroll
            for (int c = 0; c < KPT; ++c) rt[c] = (R)(target_of(c) - cbase);
        } else {
            const R rA = (R)(tA - cbase), rB = (R)tB;
#pragma unroll
            for (int c = 0; c < KPT; ++c) rt[c] = spacings ? (rA + rB * u[c]) : (rA + rB * ((R)c + u[c]));
        }
        // first child: branch-free binary search over the range; every further child advances from its
        // predecessor (expected: one parent) with a fixed 4-step search of the next 16 entries -- no
        // warp divergence -- and a loop only when the gap is longer than that
        const int lastp = range - 1;
        int pos = 0;
        for (int step = pow2_floor(range); step > 0; step >>= 1)
            if (pos + step <= range && s_cdf[pos + step - 1] <= rt[0]) pos += step;
        pos = min(pos, lastp);
        anc[0] = lo + pos;
#pragma unroll
        for (int c = 1; c < KPT; ++c) {
#pragma unroll
            for (int step = 8; step > 0; step >>= 1)
                if (s_cdf[pos + step - 1] <= rt[c]) pos += step;
            while (pos < lastp && s_cdf[pos] <= rt[c]) ++pos;
            pos = min(pos, lastp);
            anc[c] = lo + pos;
        }
    } else {
        // very uneven weights: the tile spans more parents than the staging buffer holds
#pragma unroll 1
        for (int c = 0; c < KPT; ++c) anc[c] = search_cdf<R>(target_of(c), hdr, G, fine_old, N);
    }
    R lwn[KPT];
    propagate_store<R, Model>(a, b, t, par, i0, item_off, anc, z, hdr, nws, carries, shrink, lwn);
    const bool need_ws = (nws > 0) && (a.pf == SGM_PF_FILTER || shrink || (carries && t == Tb - 1));
    tile_epilogue<R, W, NP>(lwn, i0, N, reinterpret_cast<R*>(a.fine[par ^ 1]) + item_off,
                            a.part[par ^ 1] + ((size_t)b * G + g) * PSTRIDE, a.rec[par ^ 1], a.tail[par ^ 1], item_off,
                            need_ws, nws, sh_r, sh_d);
}

// Per-(item, step) exclusive prefix of Gamma(P_g, 1) tile totals (+ one Exp(1) for the (N+1)-th
// spacing) used by the order-statistics multinomial resampler.  gam[b][t][0..G+1].
__global__ void __launch_bounds__(NT) gamma_prefix_kernel(KArgs a, double* gam_out) {
    __shared__ double sh_d[NWARP];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, G = a.G, N = a.N;
    RngKey key = a.key; key.item += (uint32_t)b;
    const int per = (G + 1 + NT - 1) / NT, g0 = tid * per;
    double loc = 0.0;
    for (int k = 0; k < per; ++k) {
        const int g2 = g0 + k;
        if (g2 <= G) loc += rng_gamma(key, (uint32_t)g2, (uint32_t)t, (g2 == G) ? 1.0 : (double)min(TILE, N - g2 * TILE));
    }
    double total;
    double run = block_excl_scan(loc, sh_d, total);
    double* out = gam_out + ((size_t)b * a.max_T + t) * (G + 2);
    for (int k = 0; k < per; ++k) {
        const int g2 = g0 + k;
        if (g2 <= G) {
            out[g2] = run;
            run += rng_gamma(key, (uint32_t)g2, (uint32_t)t, (g2 == G) ? 1.0 : (double)min(TILE, N - g2 * TILE));
        }
    }
    if (tid == 0) out[G + 1] = total;
}

// ---- final: last log-likelihood term + average_statistic (buffered_smoother.py:151-154) -----------
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_final_kernel(KArgs a) {
    __shared__ CdfHeader hdr;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int Tb = a.T_buf[b], par = Tb & 1, G = a.G;
    const int nws = stat_width<Model>(a.stat_kind);
    build_cdf_header(a.part[par] + (size_t)b * G * PSTRIDE, G, nws, hdr);
    __syncthreads();
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
