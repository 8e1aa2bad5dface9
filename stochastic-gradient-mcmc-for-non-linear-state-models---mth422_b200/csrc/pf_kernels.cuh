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

// Built by warp 0 with shuffles only (one block barrier at the end).  hdr.sbar[k] receives
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
    __syncthreads();
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
// Resampling search, two paths:
//   sorted targets (order-statistics multinomial / systematic / stratified, or INJECTED uniforms the
//   caller declares ascending): the CTA's 2048 children hit one contiguous parent range [lo, hi]; two
//   warp-cooperative searches find it, the CDF of that range is staged in shared memory in global
//   units, and each thread merges its 8 consecutive children against it (streaming gather).
//   iid targets (reference semantics): per-child coarse (shared) + fine (global) branch-free binary
//   searches, the 8 children of a thread interleaved for memory-level parallelism.
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_step_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ CdfHeader hdr;
    __shared__ R s_cdf[CAP];
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
    __shared__ int s_range[2];
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

    const R* __restrict__ fine_old = reinterpret_cast<const R*>(a.fine[par]) + item_off;
    const int i0 = g * TILE + tid * KPT;
    const int n_valid = min(TILE, N - g * TILE);             // children in this tile
    RngKey key = a.key; key.item += (uint32_t)b;
    const double total = hdr.total;

    // ---- resampling targets  u_i * total ------------------------------------------------------------
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
        // Order statistics of N iid uniforms via exponential spacings.  Within a tile the normalised
        // partial sums of P Exp(1) draws are independent of their total, which is Gamma(P, 1); the tile
        // totals are therefore drawn directly (gamma_prefix_kernel) and no cross-tile scan is needed.
        R e[KPT], run = (R)0;
        rng_uniform4(key, (uint32_t)(i0 >> 2), (uint32_t)t, STREAM_UNIFORM, e);
        rng_uniform4(key, (uint32_t)(i0 >> 2) + 1u, (uint32_t)t, STREAM_UNIFORM, e + 4);
#pragma unroll
        for (int c = 0; c < KPT; ++c) { e[c] = (i0 + c < N) ? -Mth<R>::log(e[c]) : (R)0; run += e[c]; }
        R tile_sum;
        R pre = block_excl_scan(run, sh_r, tile_sum);
        const double* gam = a.gam + ((size_t)b * a.max_T + t) * (G + 2);
        const double gl = gam[g], gmine = (gam[g + 1] - gam[g]) / (double)tile_sum, scale = total / gam[G + 1];
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
#pragma unroll
    for (int c = 0; c < KPT; ++c) if (!(target[c] < total)) target[c] = total * (1.0 - 1.2e-16);

    // ---- ancestor search ----------------------------------------------------------------------------
    int anc[KPT];
    bool staged = false;
    if (a.resample != SGM_RESAMPLE_MULTINOMIAL) {
        const int last = n_valid - 1, w_last = (last / KPT) >> 5, l_last = (last / KPT) & 31, c_last = last % KPT;
        const int warp = tid >> 5;
        if (warp == 0) {
            const int lo = warp_search_cdf<R>(__shfl_sync(FULL, target[0], 0), hdr, G, fine_old, N);
            if (tid == 0) s_range[0] = lo;
        }
        if (warp == w_last) {
            double tl = target[0];
#pragma unroll
            for (int c = 1; c < KPT; ++c) if (c == c_last) tl = target[c];
            const int hi = warp_search_cdf<R>(__shfl_sync(FULL, tl, l_last), hdr, G, fine_old, N);
            if ((tid & 31) == 0) s_range[1] = hi;
        }
        __syncthreads();
        const int lo = s_range[0], range = s_range[1] - lo + 1;
        if (range >= 1 && range <= CAP) {
            staged = true;
            const double cbase = hdr.coarse[lo / TILE];
            for (int k = tid; k < range; k += NT) {
                const int p = lo + k, gp = p / TILE;
                s_cdf[k] = (R)((hdr.coarse[gp] - cbase) + (double)fine_old[p] * hdr.e[gp]);
            }
            __syncthreads();
            // first child: branch-free binary search over the staged range; the rest merge forward
            int pos = 0;
            {
                const R rt = (R)(target[0] - cbase);
                for (int step = pow2_floor(range); step > 0; step >>= 1)
                    if (pos + step <= range && s_cdf[pos + step - 1] <= rt) pos += step;
                pos = min(pos, range - 1);
            }
            anc[0] = lo + pos;
#pragma unroll
            for (int c = 1; c < KPT; ++c) {
                const R rt = (R)(target[c] - cbase);
                if (i0 + c < N) while (pos < range - 1 && s_cdf[pos] <= rt) ++pos;
                anc[c] = lo + pos;
            }
        }
    }
    if (!staged) {
        int gsel[KPT], pos[KPT], len[KPT];
        R rr[KPT];
        const int step0 = pow2_floor(G);
#pragma unroll
        for (int c = 0; c < KPT; ++c) {
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

    const void* rec_old = a.rec[par];
    const void* tail_old = a.tail[par];
    void* rec_new = a.rec[par ^ 1];
    void* tail_new = a.tail[par ^ 1];
    R lwn[KPT];
#pragma unroll
    for (int h0 = 0; h0 < KPT; h0 += 4) {
        R ra[4][W];
#pragma unroll
        for (int c = 0; c < 4; ++c)                       // four independent parent gathers in flight
            if (i0 + h0 + c < N) load_rec<R, W>(rec_old, tail_old, item_off + anc[h0 + c], ra[c]);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const int c = h0 + c4, i = i0 + c;
            lwn[c] = (R)0;
            if (i < N) {
                R rn[W];
                Model::propagate(th, ra[c4] + NP, y, z[c], rn + NP);
                lwn[c] = Model::log_weight(th, ra[c4] + NP, rn + NP, y);
                R h[4] = {(R)0, (R)0, (R)0, (R)0};
                if (in_sub) {
                    if (a.stat_kind == SGM_STAT_SCORE) Model::score(th, ra[c4] + NP, rn + NP, y, h);
                    else if (a.stat_kind == SGM_STAT_SUFF) Model::suff(ra[c4] + NP, rn + NP, h);
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (carries) rn[q] = shrink ? (lam * ra[c4][q] + sbar[q] + h[q] * wt) : (ra[c4][q] + h[q] * wt);   // pf.py:175-179
                    else if (a.pf == SGM_PF_FILTER) rn[q] = h[q] * wt;                                               // pf.py:70-71
                    else rn[q] = (R)0;                                                                                // set by the backward kernel
                }
                store_rec<R, W>(rec_new, tail_new, item_off + i, rn);
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
    const bool need_ws = (nws > 0) && (a.pf == SGM_PF_FILTER || shrink || (carries && t == Tb - 1));
    tile_epilogue<R, W, NP>(lwn, i0, N, reinterpret_cast<R*>(a.fine[par ^ 1]) + item_off,
                            a.part[par ^ 1] + ((size_t)b * G + g) * PSTRIDE, rec_new, tail_new, item_off,
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
