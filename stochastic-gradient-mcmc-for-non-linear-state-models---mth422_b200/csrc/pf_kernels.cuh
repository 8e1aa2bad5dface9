// Kernels of the batched buffered particle filter / smoother (O(N) paths).
//
// Work decomposition: the unit of work is a WARP TILE of WT = 256 consecutive particles.  Lane l owns the
// particles 32 * c + l (c = 0..7) of its tile, so every warp-level load / store of particle data is one
// contiguous 128..512-byte access (the L1 data pipe, not HBM, was the first bound of a lane-contiguous layout).
// A warp does everything for its tile -- draw randoms, find its parents, gather, propose, reweight,
// update statistics, tile-local scan of the new weights -- with shuffles and __syncwarp only: the step
// kernel contains NO block barrier, so the 32+ resident warps of an SM hide each other's memory latency.
// All cross-tile work of a step (global max, tile offsets, log-likelihood, Gamma prefix) is done by a
// tiny per-item header kernel between two step kernels (a single warp for items with < 32 tiles); items with
// N <= 2048 and a small batch run the whole time loop in one launch out of shared memory (small_kernels.cuh), few items with
// N <= 65536 in one cooperative launch of these same device functions (coop_kernels.cuh).
//
// Data layout in HBM (caller workspace; B items, N particles, Q = ceil(N / 256) warp tiles per item):
//   rec [2][B][N][4]   R   first four components of the particle record  (stats..., then state)
//   tail[2][B][N][KT]  R   remaining KT = NX + NP - 4 components (SVM 0, LGSSM 1, GARCH 2)
//   fine[2][B][Q*256]  R   tile-local inclusive scan of exp(lw - m_tile)   ("fine" CDF), padded to whole tiles
//   lw  [2][B][N]      R   log-weights (only written when a consumer needs them)
//   sub [2][B][Q][8]   f64 per warp tile: m_tile, s_tile = sum exp(lw - m_tile), ws[0..3]
//   hdr [B][8+3(Q+2)]  f64 per item, rebuilt every step: M, total, sbar[4]; off[Q+1] exclusive prefix of
//                          tile masses in units of exp(-M); sc[Q] = exp(m_tile - M); gam[Q+2] Gamma prefix
//                          (order-statistics sampler: first target of every child tile, in units of exp(-M))
//   acc [B][16]        f64 running log-likelihood (+ filter / predictive statistic)
// rec / tail / fine / lw / sub are double-buffered on step parity: step t reads [t & 1], writes [(t+1) & 1].
//
// The global CDF of an item is never materialised: c_i = off[q] + fine_i * sc[q]  (flash-attention style
// rescaling of tile-local sums; offsets in f64, tile-local scans in R).
#pragma once
#include "blockops.cuh"
#include "models.cuh"
#include "rng.cuh"

namespace sgm {

#ifndef SGM_STEP_CTAS
#define SGM_STEP_CTAS 4
#endif
#ifndef SGM_GATHER_BATCH
#define SGM_GATHER_BATCH 4
#endif
constexpr int GB = SGM_GATHER_BATCH;   // parent gathers a lane keeps in flight
constexpr int KPT = 8;                 // consecutive particles per lane
constexpr int WT = 32 * KPT;           // 256 particles per warp tile
constexpr int TILE = NT * KPT;         // 2048 particles per CTA (8 warp tiles)
constexpr int MAX_Q = 4096;            // N <= 2^20
constexpr int WIN_BYTES = 4096;        // per-warp shared-memory window of the staged parent CDF (sorted resampling)
constexpr int SSTRIDE = 8;             // doubles per `sub` entry
constexpr int ACC_STRIDE = 16;         // doubles per `acc` entry: log-likelihood, then up to 15 filter / predictive statistics
static_assert(SGM_PRED_MAX_STEPS + 2 <= ACC_STRIDE && SGM_PRED_MAX_STEPS < SGM_PRED_SLOTS, "predictive horizons must fit acc / grad rows");
constexpr int THC_BYTES = 128;         // per-item slot for the model's derived constants
constexpr int H_M = 0, H_TOTAL = 1, H_SBAR = 2, H_SCALARS = 8;

__host__ __device__ inline size_t hdr_stride(int Q) { return (size_t)H_SCALARS + 3 * (size_t)(Q + 2); }

struct KArgs {
    int B, N, G, Q, max_T;
    int pf, rng_mode, resample, stat_kind, Ntilde, accept_reject, max_ar, manual_thresh;
    int need_lw, n2_tensor;
    int variates32;        // f64 runs: generate the random variates with the f32 transforms (see draw_normals)
    int pred_K, pred_per_horizon;  // SGM_STAT_PRED: num_steps_ahead, log-sum variant
    const double* inj_pred;
    int b0;                // first item of this launch (a batch may be split over two streams)
    const uint64_t* offset_dev;   // optional: Philox call offset read from device memory (CUDA-graph replays)
    double lambduh;
    RngKey key;            // .item holds item_id_base
    const double* obs; const int64_t* obs_off; const int32_t* T_buf; const int32_t* t1; const int32_t* tL;
    const double* step_weights; const int64_t* wts_off; const double* theta;
    const double* prior_mean; const double* prior_var;
    const double* inj_z0; const double* inj_u; const double* inj_z; const double* inj_extra; const int64_t* inj_extra_off;
    void* rec[2]; void* tail[2]; void* fine[2]; void* lw[2]; double* sub[2]; double* hdr; double* acc;
    int32_t* Jidx; int32_t* Llist[2]; int32_t* counters;
    double* pcdf; int32_t* pguide;   // PaRIS (PHILOX): flat f64 CDF [B][N] and guide table [B][N+1] of the old weights
    void* pkey;                      // PaRIS (PHILOX): [B][N][4] R per-parent score keys (u, gq, lw - M, -)
    void* n2part;          // [splits][B][N][8] R  split-J partial sums of the O(N^2) smoother
    char* thc;             // [B][THC_BYTES]   Model::Theta<R> (derived constants) written by the init kernel
    void* yw;              // [B][max_T][2] R  (y_t, statistic weight of step t: w_t inside [t1, tL), else 0)
    double* grad; double* loglik; int32_t* status;
    void* out_x; void* out_lw; void* out_stats; int32_t* trace_anc; void* trace_x; void* trace_lw; int32_t* trace_J;
};

// Philox key of item b.  The call offset normally travels in the kernel arguments; a captured CUDA graph replays
// the same arguments, so there it is read from device memory instead.
__device__ __forceinline__ RngKey item_key(const KArgs& a, int b) {
    RngKey key = a.key;
    if (a.offset_dev) {
        const uint64_t o = *a.offset_dev;
        key.offset = (uint32_t)(o & 0xffffffffu);
        key.k1 ^= (uint32_t)(o >> 32);
    }
    key.item += (uint32_t)b;
    return key;
}

// ---- record access ---------------------------------------------------------------------------
template <class R> struct alignas(4 * sizeof(R)) Vec4T { R x, y, z, w; };

template <class R, int W, class I> __device__ __forceinline__ void load_rec(const void* rec, const void* tail, I idx, R* r) {
    const Vec4T<R> v = reinterpret_cast<const Vec4T<R>*>(rec)[idx];
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    if (W == 5) r[4] = reinterpret_cast<const R*>(tail)[idx];
    if (W == 6) { const R* t = reinterpret_cast<const R*>(tail) + 2 * idx; r[4] = t[0]; r[5] = t[1]; }
}
template <class R, int W, class I> __device__ __forceinline__ void store_rec(void* rec, void* tail, I idx, const R* r) {
    Vec4T<R> v; v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3];
    reinterpret_cast<Vec4T<R>*>(rec)[idx] = v;
    if (W == 5) reinterpret_cast<R*>(tail)[idx] = r[4];
    if (W == 6) { R* t = reinterpret_cast<R*>(tail) + 2 * idx; t[0] = r[4]; t[1] = r[5]; }
}

template <class Model> __device__ __forceinline__ int stat_width(int stat_kind) {
    return stat_kind == SGM_STAT_SCORE ? Model::NP : (stat_kind == SGM_STAT_SUFF ? 3 : 0);
}
template <class R, class Model>
__device__ __forceinline__ typename Model::template Theta<R> load_thc(const KArgs& a, int b) {
    return *reinterpret_cast<const typename Model::template Theta<R>*>(a.thc + (size_t)b * THC_BYTES);
}
__device__ __forceinline__ bool uses_spacings(const KArgs& a) {
    return a.rng_mode == SGM_RNG_PHILOX && a.resample == SGM_RESAMPLE_MULTINOMIAL_SORTED;
}
__device__ __forceinline__ int pow2_floor(int x) { return 1 << (31 - __clz(max(x, 1))); }

// ---- per-item CDF header (global memory, written by pf_header_kernel) ---------------------------------
struct ItemHdr {
    const double* base; const double* off; const double* sc; const double* gam;
    int Q;
    double M, total;
};
// `local` != nullptr: a copy of the header somewhere else (the cooperative kernel keeps one per CTA -- for fewer than 32 tiles one per warp -- in shared memory)
__device__ __forceinline__ ItemHdr load_hdr(const KArgs& a, int b, const double* local = nullptr) {
    ItemHdr h;
    h.Q = a.Q;
    h.base = local ? local : a.hdr + (size_t)b * hdr_stride(a.Q);
    h.off = h.base + H_SCALARS;
    h.sc = h.off + (a.Q + 2);
    h.gam = h.sc + (a.Q + 2);
    h.M = h.base[H_M];
    h.total = h.base[H_TOTAL];
    return h;
}

// searchsorted(cdf, u, side='right') on the hierarchical CDF: first index whose cumulative mass exceeds
// target = u * total.  One thread, dependent (cached) loads.
template <class R>
__device__ __forceinline__ int search_hdr(double target, const ItemHdr& h, const R* fine, int N) {
    if (!(target < h.total)) target = h.total * (1.0 - 1.2e-16);
    int q = 0;
    for (int step = pow2_floor(h.Q); step > 0; step >>= 1)       // largest q with off[q] <= target
        if (q + step < h.Q && h.off[q + step] <= target) q += step;
    const R r = (R)((target - h.off[q]) / h.sc[q]);
    const int base = q * WT;
    const int len = min(WT, N - base);
    const R* f = fine + base;
    int pos = 0;
#pragma unroll
    for (int step = WT / 2; step > 0; step >>= 1)
        if (pos + step <= len && f[pos + step - 1] <= r) pos += step;
    return base + min(pos, len - 1);
}

// Coarse searches of two targets at once (first and last child of a warp tile): largest tile q with
// off[q] <= target, by 32-ary probing with ballots; all loads of a level are issued together so the two
// dependent-load chains overlap.  `c1` is the level-1 probe off[lane * s1c] (+inf beyond Q), prefetched by
// the caller before it draws its randoms.
__device__ __forceinline__ void warp_search_tiles(double ta, double tb, double c1, int s1c, const ItemHdr& h,
                                                  int lane, int& qa, int& qb) {
    qa = (max(__popc(__ballot_sync(FULL, c1 <= ta)), 1) - 1) * s1c;
    qb = (max(__popc(__ballot_sync(FULL, c1 <= tb)), 1) - 1) * s1c;
    int la = min(s1c, h.Q - qa), lb = min(s1c, h.Q - qb);
    while (la > 1 || lb > 1) {
        const int sa = (la + 31) >> 5, sb = (lb + 31) >> 5;
        const double va = (la > 1 && lane * sa < la) ? h.off[qa + lane * sa] : Mth<double>::inf();
        const double vb = (lb > 1 && lane * sb < lb) ? h.off[qb + lane * sb] : Mth<double>::inf();
        if (la > 1) {
            const int adv = (max(__popc(__ballot_sync(FULL, va <= ta)), 1) - 1) * sa;
            la = min(sa, la - adv); qa += adv;
        }
        if (lb > 1) {
            const int adv = (max(__popc(__ballot_sync(FULL, vb <= tb)), 1) - 1) * sb;
            lb = min(sb, lb - adv); qb += adv;
        }
    }
}

// ---- warp-tile scan helpers -------------------------------------------------------------------------------
// A warp tile is an 8 x 32 matrix of particles, slot = 32 c + lane ("row-major": what the coalesced loads /
// stores want).  A prefix scan in slot order is cheapest "lane-major" (lane l holds slots 8 l .. 8 l + 7: 7
// serial adds, ONE shuffle scan of the lane totals, 8 offset adds, instead of 8 shuffle scans), so scans
// transpose through the warp's slice `s_tr` (>= 256 R, 32-byte aligned) of shared memory.
template <class R>
__device__ __forceinline__ void lane_major_incl_scan(R* v, R& total) {           // v[k]: slot 8 lane + k, in place
#pragma unroll
    for (int k = 1; k < KPT; ++k) v[k] += v[k - 1];
    const R incl = warp_incl_scan(v[KPT - 1]);
    const R excl = incl - v[KPT - 1];
#pragma unroll
    for (int k = 0; k < KPT; ++k) v[k] += excl;
    total = __shfl_sync(FULL, incl, 31);
}
template <class R>
__device__ __forceinline__ void store_lane_major(R* dst, int lane, const R* v) {  // dst[8 lane + k] = v[k]
    Vec4T<R> a, b;
    a.x = v[0]; a.y = v[1]; a.z = v[2]; a.w = v[3]; b.x = v[4]; b.y = v[5]; b.z = v[6]; b.w = v[7];
    reinterpret_cast<Vec4T<R>*>(dst)[2 * lane] = a;
    reinterpret_cast<Vec4T<R>*>(dst)[2 * lane + 1] = b;
}
template <class R>
__device__ __forceinline__ void load_lane_major(const R* src, int lane, R* v) {
    const Vec4T<R> a = reinterpret_cast<const Vec4T<R>*>(src)[2 * lane], b = reinterpret_cast<const Vec4T<R>*>(src)[2 * lane + 1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- per-warp-tile epilogue: tile max, tile-local scan of exp(lw - m), per-tile partials ---------------
// lwn[c] belongs to particle tile_base + 32 c + lane; callers set -inf beyond N (zero weight).
// `carried` != nullptr: the caller already holds the lane's weighted statistic sums sum_c stat[c] * exp(lwn[c] - carried->m)
// (accumulated in registers while the particles were propagated, see WsCarry) -- no re-read of the new records.
template <class R> struct WsCarry { R m; R s[4]; };
template <class R, int W>
__device__ __forceinline__ void warp_tile_epilogue(const R* lwn, int tile_base, int N, int lane, R* fine_out, double* sub_out,
                                                   const void* rec_new, const void* tail_new, size_t item_off,
                                                   bool need_ws, int nws, R* s_tr, const WsCarry<R>* carried = nullptr) {
    R m = -Mth<R>::inf();
#pragma unroll
    for (int c = 0; c < KPT; ++c) m = nan_max(m, lwn[c]);
    m = warp_max(m);
    // m == -inf (every weight of the tile zero): shift by 0 instead; exp(-inf) = 0, and a NaN log-weight
    // still poisons the tile sum so that the item gets flagged
    const R msafe = (m == -Mth<R>::inf()) ? (R)0 : m;
    R w[KPT], f[KPT], total;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        w[c] = Mth<R>::exp(lwn[c] - msafe);
        s_tr[32 * c + lane] = w[c];
    }
    __syncwarp();
    load_lane_major<R>(s_tr, lane, f);
    lane_major_incl_scan<R>(f, total);
    store_lane_major<R>(fine_out + tile_base, lane, f);   // `fine` is padded to whole tiles (entries >= N repeat the total)
    __syncwarp();
    double ws[4] = {0.0, 0.0, 0.0, 0.0};
    if (carried) {
        // lane-local reference -> tile reference (exp(-inf - msafe) = 0 for a lane that owns no live particle)
        const R sc = Mth<R>::exp(carried->m - msafe);
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < nws) ws[q] = (double)warp_sum(carried->s[q] * sc);
    } else if (need_ws) {
        // weighted statistic sums of the tile (Nemeth shrinkage, filter statistic, final average): summed in R within
        // the tile (256 terms with weights <= 1), in f64 across tiles (header kernel)
        R wl[4] = {(R)0, (R)0, (R)0, (R)0};
#pragma unroll
        for (int c = 0; c < KPT; ++c) {
            const int i = tile_base + 32 * c + lane;
            if (i < N) {
                R r[W];
                load_rec<R, W>(rec_new, tail_new, item_off + i, r);
#pragma unroll
                for (int q = 0; q < 4; ++q) if (q < nws) wl[q] += r[q] * w[c];
            }
        }
        for (int q = 0; q < nws; ++q) ws[q] = (double)warp_sum(wl[q]);
    }
    if (lane == 0) {
        sub_out[0] = (double)m;
        sub_out[1] = (double)total;
        for (int q = 0; q < 4; ++q) sub_out[2 + q] = ws[q];
    }
}

// ---- init: x0 ~ N(prior_mean, prior_var), lw = 0, stats = 0  (buffered_smoother.py:67-75) ----------
template <class R, class Model>
__device__ __forceinline__ void init_body(const KArgs& a, int b, int g, R* s_tr) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N;
    const size_t item_off = (size_t)b * N;
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
            if (a.counters) for (int q = 0; q < 16; ++q) a.counters[b * 16 + q] = 0;
        }
    }
    const int q_me = g * NWARP + warp;
    if (q_me >= a.Q) return;
    const int tile_base = q_me * WT;
    const RngKey key = item_key(a, b);
    const R mean = (R)a.prior_mean[b], sd = (R)::sqrt(a.prior_var[b]);
    R lwn[KPT];
    R z[KPT];
    if (a.rng_mode == SGM_RNG_PHILOX) {
        rng_normal4(key, (uint32_t)(q_me * 64 + lane * 2), 0xffffu, z);
        rng_normal4(key, (uint32_t)(q_me * 64 + lane * 2 + 1), 0xffffu, z + 4);
    } else {
#pragma unroll
        for (int c = 0; c < KPT; ++c) { const int i = tile_base + 32 * c + lane; z[c] = (i < N) ? (R)a.inj_z0[item_off + i] : (R)0; }
    }
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
        const int i = tile_base + 32 * c + lane;
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
    warp_tile_epilogue<R, W>(lwn, tile_base, N, lane, reinterpret_cast<R*>(a.fine[0]) + (size_t)b * a.Q * WT,
                             a.sub[0] + ((size_t)b * a.Q + q_me) * SSTRIDE, a.rec[0], a.tail[0], item_off, false, 0, s_tr);
}
template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_init_kernel(KArgs a) {
    __shared__ __align__(32) R s_tr_all[NWARP][WT];
    init_body<R, Model>(a, a.b0 + blockIdx.y, blockIdx.x, s_tr_all[threadIdx.x >> 5]);
}

// ---- per-item header: everything of a step that crosses tiles ----------------------------------------
// One CTA per item, launched before step kernel t (final_pass = 0) and once after the last step (1):
//   * global max M, tile scales sc[q] = exp(m_q - M), exclusive prefix off[q] of the tile masses (f64)
//   * weighted statistic mean sbar (Nemeth shrinkage, filter statistic, final average_statistic)
//   * log-likelihood increment of the step that produced the weights (buffered_smoother.py:124-126; here
//     with the max shift: M + log(total / N)), degenerate-weight flags
//   * exclusive prefix of the per-tile Gamma(P_q, 1) draws (+ one Exp(1)) of the order-statistics sampler
//   * final: grad = average_statistic (buffered_smoother.py:151-154) or the filter statistic (pf.py:77-80)
// scalar tail of the header (one thread): totals, status flags, log-likelihood increment, final outputs
template <class Model>
__device__ __forceinline__ void header_finish(const KArgs& a, int b, int t_done, int final_pass, int nstat, double M, double total,
                                              const double* sbar, double* base, double* off, bool side_effects = true,
                                              double* acc_local = nullptr) {
    const int Q = a.Q, N = a.N;
    const double NEG_INF = -Mth<double>::inf();
    off[Q] = total;
    base[H_M] = M; base[H_TOTAL] = total;
    for (int j = 0; j < 4; ++j) base[H_SBAR + j] = sbar[j];
    if (!side_effects) return;             // a redundant per-warp copy: status / log-likelihood are the first warp's job
    if (!(total > 0.0) || !(total < Mth<double>::inf()) || !(M == M) || !(fabs(M) < Mth<double>::inf()))
        a.status[b] |= (total == 0.0 || M == NEG_INF) ? SGM_STATUS_ZERO_WEIGHT : SGM_STATUS_NAN_WEIGHT;
    double* acc = acc_local ? acc_local : a.acc + (size_t)b * ACC_STRIDE;       // acc_local: the cooperative kernel's shared-memory copy
    if (t_done >= 0) {
        if (t_done >= a.t1[b] && t_done < a.tL[b]) {
            const double wt = (a.wts_off && a.wts_off[b] >= 0) ? a.step_weights[a.wts_off[b] + (t_done - a.t1[b])] : 1.0;
            acc[0] += wt * (M + ::log(total / (double)N));
        }
        if (a.pf == SGM_PF_FILTER) for (int j = 0; j < nstat; ++j) acc[1 + j] += sbar[j];     // pf.py:77-80
    }
    if (final_pass) {
        a.loglik[b] = acc[0];
        if (a.stat_kind == SGM_STAT_PRED) {                 // K + 1 horizons in a row of SGM_PRED_SLOTS
            for (int j = 0; j < SGM_PRED_SLOTS; ++j) a.grad[(size_t)b * SGM_PRED_SLOTS + j] = (j <= a.pred_K) ? acc[1 + j] : 0.0;
        } else {
            for (int j = 0; j < 8; ++j) a.grad[(size_t)b * 8 + j] = 0.0;
            for (int j = 0; j < nstat; ++j) a.grad[(size_t)b * 8 + j] = (a.pf == SGM_PF_FILTER) ? acc[1 + j] : sbar[j];
            if (a.pf == SGM_PF_PARIS && a.counters && a.rng_mode == SGM_RNG_PHILOX) {
                // diagnostics of the device-random PaRIS sampler: accept-reject proposals made and entries that fell back to
                // the exact sampler, summed over the item's time steps (slots 6, 7 of the row: never part of a statistic)
                a.grad[(size_t)b * 8 + 6] = (double)a.counters[b * 16 + 1];
                a.grad[(size_t)b * 8 + 7] = (double)a.counters[b * 16 + 2];
            }
        }
    }
}

// Header of an item with fewer than 32 warp tiles (N < 8192): ONE warp, shuffles only -- no block barrier.
// This is the header path of items with 2048 < N < 8192 (per-step launches and the cooperative kernel).
// `local` != nullptr: EVERY warp of the CTA calls this and builds its own copy of the header there (single-launch
// kernel: no barrier and no global round trip between header and step; the copies are bit-identical -- same inputs,
// counter-based Gamma draws, fixed shuffle trees); only the first warp applies the side effects.
template <class R, class Model>
__device__ __forceinline__ void header_warp(const KArgs& a, int b, int t, int final_pass, double* local = nullptr, bool side = true) {
    if (!local && threadIdx.x >= 32) return;
    const int lane = threadIdx.x & 31;
    const int Tb = a.T_buf[b];
    const int par = final_pass ? (Tb & 1) : (t & 1);
    const int t_done = final_pass ? Tb - 1 : t - 1;
    const int Q = a.Q, N = a.N;
    const int nstat = stat_width<Model>(a.stat_kind);
    const bool shrink = (a.pf == SGM_PF_NEMETH) && (a.lambduh != 1.0);
    const int nws = (final_pass || shrink || a.pf == SGM_PF_FILTER) ? nstat : 0;
    const double* sub = a.sub[par] + (size_t)b * Q * SSTRIDE;
    double* base = local ? local : a.hdr + (size_t)b * hdr_stride(Q);
    double* off = base + H_SCALARS;
    double* sc = off + (Q + 2);
    double* gam = sc + (Q + 2);
    const double NEG_INF = -Mth<double>::inf();
    const bool has = lane < Q;
    const double* p = sub + (size_t)lane * SSTRIDE;
    const double mq = has ? p[0] : NEG_INF;
    const double M = warp_max(mq);
    const double e = (has && mq != NEG_INF) ? ::exp(mq - M) : 0.0;
    const double loc = has ? e * p[1] : 0.0;
    const double incl = warp_incl_scan(loc);
    const double total = __shfl_sync(FULL, incl, 31);
    if (has) { off[lane] = incl - loc; sc[lane] = e; }
    double sbar[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < nws; ++j) sbar[j] = warp_sum(has ? e * p[2 + j] : 0.0) / total;
    if (!final_pass && uses_spacings(a)) {
        const RngKey key = item_key(a, b);
        const double g = (lane <= Q) ? rng_gamma(key, (uint32_t)lane, (uint32_t)t, (lane == Q) ? 1.0 : (double)min(WT, N - lane * WT)) : 0.0;
        const double gincl = warp_incl_scan(g);
        const double kk = total / __shfl_sync(FULL, gincl, 31);
        if (lane <= Q) gam[lane] = (gincl - g) * kk;
        if (lane == 0) gam[Q + 1] = total;
    }
    if (lane == 0) header_finish<Model>(a, b, t_done, final_pass, nstat, M, total, sbar, base, off, side && (!local || threadIdx.x < 32));
    if (local) __syncwarp();
}

// What the cooperative kernel hands to header_body: values it loaded once (T_buf) or prepared while it waited at the grid
// barrier (the Gamma draws of the step and their block scan, none of which depends on the weights), and its shared-memory
// log-likelihood accumulator.  Same values and the same order of operations as header_body computes on its own.
struct HeaderCoop {
    int Tb;
    bool scanned;              // gv / grun / gtot hold this step's draws and their exclusive block scan
    double gv[2], grun, gtot;
    double* acc;               // ACC_STRIDE doubles in shared memory (CTA 0 of the item applies the side effects)
};
// The calling thread's Gamma draws of step t when a header thread owns at most two of the Q + 1 draws (Q <= 2 NTH - 1):
// they do not depend on the weights, so the cooperative kernel draws them while it waits at the grid barrier.
template <int NTH>
__device__ __forceinline__ void header_gammas(const KArgs& a, int b, int t, double* gv) {
    const int Q = a.Q, N = a.N;
    const int perg = (Q + 1 + NTH - 1) / NTH, g0 = threadIdx.x * perg;
    const RngKey key = item_key(a, b);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int q = g0 + k;
        gv[k] = (k < perg && q <= Q) ? rng_gamma(key, (uint32_t)q, (uint32_t)t, (q == Q) ? 1.0 : (double)min(WT, N - q * WT)) : 0.0;
    }
}

// draws + block scan of step t (all threads of the CTA; sh_d as in header_body)
template <int NTH>
__device__ __forceinline__ void header_gamma_scan(const KArgs& a, int b, int t, double* sh_d, HeaderCoop& hc) {
    header_gammas<NTH>(a, b, t, hc.gv);
    const int perg = (a.Q + 1 + NTH - 1) / NTH, g0 = threadIdx.x * perg;
    double gl = 0.0;
    gl += hc.gv[0];
    if (perg == 2 && g0 + 1 <= a.Q) gl += hc.gv[1];
    hc.grun = block_excl_scan<NTH / 32>(gl, sh_d, hc.gtot);
    hc.scanned = true;
}

// NTH = threads of the CTA: 256, or 1024 for items with more than 256 tiles (N > 65536), where the per-thread chunk of
// tiles (and with it the serial part of the scans) shrinks 4x.
template <class R, class Model, int NTH = NT>
__device__ __forceinline__ void header_body(const KArgs& a, int b, int t, int final_pass, double* sh_d, double* local = nullptr,
                                            bool side = true, const HeaderCoop* hc = nullptr) {
    // `local` != nullptr: the header goes to that (shared-memory) copy instead of the item's global one -- the cooperative
    // kernel lets every CTA of an item build its own; `side` = this caller applies the log-likelihood / status / output
    // side effects (exactly one CTA per item may)
    const int tid = threadIdx.x;
    const int Tb = hc ? hc->Tb : a.T_buf[b];
    if (!final_pass && t >= Tb) return;
    if (a.Q < 32) { if (!local || tid < 32) header_warp<R, Model>(a, b, t, final_pass, local, side); return; }
    const int par = final_pass ? (Tb & 1) : (t & 1);
    const int t_done = final_pass ? Tb - 1 : t - 1;
    const int Q = a.Q, N = a.N;
    const int nstat = stat_width<Model>(a.stat_kind);
    const bool shrink = (a.pf == SGM_PF_NEMETH) && (a.lambduh != 1.0);
    const int nws = (final_pass || shrink || a.pf == SGM_PF_FILTER) ? nstat : 0;
    const double* sub = a.sub[par] + (size_t)b * Q * SSTRIDE;
    double* base = local ? local : a.hdr + (size_t)b * hdr_stride(Q);
    double* off = base + H_SCALARS;
    double* sc = off + (Q + 2);
    double* gam = sc + (Q + 2);
    const double NEG_INF = -Mth<double>::inf();

    const int per = (Q + NTH - 1) / NTH, q0 = tid * per;
    double m = NEG_INF;
    for (int k = 0; k < per; ++k) if (q0 + k < Q) m = fmax(m, sub[(size_t)(q0 + k) * SSTRIDE]);
    const double M = block_max<NTH / 32>(m, sh_d);
    double loc = 0.0, ws[4] = {0.0, 0.0, 0.0, 0.0};
    double e_first = 0.0;                         // scale of the thread's first tile: kept for the loop below when it is the only one
    for (int k = 0; k < per; ++k) {
        const int q = q0 + k;
        if (q < Q) {
            const double* p = sub + (size_t)q * SSTRIDE;
            const double e = (p[0] == NEG_INF) ? 0.0 : ::exp(p[0] - M);
            if (k == 0) e_first = e;
            loc += e * p[1];
            for (int j = 0; j < nws; ++j) ws[j] += e * p[2 + j];
        }
    }
    double total;
    double run = block_excl_scan<NTH / 32>(loc, sh_d, total);
    if (per == 1) {
        if (q0 < Q) { off[q0] = run; sc[q0] = e_first; }
    } else {
        for (int k = 0; k < per; ++k) {
            const int q = q0 + k;
            if (q < Q) {
                const double* p = sub + (size_t)q * SSTRIDE;
                const double e = (p[0] == NEG_INF) ? 0.0 : ::exp(p[0] - M);
                off[q] = run;
                sc[q] = e;
                run += e * p[1];
            }
        }
    }
    double sbar[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < nws; ++j) sbar[j] = block_sum<NTH / 32>(ws[j], sh_d) / total;

    if (!final_pass && uses_spacings(a)) {
        const RngKey key = item_key(a, b);
        const int perg = (Q + 1 + NTH - 1) / NTH, g0 = tid * perg;
        // a thread with at most two draws keeps them in registers (gpre: already drawn by the caller) instead of drawing
        // them again for the running sums below: same values, same order of additions
        const bool cached = perg <= 2;
        const bool scanned = cached && hc && hc->scanned;
        double gv[2] = {0.0, 0.0};
        if (scanned) { gv[0] = hc->gv[0]; gv[1] = hc->gv[1]; }
        else if (cached) header_gammas<NTH>(a, b, t, gv);
        double gtot, grun;
        if (scanned) { grun = hc->grun; gtot = hc->gtot; }
        else {
            double gl = 0.0;
            if (cached) { gl += gv[0]; if (perg == 2 && g0 + 1 <= Q) gl += gv[1]; }
            else {
                for (int k = 0; k < perg; ++k) {
                    const int q = g0 + k;
                    if (q <= Q) gl += rng_gamma(key, (uint32_t)q, (uint32_t)t, (q == Q) ? 1.0 : (double)min(WT, N - q * WT));
                }
            }
            grun = block_excl_scan<NTH / 32>(gl, sh_d, gtot);
        }
        const double kk = total / gtot;                  // stored in target units: gam[q] = total * G_q / G_total
        for (int k = 0; k < perg; ++k) {
            const int q = g0 + k;
            if (q <= Q) {
                gam[q] = grun * kk;
                grun += cached ? gv[k & 1] : rng_gamma(key, (uint32_t)q, (uint32_t)t, (q == Q) ? 1.0 : (double)min(WT, N - q * WT));
            }
        }
        if (tid == 0) gam[Q + 1] = total;
    }
    if (tid == 0) header_finish<Model>(a, b, t_done, final_pass, nstat, M, total, sbar, base, off, side, hc ? hc->acc : nullptr);
}

template <class R, class Model, int NTH = NT>
__global__ void __launch_bounds__(NTH) pf_header_kernel(KArgs a, int t, int final_pass) {
    __shared__ double sh_d[NTH / 32];
    header_body<R, Model, NTH>(a, a.b0 + blockIdx.x, t, final_pass, sh_d);
}

// ---- gather parents -> propagate -> reweight -> statistic update -> store (pf.py:30-36, 168-179) -----
// Row c of the warp tile = particles tile_base + 32 c + lane: coalesced stores, and with ascending
// ancestors the parent gathers of a row are (nearly) contiguous too.
// FM = fast mode of the instantiation (see step_body): FM_GENERIC reads every flag at run time.
enum : int { FM_GENERIC = 0, FM_POY = 1, FM_SHRINK = 2, FM_FILTER = 3 };
template <class R, class Model, int FM = FM_GENERIC, bool RAGGED = false>
__device__ __forceinline__ void propagate_store(const KArgs& a, int b, int t, int par, int tile_base, int lane, size_t item_off,
                                                const int* anc, const R* z, const ItemHdr& hdr, int nws,
                                                bool carries, bool shrink, R* lwn,
                                                const typename Model::template Theta<R>& th, R y, R wt, WsCarry<R>* carry) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    constexpr bool FAST = FM != FM_GENERIC;
    constexpr bool FULLT = FAST && !RAGGED;        // every tile of the item is complete: no bound checks
    constexpr bool CARRY_WS = FM == FM_SHRINK || FM == FM_FILTER;   // weighted statistic sums accumulated in registers
    const int N = a.N;
    const bool in_sub = wt != (R)0;               // yw[1] is zero outside [t1, tL)
    const R lam = (R)a.lambduh;
    R sbar[4] = {(R)0, (R)0, (R)0, (R)0};
    if (shrink) for (int q = 0; q < nws; ++q) sbar[q] = (R)((1.0 - a.lambduh) * hdr.base[H_SBAR + q]);
    const bool tracing = !FAST && (a.need_lw || a.trace_anc || a.trace_x || a.trace_lw);
    // new statistic = keep * parent statistic + sbar + h * hs   (one FMA chain for every smoother):
    //   Poyiadjis O(N)/Nemeth: keep = lambduh, hs = w_t (pf.py:175-179); filter: keep = 0, hs = w_t
    //   (pf.py:70-71); O(N^2)/PaRIS: keep = hs = 0 (the backward kernel writes the statistic)
    const R keep = carries ? (shrink ? lam : (R)1) : (R)0;
    const R hs = (FAST || carries || a.pf == SGM_PF_FILTER) ? wt : (R)0;
    const int stat_kind = (in_sub && hs != (R)0) ? (FAST ? (int)SGM_STAT_SCORE : a.stat_kind) : (int)SGM_STAT_NONE;
    const bool plain = FM == FM_POY || (!FAST && carries && !shrink);
    // per-item base pointers: the particle index stays a 32-bit register (one IMAD.WIDE per address)
    const void* rec_old = reinterpret_cast<const Vec4T<R>*>(a.rec[par]) + item_off;
    const void* tail_old = reinterpret_cast<const R*>(a.tail[par]) + item_off * (W - 4);
    void* rec_new = reinterpret_cast<Vec4T<R>*>(a.rec[par ^ 1]) + item_off;
    void* tail_new = reinterpret_cast<R*>(a.tail[par ^ 1]) + item_off * (W - 4);
    R* lw_new = reinterpret_cast<R*>(a.lw[par ^ 1]) + item_off;
    // keep the compiler from folding the item offset back into every 64-bit address computation
    asm volatile("" : "+l"(rec_old), "+l"(tail_old), "+l"(rec_new), "+l"(tail_new), "+l"(lw_new));
    if (CARRY_WS) { carry->m = -Mth<R>::inf(); carry->s[0] = carry->s[1] = carry->s[2] = carry->s[3] = (R)0; }
#pragma unroll
    for (int h0 = 0; h0 < KPT; h0 += GB) {
        R ra[GB][W];
#pragma unroll
        for (int c = 0; c < GB; ++c)                      // GB independent parent gathers in flight
            if (FULLT || tile_base + 32 * (h0 + c) + lane < N) load_rec<R, W>(rec_old, tail_old, anc[h0 + c], ra[c]);
#pragma unroll
        for (int c4 = 0; c4 < GB; ++c4) {
            const int c = h0 + c4, i = tile_base + 32 * c + lane;
            lwn[c] = -Mth<R>::inf();
            if (FULLT || i < N) {
                R rn[W];
                Model::propagate(th, ra[c4] + NP, y, z[c], rn + NP);
                lwn[c] = Model::log_weight(th, ra[c4] + NP, rn + NP, y);
                R h[4] = {(R)0, (R)0, (R)0, (R)0};
                if (stat_kind == SGM_STAT_SCORE) Model::score(th, ra[c4] + NP, rn + NP, y, h);
                else if (stat_kind == SGM_STAT_SUFF) Model::suff(ra[c4] + NP, rn + NP, h);
                if (plain) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) rn[q] = ra[c4][q] + h[q] * hs;           // Poyiadjis O(N): one FMA
                } else if (FM == FM_FILTER) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) rn[q] = h[q] * hs;                       // pf.py:70-71: nothing carried
                } else {
#pragma unroll
                    for (int q = 0; q < NP; ++q) rn[q] = keep * ra[c4][q] + (sbar[q] + h[q] * hs);
                }
                store_rec<R, W>(rec_new, tail_new, i, rn);
                if (CARRY_WS) {
                    // sum_c stat[c] * exp(lw[c] - m) with a lane-local running reference m (rescaled to the tile maximum
                    // in the epilogue): the freshly written records are not read again
                    const R mnew = nan_max(carry->m, lwn[c]);
                    const R msf = (mnew == -Mth<R>::inf()) ? (R)0 : mnew;
                    const R so = Mth<R>::exp(carry->m - msf), e = Mth<R>::exp(lwn[c] - msf);
#pragma unroll
                    for (int q = 0; q < NP; ++q) carry->s[q] = carry->s[q] * so + rn[q] * e;
                    carry->m = mnew;
                }
                if (tracing) {
                    if (a.need_lw) lw_new[i] = lwn[c];
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

// Random numbers of a warp tile: one Philox call yields the values of rows 4h..4h+3 of a lane
// (counter index = 64 q + 2 lane + h: a function of the particle index only).
// a.variates32 (f64 runs only): the variates are generated with the f32 transforms (one Philox call per four values,
// MUFU log / sin / cos) and widened -- exactly the random numbers an f32 run of the same call draws; all arithmetic on
// the particle system stays f64.
__device__ __forceinline__ void widen8(const float* s, double* d) {
#pragma unroll
    for (int c = 0; c < KPT; ++c) d[c] = (double)s[c];
}
__device__ __forceinline__ void widen8(const float* s, float* d) {
#pragma unroll
    for (int c = 0; c < KPT; ++c) d[c] = s[c];
}
template <class R>
__device__ __forceinline__ void draw_normals(const KArgs& a, const RngKey& key, int b, int t, int q_me, int lane, R* z) {
    if (sizeof(R) == 8 && a.variates32 && a.rng_mode != SGM_RNG_INJECTED) {
        float zf[KPT];
        rng_normal4(key, (uint32_t)(q_me * 64 + lane * 2), (uint32_t)t, zf);
        rng_normal4(key, (uint32_t)(q_me * 64 + lane * 2 + 1), (uint32_t)t, zf + 4);
        widen8(zf, z);
        return;
    }
    if (a.rng_mode == SGM_RNG_INJECTED) {
        const double* zz = a.inj_z + ((size_t)b * a.max_T + t) * a.N;
#pragma unroll
        for (int c = 0; c < KPT; ++c) { const int i = q_me * WT + 32 * c + lane; z[c] = (i < a.N) ? (R)zz[i] : (R)0; }
    } else {
        rng_normal4(key, (uint32_t)(q_me * 64 + lane * 2), (uint32_t)t, z);
        rng_normal4(key, (uint32_t)(q_me * 64 + lane * 2 + 1), (uint32_t)t, z + 4);
    }
}
template <class R>
__device__ __forceinline__ void draw_uniforms(const RngKey& key, int t, int q_me, int lane, R* u, bool variates32 = false) {
    if (sizeof(R) == 8 && variates32) {
        float uf[KPT];
        rng_uniform4(key, (uint32_t)(q_me * 64 + lane * 2), (uint32_t)t, STREAM_UNIFORM, uf);
        rng_uniform4(key, (uint32_t)(q_me * 64 + lane * 2 + 1), (uint32_t)t, STREAM_UNIFORM, uf + 4);
        widen8(uf, u);
        return;
    }
    rng_uniform4(key, (uint32_t)(q_me * 64 + lane * 2), (uint32_t)t, STREAM_UNIFORM, u);
    rng_uniform4(key, (uint32_t)(q_me * 64 + lane * 2 + 1), (uint32_t)t, STREAM_UNIFORM, u + 4);
}

// Shared-memory load at [byte address + compile-time offset] (the offset lands in the LDS immediate field).
template <int OFF> __device__ __forceinline__ float lds_at(uint32_t addr, float) {
    float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ double lds_at(uint32_t addr, double) {
    double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF)); return v;
}
// ---- experiment (VERDICT W10 / W11): stage the RAW parent-CDF window with one bulk copy (TMA, cp.async.bulk + mbarrier) -------
// Compile with -DSGM_TMA_WINDOW=1.  The window's whole tiles are contiguous in `fine`, so one elected lane issues ONE 1-4 KB
// bulk copy instead of every lane's 2 LDG.128 + 8 FMA + 2 STS.128 per tile; the children then pick their parent tile by
// comparing against the <= 4 tile offsets in registers and search in tile-local units with a rescaled TARGET.
#ifndef SGM_TMA_WINDOW
#define SGM_TMA_WINDOW 0
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_window(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n .reg .pred p;\n W_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra W_DONE;\n bra W_LOOP;\n W_DONE:\n}"
                 :: "r"(b), "r"(parity) : "memory");
}

template <class R, int STEP>
__device__ __forceinline__ void search_levels(uint32_t* ad, const R* rt) {
    if constexpr (STEP > 0) {
#pragma unroll
        for (int c = 0; c < KPT; ++c)
            if (lds_at<(STEP - 1) * (int)sizeof(R)>(ad[c], (R)0) <= rt[c]) ad[c] += STEP * (int)sizeof(R);
        search_levels<R, STEP / 2>(ad, rt);
    }
}

// ---- one resample -> propagate -> reweight -> statistic-update step (pf.py:7-38, 138-181, 40-82) ---
// SORTED = false: iid resampling uniforms (exact reference semantics, pf.py:27-29): per child a binary
//   search over the tile offsets and inside one tile, the 8 children of a lane interleaved.
// SORTED = true : ascending targets (order-statistics multinomial / systematic / stratified, or INJECTED
//   uniforms the caller declares sorted).  The warp's 256 children hit ONE contiguous parent range
//   [lo, hi]: a paired warp-cooperative search finds it, the CDF of the range is staged in the warp's slice
//   of shared memory in global units, each child binary-searches that slice (8 interleaved searches per
//   lane, neighbouring lanes read neighbouring words), and parent records are gathered as a stream.
// FM != FM_GENERIC = the production configurations, checked by the host before it picks the instantiation: device
// randoms, order-statistics resampling, the model score as the statistic, no traces / exported log-weights, and
//   FM_POY    Poyiadjis O(N)  (pf = nemeth with lambduh = 1)
//   FM_SHRINK Nemeth          (lambduh < 1)
//   FM_FILTER pf = filter
// They turn the run-time flags below into constants (fewer uniform branches, constant loads and -- unless RAGGED,
// i.e. N is not a multiple of 256 -- bound checks: ~9 % of the generic kernel's instructions); the particle system
// (positions, weights, genealogy) is bit-identical to the generic instantiation.  FM_SHRINK / FM_FILTER additionally
// accumulate the per-tile weighted statistic sums in registers while they propagate (WsCarry) instead of re-reading
// the freshly written records: same sums up to f32 / f64 rounding of the statistic (not of the particle system).
// Order statistics of N iid uniforms via exponential spacings.  Within a tile the normalised partial sums of P Exp(1) draws
// are independent of their total, which is Gamma(P, 1); the tile totals are drawn directly (pf_header_kernel), so no
// cross-tile scan is needed.  Lane l draws the spacings of ranks 8 l .. 8 l + 7 (lane-major scan), then the positions are
// transposed to row-major (child slot 32 c + lane has rank 32 c + lane) for coalesced gathers.
// Exp(1) spacings up to a common factor: the positions are normalised by the tile's total, so -ln u and log2 u (same sign
// throughout, no scaling multiply, no negation) give identical positions.
template <class R>
__device__ __forceinline__ void draw_spacings(const KArgs& a, const RngKey& key, int t, int q_me, int lane, int n_valid, R* s_tr, R* u) {
    if (sizeof(R) == 8 && a.variates32) {
        float uf[KPT];
        draw_uniforms<float>(key, t, q_me, lane, uf);
#pragma unroll
        for (int k = 0; k < KPT; ++k) uf[k] = (8 * lane + k < n_valid) ? Mth<float>::log2(uf[k]) : 0.0f;
        widen8(uf, u);
    } else {
        draw_uniforms<R>(key, t, q_me, lane, u);
#pragma unroll
        for (int k = 0; k < KPT; ++k) u[k] = (8 * lane + k < n_valid) ? Mth<R>::logb(u[k]) : (R)0;
    }
    R etot;
    lane_major_incl_scan<R>(u, etot);
    const R inv = Mth<R>::rcp(etot);
#pragma unroll
    for (int k = 0; k < KPT; ++k) u[k] = Mth<R>::mul(u[k], inv);
    __syncwarp();
    store_lane_major<R>(s_tr, lane, u);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < KPT; ++c) u[c] = s_tr[32 * c + lane];
}
template <class R, class Model, bool SORTED, int FM = FM_GENERIC, bool RAGGED = false, int WINB = WIN_BYTES>
__device__ __forceinline__ void step_body(const KArgs& a, int b, int t, int q_me, int lane, R* s_cdf, const double* hdr_local = nullptr,
                                          uint64_t* mbar = nullptr) {
    constexpr bool FAST = FM != FM_GENERIC;
    static_assert(!FAST || SORTED, "the fast modes imply sorted resampling");
    R* const s_tr = s_cdf;                 // the warp's shared-memory slice doubles as the scan transposition buffer
    constexpr int NP = Model::NP, W = Model::NX + NP;
    const int N = a.N, par = t & 1;
    if (q_me >= a.Q) return;
    // every load that does not depend on the randoms is issued up front (one round trip for all of them):
    // activity flag, item header scalars, level-1 coarse probe, Gamma prefix, (y_t, w_t), model constants
    const int Tb = a.T_buf[b];
    const ItemHdr hdr = load_hdr(a, b, hdr_local);
    const int s1c = (a.Q + 31) >> 5;
    const double c1 = (SORTED && lane * s1c < a.Q) ? hdr.off[lane * s1c] : Mth<double>::inf();
    const R* ywp = reinterpret_cast<const R*>(a.yw) + ((size_t)b * a.max_T + min(t, a.max_T - 1)) * 2;
    const R y_t = ywp[0], w_t = ywp[1];
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const bool spacings = FAST || (SORTED && uses_spacings(a));
    const bool injected = !FAST && a.rng_mode == SGM_RNG_INJECTED;
    const double gam_lo = spacings ? hdr.gam[q_me] : 0.0;
    const double gam_hi = spacings ? hdr.gam[q_me + 1] : 0.0;
    if (t >= Tb) return;
    const size_t item_off = (size_t)b * N;
    const int nws = FAST ? NP : stat_width<Model>(a.stat_kind);
    const bool carries = FAST ? (FM != FM_FILTER) : (a.pf == SGM_PF_NEMETH);    // stats follow the resampled genealogy here
    const bool shrink = FAST ? (FM == FM_SHRINK) : (carries && (a.lambduh != 1.0));
    const R* fine_old = reinterpret_cast<const R*>(a.fine[par]) + (size_t)b * a.Q * WT;      // padded to whole tiles
    const int tile_base = q_me * WT;
    const int n_valid = (FAST && !RAGGED) ? WT : min(WT, N - tile_base);
    const RngKey key = item_key(a, b);
    const double total = hdr.total;
    const double tmax = total * (1.0 - 1.2e-16);
    int anc[KPT];
    R z[KPT];

    if (!SORTED) {
        double target[KPT];
        if (a.rng_mode == SGM_RNG_INJECTED) {
            const double* u = a.inj_u + ((size_t)b * a.max_T + t) * N;
#pragma unroll
            for (int c = 0; c < KPT; ++c) { const int i = tile_base + 32 * c + lane; target[c] = (i < N) ? u[i] * total : 0.0; }
        } else {
            R u[KPT];
            draw_uniforms<R>(key, t, q_me, lane, u, a.variates32 != 0);
#pragma unroll
            for (int c = 0; c < KPT; ++c) target[c] = (double)u[c] * total;
        }
        int qsel[KPT], pos[KPT], len[KPT];
        R rr[KPT];
#pragma unroll
        for (int c = 0; c < KPT; ++c) { if (!(target[c] < total)) target[c] = tmax; qsel[c] = 0; }
#pragma unroll 1
        for (int step = pow2_floor(hdr.Q); step > 0; step >>= 1) {
#pragma unroll
            for (int c = 0; c < KPT; ++c)
                if (qsel[c] + step < hdr.Q && hdr.off[qsel[c] + step] <= target[c]) qsel[c] += step;
        }
#pragma unroll
        for (int c = 0; c < KPT; ++c) {
            rr[c] = (R)((target[c] - hdr.off[qsel[c]]) / hdr.sc[qsel[c]]);
            len[c] = min(WT, N - qsel[c] * WT);
            pos[c] = 0;
        }
#pragma unroll 1
        for (int step = WT / 2; step > 0; step >>= 1) {
#pragma unroll
            for (int c = 0; c < KPT; ++c) {
                const int idx = pos[c] + step;
                if (idx <= len[c] && fine_old[qsel[c] * WT + idx - 1] <= rr[c]) pos[c] = idx;
            }
        }
#pragma unroll
        for (int c = 0; c < KPT; ++c) anc[c] = qsel[c] * WT + min(pos[c], len[c] - 1);
        draw_normals<R>(a, key, b, t, q_me, lane, z);
    } else {
        // u[c]: for the order-statistics sampler the tile-local position in (0, 1] of child 32 c + lane,
        // else its uniform.  target of that child = tA + tB * (u[c] (+ 32 c + lane))
        R u[KPT];
        double tA, tB;
        if (spacings) {
            draw_spacings<R>(a, key, t, q_me, lane, n_valid, s_tr, u);
            tA = gam_lo; tB = gam_hi - gam_lo;
        } else {
            if (injected) {
#pragma unroll
                for (int c = 0; c < KPT; ++c) u[c] = (R)0;            // targets come from inj_u (target_of)
            } else if (a.resample == SGM_RESAMPLE_SYSTEMATIC) {
                R u4[4];
                rng_uniform4(key, 0u, (uint32_t)t, STREAM_GAMMA, u4);
#pragma unroll
                for (int c = 0; c < KPT; ++c) u[c] = u4[0];
            } else {
                draw_uniforms<R>(key, t, q_me, lane, u, a.variates32 != 0);
            }
            tB = total / (double)N; tA = (double)tile_base * tB;
        }
        draw_normals<R>(a, key, b, t, q_me, lane, z);
        auto target_of = [&](int c) -> double {
            double tg;
            const int i = tile_base + 32 * c + lane;
            if (injected) tg = (i < N) ? a.inj_u[((size_t)b * a.max_T + t) * N + i] * total : 0.0;
            else if (spacings) tg = tA + tB * (double)u[c];
            else tg = tA + tB * ((double)(32 * c + lane) + (double)u[c]);
            return (tg < total) ? tg : tmax;
        };
        // ---- parent tiles of this warp tile: first child = (row 0, lane 0), last = particle n_valid - 1 ----
        const int last = n_valid - 1, l_last = last & 31, c_last = last >> 5;
        double tf, tl;
        if (injected) {
            const double* iu = a.inj_u + ((size_t)b * a.max_T + t) * N + tile_base;
            tf = iu[0] * total; tl = iu[last] * total;
        } else {
            R ul = u[KPT - 1];
            if (n_valid != WT) {
#pragma unroll
                for (int c = 0; c < KPT - 1; ++c) if (c == c_last) ul = u[c];
            }
            ul = __shfl_sync(FULL, ul, l_last);
            const R uf = __shfl_sync(FULL, u[0], 0);
            tf = spacings ? tA + tB * (double)uf : tA + tB * (double)uf;
            tl = spacings ? tA + tB * (double)ul : tA + tB * ((double)last + (double)ul);
        }
        tf = (tf < total) ? tf : tmax;
        tl = (tl < total) ? tl : tmax;
        int q_lo, q_hi;
        warp_search_tiles(tf, tl, c1, s1c, hdr, lane, q_lo, q_hi);
        constexpr int MAXT = WINB / (int)(WT * sizeof(R));              // tiles the window holds: 4 (2 for f64 in the generic and cooperative kernels)
        const int nst = q_hi - q_lo + 1;
        if (SGM_TMA_WINDOW && FAST && mbar != nullptr && nst >= 1 && nst <= MAXT) {
            const int wt = (nst <= 1) ? 1 : ((nst <= 2) ? 2 : 4);
            const int live_t = min(wt, a.Q - q_lo);
            constexpr int ES = (int)sizeof(R);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the slice was the transposition buffer (generic proxy)
            __syncwarp();
            if (lane == 0) bulk_load_window(s_cdf, fine_old + (size_t)q_lo * WT, (uint32_t)(live_t * WT * ES), mbar);
            const double cbase = hdr.off[q_lo];
            R o[MAXT], isc[MAXT];
#pragma unroll
            for (int j = 0; j < MAXT; ++j) {
                const bool live = j < live_t;
                o[j] = live ? (R)(hdr.off[q_lo + (live ? j : 0)] - cbase) : Mth<R>::inf();
                isc[j] = live ? Mth<R>::rcp((R)hdr.sc[q_lo + (live ? j : 0)]) : (R)0;
            }
            const R rA = (R)(tA - cbase), rB = (R)tB;
            R rt[KPT];
            uint32_t ad[KPT];
            const uint32_t ad0 = (uint32_t)__cvta_generic_to_shared(s_cdf);
#pragma unroll
            for (int c = 0; c < KPT; ++c) {
                const R g = Mth<R>::fma(rB, u[c], rA);                       // spacings (fast modes)
                R oj = o[0], ij = isc[0];
                uint32_t adj = ad0;
#pragma unroll
                for (int j = 1; j < MAXT; ++j)
                    if (g >= o[j]) { oj = o[j]; ij = isc[j]; adj = ad0 + j * WT * ES; }
                rt[c] = Mth<R>::mul(g - oj, ij);
                ad[c] = adj;
            }
            mbar_wait(mbar, 0u);
            search_levels<R, WT / 2>(ad, rt);
#pragma unroll
            for (int c = 0; c < KPT; ++c) anc[c] = min(q_lo * WT + (int)((ad[c] - ad0) / ES), N - 1);
            __syncwarp();
        } else if (nst >= 1 && nst <= MAXT) {
            // Stage the CDF of whole parent tiles q_lo .. q_lo + wt - 1 (wt = 1, 2 or 4: a power-of-two window, so
            // the searches below need neither bound checks nor a run-time step) in global units relative to
            // cbase = off[q_lo]; f32 is enough, the window spans a few tiles.  `fine` is padded to whole tiles, so
            // every tile is 2 aligned vector loads per lane; tiles beyond the item read as +inf.
            const int wt = (nst <= 1) ? 1 : ((nst <= 2) ? 2 : 4);
            const double cbase = hdr.off[q_lo];
            const Vec4T<R>* src = reinterpret_cast<const Vec4T<R>*>(fine_old + (size_t)q_lo * WT);
            Vec4T<R>* dst = reinterpret_cast<Vec4T<R>*>(s_cdf);
#pragma unroll
            for (int j0 = 0; j0 < MAXT; j0 += 2) {
                if (j0 < wt) {
                    Vec4T<R> f[2][2];
                    R o[2], sc[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const bool live = (j0 + j < wt) && (q_lo + j0 + j < a.Q);
                        o[j] = Mth<R>::inf(); sc[j] = (R)0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) { f[j][h].x = f[j][h].y = f[j][h].z = f[j][h].w = (R)0; }
                        if (live) {
                            f[j][0] = src[(j0 + j) * (WT / 4) + lane];
                            f[j][1] = src[(j0 + j) * (WT / 4) + 32 + lane];
                            o[j] = (R)(hdr.off[q_lo + j0 + j] - cbase);
                            sc[j] = (R)hdr.sc[q_lo + j0 + j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j0 + j < wt) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                Vec4T<R> v;
                                v.x = Mth<R>::fma(f[j][h].x, sc[j], o[j]); v.y = Mth<R>::fma(f[j][h].y, sc[j], o[j]);
                                v.z = Mth<R>::fma(f[j][h].z, sc[j], o[j]); v.w = Mth<R>::fma(f[j][h].w, sc[j], o[j]);
                                dst[(j0 + j) * (WT / 4) + 32 * h + lane] = v;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            R rt[KPT];
            if (injected) {
#pragma unroll
                for (int c = 0; c < KPT; ++c) rt[c] = (R)(target_of(c) - cbase);
            } else {
                const R rA = (R)(tA - cbase), rB = (R)tB;
#pragma unroll
                for (int c = 0; c < KPT; ++c)
                    rt[c] = spacings ? Mth<R>::fma(rB, u[c], rA) : Mth<R>::fma(rB, Mth<R>::add((R)(32 * c + lane), u[c]), rA);
            }
            // 8 independent branch-free binary searches per lane, interleaved; #{k : s_cdf[k] <= rt} is kept as a
            // shared-memory byte address: LDS [addr + immediate], compare, predicated add -- 3 instructions / level
            uint32_t ad[KPT];
            const uint32_t ad0 = (uint32_t)__cvta_generic_to_shared(s_cdf);
#pragma unroll
            for (int c = 0; c < KPT; ++c) ad[c] = ad0;
            constexpr int ES = (int)sizeof(R);
            if (MAXT >= 4 && wt == 4) {
#pragma unroll
                for (int c = 0; c < KPT; ++c) if (lds_at<(2 * WT - 1) * ES>(ad0, (R)0) <= rt[c]) ad[c] += 2 * WT * ES;
            }
            if (wt >= 2) {
#pragma unroll
                for (int c = 0; c < KPT; ++c) if (lds_at<(WT - 1) * ES>(ad[c], (R)0) <= rt[c]) ad[c] += WT * ES;
            }
            search_levels<R, WT / 2>(ad, rt);
#pragma unroll
            for (int c = 0; c < KPT; ++c) anc[c] = min(q_lo * WT + (int)((ad[c] - ad0) / ES), N - 1);
            __syncwarp();                   // the window is reused by this warp's next tile (fused kernel)
        } else {
            // very uneven weights: the tile spans more parent tiles than the staging window holds
#pragma unroll
            for (int c = 0; c < KPT; ++c) anc[c] = search_hdr<R>(target_of(c), hdr, fine_old, N);
        }
    }
    R lwn[KPT];
    WsCarry<R> carry;
    propagate_store<R, Model, FM, RAGGED>(a, b, t, par, tile_base, lane, item_off, anc, z, hdr, nws, carries, shrink, lwn, th, y_t, w_t, &carry);
    constexpr bool CARRY_WS = FM == FM_SHRINK || FM == FM_FILTER;
    const bool need_ws = (nws > 0) && ((!FAST && a.pf == SGM_PF_FILTER) || shrink || (carries && t == Tb - 1));
    warp_tile_epilogue<R, W>(lwn, tile_base, N, lane, reinterpret_cast<R*>(a.fine[par ^ 1]) + (size_t)b * a.Q * WT,
                             a.sub[par ^ 1] + ((size_t)b * a.Q + q_me) * SSTRIDE, a.rec[par ^ 1], a.tail[par ^ 1],
                             item_off, need_ws, nws, s_tr, CARRY_WS ? &carry : nullptr);
}

// CTA shape of the step kernel: its warps are independent (no block barrier), so the CTA size only sets the
// register-allocation granularity.  Generic instantiations: 8 warps x SGM_STEP_CTAS (4) CTAs / SM at 64 registers.
// FAST instantiation: 4 warps x 9 CTAs / SM at 56 registers = 36 resident warps (measured: 1.31e11 -> 1.36e11
// particle-steps/s; 10 CTAs at 48 registers and the 8-warp shape at 64 registers are both slower).
#ifndef SGM_STEP_WARPS
#define SGM_STEP_WARPS 8
#endif
#ifndef SGM_FAST_WARPS
#define SGM_FAST_WARPS 4
#endif
#ifndef SGM_FAST_CTAS
#define SGM_FAST_CTAS 9
#endif
#ifndef SGM_FAST64_WARPS
#define SGM_FAST64_WARPS 4
#endif
#ifndef SGM_FAST64_CTAS
#define SGM_FAST64_CTAS 4          /* Nemeth / filter fast modes: 4 CTAs; 5 costs GARCH Nemeth 8 % (92 registers) */
#endif
#ifndef SGM_FAST64_CTAS_POY
/* FM_POY, measured with the table-driven variate transforms (N = 2^16, 256 items, same box, % of the HBM roofline):
 * 4 CTAs -> SVM 63.0, LGSSM 69.6, GARCH 65.2;  5 -> 65.1, 73.1, 63.8-66.0;  6 (80 registers, no spills) -> 58.5, 64.7, 54.6 */
#define SGM_FAST64_CTAS_POY 5
#endif
#ifndef SGM_FAST_CTAS_SHRINK
/* f32 Nemeth-with-shrinkage fast mode (carries the weighted statistic sums in registers): 8 CTAs = 64 registers, no spill.
 * Same box, % of the HBM roofline, 9 -> 8 CTAs: SVM 79.4 -> 79.2, LGSSM 82.2 -> 84.0, GARCH 83.1 -> 89.5 (7 CTAs: 73 / 80 / 86).
 * The filter fast mode stays at 9 (8: SVM 81 -> 78.5, GARCH 99.6 -> 97). */
#define SGM_FAST_CTAS_SHRINK 8
#endif
template <class R, bool FAST, int FM = FM_GENERIC> struct StepShape {
    static constexpr int WARPS = FAST ? (sizeof(R) == 4 ? SGM_FAST_WARPS : SGM_FAST64_WARPS) : SGM_STEP_WARPS;
    static constexpr int CTAS = FAST ? (sizeof(R) == 4 ? (FM == FM_SHRINK ? SGM_FAST_CTAS_SHRINK : SGM_FAST_CTAS)
                                                       : (FM == FM_POY ? SGM_FAST64_CTAS_POY : SGM_FAST64_CTAS))
                                     : (sizeof(R) == 4 ? SGM_STEP_CTAS : 2 * 8 / SGM_STEP_WARPS);
};
template <class R, class Model, bool SORTED, int FM = FM_GENERIC, bool RAGGED = false>
__global__ void __launch_bounds__(32 * StepShape<R, FM != FM_GENERIC, FM>::WARPS, StepShape<R, FM != FM_GENERIC, FM>::CTAS) pf_step_kernel(KArgs a, int t) {
    constexpr int SW = StepShape<R, FM != FM_GENERIC, FM>::WARPS;
    // four parent tiles per warp in the fast modes (4 KB f32 / 8 KB f64); the generic f64 instantiation keeps two
    constexpr int WINB = (FM != FM_GENERIC) ? 4 * WT * (int)sizeof(R) : WIN_BYTES;
    __shared__ __align__(32) R s_cdf_all[SW][SORTED ? WINB / sizeof(R) : WT];
    const int warp = threadIdx.x >> 5;
#if SGM_TMA_WINDOW
    __shared__ __align__(8) uint64_t s_mbar[SW];
    if (FM != FM_GENERIC) {
        if ((threadIdx.x & 31) == 0) mbar_init(&s_mbar[warp], 1);
        __syncwarp();
    }
    step_body<R, Model, SORTED, FM, RAGGED, WINB>(a, a.b0 + blockIdx.y, t, blockIdx.x * SW + warp, threadIdx.x & 31, s_cdf_all[warp], nullptr,
                                                  FM != FM_GENERIC ? &s_mbar[warp] : nullptr);
#else
    step_body<R, Model, SORTED, FM, RAGGED, WINB>(a, a.b0 + blockIdx.y, t, blockIdx.x * SW + warp, threadIdx.x & 31, s_cdf_all[warp]);
#endif
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
