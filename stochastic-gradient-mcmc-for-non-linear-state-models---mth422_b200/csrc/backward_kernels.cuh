// Kernels that run after pf_step_kernel has produced the new particles of a step:
//   * Poyiadjis O(N^2)  (pf.py:84-136)   -- flash-attention-style restatement: rank-1 scores (one FMA per pair),
//     bounded-reference softmax (no running max, partial sums over parent ranges add), statistics through the moments
//     of two per-parent features; the P V contraction on the tensor cores (TF32 mma.sync, f32) or on the FP32 / FP64
//     pipe; split-J grid + deterministic finish kernel.  The (N, N) backward-weight matrix is never materialised.
//   * PaRIS            (pf.py:183-341)   -- accept-reject backward sampling with exact fallback.
//       PHILOX mode : guide-table proposals, shared-memory work queue compacted every round, multi-proposal
//                     rounds for short queues, per-child acceptance bound, warp-per-entry exact sampler.
//       INJECTED mode: one CTA per item replays the reference's round structure (stable compaction
//                     of the unresolved list L, uniforms consumed at a running offset) so that the
//                     recorded numpy stream lines up draw for draw.
//   * predictive log-likelihood statistic of the particle filter (pf.py:40-82 `logsumexp`), pf_pred_kernel.
#pragma once
#include "pf_kernels.cuh"

namespace sgm {

constexpr int JT = 512;        // old particles per shared-memory tile in the O(N^2) kernel
constexpr int EXACT_CTAS = 32; // CTAs per item for the PaRIS exact fallback

template <class R, class Model>
__device__ __forceinline__ void stat_of(const KArgs& a, const typename Model::template Theta<R>& th, const R* xa, const R* xn, R y, bool in_sub, R* h) {
    h[0] = h[1] = h[2] = h[3] = (R)0;
    if (in_sub) {
        if (a.stat_kind == SGM_STAT_SCORE) Model::score(th, xa, xn, y, h);
        else if (a.stat_kind == SGM_STAT_SUFF) Model::suff(xa, xn, h);
    }
}

// ---- Poyiadjis O(N^2) ------------------------------------------------------------------------------
// For every new particle i:  tau'_i = sum_j bw_ij (tau_j + w_t h(x_j, x'_i)),  bw_i. = softmax_j(lw_j + log q(x'_i | x_j))
// (pf.py:115-135).  Flash-attention-style restatement, the (N, N) matrix never exists:
//   * score: for all three models log q(x'_i | x_j) - log_trans_max = b_i u_j + gq_j + a_i (models.cuh pair_child /
//     pair_parent), so one FMA per pair gives the exponent:  p_ij = 2^(b'_i u_j + g_j + c_i),
//     g_j = (lw_j - M + gq_j) log2 e  (M = max_j lw_j from the item header).
//   * bounded reference instead of a running max: lw_j <= M and log q <= log_trans_max, so every exponent is <= 0
//     -- no overflow, no rescaling, and partial sums over disjoint parent ranges simply ADD (split-J grid).
//     Children whose total underflows (exponents below ~-87 nats everywhere: never seen) are redone exactly.
//   * value: h is a polynomial in x_j, so sum_j p_ij h needs only the moments of m_j (2 columns):
//     P V with V_j = [tau_j (p columns) | m_j (2) | 1], then an O(1) epilogue per child (score_moments).
// Work split: a thread owns N2_CPT children (accumulators in registers), a CTA streams tiles of N2_JT parents
// through shared memory as two float4 per parent (broadcast LDS.128), grid = (child blocks, items, parent splits).
constexpr int N2_CPT = 4;
constexpr int N2_JT = 512;
constexpr int N2_TARGET_CTAS = 8 * 148 * 3;

struct N2Plan { int child_blocks, splits, chunk; };
inline N2Plan n2_plan(int B, int N) {
    N2Plan p;
    p.child_blocks = (N + NT * N2_CPT - 1) / (NT * N2_CPT);
    const int tiles = (N + N2_JT - 1) / N2_JT;
    // >= ~8 waves of CTAs (3 resident per SM): the last, partially filled wave then costs <= ~12 %
    int want = (N2_TARGET_CTAS + B * p.child_blocks - 1) / (B * p.child_blocks);
    if (want > 64) want = 64;
    if (want > tiles) want = tiles;
    if (want < 1) want = 1;
    const int tiles_per = (tiles + want - 1) / want;
    p.chunk = tiles_per * N2_JT;
    p.splits = (tiles + tiles_per - 1) / tiles_per;
    return p;
}

// exact (running-max free, two-pass) evaluation of one child, thread-serial over all parents: the fallback of
// the bounded-reference kernel and the direct transcription of pf.py:115-135
template <class R, class Model>
__device__ void n2_exact_child(const KArgs& a, const typename Model::template Theta<R>& th, int par, size_t item_off,
                               const R* xi, R y, R wt, bool in_sub, R* out) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const R* lw_old = reinterpret_cast<const R*>(a.lw[par]) + item_off;
    R m = -Mth<R>::inf();
    for (int j = 0; j < a.N; ++j) {
        R rj[W];
        load_rec<R, W>(a.rec[par], a.tail[par], item_off + j, rj);
        m = nan_max(m, lw_old[j] + Model::log_trans(th, rj + NP, xi));
    }
    R l = (R)0, acc[4] = {(R)0, (R)0, (R)0, (R)0};
    for (int j = 0; j < a.N; ++j) {
        R rj[W], h[4];
        load_rec<R, W>(a.rec[par], a.tail[par], item_off + j, rj);
        const R p = Mth<R>::exp(lw_old[j] + Model::log_trans(th, rj + NP, xi) - m);
        stat_of<R, Model>(a, th, rj + NP, xi, y, in_sub, h);
        l += p;
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] += p * (rj[q] + h[q] * wt);
    }
    for (int q = 0; q < NP; ++q) out[q] = acc[q] / l;
}

// normalise, rebuild E_j[h] from the moments, write the child's statistic (state part of the record is kept)
template <class R, class Model>
__device__ __forceinline__ void n2_finalize(const KArgs& a, const typename Model::template Theta<R>& th, int par, size_t item_off,
                                            int i, R l, const R* acc, R y, R wt, bool in_sub) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int nws = stat_width<Model>(a.stat_kind);
    R rn[W];
    load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
    R st[4] = {(R)0, (R)0, (R)0, (R)0};
    if (l > (R)0 && l < Mth<R>::inf()) {
        const R il = (R)1 / l;
        const R Em[2] = {acc[NP] * il, acc[NP + 1] * il};
        R h[4] = {(R)0, (R)0, (R)0, (R)0};
        if (in_sub) {
            if (a.stat_kind == SGM_STAT_SCORE) Model::score_moments(th, Em, rn + NP, y, h);
            else if (a.stat_kind == SGM_STAT_SUFF) Model::suff_moments(Em, rn + NP, h);
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) st[q] = acc[q] * il + h[q] * wt;
    } else {
        n2_exact_child<R, Model>(a, th, par, item_off, rn + NP, y, wt, in_sub, st);
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) rn[q] = (q < nws) ? st[q] : (R)0;
    store_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
}

template <class R, class Model>
__global__ void __launch_bounds__(NT) poyiadjis_n2_kernel(KArgs a, int t, int splits, int chunk, R* part) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP, NV = NP + 2;
    __shared__ Vec4T<R> sA[N2_JT];          // (u_j, g_j, tau_j0, tau_j1)
    __shared__ Vec4T<R> sB[N2_JT];          // NP = 3: (tau_j2, m_j0, m_j1, -)   NP = 4: (tau_j2, tau_j3, m_j0, m_j1)
    const int b = blockIdx.y, z = blockIdx.z, tid = threadIdx.x;
    if (t >= a.T_buf[b]) return;
    const int N = a.N, par = t & 1;
    const size_t item_off = (size_t)b * N;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const R M = (R)a.hdr[(size_t)b * hdr_stride(a.Q) + H_M];          // max of the OLD log-weights (header of step t)
    const R L2E = (R)1.4426950408889634;
    const R* lw_old = reinterpret_cast<const R*>(a.lw[par]) + item_off;

    R bc[N2_CPT], cc[N2_CPT], l[N2_CPT], acc[N2_CPT][NV];
#pragma unroll
    for (int c = 0; c < N2_CPT; ++c) {
        const int i = (blockIdx.x * N2_CPT + c) * NT + tid;
        R rn[W];
#pragma unroll
        for (int q = 0; q < W; ++q) rn[q] = (R)0;
        if (i < N) load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
        R bb, aa;
        Model::pair_child(th, rn + NP, bb, aa);
        bc[c] = bb * L2E; cc[c] = aa * L2E; l[c] = (R)0;
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[c][q] = (R)0;
    }
    const int j_begin = z * chunk, j_end = min(N, j_begin + chunk);
    for (int j0 = j_begin; j0 < j_end; j0 += N2_JT) {
        __syncthreads();
        for (int j = tid; j < N2_JT; j += NT) {
            Vec4T<R> A, Bv;
            A.x = (R)0; A.y = -Mth<R>::inf(); A.z = A.w = (R)0; Bv.x = Bv.y = Bv.z = Bv.w = (R)0;
            if (j0 + j < j_end) {
                R rj[W], u, gq, m[2];
                load_rec<R, W>(a.rec[par], a.tail[par], item_off + j0 + j, rj);
                Model::pair_parent(th, rj + NP, u, gq, m);
                A.x = u; A.y = (lw_old[j0 + j] - M + gq) * L2E; A.z = rj[0]; A.w = rj[1];
                if (NP == 3) { Bv.x = rj[2]; Bv.y = m[0]; Bv.z = m[1]; }
                else { Bv.x = rj[2]; Bv.y = rj[NP - 1]; Bv.z = m[0]; Bv.w = m[1]; }
            }
            sA[j] = A; sB[j] = Bv;
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < N2_JT; ++j) {
            const Vec4T<R> A = sA[j], Bv = sB[j];
#pragma unroll
            for (int c = 0; c < N2_CPT; ++c) {
                const R p = Mth<R>::exp2(bc[c] * A.x + A.y + cc[c]);
                l[c] += p;
                acc[c][0] += p * A.z; acc[c][1] += p * A.w; acc[c][2] += p * Bv.x; acc[c][3] += p * Bv.y; acc[c][4] += p * Bv.z;
                if (NV == 6) acc[c][NV - 1] += p * Bv.w;
            }
        }
    }
    const R y = (R)a.obs[a.obs_off[b] + t];
    const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
    const R wt = in_sub ? ((a.wts_off && a.wts_off[b] >= 0) ? (R)a.step_weights[a.wts_off[b] + (t - a.t1[b])] : (R)1) : (R)0;
#pragma unroll
    for (int c = 0; c < N2_CPT; ++c) {
        const int i = (blockIdx.x * N2_CPT + c) * NT + tid;
        if (i >= N) continue;
        if (splits == 1) {
            n2_finalize<R, Model>(a, th, par, item_off, i, l[c], acc[c], y, wt, in_sub);
        } else {
            Vec4T<R> p0, p1;
            p0.x = acc[c][0]; p0.y = acc[c][1]; p0.z = acc[c][2]; p0.w = acc[c][3];
            p1.x = acc[c][4]; p1.y = (NV == 6) ? acc[c][NV - 1] : (R)0; p1.z = (R)0; p1.w = l[c];
            Vec4T<R>* dst = reinterpret_cast<Vec4T<R>*>(part) + (((size_t)z * a.B + b) * N + i) * 2;
            dst[0] = p0; dst[1] = p1;
        }
    }
}

// split-J: deterministic sum of the per-split partials (fixed order), then the same finalisation
template <class R, class Model>
__global__ void __launch_bounds__(NT) poyiadjis_n2_finish_kernel(KArgs a, int t, int splits, const R* part) {
    constexpr int NP = Model::NP, NV = NP + 2;
    const int b = blockIdx.y, i = blockIdx.x * NT + threadIdx.x;
    if (t >= a.T_buf[b] || i >= a.N) return;
    const int N = a.N, par = t & 1;
    const size_t item_off = (size_t)b * N;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    R acc[6] = {(R)0, (R)0, (R)0, (R)0, (R)0, (R)0}, l = (R)0;
    for (int z = 0; z < splits; ++z) {
        const Vec4T<R>* src = reinterpret_cast<const Vec4T<R>*>(part) + (((size_t)z * a.B + b) * N + i) * 2;
        const Vec4T<R> p0 = src[0], p1 = src[1];
        acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w; acc[4] += p1.x; acc[5] += p1.y; l += p1.w;
    }
    const R y = (R)a.obs[a.obs_off[b] + t];
    const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
    const R wt = in_sub ? ((a.wts_off && a.wts_off[b] >= 0) ? (R)a.step_weights[a.wts_off[b] + (t - a.t1[b])] : (R)1) : (R)0;
    (void)NV;
    n2_finalize<R, Model>(a, th, par, item_off, i, l, acc, y, wt, in_sub);
}

// ---- Poyiadjis O(N^2), tensor-core variant (f32 particles) -----------------------------------------------
// Same algebra; the P V contraction (V = [tau | m | 1], 8 columns) runs on the tensor cores:
// mma.sync.m16n8k8 TF32 with FP32 accumulators, A = exp-weights P (16 children x 8 parents) produced in
// registers straight into the A-fragment layout, B = V tile (8 parents x 8 columns) from shared memory.
// The FP32-pipe kernel above spends 6 of its 8 FMA-pipe instructions per pair on P V and is bound by that
// pipe (ncu: profiles/); here a pair costs FFMA + FADD + MUFU.EX2 + 1/64 MMA and the kernel is MUFU-bound.
// Precision: V is split hi + lo (two TF32 MMAs: ~21 mantissa bits); P keeps 10 mantissa bits (rounded to
// nearest by adding half an ulp before the tensor core truncates), and the SAME rounded p feeds numerator
// and denominator (the ones-column), so the rounding cancels to first order in tau' = sum p V / sum p.
// Results always go through the split-J partial buffer + poyiadjis_n2_finish_kernel.
constexpr int N2T_MT = 4;                // m16 child tiles per warp  -> 64 children / warp, 512 / CTA
constexpr int N2T_JT = 512;              // parents per shared-memory tile

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round x to the nearest integer n with the magic-number add,
// degree-3 minimax polynomial for 2^f on f = x - n in [-0.5, 0.5] (max relative error 1.0e-4, below half an ulp of
// the TF32 mantissa the result is rounded to), n added into the exponent field.  The tensor-core kernel is
// bound by the MUFU (XU) pipe -- 16 lanes / SM -- so it evaluates N2T_NPOLY of every 16 exponentials this
// way to balance the XU pipe against instruction issue (the FA4 trick).
#ifndef SGM_N2_NPOLY
#define SGM_N2_NPOLY 0          /* measured on B200 (N = 65536): 0 -> 3.41e12, 2 -> 3.32e12, 3 -> 3.19e12, 4 -> 3.09e12 pair-steps/s */
#endif
#ifndef SGM_N2T_CTAS
#define SGM_N2T_CTAS 3
#endif
constexpr int N2T_NPOLY = SGM_N2_NPOLY;
__device__ __forceinline__ float exp2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: n sits in the low mantissa bits of t
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05500893f, 0.24221095f);
    p = fmaf(p, f, 0.6932829f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ uint32_t tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32_16x8x8(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <class Model>
__global__ void __launch_bounds__(NT, SGM_N2T_CTAS) poyiadjis_n2_tc_kernel(KArgs a, int t, int chunk, float* part) {
    typedef float R;
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ float2 s_ug[N2T_JT];              // (u_j, g_j)
    __shared__ __align__(16) uint32_t s_vh[N2T_JT][8];         // V_j rounded to TF32
    __shared__ __align__(16) uint32_t s_vl[N2T_JT][8];         // TF32(V_j - hi)
    const int b = blockIdx.y, z = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (t >= a.T_buf[b]) return;
    const int N = a.N, par = t & 1;
    const size_t item_off = (size_t)b * N;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const R M = (R)a.hdr[(size_t)b * hdr_stride(a.Q) + H_M];
    const R L2E = 1.4426950408889634f;
    const R* lw_old = reinterpret_cast<const R*>(a.lw[par]) + item_off;
    const int g = lane >> 2, tq = lane & 3;
    const int child0 = (blockIdx.x * NWARP + warp) * (16 * N2T_MT);

    R bc[N2T_MT][2], cc[N2T_MT][2], C[N2T_MT][4];
#pragma unroll
    for (int m = 0; m < N2T_MT; ++m) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = child0 + 16 * m + g + 8 * r;
            R rn[W];
#pragma unroll
            for (int q = 0; q < W; ++q) rn[q] = (R)0;
            if (i < N) load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
            R bb, aa;
            Model::pair_child(th, rn + NP, bb, aa);
            bc[m][r] = bb * L2E; cc[m][r] = aa * L2E;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) C[m][q] = 0.f;
    }
    const int j_begin = z * chunk, j_end = min(N, j_begin + chunk);
    for (int j0 = j_begin; j0 < j_end; j0 += N2T_JT) {
        __syncthreads();
        for (int j = tid; j < N2T_JT; j += NT) {
            float2 ug = make_float2(0.f, -Mth<R>::inf());
            R v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (j0 + j < j_end) {
                R rj[W], u, gq, mm[2];
                load_rec<R, W>(a.rec[par], a.tail[par], item_off + j0 + j, rj);
                Model::pair_parent(th, rj + NP, u, gq, mm);
                ug.x = u; ug.y = (lw_old[j0 + j] - M + gq) * L2E;
#pragma unroll
                for (int q = 0; q < NP; ++q) v[q] = rj[q];
                v[NP] = mm[0]; v[NP + 1] = mm[1]; v[7] = 1.f;        // column order = partial-buffer layout
            }
            s_ug[j] = ug;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                hi[q] = tf32_rna(v[q]);
                lo[q] = tf32_rna(v[q] - __uint_as_float(hi[q]));
            }
            uint4* dh = reinterpret_cast<uint4*>(s_vh[j]);
            uint4* dl = reinterpret_cast<uint4*>(s_vl[j]);
            dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
        __syncthreads();
#pragma unroll 2
        for (int k0 = 0; k0 < N2T_JT; k0 += 8) {
            const float2 ug0 = s_ug[k0 + tq], ug1 = s_ug[k0 + tq + 4];
            const uint32_t bh0 = s_vh[k0 + tq][g], bh1 = s_vh[k0 + tq + 4][g];
            const uint32_t bl0 = s_vl[k0 + tq][g], bl1 = s_vl[k0 + tq + 4][g];
#pragma unroll
            for (int m = 0; m < N2T_MT; ++m) {
                uint32_t A[4];
                // +0x1000: half an ulp of the 10-bit TF32 mantissa, so the tensor core's truncation rounds to nearest
                A[0] = __float_as_uint(Mth<R>::exp2(bc[m][0] * ug0.x + ug0.y + cc[m][0])) + 0x1000u;
                A[1] = __float_as_uint(Mth<R>::exp2(bc[m][1] * ug0.x + ug0.y + cc[m][1])) + 0x1000u;
                A[2] = __float_as_uint(Mth<R>::exp2(bc[m][0] * ug1.x + ug1.y + cc[m][0])) + 0x1000u;
                const R x3 = bc[m][1] * ug1.x + ug1.y + cc[m][1];
                A[3] = __float_as_uint(m < N2T_NPOLY ? exp2_poly(x3) : Mth<R>::exp2(x3)) + 0x1000u;
                mma_tf32_16x8x8(C[m], A, bh0, bh1);
                mma_tf32_16x8x8(C[m], A, bl0, bl1);
            }
        }
    }
    // C fragment: rows g / g + 8, columns 2 tq, 2 tq + 1  ->  partial buffer [z][b][i][8]
#pragma unroll
    for (int m = 0; m < N2T_MT; ++m) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = child0 + 16 * m + g + 8 * r;
            if (i < N) {
                float2* dst = reinterpret_cast<float2*>(part + (((size_t)z * a.B + b) * N + i) * 8) + tq;
                *dst = make_float2(C[m][2 * r], C[m][2 * r + 1]);
            }
        }
    }
}

// Per-warp-tile weighted statistic sums for the final average (only on an item's last step).
template <class R, class Model>
__global__ void __launch_bounds__(NT) stat_ws_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (t != a.T_buf[b] - 1) return;
    const int q_me = g * NWARP + warp;
    if (q_me >= a.Q) return;
    const int N = a.N, par = (t & 1) ^ 1;          // buffers written by step t
    const size_t item_off = (size_t)b * N;
    const int nws = stat_width<Model>(a.stat_kind);
    double* sub = a.sub[par] + ((size_t)b * a.Q + q_me) * SSTRIDE;
    const R m = (R)sub[0];
    const R msafe = (m == -Mth<R>::inf()) ? (R)0 : m;
    const R* lw = reinterpret_cast<const R*>(a.lw[par]) + item_off;
    double ws[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < KPT; ++c) {
        const int i = q_me * WT + 32 * c + lane;
        if (i < N) {
            R r[W];
            load_rec<R, W>(a.rec[par], a.tail[par], item_off + i, r);
            const R w = Mth<R>::exp(lw[i] - msafe);
            for (int q = 0; q < nws; ++q) ws[q] += (double)(r[q] * w);
        }
    }
    for (int q = 0; q < nws; ++q) ws[q] = warp_sum(ws[q]);
    __syncwarp();
    if (lane == 0) for (int q = 0; q < 4; ++q) sub[2 + q] = ws[q];
}

inline N2Plan n2_plan_tc(int B, int N) {
    N2Plan p;
    p.child_blocks = (N + NWARP * 16 * N2T_MT - 1) / (NWARP * 16 * N2T_MT);
    const int tiles = (N + N2T_JT - 1) / N2T_JT;
    int want = (N2_TARGET_CTAS + B * p.child_blocks - 1) / (B * p.child_blocks);
    if (want > 64) want = 64;
    if (want > tiles) want = tiles;
    if (want < 1) want = 1;
    const int tiles_per = (tiles + want - 1) / want;
    p.chunk = tiles_per * N2T_JT;
    p.splits = (tiles + tiles_per - 1) / tiles_per;
    return p;
}
inline bool n2_use_tensor(int dtype, int n2_mode) { return dtype == SGM_F32 && n2_mode != SGM_N2_FP32_PIPE; }
inline size_t n2_partial_bytes(int dtype, int n2_mode, int B, int N) {
    if (n2_use_tensor(dtype, n2_mode)) return (size_t)n2_plan_tc(B, N).splits * B * N * 8 * 4;
    const N2Plan p = n2_plan(B, N);
    return p.splits > 1 ? (size_t)p.splits * B * N * 8 * (dtype == SGM_F64 ? 8 : 4) : 0;
}

template <class R, class Model> struct N2TcLaunch {
    static bool run(const KArgs&, int, cudaStream_t, int&) { return false; }
};
template <class Model> struct N2TcLaunch<float, Model> {
    static bool run(const KArgs& a, int t, cudaStream_t stream, int& n) {
        const N2Plan p = n2_plan_tc(a.B, a.N);
        float* part = reinterpret_cast<float*>(a.n2part);
        poyiadjis_n2_tc_kernel<Model><<<dim3(p.child_blocks, a.B, p.splits), NT, 0, stream>>>(a, t, p.chunk, part);
        poyiadjis_n2_finish_kernel<float, Model><<<dim3((a.N + NT - 1) / NT, a.B), NT, 0, stream>>>(a, t, p.splits, part);
        n += 2;
        return true;
    }
};

template <class R, class Model>
int launch_poyiadjis_n2(const KArgs& a, int t, cudaStream_t stream) {
    int n = 0;
    if (!(a.n2_tensor && N2TcLaunch<R, Model>::run(a, t, stream, n))) {
        const N2Plan p = n2_plan(a.B, a.N);
        R* part = reinterpret_cast<R*>(a.n2part);
        poyiadjis_n2_kernel<R, Model><<<dim3(p.child_blocks, a.B, p.splits), NT, 0, stream>>>(a, t, p.splits, p.chunk, part);
        ++n;
        if (p.splits > 1) {
            poyiadjis_n2_finish_kernel<R, Model><<<dim3((a.N + NT - 1) / NT, a.B), NT, 0, stream>>>(a, t, p.splits, part);
            ++n;
        }
    }
    stat_ws_kernel<R, Model><<<dim3(a.G, a.B), NT, 0, stream>>>(a, t);
    return n + 1;
}

// ---- k-step-ahead predictive log-likelihood (pf = 'filter', `logsumexp`) ----------------------------
// After step t has produced the new particles and log-weights: per particle h_ik = log Pr(y_{t+k} | x'_i), k = 0..K
// (models.cuh pred_*), then  stat_k += max_i h_ik + log( sum over i [and, as in the reference pf.py:73-76, over ALL k]
// of exp(h_ik - max_k) W_i ),  W = normalised new weights.  Outside [t1, tL) the statistic function is identically 0,
// for which the reference's all-horizon sum adds log(K + 1) to every entry.  One CTA per item, two passes over the
// particles (h is recomputed -- counter-based / injected randoms); an evaluation metric, not a throughput path.
template <class R, class Model>
__device__ __forceinline__ void pred_h(const KArgs& a, const typename Model::template Theta<R>& th, const RngKey& key, int b, int t,
                                       int i, int kmax, R wt, const double* obs, const R* xn, R* h) {
    constexpr int PS = SGM_PRED_SLOTS;
    R ps[2];
    Model::pred_begin(xn, ps);
#pragma unroll
    for (int k = 0; k < PS; ++k) h[k] = (R)0;
    for (int k = 0; k <= kmax; ++k) {
        R z = (R)0;
        if (Model::PRED_RNG) {
            if (a.rng_mode == SGM_RNG_INJECTED) z = (R)a.inj_pred[(((size_t)b * a.max_T + t) * PS + k) * a.N + i];
            else rng_normal1(key, (uint32_t)i, (uint32_t)t, STREAM_PRED, (uint32_t)k, z);
        }
        h[k] = Model::pred_ll(th, ps, (R)obs[t + k], z) * wt;
        Model::pred_next(th, ps, z);
    }
}

template <class R, class Model>
__global__ void __launch_bounds__(NT) pf_pred_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP, PS = SGM_PRED_SLOTS;
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int Tb = a.T_buf[b];
    if (t >= Tb) return;
    const int K = a.pred_K, N = a.N, par = t & 1;
    double* acc = a.acc + (size_t)b * ACC_STRIDE;
    const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
    if (!in_sub) {
        if (tid == 0 && !a.pred_per_horizon) for (int k = 0; k <= K; ++k) acc[1 + k] += ::log((double)(K + 1));
        return;
    }
    const R wt = (a.wts_off && a.wts_off[b] >= 0) ? (R)a.step_weights[a.wts_off[b] + (t - a.t1[b])] : (R)1;
    const size_t item_off = (size_t)b * N;
    const R* lw_new = reinterpret_cast<const R*>(a.lw[par ^ 1]) + item_off;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const double* obs = a.obs + a.obs_off[b];
    const int kmax = min(K, Tb - 1 - t);
    const RngKey key = item_key(a, b);
    R mx[PS], mlw = -Mth<R>::inf();
#pragma unroll
    for (int k = 0; k < PS; ++k) mx[k] = -Mth<R>::inf();
    for (int i = tid; i < N; i += NT) {
        R rn[W], h[PS];
        load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
        pred_h<R, Model>(a, th, key, b, t, i, kmax, wt, obs, rn + NP, h);
#pragma unroll
        for (int k = 0; k < PS; ++k) mx[k] = nan_max(mx[k], h[k]);
        mlw = nan_max(mlw, lw_new[i]);
    }
#pragma unroll
    for (int k = 0; k < PS; ++k) if (k <= K) mx[k] = block_max(mx[k], sh_r);      // K is uniform over the CTA
    mlw = block_max(mlw, sh_r);
    double s[PS], Wsum = 0.0;
#pragma unroll
    for (int k = 0; k < PS; ++k) s[k] = 0.0;
    for (int i = tid; i < N; i += NT) {
        R rn[W], h[PS];
        load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
        pred_h<R, Model>(a, th, key, b, t, i, kmax, wt, obs, rn + NP, h);
        const double w = (double)Mth<R>::exp(lw_new[i] - mlw);
        Wsum += w;
#pragma unroll
        for (int k = 0; k < PS; ++k) if (k <= K) s[k] += (double)Mth<R>::exp(h[k] - mx[k]) * w;
    }
#pragma unroll
    for (int k = 0; k < PS; ++k) if (k <= K) s[k] = block_sum(s[k], sh_d);
    Wsum = block_sum(Wsum, sh_d);
    if (tid == 0) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < PS; ++k) if (k <= K) tot += s[k] / Wsum;
#pragma unroll
        for (int k = 0; k < PS; ++k) if (k <= K) acc[1 + k] += (double)mx[k] + ::log(a.pred_per_horizon ? s[k] / Wsum : tot);
    }
}

// ---- PaRIS ------------------------------------------------------------------------------------------
// Exact draw(s) from the backward kernel of child i:  J ~ Cat(softmax_j(lw_j + log q(x'_i | x_j)))
// (pf.py:328-339 and the naive branch :226-237).  Block-cooperative, fixed summation order.
template <class R, class Model>
__device__ __forceinline__ void exact_backward_sample(const KArgs& a, const typename Model::template Theta<R>& th, int par,
                                                      size_t item_off, const R* xi, int nu, const double* u, int32_t* Jout,
                                                      int Jstride, R* sh_r, double* sh_d) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int N = a.N, tid = threadIdx.x;
    const R* lw_old = reinterpret_cast<const R*>(a.lw[par]) + item_off;
    const int per = (N + NT - 1) / NT;
    const int j0 = tid * per, j1 = min(N, j0 + per);
    R m = -Mth<R>::inf();
    for (int j = j0; j < j1; ++j) {
        R rj[W];
        load_rec<R, W>(a.rec[par], a.tail[par], item_off + j, rj);
        m = nan_max(m, lw_old[j] + Model::log_trans(th, rj + NP, xi));
    }
    m = block_max(m, sh_r);
    double loc = 0.0;
    for (int j = j0; j < j1; ++j) {
        R rj[W];
        load_rec<R, W>(a.rec[par], a.tail[par], item_off + j, rj);
        loc += (double)Mth<R>::exp(lw_old[j] + Model::log_trans(th, rj + NP, xi) - m);
    }
    double total;
    const double pre = block_excl_scan(loc, sh_d, total);
    for (int k = 0; k < nu; ++k) {
        double target = u[k] * total;
        if (!(target < total)) target = total * (1.0 - 1.2e-16);
        if (tid == 0) Jout[k * Jstride] = N - 1;
    }
    __syncthreads();
    for (int k = 0; k < nu; ++k) {
        double target = u[k] * total;
        if (!(target < total)) target = total * (1.0 - 1.2e-16);
        if (target >= pre && target < pre + loc) {
            double run = pre;
            int found = j1 - 1;
            for (int j = j0; j < j1; ++j) {
                R rj[W];
                load_rec<R, W>(a.rec[par], a.tail[par], item_off + j, rj);
                run += (double)Mth<R>::exp(lw_old[j] + Model::log_trans(th, rj + NP, xi) - m);
                if (run > target) { found = j; break; }
            }
            Jout[k * Jstride] = found;
        }
    }
    __syncthreads();
}

// ---- PHILOX mode (pf.py:260-341 in law) -------------------------------------------------------------------
// The accept-reject sampler proposes I ~ Cat(softmax(lw)) and accepts with probability q(x'_i | x_I) / q_max; the
// exact sampler draws from softmax_j(lw_j + log q(x'_i | x_j)) directly.  Both return exact draws of the backward
// kernel, so the cap on the number of proposals only moves cost between them (it is not part of the law): entries
// still unresolved after PARIS_CAP rounds go to the exact sampler.
//   paris_guide_kernel : flat f64 CDF of the OLD weights + a guide table (bucket k -> first particle whose
//                        cumulative mass exceeds k / N of the total), so a proposal is ~4 dependent loads
//                        instead of a 16-probe binary search over the hierarchical CDF.
//   paris_ar_kernel    : one CTA owns PARIS_CH children x Ntilde replicates; the unresolved entries live in a
//                        shared-memory work queue that is compacted every round (the reference's shrinking list
//                        L, pf.py:292-325), so threads never idle behind one unlucky child.
//   paris_exact_kernel : one WARP per leftover entry, two passes over the parents (total, then the crossing).
constexpr int PARIS_CH = 1024;         // children per CTA of the accept-reject kernel
#ifndef SGM_PARIS_CAP
#define SGM_PARIS_CAP 512
#endif
constexpr int PARIS_CAP = SGM_PARIS_CAP;   // proposals per entry before it falls back to the exact sampler (x Ntilde <= 8 fits the 12-bit Philox sub-counter)
#ifndef SGM_PARIS_U
#define SGM_PARIS_U 2      /* A/B (64 items, N = 2^14, ms per 60-step gradient; U = 1 / 2 / 4): GARCH 27.3 / 24.1 / 23.8, SVM 27.0 / 25.0 / 27.1, LGSSM 23.0 / 20.4 / 22.1 */
#endif
constexpr int PARIS_U = SGM_PARIS_U;   // queue entries a thread works on at once (memory-level parallelism)
constexpr int PARIS_MAXQ = 4096;       // queue capacity >= children per CTA * Ntilde (Ntilde > 4: fewer children per CTA)

template <class R, class Model>
__global__ void __launch_bounds__(NT) paris_guide_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int b = blockIdx.y, i = blockIdx.x * NT + threadIdx.x;
    if (t >= a.T_buf[b] || i > a.N) return;
    const int N = a.N, par = t & 1;
    const size_t item_off = (size_t)b * N;
    const ItemHdr hdr = load_hdr(a, b);
    const R* fine_old = reinterpret_cast<const R*>(a.fine[par]) + (size_t)b * a.Q * WT;
    double* cdf = a.pcdf + item_off;
    int32_t* guide = a.pguide + (size_t)b * (N + 1);
    if (i < N) {
        const int q = i / WT;
        cdf[i] = hdr.off[q] + (double)fine_old[i] * hdr.sc[q];
        guide[i] = search_hdr<R>((double)i * (hdr.total / (double)N), hdr, fine_old, N);
        // per-parent key of the rank-1 score (models.cuh pair_parent): log q(x'|x_j) - log_trans_max = b u_j + gq_j + a
        const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
        R rj[W], u, gq, mm[2];
        load_rec<R, W>(a.rec[par], a.tail[par], item_off + i, rj);
        Model::pair_parent(th, rj + NP, u, gq, mm);
        Vec4T<R> k;
        k.x = u; k.y = gq; k.z = reinterpret_cast<const R*>(a.lw[par])[item_off + i] - (R)hdr.M; k.w = (R)0;
        reinterpret_cast<Vec4T<R>*>(a.pkey)[item_off + i] = k;
    } else {
        guide[N] = N - 1;
    }
}

template <class R, class Model>
__global__ void __launch_bounds__(NT) paris_ar_kernel(KArgs a, int t, int ch) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ R s_x[2][PARIS_CH];                 // per child (b_i, a_i) of the rank-1 score
    __shared__ uint16_t s_q[2][PARIS_MAXQ];
    __shared__ int s_n[2];
    const int b = blockIdx.y, tid = threadIdx.x;
    if (t >= a.T_buf[b]) return;
    const int N = a.N, par = t & 1, Nt = a.Ntilde;
    const size_t item_off = (size_t)b * N;
    const int c0 = blockIdx.x * ch, nch = min(ch, N - c0);
    if (nch <= 0) return;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const double total = a.hdr[(size_t)b * hdr_stride(a.Q) + H_TOTAL];
    const Vec4T<R>* pkey = reinterpret_cast<const Vec4T<R>*>(a.pkey) + item_off;
    const double* cdf = a.pcdf + item_off;
    const int32_t* guide = a.pguide + (size_t)b * (N + 1);
    const RngKey key = item_key(a, b);
    for (int k = tid; k < nch; k += NT) {
        R rn[W], bb, aa;
        load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + c0 + k, rn);
        Model::pair_child(th, rn + NP, bb, aa);
        s_x[0][k] = bb; s_x[1][k] = aa - Model::pair_bound(th, rn + NP);       // accept w.p. q(x'|x_I) / sup_j q(x'|x_j)
    }
    int qlen = nch * Nt;
    for (int e = tid; e < qlen; e += NT) s_q[0][e] = (uint16_t)(((e / Nt) << 3) | (e % Nt));      // entry = (child << 3) | replicate
    if (tid == 0) { s_n[0] = qlen; s_n[1] = 0; }
    __syncthreads();
    // Proposal budget per entry (PARIS_CAP; the reference's max_accept_reject only lowers it).  While the queue is
    // long every thread tries one proposal of one entry per round; once it is shorter than the CTA, groups of
    // tpr = 2 .. 32 lanes try tpr consecutive proposals of the SAME entry at once and the first accepted one in
    // proposal order wins -- the same draw sequential accept-reject would have taken, in 1 / tpr of the rounds.
    const int cap = a.accept_reject ? min(a.max_ar, PARIS_CAP) : 0;
    int cur = 0, tries_done = 0;
    int n_prop = 0;                                  // proposals made by this thread (diagnostic: grad[6] of the item)
    while (tries_done < cap && qlen > 0) {
        int tpr = 1;
        while (tpr < 32 && qlen * tpr * 2 <= NT && tries_done + tpr * 2 <= cap) tpr *= 2;
        const int r = tid & (tpr - 1), gshift = (tid & 31) & ~(tpr - 1), per_round = NT / tpr;
        // PARIS_U entries per thread and pass with their dependent load chains (guide -> CDF bisection -> key) issued side by
        // side: the kernel is bound by L2 latency, not by bandwidth or issue (ncu round 2: issue active 25 %, long-scoreboard
        // stalls 16.7 per issue with all 64 warps resident), so memory-level parallelism per thread is the lever
        for (int base = 0; base < qlen; base += per_round * PARIS_U) {          // warp-uniform trip count: full-warp ballots
            bool live[PARIS_U], ok[PARIS_U];
            int entry[PARIS_U], I[PARIS_U], hi[PARIS_U], kk[PARIS_U], jt[PARIS_U];
            double target[PARIS_U];
            R ua[PARIS_U];
#pragma unroll
            for (int u = 0; u < PARIS_U; ++u) {
                const int e = base + u * per_round + tid / tpr;
                live[u] = e < qlen;
                ok[u] = false;
                entry[u] = 0; I[u] = 0; hi[u] = 0; kk[u] = 0; jt[u] = 0; target[u] = 0.0; ua[u] = (R)2;
                if (live[u]) {
                    ++n_prop;
                    entry[u] = s_q[cur][e];
                    kk[u] = entry[u] >> 3;
                    jt[u] = entry[u] & 7;
                    const uint4 raw = rng_raw(key, (uint32_t)(c0 + kk[u]), (uint32_t)t, STREAM_PARIS, (uint32_t)(jt[u] * PARIS_CAP + tries_done + r));
                    const double uu = u01d(raw.x, raw.y);
                    target[u] = uu * total;
                    if (!(target[u] < total)) target[u] = total * (1.0 - 1.2e-16);
                    const int kb = min((int)(uu * (double)N), N - 1);
                    I[u] = guide[kb];
                    hi[u] = guide[kb + 1];
                    ua[u] = (R)u01d(raw.z, raw.w);
                }
            }
            // first particle in [guide[kb], guide[kb + 1]] whose cumulative mass exceeds the target: usually 0-2
            // bisections; a bucket holds many particles only where the weights are tiny (degenerate steps)
            bool any = true;
            while (any) {
                any = false;
#pragma unroll
                for (int u = 0; u < PARIS_U; ++u) {
                    if (I[u] < hi[u]) {
                        const int mid = (I[u] + hi[u]) >> 1;
                        if (cdf[mid] <= target[u]) I[u] = mid + 1; else hi[u] = mid;
                        any = any || (I[u] < hi[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PARIS_U; ++u) {
                if (live[u]) {
                    const Vec4T<R> pk = pkey[I[u]];
                    const R thr = Mth<R>::exp(s_x[0][kk[u]] * pk.x + pk.y + s_x[1][kk[u]]);        // q(x'_i | x_I) / (per-child bound)
                    ok[u] = ua[u] <= thr;
                }
            }
#pragma unroll
            for (int u = 0; u < PARIS_U; ++u) {
                // first accepted proposal of the group, in proposal order
                const unsigned grp = (__ballot_sync(FULL, ok[u]) >> gshift) & (tpr == 32 ? 0xffffffffu : ((1u << tpr) - 1u));
                if (live[u]) {
                    if (grp) { if (r == __ffs(grp) - 1) a.Jidx[(item_off + c0 + kk[u]) * Nt + jt[u]] = I[u]; }
                    else if (r == 0) s_q[cur ^ 1][atomicAdd(&s_n[cur ^ 1], 1)] = (uint16_t)entry[u];
                }
            }
        }
        tries_done += tpr;
        __syncthreads();
        qlen = s_n[cur ^ 1];
        __syncthreads();
        if (tid == 0) s_n[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    n_prop = warp_sum(n_prop);
    if ((tid & 31) == 0) atomicAdd(a.counters + b * 16 + 1, n_prop);
    if (qlen > 0) {                                                  // leftovers -> exact sampler
        __shared__ int s_base;
        if (tid == 0) { s_base = atomicAdd(a.counters + b * 16, qlen); atomicAdd(a.counters + b * 16 + 2, qlen); }
        __syncthreads();
        for (int e = tid; e < qlen; e += NT) {
            const int entry = s_q[cur][e], k = entry >> 3, jt = entry & 7;
            a.Llist[0][item_off * Nt + s_base + e] = (c0 + k) * Nt + jt;
        }
    }
}

// one warp per leftover (child, replicate): J ~ Cat(softmax_j(lw_j + log q(x'_i | x_j)))   (pf.py:328-339).
// Weights relative to the bounded reference M + log_trans_max (every exponent <= 0: no max pass), evaluated from
// the per-parent keys with one FMA + one exp.  Pass 1 streams all parents (coalesced), keeping the total of every
// 2048-parent segment; the segment holding the target is then re-scanned with warp prefix sums.
constexpr int PSEG = 2048;
// Round 2: a whole CTA per entry (warp w takes the segments w, w + 8, ...) instead of one warp per entry -- the leftover
// lists are short (tens of entries per item), so one-warp-per-entry left 3/4 of the launched warps idle while a few ran
// 512 dependent iterations each (ncu round 2: 89 us per step, issue active 12 %, warps active 26 %).
template <class R, class Model>
__global__ void __launch_bounds__(NT) paris_exact_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ double s_seg[MAX_Q * WT / PSEG];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (t >= a.T_buf[b]) return;
    const int N = a.N, par = t & 1, Nt = a.Ntilde;
    const size_t item_off = (size_t)b * N;
    const int count = a.counters[b * 16];
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const Vec4T<R>* pkey = reinterpret_cast<const Vec4T<R>*>(a.pkey) + item_off;
    const RngKey key = item_key(a, b);
    const int nseg = (N + PSEG - 1) / PSEG;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int entry = a.Llist[0][item_off * Nt + e];
        const int i = entry / Nt, jt = entry - i * Nt;
        R rn[W], bb, aa;
        load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
        Model::pair_child(th, rn + NP, bb, aa);
        for (int sg = warp; sg < nseg; sg += NWARP) {
            R loc = (R)0;
            const int jend = min(N, (sg + 1) * PSEG);
            if (jend - sg * PSEG == PSEG) {
                // full segment: 8 independent key loads in flight per lane
#pragma unroll 8
                for (int j = sg * PSEG + lane; j < (sg + 1) * PSEG; j += 32) {
                    const Vec4T<R> pk = pkey[j];
                    loc += Mth<R>::exp(bb * pk.x + pk.y + pk.z + aa);
                }
            } else {
                for (int j = sg * PSEG + lane; j < jend; j += 32) {
                    const Vec4T<R> pk = pkey[j];
                    loc += Mth<R>::exp(bb * pk.x + pk.y + pk.z + aa);
                }
            }
            const double st = warp_sum((double)loc);
            if (lane == 0) s_seg[sg] = st;
        }
        __syncthreads();
        if (warp == 0) {
            double tot = 0.0;
            for (int sg = 0; sg < nseg; ++sg) tot += s_seg[sg];          // fixed order: independent of the warp schedule
            const uint4 raw = rng_raw(key, (uint32_t)i, (uint32_t)t, STREAM_EXACT, (uint32_t)jt);
            double target = u01d(raw.x, raw.y) * tot;
            if (!(target < tot)) target = tot * (1.0 - 1.2e-16);
            // segment of the crossing (sequential over <= 512 segment totals, warp-uniform)
            int sg = 0;
            double run = 0.0;
            while (sg < nseg - 1 && run + s_seg[sg] <= target) { run += s_seg[sg]; ++sg; }
            int J = min(N, (sg + 1) * PSEG) - 1;                         // rounding guard: last parent of the segment
            bool done = false;
            const int jend = min(N, (sg + 1) * PSEG);
            for (int j0 = sg * PSEG; j0 < jend && !done; j0 += 32) {
                const int j = j0 + lane;
                double w = 0.0;
                if (j < jend) { const Vec4T<R> pk = pkey[j]; w = (double)Mth<R>::exp(bb * pk.x + pk.y + pk.z + aa); }
                const double incl = warp_incl_scan(w);
                const unsigned hit = __ballot_sync(FULL, (j < jend) && (run + incl > target));
                if (hit) { J = j0 + __ffs(hit) - 1; done = true; }
                run += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) a.Jidx[(item_off + i) * Nt + jt] = J;
        }
        __syncthreads();
    }
}

// INJECTED mode: replay of accept_reject_based_backward_sampling (pf.py:260-341) / the naive branch
// (pf.py:226-237) with the recorded uniforms, one CTA per item.
template <class R, class Model>
__global__ void __launch_bounds__(NT) paris_injected_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    __shared__ R sh_r[NWARP];
    __shared__ double sh_d[NWARP];
    __shared__ int sh_i[NWARP];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (t >= a.T_buf[b]) return;
    const int N = a.N, par = t & 1, Nt = a.Ntilde;
    const size_t item_off = (size_t)b * N;
    const ItemHdr hdr = load_hdr(a, b);            // header of the OLD weights (built before step t)
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const R ltmax = Model::log_trans_max(th);
    const R* fine_old = reinterpret_cast<const R*>(a.fine[par]) + (size_t)b * a.Q * WT;
    const double* extra = a.inj_extra + a.inj_extra_off[(size_t)b * a.max_T + t];
    int32_t* Jb = a.Jidx + item_off * Nt;
    int64_t off = 0;
    if (!a.accept_reject) {
        for (int i = 0; i < N; ++i) {
            R rn[W];
            load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
            exact_backward_sample<R, Model>(a, th, par, item_off, rn + NP, Nt, extra + off, Jb + (size_t)i * Nt, 1, sh_r, sh_d);
            off += Nt;
        }
        return;
    }
    int32_t* Lcur = a.Llist[0] + item_off * Nt;
    int32_t* Lnext = a.Llist[1] + item_off;
    for (int jt = 0; jt < Nt; ++jt) {
        for (int k = tid; k < N; k += NT) Lcur[k] = k;
        __syncthreads();
        int size_L = N;
        bool converged = false;
        for (int round = 0; round < a.max_ar; ++round) {
            if (size_L == 0) { converged = true; break; }
            if (size_L <= a.manual_thresh) break;
            int new_count = 0;
            for (int base = 0; base < size_L; base += NT) {
                const int k = base + tid;
                int flag = 0, i = 0;
                if (k < size_L) {
                    i = Lcur[k];
                    const int I = search_hdr<R>(extra[off + k] * hdr.total, hdr, fine_old, N);
                    R ra[W], rn[W];
                    load_rec<R, W>(a.rec[par], a.tail[par], item_off + I, ra);
                    load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
                    const R thr = Mth<R>::exp(Model::log_trans(th, ra + NP, rn + NP) - ltmax);
                    if ((R)extra[off + size_L + k] <= thr) Jb[(size_t)i * Nt + jt] = I; else flag = 1;
                }
                int chunk_total;
                const int pos = block_excl_scan(flag, sh_i, chunk_total);
                if (flag) Lnext[new_count + pos] = i;
                new_count += chunk_total;
            }
            off += 2 * (int64_t)size_L;
            size_L = new_count;
            int32_t* tmp = Lcur; Lcur = Lnext; Lnext = tmp;
            __syncthreads();
        }
        if (!converged) {
            if (size_L > 0 && tid == 0 && size_L > a.manual_thresh) a.status[b] |= SGM_STATUS_AR_OVERFLOW;
            for (int e = 0; e < size_L; ++e) {
                const int i = Lcur[e];
                R rn[W];
                load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
                exact_backward_sample<R, Model>(a, th, par, item_off, rn + NP, 1, extra + off + e, Jb + (size_t)i * Nt + jt, 1, sh_r, sh_d);
            }
            off += size_L;
        }
        __syncthreads();
    }
}

// statistics update  tau'_i = mean_k( tau[J_ik] + h(x[J_ik], x'_i) * scale )   (pf.py:239-256)
template <class R, class Model>
__global__ void __launch_bounds__(NT) paris_update_kernel(KArgs a, int t) {
    constexpr int NX = Model::NX, NP = Model::NP, W = NX + NP;
    const int b = blockIdx.y, tid = threadIdx.x;
    if (t >= a.T_buf[b]) return;
    const int N = a.N, par = t & 1, Nt = a.Ntilde;
    const size_t item_off = (size_t)b * N;
    const int i = blockIdx.x * NT + tid;
    if (blockIdx.x == 0 && tid == 0) a.counters[b * 16] = 0;
    if (i >= N) return;
    const typename Model::template Theta<R> th = load_thc<R, Model>(a, b);
    const R y = (R)a.obs[a.obs_off[b] + t];
    const bool in_sub = (t >= a.t1[b]) && (t < a.tL[b]);
    const R wt = in_sub ? ((a.wts_off && a.wts_off[b] >= 0) ? (R)a.step_weights[a.wts_off[b] + (t - a.t1[b])] : (R)1) : (R)0;
    const int nws = stat_width<Model>(a.stat_kind);
    R rn[W];
    load_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
    R acc[4] = {(R)0, (R)0, (R)0, (R)0};
    for (int k = 0; k < Nt; ++k) {
        const int J = a.Jidx[(item_off + i) * Nt + k];
        if (a.trace_J) a.trace_J[(((size_t)b * a.max_T + t) * N + i) * Nt + k] = J;
        R ra[W], h[4];
        load_rec<R, W>(a.rec[par], a.tail[par], item_off + J, ra);
        stat_of<R, Model>(a, th, ra + NP, rn + NP, y, in_sub, h);
        for (int q = 0; q < NP; ++q) acc[q] += ra[q] + h[q] * wt;
    }
    for (int q = 0; q < NP; ++q) rn[q] = (q < nws) ? acc[q] / (R)Nt : (R)0;
    store_rec<R, W>(a.rec[par ^ 1], a.tail[par ^ 1], item_off + i, rn);
}

template <class R, class Model>
int launch_paris(const KArgs& a, int t, cudaStream_t stream) {
    int n = 0;
    if (a.rng_mode == SGM_RNG_INJECTED) {
        paris_injected_kernel<R, Model><<<a.B, NT, 0, stream>>>(a, t); ++n;
    } else {
        const int ch = (PARIS_CH * a.Ntilde <= PARIS_MAXQ) ? PARIS_CH : PARIS_MAXQ / a.Ntilde;
        paris_guide_kernel<R, Model><<<dim3(a.N / NT + 1, a.B), NT, 0, stream>>>(a, t); ++n;
        paris_ar_kernel<R, Model><<<dim3((a.N + ch - 1) / ch, a.B), NT, 0, stream>>>(a, t, ch); ++n;
        paris_exact_kernel<R, Model><<<dim3(EXACT_CTAS, a.B), NT, 0, stream>>>(a, t); ++n;
    }
    paris_update_kernel<R, Model><<<dim3((a.N + NT - 1) / NT, a.B), NT, 0, stream>>>(a, t); ++n;
    stat_ws_kernel<R, Model><<<dim3(a.G, a.B), NT, 0, stream>>>(a, t); ++n;
    return n;
}

}  // namespace sgm
