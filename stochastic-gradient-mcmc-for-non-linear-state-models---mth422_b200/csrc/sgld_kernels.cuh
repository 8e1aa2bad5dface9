// Device-resident SG-MCMC iteration around the batched particle filter: the two tiny kernels that bracket one
// sgm_pf_run-style launch sequence per iteration, so that K iterations of many chains run without a host round trip.
//
//   sgld_prepare_kernel  one thread per work item: sequence choice (Seq samplers), window draw + per-step importance
//                        weights (sgmcmc_sampler.py:1969-2017 random_subsequence_and_weights, :259-288 buffers), model
//                        scalars theta from the chain's current parameters (what the Parameters object computes:
//                        Qinv = LQinv^2 + 1e-16, variables/covariance.py:141-146; GARCH alpha / beta / gamma,
//                        variables/garch_var.py:69-91), prior of x_0
//   sgld_update_kernel   one CTA: per chain the minibatch mean of the item gradients (sgmcmc_sampler.py:411-418; Seq:
//                        :1264-1282 sum over sequences, T / S rescale), grad log-prior (covariance.py:229-243,
//                        matrices.py:575-590, garch_var.py:150-163), the SGLD / SGRLD / SGD step
//                        (sgmcmc_sampler.py:549-567, :613-640, :466-480) and project_parameters (:650-656 with the
//                        models' defaults: |A| <= 0.9999, positive Cholesky diagonals, LGSSM C = 1)
//
// All of it is scalar float64 algebra (n = m = 1 models).  Parameter slots follow var_dict order:
//   SVM   [A, LQinv, LRinv]            hyper [mean_A, var_col_A, df_Q, scale_Q, df_R, scale_R]
//   LGSSM [A, C, LQinv, LRinv]         hyper [mean_A, var_col_A, mean_C, var_col_C, df_Q, scale_Q, df_R, scale_R]
//   GARCH [log_mu, logit_phi, logit_lambduh, LRinv]
//                                      hyper [scale_mu, shape_mu, alpha_phi, beta_phi, alpha_lam, beta_lam, df_R, scale_R]
#pragma once
#include "rng.cuh"
#include "../../include/sgmpf.h"

namespace sgm {

enum : uint32_t { STREAM_SGLD_WINDOW = 6, STREAM_SGLD_NOISE = 7, STREAM_SGLD_SEQ = 8 };

struct SgldArgs {
    int model, method, C, M, nsel, n_seqs, S, Bf, partition, project, prior_x0, injected;
    int ipc;                    // items per chain = nsel * M
    int Smax;                   // weights row stride
    int pick_all;               // num_sequences == -1
    double epsilon, T_total;
    RngKey key;                 // .item = global index of chain 0
    const uint64_t* offset_dev;
    uint64_t* offset_rw;        // same location, advanced by the update kernel
    const double* obs; const int64_t* seq_off;
    double* params; const double* hyper; const double* prior_mean; const double* prior_var;
    int32_t* chain_status;
    double* trace; int trace_every, trace_rows;
    int64_t* iter_dev;          // iterations completed by this descriptor's chains (trace row index)
    const int32_t* inj_start; const int32_t* inj_seq; const double* inj_noise;
    // per-item arrays handed to the particle filter
    int64_t* obs_off; int32_t* T_buf; int32_t* t1; int32_t* tL; int64_t* wts_off; double* weights; double* theta;
    double* item_pm; double* item_pv; int32_t* item_seq;
    const double* grad; const double* loglik; const int32_t* item_status;
};

__device__ __forceinline__ RngKey sgld_key(const SgldArgs& a, int chain, uint64_t o) {
    RngKey k = a.key;
    k.offset = (uint32_t)(o & 0xffffffffu);
    k.k1 ^= (uint32_t)(o >> 32);
    k.item += (uint32_t)chain;
    return k;
}
__device__ __forceinline__ double sgld_uniform(const RngKey& k, uint32_t index, uint32_t stream, uint32_t sub) {
    const uint4 r = rng_raw(k, index, 0u, stream, sub);
    return u01d(r.x, r.y);
}

// theta (SGM_THETA_STRIDE doubles, layout of include/sgmpf.h) from the chain's parameter slots
__device__ inline void sgld_theta(int model, const double* p, double* th) {
    for (int q = 0; q < SGM_THETA_STRIDE; ++q) th[q] = 0.0;
    if (model == SGM_MODEL_SVM) {
        const double Qinv = p[1] * p[1] + 1e-16, Rinv = p[2] * p[2] + 1e-16;
        th[0] = p[0]; th[1] = p[1]; th[2] = Qinv; th[3] = p[2]; th[4] = Rinv; th[10] = 1.0 / Qinv; th[11] = 1.0 / Rinv;
    } else if (model == SGM_MODEL_LGSSM) {
        const double Qinv = p[2] * p[2] + 1e-16, Rinv = p[3] * p[3] + 1e-16;
        th[0] = p[0]; th[1] = p[2]; th[2] = Qinv; th[3] = p[1]; th[4] = p[3]; th[5] = Rinv; th[10] = 1.0 / Qinv; th[11] = 1.0 / Rinv;
    } else {
        const double mu = ::exp(p[0]), phi = 1.0 / (1.0 + ::exp(-p[1])), lam = 1.0 / (1.0 + ::exp(-p[2]));
        const double Rinv = p[3] * p[3] + 1e-16;
        th[0] = mu * (1.0 - phi); th[1] = phi * lam; th[2] = phi * (1.0 - lam); th[3] = mu; th[4] = phi; th[5] = lam;
        th[6] = p[3]; th[7] = Rinv; th[8] = 1.0 / Rinv;
    }
}

// sequences of chain c for this iteration: position s of the draw (np.random.choice(idx, num_sequences, replace=False),
// sgmcmc_sampler.py:1261-1263; here: successive uniform picks, repeats rejected -- the same law)
__device__ inline int sgld_pick_sequence(const SgldArgs& a, const RngKey& k, int c, int s, int64_t it) {
    if (a.pick_all) return s;
    if (a.injected) return a.inj_seq[((size_t)it * a.C + c) * a.nsel + s];
    int chosen[8];
    for (int i = 0; i <= s; ++i) {
        for (uint32_t attempt = 0;; ++attempt) {
            const int q = min(a.n_seqs - 1, (int)(sgld_uniform(k, (uint32_t)i, STREAM_SGLD_SEQ, attempt) * a.n_seqs));
            bool dup = false;
            for (int j = 0; j < i; ++j) dup = dup || (chosen[j] == q);
            if (!dup || attempt > 4096u) { chosen[i] = q; break; }
        }
    }
    return chosen[s];
}

// work item b of iteration k_call (Philox call offset `off`): window, weights, theta, x_0 prior -> the item arrays
// (tid, nth): the nth threads that call this together for the SAME item (persistent kernels: the whole CTA) share the stores --
// the S window weights are strided over them, thread 0 writes the scalars; every thread recomputes the few scalars itself
__device__ inline void sgld_prepare_item(const SgldArgs& a, int b, int k_call, uint64_t off, int tid = 0, int nth = 1) {
    const int B = a.C * a.ipc;
    const int c = b / a.ipc, j = b % a.ipc, s = j / a.M;
    const RngKey key = sgld_key(a, c, off);
    const int q = sgld_pick_sequence(a, key, c, s, k_call);
    const int64_t base = a.seq_off[q];
    const int Tq = (int)(a.seq_off[q + 1] - base);
    const int S = a.S, Bf = (a.Bf < 0) ? Tq : a.Bf;
    int start = 0, end = Tq;
    bool weighted = false;
    if (S != -1 && Tq - S > 0) {                                  // sgmcmc_sampler.py:270-276
        weighted = true;
        const int nstart = (a.partition == SGM_PARTITION_STRICT) ? Tq / S : Tq - S + 1;
        int r;
        if (a.injected) r = a.inj_start[(size_t)k_call * B + b];
        else r = min(nstart - 1, (int)(sgld_uniform(key, (uint32_t)j, STREAM_SGLD_WINDOW, 0u) * nstart));
        start = (a.partition == SGM_PARTITION_STRICT) ? r * S : r;
        end = start + S;
        double* w = a.weights + (size_t)b * a.Smax;
        for (int i = tid; i < S; i += nth) {
            double wt;
            if (a.partition == SGM_PARTITION_UNIFORM) {           // :1994-2008
                const int t = start + i, cap = min(S, Tq - S + 1);
                double num;
                if (end <= 2 * S) num = (double)min(t + 1, cap);
                else if (start >= Tq - 2 * S - 1) num = (double)min(Tq - t, cap);
                else num = (double)S;
                wt = 1.0 * (double)(Tq - S + 1) / num;
            } else {
                wt = 1.0 * (double)Tq / (double)S;                // 'strict' / 'naive'
            }
            w[i] = wt;
        }
    }
    const int left = max(0, start - Bf), right = min(Tq, end + Bf);
    const double* p = a.params + (size_t)c * SGM_PARAM_STRIDE;
    double th[SGM_THETA_STRIDE];
    sgld_theta(a.model, p, th);
    for (int i = tid; i < SGM_THETA_STRIDE; i += nth) a.theta[(size_t)b * SGM_THETA_STRIDE + i] = th[i];
    if (tid != 0) return;
    a.obs_off[b] = base + left;
    a.T_buf[b] = right - left;
    a.t1[b] = start - left;
    a.tL[b] = end - left;
    a.wts_off[b] = weighted ? (int64_t)b * a.Smax : -1;
    a.item_seq[b] = q;
    if (a.prior_x0 == 1) {                                        // garch/helper.py:324-332
        a.item_pm[b] = 0.0;
        a.item_pv[b] = th[0] / (1.0 - th[1] - th[2]);
    } else {
        a.item_pm[b] = a.prior_mean[c];
        a.item_pv[b] = a.prior_var[c];
    }
}

static __global__ void __launch_bounds__(128) sgld_prepare_kernel(SgldArgs a, int k_call) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    if (b >= a.C * a.ipc) return;
    sgld_prepare_item(a, b, k_call, *a.offset_dev);
}

__device__ __forceinline__ double sgld_reflect(double L) {        // covariance.py:67-79 (1 x 1 Cholesky of L L^T + 1e-16)
    return (L < 0.0) ? ::sqrt(L * L + 1e-16) : L;
}
__device__ __forceinline__ double sgld_thresh(double A) {         // _utils.py:149-172 (1 x 1), cutoff 0.9999
    const double rho = fabs(A);
    return (rho > 0.9999) ? A * (0.9999 / rho) : A;
}

// chain c: gradient of the minibatch -> parameter update of iteration k_call (Philox call offset s_off, s_iter
// iterations completed before this one)
__device__ inline void sgld_update_chain(const SgldArgs& a, int c, int k_call, uint64_t s_off, int64_t s_iter) {
    const int NPAR = (a.model == SGM_MODEL_SVM) ? 3 : 4;
    {
        const RngKey key = sgld_key(a, c, s_off);
        double* p = a.params + (size_t)c * SGM_PARAM_STRIDE;
        const double* h = a.hyper + (size_t)c * SGM_HYPER_STRIDE;
        // ---- likelihood part: sum over the chain's sequences of the minibatch means (+ T / S rescale) ----
        double col[4] = {0.0, 0.0, 0.0, 0.0};
        double S_sel = 0.0;
        int bad = 0;
        for (int s = 0; s < a.nsel; ++s) {
            const int b0 = c * a.ipc + s * a.M;
            const int q = a.item_seq[b0];
            S_sel += (double)(a.seq_off[q + 1] - a.seq_off[q]);
            for (int m = 0; m < a.M; ++m) {
                const int b = b0 + m;
                bad |= a.item_status[b] & (SGM_STATUS_NAN_WEIGHT | SGM_STATUS_ZERO_WEIGHT);
                for (int i = 0; i < NPAR; ++i) col[i] += a.grad[(size_t)b * 8 + i] * 1.0 / (double)a.M;
            }
        }
        if (!a.pick_all) { const double sc = a.T_total / S_sel; for (int i = 0; i < NPAR; ++i) col[i] *= sc; }
        double gl[4], gp[4], z[4];
        // ---- map gradient columns to parameter slots, grad log-prior ----
        if (a.model == SGM_MODEL_SVM) {                           // columns [LRinv, LQinv, A]
            gl[0] = col[2]; gl[1] = col[1]; gl[2] = col[0];
            const double Qinv = p[1] * p[1] + 1e-16;
            gp[0] = -1.0 * (Qinv * (p[0] - h[0])) * (1.0 / h[1]);
            gp[1] = (h[2] - 2.0) * (1.0 / p[1]) - p[1] / h[3];
            gp[2] = (h[4] - 2.0) * (1.0 / p[2]) - p[2] / h[5];
        } else if (a.model == SGM_MODEL_LGSSM) {                  // columns [LRinv, LQinv, C, A]
            gl[0] = col[3]; gl[1] = col[2]; gl[2] = col[1]; gl[3] = col[0];
            const double Qinv = p[2] * p[2] + 1e-16, Rinv = p[3] * p[3] + 1e-16;
            gp[0] = -1.0 * (Qinv * (p[0] - h[0])) * (1.0 / h[1]);
            gp[1] = -1.0 * (Rinv * (p[1] - h[2])) * (1.0 / h[3]);
            gp[2] = (h[4] - 2.0) * (1.0 / p[2]) - p[2] / h[5];
            gp[3] = (h[6] - 2.0) * (1.0 / p[3]) - p[3] / h[7];
        } else {                                                  // columns [LRinv, log_mu, logit_phi, logit_lambduh]
            gl[0] = col[1]; gl[1] = col[2]; gl[2] = col[3]; gl[3] = col[0];
            const double mu = ::exp(p[0]), phi = 1.0 / (1.0 + ::exp(-p[1])), lam = 1.0 / (1.0 + ::exp(-p[2]));
            gp[0] = -h[1] - 1.0 + h[0] / mu;
            gp[1] = ((h[2] - 1.0) / (1.0 + phi) - (h[3] - 1.0) / (1.0 - phi)) * phi * (1.0 - phi);
            gp[2] = ((h[4] - 1.0) / (1.0 + lam) - (h[5] - 1.0) / (1.0 - lam)) * lam * (1.0 - lam);
            gp[3] = (h[6] - 2.0) * (1.0 / p[3]) - p[3] / h[7];
        }
        for (int i = 0; i < NPAR; ++i) {
            if (a.method == SGM_STEP_SGD) z[i] = 0.0;
            else if (a.injected) z[i] = a.inj_noise[((size_t)k_call * a.C + c) * SGM_PARAM_STRIDE + i];
            else rng_normal1(key, (uint32_t)i, 0u, STREAM_SGLD_NOISE, 0u, z[i]);
        }
        const double scale = 1.0 / a.T_total, eps = a.epsilon, s2e = ::sqrt(2.0 * eps);
        double pn[4];
        if (a.method == SGM_STEP_SGRLD && a.model == SGM_MODEL_LGSSM) {
            // LGSSMPreconditioner (lgssm/parameters.py:58-67; matrices.py:632-656, 1094-1125; covariance.py:286-317), 1 x 1
            const double LQ = p[2], LR = p[3], Qinv = LQ * LQ + 1e-16, Rinv = LR * LR + 1e-16;
            const double ng[4] = {gp[0] + gl[0], gp[1] + gl[1], gp[2] + gl[2], gp[3] + gl[3]};
            const double delta[4] = {(1.0 / Qinv) * ng[0] * scale, (1.0 / Rinv) * ng[1] * scale,
                                     0.5 * Qinv * ng[2] * scale, 0.5 * Rinv * ng[3] * scale};
            const double rs = ::sqrt(scale);
            const double noise[4] = {z[0] / LQ * rs, z[1] / LR * rs, ::sqrt(0.5) * LQ * z[2] * rs, ::sqrt(0.5) * LR * z[3] * rs};
            const double corr[4] = {0.0, 0.0, LQ * scale, LR * scale};
            for (int i = 0; i < 4; ++i) pn[i] = p[i] + (eps * (delta[i] + corr[i]) + s2e * noise[i]);
        } else {
            const double sd = ::sqrt(scale);
            for (int i = 0; i < NPAR; ++i) {
                const double g = (gp[i] + gl[i]) / a.T_total;
                pn[i] = (a.method == SGM_STEP_SGD) ? p[i] + eps * g : p[i] + (eps * g + s2e * (0.0 + sd * z[i]));
            }
        }
        for (int i = 0; i < NPAR; ++i) bad |= (pn[i] != pn[i]) ? SGM_STATUS_NAN_WEIGHT : 0;
        if (a.project) {
            if (a.model == SGM_MODEL_SVM) { pn[0] = sgld_thresh(pn[0]); pn[1] = sgld_reflect(pn[1]); pn[2] = sgld_reflect(pn[2]); }
            else if (a.model == SGM_MODEL_LGSSM) { pn[0] = sgld_thresh(pn[0]); pn[1] = 1.0; pn[2] = sgld_reflect(pn[2]); pn[3] = sgld_reflect(pn[3]); }
            else pn[3] = sgld_reflect(pn[3]);
        }
        if (a.model == SGM_MODEL_SVM && fabs(pn[0]) > 1.0) bad |= SGM_STATUS_NAN_WEIGHT;     // svm/kernels.py:8-10
        // a chain whose gradient went non-finite keeps its last good parameters and is flagged
        if (!bad) for (int i = 0; i < NPAR; ++i) p[i] = pn[i];
        if (bad) a.chain_status[c] |= bad;
        if (a.trace && a.trace_every > 0) {
            const int64_t done = s_iter + 1;
            if (done % a.trace_every == 0) {
                const int64_t row = done / a.trace_every;
                if (row < a.trace_rows) for (int i = 0; i < SGM_PARAM_STRIDE; ++i)
                    a.trace[((size_t)row * a.C + c) * SGM_PARAM_STRIDE + i] = (i < NPAR) ? p[i] : 0.0;
            }
        }
    }
}

static __global__ void __launch_bounds__(256) sgld_update_kernel(SgldArgs a, int k_call) {
    __shared__ uint64_t s_off;
    __shared__ int64_t s_iter;
    if (threadIdx.x == 0) { s_off = *a.offset_dev; s_iter = *a.iter_dev; }
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += 256) sgld_update_chain(a, c, k_call, s_off, s_iter);
    __syncthreads();
    if (threadIdx.x == 0) { *a.offset_rw = s_off + 1; *a.iter_dev = s_iter + 1; }
}

// after a persistent launch that ran K iterations inside one kernel
static __global__ void sgld_advance_kernel(SgldArgs a, int K) {
    *a.offset_rw = *a.offset_dev + (uint64_t)K;
    *a.iter_dev = *a.iter_dev + K;
}

}  // namespace sgm
