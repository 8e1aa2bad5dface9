// Per-model device functors: proposal (rv), incremental log-weight (reweight), transition log-density,
// complete-data score and sufficient statistics.  Each cites the reference lines it restates
// (paths relative to /root/reference/sgmcmc_ssm/).  Everything is templated on the arithmetic type R.
#pragma once
#include <cuda_runtime.h>
#include "../../include/sgmpf.h"
#include "fastlog.cuh"

namespace sgm {

template <class R> struct Mth;
template <> struct Mth<float> {
    // single MUFU instructions (flush-to-zero, ~2 ulp): the f32 path is the throughput path; the f64
    // instantiation below is the parity path
    static __device__ __forceinline__ float exp(float x) {
        float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(x, 1.4426950408889634f))); return r;
    }
    static __device__ __forceinline__ float exp2(float x) {
        float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
    }
    static __device__ __forceinline__ float log(float x) {
        float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return __fmul_rn(r, 0.6931471805599453f);
    }
    static __device__ __forceinline__ float log2(float x) {
        float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
    }
    static __device__ __forceinline__ float logb(float x) { return log2(x); }
    static __device__ __forceinline__ float sqrt(float x) {
        float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
    }
    static __device__ __forceinline__ float rcp(float x) {
        float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
    }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    // explicitly rounded primitives: never contracted or re-associated by the compiler, so the particle system
    // (and with it the genealogy) is bit-identical across the instantiations of a kernel
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct Mth<double> {
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double exp2(double x) { return ::exp2(x); }
    static __device__ __forceinline__ double log2(double x) { return ::log2(x); }
    // any fixed-base logarithm serves the exponential spacings (they are normalised by their sum): the table-driven natural
    // one (fastlog.cuh; ::log2 was 21 % of the f64 step kernel's instructions, profiles/ncu_r02_f64_step_kernel_source_lines.txt)
    static __device__ __forceinline__ double logb(double x) { return fast_log(x); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};

constexpr double LOG_2PI_D = 1.8378770664093453;

// ------------------------------------------------------------------------------------------------
// Latent-Gaussian AR(1) state shared by LGSSM and SVM (particle_filters/kernels.py:82-138, n = 1)
// ------------------------------------------------------------------------------------------------
template <class R> struct GaussTheta {
    R A, LQinv, Qinv, C, LRinv, Rinv;
    R invLQ, invLR, logLQinv, logLRinv;     // LQinv**-1, LRinv**-1, log LQinv, log LRinv
    R opt_sd, opt_iprec, opt_ivar, opt_lvar; // LGSSM optimal kernel: prec**-0.5, 1/prec, 1/(1/Qinv+1/Rinv), log(1/Qinv+1/Rinv)
    R Qv, Rv;                                // parameters.Q, parameters.R (predictive statistic only)
};

struct SvmPrior {
    static constexpr int NX = 1, NP = 3, MODEL = SGM_MODEL_SVM;
    template <class R> using Theta = GaussTheta<R>;
    template <class R> static __device__ __forceinline__ Theta<R> load(const double* th) {
        Theta<R> t;   // theta: A, LQinv, Qinv, LRinv, Rinv
        t.A = (R)th[0]; t.LQinv = (R)th[1]; t.Qinv = (R)th[2]; t.LRinv = (R)th[3]; t.Rinv = (R)th[4]; t.C = (R)1;
        t.invLQ = (R)(1.0 / th[1]); t.invLR = (R)(1.0 / th[3]); t.logLQinv = (R)::log(th[1]); t.logLRinv = (R)::log(th[3]);
        t.opt_sd = t.opt_iprec = t.opt_ivar = t.opt_lvar = (R)0;
        t.Qv = (R)th[10]; t.Rv = (R)th[11];
        return t;
    }
    // svm/kernels.py:34-37
    template <class R> static __device__ __forceinline__ void propagate(const Theta<R>& t, const R* xa, R y, R z, R* xn) {
        xn[0] = Mth<R>::fma(t.invLQ, z, Mth<R>::mul(xa[0], t.A));
    }
    // svm/kernels.py:57-62
    template <class R> static __device__ __forceinline__ R log_weight(const Theta<R>& t, const R* xa, const R* xn, R y) {
        // terms that do not depend on the particle first, so that they are shared by the 8 particles of a lane
        const R k = Mth<R>::mul(Mth<R>::mul((R)-0.5, t.Rinv), Mth<R>::mul(y, y));
        const R c0 = Mth<R>::add((R)(-0.5 * LOG_2PI_D), t.logLRinv);
        return Mth<R>::fma(k, Mth<R>::exp(-xn[0]), Mth<R>::fma((R)-0.5, xn[0], c0));
    }
    // svm/helper.py:342-348 ; order [dLRinv, dLQinv, dA]
    template <class R> static __device__ __forceinline__ void score(const Theta<R>& t, const R* xa, const R* xn, R y, R* h) {
        const R d = xn[0] - t.A * xa[0];
        h[2] = t.Qinv * d * xa[0];
        h[1] = t.invLQ - (d * d) * t.LQinv;
        h[0] = t.invLR - ((y * y) * t.LRinv) * Mth<R>::exp(-xn[0]);       // 1/LRinv - (y^2 / exp(x')) LRinv
    }
    // particle_filters/kernels.py:121-126
    template <class R> static __device__ __forceinline__ R log_trans(const Theta<R>& t, const R* xa, const R* xn) {
        const R d = xn[0] - t.A * xa[0];
        return (R)-0.5 * (d * d) * t.Qinv + (R)(-0.5 * LOG_2PI_D) + t.logLQinv;
    }
    // particle_filters/kernels.py:134-138
    template <class R> static __device__ __forceinline__ R log_trans_max(const Theta<R>& t) {
        return (R)(-0.5 * LOG_2PI_D) + t.logLQinv;
    }
    // per-parent part of log_trans, hoisted out of the O(N^2) pair loop
    template <class R> static __device__ __forceinline__ void jkey(const Theta<R>& t, const R* xa, R* k) {
        k[0] = t.A * xa[0]; k[1] = (R)(-0.5 * LOG_2PI_D) + t.logLQinv;
    }
    template <class R> static __device__ __forceinline__ R log_trans_key(const Theta<R>& t, const R* k, const R* xn) {
        const R d = xn[0] - k[0];
        return (R)-0.5 * (d * d) * t.Qinv + k[1];
    }
    // lgssm/helper.py:1338-1363 (n = 1): [x', x'^2, x x']
    template <class R> static __device__ __forceinline__ void suff(const R* xa, const R* xn, R* h) {
        h[0] = xn[0]; h[1] = xn[0] * xn[0]; h[2] = xa[0] * xn[0];
    }
    template <class R> static __device__ __forceinline__ void init(R mean, R sd, R z, R* x) { x[0] = mean + sd * z; }

    // ---- O(N^2) backward weights as a rank-1 score (pf.py:115-135) --------------------
    // log q(x'_i | x_j) - log_trans_max = b_i u_j + gq_j + a_i :  -0.5 Qinv (x' - A x)^2 expanded.  The statistic
    // h(x_j, x'_i) is a polynomial in x_j, so sum_j p_ij h needs only the p-weighted moments of m_j = (x_j, x_j^2).
    template <class R> static __device__ __forceinline__ void pair_child(const Theta<R>& t, const R* xn, R& b, R& a) {
        b = t.Qinv * xn[0]; a = (R)-0.5 * t.Qinv * (xn[0] * xn[0]);
    }
    template <class R> static __device__ __forceinline__ void pair_parent(const Theta<R>& t, const R* xa, R& u, R& gq, R* m) {
        u = t.A * xa[0]; gq = (R)-0.5 * t.Qinv * (u * u);
        m[0] = xa[0]; m[1] = xa[0] * xa[0];
    }
    // Per-child upper bound of log q(x'_i | x_j) - log_trans_max over ALL admissible parents (PaRIS accept-reject,
    // device-random mode): the tighter the bound, the higher the acceptance rate; the sampler stays exact for any
    // bound that dominates.  Gaussian AR(1): parents can sit anywhere, the global maximum is the bound.
    template <class R> static __device__ __forceinline__ R pair_bound(const Theta<R>& t, const R* xn) { return (R)0; }
    // E_j[h(x_j, x'_i)] from E[m]   (svm/helper.py:342-348 with x_j -> its moments)
    template <class R> static __device__ __forceinline__ void score_moments(const Theta<R>& t, const R* Em, const R* xn, R y, R* h) {
        h[2] = t.Qinv * (xn[0] * Em[0] - t.A * Em[1]);
        h[1] = t.invLQ - t.LQinv * (xn[0] * xn[0] - (R)2 * t.A * xn[0] * Em[0] + t.A * t.A * Em[1]);
        h[0] = t.invLR - ((y * y) * t.LRinv) * Mth<R>::exp(-xn[0]);
        h[3] = (R)0;
    }
    template <class R> static __device__ __forceinline__ void suff_moments(const R* Em, const R* xn, R* h) {
        h[0] = xn[0]; h[1] = xn[0] * xn[0]; h[2] = Em[0] * xn[0]; h[3] = (R)0;
    }
    // ---- k-step-ahead predictive log-likelihood (svm/helper.py:352-395, Ntilde = 1) ----------------------------
    // predictive state ps = (mean, cov) of x_{t+k} given the particle; one normal per horizon (also at k = 0,
    // where it multiplies sqrt(0)); y ~ N(0, R exp(x_mc))
    static constexpr bool PRED_RNG = true;
    template <class R> static __device__ __forceinline__ void pred_begin(const R* xn, R* ps) { ps[0] = xn[0]; ps[1] = (R)0; }
    template <class R> static __device__ __forceinline__ R pred_ll(const Theta<R>& t, const R* ps, R yk, R z) {
        const R x_mc = ps[0] + Mth<R>::sqrt(ps[1]) * z;
        const R y_cov = t.Rv * Mth<R>::exp(x_mc);
        return (R)-0.5 * (yk * yk) / y_cov + (R)(-0.5 * LOG_2PI_D) - (R)0.5 * Mth<R>::log(y_cov);
    }
    template <class R> static __device__ __forceinline__ void pred_next(const Theta<R>& t, R* ps, R z) {
        ps[0] = t.A * ps[0]; ps[1] = t.Qv + (t.A * t.A) * ps[1];
    }
};

struct LgssmPrior {
    static constexpr int NX = 1, NP = 4, MODEL = SGM_MODEL_LGSSM;
    template <class R> using Theta = GaussTheta<R>;
    template <class R> static __device__ __forceinline__ Theta<R> load(const double* th) {
        Theta<R> t;   // theta: A, LQinv, Qinv, C, LRinv, Rinv
        t.A = (R)th[0]; t.LQinv = (R)th[1]; t.Qinv = (R)th[2]; t.C = (R)th[3]; t.LRinv = (R)th[4]; t.Rinv = (R)th[5];
        t.invLQ = (R)(1.0 / th[1]); t.invLR = (R)(1.0 / th[4]); t.logLQinv = (R)::log(th[1]); t.logLRinv = (R)::log(th[4]);
        const double prec = th[2] + th[3] * th[3] * th[5];          // lgssm/kernels.py:91-93
        const double var = 1.0 / th[2] + 1.0 / th[5];                // lgssm/kernels.py:118
        t.opt_sd = (R)(1.0 / ::sqrt(prec)); t.opt_iprec = (R)(1.0 / prec); t.opt_ivar = (R)(1.0 / var); t.opt_lvar = (R)::log(var);
        t.Qv = (R)th[10]; t.Rv = (R)th[11];
        return t;
    }
    // lgssm/kernels.py:29-33
    template <class R> static __device__ __forceinline__ void propagate(const Theta<R>& t, const R* xa, R y, R z, R* xn) {
        xn[0] = Mth<R>::fma(t.invLQ, z, Mth<R>::mul(xa[0], t.A));
    }
    // lgssm/kernels.py:58-62
    template <class R> static __device__ __forceinline__ R log_weight(const Theta<R>& t, const R* xa, const R* xn, R y) {
        const R d = Mth<R>::fma(-t.C, xn[0], y);
        return Mth<R>::add(Mth<R>::fma(Mth<R>::mul((R)-0.5, Mth<R>::mul(d, d)), t.Rinv, (R)(-0.5 * LOG_2PI_D)), t.logLRinv);
    }
    // lgssm/helper.py:1270-1277 ; order [dLRinv, dLQinv, dC, dA]
    template <class R> static __device__ __forceinline__ void score(const Theta<R>& t, const R* xa, const R* xn, R y, R* h) {
        const R d = xn[0] - t.A * xa[0];
        h[3] = t.Qinv * d * xa[0];
        h[1] = t.invLQ - (d * d) * t.LQinv;
        const R dy = y - t.C * xn[0];
        h[2] = t.Rinv * dy * xn[0];
        h[0] = t.invLR - (dy * dy) * t.LRinv;
    }
    template <class R> static __device__ __forceinline__ R log_trans(const Theta<R>& t, const R* xa, const R* xn) {
        return SvmPrior::log_trans(t, xa, xn);
    }
    template <class R> static __device__ __forceinline__ R log_trans_max(const Theta<R>& t) { return SvmPrior::log_trans_max(t); }
    template <class R> static __device__ __forceinline__ void jkey(const Theta<R>& t, const R* xa, R* k) { SvmPrior::jkey(t, xa, k); }
    template <class R> static __device__ __forceinline__ R log_trans_key(const Theta<R>& t, const R* k, const R* xn) {
        return SvmPrior::log_trans_key(t, k, xn);
    }
    template <class R> static __device__ __forceinline__ void suff(const R* xa, const R* xn, R* h) { SvmPrior::suff(xa, xn, h); }
    template <class R> static __device__ __forceinline__ void init(R mean, R sd, R z, R* x) { x[0] = mean + sd * z; }
    template <class R> static __device__ __forceinline__ void pair_child(const Theta<R>& t, const R* xn, R& b, R& a) { SvmPrior::pair_child(t, xn, b, a); }
    template <class R> static __device__ __forceinline__ void pair_parent(const Theta<R>& t, const R* xa, R& u, R& gq, R* m) { SvmPrior::pair_parent(t, xa, u, gq, m); }
    template <class R> static __device__ __forceinline__ R pair_bound(const Theta<R>& t, const R* xn) { return (R)0; }
    // lgssm/helper.py:1270-1277 with x_j -> its moments
    template <class R> static __device__ __forceinline__ void score_moments(const Theta<R>& t, const R* Em, const R* xn, R y, R* h) {
        h[3] = t.Qinv * (xn[0] * Em[0] - t.A * Em[1]);
        h[1] = t.invLQ - t.LQinv * (xn[0] * xn[0] - (R)2 * t.A * xn[0] * Em[0] + t.A * t.A * Em[1]);
        const R dy = y - t.C * xn[0];
        h[2] = t.Rinv * dy * xn[0];
        h[0] = t.invLR - (dy * dy) * t.LRinv;
    }
    template <class R> static __device__ __forceinline__ void suff_moments(const R* Em, const R* xn, R* h) { SvmPrior::suff_moments(Em, xn, h); }
    // lgssm/helper.py:1281-1336 (m = n = 1): analytic, no random numbers
    static constexpr bool PRED_RNG = false;
    template <class R> static __device__ __forceinline__ void pred_begin(const R* xn, R* ps) { ps[0] = xn[0]; ps[1] = (R)0; }
    template <class R> static __device__ __forceinline__ R pred_ll(const Theta<R>& t, const R* ps, R yk, R z) {
        const R diff = yk - ps[0] * t.C;
        const R y_cov = t.Rv + t.C * ps[1] * t.C;
        return (R)-0.5 * (diff * diff) / y_cov + (R)(-0.5 * LOG_2PI_D) - (R)0.5 * Mth<R>::log(y_cov);
    }
    template <class R> static __device__ __forceinline__ void pred_next(const Theta<R>& t, R* ps, R z) {
        ps[0] = t.A * ps[0]; ps[1] = t.Qv + (t.A * t.A) * ps[1];
    }
};

struct LgssmOptimal : LgssmPrior {
    // lgssm/kernels.py:87-97
    template <class R> static __device__ __forceinline__ void propagate(const Theta<R>& t, const R* xa, R y, R z, R* xn) {
        const R mp = Mth<R>::fma(Mth<R>::mul(xa[0], t.A), t.Qinv, Mth<R>::mul(Mth<R>::mul(y, t.C), t.Rinv));
        xn[0] = Mth<R>::fma(t.opt_sd, z, Mth<R>::mul(mp, t.opt_iprec));
    }
    // lgssm/kernels.py:117-120 (ignores C, as the reference does)
    template <class R> static __device__ __forceinline__ R log_weight(const Theta<R>& t, const R* xa, const R* xn, R y) {
        const R d = Mth<R>::fma(-t.A, xa[0], y);
        const R c0 = Mth<R>::fma((R)-0.5, t.opt_lvar, (R)(-0.5 * LOG_2PI_D));
        return Mth<R>::fma(Mth<R>::mul((R)-0.5, Mth<R>::mul(d, d)), t.opt_ivar, c0);
    }
};

// ------------------------------------------------------------------------------------------------
// GARCH(1,1) + noise; particle = (x, sigma2)  (models/garch/kernels.py)
// ------------------------------------------------------------------------------------------------
template <class R> struct GarchTheta {
    R alpha, beta, gamma, mu, phi, lam, LRinv, Rinv, Rv, invLR, logLRinv, ltmax;
};

struct GarchPrior {
    static constexpr int NX = 2, NP = 4, MODEL = SGM_MODEL_GARCH;
    template <class R> using Theta = GarchTheta<R>;
    template <class R> static __device__ __forceinline__ Theta<R> load(const double* th) {
        Theta<R> t;   // theta: alpha, beta, gamma, mu, phi, lambduh, LRinv, Rinv, R
        t.alpha = (R)th[0]; t.beta = (R)th[1]; t.gamma = (R)th[2]; t.mu = (R)th[3]; t.phi = (R)th[4]; t.lam = (R)th[5];
        t.LRinv = (R)th[6]; t.Rinv = (R)th[7]; t.Rv = (R)th[8];
        t.invLR = (R)(1.0 / th[6]); t.logLRinv = (R)::log(th[6]);
        t.ltmax = (R)(-0.5 * LOG_2PI_D - 0.5 * ::log(th[0]));       // garch/kernels.py:41-43
        return t;
    }
    template <class R> static __device__ __forceinline__ R sigma2_next(const Theta<R>& t, const R* xa) {
        return Mth<R>::fma(t.gamma, xa[1], Mth<R>::fma(t.beta, Mth<R>::mul(xa[0], xa[0]), t.alpha));
    }
    // garch/kernels.py:60-68
    template <class R> static __device__ __forceinline__ void propagate(const Theta<R>& t, const R* xa, R y, R z, R* xn) {
        const R s2 = sigma2_next(t, xa);
        xn[0] = Mth<R>::mul(Mth<R>::sqrt(s2), z);
        xn[1] = s2;
    }
    // garch/kernels.py:82-89
    template <class R> static __device__ __forceinline__ R log_weight(const Theta<R>& t, const R* xa, const R* xn, R y) {
        const R d = Mth<R>::add(y, -xn[0]);
        return Mth<R>::add(Mth<R>::fma(Mth<R>::mul((R)-0.5, Mth<R>::mul(d, d)), t.Rinv, (R)(-0.5 * LOG_2PI_D)), t.logLRinv);
    }
    // garch/helper.py:350-372 ; order [dLRinv, dlog_mu, dlogit_phi, dlogit_lambduh]
    template <class R> static __device__ __forceinline__ void score(const Theta<R>& t, const R* xa, const R* xn, R y, R* h) {
        const R v = xn[1];
        const R iv = Mth<R>::rcp(v);
        const R gv = (R)-0.5 * (v - xn[0] * xn[0]) * (iv * iv);
        const R xa2 = xa[0] * xa[0];
        h[1] = gv * ((R)1 - t.phi) * t.mu;
        h[2] = gv * (-t.mu + t.lam * xa2 + ((R)1 - t.lam) * xa[1]) * ((R)1 - t.phi) * t.phi;
        h[3] = gv * t.phi * (xa2 - xa[1]) * ((R)1 - t.lam) * t.lam;
        const R dy = y - xn[0];
        h[0] = t.invLR - (dy * dy) * t.LRinv;
    }
    // garch/kernels.py:28-34 : sigma2 from the candidate parent
    template <class R> static __device__ __forceinline__ R log_trans(const Theta<R>& t, const R* xa, const R* xn) {
        const R s2 = sigma2_next(t, xa);
        return (R)-0.5 * (xn[0] * xn[0]) * Mth<R>::rcp(s2) - (R)(0.5 * LOG_2PI_D) - (R)0.5 * Mth<R>::log(s2);
    }
    template <class R> static __device__ __forceinline__ R log_trans_max(const Theta<R>& t) { return t.ltmax; }
    template <class R> static __device__ __forceinline__ void jkey(const Theta<R>& t, const R* xa, R* k) {
        const R s2 = sigma2_next(t, xa);
        k[0] = Mth<R>::rcp(s2); k[1] = -(R)(0.5 * LOG_2PI_D) - (R)0.5 * Mth<R>::log(s2);
    }
    template <class R> static __device__ __forceinline__ R log_trans_key(const Theta<R>& t, const R* k, const R* xn) {
        return (R)-0.5 * (xn[0] * xn[0]) * k[0] + k[1];
    }
    // garch/helper.py:414-434: [x', x'^2, x'^4]
    template <class R> static __device__ __forceinline__ void suff(const R* xa, const R* xn, R* h) {
        const R x2 = xn[0] * xn[0];
        h[0] = xn[0]; h[1] = x2; h[2] = x2 * x2;
    }
    // garch/kernels.py:99-104 : sigma2_0 = 0
    template <class R> static __device__ __forceinline__ void init(R mean, R sd, R z, R* x) { x[0] = mean + sd * z; x[1] = (R)0; }

    // O(N^2) backward weights: log q(x'_i | x_j) - log_trans_max = (-0.5 x'^2) (1 / s2_j) + (-0.5 log s2_j + 0.5 log alpha)
    template <class R> static __device__ __forceinline__ void pair_child(const Theta<R>& t, const R* xn, R& b, R& a) {
        b = (R)-0.5 * (xn[0] * xn[0]); a = (R)0;
    }
    template <class R> static __device__ __forceinline__ void pair_parent(const Theta<R>& t, const R* xa, R& u, R& gq, R* m) {
        const R s2 = sigma2_next(t, xa);
        u = Mth<R>::rcp(s2);
        gq = (R)-0.5 * Mth<R>::log(s2) - (R)(0.5 * LOG_2PI_D) - t.ltmax;
        const R xa2 = xa[0] * xa[0];
        m[0] = -t.mu + t.lam * xa2 + ((R)1 - t.lam) * xa[1]; m[1] = xa2 - xa[1];
    }
    // sup over s2 >= alpha of  -0.5 x'^2 / s2 - 0.5 log(s2 / alpha)  is attained at s2 = max(x'^2, alpha): for a child
    // in the tails the global maximum (s2 = alpha, x' = 0; garch/kernels.py:131-133) under-estimates the acceptance
    // probability by a factor exp(x'^2 / (2 alpha))-ish, which is what makes plain accept-reject stall there.
    template <class R> static __device__ __forceinline__ R pair_bound(const Theta<R>& t, const R* xn) {
        const R x2 = xn[0] * xn[0];
        const R s2 = x2 > t.alpha ? x2 : t.alpha;
        return (R)-0.5 * x2 * Mth<R>::rcp(s2) - (R)0.5 * Mth<R>::log(s2 * Mth<R>::rcp(t.alpha));
    }
    // garch/helper.py:350-372 with the candidate parent's terms replaced by their moments
    template <class R> static __device__ __forceinline__ void score_moments(const Theta<R>& t, const R* Em, const R* xn, R y, R* h) {
        const R v = xn[1];
        const R iv = Mth<R>::rcp(v);
        const R gv = (R)-0.5 * (v - xn[0] * xn[0]) * (iv * iv);
        h[1] = gv * ((R)1 - t.phi) * t.mu;
        h[2] = gv * Em[0] * ((R)1 - t.phi) * t.phi;
        h[3] = gv * t.phi * Em[1] * ((R)1 - t.lam) * t.lam;
        const R dy = y - xn[0];
        h[0] = t.invLR - (dy * dy) * t.LRinv;
    }
    template <class R> static __device__ __forceinline__ void suff_moments(const R* Em, const R* xn, R* h) {
        const R x2 = xn[0] * xn[0];
        h[0] = xn[0]; h[1] = x2; h[2] = x2 * x2; h[3] = (R)0;
    }
    // garch/helper.py:374-412: y ~ N(x_pred, R); x_pred advanced with the PRIOR kernel (one normal per horizon)
    static constexpr bool PRED_RNG = true;
    template <class R> static __device__ __forceinline__ void pred_begin(const R* xn, R* ps) { ps[0] = xn[0]; ps[1] = xn[1]; }
    template <class R> static __device__ __forceinline__ R pred_ll(const Theta<R>& t, const R* ps, R yk, R z) {
        const R diff = yk - ps[0];
        return (R)-0.5 * (diff * diff) / t.Rv + (R)(-0.5 * LOG_2PI_D) - (R)0.5 * Mth<R>::log(t.Rv);
    }
    template <class R> static __device__ __forceinline__ void pred_next(const Theta<R>& t, R* ps, R z) {
        const R s2 = t.alpha + t.beta * (ps[0] * ps[0]) + t.gamma * ps[1];
        ps[0] = Mth<R>::sqrt(s2) * z; ps[1] = s2;
    }
};

struct GarchOptimal : GarchPrior {
    // garch/kernels.py:147-158
    template <class R> static __device__ __forceinline__ void propagate(const Theta<R>& t, const R* xa, R y, R z, R* xn) {
        const R s2 = sigma2_next(t, xa);
        const R var = Mth<R>::rcp(Mth<R>::add(t.Rinv, Mth<R>::rcp(s2)));
        const R mean = Mth<R>::mul(var, Mth<R>::mul(y, t.Rinv));
        xn[0] = Mth<R>::fma(Mth<R>::sqrt(var), z, mean);
        xn[1] = s2;
    }
    // garch/kernels.py:172-180
    template <class R> static __device__ __forceinline__ R log_weight(const Theta<R>& t, const R* xa, const R* xn, R y) {
        const R var = Mth<R>::add(xn[1], t.Rv);
        const R q = Mth<R>::fma(Mth<R>::mul((R)-0.5, Mth<R>::mul(y, y)), Mth<R>::rcp(var), (R)(-0.5 * LOG_2PI_D));
        return Mth<R>::fma((R)-0.5, Mth<R>::log(var), q);
    }
};

}  // namespace sgm
