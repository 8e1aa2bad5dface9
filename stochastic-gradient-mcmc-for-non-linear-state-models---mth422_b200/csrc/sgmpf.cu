// C-ABI entry points (include/sgmpf.h) and launch orchestration of the batched buffered particle
// filter.  One sgm_pf_run() = init kernel + max_T step launches (+ backward kernels for the
// O(N^2) / PaRIS smoothers) + final reduce, all asynchronous on the caller's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sgmpf.h"
#include "pf_kernels.cuh"
#include "backward_kernels.cuh"
#include "ksd_kernel.cuh"

using namespace sgm;

namespace {

thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

int fail(int code, const char* fmt, const char* extra = "") {
    snprintf(g_err, sizeof(g_err), fmt, extra);
    return code;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout {
    size_t rec[2], tail[2], fine[2], lw[2], sub[2], hdr, acc, thc, yw, Jidx, Llist[2], counters, pcdf, pguide, pkey, n2part, total;
};

bool backward_pf(int pf) { return pf == SGM_PF_POY_N2 || pf == SGM_PF_PARIS; }

int validate(const sgm_pf_desc* d) {
    if (!d) return fail(SGM_ERR_INVALID, "null descriptor");
    if (d->struct_bytes != (int32_t)sizeof(sgm_pf_desc)) return fail(SGM_ERR_INVALID, "sgm_pf_desc size mismatch (header / library out of sync)");
    if (d->model < 0 || d->model > SGM_MODEL_GARCH) return fail(SGM_ERR_INVALID, "unknown model");
    if (d->kernel != SGM_KERNEL_PRIOR && d->kernel != SGM_KERNEL_OPTIMAL) return fail(SGM_ERR_INVALID, "unknown kernel");
    if (d->model == SGM_MODEL_SVM && d->kernel == SGM_KERNEL_OPTIMAL) return fail(SGM_ERR_UNSUPPORTED, "SVM optimal kernel not analytic");
    if (d->pf < 0 || d->pf > SGM_PF_FILTER) return fail(SGM_ERR_INVALID, "unknown pf");
    if (d->dtype != SGM_F32 && d->dtype != SGM_F64) return fail(SGM_ERR_INVALID, "unknown dtype");
    if (d->rng_mode != SGM_RNG_PHILOX && d->rng_mode != SGM_RNG_INJECTED) return fail(SGM_ERR_INVALID, "unknown rng_mode");
    if (d->resample < 0 || d->resample > SGM_RESAMPLE_STRATIFIED) return fail(SGM_ERR_INVALID, "unknown resample");
    if (d->stat_kind < 0 || d->stat_kind > SGM_STAT_PRED) return fail(SGM_ERR_INVALID, "unknown stat_kind");
    if (d->stat_kind == SGM_STAT_PRED) {
        if (d->pf != SGM_PF_FILTER) return fail(SGM_ERR_INVALID, "Only can use pf = 'filter' since we are filtering");
        if (d->pred_steps_ahead < 0 || d->pred_steps_ahead > 7) return fail(SGM_ERR_UNSUPPORTED, "num_steps_ahead must be in [0, 7]");
        if (d->rng_mode == SGM_RNG_INJECTED && d->model != SGM_MODEL_LGSSM && !d->inj_pred)
            return fail(SGM_ERR_INVALID, "INJECTED predictive statistic needs inj_pred");
    }
    if (d->n_items < 1 || d->n_items > 65535) return fail(SGM_ERR_INVALID, "n_items must be in [1, 65535]");
    if (d->n_particles < 1 || d->n_particles > WT * MAX_Q) return fail(SGM_ERR_INVALID, "n_particles must be in [1, 2^20]");
    if (d->max_T < 0 || d->max_T > 65000) return fail(SGM_ERR_INVALID, "max_T out of range");
    if (!d->obs || !d->obs_off || !d->T_buf || !d->t1 || !d->tL || !d->theta || !d->prior_mean || !d->prior_var)
        return fail(SGM_ERR_INVALID, "missing per-item input array");
    if (!d->grad || !d->loglik || !d->status) return fail(SGM_ERR_INVALID, "missing output array");
    if (d->rng_mode == SGM_RNG_INJECTED && (!d->inj_z0 || (d->max_T > 0 && (!d->inj_u || !d->inj_z))))
        return fail(SGM_ERR_INVALID, "INJECTED rng_mode needs inj_z0 / inj_u / inj_z");
    if (d->pf == SGM_PF_PARIS) {
        if (d->Ntilde < 1 || d->Ntilde > 8) return fail(SGM_ERR_INVALID, "Ntilde must be in [1, 8]");
        if (d->rng_mode == SGM_RNG_INJECTED && (!d->inj_extra || !d->inj_extra_off))
            return fail(SGM_ERR_INVALID, "INJECTED PaRIS needs inj_extra / inj_extra_off");
    }
    if (d->step_weights && !d->wts_off) return fail(SGM_ERR_INVALID, "step_weights given without wts_off");
    if (d->n2_mode < SGM_N2_AUTO || d->n2_mode > SGM_N2_TENSOR || d->reserved0 != 0) return fail(SGM_ERR_INVALID, "unknown n2_mode");
    if (d->n2_mode == SGM_N2_TENSOR && d->dtype != SGM_F32) return fail(SGM_ERR_UNSUPPORTED, "the tensor-core O(N^2) smoother needs dtype f32");
    return SGM_OK;
}

int state_dim(int model) { return model == SGM_MODEL_GARCH ? 2 : 1; }
int score_dim(int model) { return model == SGM_MODEL_SVM ? 3 : 4; }

Layout make_layout(const sgm_pf_desc* d) {
    Layout L;
    memset(&L, 0, sizeof(L));
    const size_t es = d->dtype == SGM_F64 ? 8 : 4;
    const size_t B = d->n_items, N = d->n_particles, Q = (N + WT - 1) / WT;
    const size_t KT = state_dim(d->model) + score_dim(d->model) - 4;
    const bool need_lw = backward_pf(d->pf) || d->out_lw || d->trace_lw || d->stat_kind == SGM_STAT_PRED;
    size_t off = 0;
    for (int k = 0; k < 2; ++k) { L.rec[k] = off; off = align_up(off + B * N * 4 * es); }
    for (int k = 0; k < 2; ++k) { L.tail[k] = off; off = align_up(off + B * N * KT * es); }
    for (int k = 0; k < 2; ++k) { L.fine[k] = off; off = align_up(off + B * Q * WT * es); }
    for (int k = 0; k < 2; ++k) { L.lw[k] = off; off = align_up(off + (need_lw ? B * N * es : 0)); }
    for (int k = 0; k < 2; ++k) { L.sub[k] = off; off = align_up(off + B * Q * SSTRIDE * 8); }
    L.hdr = off; off = align_up(off + B * hdr_stride((int)Q) * 8);
    L.acc = off; off = align_up(off + B * ACC_STRIDE * 8);
    L.thc = off; off = align_up(off + B * THC_BYTES);
    L.yw = off; off = align_up(off + B * (size_t)(d->max_T > 0 ? d->max_T : 1) * 2 * es);
    if (d->pf == SGM_PF_PARIS) {
        L.Jidx = off; off = align_up(off + B * N * (size_t)d->Ntilde * 4);
        L.Llist[0] = off; off = align_up(off + B * N * (size_t)d->Ntilde * 4);
        L.Llist[1] = off; off = align_up(off + B * N * 4);
        L.counters = off; off = align_up(off + B * 16 * 4);
        L.pcdf = off; off = align_up(off + B * N * 8);
        L.pguide = off; off = align_up(off + B * (N + 1) * 4);
        L.pkey = off; off = align_up(off + B * N * 4 * es);
    }
    if (d->pf == SGM_PF_POY_N2) {
        L.n2part = off; off = align_up(off + n2_partial_bytes(d->dtype, d->n2_mode, (int)B, (int)N));
    }
    L.total = off;
    return L;
}

template <class R, class Model>
int run_impl(const sgm_pf_desc* d, cudaStream_t stream) {
    const Layout L = make_layout(d);
    if (!d->workspace || d->workspace_bytes < L.total) return fail(SGM_ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)d->workspace & 255) != 0) return fail(SGM_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    char* ws = reinterpret_cast<char*>(d->workspace);
    KArgs a;
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

    const dim3 grid(a.G, a.B), block(NT);
    int64_t launches = 0;
    const bool pred = d->stat_kind == SGM_STAT_PRED;
    const bool fused = (a.Q <= NWARP) && !backward_pf(d->pf) && !pred;
    if (fused) {
        // small N: the whole time loop of an item in one launch (one CTA per item)
        if (d->ev_steps_begin) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_begin), stream);
        if (d->resample == SGM_RESAMPLE_MULTINOMIAL) pf_fused_kernel<R, Model, false><<<a.B, block, 0, stream>>>(a);
        else pf_fused_kernel<R, Model, true><<<a.B, block, 0, stream>>>(a);
        ++launches;
        if (d->ev_steps_end) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_end), stream);
    } else {
        // Two-stream pipelining (O(N) smoothers, big batches): the halves alternate on `stream` / `aux_stream`, so
        // the one-CTA-per-item header kernel and the launch gap of one half overlap the step kernel of the other.
        const bool piped = d->aux_stream && d->ev_aux_fork && d->ev_aux_join && !backward_pf(d->pf) && !pred &&
                           (int64_t)a.B * a.G >= 2 * 148 * 4 && a.B >= 2;
        const int nh = piped ? 2 : 1;
        // the production configuration runs the step kernel with its run-time flags folded to constants
        const bool fast_path = sizeof(R) == 4 && d->rng_mode == SGM_RNG_PHILOX && d->resample == SGM_RESAMPLE_MULTINOMIAL_SORTED &&
                               d->pf == SGM_PF_NEMETH && d->lambduh == 1.0 && d->stat_kind == SGM_STAT_SCORE &&
                               a.N % WT == 0 && !a.need_lw && !d->trace_anc && !d->trace_x && !d->trace_lw;
        cudaStream_t sh[2] = {stream, piped ? reinterpret_cast<cudaStream_t>(d->aux_stream) : stream};
        KArgs ah[2] = {a, a};
        int nb[2] = {piped ? a.B / 2 : a.B, piped ? a.B - a.B / 2 : 0};
        ah[1].b0 = nb[0];
        if (piped) {
            cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_aux_fork), stream);
            cudaStreamWaitEvent(sh[1], reinterpret_cast<cudaEvent_t>(d->ev_aux_fork), 0);
        }
        for (int h = 0; h < nh; ++h) { pf_init_kernel<R, Model><<<dim3(a.G, nb[h]), block, 0, sh[h]>>>(ah[h]); ++launches; }
        if (d->ev_steps_begin) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_begin), stream);
        for (int t = 0; t < a.max_T; ++t) {
            for (int h = 0; h < nh; ++h) {
                const dim3 gh((a.Q + STEP_WARPS - 1) / STEP_WARPS, nb[h]), bs(32 * STEP_WARPS);
                pf_header_kernel<R, Model><<<nb[h], block, 0, sh[h]>>>(ah[h], t, 0); ++launches;
                if (d->resample == SGM_RESAMPLE_MULTINOMIAL) pf_step_kernel<R, Model, false><<<gh, bs, 0, sh[h]>>>(ah[h], t);
                else if (fast_path) pf_step_kernel<R, Model, true, true><<<gh, bs, 0, sh[h]>>>(ah[h], t);
                else pf_step_kernel<R, Model, true><<<gh, bs, 0, sh[h]>>>(ah[h], t);
                ++launches;
            }
            if (pred) { pf_pred_kernel<R, Model><<<a.B, block, 0, stream>>>(a, t); ++launches; }
            if (d->pf == SGM_PF_POY_N2) launches += launch_poyiadjis_n2<R, Model>(a, t, stream);
            else if (d->pf == SGM_PF_PARIS) launches += launch_paris<R, Model>(a, t, stream);
        }
        for (int h = 0; h < nh; ++h) { pf_header_kernel<R, Model><<<nb[h], block, 0, sh[h]>>>(ah[h], a.max_T, 1); ++launches; }
        if (piped) {
            cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_aux_join), sh[1]);
            cudaStreamWaitEvent(stream, reinterpret_cast<cudaEvent_t>(d->ev_aux_join), 0);
        }
        if (d->ev_steps_end) cudaEventRecord(reinterpret_cast<cudaEvent_t>(d->ev_steps_end), stream);
    }
    if (d->out_x || d->out_lw || d->out_stats) {
        pf_export_kernel<R, Model><<<dim3((a.N + NT - 1) / NT, a.B), block, 0, stream>>>(a); ++launches;
    }
    g_launches = launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
    return SGM_OK;
}

template <class R>
int run_model(const sgm_pf_desc* d, cudaStream_t s) {
    switch (d->model) {
        case SGM_MODEL_SVM: return run_impl<R, SvmPrior>(d, s);
        case SGM_MODEL_LGSSM: return d->kernel == SGM_KERNEL_PRIOR ? run_impl<R, LgssmPrior>(d, s) : run_impl<R, LgssmOptimal>(d, s);
        default: return d->kernel == SGM_KERNEL_PRIOR ? run_impl<R, GarchPrior>(d, s) : run_impl<R, GarchOptimal>(d, s);
    }
}

}  // namespace

extern "C" {

int sgm_version(void) { return SGM_VERSION; }

const char* sgm_last_error(void) { return g_err; }

int64_t sgm_last_launch_count(void) { return g_launches; }

int sgm_device_check(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        cudaGetLastError();
        return fail(SGM_ERR_DEVICE, "no usable CUDA device");
    }
    if (p.major != 10) return fail(SGM_ERR_DEVICE, "this library is built for sm_100a (compute capability 10.x) only");
    return SGM_OK;
}

int sgm_state_dim(int32_t model) {
    if (model < 0 || model > SGM_MODEL_GARCH) return SGM_ERR_INVALID;
    return state_dim(model);
}

int sgm_stat_dim(int32_t model, int32_t stat_kind) {
    if (model < 0 || model > SGM_MODEL_GARCH) return SGM_ERR_INVALID;
    if (stat_kind == SGM_STAT_SCORE) return score_dim(model);
    if (stat_kind == SGM_STAT_SUFF) return 3;
    if (stat_kind == SGM_STAT_NONE) return 0;
    if (stat_kind == SGM_STAT_PRED) return 8;          /* upper bound: num_steps_ahead + 1 <= 8 */
    return SGM_ERR_INVALID;
}

int sgm_ksd_imq(const double* x, const double* gradlogp, int32_t num_points, int32_t dim, double c, double beta,
                double* partial, void* stream) {
    if (!x || !gradlogp || !partial) return fail(SGM_ERR_INVALID, "sgm_ksd_imq: null pointer");
    if (num_points < 1 || dim < 1 || dim > KSD_MAXD) return fail(SGM_ERR_UNSUPPORTED, "sgm_ksd_imq: dim must be in [1, 8]");
    g_err[0] = 0;
    const int blocks = (num_points + NT - 1) / NT;
    ksd_imq_kernel<<<blocks, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, gradlogp, num_points, dim, c * c, beta, partial);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
    return SGM_OK;
}

uint64_t sgm_pf_workspace_bytes(const sgm_pf_desc* d) {
    if (validate(d) != SGM_OK) return 0;
    return (uint64_t)make_layout(d).total;
}

int sgm_pf_run(const sgm_pf_desc* d, void* stream) {
    const int v = validate(d);
    if (v != SGM_OK) return v;
    g_err[0] = 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return d->dtype == SGM_F64 ? run_model<double>(d, s) : run_model<float>(d, s);
}

}  // extern "C"
