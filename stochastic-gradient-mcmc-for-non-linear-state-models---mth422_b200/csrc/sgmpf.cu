// C-ABI entry points (include/sgmpf.h): descriptor validation, workspace layout, dtype dispatch.  The launch
// orchestration (one sgm_pf_run() = init kernel + max_T step launches (+ backward kernels for the O(N^2) / PaRIS
// smoothers) + final reduce, all asynchronous on the caller's stream) lives in run_impl.cuh and is instantiated per
// arithmetic type in sgmpf_f32.cu / sgmpf_f64.cu.
#include "host_common.cuh"
#include "ksd_kernel.cuh"

using namespace sgm;

namespace {


thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

}  // namespace

namespace sgmhost {
int fail(int code, const char* fmt, const char* extra) {
    snprintf(g_err, sizeof(g_err), fmt, extra);
    return code;
}
void set_launch_count(int64_t n) { g_launches = n; }
}  // namespace sgmhost

using namespace sgmhost;

namespace {

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
        if (d->pred_steps_ahead < 0 || d->pred_steps_ahead > SGM_PRED_MAX_STEPS)
            return fail(SGM_ERR_UNSUPPORTED, "num_steps_ahead must be in [0, 14]");
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
    if (d->n2_mode < SGM_N2_AUTO || d->n2_mode > SGM_N2_TENSOR) return fail(SGM_ERR_INVALID, "unknown n2_mode");
    if (d->variates != SGM_VARIATES_NATIVE && d->variates != SGM_VARIATES_F32) return fail(SGM_ERR_INVALID, "unknown variates");
    if (d->n2_mode == SGM_N2_TENSOR && d->dtype != SGM_F32) return fail(SGM_ERR_UNSUPPORTED, "the tensor-core O(N^2) smoother needs dtype f32");
    return SGM_OK;
}


}  // namespace

Layout sgmhost::make_layout(const sgm_pf_desc* d) {
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
    if (stat_kind == SGM_STAT_PRED) return SGM_PRED_MAX_STEPS + 1;          /* upper bound: num_steps_ahead + 1 */
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
    return d->dtype == SGM_F64 ? sgmhost::run_model<double>(d, s) : sgmhost::run_model<float>(d, s);
}

}  // extern "C"
