// C-ABI entry points (include/sgmpf.h): descriptor validation, workspace layout, dtype dispatch.  The launch
// orchestration (one sgm_pf_run() = init kernel + max_T step launches (+ backward kernels for the O(N^2) / PaRIS
// smoothers) + final reduce, all asynchronous on the caller's stream) lives in run_impl.cuh and is instantiated per
// arithmetic type in sgmpf_f32.cu / sgmpf_f64.cu.
#include "host_common.cuh"
#include "ksd_kernel.cuh"
#include "sgld_kernels.cuh"

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
    if (d->path < SGM_PATH_AUTO || d->path > SGM_PATH_STEPS || d->reserved1 != 0) return fail(SGM_ERR_INVALID, "unknown path");
    if (d->path == SGM_PATH_CLUSTER && (d->n_particles <= 256 || d->n_particles > 16384 || d->pf == SGM_PF_POY_N2 || d->pf == SGM_PF_PARIS ||
                                        d->stat_kind == SGM_STAT_PRED || (int64_t)d->n_items * 2 > 148))
        return fail(SGM_ERR_UNSUPPORTED, "path = CLUSTER needs 256 < N <= 16384, an O(N) smoother and items x cluster size <= 148");
    if (d->path == SGM_PATH_SMALL && (d->n_particles > 2048 || d->pf == SGM_PF_POY_N2 || d->pf == SGM_PF_PARIS || d->stat_kind == SGM_STAT_PRED))
        return fail(SGM_ERR_UNSUPPORTED, "path = SMALL needs N <= 2048 and an O(N) smoother");
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

namespace {

struct SgldLayout {
    size_t obs_off, T_buf, t1, tL, wts_off, weights, theta, pm, pv, item_seq, grad, loglik, status, pf_ws, total;
    int B, max_T, nsel, Smax;
};

int sgld_geometry(const sgm_sgld_desc* d, int& B, int& max_T, int& nsel, int& Smax) {
    if (!d) return fail(SGM_ERR_INVALID, "null descriptor");
    if (d->struct_bytes != (int32_t)sizeof(sgm_sgld_desc)) return fail(SGM_ERR_INVALID, "sgm_sgld_desc size mismatch (header / library out of sync)");
    if (d->method < SGM_STEP_SGLD || d->method > SGM_STEP_SGD) return fail(SGM_ERR_INVALID, "unknown method");
    if (d->method == SGM_STEP_SGRLD && d->pf.model != SGM_MODEL_LGSSM)
        return fail(SGM_ERR_UNSUPPORTED, "No Default Preconditioner: SGRLD on the device covers the LGSSM preconditioner only");
    if (d->n_chains < 1 || d->minibatch < 1 || d->n_seqs < 1 || d->n_iters < 0) return fail(SGM_ERR_INVALID, "n_chains / minibatch / n_seqs / n_iters out of range");
    if (d->num_sequences != -1 && (d->num_sequences < 1 || d->num_sequences > 8 || d->num_sequences > d->n_seqs))
        return fail(SGM_ERR_UNSUPPORTED, "num_sequences must be -1 or in [1, min(8, n_seqs)]");
    if (d->partition < SGM_PARTITION_UNIFORM || d->partition > SGM_PARTITION_STRICT) return fail(SGM_ERR_INVALID, "unknown partition");
    if (d->subsequence_length == 0 || d->subsequence_length < -1 || d->buffer_length < -1 || d->max_seq_len < 1)
        return fail(SGM_ERR_INVALID, "subsequence_length / buffer_length / max_seq_len out of range");
    if (d->pf.stat_kind != SGM_STAT_SCORE) return fail(SGM_ERR_INVALID, "the SG-MCMC loop needs stat_kind = SGM_STAT_SCORE");
    if (d->pf.pf == SGM_PF_PARIS && d->pf.rng_mode == SGM_RNG_INJECTED) return fail(SGM_ERR_UNSUPPORTED, "INJECTED PaRIS inside the device loop");
    nsel = d->num_sequences == -1 ? d->n_seqs : d->num_sequences;
    const int64_t items = (int64_t)d->n_chains * nsel * d->minibatch;
    if (items > 65535) return fail(SGM_ERR_INVALID, "more than 65535 work items per iteration");
    B = (int)items;
    const int S = d->subsequence_length;
    if (S == -1 || d->buffer_length == -1) max_T = d->max_seq_len;
    else { const int64_t w = (int64_t)S + 2 * (int64_t)d->buffer_length; max_T = (int)(w < d->max_seq_len ? w : d->max_seq_len); }
    // a sequence shorter than S runs whole
    if (S != -1 && max_T < (S < d->max_seq_len ? S : d->max_seq_len)) max_T = S < d->max_seq_len ? S : d->max_seq_len;
    Smax = S == -1 ? 1 : S;
    return SGM_OK;
}

void sgld_pf_template(const sgm_sgld_desc* d, int B, int max_T, sgm_pf_desc& pf) {
    pf = d->pf;
    pf.struct_bytes = (int32_t)sizeof(sgm_pf_desc);
    pf.n_items = B; pf.max_T = max_T;
    pf.out_x = pf.out_lw = pf.out_stats = nullptr;
    pf.trace_anc = nullptr; pf.trace_x = pf.trace_lw = nullptr; pf.trace_J = nullptr;
    pf.ev_steps_begin = pf.ev_steps_end = nullptr;
    // placeholders so that the layout / validation see a complete descriptor
    static const double dummy = 0.0;
    pf.obs = &dummy; pf.obs_off = reinterpret_cast<const int64_t*>(&dummy); pf.T_buf = pf.t1 = pf.tL = reinterpret_cast<const int32_t*>(&dummy);
    pf.theta = pf.prior_mean = pf.prior_var = &dummy;
    pf.step_weights = &dummy; pf.wts_off = reinterpret_cast<const int64_t*>(&dummy);
    pf.grad = pf.loglik = const_cast<double*>(&dummy); pf.status = reinterpret_cast<int32_t*>(const_cast<double*>(&dummy));
}

int sgld_layout(const sgm_sgld_desc* d, SgldLayout& L) {
    memset(&L, 0, sizeof(L));
    int rc = sgld_geometry(d, L.B, L.max_T, L.nsel, L.Smax);
    if (rc != SGM_OK) return rc;
    sgm_pf_desc pf;
    sgld_pf_template(d, L.B, L.max_T, pf);
    rc = validate(&pf);
    if (rc != SGM_OK) return rc;
    const size_t B = L.B;
    size_t off = 0;
    L.obs_off = off; off = align_up(off + B * 8);
    L.T_buf = off; off = align_up(off + B * 4);
    L.t1 = off; off = align_up(off + B * 4);
    L.tL = off; off = align_up(off + B * 4);
    L.wts_off = off; off = align_up(off + B * 8);
    L.weights = off; off = align_up(off + B * (size_t)L.Smax * 8);
    L.theta = off; off = align_up(off + B * SGM_THETA_STRIDE * 8);
    L.pm = off; off = align_up(off + B * 8);
    L.pv = off; off = align_up(off + B * 8);
    L.item_seq = off; off = align_up(off + B * 4);
    L.grad = off; off = align_up(off + B * 8 * 8);
    L.loglik = off; off = align_up(off + B * 8);
    L.status = off; off = align_up(off + B * 4);
    L.pf_ws = off; off = align_up(off + make_layout(&pf).total);
    L.total = off;
    return SGM_OK;
}

}  // namespace

extern "C" {

int32_t sgm_sgld_items(const sgm_sgld_desc* d) {
    int B, max_T, nsel, Smax;
    const int rc = sgld_geometry(d, B, max_T, nsel, Smax);
    return rc != SGM_OK ? rc : B;
}
int32_t sgm_sgld_max_steps(const sgm_sgld_desc* d) {
    int B, max_T, nsel, Smax;
    const int rc = sgld_geometry(d, B, max_T, nsel, Smax);
    return rc != SGM_OK ? rc : max_T;
}
uint64_t sgm_sgld_workspace_bytes(const sgm_sgld_desc* d) {
    SgldLayout L;
    if (sgld_layout(d, L) != SGM_OK) return 0;
    return (uint64_t)L.total;
}

int sgm_sgld_run(const sgm_sgld_desc* d, void* stream) {
    SgldLayout L;
    int rc = sgld_layout(d, L);
    if (rc != SGM_OK) return rc;
    g_err[0] = 0;
    if (!d->obs || !d->seq_off || !d->params || !d->hyper || !d->chain_status || !d->offset_dev || !d->iter_dev)
        return fail(SGM_ERR_INVALID, "missing array (obs / seq_off / params / hyper / chain_status / offset_dev / iter_dev)");
    if (d->prior_x0 == 0 && (!d->prior_mean || !d->prior_var)) return fail(SGM_ERR_INVALID, "prior_mean / prior_var missing");
    if (d->prior_x0 != 0 && (d->prior_x0 != 1 || d->pf.model != SGM_MODEL_GARCH)) return fail(SGM_ERR_INVALID, "prior_x0 = 1 is the GARCH stationary prior");
    const bool injected = d->pf.rng_mode == SGM_RNG_INJECTED;
    if (injected && ((d->subsequence_length != -1 && !d->inj_start) || (d->method != SGM_STEP_SGD && !d->inj_noise) ||
                     (d->num_sequences != -1 && !d->inj_seq)))
        return fail(SGM_ERR_INVALID, "INJECTED mode needs inj_start / inj_noise (/ inj_seq)");
    if (!d->workspace || d->workspace_bytes < L.total) return fail(SGM_ERR_WORKSPACE, "workspace too small");
    if (((uintptr_t)d->workspace & 255) != 0) return fail(SGM_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    char* ws = reinterpret_cast<char*>(d->workspace);

    SgldArgs a;
    memset(&a, 0, sizeof(a));
    a.model = d->pf.model; a.method = d->method; a.C = d->n_chains; a.M = d->minibatch; a.nsel = L.nsel; a.n_seqs = d->n_seqs;
    a.S = d->subsequence_length; a.Bf = d->buffer_length; a.partition = d->partition; a.project = d->project;
    a.prior_x0 = d->prior_x0; a.injected = injected ? 1 : 0; a.ipc = L.nsel * d->minibatch; a.Smax = L.Smax;
    a.pick_all = d->num_sequences == -1 ? 1 : 0;
    a.epsilon = d->epsilon; a.T_total = d->T_total;
    a.key.k0 = (uint32_t)(d->pf.seed & 0xffffffffu); a.key.k1 = (uint32_t)(d->pf.seed >> 32);
    a.key.item = (uint32_t)(d->pf.item_id_base / (a.ipc > 0 ? a.ipc : 1)); a.key.offset = 0;
    a.offset_dev = d->offset_dev; a.offset_rw = d->offset_dev; a.iter_dev = d->iter_dev;
    a.obs = d->obs; a.seq_off = d->seq_off; a.params = d->params; a.hyper = d->hyper;
    a.prior_mean = d->prior_mean; a.prior_var = d->prior_var; a.chain_status = d->chain_status;
    a.trace = d->trace; a.trace_every = d->trace_every; a.trace_rows = d->trace_rows;
    a.inj_start = d->inj_start; a.inj_seq = d->inj_seq; a.inj_noise = d->inj_noise;
    a.obs_off = reinterpret_cast<int64_t*>(ws + L.obs_off); a.T_buf = reinterpret_cast<int32_t*>(ws + L.T_buf);
    a.t1 = reinterpret_cast<int32_t*>(ws + L.t1); a.tL = reinterpret_cast<int32_t*>(ws + L.tL);
    a.wts_off = reinterpret_cast<int64_t*>(ws + L.wts_off); a.weights = reinterpret_cast<double*>(ws + L.weights);
    a.theta = reinterpret_cast<double*>(ws + L.theta); a.item_pm = reinterpret_cast<double*>(ws + L.pm);
    a.item_pv = reinterpret_cast<double*>(ws + L.pv); a.item_seq = reinterpret_cast<int32_t*>(ws + L.item_seq);
    a.grad = reinterpret_cast<double*>(ws + L.grad); a.loglik = reinterpret_cast<double*>(ws + L.loglik);
    a.item_status = reinterpret_cast<int32_t*>(ws + L.status);

    sgm_pf_desc pf;
    sgld_pf_template(d, L.B, L.max_T, pf);
    pf.obs = d->obs; pf.obs_off = a.obs_off; pf.T_buf = a.T_buf; pf.t1 = a.t1; pf.tL = a.tL;
    pf.step_weights = a.weights; pf.wts_off = a.wts_off; pf.theta = a.theta; pf.prior_mean = a.item_pm; pf.prior_var = a.item_pv;
    pf.grad = reinterpret_cast<double*>(ws + L.grad); pf.loglik = reinterpret_cast<double*>(ws + L.loglik);
    pf.status = reinterpret_cast<int32_t*>(ws + L.status);
    pf.workspace = ws + L.pf_ws; pf.workspace_bytes = L.total - L.pf_ws;
    pf.offset_dev = d->offset_dev;

    int64_t launches = 0;
    const size_t N = (size_t)pf.n_particles, B = (size_t)L.B, T = (size_t)L.max_T;
    if (d->n_iters > 0 && !d->no_persistent) {
        // one item per chain and N <= 2048: all iterations inside one persistent launch
        rc = pf.dtype == SGM_F64 ? sgmhost::run_sgld_persistent<double>(&pf, a, d->n_iters, s)
                                 : sgmhost::run_sgld_persistent<float>(&pf, a, d->n_iters, s);
        if (rc < 0) return rc;
        if (rc == 1) return SGM_OK;
    }
    for (int k = 0; k < d->n_iters; ++k) {
        sgld_prepare_kernel<<<(L.B + 127) / 128, 128, 0, s>>>(a, k);
        if (injected) {
            pf.inj_z0 = d->pf.inj_z0 + (size_t)k * B * N;
            pf.inj_u = d->pf.inj_u + (size_t)k * B * T * N;
            pf.inj_z = d->pf.inj_z + (size_t)k * B * T * N;
        }
        rc = pf.dtype == SGM_F64 ? sgmhost::run_model<double>(&pf, s) : sgmhost::run_model<float>(&pf, s);
        if (rc != SGM_OK) return rc;
        launches += g_launches + 2;
        sgld_update_kernel<<<1, 256, 0, s>>>(a, k);
    }
    g_launches = launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
    return SGM_OK;
}

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

namespace {
__global__ void selftest_math_kernel(int what, const double* x, double* y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (what == SGM_SELFTEST_LOG) y[i] = sgm::fast_log(x[i]);
    else {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(x[i]);
        sgm::fast_sincos2pi((uint32_t)(bits >> 32), (uint32_t)bits, &y[2 * i], &y[2 * i + 1]);
    }
}
}  // namespace

int sgm_selftest_math(int32_t what, const double* x, double* y, int64_t n, void* stream) {
    if (!x || !y || n < 0) return fail(SGM_ERR_INVALID, "sgm_selftest_math: null pointer or negative size");
    if (what != SGM_SELFTEST_LOG && what != SGM_SELFTEST_SINCOS2PI) return fail(SGM_ERR_INVALID, "sgm_selftest_math: unknown function");
    if (n == 0) return SGM_OK;
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(what, x, y, n);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SGM_ERR_CUDA, "CUDA launch failed: %s", cudaGetErrorString(e));
    return SGM_OK;
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
