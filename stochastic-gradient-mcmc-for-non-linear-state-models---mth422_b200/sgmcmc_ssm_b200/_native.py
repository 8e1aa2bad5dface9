"""ctypes binding of the C-ABI declared in include/sgmpf.h.

There is NO fallback: if libsgmpf.so is missing or the device is not a B200 the import of the
compute path fails loudly."""
import ctypes
import os

from .build import LIB_PATH

c_i32, c_i64, c_u64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p

MODEL = dict(lgssm=0, svm=1, garch=2)
KERNEL = dict(prior=0, optimal=1)
PF = dict(nemeth=0, poyiadjis_N=0, poyiadjis_N2=1, paris=2, filter=3)
DTYPE = dict(f32=0, f64=1)
RNG = dict(philox=0, injected=1)
RESAMPLE = dict(multinomial=0, multinomial_sorted=1, sorted=1, systematic=2, stratified=3)
STAT = dict(score=0, suff=1, none=2, pred=3)
N2_MODE = dict(auto=0, fp32_pipe=1, tensor=2)
VARIATES = dict(native=0, f32=1)
PATH = dict(auto=0, tiles=1, small=2, cluster=3, steps=4)
STATUS_NAN_WEIGHT, STATUS_ZERO_WEIGHT, STATUS_AR_OVERFLOW = 1, 2, 4
THETA_STRIDE = 12
PRED_SLOTS, PRED_MAX_STEPS = 16, 14           # SGM_PRED_SLOTS, SGM_PRED_MAX_STEPS (include/sgmpf.h)
ERR = {-1: ValueError, -2: NotImplementedError, -3: MemoryError, -4: RuntimeError, -5: RuntimeError}


class SgmPfDesc(ctypes.Structure):
    _fields_ = [
        ("struct_bytes", c_i32), ("model", c_i32), ("kernel", c_i32), ("pf", c_i32), ("dtype", c_i32),
        ("rng_mode", c_i32), ("resample", c_i32), ("stat_kind", c_i32),
        ("n_items", c_i32), ("n_particles", c_i32), ("max_T", c_i32), ("Ntilde", c_i32),
        ("accept_reject", c_i32), ("max_accept_reject", c_i32), ("manual_sample_threshold", c_i32),
        ("item_id_base", c_i32), ("n2_mode", c_i32), ("pred_steps_ahead", c_i32),
        ("pred_per_horizon", c_i32), ("variates", c_i32), ("path", c_i32), ("reserved1", c_i32),
        ("lambduh", c_f64), ("seed", c_u64), ("offset", c_u64), ("offset_dev", c_vp),
        ("obs", c_vp), ("obs_off", c_vp), ("T_buf", c_vp), ("t1", c_vp), ("tL", c_vp),
        ("step_weights", c_vp), ("wts_off", c_vp), ("theta", c_vp), ("prior_mean", c_vp), ("prior_var", c_vp),
        ("inj_z0", c_vp), ("inj_u", c_vp), ("inj_z", c_vp), ("inj_extra", c_vp), ("inj_extra_off", c_vp), ("inj_pred", c_vp),
        ("grad", c_vp), ("loglik", c_vp), ("status", c_vp),
        ("out_x", c_vp), ("out_lw", c_vp), ("out_stats", c_vp),
        ("trace_anc", c_vp), ("trace_x", c_vp), ("trace_lw", c_vp), ("trace_J", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_u64),
        ("aux_stream", c_vp), ("ev_aux_fork", c_vp), ("ev_aux_join", c_vp),
        ("ev_steps_begin", c_vp), ("ev_steps_end", c_vp),
    ]


STEP = dict(SGLD=0, SGRLD=1, SGD=2)
PARTITION = dict(uniform=0, naive=1, strict=2)
PARAM_STRIDE, HYPER_STRIDE = 8, 16          # SGM_PARAM_STRIDE, SGM_HYPER_STRIDE


class SgmSgldDesc(ctypes.Structure):
    _fields_ = [
        ("struct_bytes", c_i32), ("method", c_i32), ("n_chains", c_i32), ("minibatch", c_i32), ("n_seqs", c_i32),
        ("num_sequences", c_i32), ("subsequence_length", c_i32), ("buffer_length", c_i32), ("partition", c_i32),
        ("n_iters", c_i32), ("project", c_i32), ("prior_x0", c_i32), ("trace_every", c_i32), ("trace_rows", c_i32),
        ("max_seq_len", c_i32), ("no_persistent", c_i32),
        ("epsilon", c_f64), ("T_total", c_f64),
        ("obs", c_vp), ("seq_off", c_vp), ("params", c_vp), ("hyper", c_vp), ("prior_mean", c_vp), ("prior_var", c_vp),
        ("chain_status", c_vp), ("trace", c_vp), ("offset_dev", c_vp), ("iter_dev", c_vp),
        ("inj_start", c_vp), ("inj_seq", c_vp), ("inj_noise", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_u64),
        ("pf", SgmPfDesc),
    ]


EXPORTS = ["sgm_version", "sgm_device_check", "sgm_last_error", "sgm_stat_dim", "sgm_state_dim",
           "sgm_pf_workspace_bytes", "sgm_pf_run", "sgm_last_launch_count", "sgm_ksd_imq",
           "sgm_sgld_items", "sgm_sgld_max_steps", "sgm_sgld_workspace_bytes", "sgm_sgld_run", "sgm_selftest_math"]
_lib = None


def load():
    """Load libsgmpf.so (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "sgmcmc_ssm_b200: CUDA library not built ({0} missing). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `python -m sgmcmc_ssm_b200.build`. There is no CPU fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    lib.sgm_version.restype = c_i32
    lib.sgm_device_check.restype = c_i32
    lib.sgm_last_error.restype = ctypes.c_char_p
    lib.sgm_stat_dim.restype = c_i32
    lib.sgm_stat_dim.argtypes = [c_i32, c_i32]
    lib.sgm_state_dim.restype = c_i32
    lib.sgm_state_dim.argtypes = [c_i32]
    lib.sgm_pf_workspace_bytes.restype = c_u64
    lib.sgm_pf_workspace_bytes.argtypes = [ctypes.POINTER(SgmPfDesc)]
    lib.sgm_pf_run.restype = c_i32
    lib.sgm_pf_run.argtypes = [ctypes.POINTER(SgmPfDesc), c_vp]
    lib.sgm_last_launch_count.restype = c_i64
    lib.sgm_sgld_items.restype = c_i32
    lib.sgm_sgld_items.argtypes = [ctypes.POINTER(SgmSgldDesc)]
    lib.sgm_sgld_max_steps.restype = c_i32
    lib.sgm_sgld_max_steps.argtypes = [ctypes.POINTER(SgmSgldDesc)]
    lib.sgm_sgld_workspace_bytes.restype = c_u64
    lib.sgm_sgld_workspace_bytes.argtypes = [ctypes.POINTER(SgmSgldDesc)]
    lib.sgm_sgld_run.restype = c_i32
    lib.sgm_sgld_run.argtypes = [ctypes.POINTER(SgmSgldDesc), c_vp]
    lib.sgm_selftest_math.restype = c_i32
    lib.sgm_selftest_math.argtypes = [c_i32, c_vp, c_vp, c_i64, c_vp]
    lib.sgm_ksd_imq.restype = c_i32
    lib.sgm_ksd_imq.argtypes = [c_vp, c_vp, c_i32, c_i32, c_f64, c_f64, c_vp, c_vp]
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().sgm_last_error().decode()
        raise ERR.get(code, RuntimeError)(msg or "sgmpf error {0}".format(code))
