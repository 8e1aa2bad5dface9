"""Host side of the C-ABI: packs a batch of buffered-subsequence work items, launches
`sgm_pf_run` on the current CUDA stream and reads back gradients / log-likelihoods.

PyTorch is used for device memory, pinned staging buffers and streams only.  One call moves ONE
packed host buffer to the device and ONE packed result buffer back.
"""
import ctypes
import os

import numpy as np
import torch

from . import _native as nat

_SEED = [None]
_CALLS = [0]


class Config(object):
    """Library-wide defaults; every entry can be overridden per call through **kwargs
    (the reference passes unknown kwargs through every layer, so the drop-in accepts them too)."""
    dtype = "f32"                    # particle arithmetic / storage: 'f32' or 'f64'
    rng = "philox"                   # 'philox' (device) or 'injected' (numpy legacy stream, parity mode)
    resample = "multinomial_sorted"  # 'multinomial' | 'multinomial_sorted' | 'systematic' | 'stratified'
    device = None                    # torch device; default = current CUDA device
    two_streams = True               # split big O(N) batches over two streams (hides the per-step header kernel)
    cuda_graphs = True               # replay a captured CUDA graph for small launch-bound batches (one wave of CTAs)
    device_loop = True               # sampler.fit(...) runs whole SG-MCMC iterations on the device when it can (device_loop.py)
    variates = "native"              # dtype 'f64' + device randoms: 'native' (53-bit variates) or 'f32' (the f32 path's variates, widened)


config = Config()


def set_seed(seed):
    """Seed of the device Philox streams (call counter restarts)."""
    _SEED[0] = int(seed) & (2 ** 64 - 1)
    _CALLS[0] = 0


def _next_seed_offset():
    """(seed, call counter) of the next device-random call.  An unseeded process seeds itself from the OS -- never from
    np.random: the global numpy stream belongs to the caller (the reference's samplers draw windows and SGLD noise from
    it, and parity runs replay it draw for draw)."""
    if _SEED[0] is None:
        _SEED[0] = int.from_bytes(os.urandom(8), "little")
    _CALLS[0] += 1
    return _SEED[0], _CALLS[0]


def skip_call():
    """Advance the call counter without launching (a rank whose shard of a distributed minibatch is empty), so the
    Philox call offsets stay aligned across ranks."""
    _next_seed_offset()


def _device(device=None):
    if device is None:
        device = config.device
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("sgmcmc_ssm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


class _DeviceState(object):
    """Grow-only buffers cached per device."""

    def __init__(self, device):
        self.device = device
        self.workspace = None
        self.pin_in = None
        self.dev_in = None
        self.pin_out = None
        self.dev_out = None
        self.checked = False
        self.generation = 0        # bumped by every download(): a PFResult whose staging buffer was reused says so
        self.aux = None            # (stream, fork event, join event) for the two-stream pipelining of big batches

    @staticmethod
    def _grow(t, nbytes, **kw):
        if t is None or t.numel() < nbytes:
            t = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, **kw)
        return t

    def buffers(self, ws_bytes, in_bytes, out_bytes):
        before = tuple(None if t is None else t.data_ptr() for t in (self.workspace, self.dev_in, self.dev_out))
        self._buffers(ws_bytes, in_bytes, out_bytes)
        after = tuple(t.data_ptr() for t in (self.workspace, self.dev_in, self.dev_out))
        if before != after:
            _GRAPHS.clear()            # captured graphs hold the old device addresses

    def _buffers(self, ws_bytes, in_bytes, out_bytes):
        self.workspace = self._grow(self.workspace, ws_bytes + 256, device=self.device)
        self.pin_in = self._grow(self.pin_in, in_bytes, pin_memory=True)
        self.dev_in = self._grow(self.dev_in, in_bytes + 256, device=self.device)
        self.pin_out = self._grow(self.pin_out, out_bytes, pin_memory=True)
        self.dev_out = self._grow(self.dev_out, out_bytes + 256, device=self.device)


_STATES = {}
_GRAPHS = {}        # graph key -> (torch.cuda.CUDAGraph, device offset tensor, launches per replay)


def _state(device):
    key = (device.type, device.index)
    if key not in _STATES:
        _STATES[key] = _DeviceState(device)
    st = _STATES[key]
    if not st.checked:
        with torch.cuda.device(device):
            nat.check(nat.load().sgm_device_check())
        st.checked = True
    return st


def cluster_plan(N, B, forced=False):
    """Mirror of cluster_plan() in csrc/cluster_impl.cuh: (cluster size, particles per CTA) when a batch of B items with N
    particles runs the thread-block-cluster kernel (only when forced with path='cluster': path='auto' prefers the
    shared-memory kernel up to N = 2048 and the cooperative tile kernel above), else None."""
    if N <= 256 or not forced:
        return None
    for nl in ((256, 512, 1024, 2048) if forced else (256,)):
        C = 2
        while C * nl < N:
            C *= 2
        if C <= 8 and B * C <= 148:
            return C, nl
    return None


def _align(x, a=256):
    return (x + a - 1) // a * a


def _aligned_ptr(t):
    return _align(t.data_ptr())


class PFItems(object):
    """A batch of independent work items (chain x sequence x subsequence)."""

    def __init__(self):
        self.obs, self.t1, self.tL, self.weights, self.theta, self.prior_mean, self.prior_var = [], [], [], [], [], [], []

    def add(self, observations, theta, t1=0, tL=None, weights=None, prior_mean=0.0, prior_var=1.0):
        obs = np.ascontiguousarray(np.asarray(observations, dtype=np.float64).reshape(-1))
        T = obs.shape[0]
        tL = T if tL is None else int(tL)
        if weights is not None:
            weights = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1))
            if weights.shape[0] < tL - int(t1):
                raise ValueError("weights shorter than the subsequence")
        th = np.zeros(nat.THETA_STRIDE)
        th[:len(theta)] = np.asarray(theta, dtype=np.float64)
        self.obs.append(obs); self.t1.append(int(t1)); self.tL.append(tL); self.weights.append(weights)
        self.theta.append(th); self.prior_mean.append(float(prior_mean)); self.prior_var.append(float(prior_var))
        return self

    def __len__(self):
        return len(self.obs)

    def pack(self):
        """Flatten to the arrays the C-ABI descriptor points at."""
        B = len(self.obs)
        T_buf = np.array([o.shape[0] for o in self.obs], dtype=np.int32)
        has_w = any(w is not None for w in self.weights)
        wts_off = np.full(B, -1, dtype=np.int64)
        wlen = 0
        if has_w:
            for b, w in enumerate(self.weights):
                if w is not None:
                    wts_off[b] = wlen
                    wlen += w.shape[0]
        return PackedItems(
            obs_flat=np.concatenate(self.obs) if B else np.zeros(0), T_buf=T_buf,
            t1=np.array(self.t1, dtype=np.int32), tL=np.array(self.tL, dtype=np.int32),
            wts_flat=np.concatenate([w for w in self.weights if w is not None]) if has_w else None, wts_off=wts_off,
            theta=np.stack(self.theta) if B else np.zeros((0, nat.THETA_STRIDE)),
            prior_mean=np.array(self.prior_mean, dtype=np.float64), prior_var=np.array(self.prior_var, dtype=np.float64))


class PackedItems(object):
    """The same batch as flat arrays (what a vectorised caller builds directly, skipping the per-item
    Python work of PFItems.add): obs_flat (sum T_buf,), T_buf / t1 / tL (B,), wts_flat + wts_off (-1 = all
    ones) or wts_flat None, theta (B, 12) or (12,) shared by all items, prior_mean / prior_var (B,) or scalars."""

    def __init__(self, obs_flat, T_buf, t1, tL, wts_flat, wts_off, theta, prior_mean, prior_var):
        self.T_buf = np.ascontiguousarray(T_buf, dtype=np.int32)
        B = self.T_buf.shape[0]
        self.obs_flat = np.ascontiguousarray(obs_flat, dtype=np.float64).reshape(-1)
        self.t1 = np.ascontiguousarray(t1, dtype=np.int32)
        self.tL = np.ascontiguousarray(tL, dtype=np.int32)
        self.wts_flat = None if wts_flat is None else np.ascontiguousarray(wts_flat, dtype=np.float64).reshape(-1)
        self.wts_off = (np.full(B, -1, dtype=np.int64) if wts_off is None else np.ascontiguousarray(wts_off, dtype=np.int64))
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            row = np.zeros(nat.THETA_STRIDE)
            row[:theta.shape[0]] = theta
            theta = np.broadcast_to(row, (B, nat.THETA_STRIDE))
        self.theta = theta
        self.prior_mean = np.broadcast_to(np.asarray(prior_mean, dtype=np.float64), (B,))
        self.prior_var = np.broadcast_to(np.asarray(prior_var, dtype=np.float64), (B,))
        if self.obs_flat.shape[0] != int(self.T_buf.sum()):
            raise ValueError("obs_flat does not match T_buf")

    def __len__(self):
        return self.T_buf.shape[0]

    def pack(self):
        return self

    @staticmethod
    def concat(parts):
        """One batch from several (e.g. one per chain); items keep their order."""
        parts = [p.pack() for p in parts]
        any_w = any(p.wts_flat is not None for p in parts)
        wts_flat, wts_off, base = [], [], 0
        for p in parts:
            if p.wts_flat is None:
                wts_off.append(np.full(len(p), -1, dtype=np.int64))
            else:
                wts_off.append(np.where(p.wts_off >= 0, p.wts_off + base, -1))
                wts_flat.append(p.wts_flat)
                base += p.wts_flat.shape[0]
        return PackedItems(np.concatenate([p.obs_flat for p in parts]), np.concatenate([p.T_buf for p in parts]),
                           np.concatenate([p.t1 for p in parts]), np.concatenate([p.tL for p in parts]),
                           np.concatenate(wts_flat) if any_w else None, np.concatenate(wts_off),
                           np.concatenate([p.theta for p in parts]), np.concatenate([p.prior_mean for p in parts]),
                           np.concatenate([p.prior_var for p in parts]))

    def slice(self, lo, hi):
        """Items [lo, hi) (the shard of one rank)."""
        obs_off = np.concatenate([[0], np.cumsum(self.T_buf)])
        wts_flat, wts_off = self.wts_flat, self.wts_off[lo:hi]
        if wts_flat is not None and hi > lo:
            # a weighted item's slice ends where the next weighted item's begins (or at the end)
            used = wts_off[wts_off >= 0]
            if used.size:
                nxt = self.wts_off[hi:]
                nxt = nxt[nxt >= 0]
                w_lo, w_hi = int(used.min()), int(nxt.min()) if nxt.size else wts_flat.shape[0]
                wts_flat = wts_flat[w_lo:w_hi]
                wts_off = np.where(wts_off >= 0, wts_off - w_lo, -1)
            else:
                wts_flat = None
        return PackedItems(self.obs_flat[obs_off[lo]:obs_off[hi]], self.T_buf[lo:hi], self.t1[lo:hi], self.tL[lo:hi],
                           wts_flat, wts_off, self.theta[lo:hi], self.prior_mean[lo:hi], self.prior_var[lo:hi])


class PFResult(object):
    """Handle of an asynchronous sgm_pf_run; `.wait()` synchronises and parses the packed result."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self._done = False

    def wait(self, check=True):
        if self._done:
            return self
        if self._state.generation != self._generation:
            raise RuntimeError("this PFResult was not waited for before the next run_pf on the same device reused its "
                               "staging buffer (one call may be outstanding per device)")
        self._event.synchronize()
        B = self.B
        out = self._pin_out.numpy()
        gb = 8 * self.grad_slots                         # bytes of one row of `grad`
        rows = out[:B * gb].view(np.float64).reshape(B, self.grad_slots)
        self.grad = rows[:, :max(self.p, 0)].copy()
        self.diag = rows[:, 6:8].copy() if self.grad_slots == 8 else None      # PaRIS (device randoms): proposals, exact fallbacks
        self.loglik = out[B * gb:B * (gb + 8)].view(np.float64).copy()
        self.status = out[B * (gb + 8):B * (gb + 12)].view(np.int32).copy()
        self._done = True
        if check:
            bad = self.status & (nat.STATUS_NAN_WEIGHT | nat.STATUS_ZERO_WEIGHT)
            if np.any(bad):
                raise ValueError("probabilities contain NaN (degenerate particle weights in item(s) {0})".format(
                    np.nonzero(bad)[0].tolist()))
        return self

    def tensor(self, name):
        """Optional device outputs (x, lw, stats, anc, trace_x, trace_lw, J) as torch tensors."""
        self.wait()
        return self._extra[name]


class PreparedPF(object):
    """A packed batch: host staging buffer filled, descriptor built.  `upload()` issues the single
    H2D copy, `launch()` calls sgm_pf_run on the current stream (re-launchable: the bench times
    launches with inputs resident in HBM), `download()` issues the single D2H copy."""

    def __init__(self, model, kernel, pf, items, N, dtype=None, rng=None, resample=None, stat_kind="score",
                 lambduh=None, Ntilde=2, accept_reject=True, max_accept_reject=None,
                 manual_sample_threshold=None, seed=None, offset=None, item_id_base=0, injected=None,
                 want=(), device=None, n2_mode="auto", num_steps_ahead=5, per_horizon=False, variates=None, path="auto"):
        lib = self.lib = nat.load()
        device = self.device = _device(device)
        st = self.st = _state(device)
        dtype = dtype or config.dtype
        rng = rng or config.rng
        if resample is None:
            # recorded uniforms come in the reference's (iid) order: only the generic search is valid for them
            resample = "multinomial" if rng == "injected" else config.resample
        if pf not in nat.PF:
            raise ValueError("Unrecognized pf = {0}".format(pf))
        if lambduh is None:
            lambduh = 0.95                                           # pf.py:140
        if pf == "poyiadjis_N":
            lambduh = 1.0                                            # buffered_smoother.py:175-180
        B = self.B = len(items)
        if B == 0:
            raise ValueError("empty batch")
        N = self.N = int(N)
        model_id, kernel_id = nat.MODEL[model], nat.KERNEL[kernel]
        self.p = lib.sgm_stat_dim(model_id, nat.STAT[stat_kind])
        if stat_kind == "pred":
            if not 0 <= int(num_steps_ahead) <= nat.PRED_MAX_STEPS:
                raise NotImplementedError("num_steps_ahead must be in [0, %d] on the CUDA path" % nat.PRED_MAX_STEPS)
            self.p = int(num_steps_ahead) + 1
        self.n = n = lib.sgm_state_dim(model_id)
        NPrec = lib.sgm_stat_dim(model_id, 0)
        pk = items.pack()
        T_buf = self.T_buf = pk.T_buf
        max_T = self.max_T = int(T_buf.max())
        t_real = torch.float64 if dtype == "f64" else torch.float32

        # ---- layout of the packed host buffer ----------------------------------------------------
        obs_off = np.zeros(B, dtype=np.int64)
        obs_off[1:] = np.cumsum(T_buf[:-1])
        n_obs = int(pk.obs_flat.shape[0])
        has_w = pk.wts_flat is not None
        wts_off = pk.wts_off
        wlen = pk.wts_flat.shape[0] if has_w else 0
        # fixed-size sections first and the variable-length ones padded to their maxima, so that batches of the same
        # shape (B, max_T) always land on the same device addresses -- the CUDA-graph cache key holds these pointers
        sections = [("theta", B * nat.THETA_STRIDE * 8), ("prior_mean", B * 8), ("prior_var", B * 8), ("obs_off", B * 8),
                    ("wts_off", B * 8), ("T_buf", B * 4), ("t1", B * 4), ("tL", B * 4),
                    ("obs", max(n_obs, B * max_T) * 8), ("wts", max(wlen, B * max_T, 1) * 8)]
        offs, tot = {}, 0
        for name, nb in sections:
            offs[name] = tot
            tot = _align(tot + nb, 16)
        gs = self.grad_slots = nat.PRED_SLOTS if stat_kind == "pred" else 8
        self.in_bytes, self.out_bytes = tot, B * (8 * gs + 12) + 64

        desc = self.desc = nat.SgmPfDesc()
        desc.struct_bytes = ctypes.sizeof(nat.SgmPfDesc)
        desc.model, desc.kernel, desc.pf, desc.dtype = model_id, kernel_id, nat.PF[pf], nat.DTYPE[dtype]
        desc.rng_mode, desc.resample, desc.stat_kind = nat.RNG[rng], nat.RESAMPLE[resample], nat.STAT[stat_kind]
        desc.n_items, desc.n_particles, desc.max_T = B, N, max_T
        desc.Ntilde, desc.accept_reject = int(Ntilde), int(bool(accept_reject))
        desc.max_accept_reject = -1 if max_accept_reject is None else int(max_accept_reject)
        desc.manual_sample_threshold = -1 if manual_sample_threshold is None else int(manual_sample_threshold)
        desc.item_id_base = int(item_id_base)
        desc.n2_mode = nat.N2_MODE[n2_mode]
        desc.variates = nat.VARIATES[variates or config.variates]
        desc.path = nat.PATH[path]
        desc.pred_steps_ahead, desc.pred_per_horizon = int(num_steps_ahead), int(bool(per_horizon))
        desc.lambduh = float(lambduh)
        if rng == "injected":
            seed, offset = 0, 0                          # no device randoms: the process-wide Philox state is untouched
        elif seed is None or offset is None:
            s_, o_ = _next_seed_offset()
            seed = s_ if seed is None else seed
            offset = o_ if offset is None else offset
        desc.seed, desc.offset = int(seed) & (2 ** 64 - 1), int(offset)

        with torch.cuda.device(device):
            extra = self.extra = {}

            def opt(name, shape, tdtype):
                if name in want:
                    extra[name] = torch.empty(shape, dtype=tdtype, device=device)
                    return extra[name].data_ptr()
                return None

            desc.out_x = opt("x", (B, N, n), t_real)
            desc.out_lw = opt("lw", (B, N), t_real)
            desc.out_stats = opt("stats", (B, N, NPrec), t_real)       # record width
            desc.trace_anc = opt("anc", (B, max_T, N), torch.int32)
            desc.trace_x = opt("trace_x", (B, max_T + 1, N, n), t_real)
            desc.trace_lw = opt("trace_lw", (B, max_T + 1, N), t_real)
            desc.trace_J = opt("J", (B, max_T, N, int(Ntilde)), torch.int32)

            keep = self.keep = []
            if rng == "injected":
                if injected is None:
                    raise ValueError("rng='injected' needs the recorded randoms")

                def dev64(a, shape):
                    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(shape)
                    tns = torch.from_numpy(a).to(device)
                    keep.append(tns)
                    return tns.data_ptr()

                desc.inj_z0 = dev64(injected["z0"], (B, N))
                u_inj = np.asarray(injected["u"], dtype=np.float64).reshape(B, max_T, N)
                if resample != "multinomial" and N > 1 and np.any(np.diff(u_inj, axis=-1) < 0):
                    # the streaming search stages the CDF window of a warp tile from its first and last target only
                    raise ValueError("rng='injected' with resample='{0}' needs ascending uniforms per step; "
                                     "use resample='multinomial' for uniforms in the reference's order".format(resample))
                desc.inj_u = dev64(u_inj, (B, max_T, N))
                desc.inj_z = dev64(injected["z"], (B, max_T, N))
                if injected.get("zp") is not None:       # predictive-statistic normals, (B, max_T, PRED_SLOTS, N)
                    desc.inj_pred = dev64(injected["zp"], (B, max_T, nat.PRED_SLOTS, N))
                if injected.get("extra") is not None:
                    flat = np.concatenate([np.asarray(e, dtype=np.float64).ravel() for e in injected["extra"]] + [np.zeros(1)])
                    lens = np.array([np.asarray(e).size for e in injected["extra"]], dtype=np.int64)
                    eoff = np.zeros(B * max_T, dtype=np.int64)
                    eoff[1:] = np.cumsum(lens)[:-1]
                    desc.inj_extra = dev64(flat, (-1,))
                    toff = torch.from_numpy(eoff).to(device)
                    keep.append(toff)
                    desc.inj_extra_off = toff.data_ptr()

            # workspace size depends on the scalar fields and on which optional outputs are set
            desc.obs = desc.obs_off = desc.T_buf = desc.t1 = desc.tL = desc.theta = desc.prior_mean = desc.prior_var = 1
            desc.grad = desc.loglik = desc.status = 1
            ws_bytes = self.ws_bytes = int(lib.sgm_pf_workspace_bytes(ctypes.byref(desc)))
            if ws_bytes == 0:
                nat.check(lib.sgm_pf_run(ctypes.byref(desc), None) or -1)
            st.buffers(ws_bytes, self.in_bytes, self.out_bytes)

            host = st.pin_in.numpy()

            def view(name, dt, count):
                return host[offs[name]:offs[name] + count * np.dtype(dt).itemsize].view(dt)

            view("obs", np.float64, n_obs)[:] = pk.obs_flat
            if has_w:
                view("wts", np.float64, wlen)[:] = pk.wts_flat
            view("theta", np.float64, B * nat.THETA_STRIDE).reshape(B, nat.THETA_STRIDE)[:] = pk.theta
            view("prior_mean", np.float64, B)[:] = pk.prior_mean
            view("prior_var", np.float64, B)[:] = pk.prior_var
            view("obs_off", np.int64, B)[:] = obs_off
            view("wts_off", np.int64, B)[:] = wts_off
            view("T_buf", np.int32, B)[:] = T_buf
            view("t1", np.int32, B)[:] = pk.t1
            view("tL", np.int32, B)[:] = pk.tL

            base_in = self.base_in = _aligned_ptr(st.dev_in)
            for name in ("obs", "theta", "prior_mean", "prior_var", "obs_off", "wts_off", "T_buf", "t1", "tL"):
                setattr(desc, name, base_in + offs[name])
            desc.step_weights = (base_in + offs["wts"]) if has_w else None
            base_out = self.base_out = _aligned_ptr(st.dev_out)
            desc.grad, desc.loglik, desc.status = base_out, base_out + B * 8 * gs, base_out + B * (8 * gs + 8)
            desc.workspace = _aligned_ptr(st.workspace)
            desc.workspace_bytes = ws_bytes
            if config.two_streams:
                if st.aux is None:
                    st.aux = (torch.cuda.Stream(device), torch.cuda.Event(), torch.cuda.Event())
                    for ev in st.aux[1:]:
                        ev.record()                      # creates the handles
                desc.aux_stream = st.aux[0].cuda_stream
                desc.ev_aux_fork, desc.ev_aux_join = st.aux[1].cuda_event, st.aux[2].cuda_event
        self.launches = 0
        self.particle_steps = int(N) * int(T_buf.sum())

    def upload(self):
        st = self.st
        with torch.cuda.device(self.device):
            sh = self.base_in - st.dev_in.data_ptr()
            st.dev_in[sh:sh + self.in_bytes].copy_(st.pin_in[:self.in_bytes], non_blocking=True)
        return self

    def launch(self, offset=None, step_events=None):
        """sgm_pf_run on the current stream.  step_events = (begin, end) torch events recorded around
        the step-kernel launches (they must have been recorded once so their handles exist)."""
        if offset is not None:
            self.desc.offset = int(offset)
        if step_events is not None:
            self.desc.ev_steps_begin, self.desc.ev_steps_end = step_events[0].cuda_event, step_events[1].cuda_event
        else:
            self.desc.ev_steps_begin = self.desc.ev_steps_end = None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            nat.check(self.lib.sgm_pf_run(ctypes.byref(self.desc), ctypes.c_void_p(stream.cuda_stream)))
        self.launches = int(self.lib.sgm_last_launch_count())
        return self

    # ---- CUDA-graph replay for launch-bound batches ----------------------------------------------------------
    def graph_eligible(self):
        """2 * max_T + 2 (+ backward kernels) tiny launches per call: worth a graph when the whole batch is at most one wave of
        CTAs and the call is not a single launch anyway (shared-memory / cluster / cooperative kernels of the O(N) smoothers).  Device randoms only (a graph bakes the
        kernel arguments in; the Philox call offset is then read from device memory, sgm_pf_desc.offset_dev)."""
        d = self.desc
        if not config.cuda_graphs or d.rng_mode != nat.RNG["philox"] or self.extra or self.N <= 2048:
            return False
        if d.pf in (nat.PF["nemeth"], nat.PF["filter"]) and d.stat_kind != nat.STAT["pred"] and self.N <= 65536:
            # the O(N) smoothers run such a batch in ONE launch (cluster kernel or the cooperative form of the tile kernels,
            # csrc/coop_kernels.cuh): nothing to put into a graph
            return False
        return self.B * ((self.N + 2047) // 2048) <= 148

    def _graph_key(self):
        d = self.desc
        skip = ("offset", "offset_dev", "ev_steps_begin", "ev_steps_end", "aux_stream", "ev_aux_fork", "ev_aux_join")
        return (self.device.index,) + tuple(getattr(d, name) for name, _ in d._fields_ if name not in skip)

    def launch_graph(self, offset=None):
        """Same work as launch(), replayed from a cached CUDA graph (captured on first use of this exact shape)."""
        if offset is not None:
            self.desc.offset = int(offset)
        key = self._graph_key()
        with torch.cuda.device(self.device):
            entry = _GRAPHS.get(key)
            if entry is None:
                off_t = torch.zeros(1, dtype=torch.int64, device=self.device)
                self.desc.offset_dev = off_t.data_ptr()
                self.desc.aux_stream = self.desc.ev_aux_fork = self.desc.ev_aux_join = None
                self.desc.ev_steps_begin = self.desc.ev_steps_end = None
                self.launch()                                   # warm-up outside the capture (module load, checks)
                graph = torch.cuda.CUDAGraph()
                # thread_local: other threads of the process (e.g. the NCCL watchdog under torchrun) may touch CUDA
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    self.launch()
                if len(_GRAPHS) >= 32:
                    _GRAPHS.pop(next(iter(_GRAPHS)))
                entry = _GRAPHS[key] = (graph, off_t, self.launches)
            graph, off_t, self.launches = entry
            off_t.fill_(int(self.desc.offset) & (2 ** 63 - 1))
            graph.replay()
        return self

    def download(self):
        st = self.st
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            so = self.base_out - st.dev_out.data_ptr()
            st.pin_out[:self.out_bytes].copy_(st.dev_out[so:so + self.out_bytes], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            st.generation += 1
        return PFResult(_state=st, _generation=st.generation, seed=int(self.desc.seed), offset=int(self.desc.offset), B=self.B, N=self.N, p=self.p, n=self.n, max_T=self.max_T, T_buf=self.T_buf, _event=ev,
                        _pin_out=st.pin_out, _extra=self.extra, _keep=self.keep, launches=self.launches,
                        h2d_bytes=self.in_bytes, d2h_bytes=self.out_bytes, particle_steps=self.particle_steps,
                        ws_bytes=self.ws_bytes, grad_slots=self.grad_slots)


def run_pf(model, kernel, pf, items, N, sync=True, check=True, while_running=None, **kwargs):
    """Run the buffered particle filter / smoother for a batch of work items on the GPU: pack, one H2D
    copy, the whole t-loop through the C-ABI, one D2H copy.

    Mirrors particle_filters/buffered_smoother.py:156-199 (`pf` dispatch) for a whole batch.
    Returns a PFResult with .grad (B, p), .loglik (B,), .status (B,).  One call may be outstanding per
    device (staging buffers are shared).  `while_running`: host work to do between the launch and the wait
    (the samplers evaluate the prior gradient there)."""
    prep = PreparedPF(model, kernel, pf, items, N, **kwargs).upload()
    res = (prep.launch_graph() if prep.graph_eligible() else prep.launch()).download()
    if while_running is not None:
        while_running()
    if sync:
        res.wait(check=check)
    return res


def run_pf_sum(model, kernel, pf, items, N, allreduce=False, while_running=None, **kwargs):
    """Sum over the items of a batch of their gradient estimates, reduced ON THE DEVICE (and, with allreduce=True, over
    the ranks of the process group by ONE in-place NCCL all-reduce of the same device buffer), then one small D2H copy:
    the minibatch path of the samplers (sgmcmc_sampler.py:411-418).  `items` may be None on a rank whose shard is empty.
    Returns (sums (p,) float64, info dict)."""
    import torch.distributed as dist
    device = _device(kwargs.get("device"))
    info = dict(particle_steps=0, h2d_bytes=0, d2h_bytes=0, launches=0)
    with torch.cuda.device(device):
        if items is not None and len(items) > 0:
            prep = PreparedPF(model, kernel, pf, items, N, **kwargs).upload()
            (prep.launch_graph() if prep.graph_eligible() else prep.launch())
            st = prep.st
            so = prep.base_out - st.dev_out.data_ptr()
            B, gs = prep.B, prep.grad_slots
            g = st.dev_out[so:so + B * 8 * gs].view(torch.float64).view(B, gs)
            status = st.dev_out[so + B * (8 * gs + 8):so + B * (8 * gs + 12)].view(torch.int32)
            buf = torch.empty(gs + 1, dtype=torch.float64, device=device)
            torch.sum(g, dim=0, out=buf[:gs])
            buf[gs] = (status & (nat.STATUS_NAN_WEIGHT | nat.STATUS_ZERO_WEIGHT)).ne(0).sum().to(torch.float64)
            info.update(particle_steps=prep.particle_steps, h2d_bytes=prep.in_bytes, launches=prep.launches + 2, p=prep.p)
        else:
            skip_call()
            gs = nat.PRED_SLOTS if kwargs.get("stat_kind") == "pred" else 8
            buf = torch.zeros(gs + 1, dtype=torch.float64, device=device)
        if allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            info["launches"] += 1
        info["allreduced"] = bool(allreduce)
        if while_running is not None:
            while_running()
        host = buf.cpu().numpy()                      # the only synchronisation and D2H copy of the call
        info["d2h_bytes"] = int(host.nbytes)
    if host[gs] != 0:
        raise ValueError("probabilities contain NaN (degenerate particle weights in {0} item(s))".format(int(host[gs])))
    return host[:gs], info
