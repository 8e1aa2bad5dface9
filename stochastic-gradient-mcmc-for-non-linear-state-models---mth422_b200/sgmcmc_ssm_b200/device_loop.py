"""Device-resident SG-MCMC loop: K iterations of C independent chains per host call (C-ABI `sgm_sgld_run`).

The reference advances one chain by one iteration per Python call: draw a window, run the particle filters, add the
prior gradient, draw the Langevin noise, update, project (sgmcmc_sampler.py:390-464, 529-567, 613-640, 650-656) -- for
the N ~ 10^3 particle counts of its own scripts the Python around the filter costs more than the filter.  Here the
whole iteration lives on the device; the host enqueues K iterations (optionally as a replayed CUDA graph) and reads
the parameters back when asked.

Randomness: `rng='philox'` draws windows, filter randoms and Langevin noise from counter-based device streams
(chain- and GPU-count-invariant).  `rng='injected'` is the parity mode: the host consumes the GLOBAL numpy legacy
stream exactly in the reference's order (window draws, per-item filter draws, noise normals) and ships the values, so
`np.random.seed(s); DeviceChains([sampler], ...).run(K)` lands on the reference's parameters after K iterations.
"""
import ctypes

import numpy as np
import torch

from . import _native as nat
from . import engine
from .particle_filters.buffered_smoother import _draw_injected

PARAM_SLOTS = {"svm": ("A", "LQinv_vec", "LRinv_vec"), "lgssm": ("A", "C", "LQinv_vec", "LRinv_vec"),
               "garch": ("log_mu", "logit_phi", "logit_lambduh", "LRinv_vec")}
HYPER_SLOTS = {"svm": ("mean_A", "var_col_A", "df_Qinv", "scale_Qinv", "df_Rinv", "scale_Rinv"),
               "lgssm": ("mean_A", "var_col_A", "mean_C", "var_col_C", "df_Qinv", "scale_Qinv", "df_Rinv", "scale_Rinv"),
               "garch": ("scale_mu", "shape_mu", "alpha_phi", "beta_phi", "alpha_lambduh", "beta_lambduh", "df_Rinv", "scale_Rinv")}
_METHODS = {"SGLD": "SGLD", "SGRLD": "SGRLD", "SGD": "SGD"}


def _f(a):
    return float(np.ravel(a)[0])


cluster_plan = engine.cluster_plan


def _supported_options(options):
    """project_parameters runs with the models' defaults on the device; custom projection options stay on the host."""
    return all(k == "partition_style" for k in (options or {}))


class DeviceChains(object):
    """C chains (samplers of one model sharing their observations) advanced on the device."""

    def __init__(self, samplers, method="SGLD", epsilon=0.1, subsequence_length=-1, buffer_length=0, minibatch_size=1,
                 num_sequences=None, pf="poyiadjis_N", N=None, num_samples=None, kernel=None, preconditioner=None,
                 dtype=None, rng=None, resample=None, variates=None, seed=None, lambduh=None, Ntilde=2,
                 trace_every=0, max_trace_rows=0, project=True, device=None, chain_id_base=0, kind="pf",
                 persistent=True, path="auto", **unused):
        if kind != "pf":
            raise NotImplementedError("the device loop covers the particle-filter gradient (kind='pf')")
        self.samplers = list(samplers)
        if not self.samplers:
            raise ValueError("no chains")
        s0 = self.samplers[0]
        models = {s.message_helper._model for s in self.samplers}
        if len(models) != 1:
            raise ValueError("all chains must share the model")
        self.model = models.pop()
        if method not in _METHODS:
            raise ValueError("Unrecognized iter_type {0}".format(method))
        if method == "SGRLD":
            from .models.lgssm import LGSSMPreconditioner
            if self.model != "lgssm" or not isinstance(preconditioner, LGSSMPreconditioner):
                raise NotImplementedError("No Default Preconditioner for {0} on the device".format(self.model))
        for s in self.samplers:
            if not _supported_options(getattr(s, "options", {})):
                raise NotImplementedError("custom project_parameters options are not available in the device loop")
            if getattr(s.parameters, "n", 1) != 1 or getattr(s.parameters, "m", 1) != 1:
                raise NotImplementedError("the CUDA particle-filter path covers n = m = 1")
        N = num_samples if N is None else N
        if N is None:
            raise TypeError("N (number of particles) must be given for kind='pf'")
        self.lib = lib = nat.load()
        self.device = device = engine._device(device)
        engine._state(device)
        self.method, self.N = method, int(N)
        self.rng = rng or engine.config.rng
        dtype = dtype or engine.config.dtype
        if resample is None:
            resample = "multinomial" if self.rng == "injected" else engine.config.resample
        K = s0.message_helper._get_kernel(kernel)
        for s in self.samplers:
            s.message_helper._get_kernel(kernel).set_parameters(s.parameters)          # |A| > 1 raises here
        C = self.C = len(self.samplers)
        self.slots, self.hslots = PARAM_SLOTS[self.model], HYPER_SLOTS[self.model]
        self.S, self.Bf, self.M = int(subsequence_length), int(buffer_length), int(minibatch_size)
        self.partition = (getattr(s0, "options", {}) or {}).get("partition_style") or "uniform"

        # ---- observations: one series, or a list of sequences (Seq samplers) ----
        obs = s0.observations
        self.is_seq = isinstance(obs, (list, tuple))
        seqs = [np.asarray(o, dtype=np.float64) for o in obs] if self.is_seq else [np.asarray(obs, dtype=np.float64)]
        for q in seqs:
            if q.ndim == 2 and q.shape[1] != 1:
                raise NotImplementedError("the CUDA particle-filter path covers m = 1 observations")
        self.seq_lens = np.array([q.shape[0] for q in seqs], dtype=np.int64)
        self.num_sequences = -1 if (num_sequences is None or num_sequences == -1 or not self.is_seq) else int(num_sequences)
        seq_off = np.concatenate([[0], np.cumsum(self.seq_lens)]).astype(np.int64)
        self.T_total = float(self.seq_lens.sum())

        with torch.cuda.device(device):
            self.d_obs = torch.from_numpy(np.concatenate([q.reshape(-1) for q in seqs])).to(device)
            self.d_seq_off = torch.from_numpy(seq_off).to(device)
            self.d_params = torch.zeros((C, nat.PARAM_STRIDE), dtype=torch.float64, device=device)
            self.d_hyper = torch.zeros((C, nat.HYPER_STRIDE), dtype=torch.float64, device=device)
            self.d_status = torch.zeros(C, dtype=torch.int32, device=device)
            self.d_offset = torch.zeros(1, dtype=torch.int64, device=device)
            self.d_iter = torch.zeros(1, dtype=torch.int64, device=device)
            self.trace_every = int(trace_every)
            self.trace_rows = int(max_trace_rows) + 1 if trace_every else 0
            self.d_trace = (torch.zeros((self.trace_rows, C, nat.PARAM_STRIDE), dtype=torch.float64, device=device)
                            if self.trace_rows else None)

            d = self.desc = nat.SgmSgldDesc()
            d.struct_bytes = ctypes.sizeof(nat.SgmSgldDesc)
            d.method = nat.STEP[method]
            d.n_chains, d.minibatch, d.n_seqs, d.num_sequences = C, self.M, len(seqs), self.num_sequences
            d.subsequence_length, d.buffer_length = self.S, self.Bf
            d.partition = nat.PARTITION[self.partition]
            d.project = int(bool(project))
            d.no_persistent = int(not persistent)
            d.trace_every, d.trace_rows = self.trace_every, self.trace_rows
            d.max_seq_len = int(self.seq_lens.max())
            d.epsilon, d.T_total = float(epsilon), self.T_total
            d.obs, d.seq_off = self.d_obs.data_ptr(), self.d_seq_off.data_ptr()
            d.params, d.hyper = self.d_params.data_ptr(), self.d_hyper.data_ptr()
            d.chain_status = self.d_status.data_ptr()
            d.trace = self.d_trace.data_ptr() if self.d_trace is not None else None
            d.offset_dev, d.iter_dev = self.d_offset.data_ptr(), self.d_iter.data_ptr()
            # x_0 prior: GARCH with no forward message uses the stationary variance of the CURRENT parameters
            helper = s0.message_helper
            if self.model == "garch" and helper.default_forward_message is None:
                d.prior_x0 = 1
                self.d_prior = None
            else:
                d.prior_x0 = 0
                pm = [s.message_helper._prior_moments(None, s.parameters) for s in self.samplers]
                self.d_prior = torch.tensor(pm, dtype=torch.float64, device=device).t().contiguous()
                d.prior_mean, d.prior_var = self.d_prior[0].data_ptr(), self.d_prior[1].data_ptr()

            pf_d = d.pf
            pf_d.struct_bytes = ctypes.sizeof(nat.SgmPfDesc)
            if pf not in nat.PF:
                raise ValueError("Unrecognized pf = {0}".format(pf))
            if lambduh is None:
                lambduh = 0.95
            if pf == "poyiadjis_N":
                lambduh = 1.0
            pf_d.model, pf_d.kernel, pf_d.pf, pf_d.dtype = nat.MODEL[K.model], nat.KERNEL[K.kernel], nat.PF[pf], nat.DTYPE[dtype]
            pf_d.rng_mode, pf_d.resample, pf_d.stat_kind = nat.RNG[self.rng], nat.RESAMPLE[resample], nat.STAT["score"]
            pf_d.n_particles, pf_d.Ntilde, pf_d.accept_reject = self.N, int(Ntilde), 1
            pf_d.max_accept_reject = pf_d.manual_sample_threshold = -1
            pf_d.variates = nat.VARIATES[variates or engine.config.variates]
            pf_d.path = nat.PATH[path]
            pf_d.lambduh = float(lambduh)
            if self.rng == "injected":
                seed, offset = 0, 0
            else:
                s_, offset = engine._next_seed_offset()
                seed = s_ if seed is None else seed
            pf_d.seed = int(seed) & (2 ** 64 - 1)
            self.d_offset.fill_(int(offset) << 20)            # leaves the low bits to the per-iteration increments
            B = self.B = int(lib.sgm_sgld_items(ctypes.byref(d)))
            if B < 0:
                nat.check(B)
            self.max_T = int(lib.sgm_sgld_max_steps(ctypes.byref(d)))
            self.ipc = B // C
            # one work item per chain and a shared-memory-sized particle system: the library runs all iterations of a
            # call inside one persistent kernel (no per-iteration launches to put into a graph)
            fast = pf == "poyiadjis_N" and self.rng == "philox" and resample in ("multinomial", "multinomial_sorted", "sorted")
            self.cluster = bool(persistent) and self.ipc == 1 and fast and path in ("auto", "cluster") and (
                cluster_plan(self.N, B, path == "cluster") is not None)
            self.persistent = bool(persistent) and self.ipc == 1 and pf in ("nemeth", "poyiadjis_N", "filter") and (
                self.N <= 2048 or self.cluster)
            pf_d.item_id_base = int(chain_id_base) * self.ipc
            if engine.config.two_streams:
                st = engine._state(device)
                if st.aux is None:
                    st.aux = (torch.cuda.Stream(device), torch.cuda.Event(), torch.cuda.Event())
                    for ev in st.aux[1:]:
                        ev.record()
                pf_d.aux_stream = st.aux[0].cuda_stream
                pf_d.ev_aux_fork, pf_d.ev_aux_join = st.aux[1].cuda_event, st.aux[2].cuda_event
            self._inj_keep = []
            if self.rng == "injected":                        # placeholders so that the size query validates
                dummy = torch.zeros(8, dtype=torch.float64, device=device)
                self._inj_keep.append(dummy)
                pf_d.inj_z0 = pf_d.inj_u = pf_d.inj_z = dummy.data_ptr()
            ws = int(lib.sgm_sgld_workspace_bytes(ctypes.byref(d)))
            if ws == 0:
                d.n_iters = 0
                nat.check(lib.sgm_sgld_run(ctypes.byref(d), None) or -1)
            self.d_ws = torch.empty(ws + 256, dtype=torch.uint8, device=device)
            d.workspace, d.workspace_bytes = engine._aligned_ptr(self.d_ws), ws
        self.launches = 0
        self.iterations = 0
        self._graphs = {}
        self.push_parameters()

    # ---- parameters <-> device ---------------------------------------------------------------------------------
    def push_parameters(self):
        P = np.zeros((self.C, nat.PARAM_STRIDE))
        H = np.zeros((self.C, nat.HYPER_STRIDE))
        for c, s in enumerate(self.samplers):
            for i, k in enumerate(self.slots):
                P[c, i] = _f(s.parameters.var_dict[k])
            for i, k in enumerate(self.hslots):
                H[c, i] = _f(s.prior.hyperparams[k])
        self.d_params.copy_(torch.from_numpy(P))
        self.d_hyper.copy_(torch.from_numpy(H))
        if self.d_trace is not None:
            self.d_trace[0].copy_(self.d_params)

    def pull_parameters(self, check=True):
        """Copy the chains' parameters back into the sampler objects (one D2H copy); flagged chains raise like the
        reference ("NaNs in gradient", sgmcmc_sampler.py:420-421)."""
        P = self.d_params.cpu().numpy()
        status = self.d_status.cpu().numpy()
        for c, s in enumerate(self.samplers):
            for i, k in enumerate(self.slots):
                s.parameters.var_dict[k] = np.full_like(np.asarray(s.parameters.var_dict[k], dtype=float), P[c, i])
        if check and np.any(status):
            raise ValueError("NaNs in gradient (chains {0})".format(np.nonzero(status)[0].tolist()))
        return [s.parameters for s in self.samplers]

    def trace(self):
        """(rows, C, n_params) parameters after 0, trace_every, 2 trace_every, ... iterations."""
        rows = min(self.trace_rows, self.iterations // self.trace_every + 1) if self.trace_every else 0
        return self.d_trace[:rows, :, :len(self.slots)].cpu().numpy()

    # ---- parity mode: consume the numpy stream like the reference -------------------------------------------
    def _draw_injected(self, K):
        """Per iteration and chain, in the reference's order: [sequence picks,] window starts, the filters' draws item
        by item, then the noise normals (sgmcmc_sampler.py:397-406, 1261-1277, 542-546)."""
        C, M, S, B, N, T = self.C, self.M, self.S, self.B, self.N, self.max_T
        nsel = self.ipc // M
        starts = np.zeros((K, B), dtype=np.int32)
        seqs = np.zeros((K, C, max(nsel, 1)), dtype=np.int32)
        noise = np.zeros((K, C, nat.PARAM_STRIDE))
        z0 = np.zeros((K, B, N)); u = np.zeros((K, B, T, N)); z = np.zeros((K, B, T, N))
        n_seqs = len(self.seq_lens)
        for k in range(K):
            for c in range(C):
                if self.num_sequences != -1:
                    pick = np.random.choice(np.arange(n_seqs), self.num_sequences, replace=False)
                else:
                    pick = np.arange(n_seqs)
                seqs[k, c, :len(pick)] = pick
                for si, q in enumerate(pick):
                    Tq = int(self.seq_lens[q])
                    b0 = c * self.ipc + si * M
                    T_list = []
                    for m in range(M):
                        if S != -1 and Tq - S > 0:
                            if self.partition == "strict":
                                r = int(np.random.choice(np.arange(0, Tq // S)))
                                start = r * S
                            else:
                                r = start = int(np.random.randint(0, Tq - S + 1))
                            end = start + S
                        else:
                            r, start, end = 0, 0, Tq
                        starts[k, b0 + m] = r
                        Bf = Tq if self.Bf == -1 else self.Bf
                        T_list.append(min(Tq, end + Bf) - max(0, start - Bf))
                    dr = _draw_injected(N, T_list)
                    for m in range(M):
                        z0[k, b0 + m] = dr["z0"][m]
                        u[k, b0 + m, :dr["u"].shape[1]] = dr["u"][m]
                        z[k, b0 + m, :dr["z"].shape[1]] = dr["z"][m]
                if self.method != "SGD":
                    for i in range(len(self.slots)):
                        noise[k, c, i] = np.random.normal(loc=0, size=(1,))[0]
        return starts, seqs, noise, z0, u, z

    # ---- run ------------------------------------------------------------------------------------------------
    def _enqueue(self, K):
        self.desc.n_iters = int(K)
        stream = torch.cuda.current_stream(self.device)
        nat.check(self.lib.sgm_sgld_run(ctypes.byref(self.desc), ctypes.c_void_p(stream.cuda_stream)))
        self.launches = int(self.lib.sgm_last_launch_count())

    def run(self, num_iters, graph=None, chunk=8):
        """Enqueue `num_iters` iterations on the current stream (asynchronous).  graph=True replays a captured CUDA
        graph of `chunk` iterations (the Philox counter lives in device memory, so every replay draws fresh randoms);
        default: graphs for device randoms when enabled in engine.config."""
        num_iters = int(num_iters)
        with torch.cuda.device(self.device):
            if self.rng == "injected":
                starts, seqs, noise, z0, u, z = self._draw_injected(num_iters)
                keep = [torch.from_numpy(x).to(self.device) for x in (starts, seqs, noise, z0, u, z)]
                self._inj_keep = keep
                d = self.desc
                d.inj_start, d.inj_seq, d.inj_noise = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
                d.pf.inj_z0, d.pf.inj_u, d.pf.inj_z = keep[3].data_ptr(), keep[4].data_ptr(), keep[5].data_ptr()
                self._enqueue(num_iters)
            else:
                if graph is None:
                    graph = engine.config.cuda_graphs and not self.persistent
                done = 0
                if graph and num_iters >= chunk:
                    g = self._graphs.get(chunk)
                    if g is None:
                        saved = (self.desc.pf.aux_stream, self.desc.pf.ev_aux_fork, self.desc.pf.ev_aux_join)
                        self.desc.pf.aux_stream = self.desc.pf.ev_aux_fork = self.desc.pf.ev_aux_join = None
                        self._enqueue(1)                          # warm-up outside the capture (module load)
                        done += 1
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, capture_error_mode="thread_local"):
                            self._enqueue(chunk)
                        self._graphs[chunk] = g
                        self._graph_launches = self.launches
                        self.desc.pf.aux_stream, self.desc.pf.ev_aux_fork, self.desc.pf.ev_aux_join = saved
                    while num_iters - done >= chunk:
                        g.replay()
                        done += chunk
                if num_iters - done > 0:
                    self._enqueue(num_iters - done)
        self.iterations += num_iters
        return self

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()
        return self
