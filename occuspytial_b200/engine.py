"""Thin Python owner of one device engine (one GPU's shard of lock-step chains).

PyTorch is used for exactly three things here: allocating device memory (the engine's workspace,
the problem arrays, the record buffers), host<->device copies, and the CUDA stream.  All compute
goes through the C ABI in ``_lib`` (include/occu_b200.h).
"""
import ctypes as C
import threading

import numpy as np

from . import _lib

_MODEL_IDS = {'logit_rsr': _lib.LOGIT_RSR, 'probit_rsr': _lib.PROBIT_RSR, 'logit_icar': _lib.LOGIT_ICAR}

# one pinned staging buffer per device and concurrent engine slot, kept for the life of the process: page-locking memory is a slow driver call
# (tens of milliseconds when several processes do it at once), not something to repeat in every sample() call
_PINNED = {}


def _pinned_buffer(torch, device, slot, count, kind='h2d'):
    """``count`` float64 elements of page-locked host memory, cached per (device, slot, direction)."""
    key = (str(device), slot, kind)   # engines that run at the same time (one per host thread) use different slots
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < count:
        grow = 0 if buf is None else 2 * buf.numel()
        buf = torch.empty(max(count, grow, 1 << 16), dtype=torch.float64, pin_memory=True)
        _PINNED[key] = buf
    return buf[:count]


_D2H_STAGING_MAX = 1 << 27   # results above 128 MiB take the plain pageable copy: page-locking (and keeping) that much host memory costs more than it saves on a run that long


# ... and one CUDA stream per device and slot: torch's caching allocator keeps freed blocks per stream, so an engine that runs on
# the stream of its predecessor gets the staging / record / summary buffers of the previous sample() call back instead of
# going through cudaMalloc again (a fresh stream per engine did, on every call)
_STREAMS = {}


def _engine_stream(torch, device, slot):
    key = (str(device), slot)
    st = _STREAMS.get(key)
    if st is None:
        st = _STREAMS[key] = torch.cuda.Stream(device=device)
    return st


# Slots are handed out per device under a lock: two engines alive at the same time -- the per-device threads of one
# sample() call, or sample() calls from different host threads -- never share a stream or a staging buffer.
_SLOT_LOCK = threading.Lock()
_BUSY_SLOTS = {}


def _acquire_slot(device):
    key = str(device)
    with _SLOT_LOCK:
        busy = _BUSY_SLOTS.setdefault(key, set())
        slot = 0
        while slot in busy:
            slot += 1
        busy.add(slot)
    return slot


def _release_slot(device, slot):
    with _SLOT_LOCK:
        _BUSY_SLOTS.get(str(device), set()).discard(slot)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.EngineError('occuspytial_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch


class DeviceEngine:
    """Engine for ``chains`` lock-step chains of one model on one CUDA device.

    Parameters
    ----------
    model : {'logit_rsr', 'probit_rsr', 'logit_icar'}
    problem : dict
        Host (NumPy) arrays: X (n, p), W_rows (V, r), y_rows (V,), visit_ptr (n + 1,), site_class (n,),
        and for RSR models K (n, q), Qrsr (q, q); for ICAR Qband (n, bw + 1); hyper-parameters
        tau_rate, tau_shape, a_mu, a_prec, b_mu, b_prec.
    chains, chain_offset, seed : int
        Number of chains on this device, global id of the first one, Philox key.
    """

    def __init__(self, model, problem, chains, seed, chain_offset=0, device=None, debug=False, pg_fp64=False,
                 gram_int8=False, workspace=None, slot=None):
        import time as _time
        t_begin = _time.perf_counter()
        torch = _torch()
        self.lib = _lib.load()
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.model = model
        self.chains = int(chains)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.slot = None   # assigned below, once the device is known (``slot=`` is accepted for compatibility and ignored)
        X = np.ascontiguousarray(problem['X'], dtype=np.float64)
        self.n, self.p = X.shape
        W_rows = np.ascontiguousarray(problem['W_rows'], dtype=np.float64)
        self.V, self.r = W_rows.shape
        self.q = int(problem['K'].shape[1]) if problem.get('K') is not None else 0
        self.eta_dim = self.q if self.q else self.n
        self.icar_bw = int(problem.get('icar_bw', 0))
        with torch.cuda.device(self.device):
            self.slot = _acquire_slot(self.device)
            self.stream = _engine_stream(torch, self.device, self.slot)
            dev = {}

            def up(name, arr, dtype):
                a = np.ascontiguousarray(arr, dtype=dtype)
                self.h2d_bytes += a.nbytes
                t = torch.from_numpy(a).to(self.device, non_blocking=False)
                dev[name] = t
                return t.data_ptr()

            P = _lib.Problem()
            P.abi_version = _lib.ABI_VERSION
            P.model = _MODEL_IDS[model]
            P.n, P.p, P.r, P.q, P.V = self.n, self.p, self.r, self.q, self.V
            P.chains, P.chain_offset = self.chains, int(chain_offset)
            P.debug, P.pg_fp64 = int(bool(debug)), int(bool(pg_fp64))
            P.gram_int8 = int(bool(gram_int8))
            self.gram_int8 = bool(gram_int8) and model == 'logit_rsr'
            P.icar_bw = int(problem.get('icar_bw', 0))
            P.seed = int(seed) & (2**64 - 1)
            P.X = up('X', X, np.float64)
            P.K = up('K', problem['K'], np.float64) if self.q else None
            P.Qrsr = up('Qrsr', problem['Qrsr'], np.float64) if self.q else None
            P.Qband = up('Qband', problem['Qband'], np.float64) if problem.get('Qband') is not None else None
            P.W = up('W', W_rows, np.float64)
            P.y = up('y', problem['y_rows'], np.uint8)
            P.visit_ptr = up('visit_ptr', problem['visit_ptr'], np.int32)
            P.site_class = up('site_class', problem['site_class'], np.uint8)
            P.tau_rate, P.tau_shape = float(problem['tau_rate']), float(problem['tau_shape'])
            self._host = [np.ascontiguousarray(problem[k], dtype=np.float64) for k in ('a_mu', 'a_prec', 'b_mu', 'b_prec')]
            P.a_mu, P.a_prec, P.b_mu, P.b_prec = (a.ctypes.data for a in self._host)
            self._problem = P
            t_upload = _time.perf_counter()
            nbytes = self.lib.occu_workspace_bytes(C.byref(P))
            if nbytes == 0:
                _lib.check(-1, 'occu_workspace_bytes')
            # The workspace is caller-provided memory (include/occu_b200.h): a sampler hands the block of its previous
            # run back in (`workspace=`) so that repeated sample() calls do not go through cudaMalloc again.  A fresh one
            # is allocated in the context of the engine's stream, the only stream that touches it.
            if workspace is not None and workspace.device == self.device and workspace.numel() >= nbytes + 256:
                self.workspace = workspace
            else:
                with torch.cuda.stream(self.stream):
                    self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            t_alloc = _time.perf_counter()
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            handle = C.c_void_p()
            _lib.check(self.lib.occu_create(C.byref(P), base, nbytes, self.stream.cuda_stream, C.byref(handle)),
                       'occu_create')
            self.handle = handle
            self.stream.synchronize()
            del dev  # the engine copied the problem into its workspace
            t_end = _time.perf_counter()
            self.setup_seconds = {'upload': t_upload - t_begin, 'alloc': t_alloc - t_upload, 'create': t_end - t_alloc}

    # ---- state ---------------------------------------------------------------------------------
    def _field_shape(self, field):
        return {_lib.F_ALPHA: (self.r,), _lib.F_BETA: (self.p,), _lib.F_ETA: (self.eta_dim,), _lib.F_TAU: (),
                _lib.F_Z: (self.n,), _lib.F_OMEGA: (self.n,), _lib.F_PSI: (self.n,), _lib.F_EPS: (self.n,),
                _lib.F_ZPROB: (self.n,)}[field]

    def set_field(self, field, values):
        torch = self.torch
        dtype = np.uint8 if field == _lib.F_Z else np.float64
        arr = np.array(np.broadcast_to(np.asarray(values, dtype=dtype), (self.chains,) + self._field_shape(field)),
                       order='C', copy=True)
        self.h2d_bytes += arr.nbytes
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            t = torch.from_numpy(arr).to(self.device)
            _lib.check(self.lib.occu_set_field(self.handle, field, t.data_ptr(), self.stream.cuda_stream), 'occu_set_field')
            self.stream.synchronize()

    def get_field(self, field, chain=None):
        """Chain-major host copy of a state field, (chains, dim); ``chain=k`` copies only that chain, (1, dim)."""
        torch = self.torch
        dtype = torch.uint8 if field == _lib.F_Z else torch.float64
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if chain is None:
                t = torch.empty((self.chains,) + self._field_shape(field), dtype=dtype, device=self.device)
                _lib.check(self.lib.occu_get_field(self.handle, field, t.data_ptr(), self.stream.cuda_stream), 'occu_get_field')
            else:
                t = torch.empty((1,) + self._field_shape(field), dtype=dtype, device=self.device)
                _lib.check(self.lib.occu_get_field_chain(self.handle, field, int(chain), t.data_ptr(), self.stream.cuda_stream),
                           'occu_get_field_chain')
            self.stream.synchronize()
            out = t.cpu().numpy()
        self.d2h_bytes += out.nbytes
        return out

    def set_start(self, start):
        """start: dict of chain-major arrays alpha (C, r), beta (C, p), tau (C,), eta (C, q|n)[, eps (C, n)].

        The fields are packed into one pinned host buffer and cross PCIe in a single copy; the engine scatters them
        into its chain-minor state from device pointers into that block."""
        torch = self.torch
        fields = [(_lib.F_ALPHA, 'alpha'), (_lib.F_BETA, 'beta'), (_lib.F_TAU, 'tau'), (_lib.F_ETA, 'eta')]
        if self.model == 'probit_rsr':
            fields.append((_lib.F_EPS, 'eps'))
        parts, offsets, off = [], [], 0
        for field, name in fields:
            arr = np.broadcast_to(np.asarray(start[name], dtype=np.float64), (self.chains,) + self._field_shape(field))
            parts.append(arr)
            offsets.append(off)
            off += arr.size
        import time as _time
        t0 = _time.perf_counter()
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            host = _pinned_buffer(torch, self.device, self.slot, off)
            t1 = _time.perf_counter()
            hv = host.numpy()
            for arr, o in zip(parts, offsets):
                hv[o:o + arr.size] = arr.reshape(-1)
            t1a = _time.perf_counter()
            dev = host.to(self.device, non_blocking=True)
            t1b = _time.perf_counter()
            self.h2d_bytes += off * 8
            _lib.check(self.lib.occu_reset(self.handle, self.stream.cuda_stream), 'occu_reset')
            t1c = _time.perf_counter()
            for (field, _), o in zip(fields, offsets):
                _lib.check(self.lib.occu_set_field(self.handle, field, dev.data_ptr() + 8 * o, self.stream.cuda_stream),
                           'occu_set_field')
            t2 = _time.perf_counter()
            self.stream.synchronize()   # the staging buffer is free again after this
        self.setup_seconds.update(start_pin=t1 - t0, start_pack=t1a - t1, start_h2d=t1b - t1a, start_reset=t1c - t1b,
                                  start_scatter=t2 - t1c, start_sync=_time.perf_counter() - t2)

    def draw_start(self, seed, eta_scale=1.0, centre_eta=False):
        """Default start values drawn on the device (include/occu_b200.h: occu_draw_start); nothing crosses PCIe."""
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.occu_draw_start(self.handle, int(seed) & (2**64 - 1), float(eta_scale), int(bool(centre_eta)),
                                                self.stream.cuda_stream), 'occu_draw_start')

    # ---- sampling ------------------------------------------------------------------------------
    def alloc_record(self, capacity):
        torch = self.torch
        # the zero fill runs on the engine's stream, the same one the sweeps record on (a fill on the default stream
        # could land after the first recorded draws: torch streams do not synchronise with each other)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            return {
                'alpha': torch.zeros((self.chains, capacity, self.r), dtype=torch.float64, device=self.device),
                'beta': torch.zeros((self.chains, capacity, self.p), dtype=torch.float64, device=self.device),
                'tau': torch.zeros((self.chains, capacity), dtype=torch.float64, device=self.device),
            }

    def sweep(self, n_iters, record=None, rec_start=0):
        """Advance every chain by ``n_iters`` iterations; returns the number of failed chains."""
        if record is None:
            ra = rb = rt = None
            cap = 0
        else:
            ra, rb, rt = record['alpha'].data_ptr(), record['beta'].data_ptr(), record['tau'].data_ptr()
            cap = record['tau'].shape[1]
        with self.torch.cuda.device(self.device):
            return _lib.check(self.lib.occu_sweep(self.handle, int(n_iters), ra, rb, rt, int(rec_start), int(cap),
                                                  self.stream.cuda_stream), 'occu_sweep')

    def sweep_enqueue(self, n_iters, record=None, rec_start=0):
        """Enqueue ``n_iters`` iterations on this engine's stream without waiting (pair with ``finish``)."""
        if record is None:
            ra = rb = rt = None
            cap = 0
        else:
            ra, rb, rt = record['alpha'].data_ptr(), record['beta'].data_ptr(), record['tau'].data_ptr()
            cap = record['tau'].shape[1]
        with self.torch.cuda.device(self.device):
            _lib.check(self.lib.occu_sweep_enqueue(self.handle, int(n_iters), ra, rb, rt, int(rec_start), int(cap),
                                                   self.stream.cuda_stream), 'occu_sweep_enqueue')

    def finish(self):
        """Wait for everything enqueued on this engine; returns the number of failed chains."""
        with self.torch.cuda.device(self.device):
            return _lib.check(self.lib.occu_finish(self.handle, self.stream.cuda_stream), 'occu_finish')

    def chain_summaries(self, record, draws, max_lag):
        """Per-chain summary blocks of the recorded draws, computed where the draws are: {name: device tensor
        (chains, dim, 6 + max_lag + 1)} with {mean, M2, split-half means / M2s, autocovariances 0..max_lag}
        (occu_chain_summaries).  ``record`` must hold exactly ``draws`` draws per chain (capacity == draws)."""
        torch = self.torch
        out = {}
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            for name, rec in record.items():
                dim = rec.shape[2] if rec.dim() == 3 else 1
                assert rec.shape[1] == draws
                t = torch.empty((self.chains, dim, 6 + max_lag + 1), dtype=torch.float64, device=self.device)
                _lib.check(self.lib.occu_chain_summaries(rec.data_ptr(), self.chains, int(draws), int(dim), int(max_lag),
                                                         t.data_ptr(), self.stream.cuda_stream), 'occu_chain_summaries')
                out[name] = t
            self.stream.synchronize()
        return out

    def to_host(self, tensors):
        """Host (NumPy) copies of float64 device tensors: one packed transfer into the cached page-locked staging buffer on
        the engine's stream, one synchronisation, then a host copy of each part (the staging buffer is reused by the next
        call).  Pageable ``tensor.cpu()`` copies cost several times as much for the few MB a short run returns."""
        torch = self.torch
        counts = [int(t.numel()) for t in tensors]
        total = sum(-(-c // 8) * 8 for c in counts)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if total * 8 > _D2H_STAGING_MAX or any(t.dtype != torch.float64 for t in tensors):
                self.stream.synchronize()
                out = [t.cpu().numpy() for t in tensors]
            else:
                stage = _pinned_buffer(torch, self.device, self.slot, max(total, 1), kind='d2h')
                views, off = [], 0
                for t, c in zip(tensors, counts):
                    dst = stage[off:off + c].view(t.shape)
                    dst.copy_(t, non_blocking=True)
                    views.append(dst)
                    off += -(-c // 8) * 8
                self.stream.synchronize()
                out = [v.numpy().copy() for v in views]
        self.d2h_bytes += 8 * sum(counts)
        return out

    @property
    def launches(self):
        return int(self.lib.occu_launch_count(self.handle))

    KERNEL_CLASSES = ('psi_gemm', 'site_pass', 'gram_gemm', 'reduce', 'solve')

    def set_profiling(self, on):
        _lib.check(self.lib.occu_set_profiling(self.handle, int(bool(on))), 'occu_set_profiling')

    def kernel_ms(self):
        """{class: (accumulated ms, launches)} since profiling was switched on."""
        n = len(self.KERNEL_CLASSES)
        ms = (C.c_double * n)()
        calls = (C.c_longlong * n)()
        _lib.check(self.lib.occu_get_kernel_ms(self.handle, ms, calls, n), 'occu_get_kernel_ms')
        return {k: (ms[i], int(calls[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    # ---- deterministic-piece hooks (tests) -------------------------------------------------------
    def debug_gaussian_half(self, normals=None, tau=None, omega_a=None):
        torch = self.torch
        C_, q, p, r = self.chains, self.q, self.p, self.r
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            def dev(a, shape):
                if a is None:
                    return None
                a = np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), shape), order='C', copy=True)
                return torch.from_numpy(a).to(self.device)
            d = self.eta_dim
            tn, tt, to = dev(normals, (C_, d + p + r)), dev(tau, (C_,)), dev(omega_a, (C_, self.V))
            if q:
                eta_shapes = dict(eta_prec=(C_, q, q), eta_chol=(C_, q, q))
            else:  # ICAR: Lambda is banded; the factor comes back in column-band form, Lambda itself is not dumped
                eta_shapes = dict(eta_prec=(1,), eta_chol=(C_, d, self.icar_bw + 1))
            shapes = dict(eta_b=(C_, d), eta_mean=(C_, d), beta_prec=(C_, p, p), beta_b=(C_, p),
                          alpha_prec=(C_, r, r), alpha_b=(C_, r), tau_rate=(C_,), **eta_shapes)
            bufs = {k: torch.zeros(s, dtype=torch.float64, device=self.device) for k, s in shapes.items()}
            out = _lib.DebugOut(**{k: (t.data_ptr() if (q or k != 'eta_prec') else None) for k, t in bufs.items()})
            fails = _lib.check(self.lib.occu_debug_gaussian_half(
                self.handle, None if tn is None else tn.data_ptr(), None if tt is None else tt.data_ptr(),
                None if to is None else to.data_ptr(), C.byref(out), self.stream.cuda_stream), 'occu_debug_gaussian_half')
            self.stream.synchronize()
            res = {k: t.cpu().numpy() for k, t in bufs.items()}
        res['fails'] = fails
        return res

    def debug_site_half(self, keep_z=False, u_site=None, u_visit=None, nrm_eps=None):
        torch = self.torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            def dev(a, shape):
                if a is None:
                    return None
                a = np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), shape), order='C', copy=True)
                return torch.from_numpy(a).to(self.device)
            ts, tv, te = dev(u_site, (self.chains, self.n)), dev(u_visit, (self.chains, self.V)), dev(nrm_eps, (self.chains, self.n))
            ptr = lambda t: None if t is None else t.data_ptr()
            _lib.check(self.lib.occu_debug_site_half(self.handle, int(bool(keep_z)), ptr(ts), ptr(tv), ptr(te),
                                                     self.stream.cuda_stream), 'occu_debug_site_half')
            self.stream.synchronize()

    def close(self, keep_workspace=False):
        """Destroy the engine; ``keep_workspace=True`` returns the workspace tensor for reuse by a later engine."""
        ws = None
        if getattr(self, 'handle', None) is not None:
            self.lib.occu_destroy(self.handle)
            self.handle = None
            ws = self.workspace if keep_workspace else None
            self.workspace = None
        if getattr(self, 'slot', None) is not None:
            _release_slot(self.device, self.slot)
            self.slot = None
        return ws

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
