"""Engine: one GPU, one model structure, any number of stars (light curves).

Thin object wrapper over the C ABI; numpy arrays in, numpy arrays out.  Mirrors
the three reference seams: log-posterior of a batch of
walkers, batman-style transit flux, and the emcee stretch-move run.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GptError, ModelDesc, Prior

_dp = C.POINTER(C.c_double)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class Engine:
    def __init__(self, device=0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        rc = self._lib.gpt_engine_create(int(device), C.byref(h))
        if rc != 0:
            raise GptError(f"gpt_engine_create(device={device}) failed with {rc}: no usable CUDA device "
                           "(this package has no CPU fallback)")
        self._h = h
        self.device = int(device)
        self.ndim = None
        self._nstars = 0
        self._N = {}
        self._comm = None

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpt_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GptError(f"[{rc}] {self._lib.gpt_last_error(self._h).decode()}")

    def set_option(self, key, value):
        self._ck(self._lib.gpt_set_option(self._h, key.encode(), float(value)))

    def stat(self, key):
        v = C.c_double()
        self._ck(self._lib.gpt_get_stat(self._h, key.encode(), C.byref(v)))
        return v.value

    # -- model + data
    def set_model(self, spec):
        """spec: dict(comp_type, prior_kind, prior_a, prior_b, has_transit, tr_fixed, tr_free,
        supersample, exp_time, ellip_mode) -- see gptransits_b200.model.build_spec."""
        ct = np.ascontiguousarray(spec["comp_type"], dtype=np.int32)
        kind = np.asarray(spec["prior_kind"], dtype=np.int32)
        pa, pb = _d(spec["prior_a"]), _d(spec["prior_b"])
        pri = (Prior * max(1, kind.size))()
        for i in range(kind.size):
            pri[i].kind, pri[i].a, pri[i].b = int(kind[i]), float(pa[i]), float(pb[i])
        d = ModelDesc()
        d.ncomp = ct.size
        d.comp_type = ct.ctypes.data_as(C.POINTER(C.c_int32))
        d.has_transit = int(spec.get("has_transit", 0))
        fixed = _d(spec.get("tr_fixed", np.zeros(9)))
        free = np.asarray(spec.get("tr_free", np.zeros(9)), dtype=np.uint8)
        for i in range(9):
            d.tr_fixed[i] = float(fixed[i])
            d.tr_free[i] = int(free[i])
        d.supersample = int(spec.get("supersample", 1))
        d.exp_time_days = float(spec.get("exp_time", 0.0))
        d.ellip_mode = int(spec.get("ellip_mode", 0))
        d.inc_mode = int(spec.get("inc_mode", 0))
        d.ndim = kind.size
        d.priors = pri
        self._ck(self._lib.gpt_model_set(self._h, C.byref(d)))
        self.ndim = int(kind.size)

    def upload_star(self, star, t_days, y, yerr=None):
        t, y = _d(t_days), _d(y)
        if t.shape != y.shape or t.ndim != 1:
            raise ValueError("t_days and y must be 1-D arrays of the same length")
        e = None if yerr is None else _d(np.broadcast_to(yerr, t.shape))
        self._ck(self._lib.gpt_star_upload(self._h, int(star), _p(t), _p(y), _p(e) if e is not None else None,
                                           t.size))
        self._N[int(star)] = t.size
        self._nstars = max(self._nstars, int(star) + 1)

    def upload_stars(self, star0, t_days, ys, yerr=None):
        """Stars star0 .. star0+len(ys)-1 observed on one time grid: ys [nstars, N], one yerr for all."""
        t, ys = _d(t_days), _d(ys)
        if ys.ndim != 2 or t.ndim != 1 or ys.shape[1] != t.size:
            raise ValueError("ys must be [nstars, N] on the time grid t_days [N]")
        e = None if yerr is None else _d(np.broadcast_to(yerr, t.shape))
        self._ck(self._lib.gpt_stars_upload(self._h, int(star0), ys.shape[0], _p(t), _p(ys),
                                            _p(e) if e is not None else None, t.size))
        for s in range(int(star0), int(star0) + ys.shape[0]):
            self._N[s] = t.size
        self._nstars = max(self._nstars, int(star0) + ys.shape[0])

    def clear_stars(self):
        self._ck(self._lib.gpt_star_clear(self._h))
        self._N = {}
        self._nstars = 0

    # -- posterior seam
    def log_prob(self, theta, star=0, nstars=1, return_parts=False):
        """theta[(nstars,) W, ndim] -> lp[(nstars,) W] (and parts[..., 4] = lnprior, quad, logdet, lnL)."""
        theta = _d(theta)
        shape = theta.shape[:-1]
        if theta.shape[-1] != self.ndim:
            raise ValueError(f"theta last axis must be ndim={self.ndim}")
        flat = theta.reshape(-1, self.ndim)
        if flat.shape[0] % nstars:
            raise ValueError("number of walkers must be a multiple of nstars")
        W = flat.shape[0] // nstars
        lp = np.empty(flat.shape[0])
        parts = np.empty((flat.shape[0], 4)) if return_parts else None
        self._ck(self._lib.gpt_log_prob_multi(self._h, int(star), int(nstars), _p(flat), W, _p(lp),
                                              _p(parts) if return_parts else None))
        lp = lp.reshape(shape)
        if return_parts:
            return lp, parts.reshape(shape + (4,))
        return lp

    def transit_flux(self, pars9, star=0):
        """pars9[W, 9] = per,t0,rp,a,inc[deg],ecc,w[deg],u1,u2 (batman semantics) -> flux[W, N]."""
        p = _d(pars9).reshape(-1, 9)
        out = np.empty((p.shape[0], self._N[int(star)]))
        self._ck(self._lib.gpt_transit_flux(self._h, int(star), _p(p), p.shape[0], _p(out)))
        return out

    # -- sampler seam
    def sample(self, init, nsteps, seed=42, step0=0, a=2.0, star=0, nstars=1):
        """emcee-style run.  init[(nstars,) W, ndim] -> chain[(nstars,) W, nsteps, ndim],
        logp[(nstars,) nsteps, W], naccept[(nstars,) W]."""
        init = _d(init)
        flat = init.reshape(nstars, -1, self.ndim)
        W = flat.shape[1]
        chain = np.empty((nstars, W, nsteps, self.ndim))
        logp = np.empty((nstars, nsteps, W))
        nacc = np.zeros((nstars, W), dtype=np.int64)
        self._ck(self._lib.gpt_sample(self._h, int(star), int(nstars), _p(flat), W, int(nsteps), int(seed),
                                      int(step0), float(a), _p(chain), _p(logp),
                                      nacc.ctypes.data_as(C.POINTER(C.c_int64))))
        if init.ndim == 2:
            return chain[0], logp[0], nacc[0]
        return chain, logp, nacc

    def sampler_begin(self, init, seed=42, step0=0, a=2.0, star=0, nstars=1, star_ids=None):
        flat = _d(init).reshape(nstars, -1, self.ndim)
        ids = None
        if star_ids is not None:
            ids = np.ascontiguousarray(star_ids, dtype=np.int32)
        self._ck(self._lib.gpt_sampler_begin(self._h, int(star), int(nstars),
                                             ids.ctypes.data_as(C.POINTER(C.c_int32)) if ids is not None else None,
                                             _p(flat), flat.shape[1], int(seed), int(step0), float(a)))
        self._sW, self._sns = flat.shape[1], nstars

    def sampler_run(self, nsteps, keep=True):
        self._ck(self._lib.gpt_sampler_run(self._h, int(nsteps), 1 if keep else 0))

    def sampler_read(self, first, count):
        chain = np.empty((self._sns, self._sW, count, self.ndim))
        logp = np.empty((self._sns, count, self._sW))
        self._ck(self._lib.gpt_sampler_read(self._h, int(first), int(count), _p(chain), _p(logp)))
        return chain, logp

    def sampler_state(self):
        x = np.empty((self._sns, self._sW, self.ndim))
        lp = np.empty((self._sns, self._sW))
        na = np.zeros((self._sns, self._sW), dtype=np.int64)
        self._ck(self._lib.gpt_sampler_state(self._h, _p(x), _p(lp), na.ctypes.data_as(C.POINTER(C.c_int64))))
        return x, lp, na

    def sampler_rhat(self, first, count):
        """Gelman-Rubin V/W per (star, parameter) over kept steps [first, first+count), on the device."""
        out = np.empty((self._sns, self.ndim))
        self._ck(self._lib.gpt_sampler_rhat(self._h, int(first), int(count), _p(out)))
        return out

    def sampler_geweke(self, first, count, frac1=0.1, frac2=0.5, bins=10):
        """Geweke z-scores (convergence.geweke's conventions) of kept steps [first, first+count) on the device:
        -> starts[bins], z[nstars, bins, W, ndim]."""
        starts = np.ascontiguousarray(np.linspace(0, int(count) / 2, num=bins).astype(np.int64))
        z = np.empty((self._sns, bins, self._sW, self.ndim))
        self._ck(self._lib.gpt_sampler_geweke(self._h, int(first), int(count), int(bins),
                                              starts.ctypes.data_as(C.POINTER(C.c_int64)), float(frac1), float(frac2), _p(z)))
        return starts, z

    def sampler_end(self):
        self._ck(self._lib.gpt_sampler_end(self._h))

    # -- one star sharded by walker across ranks (parallel.run_walker_sharded)
    def sampler_reserve(self, nsteps):
        self._ck(self._lib.gpt_sampler_reserve(self._h, int(nsteps)))

    def sampler_lq_device(self):
        """(device pointer, count) of the proposal log-probabilities lq[W/2] of the current half-step."""
        ptr, n = C.c_void_p(), C.c_int64()
        self._ck(self._lib.gpt_sampler_lq_device(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def device_copy(self, host, device_ptr, to_device):
        """Copy between a contiguous float64 host array and device memory of this engine (engine stream, synchronous)."""
        host = np.ascontiguousarray(host, dtype=np.float64) if to_device else host
        if host.nbytes:
            if to_device:
                self._ck(self._lib.gpt_device_copy(self._h, C.c_void_p(int(device_ptr)), host.ctypes.data_as(C.c_void_p), host.nbytes, 1))
            else:
                self._ck(self._lib.gpt_device_copy(self._h, host.ctypes.data_as(C.c_void_p), C.c_void_p(int(device_ptr)), host.nbytes, 0))
        return host

    def sampler_half_eval(self, half, lo, hi):
        self._ck(self._lib.gpt_sampler_half_eval(self._h, int(half), int(lo), int(hi)))

    def sampler_half_accept(self, half, keep=True):
        self._ck(self._lib.gpt_sampler_half_accept(self._h, int(half), 1 if keep else 0))

    def comm_init(self, dist):
        """Join the NCCL communicator of the walker-sharded sampler: rank 0 draws the id, torch.distributed (any
        backend) carries its 128 bytes to the other ranks, every rank joins.  Returns (rank, world)."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            self._ck(self._lib.gpt_comm_id(C.cast(buf, C.c_void_p)))
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        tid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        dist.broadcast(tid, src=0)
        raw = bytes(tid.cpu().tolist())
        buf = (C.c_uint8 * 128).from_buffer_copy(raw)
        self._ck(self._lib.gpt_comm_init(self._h, C.cast(buf, C.c_void_p), rank, world))
        self._comm = (rank, world)
        return rank, world

    def comm_destroy(self):
        self._ck(self._lib.gpt_comm_destroy(self._h))
        self._comm = None

    def sampler_run_sharded(self, nsteps, keep=True, timed=False):
        """nsteps with the active half sharded by walker over the communicator of comm_init; everything,
        the all-gather included, runs on the engine's stream.  timed: -> per-step device milliseconds."""
        ms = (C.c_float * int(nsteps))() if timed else None
        self._ck(self._lib.gpt_sampler_run_sharded(self._h, int(nsteps), 1 if keep else 0, ms))
        return np.array(ms[:], dtype=np.float64) if timed else None

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._ck(self._lib.gpt_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    # -- measurement helpers (bench.py)
    def sampler_run_timed(self, nsteps, keep=False, flush_l2=True):
        """-> per-step milliseconds (CUDA events on the engine stream)."""
        ms = (C.c_float * int(nsteps))()
        self._ck(self._lib.gpt_sampler_run_timed(self._h, int(nsteps), 1 if keep else 0, 1 if flush_l2 else 0, ms))
        return np.array(ms[:], dtype=np.float64)

    def log_prob_resident_timed(self, theta, reps, star=0, nstars=1, flush_l2=True):
        """Upload theta once, evaluate it `reps` times from HBM.  -> (lp, ms per rep)."""
        flat = _d(theta).reshape(-1, self.ndim)
        W = flat.shape[0] // nstars
        d_theta, d_lp = C.c_void_p(), C.c_void_p()
        self._ck(self._lib.gpt_device_alloc(self._h, flat.nbytes, C.byref(d_theta)))
        self._ck(self._lib.gpt_device_alloc(self._h, flat.shape[0] * 8, C.byref(d_lp)))
        try:
            self._ck(self._lib.gpt_device_copy(self._h, d_theta, flat.ctypes.data_as(C.c_void_p), flat.nbytes, 1))
            ms = (C.c_float * int(reps))()
            self._ck(self._lib.gpt_log_prob_device_timed(self._h, int(star), int(nstars), d_theta, W, d_lp,
                                                         int(reps), 1 if flush_l2 else 0, ms))
            lp = np.empty(flat.shape[0])
            self._ck(self._lib.gpt_device_copy(self._h, lp.ctypes.data_as(C.c_void_p), d_lp, lp.nbytes, 0))
        finally:
            self._lib.gpt_device_free(self._h, d_theta)
            self._lib.gpt_device_free(self._h, d_lp)
        return lp, np.array(ms[:], dtype=np.float64)

    # -- GP conditional prediction (celerite GP.predict)
    def gp_predict(self, theta, x_days, star=0, return_var=True):
        """mu*, var* of the GP at times x_days for ONE walker vector theta[ndim]; the transit model of
        theta (if any) is subtracted from the star's data first."""
        th = _d(theta).reshape(-1)
        if th.size != self.ndim:
            raise ValueError(f"theta must have ndim={self.ndim} entries")
        x = _d(x_days).reshape(-1)
        mu = np.empty(x.size)
        var = np.empty(x.size) if return_var else None
        self._ck(self._lib.gpt_gp_predict(self._h, int(star), _p(th), _p(x), x.size, _p(mu),
                                          _p(var) if return_var else None))
        return (mu, var) if return_var else mu
