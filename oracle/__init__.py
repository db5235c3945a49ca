"""CPU oracle for the gptransits posterior hot path -- TEST INFRASTRUCTURE ONLY.

Thin ctypes wrapper over ``gpt_oracle.c`` (see that file's header for what it
restates and its parity status: *parity unpinned* -- the reference has no tests
and its numerical dependencies are absent from this image).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  The product package
``gptransits_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgpt_oracle.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)


def build(force=False):
    """Compile gpt_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "gpt_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libgpt_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(_LIB_PATH):
                raise
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_logpdf.restype = C.c_double
        _lib.orc_logpdf.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        _lib.orc_lnprior.restype = C.c_double
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def logpdf(kind, a, b, x):
    return lib().orc_logpdf(int(kind), float(a), float(b), float(x))


def lnprior(kind, pa, pb, x):
    kind = np.ascontiguousarray(kind, dtype=np.int32)
    pa, pb, x = _d(pa), _d(pb), _d(x)
    f = lib().orc_lnprior
    f.argtypes = [c_int_p, c_double_p, c_double_p, c_double_p, C.c_int]
    return f(_ip(kind), _dp(pa), _dp(pb), _dp(x), x.size)


def gp_celerite_params(comp_type, theta):
    ct = np.ascontiguousarray(comp_type, dtype=np.int32)
    theta = _d(theta)
    out = np.zeros(4 * max(1, ct.size))
    n = lib().orc_gp_celerite_params(_ip(ct), C.c_int(ct.size), _dp(theta), _dp(out))
    return out[:n].copy()


def gp_coeffs(comp_type, theta):
    """-> dict(jitter, a_real, c_real, a_comp, b_comp, c_comp, d_comp)."""
    ct = np.ascontiguousarray(comp_type, dtype=np.int32)
    theta = _d(theta)
    out = np.zeros(8 * max(1, ct.size) + 3)
    lib().orc_gp_coeffs(_ip(ct), C.c_int(ct.size), _dp(theta), _dp(out))
    jit, nr, nc = out[0], int(out[1]), int(out[2])
    o = 3
    r = {"jitter": jit}
    for name, n in (("a_real", nr), ("c_real", nr), ("a_comp", nc), ("b_comp", nc),
                    ("c_comp", nc), ("d_comp", nc)):
        r[name] = out[o:o + n].copy()
        o += n
    return r


def celerite_loglike(comp_type, theta_gp, x, diag, resid):
    """-> (status, quad, logdet); x in the GP's time unit, diag = yerr**2."""
    ct = np.ascontiguousarray(comp_type, dtype=np.int32)
    theta_gp, x, diag, resid = _d(theta_gp), _d(x), _d(diag), _d(resid)
    quad, logdet = C.c_double(), C.c_double()
    st = lib().orc_celerite_loglike(_ip(ct), C.c_int(ct.size), _dp(theta_gp), C.c_int(x.size),
                                    _dp(x), _dp(diag), _dp(resid), C.byref(quad), C.byref(logdet))
    return st, quad.value, logdet.value


def celerite_factor_D(comp_type, theta_gp, x, diag):
    ct = np.ascontiguousarray(comp_type, dtype=np.int32)
    theta_gp, x, diag = _d(theta_gp), _d(x), _d(diag)
    D = np.zeros(x.size)
    st = lib().orc_celerite_factor_D(_ip(ct), C.c_int(ct.size), _dp(theta_gp), C.c_int(x.size),
                                     _dp(x), _dp(diag), _dp(D))
    return st, D


def rsky(t, tc, per, a, inc_rad, ecc, omega_rad):
    t = _d(t)
    out = np.zeros(t.size)
    f = lib().orc_rsky
    f.argtypes = [c_double_p, C.c_int] + [C.c_double] * 6 + [c_double_p]
    f(_dp(t), t.size, tc, per, a, inc_rad, ecc, omega_rad, _dp(out))
    return out


def quadratic_ld(ds, p, u1, u2, mode=0):
    ds = _d(ds)
    out = np.zeros(ds.size)
    f = lib().orc_quadratic_ld
    f.argtypes = [c_double_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, c_double_p]
    f(_dp(ds), ds.size, p, u1, u2, mode, _dp(out))
    return out


def light_curve(t_days, pars9, supersample=1, exp_time=0.0, mode=0):
    """batman.TransitModel(...).light_curve with pars9 = per,t0,rp,a,inc[deg],ecc,w[deg],u1,u2."""
    t, pars9 = _d(t_days), _d(pars9)
    out = np.zeros(t.size)
    f = lib().orc_light_curve
    f.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, C.c_double, C.c_int, c_double_p]
    f(_dp(t), t.size, _dp(pars9), int(supersample), float(exp_time), int(mode), _dp(out))
    return out


def _spec_args(spec):
    ct = np.ascontiguousarray(spec["comp_type"], dtype=np.int32)
    fixed = _d(spec.get("tr_fixed", np.zeros(9)))
    free = np.ascontiguousarray(spec.get("tr_free", np.zeros(9)), dtype=np.uint8)
    kind = np.ascontiguousarray(spec["prior_kind"], dtype=np.int32)
    pa, pb = _d(spec["prior_a"]), _d(spec["prior_b"])
    keep = (ct, fixed, free, kind, pa, pb)
    args = [C.c_int(ct.size), _ip(ct), C.c_int(int(spec.get("has_transit", 0))), _dp(fixed),
            free.ctypes.data_as(c_ubyte_p), C.c_int(kind.size), _ip(kind), _dp(pa), _dp(pb),
            C.c_int(int(spec.get("supersample", 1))), C.c_double(float(spec.get("exp_time", 0.0))),
            C.c_int(int(spec.get("ellip_mode", 0))), C.c_int(int(spec.get("inc_mode", 0)))]
    return args, keep, kind.size


def _star_args(t_days, y, yerr):
    t, y = _d(t_days), _d(y)
    e = None if yerr is None else _d(yerr)
    keep = (t, y, e)
    return [C.c_int(t.size), _dp(t), _dp(y), _dp(e) if e is not None else None], keep


def log_prob(spec, t_days, y, yerr, theta, nthreads=1, return_parts=False):
    """Posterior of model.py:299-325 for theta[W, ndim].  yerr=None -> celerite default 1.123e-12."""
    margs, keep1, ndim = _spec_args(spec)
    sargs, keep2 = _star_args(t_days, y, yerr)
    theta = _d(theta).reshape(-1, ndim)
    W = theta.shape[0]
    lp = np.zeros(W)
    parts = np.zeros((W, 4))
    lib().orc_log_prob(*margs, *sargs, _dp(theta), C.c_int(W), _dp(lp), _dp(parts), C.c_int(nthreads))
    del keep1, keep2
    return (lp, parts) if return_parts else lp


def sample(spec, t_days, y, yerr, init, nsteps, seed=42, star=0, step0=0, a=2.0, nthreads=1):
    """emcee-style stretch-move run with the Philox stream of DESIGN.md.
    -> chain[W, nsteps, ndim], logp[nsteps, W], naccept[W]"""
    margs, keep1, ndim = _spec_args(spec)
    sargs, keep2 = _star_args(t_days, y, yerr)
    init = _d(init).reshape(-1, ndim)
    W = init.shape[0]
    chain = np.zeros((W, nsteps, ndim))
    logp = np.zeros((nsteps, W))
    nacc = np.zeros(W, dtype=np.int64)
    st = lib().orc_sample(*margs, *sargs, C.c_int(W), _dp(init), C.c_long(nsteps), C.c_uint64(seed),
                          C.c_uint32(star), C.c_uint64(step0), C.c_double(a), _dp(chain), _dp(logp),
                          nacc.ctypes.data_as(C.POINTER(C.c_long)), C.c_int(nthreads))
    if st != 0:
        raise ValueError("orc_sample: W must be even and >= 2")
    del keep1, keep2
    return chain, logp, nacc


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().orc_philox4x32_10(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p), out.ctypes.data_as(u32p))
    return out
