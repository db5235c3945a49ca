"""Deterministic synthetic light curves of the BASELINE.json shapes.

C3: Kepler long cadence, N = 65 536, 2 granulation + 1 oscillation bump + white noise + transit.
C4: 1024 stars x N = 4096.   C5: TESS 2-min cadence, N = 1 048 576.
The GP realisation is drawn with the celerite recursion itself (z_n = sqrt(D_n) eps_n,
y_n = z_n + U_n f_n), vectorised over stars; the transit is added by the caller from the engine's
own flux kernel (bench.py) so that no CPU transit code lives in the product package.
"""
import numpy as np

TRUTH_GP = np.array([300.0, 50.0, 150.0, 15.0, 250.0, 5.0, 150.0, 100.0])  # a1,b1,a2,b2,P_g,Q,nu_max,jitter
TRUTH_TR = np.array([5.9, 3.0, 0.03, 10.0, 88.5])                          # P,t0,Rrat,aR,inc[deg]
TR_FIXED = {"e": 0.0, "w": 90.0, "u1": 0.4996, "u2": 0.0847}
SIGMA = 50.0


def kepler_times(N, cadence_min=29.4244, gap_every=4400, gap_days=1.0):
    n = np.arange(N)
    return n * (cadence_min / 1440.0) + (n // gap_every) * gap_days


def tess_times(N):
    return kepler_times(N, cadence_min=2.0, gap_every=9864, gap_days=1.0)


def _u(a, b):
    return [False, None, "uniform", float(a), float(b)]


def c3_config(width=0.5):
    """Uniform priors bracketing the truth by +-50 %; inclination U(86, 90) degrees.  The fifth transit
    parameter reaches batman's `inc` verbatim (the reference's mapping, transit.py:103)."""
    g = TRUTH_GP
    lo, hi = 1 - width, 1 + width
    gp = [
        {"type": "granulation", "params": {"values": {"a_gran": _u(g[0] * lo, g[0] * hi), "b_gran": _u(g[1] * lo, g[1] * hi)}}},
        {"type": "granulation", "params": {"values": {"a_gran": _u(g[2] * lo, g[2] * hi), "b_gran": _u(g[3] * lo, g[3] * hi)}}},
        {"type": "oscillation_bump", "params": {"values": {"P_g": _u(g[4] * lo, g[4] * hi), "Q": _u(g[5] * lo, g[5] * hi),
                                                           "nu_max": _u(g[6] * lo, g[6] * hi)}}},
        {"type": "white_noise", "params": {"values": {"jitter": _u(g[7] * lo, g[7] * hi)}}},
    ]
    t = TRUTH_TR
    transit = [{"params": {"values": {
        "P": _u(t[0] * lo, t[0] * hi), "t0": _u(t[1] * lo, t[1] * hi), "Rrat": _u(t[2] * lo, t[2] * hi),
        "aR": _u(t[3] * lo, t[3] * hi), "cosi": _u(86.0, 90.0),
        "e": [True, TR_FIXED["e"], "uniform", 0, 1], "w": [True, TR_FIXED["w"], "uniform", 0, 90],
        "u1": [True, TR_FIXED["u1"], "uniform", 0, 1], "u2": [True, TR_FIXED["u2"], "uniform", 0, 1]}}}]
    return {"gp": gp, "transit": transit}


def sho_coefficients(theta_gp):
    """(a, b, c, d) of the three SHO terms and the jitter variance for TRUTH_GP-shaped vectors [..., 8]."""
    th = np.asarray(theta_gp, dtype=float)
    out = []
    for (amp, freq) in ((th[..., 0], th[..., 1]), (th[..., 2], th[..., 3])):
        S0 = (np.sqrt(2) / (2 * np.pi)) * amp ** 2 / freq
        w0 = 2 * np.pi * freq
        Q = 1 / np.sqrt(2)
        f = np.sqrt(4 * Q * Q - 1)
        out.append((S0 * w0 * Q, S0 * w0 * Q / f, 0.5 * w0 / Q, 0.5 * w0 / Q * f))
    Pg, Q, numax = th[..., 4], th[..., 5], th[..., 6]
    S0, w0 = Pg / (4 * Q * Q), 2 * np.pi * numax
    f = np.sqrt(4 * Q * Q - 1)
    out.append((S0 * w0 * Q, S0 * w0 * Q / f, 0.5 * w0 / Q, 0.5 * w0 / Q * f))
    a, b, c, d = (np.stack([o[i] for o in out], axis=-1) for i in range(4))
    return a, b, c, d, th[..., 7] ** 2


def draw_gp(t_days, theta_gp, sigma, rng):
    """Realisations of the GP + white noise for a batch of stars: theta_gp [B, 8] -> y [B, N] (ppm)."""
    th = np.atleast_2d(theta_gp)
    B, N = th.shape[0], t_days.size
    x = t_days * 0.0864
    a, b, c, d, jit2 = sho_coefficients(th)          # [B, 3]
    asum = jit2 + a.sum(axis=1)
    J = 6
    P = np.zeros((B, J, J))
    g = np.zeros((B, J))
    y = np.empty((B, N))
    eps = rng.standard_normal((B, N))
    U = np.empty((B, J))
    V = np.empty((B, J))
    phi = np.ones((B, J))
    for n in range(N):
        if n:
            e = np.exp(-c * (x[n] - x[n - 1]))
            phi[:, 0::2] = e
            phi[:, 1::2] = e
        cd, sd = np.cos(d * x[n]), np.sin(d * x[n])
        U[:, 0::2] = a * cd + b * sd
        U[:, 1::2] = a * sd - b * cd
        V[:, 0::2] = cd
        V[:, 1::2] = sd
        S = P * phi[:, :, None] * phi[:, None, :]
        f = phi * g
        tv = np.einsum("bij,bj->bi", S, U)
        D = sigma ** 2 + asum - np.einsum("bi,bi->b", U, tv)
        w = (V - tv) / D[:, None]
        z = np.sqrt(D) * eps[:, n]
        y[:, n] = z + np.einsum("bi,bi->b", U, f)
        g = f + w * z[:, None]
        P = S + D[:, None, None] * w[:, :, None] * w[:, None, :]
    return y


def make_star(N=65536, seed=2452144, cadence="kepler", theta_gp=TRUTH_GP):
    """-> t_days, gp_ppm (zero-mean GP + noise realisation), sigma array."""
    rng = np.random.default_rng(seed)
    t = kepler_times(N) if cadence == "kepler" else tess_times(N)
    y = draw_gp(t, np.asarray(theta_gp)[None, :], SIGMA, rng)[0]
    return t, y, np.full(N, SIGMA)


def make_stars(nstars, N=4096, seed0=1000):
    """C4: per-star truths drawn from the C3 priors; one shared time grid with one gap."""
    t = kepler_times(N, gap_every=N // 2)
    rng = np.random.default_rng(seed0)
    th = TRUTH_GP * rng.uniform(0.5, 1.5, size=(nstars, 8))
    y = draw_gp(t, th, SIGMA, rng)
    return t, y, np.full(N, SIGMA), th
