"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Tolerance: 1e-9 relative on quad, logdet, lnL and on the transit flux (BASELINE.json north_star),
checked SEPARATELY for quad and logdet (with the 1e6 ppm baseline the total
lnL would hide a log-det error).
"""
import numpy as np
import pytest

import oracle
from gptransits_b200 import BatmanModel, GPModel, build_spec

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def assert_parts_close(got, want, rtol=RTOL):
    # columns: lnprior, quad, logdet, lnL
    for col, name in enumerate(("lnprior", "quad", "logdet", "lnL")):
        g, w = got[..., col], want[..., col]
        fin = np.isfinite(w)
        assert np.array_equal(np.isfinite(g), fin), name
        if fin.any():
            assert rel(g[fin], w[fin]) < rtol, (name, rel(g[fin], w[fin]))


def kic_setup(kic_lc, kic_config, baseline_removed=False):
    t, f, e = kic_lc
    y = (f - 1) * 1e6 if baseline_removed else f * 1e6
    gp = GPModel(kic_config["gp"])
    return t, y, e * 1e6, build_spec(gp, None), gp


@pytest.mark.parametrize("baseline_removed", [False, True])
def test_kic_gp_only(engine, kic_lc, kic_config, golden, baseline_removed):
    t, y, yerr, spec, gp = kic_setup(kic_lc, kic_config, baseline_removed)
    theta = np.array(golden["kic-012008916"]["theta_gp"])
    theta = np.vstack([theta, golden["kic-012008916"]["theta_gp_outside"]])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want_parts = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True)
    for chunks in (1, 2, 5, 0):
        engine.set_option("chunks", chunks)
        lp, parts = engine.log_prob(theta, return_parts=True)
        assert_parts_close(parts, want_parts)
        assert lp[-1] == -np.inf and want_lp[-1] == -np.inf
        assert rel(lp[:-1], want_lp[:-1]) < RTOL
    engine.set_option("chunks", 0)


def test_kic_survey_golden_theta(engine, kic_lc, kic_config):
    """theta_gp row, logdet = 14345.578386664782, quad = 981.5028166301665."""
    t, y, yerr, spec, gp = kic_setup(kic_lc, kic_config, baseline_removed=True)
    theta = np.array([[249.74246655, 5.19060939, 105.80836122, 368.63266001, 68.27098486, 88.68944315]])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    _, parts = engine.log_prob(theta, return_parts=True)
    assert abs(parts[0, 2] / 14345.578386664782 - 1) < RTOL
    assert abs(parts[0, 1] / 981.5028166301665 - 1) < RTOL
    assert abs(parts[0, 3] / -8850.809186547911 - 1) < RTOL


def sim_setup(sim_lc, sim_config, inc_mapping="reference"):
    t, f, e = sim_lc
    np.random.seed(42)
    gp = GPModel(sim_config["gp"])
    tr = BatmanModel(sim_config["transit"][0])
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr, {"inc_mapping": inc_mapping})
    return t, f * 1e6, None, spec, gp, tr


@pytest.mark.parametrize("inc_mapping", ["reference", "cosi"])
def test_sim_gp_plus_transit(engine, sim_lc, sim_config, golden, inc_mapping):
    t, y, yerr, spec, gp, tr = sim_setup(sim_lc, sim_config, inc_mapping)
    g = golden["sim-2452144"]
    theta = np.hstack([np.array(g["theta_gp"]), np.array(g["theta_tr"])])
    # realistic transiting geometry for the cosi mapping: b = aR*cosi < 1
    extra = theta[:4].copy()
    extra[:, 6:] = [[5.9, 3.0, 0.03, 10.0, 0.02], [5.8, 3.1, 0.08, 8.0, 0.11], [6.0, 2.9, 0.05, 12.0, 0.08],
                    [5.9, 3.0, 0.099, 3.0, 0.3]]
    theta = np.vstack([theta, extra])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want_parts = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True)
    for chunks in (1, 3, 0):
        engine.set_option("chunks", chunks)
        lp, parts = engine.log_prob(theta, return_parts=True)
        assert_parts_close(parts, want_parts)
    engine.set_option("chunks", 0)


def test_transit_flux_matches_batman_restatement(engine, sim_lc, sim_config):
    t, y, yerr, spec, gp, tr = sim_setup(sim_lc, sim_config)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    rng = np.random.default_rng(7)
    pars = []
    for _ in range(24):
        pars.append([rng.uniform(2, 8), rng.uniform(2.8, 3.3), 10 ** rng.uniform(-2.5, -0.5), rng.uniform(3, 15),
                     rng.uniform(84, 90), 0.0, 90.0, 0.4996, 0.0847])
    # eccentric orbits, grazing and large planets, inclination past 90, negative rp, face-on
    pars += [[5.9, 3.0, 0.1, 9.0, 87.0, 0.3, 40.0, 0.3, 0.2], [3.3, 3.1, 0.05, 6.0, 89.0, 0.6, 200.0, 0.5, 0.1],
             [5.9, 3.0, 0.3, 4.0, 78.0, 0.0, 90.0, 0.4, 0.1], [5.9, 3.0, 1.2, 3.0, 85.0, 0.0, 90.0, 0.4, 0.1],
             [5.9, 3.0, 0.05, 10.0, 93.0, 0.0, 90.0, 0.4, 0.1], [5.9, 3.0, -0.05, 10.0, 88.0, 0.0, 90.0, 0.4, 0.1],
             [5.9, 3.0, 0.05, 1.02, 0.5, 0.0, 90.0, 0.4996, 0.0847], [5.9, 3.0, 0.05, 10.0, 88.0, 0.0, 90.0, 0.0, 0.0],
             [0.7, 3.0, 0.1, 2.0, 80.0, 0.1, 10.0, 0.6, 0.0]]
    pars = np.array(pars)
    for mode in (0, 1):
        spec_m = dict(spec, ellip_mode=mode)
        engine.set_model(spec_m)
        flux = {}
        # strict = batman's own operation order; 0 = the throughput arrangement of the same formulas
        for strict in (0, 1):
            engine.set_option("transit_strict", strict)
            got = flux[strict] = engine.transit_flux(pars)
            for k, p in enumerate(pars):
                want = oracle.light_curve(t, p, supersample=6, exp_time=spec["exp_time"], mode=mode)
                assert np.array_equal(np.isnan(got[k]), np.isnan(want)), k
                ok = ~np.isnan(want)
                assert np.max(np.abs(got[k][ok] - want[ok])) < 1e-9, (mode, strict, k, np.max(np.abs(got[k][ok] - want[ok])))
            assert np.any(got < 1.0)
        ok = ~np.isnan(flux[0])
        assert np.max(np.abs(flux[0][ok] - flux[1][ok])) < 1e-12
    engine.set_option("transit_strict", 0)
    engine.set_model(spec)


def synthetic_star(N, seed, cadence_min=29.4244, gap_every=4400):
    rng = np.random.default_rng(seed)
    t = np.arange(N) * (cadence_min / 1440.0) + (np.arange(N) // gap_every) * 1.0
    y = 1e6 + 300 * np.sin(t * 3.1) + rng.normal(0, 120, N)
    yerr = np.full(N, 50.0) * rng.uniform(0.9, 1.1, N)
    return t, y, yerr


C3_CONFIG_GP = [
    {"type": "granulation", "params": {"values": {"a_gran": [False, None, "uniform", 150, 450],
                                                  "b_gran": [False, None, "uniform", 25, 75]}}},
    {"type": "granulation", "params": {"values": {"a_gran": [False, None, "uniform", 75, 225],
                                                  "b_gran": [False, None, "uniform", 7.5, 22.5]}}},
    {"type": "oscillation_bump", "params": {"values": {"P_g": [False, None, "uniform", 125, 375],
                                                       "Q": [False, None, "uniform", 2.5, 7.5],
                                                       "nu_max": [False, None, "uniform", 75, 225]}}},
    {"type": "white_noise", "params": {"values": {"jitter": [False, None, "uniform", 50, 150]}}},
]
C3_CONFIG_TR = {"params": {"values": {
    "P": [False, None, "uniform", 2.95, 8.85], "t0": [False, None, "uniform", 1.5, 4.5],
    "Rrat": [False, None, "uniform", 0.015, 0.045], "aR": [False, None, "uniform", 5.0, 15.0],
    "cosi": [False, None, "uniform", 86.0, 90.0],   # reference mapping: the value IS batman's inc in degrees
    "e": [True, 0.0, "uniform", 0, 1], "w": [True, 90.0, "uniform", 0, 90],
    "u1": [True, 0.4996, "uniform", 0, 1], "u2": [True, 0.0847, "uniform", 0, 1]}}}


@pytest.mark.parametrize("N,irregular", [(8192, False), (8192, True), (20000, False)])
def test_synthetic_three_terms_chunked(engine, N, irregular):
    t, y, yerr = synthetic_star(N, 11)
    if irregular:
        rng = np.random.default_rng(5)
        t = t + rng.uniform(-1e-7, 1e-7, N)         # barycentric-style timing jitter
        t[N // 3:] += 0.0123                          # odd gap
    np.random.seed(3)
    gp = GPModel(C3_CONFIG_GP)
    tr = BatmanModel(C3_CONFIG_TR)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    theta = np.hstack([gp.sample_prior(16), tr.sample_prior(16)])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want_parts = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    for chunks, halo in ((1, 0), (8, 0), (16, 0), (0, 0), (8, 64)):
        engine.set_option("chunks", chunks)
        engine.set_option("halo", halo)
        lp, parts = engine.log_prob(theta, return_parts=True)
        assert_parts_close(parts, want_parts)
    engine.set_option("chunks", 0)
    engine.set_option("halo", 0)
    assert np.all(np.isfinite(want_lp))


def test_short_halo_is_caught_by_verification(engine):
    """A deliberately useless halo must be detected and repaired by the exact second pass."""
    t, y, yerr = synthetic_star(8192, 2)
    np.random.seed(3)
    gp = GPModel(C3_CONFIG_GP)
    spec = build_spec(gp, None)
    theta = gp.sample_prior(8)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want_parts = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True)
    engine.set_option("chunks", 16)
    engine.set_option("halo", 2)
    before = engine.stat("pass2_chunks")
    lp, parts = engine.log_prob(theta, return_parts=True)
    assert engine.stat("pass2_chunks") > before
    assert_parts_close(parts, want_parts)
    # with verification off the same setting is visibly wrong: the check is what protects parity
    engine.set_option("verify", 0)
    lp2, parts2 = engine.log_prob(theta, return_parts=True)
    assert rel(parts2[:, 1], want_parts[:, 1]) > 1e-9
    engine.set_option("verify", 1)
    engine.set_option("chunks", 0)
    engine.set_option("halo", 0)


def test_edge_cases(engine):
    # tiny N, single walker, white-noise-only model, transit-only model, Q < 1/2, non-finite input
    wn = [{"type": "white_noise", "params": {"values": {"jitter": [False, None, "uniform", 1, 300]}}}]
    for N in (1, 2, 3, 129, 257):
        t, y, yerr = synthetic_star(N, N)
        for cfg in (C3_CONFIG_GP[:1] + wn, wn):
            gp = GPModel(cfg)
            spec = build_spec(gp, None)
            np.random.seed(N)
            theta = gp.sample_prior(3)
            engine.clear_stars()
            engine.set_model(spec)
            engine.upload_star(0, t, y, yerr)
            _, parts = engine.log_prob(theta, return_parts=True)
            _, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True)
            assert_parts_close(parts, want)
            # known-answer identity: white noise only => independent Gaussians
            if cfg is wn:
                var = yerr ** 2 + theta[:, :1] ** 2
                lnl = -0.5 * np.sum(y ** 2 / var + np.log(var) + np.log(2 * np.pi), axis=1)
                assert rel(parts[:, 3], lnl) < RTOL
    # transit only (chi^2, model.py:328-343 with the prior sign fixed)
    t, y, yerr = synthetic_star(600, 9)
    tr = BatmanModel(C3_CONFIG_TR)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(None, tr)
    np.random.seed(1)
    theta = tr.sample_prior(5)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y - 1e6, yerr)
    lp, parts = engine.log_prob(theta, return_parts=True)
    want_lp, want = oracle.log_prob(spec, t, y - 1e6, yerr, theta, return_parts=True)
    assert rel(lp, want_lp) < RTOL
    # Q < 1/2: the SHO term splits into two real celerite terms
    lowq = [{"type": "oscillation_bump", "params": {"values": {"P_g": [False, None, "uniform", 10, 500],
                                                             "Q": [False, None, "uniform", 0.05, 2.0],
                                                             "nu_max": [False, None, "uniform", 50, 200]}}}] + wn
    gp = GPModel(lowq)
    spec = build_spec(gp, None)
    theta = np.array([[200.0, 0.3, 120.0, 80.0], [100.0, 0.1, 60.0, 50.0], [300.0, 1.5, 150.0, 90.0],
                      [300.0, 0.49999, 150.0, 90.0], [np.nan, 1.0, 100.0, 50.0], [300.0, 5.0, 150.0, 90.0]])
    t, y, yerr = synthetic_star(700, 4)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y - 1e6, yerr)
    lp, parts = engine.log_prob(theta, return_parts=True)
    want_lp, want = oracle.log_prob(spec, t, y - 1e6, yerr, theta, return_parts=True)
    assert lp[4] == -np.inf and want_lp[4] == -np.inf and lp[5] == -np.inf
    ok = np.isfinite(want_lp)
    assert rel(lp[ok], want_lp[ok]) < 1e-8      # Q -> 1/2 is ill-conditioned (1/f blows up)
    assert_parts_close(parts[:3], want[:3])


def test_multi_star_batch(engine):
    np.random.seed(5)
    gp = GPModel(C3_CONFIG_GP)
    spec = build_spec(gp, None)
    engine.clear_stars()
    engine.set_model(spec)
    stars = [synthetic_star(1024, 100 + s) for s in range(5)]
    for s, (t, y, yerr) in enumerate(stars):
        engine.upload_star(s, t, y, yerr)
    theta = gp.sample_prior(5 * 6).reshape(5, 6, -1)
    lp = engine.log_prob(theta, star=0, nstars=5)
    for s, (t, y, yerr) in enumerate(stars):
        want = oracle.log_prob(spec, t, y, yerr, theta[s])
        assert rel(lp[s], want) < RTOL
    # a sub-range of stars
    lp2 = engine.log_prob(theta[2:4], star=2, nstars=2)
    assert rel(lp2, lp[2:4]) < 1e-12     # chunk geometry depends on the batch size


def test_sampler_matches_oracle_stream(engine, sim_lc, sim_config):
    """Same Philox stream, same proposals, same accept decisions: the chains agree step by step."""
    t, y, yerr, spec, gp, tr = sim_setup(sim_lc, sim_config, "cosi")
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    np.random.seed(42)
    W, nsteps = 24, 12
    init = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
    init[:, 9] = np.random.uniform(5, 15, W)
    init[:, 10] = np.random.uniform(0, 0.05, W)
    chain, logp, nacc = engine.sample(init, nsteps, seed=1234)
    ochain, ologp, onacc = oracle.sample(spec, t, y, yerr, init, nsteps, seed=1234)
    assert chain.shape == (W, nsteps, 11) and logp.shape == (nsteps, W)
    assert np.array_equal(nacc, onacc)
    assert rel(chain, ochain) < 1e-9
    fin = np.isfinite(ologp)
    assert np.array_equal(np.isfinite(logp), fin)
    assert rel(logp[fin], ologp[fin]) < 1e-9
    assert 0 < nacc.sum() < W * nsteps
    # resume: 12 steps == 5 steps + 7 steps from step0=5
    c1, l1, _ = engine.sample(init, 5, seed=1234)
    c2, l2, _ = engine.sample(c1[:, -1, :], 7, seed=1234, step0=5)
    assert np.array_equal(np.concatenate([c1, c2], axis=1), chain)


def test_sampler_recovers_gaussian_posterior(engine):
    """Statistical check: white-noise-only model has a known posterior for the jitter."""
    rng = np.random.default_rng(0)
    N = 400
    t = np.arange(N) * 0.02
    y = rng.normal(0, 100.0, N)
    wn = [{"type": "white_noise", "params": {"values": {"jitter": [False, None, "uniform", 1, 1000]}}},
          {"type": "granulation", "params": {"values": {"a_gran": [False, None, "uniform", 0.01, 50],
                                                        "b_gran": [False, None, "uniform", 10, 100]}}}]
    gp = GPModel(wn)
    spec = build_spec(gp, None)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, np.full(N, 1e-3))
    np.random.seed(1)
    W = 32
    init = gp.sample_prior(W)
    chain, logp, nacc = engine.sample(init, 600, seed=7)
    jit = chain[:, 300:, 0].ravel()
    # jitter posterior ~ sample std of y (the weak granulation term absorbs little)
    assert abs(np.median(jit) - y.std()) < 8.0
    acc = nacc.mean() / 600
    assert 0.1 < acc < 0.8


def test_model_surface_runs_kic_example(kic_lc, kic_config, tmp_path):
    from gptransits_b200 import LightCurve, Model
    t, f, e = kic_lc
    cfg = dict(kic_config, settings=dict(kic_config["settings"], num_steps=40, save=True, checkpoint_every=25))
    lc = LightCurve(t, f * 1e6, e * 1e6)
    m = Model(lc, cfg, wdir=tmp_path, quiet=True)
    m.run()
    assert m.chain.shape == (24, 40, 6) and m.log_prob.shape == (40, 24)
    assert (tmp_path / "data" / "chain.pk").is_file()
    lp0 = m.lnlike_func(m.chain[0, -1])
    assert abs(lp0 - m.log_prob[-1, 0]) <= 1e-9 * abs(lp0)
    # resume: asking for more steps continues from the stored state
    cfg2 = dict(cfg, settings=dict(cfg["settings"], num_steps=60))
    m2 = Model(lc, cfg2, wdir=tmp_path, quiet=True)
    m2.run()
    assert m2.chain.shape == (24, 60, 6)
    assert np.array_equal(m2.chain[:, :40], m.chain)
    # append-only checkpoints: 25 + 15 steps of the first run, 20 of the second, each block written once
    blocks = sorted(p.name for p in (tmp_path / "data" / "chain_state").glob("block_*.npz"))
    assert blocks == ["block_000000.npz", "block_000001.npz", "block_000002.npz"]
    stats = m2.statistics()
    assert stats["params"]["median"].shape == (6,)
    m.close()
    m2.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes
def c3_star(N, seed=2452144, cadence="kepler"):
    from gptransits_b200 import synth
    t, y, s = synth.make_star(N, seed=seed, cadence=cadence)
    return t, 1e6 + y, s


@pytest.mark.parametrize("N,cadence,W", [(65536, "kepler", 6), (1048576, "tess", 3)])
def test_full_size_against_oracle_and_sequential(engine, N, cadence, W):
    """C3 (N = 65 536) and C5 (N = 1 048 576): the time-parallel path against the oracle directly, and
    against the engine's own sequential evaluation (chunks = 1), quad and logdet separately."""
    from gptransits_b200 import synth
    t, y, yerr = c3_star(N, cadence=cadence)
    cfg = synth.c3_config()
    np.random.seed(7)
    gp = GPModel(cfg["gp"])
    tr = BatmanModel(cfg["transit"][0])
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    theta = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    engine.set_option("chunks", 0)
    engine.set_option("halo", 0)
    lp, parts = engine.log_prob(theta, return_parts=True)
    assert engine.stat("chunks") > 1
    want_lp, want_parts = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    assert_parts_close(parts, want_parts)
    engine.set_option("chunks", 1)
    lp1, parts1 = engine.log_prob(theta, return_parts=True)
    engine.set_option("chunks", 0)
    assert_parts_close(parts, parts1, rtol=1e-10)
    # linearity property of the solve: quad(2 r) = 4 quad(r) needs the transit-free model
    spec_gp = build_spec(gp, None)
    engine.set_model(spec_gp)
    engine.upload_star(0, t, y - 1e6, yerr)
    _, pa = engine.log_prob(theta[:, :gp.ndim], return_parts=True)
    engine.upload_star(0, t, 2 * (y - 1e6), yerr)
    _, pb = engine.log_prob(theta[:, :gp.ndim], return_parts=True)
    assert rel(pb[:, 1], 4 * pa[:, 1]) < 1e-10 and rel(pb[:, 2], pa[:, 2]) < 1e-12


def test_c4_many_stars_batch(engine):
    """C4 shape (reduced count): stars x walkers in one launch, each star against the oracle."""
    from gptransits_b200 import synth
    nstars, W = 24, 8
    t, ys, yerr, truths = synth.make_stars(nstars, 4096, seed0=1000)
    cfg = synth.c3_config()
    np.random.seed(11)
    gp = GPModel(cfg["gp"])
    tr = BatmanModel(cfg["transit"][0])
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    engine.clear_stars()
    engine.set_model(spec)
    for s in range(nstars):
        engine.upload_star(s, t, 1e6 + ys[s], yerr)
    theta = np.hstack([gp.sample_prior(nstars * W), tr.sample_prior(nstars * W)]).reshape(nstars, W, -1)
    lp, parts = engine.log_prob(theta, nstars=nstars, return_parts=True)
    for s in range(0, nstars, 5):
        _, want = oracle.log_prob(spec, t, 1e6 + ys[s], yerr, theta[s], return_parts=True)
        assert_parts_close(parts[s], want)


def test_sampler_posterior_agrees_with_cpu_sampler(engine, kic_lc, kic_config):
    """Statistical parity on the bundled KIC example (BASELINE config 1): GPU and CPU stretch-move
    runs with DIFFERENT seeds give the same posterior medians and 68 % widths within Monte-Carlo error."""
    t, y, yerr, spec, gp = kic_setup(kic_lc, kic_config, baseline_removed=True)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    np.random.seed(42)
    W, nsteps, burn = 24, 2500, 1000
    init = gp.sample_prior(W)
    chain, logp, nacc = engine.sample(init, nsteps, seed=101)
    ochain, ologp, onacc = oracle.sample(spec, t, y, yerr, init, nsteps, seed=202, nthreads=8)
    a = chain[:, burn:, :].reshape(-1, gp.ndim)
    b = ochain[:, burn:, :].reshape(-1, gp.ndim)
    med_a, med_b = np.median(a, axis=0), np.median(b, axis=0)
    wid_a = np.percentile(a, 84.15, axis=0) - np.percentile(a, 15.85, axis=0)
    wid_b = np.percentile(b, 84.15, axis=0) - np.percentile(b, 15.85, axis=0)
    # ~36 000 correlated samples each (tau ~ 60 steps): standard error of a median ~ 0.06 sigma
    assert np.all(np.abs(med_a - med_b) < 0.35 * 0.5 * (wid_a + wid_b)), (med_a, med_b, wid_a)
    assert np.all(np.abs(wid_a / wid_b - 1) < 0.3), (wid_a, wid_b)
    assert abs(nacc.mean() - onacc.mean()) / nsteps < 0.05


def test_gp_predict_against_dense_linear_algebra(engine, sim_lc, sim_config):
    """celerite GP.predict semantics: mu* = K*^T K^-1 r, var* = k(0) - K*^T K^-1 K* (dense check)."""
    from scipy.linalg import cho_factor, cho_solve
    t, y, yerr, spec, gp, tr = sim_setup(sim_lc, sim_config, "cosi")
    sel = slice(0, 400)
    t, y = t[sel], y[sel] - 1e6
    yerr = np.full(t.size, 40.0)
    theta = np.array([249.7, 5.19, 105.8, 368.6, 68.3, 88.7, 5.9, 3.0, 0.05, 10.0, 0.02])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    xs = np.concatenate([t[:50] + 0.004, np.linspace(t[0] - 0.3, t[-1] + 0.3, 123)])
    mu, var = engine.gp_predict(theta, xs)
    # dense reference
    coef = oracle.gp_coeffs(spec["comp_type"], theta[:6])

    def kern(tau):
        tau = np.abs(tau)
        k = np.zeros_like(tau)
        for a, b, c, d in zip(coef["a_comp"], coef["b_comp"], coef["c_comp"], coef["d_comp"]):
            k += np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau))
        return k
    x = t * 0.0864
    K = kern(x[:, None] - x[None, :]) + np.diag(yerr ** 2 + coef["jitter"])
    pars = tr.batman_params(theta[6:], inc_mode=1)
    resid = y - (oracle.light_curve(t, pars, 6, spec["exp_time"]) - 1) * 1e6
    Ks = kern(xs[:, None] * 0.0864 - x[None, :])
    cf = cho_factor(K, lower=True)
    mu_d = Ks @ cho_solve(cf, resid)
    var_d = kern(np.zeros(1))[0] - np.einsum("mn,nm->m", Ks, cho_solve(cf, Ks.T))
    assert np.max(np.abs(mu - mu_d)) < 1e-8 * np.max(np.abs(mu_d))
    assert np.max(np.abs(var - var_d)) < 1e-8 * np.max(np.abs(var_d))
    assert np.all(var > 0)
    # outside the prior support -> error, like an invalid kernel in celerite
    bad = theta.copy()
    bad[1] = 0.2
    with pytest.raises(Exception):
        engine.gp_predict(bad, xs)


def test_four_terms_and_odd_walker_counts(engine):
    """J = 8 (four SHO terms, the largest register-resident state) and walker counts that are not a
    multiple of the warp / CTA size (300 walkers = 3 CTAs per chunk, last one partly empty)."""
    cfg4 = C3_CONFIG_GP[:3] + [
        {"type": "oscillation_bump", "params": {"values": {"P_g": [False, None, "uniform", 20, 80],
                                                           "Q": [False, None, "uniform", 1.5, 4.0],
                                                           "nu_max": [False, None, "uniform", 230, 270]}}},
        C3_CONFIG_GP[3]]
    t, y, yerr = synthetic_star(6000, 21)
    np.random.seed(9)
    gp = GPModel(cfg4)
    spec = build_spec(gp, None)
    theta = gp.sample_prior(37)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    for chunks in (1, 7, 0):
        engine.set_option("chunks", chunks)
        _, parts = engine.log_prob(theta, return_parts=True)
        assert_parts_close(parts, want)
    engine.set_option("chunks", 0)
    # 300 walkers, three-term model with transit
    tr = BatmanModel(C3_CONFIG_TR)
    tr.init_model(t, t[1] - t[0], 6)
    gp3 = GPModel(C3_CONFIG_GP)
    spec3 = build_spec(gp3, tr)
    theta3 = np.hstack([gp3.sample_prior(300), tr.sample_prior(300)])
    engine.set_model(spec3)
    engine.upload_star(0, t, y, yerr)
    lp = engine.log_prob(theta3)
    want3 = oracle.log_prob(spec3, t, y, yerr, theta3, nthreads=8)
    assert rel(lp, want3) < RTOL


def test_device_gelman_rubin_matches_host(engine, kic_lc, kic_config):
    from gptransits_b200 import convergence
    t, y, yerr, spec, gp = kic_setup(kic_lc, kic_config, baseline_removed=True)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    np.random.seed(4)
    init = gp.sample_prior(24)
    engine.sampler_begin(init, seed=11)
    engine.sampler_run(120, keep=True)
    chain, _ = engine.sampler_read(0, 120)
    rhat = engine.sampler_rhat(40, 80)
    engine.sampler_end()
    want, _, _ = convergence.gelman_rubin(chain[0][:, 40:120, :])
    assert rel(rhat[0], want) < 1e-10


def test_eccentric_and_odd_transit_parameters_through_the_posterior(engine):
    """Free e and w (Kepler solve + eccentric in-transit window) and parameter corners that normal
    priors allow (negative period, a < 1, planet larger than the star, inclination past 90 degrees):
    the sparse transit path feeding the GP must agree with the oracle, not just gpt_transit_flux."""
    t, y, yerr = synthetic_star(3000, 31)
    tr_cfg = {"params": {"values": {
        "P": [False, None, "normal", 5.9, 2.0], "t0": [False, None, "normal", 3.0, 1.0],
        "Rrat": [False, None, "normal", 0.1, 0.3], "aR": [False, None, "normal", 8.0, 4.0],
        "cosi": [False, None, "normal", 88.0, 3.0],
        "e": [False, None, "uniform", 0.0, 0.7], "w": [False, None, "uniform", 0.0, 360.0],
        "u1": [False, None, "uniform", 0.0, 1.0], "u2": [False, None, "uniform", 0.0, 0.5]}}}
    np.random.seed(17)
    gp = GPModel(C3_CONFIG_GP[:2] + C3_CONFIG_GP[3:])
    tr = BatmanModel(tr_cfg)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    theta = np.hstack([gp.sample_prior(40), tr.sample_prior(40)])
    k = gp.ndim
    theta[0, k:k + 5] = [-4.0, 3.0, 0.1, 8.0, 88.0]     # negative period
    theta[1, k:k + 5] = [5.9, 3.0, 0.1, 0.8, 88.0]      # a < 1
    theta[2, k:k + 5] = [5.9, 3.0, 1.3, 3.0, 85.0]      # planet larger than the star
    theta[3, k:k + 5] = [5.9, 3.0, 0.1, 8.0, 93.0]      # inclination past 90
    theta[4, k:k + 5] = [5.9, 3.0, -0.1, 8.0, 88.0]     # negative radius ratio (batman takes |rp|)
    theta[5, k + 5] = 0.0                                # exactly circular with free w
    theta[6, k + 5] = 5e-6                               # batman's e < 1e-5 shortcut
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    for chunks in (1, 4, 0):
        engine.set_option("chunks", chunks)
        lp, parts = engine.log_prob(theta, return_parts=True)
        assert_parts_close(parts, want)
    engine.set_option("chunks", 0)
    assert np.isfinite(want_lp).sum() >= 30


def test_sampler_speculative_block_is_replayed_with_repair(engine):
    """The sampler runs blocks of steps without the repair pass and replays a block when the chunk
    verification flagged anything (the blocks after that one run with the repair pass until one comes back
    clean).  With a useless halo the first block is replayed; the chain must still be the oracle's, and
    equal to a run with the speculation switched off."""
    t, y, yerr = synthetic_star(6000, 4)
    np.random.seed(8)
    gp = GPModel(C3_CONFIG_GP)
    tr = BatmanModel(C3_CONFIG_TR)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    W, nsteps = 32, 7
    init = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    ochain, ologp, onacc = oracle.sample(spec, t, y, yerr, init, nsteps, seed=77)
    engine.set_option("chunks", 12)
    engine.set_option("halo", 2)
    engine.set_option("adapt_every", 3)          # blocks of 3, 3 and 1 steps
    engine.set_option("speculate", 2)            # speculate whenever the previous block was clean
    try:
        before = engine.stat("spec_replays")
        chain, logp, nacc = engine.sample(init, nsteps, seed=77)
        assert engine.stat("spec_replays") == before + 1
        assert np.array_equal(nacc, onacc)
        assert rel(chain, ochain) < 1e-9
        fin = np.isfinite(ologp)
        assert rel(logp[fin], ologp[fin]) < 1e-9
        engine.set_option("speculate", 0)
        chain2, logp2, nacc2 = engine.sample(init, nsteps, seed=77)
        assert np.array_equal(chain2, chain) and np.array_equal(logp2, logp) and np.array_equal(nacc2, nacc)
    finally:
        engine.set_option("speculate", 1)
        engine.set_option("adapt_every", 32)
        engine.set_option("chunks", 0)
        engine.set_option("halo", 0)


@pytest.mark.parametrize("jitter", [0.0, 3e-11, 5e-9, 5e-7, 1e-5])
def test_regular_step_treatments_agree(engine, jitter):
    """The fast loop advances the generators by the tile's constant cadence when the accumulated timing
    deviation is negligible (order 0) and by a series of order 1..3 in the per-step deviation otherwise;
    option series_order_min forces the higher treatments, 4 being the general loop with exp / sincos.
    All of them must give the oracle's numbers on an exact grid and on grids with timing jitter."""
    t, y, yerr = synthetic_star(9000, 21)
    if jitter:
        t = t + np.random.default_rng(2).uniform(-jitter, jitter, t.size)
    np.random.seed(5)
    gp = GPModel(C3_CONFIG_GP)
    tr = BatmanModel(C3_CONFIG_TR)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    theta = np.hstack([gp.sample_prior(32), tr.sample_prior(32)])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    try:
        for order in (0, 1, 2, 3, 4):
            engine.set_option("series_order_min", order)
            for chunks in (1, 6):
                engine.set_option("chunks", chunks)
                lp, parts = engine.log_prob(theta, return_parts=True)
                assert_parts_close(parts, want)
    finally:
        engine.set_option("series_order_min", 0)
        engine.set_option("chunks", 0)


@pytest.mark.parametrize("bmax", [300.0, 700.0, 1500.0, 4000.0])
def test_fast_decay_shortens_the_scaled_runs(engine, bmax):
    """The scaled state divides by s = exp(-c (x - x_anchor)); a mode that decays fast (large granulation
    frequency: c dt0 up to ~30 here) forces shorter runs between re-anchorings (128 -> 8 steps) and, past
    that, the general loop.  Every run length must reproduce the oracle."""
    t, y, yerr = synthetic_star(5000, 9)
    gp_cfg = [
        {"type": "granulation", "params": {"values": {"a_gran": [False, None, "uniform", 100, 400],
                                                      "b_gran": [False, None, "uniform", 0.5 * bmax, bmax]}}},
        {"type": "granulation", "params": {"values": {"a_gran": [False, None, "uniform", 75, 225],
                                                      "b_gran": [False, None, "uniform", 7.5, 22.5]}}},
        {"type": "white_noise", "params": {"values": {"jitter": [False, None, "uniform", 50, 150]}}},
    ]
    np.random.seed(12)
    gp = GPModel(gp_cfg)
    spec = build_spec(gp, None)
    theta = gp.sample_prior(32)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    assert np.all(np.isfinite(want_lp))
    try:
        for chunks in (1, 5, 0):
            engine.set_option("chunks", chunks)
            lp, parts = engine.log_prob(theta, return_parts=True)
            assert_parts_close(parts, want)
    finally:
        engine.set_option("chunks", 0)


def test_long_transits_overflow_the_scan_buffer(engine):
    """Close-in orbits (a/R* of 1.2-2.5: a fifth to a third of all samples in transit) need more room than
    the epoch scan's shared-memory buffer (an eighth of the light curve): the direct-append path of the
    scan must queue exactly the same samples."""
    t, y, yerr = synthetic_star(6000, 13)
    tr_cfg = {"params": {"values": {
        "P": [False, None, "uniform", 1.5, 4.0], "t0": [False, None, "uniform", 1.0, 2.0],
        "Rrat": [False, None, "uniform", 0.02, 0.12], "aR": [False, None, "uniform", 1.2, 2.5],
        "cosi": [False, None, "uniform", 80.0, 90.0],
        "e": [True, 0.0, "uniform", 0, 1], "w": [True, 90.0, "uniform", 0, 90],
        "u1": [True, 0.4996, "uniform", 0, 1], "u2": [True, 0.0847, "uniform", 0, 1]}}}
    np.random.seed(23)
    gp = GPModel(C3_CONFIG_GP)
    tr = BatmanModel(tr_cfg)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    theta = np.hstack([gp.sample_prior(48), tr.sample_prior(48)])
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    want_lp, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True, nthreads=8)
    assert np.isfinite(want_lp).sum() >= 40
    try:
        for fuse in (1, 0):
            engine.set_option("fuse_scan", fuse)
            lp, parts = engine.log_prob(theta, return_parts=True)
            assert_parts_close(parts, want)
            assert engine.stat("transit_items") > 48 * 6000 / 8      # more than the buffers of all walkers hold
    finally:
        engine.set_option("fuse_scan", 1)


def test_walker_sharded_half_step_calls_reproduce_the_sampler(engine, sim_lc, sim_config):
    """gpt_sampler_half_eval / gpt_sampler_half_accept (the walker-sharded multi-GPU protocol) with one rank
    owning every walker must give the device sampler's chain (the log-probabilities up to the rounding of a
    different halo length: the halo adapts between the two runs)."""
    from gptransits_b200.parallel import run_walker_sharded
    t, y, yerr, spec, gp, tr = sim_setup(sim_lc, sim_config, "cosi")
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    np.random.seed(4)
    W, nsteps = 24, 9
    init = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
    init[:, 9] = np.random.uniform(5, 15, W)
    init[:, 10] = np.random.uniform(0, 0.05, W)
    want, wlogp, _ = engine.sample(init, nsteps, seed=21)
    chain, logp = run_walker_sharded(engine, None, init, nsteps, seed=21)
    assert np.array_equal(chain, want)
    fin = np.isfinite(wlogp)
    assert np.array_equal(np.isfinite(logp), fin) and rel(logp[fin], wlogp[fin]) < 1e-9
