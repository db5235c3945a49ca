"""GPU parity, round 2: the headline (bench) geometry, batman's inverse transit, wide ensembles."""
import numpy as np
import pytest

import oracle
from gptransits_b200 import BatmanModel, GPModel, build_spec, synth
from test_gpu_parity import C3_CONFIG_GP, C3_CONFIG_TR, assert_parts_close, rel, synthetic_star

pytestmark = pytest.mark.gpu


def c3_model(t, with_transit=True, seed=7, W=32):
    cfg = synth.c3_config()
    np.random.seed(seed)
    gp = GPModel(cfg["gp"])
    tr = None
    if with_transit:
        tr = BatmanModel(cfg["transit"][0])
        tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(gp, tr)
    blocks = [gp.sample_prior(W)] + ([tr.sample_prior(W)] if tr is not None else [])
    return spec, np.hstack(blocks)


def test_c3_half_ensemble_at_the_bench_geometry(engine):
    """The headline shape: N = 65 536, 128 walkers per launch (a half-ensemble of 256), automatic geometry
    and adapted halo; 16 of the walkers against the oracle, quad and log-det separately."""
    t, y, yerr = synth.make_star(65536, seed=2452144)
    y = 1e6 + y
    spec, theta = c3_model(t, True, seed=42, W=128)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    for _ in range(3):                                 # let the halo adapt as it does in the sampler
        lp, parts = engine.log_prob(theta, return_parts=True)
    assert engine.stat("chunks") >= 32
    idx = np.arange(0, 128, 8)
    _, want = oracle.log_prob(spec, t, y, yerr, theta[idx], return_parts=True, nthreads=8)
    assert_parts_close(parts[idx], want)
    # and the whole 256-walker ensemble through the sampler: stored log-probabilities are those of the
    # stored positions
    spec, init = c3_model(t, True, seed=43, W=256)
    chain, logp, _ = engine.sample(init, 3, seed=5)
    sub = np.arange(0, 256, 32)
    want_lp = oracle.log_prob(spec, t, y, yerr, chain[sub, -1, :], nthreads=8)
    assert rel(logp[-1, sub], want_lp) < 1e-9


def test_inverse_transit_for_negative_radius_ratio(engine, sim_lc):
    """batman returns 2 - lc for rp < 0 (an "inverse transit"); a normal prior on Rrat reaches it."""
    t = sim_lc[0]
    pars = np.array([[5.778, 3.233, 0.05, 13.83, 88.9, 0.0, 90.0, 0.4996, 0.0847],
                     [5.778, 3.233, -0.05, 13.83, 88.9, 0.0, 90.0, 0.4996, 0.0847]])
    spec = build_spec(GPModel(C3_CONFIG_GP), None)
    tr = BatmanModel(C3_CONFIG_TR)
    tr.init_model(t, t[1] - t[0], 6)
    spec = build_spec(GPModel(C3_CONFIG_GP), tr)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, np.zeros(t.size), None)
    flux = engine.transit_flux(pars)
    for k in range(2):
        want = oracle.light_curve(t, pars[k], 6, spec["exp_time"])
        assert np.max(np.abs(flux[k] - want)) < 1e-9
    assert np.max(np.abs(flux[1] - (2.0 - flux[0]))) < 1e-15 and flux[1].max() > 1.001
    # through the posterior: Rrat ~ normal(0, 0.05)
    cfg = {"params": {"values": dict(C3_CONFIG_TR["params"]["values"], Rrat=[False, None, "normal", 0.0, 0.05])}}
    tr2 = BatmanModel(cfg)
    tr2.init_model(t, t[1] - t[0], 6)
    gp = GPModel(C3_CONFIG_GP)
    spec2 = build_spec(gp, tr2)
    np.random.seed(31)
    theta = np.hstack([gp.sample_prior(48), tr2.sample_prior(48)])
    assert (theta[:, gp.ndim + 2] < 0).sum() >= 10
    rng = np.random.default_rng(3)
    y = 1e6 + rng.normal(0, 150, t.size)
    engine.set_model(spec2)
    engine.upload_star(0, t, y, None)
    _, parts = engine.log_prob(theta, return_parts=True)
    _, want = oracle.log_prob(spec2, t, y, None, theta, return_parts=True, nthreads=8)
    assert_parts_close(parts, want)


def test_device_gelman_rubin_with_512_walkers(engine):
    """More walkers than the diagnostic kernel's block size (its shared-memory rows must not alias)."""
    from gptransits_b200 import convergence
    rng = np.random.default_rng(1)
    N = 300
    t = np.arange(N) * 0.02
    y = rng.normal(0, 50.0, N)
    wn = [{"type": "white_noise", "params": {"values": {"jitter": [False, None, "uniform", 1, 300]}}}]
    gp = GPModel(wn)
    spec = build_spec(gp, None)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, np.full(N, 1e-3))
    np.random.seed(2)
    init = gp.sample_prior(512)
    engine.sampler_begin(init, seed=5)
    engine.sampler_run(40, keep=True)
    chain, _ = engine.sampler_read(0, 40)
    rhat = engine.sampler_rhat(10, 30)
    engine.sampler_end()
    want, _, _ = convergence.gelman_rubin(chain[0][:, 10:40, :])
    assert rel(rhat[0], want) < 1e-10

def test_flagged_chunks_are_repaired_whatever_the_geometry(engine):
    """A flagged chunk is redone from the carry its predecessor left, unless the predecessor was flagged too: then
    nothing vouches for that carry and the walker is recomputed whole.  A useless halo gets chunks flagged all
    over, with chunks shorter and longer than the halo; the results must still be the oracle's, through the
    posterior call and through the sampler (asynchronous repair)."""
    t, y, yerr = synthetic_star(4096, 6)
    np.random.seed(9)
    gp = GPModel(C3_CONFIG_GP)
    spec = build_spec(gp, None)
    theta = gp.sample_prior(16)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, y, yerr)
    _, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True)
    try:
        engine.set_option("chunks", 64)
        engine.set_option("halo", 400)                 # chunks of 64 samples behind a halo of 400
        _, parts = engine.log_prob(theta, return_parts=True)
        assert engine.stat("chunks") == 64
        assert_parts_close(parts, want)
        engine.set_option("halo", 8)                   # far too short: chunks flagged all over
        before = engine.stat("pass2_chunks")
        _, parts = engine.log_prob(theta, return_parts=True)
        assert engine.stat("pass2_chunks") > before + 16
        assert_parts_close(parts, want)
        # chunks longer than the (useless) halo: the chunk-by-chunk repair
        engine.set_option("chunks", 8)
        engine.set_option("halo", 2)
        before = engine.stat("pass2_chunks")
        _, parts = engine.log_prob(theta, return_parts=True)
        assert engine.stat("pass2_chunks") > before
        assert_parts_close(parts, want)
        # the same through the sampler (asynchronous repair launches), both repair forms
        ochain, ologp, onacc = oracle.sample(spec, t, y, yerr, theta, 4, seed=13)
        engine.set_option("speculate", 0)
        for chunks, halo in ((64, 8), (8, 2)):
            engine.set_option("chunks", chunks)
            engine.set_option("halo", halo)
            chain, logp, nacc = engine.sample(theta, 4, seed=13)
            assert np.array_equal(nacc, onacc) and rel(chain, ochain) < 1e-9
    finally:
        engine.set_option("speculate", 1)
        engine.set_option("chunks", 0)
        engine.set_option("halo", 0)


def test_sharded_sampler_entry_point_on_one_rank(engine, kic_lc, kic_config):
    """gpt_sampler_run_sharded with a single rank (no communicator) is the device sampler: the same chain as
    gpt_sample, step by step -- the all-gather is the only thing more ranks add (tests/test_gpu_multi.py)."""
    from gptransits_b200.parallel import run_walker_sharded
    t, f, e = kic_lc
    gp = GPModel(kic_config["gp"])
    spec = build_spec(gp, None)
    np.random.seed(3)
    init = gp.sample_prior(24)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, (f - 1) * 1e6, e * 1e6)
    want, wlogp, _ = engine.sample(init, 6, seed=9)
    chain, logp, ms = run_walker_sharded(engine, None, init, 6, seed=9, timed=True)
    assert ms.shape == (6,) and np.all(ms > 0)
    assert rel(chain, want) < 1e-9 and rel(logp, wlogp) < 1e-9


def test_chain_egress_while_sampling_equals_the_read_at_the_end(engine, kic_lc, kic_config):
    """gpt_sample streams finished step ranges to the host buffers on a second stream while the sampler runs
    (SURVEY.md 8 f3); the result is the chain a single read at the end returns -- with the repair pass in line
    and with speculative blocks (which are handed over only once their record is clean)."""
    t, f, e = kic_lc
    gp = GPModel(kic_config["gp"])
    spec = build_spec(gp, None)
    np.random.seed(5)
    init = gp.sample_prior(24)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, (f - 1) * 1e6, e * 1e6)
    try:
        engine.set_option("halo", 640)   # a fixed geometry: the runs below are then bit-identical
        engine.set_option("chunks", 4)
        engine.set_option("async_egress", 0)
        want, wlogp, wacc = engine.sample(init, 150, seed=11)
        for speculate in (1, 2):
            engine.set_option("async_egress", 1)
            engine.set_option("speculate", speculate)
            chain, logp, nacc = engine.sample(init, 150, seed=11)
            assert engine.stat("egress_jobs") >= 2
            assert np.array_equal(chain, want) and np.array_equal(logp, wlogp) and np.array_equal(nacc, wacc)
    finally:
        engine.set_option("speculate", 1)
        engine.set_option("async_egress", 1)
        engine.set_option("halo", 0)
        engine.set_option("chunks", 0)


def test_batch_upload_on_one_grid_equals_star_by_star(engine):
    """gpt_stars_upload (one time grid, fluxes as a matrix) leaves the engine in the state the per-star uploads
    leave it in: identical log-probabilities, also after one star of the batch is replaced on its own."""
    from gptransits_b200 import synth
    np.random.seed(8)
    gp = GPModel(synth.c3_config()["gp"])
    spec = build_spec(gp, None)
    t, y, yerr, th = synth.make_stars(6, 2048, seed0=77)
    theta = gp.sample_prior(6 * 8).reshape(6, 8, -1)
    engine.clear_stars()
    engine.set_model(spec)
    engine.set_option("chunks", 1)   # one geometry for every call below: the comparisons are then bit for bit
    try:
        for s in range(6):
            engine.upload_star(s, t, y[s], yerr)
        want = engine.log_prob(theta, star=0, nstars=6)
        assert np.all(np.isfinite(want))
        engine.clear_stars()
        engine.upload_stars(0, t, y, yerr)
        got = engine.log_prob(theta, star=0, nstars=6)
        assert np.array_equal(got, want)
        engine.upload_star(3, t, y[0], yerr)
        got3 = engine.log_prob(theta[3:4], star=3, nstars=1)
        engine.clear_stars()
        engine.upload_star(0, t, y[0], yerr)
        assert np.array_equal(got3, engine.log_prob(theta[3:4], star=0, nstars=1))
    finally:
        engine.set_option("chunks", 0)


def test_device_geweke_matches_the_host_definition(engine, kic_lc, kic_config):
    """gpt_sampler_geweke on the device-resident chain == convergence.geweke (the reference's definition,
    convergence.py:75-99) on the same chain pulled to the host."""
    from gptransits_b200.convergence import geweke
    t, f, e = kic_lc
    gp = GPModel(kic_config["gp"])
    spec = build_spec(gp, None)
    np.random.seed(6)
    init = gp.sample_prior(24)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_star(0, t, (f - 1) * 1e6, e * 1e6)
    engine.sampler_begin(init, seed=4)
    try:
        engine.sampler_run(120, keep=True)
        chain, _ = engine.sampler_read(0, 120)
        for first, count, bins in ((0, 120, 10), (17, 90, 7)):
            starts, z = engine.sampler_geweke(first, count, bins=bins)
            hs, hz = geweke(chain[0][:, first:first + count, :], bins=bins)
            assert np.array_equal(starts, hs)
            ok = np.isfinite(hz)
            assert np.array_equal(np.isfinite(z[0]), ok)
            assert np.max(np.abs(z[0][ok] - hz[ok]) / np.maximum(np.abs(hz[ok]), 1e-300)) < 1e-9
    finally:
        engine.sampler_end()


def test_large_batch_cut_into_a_few_chunks(engine):
    """Thousands of walkers cut into a few chunks (what one rank of the C4 batch runs on 4 or 8 GPUs) go through
    the warp-per-walker finalize: parity with the oracle with a sufficient halo, with a useless one (every
    boundary flagged, chunk-by-chunk repair and whole-walker recomputation), and through the sampler."""
    from gptransits_b200 import synth
    np.random.seed(21)
    gp = GPModel(synth.c3_config()["gp"])
    spec = build_spec(gp, None)
    ns, W, N = 32, 32, 2048
    t, y, yerr, th = synth.make_stars(ns, N, seed0=91)
    theta = gp.sample_prior(ns * W).reshape(ns, W, -1)
    engine.clear_stars()
    engine.set_model(spec)
    engine.upload_stars(0, t, y, yerr)
    check = (0, 13, 31)
    want = {s: oracle.log_prob(spec, t, y[s], yerr, theta[s], return_parts=True)[1] for s in check}
    try:
        for chunks, halo, flagged in ((4, 448, False), (4, 4, True), (2, 0, False)):
            engine.set_option("chunks", chunks)
            engine.set_option("halo", halo)
            before = engine.stat("pass2_chunks")
            _, parts = engine.log_prob(theta, star=0, nstars=ns, return_parts=True)
            assert engine.stat("chunks") == chunks
            if flagged:
                assert engine.stat("pass2_chunks") > before
            for s in check:
                assert_parts_close(parts[s], want[s])
        # the sampler's fused accept on the same finalize kernel, with the repair launches in line
        engine.set_option("chunks", 4)
        engine.set_option("halo", 4)
        engine.set_option("speculate", 0)
        chain, logp, nacc = engine.sample(theta, 3, seed=17, nstars=ns)
        for s in check:
            ochain, ologp, onacc = oracle.sample(spec, t, y[s], yerr, theta[s], 3, seed=17, star=s)
            assert np.array_equal(nacc[s], onacc) and rel(chain[s], ochain) < 1e-9
    finally:
        engine.set_option("speculate", 1)
        engine.set_option("chunks", 0)
        engine.set_option("halo", 0)


def test_scratch_cap_slices_large_evaluations(engine, sim_lc, sim_config):
    """An evaluation whose transit-model scratch (16 bytes per walker and sample) exceeds "scratch_cap_mb" runs as
    sub-batches -- whole stars, or slices of one star's walkers -- with the same results."""
    from test_gpu_parity import sim_setup
    t, y, yerr, spec, gp, tr = sim_setup(sim_lc, sim_config, "cosi")
    np.random.seed(12)
    W, ns = 96, 3
    theta = np.hstack([gp.sample_prior(W * ns), tr.sample_prior(W * ns)]).reshape(ns, W, -1)
    theta[..., 9] = np.random.uniform(5, 15, (ns, W))
    theta[..., 10] = np.random.uniform(0, 0.05, (ns, W))
    engine.clear_stars()
    engine.set_model(spec)
    for s in range(ns):
        engine.upload_star(s, t, y + 3.0 * s, yerr)
    try:
        engine.set_option("chunks", 1)
        want, wparts = engine.log_prob(theta, star=0, nstars=ns, return_parts=True)
        assert np.isfinite(want).sum() > ns * W // 2
        per_walker_mb = 16.0 * t.size / 2**20
        launches = engine.stat("launches")
        full = None
        for cap_walkers in (2 * W, 40):       # whole stars per sub-batch; slices of a star's walkers
            engine.set_option("scratch_cap_mb", max(1, int(per_walker_mb * cap_walkers)))
            before = engine.stat("launches")
            got, parts = engine.log_prob(theta, star=0, nstars=ns, return_parts=True)
            if full is None:
                full = before - launches
            assert engine.stat("launches") - before > full     # more than one sub-batch ran
            assert np.array_equal(np.isfinite(got), np.isfinite(want))
            assert_parts_close(parts, wparts, rtol=1e-12)
    finally:
        engine.set_option("scratch_cap_mb", 8192)
        engine.set_option("chunks", 0)


def test_transit_mask_on_edge_shapes(engine):
    """GP + transit on light curves whose length is not a multiple of the 32-bit mask words or of the 128-sample
    tiles, with short periods (several windows per tile), a walker that is in transit everywhere (a/R* ~ 1.6: the
    window covers the whole orbit) and a NaN transit parameter; one chunk and forced chunks with a halo."""
    import copy
    cfg_tr = copy.deepcopy(C3_CONFIG_TR)
    v = cfg_tr["params"]["values"]
    v["P"] = [False, None, "uniform", 0.2, 9.0]
    v["t0"] = [False, None, "uniform", 0.0, 5.0]
    v["Rrat"] = [False, None, "uniform", 0.01, 0.2]
    v["aR"] = [False, None, "uniform", 1.2, 15.0]
    v["cosi"] = [False, None, "uniform", 70.0, 90.0]
    gp = GPModel(C3_CONFIG_GP)
    try:
        for N, chunks, halo in ((1, 0, 0), (2, 0, 0), (31, 0, 0), (33, 0, 0), (127, 0, 0), (129, 0, 0), (257, 0, 0),
                                (1000, 1, 0), (1000, 5, 300), (4101, 9, 333)):
            t, y, yerr = synthetic_star(N, 40 + N)
            tr = BatmanModel(cfg_tr)
            tr.init_model(t, (t[1] - t[0]) if N > 1 else 0.0204, 6)
            spec = build_spec(gp, tr)
            np.random.seed(N)
            W = 9
            theta = np.hstack([gp.sample_prior(W), tr.sample_prior(W)])
            ng = gp.ndim
            theta[0, ng:ng + 5] = [0.31, 0.1, 0.1, 4.0, 88.0]      # several transits per tile
            theta[1, ng:ng + 5] = [0.9, 0.2, 0.15, 1.6, 75.0]      # in transit (almost) everywhere
            theta[2, ng:ng + 5] = [3.0, 1.0, 0.05, 10.0, 89.0]
            theta[3, ng + 3] = np.nan                              # NaN semi-major axis
            engine.clear_stars()
            engine.set_model(spec)
            y = y - 1e6            # no baseline: the transit depth is then a visible share of the residual
            engine.upload_star(0, t, y, yerr)
            engine.set_option("chunks", chunks)
            engine.set_option("halo", halo)
            lp, parts = engine.log_prob(theta, return_parts=True)
            want_lp, want = oracle.log_prob(spec, t, y, yerr, theta, return_parts=True)
            assert np.array_equal(np.isfinite(lp), np.isfinite(want_lp)), N
            assert_parts_close(parts, want)
    finally:
        engine.set_option("chunks", 0)
        engine.set_option("halo", 0)


def test_reupload_on_the_same_time_stamps_keeps_the_grid_analysis(engine):
    """New fluxes / error bars on bit-identical time stamps in the slot last uploaded take the short path of
    gpt_star_upload (counter "grid_reuse"); the result is the one a full upload gives, and any change of the time
    stamps, of the length or of the presence of error bars takes the full path."""
    np.random.seed(14)
    gp = GPModel(C3_CONFIG_GP)
    spec = build_spec(gp, None)
    theta = gp.sample_prior(8)
    t, y, yerr = synthetic_star(3000, 3)
    y2, yerr2 = y + 40.0 * np.sin(t), yerr * 1.3
    engine.clear_stars()
    engine.set_model(spec)
    engine.set_option("chunks", 1)
    try:
        engine.upload_star(0, t, y2, yerr2)
        want = engine.log_prob(theta)
        engine.clear_stars()
        engine.upload_star(0, t, y, yerr)
        n0 = engine.stat("grid_reuse")
        engine.upload_star(0, t.copy(), y2, yerr2)          # same stamps, another buffer
        assert engine.stat("grid_reuse") == n0 + 1
        assert np.array_equal(engine.log_prob(theta), want)
        t3 = t.copy()
        t3[1234] += 1e-9
        engine.upload_star(0, t3, y2, yerr2)                # a changed stamp
        engine.upload_star(0, t3, y2, None)                 # no error bars
        engine.upload_star(0, t3[:-1], y2[:-1], None)       # another length
        assert engine.stat("grid_reuse") == n0 + 1
        _, parts = engine.log_prob(theta, return_parts=True)
        _, wparts = oracle.log_prob(spec, t3[:-1], y2[:-1], None, theta, return_parts=True)
        assert_parts_close(parts, wparts)
    finally:
        engine.set_option("chunks", 0)
