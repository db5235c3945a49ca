"""Host-side logic: the mirror of the reference's classes, config handling, diagnostics, the C ABI
surface.  No GPU, no compute calls into the CUDA library."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from gptransits_b200 import BatmanModel, ConfigError, GPModel, GptError, LightCurve, Settings, build_spec
from gptransits_b200 import _lib, convergence, stats
from gptransits_b200.config import load_config
from gptransits_b200.parallel import HalfEnsembleSampler, rng_words, star_block

ROOT = Path(__file__).resolve().parent.parent


def test_prior_sampling_reproduces_reference_draws(golden, sim_config, kic_config):
    """np.random.seed(42) then GPModel.sample_prior(8) / BatmanModel.sample_prior(8): same numbers as
    the reference's own classes produced (tests/golden/make_golden.py)."""
    for case, cfg in (("sim-2452144", sim_config), ("kic-012008916", kic_config)):
        g = golden[case]
        np.random.seed(42)
        gp = GPModel(cfg["gp"])
        assert np.array_equal(gp.sample_prior(8), np.array(g["theta_gp"]))
        assert list(gp.get_parameters_names()) == g["gp_names"]
        assert [float(gp.lnprior(np.array(t))) for t in g["theta_gp"]] == pytest.approx(g["gp_lnprior"], abs=1e-13)
        for t, want in zip(g["theta_gp"], g["gp_celerite_params"]):
            assert np.allclose(gp.get_parameters_celerite(t), want, rtol=0, atol=1e-14)
        if "transit" in cfg:
            tr = BatmanModel(cfg["transit"][0])
            assert np.array_equal(tr.sample_prior(8), np.array(g["theta_tr"]))
            assert tr.limb_dark == g["tr_limb_dark"]


def test_spec_layout(sim_config):
    np.random.seed(0)
    gp = GPModel(sim_config["gp"])
    tr = BatmanModel(sim_config["transit"][0])
    tr.init_model(np.arange(50) / 48.0, 1 / 48.0, 6)
    spec = build_spec(gp, tr)
    assert spec["comp_type"].tolist() == [1, 0, 2]
    assert spec["prior_kind"].tolist() == [0] * 8 + [2, 2, 0]
    assert spec["tr_free"].tolist() == [1, 1, 1, 1, 1, 0, 0, 0, 0]
    assert spec["tr_fixed"][5:].tolist() == [0.0, 90.0, 0.4996, 0.0847]
    assert spec["supersample"] == 6 and abs(spec["exp_time"] - (1 / 48.0 - 1 / 288.0)) < 1e-17
    assert build_spec(gp, tr, {"inc_mapping": "cosi", "ellip": "exact"})["inc_mode"] == 1


def test_config_errors_and_defect_fixes():
    bad = [{"type": "nonsense", "params": {"values": {}}}]
    with pytest.raises(ConfigError):
        GPModel(bad)
    with pytest.raises(ConfigError):
        GPModel([{"params": {}}])
    # appendix A10: reciprocal priors are usable for GP parameters; lower bound 0 is rejected
    ok = [{"type": "white_noise", "params": {"values": {"jitter": [False, None, "reciprocal", 1.0, 100.0]}}}]
    assert GPModel(ok).component_vector[0].prior_table == [(2, 1.0, 100.0)]
    with pytest.raises(ConfigError):
        GPModel([{"type": "white_noise", "params": {"values": {"jitter": [False, None, "reciprocal", 0, 1.0]}}}])
    with pytest.raises(ConfigError):
        GPModel([{"type": "white_noise", "params": {"values": {"jitter": [True, 3.0, "uniform", 0, 1.0]}}}])


def test_example_configs_load(tmp_path):
    """The `.py` module form of the bundled examples (json.load in the reference cannot read them)."""
    text = '''
import os
from gptransits import Settings
settings = Settings()
settings.seed = 7
settings.num_steps = 123
gp = [{"type": "white_noise", "params": {"values": {"jitter": [False, None, "uniform", 1, 2]}}}]
transit = {"type": "batman_model", "params": {"values": {
    "P": [False, None, "uniform", 5.7, 6.2], "t0": [True, 3.0, "uniform", 0, 1], "Rrat": [True, 0.03, "uniform", 0, 1],
    "aR": [True, 10.0, "uniform", 0, 1], "cosib": [True, 88.0, "uniform", 0, 1], "e": [True, 0.0, "uniform", 0, 1],
    "w": [True, 90, "uniform", 0, 90], "u1": [True, 0.5, "uniform", 0, 1], "u2": [True, 0.1, "uniform", 0, 1]}}}
'''
    path = tmp_path / "config.py"
    path.write_text(text)
    cfg = load_config(str(path), tmp_path)
    assert cfg["settings"]["seed"] == 7 and cfg["settings"]["num_steps"] == 123
    assert isinstance(cfg["transit"], list)
    tr = BatmanModel(cfg["transit"][0])     # `cosib` accepted as `cosi`
    assert tr.ndim == 1 and tr.pars[4] == 88.0
    (tmp_path / "config.json").write_text(json.dumps({"gp": cfg["gp"], "settings": {"seed": 1}}))
    assert load_config(None, tmp_path)["settings"]["seed"] == 1
    with pytest.raises(FileNotFoundError):
        load_config(str(tmp_path / "nope.json"), tmp_path)
    assert isinstance(Settings().as_dict(), dict)


def test_convergence_and_stats():
    rng = np.random.default_rng(0)
    chain = rng.normal(size=(12, 400, 3)) * [1.0, 2.0, 0.5] + [0.0, 5.0, -1.0]
    rhat, V, Wv = convergence.gelman_rubin(chain)
    assert rhat.shape == (3,) and np.all(np.abs(rhat - 1) < 0.05)
    assert np.allclose(V / Wv, rhat)
    assert convergence.gelman_brooks(chain).shape == (3,)
    starts, z = convergence.geweke(chain)
    assert z.shape == (10, 12, 3) and starts[0] == 0 and starts[-1] == 200
    lo, hi = stats.hpd(chain, 0.683)
    assert np.all(np.abs((hi - lo) / 2 / np.array([1.0, 2.0, 0.5]) - 1) < 0.1)
    post = rng.normal(size=(400, 12))
    post[17, 5] = 99.0
    assert np.array_equal(stats.mapv(chain, post), chain[5, 17])
    assert np.all(np.abs(stats.mode(chain) - [0.0, 5.0, -1.0]) < [0.4, 0.8, 0.2])


def test_lightcurve_container():
    lc = LightCurve([0, 1, 2], [1.0, 1.1, 0.9], [0.1, 0.1, 0.1])
    assert lc.time.dtype == float and lc.flux_err.shape == (3,)
    with pytest.raises(ValueError):
        LightCurve([0, 1], [1.0])


def header_symbols():
    text = (ROOT / "include" / "gptransits_b200.h").read_text()
    return sorted(set(re.findall(r"\b(gpt_[a-z0-9_]+)\s*\(", text)))


def test_c_abi_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/*.h declares."""
    syms = header_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in _lib.lib().gpt_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gptransits_b200 import Engine
    with pytest.raises(GptError):
        Engine(0)


def test_product_package_never_touches_the_oracle():
    for path in (ROOT / "gptransits_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h"):
            text = path.read_text()
            assert "import oracle" not in text and "from oracle" not in text and "gpt_oracle" not in text, path


def test_star_blocks_cover_everything():
    for nstars in (1, 7, 8, 1024):
        for world in (1, 2, 3, 8):
            blocks = [star_block(nstars, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == nstars
            assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_host_rng_matches_oracle_philox():
    for args in ((42, 0, 0, 0, 0), (2 ** 40 + 17, 3, 2 ** 33 + 5, 255, 1), (7, 1023, 12345, 9, 2)):
        seed, star, step, walker, purpose = args
        ctr = [walker, step & 0xFFFFFFFF, ((step >> 32) & 0x0FFFFFFF) | (purpose << 28), star]
        want = oracle.philox4x32_10(ctr, [seed & 0xFFFFFFFF, seed >> 32]).tolist()
        assert rng_words(seed, star, step, walker, purpose) == want


def test_half_ensemble_sampler_equals_oracle_sampler(kic_lc, kic_config):
    """The walker-sharded host protocol (single process here) reproduces the oracle's chain exactly."""
    t, f, e = kic_lc
    sel = slice(0, 200)
    gp = GPModel(kic_config["gp"])
    spec = build_spec(gp, None)
    np.random.seed(1)
    init = gp.sample_prior(12)
    y, yerr = (f[sel] - 1) * 1e6, e[sel] * 1e6

    def fn(theta):
        return oracle.log_prob(spec, t[sel], y, yerr, theta)
    chain, logp = HalfEnsembleSampler(fn, init, seed=5).run(6)
    ochain, ologp, _ = oracle.sample(spec, t[sel], y, yerr, init, 6, seed=5)
    assert np.array_equal(chain, ochain) and np.array_equal(logp, ologp)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on the small C2 workload."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_chain_store_appends_blocks_and_resumes(tmp_path):
    """The checkpoint store: blocks are written once each, the state survives a reload,
    a legacy single-file checkpoint is read as the first block, and nothing is left half-written."""
    from gptransits_b200.chainstore import ChainStore
    rng = np.random.default_rng(0)
    W, ndim = 6, 3
    legacy = tmp_path / "chain_state.npz"
    c0, l0 = rng.normal(size=(W, 5, ndim)), rng.normal(size=(5, W))
    np.savez(legacy, iteration=5, chain=c0, log_prob=l0)
    store = ChainStore(tmp_path / "chain_state", legacy=legacy)
    assert store.exists()
    st = store.load()
    assert st["iteration"] == 5 and np.array_equal(st["chain"], c0) and not st["converged"]
    blocks = [(rng.normal(size=(W, n, ndim)), rng.normal(size=(n, W))) for n in (4, 7, 2)]
    it = 5
    for k, (c, l) in enumerate(blocks):
        store.append(it, c, l, it + c.shape[1], converged=(k == 2))
        it += c.shape[1]
    store.close()
    names = sorted(p.name for p in (tmp_path / "chain_state").iterdir())
    assert names == ["block_000000.npz", "block_000001.npz", "block_000002.npz", "state.json"]
    mtimes = {n: (tmp_path / "chain_state" / n).stat().st_mtime_ns for n in names[:2]}
    st = ChainStore(tmp_path / "chain_state", legacy=legacy).load()
    assert st["iteration"] == 18 and st["converged"]
    assert np.array_equal(st["chain"], np.concatenate([c0] + [b[0] for b in blocks], axis=1))
    assert np.array_equal(st["log_prob"], np.concatenate([l0] + [b[1] for b in blocks], axis=0))
    # a later append touches only the new block and the state file
    store = ChainStore(tmp_path / "chain_state", legacy=legacy)
    store.append(18, blocks[0][0], blocks[0][1], 22)
    store.close()
    for n, m in mtimes.items():
        assert (tmp_path / "chain_state" / n).stat().st_mtime_ns == m
    assert ChainStore(tmp_path / "chain_state", legacy=legacy).load()["iteration"] == 22
    store.reset()
    assert not store.exists()
