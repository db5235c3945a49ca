"""The oracle against every fixed number available for this path: golden vectors produced by
running the reference's own component.py/gp.py/transit.py (tests/golden/make_golden.py), the
values recorded in SURVEY.md section 8(c), scipy.stats, and the Random123 Philox vectors."""
import numpy as np
import pytest
from scipy.stats import norm, reciprocal, uniform

import oracle
from gptransits_b200 import BatmanModel, GPModel, build_spec

COMP = {"sim-2452144": [1, 0, 2], "kic-012008916": [1, 0, 2]}   # osc bump, granulation, white noise


@pytest.mark.parametrize("case", ["sim-2452144", "kic-012008916"])
def test_celerite_parameter_transform_matches_reference(golden, case):
    g = golden[case]
    for theta, want in zip(g["theta_gp"], g["gp_celerite_params"]):
        got = oracle.gp_celerite_params(COMP[case], theta)
        assert np.max(np.abs(got - np.array(want))) < 1e-14


def spec_for(config):
    gp = GPModel(config["gp"])
    tr = BatmanModel(config["transit"][0]) if "transit" in config else None
    return build_spec(gp, tr), gp, tr


def test_lnprior_matches_reference(golden, sim_config, kic_config):
    for case, cfg in (("sim-2452144", sim_config), ("kic-012008916", kic_config)):
        spec, gp, tr = spec_for(cfg)
        g = golden[case]
        n = gp.ndim
        for theta, want in zip(g["theta_gp"], g["gp_lnprior"]):
            got = oracle.lnprior(spec["prior_kind"][:n], spec["prior_a"][:n], spec["prior_b"][:n], theta)
            assert abs(got - want) < 1e-13
        got = oracle.lnprior(spec["prior_kind"][:n], spec["prior_a"][:n], spec["prior_b"][:n], g["theta_gp_outside"])
        assert got == g["gp_lnprior_outside"] == -np.inf
        if tr is not None:
            for theta, want in zip(g["theta_tr"], g["tr_lnprior"]):
                got = oracle.lnprior(spec["prior_kind"][n:], spec["prior_a"][n:], spec["prior_b"][n:], theta)
                assert abs(got - want) < 1e-13


def test_prior_logpdfs_match_scipy(golden):
    p = golden["priors"]
    for kind, name in ((0, "uniform"), (1, "normal"), (2, "reciprocal")):
        a, b = p["args"][name]
        for x, want in zip(p["x"], p[name]):
            got = oracle.logpdf(kind, a, b, x)
            assert (got == want) if np.isinf(want) else abs(got - want) < 1e-14
    rng = np.random.default_rng(0)
    for x in rng.uniform(-1, 9, 200):
        assert oracle.logpdf(0, 2.0, 7.5, x) == uniform(2.0, 5.5).logpdf(x)
        assert abs(oracle.logpdf(1, 1.5, 0.25, x) - norm(1.5, 0.25).logpdf(x)) < 1e-12
    for x in 10 ** rng.uniform(-4, 0, 200):
        want = reciprocal(0.001, 0.1).logpdf(x)
        got = oracle.logpdf(2, 0.001, 0.1, x)
        assert (got == want) if np.isinf(want) else abs(got - want) < 1e-13
    assert np.isnan(oracle.logpdf(0, 0.0, 1.0, np.nan))
    # closed support at both ends
    assert oracle.logpdf(0, 2.0, 7.5, 2.0) == oracle.logpdf(0, 2.0, 7.5, 7.5) == -np.log(5.5)


def test_kic_example_known_answers(kic_lc):
    """independent semiseparable + dense Cholesky numbers for the KIC example."""
    t, f, e = kic_lc
    theta = [249.74246655, 5.19060939, 105.80836122, 368.63266001, 68.27098486, 88.68944315]
    cel = oracle.gp_celerite_params([1, 0, 2], theta)
    assert np.max(np.abs(cel - [0.84043368, 1.64685111, 6.49950661, 6.10481296, 6.06136192, 4.48514086])) < 1e-8
    x, diag = t * 0.0864, (e * 1e6) ** 2
    st, quad, logdet = oracle.celerite_loglike([1, 0, 2], theta, x, diag, (f - 1) * 1e6)
    assert st == 0
    assert abs(logdet / 14345.578386664782 - 1) < 1e-12
    assert abs(quad / 981.5028166301665 - 1) < 1e-12
    lnl = -0.5 * (quad + logdet + t.size * np.log(2 * np.pi))
    assert abs(lnl / -8850.809186547911 - 1) < 1e-12
    st, quad, logdet = oracle.celerite_loglike([1, 0, 2], theta, x, diag, f * 1e6)   # as shipped, model.py:164
    assert abs(quad / 2508933258.507857 - 1) < 1e-12
    assert abs(-0.5 * (quad + logdet + t.size * np.log(2 * np.pi)) / -1254474989.3117068 - 1) < 1e-12


MA_GOLDEN = {  # (p, d) -> (batman/Hastings mode, exact mode), u = (0.4996, 0.0847)
    (0.026, 0.0): (0.999175025430051, 0.999175025428590), (0.026, 0.5): (0.999231538024111, 0.999231535418748),
    (0.026, 0.973): (0.999538214114249, 0.999538206369038), (0.026, 1.0): (0.999792211548052, 0.999792207900558),
    (0.1, 0.2): (0.987934961841597, 0.987934960083260), (0.1, 0.95): (0.993872157239854, 0.993872157140070),
    (0.1, 1.0): (0.996487662982521, 0.996487659824593), (0.3, 0.5): (0.899579728498162, 0.899579731336630),
    (0.3, 0.85): (0.935347271642844, 0.935347274004294), (0.3, 1.299): (0.999995180361942, 0.999995180144707),
}


def test_mandel_agol_known_answers():
    for (p, d), (hast, exact) in MA_GOLDEN.items():
        assert abs(oracle.quadratic_ld([d], p, 0.4996, 0.0847, mode=0)[0] - hast) < 2e-15
        assert abs(oracle.quadratic_ld([d], p, 0.4996, 0.0847, mode=1)[0] - exact) < 2e-15
    assert np.all(oracle.quadratic_ld([1.026, 1.5, 100.0], 0.026, 0.4996, 0.0847) == 1.0)


def test_transit_parameter_scatter_matches_reference(golden, sim_config):
    g = golden["sim-2452144"]
    tr = BatmanModel(sim_config["transit"][0])
    assert tr.mask.astype(int).tolist() == g["tr_mask"]
    assert list(tr.get_parameters_names()) == g["tr_names"]
    for theta, want in zip(g["theta_tr"], g["tr_batman_params"]):
        assert np.array_equal(tr.batman_params(theta), np.array(want))   # inc := cosi verbatim (transit.py:103)


def test_philox_known_answers():
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert oracle.philox4x32_10(ctr, key).tolist() == want
