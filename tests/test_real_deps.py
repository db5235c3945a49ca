"""Oracle <-> the reference's real numerical dependencies (celerite, batman-package, emcee).

None of the three is installable in the authoring image (no network, not in the wheelhouse), so every
test here SKIPS today and the oracle's parity for rows G1/G2/T2/T3/S1 stays "unpinned" (DESIGN.md
section 6).  The day the packages (or a baseline/_ref install) are importable these tests pin the
restatement against the arithmetic the reference really calls: model.py:183-187,307,310 (celerite),
transit.py:72-76,91,126 (batman), model.py:270-284 (emcee).
"""
import os
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT

_REF = os.path.join(str(ROOT), "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.insert(0, _REF)

KIC_THETA = np.array([249.74246655, 5.19060939, 105.80836122, 368.63266001, 68.27098486, 88.68944315])


def test_celerite_log_likelihood(kic_lc):
    celerite = pytest.importorskip("celerite")
    from celerite import terms
    t, f, e = kic_lc
    x, yerr, y = t * 0.0864, e * 1e6, (f - 1) * 1e6
    Pg, Q, numax, a, b, jit = KIC_THETA
    k = terms.SHOTerm(log_S0=np.log(Pg / (4 * Q ** 2)), log_Q=np.log(Q), log_omega0=np.log(2 * np.pi * numax))
    g = terms.SHOTerm(log_S0=np.log(np.sqrt(2) / (2 * np.pi) * a ** 2 / b), log_Q=np.log(1 / np.sqrt(2)),
                      log_omega0=np.log(2 * np.pi * b))
    kernel = k + g + terms.JitterTerm(log_sigma=np.log(jit))
    gp = celerite.GP(kernel)
    gp.compute(x, yerr=yerr)
    want = gp.log_likelihood(y)
    st, quad, logdet = oracle.celerite_loglike([1, 0, 2], KIC_THETA, x, yerr ** 2, y)
    assert st == 0
    got = -0.5 * (quad + logdet + x.size * np.log(2 * np.pi))
    assert got == pytest.approx(want, rel=1e-9)
    assert logdet == pytest.approx(gp.solver.log_determinant(), rel=1e-9)


@pytest.mark.parametrize("pars", [
    (5.778, 3.233, 0.0261, 13.83, 88.9, 0.0, 90.0, 0.4996, 0.0847),
    (3.1, 1.2, 0.11, 7.5, 86.0, 0.3, 40.0, 0.3, 0.2),
    (5.778, 3.233, -0.05, 13.83, 88.9, 0.0, 90.0, 0.4996, 0.0847),   # rp < 0: batman's inverse transit
])
def test_batman_light_curve(sim_lc, pars):
    batman = pytest.importorskip("batman")
    t = sim_lc[0]
    p = batman.TransitParams()
    p.per, p.t0, p.rp, p.a, p.inc, p.ecc, p.w = pars[:7]
    p.u, p.limb_dark = [pars[7], pars[8]], "quadratic"
    cad = t[1] - t[0]
    exp_time = cad * (1 - 1 / 6)
    m = batman.TransitModel(p, t, supersample_factor=6, exp_time=exp_time)
    want = m.light_curve(p)
    got = oracle.light_curve(t, np.array(pars), 6, exp_time)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


def test_emcee_stretch_move_statistics():
    """emcee's MT19937 stream cannot be reproduced by a counter-based RNG: the pin is statistical -- emcee
    driving the oracle's posterior and the oracle's own sampler give the same moments and acceptance."""
    emcee = pytest.importorskip("emcee")
    rng = np.random.default_rng(0)
    N, W, nsteps = 200, 16, 3000
    t = np.arange(N) * 0.02
    y = rng.normal(0, 50.0, N)
    yerr = np.full(N, 1e-3)
    spec = {"comp_type": [2], "prior_kind": [0], "prior_a": [1.0], "prior_b": [500.0]}
    init = rng.uniform(20, 100, size=(W, 1))
    s = emcee.EnsembleSampler(W, 1, lambda x: float(oracle.log_prob(spec, t, y, yerr, x[None, :])[0]))
    s.run_mcmc(init, nsteps)
    ref = s.get_chain(discard=500, flat=True)[:, 0]
    chain, _, nacc = oracle.sample(spec, t, y, yerr, init, nsteps, seed=11)
    own = chain[:, 500:, 0].ravel()
    assert abs(own.mean() - ref.mean()) < 0.3
    assert abs(own.std() / ref.std() - 1) < 0.1
    assert abs(nacc.mean() / nsteps - np.mean(s.acceptance_fraction)) < 0.05
