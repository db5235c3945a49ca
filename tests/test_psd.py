"""Kernel power spectrum: component.py:59-63,181-194, gp.py:70-86."""
import json

import numpy as np
import pytest
from scipy.integrate import quad

import oracle
from conftest import GOLDEN, KIC_CONFIG, SIM_CONFIG
from gptransits_b200 import GPModel
from gptransits_b200.component import Granulation, OscillationBump, WhiteNoise


def _lc_time(key):
    with np.load(GOLDEN / "example_lightcurves.npz") as z:
        return z[f"{key}_time"]


@pytest.mark.parametrize("name,key,cfg", [("sim-2452144", "sim", SIM_CONFIG), ("kic-012008916", "kic", KIC_CONFIG)])
def test_get_psd_matches_reference_run(name, key, cfg):
    """GPModel.get_psd against the reference's own get_psd executed on the bundled examples
    (tests/golden/make_psd_golden.py): frequency grid, per-component spectra, both keyword sets."""
    gold = json.loads((GOLDEN / "psd_golden.json").read_text())[name]
    t = _lc_time(key)
    gp = GPModel(cfg["gp"])
    for case in gold:
        freq, psd = gp.get_psd(np.array(case["theta"]), t, **case["kwargs"])
        assert freq.size == case["nfreq"]
        assert [freq[0], freq[-1]] == pytest.approx(case["freq_first_last"], rel=1e-15, abs=0)
        idx = np.array(case["idx"])
        np.testing.assert_allclose(freq[idx], case["freq"], rtol=1e-14, atol=0)
        assert psd.shape == (len(gp.component_vector), freq.size)
        np.testing.assert_allclose(psd[:, idx], case["psd"], rtol=1e-12, atol=0)


def _kernel_ft(comp_type, theta, freq_uhz):
    """(1 / sqrt(2 pi)) int k(|tau|) exp(-i w tau) dtau of the celerite kernel the ORACLE builds for these
    parameters, k(tau) = sum exp(-c tau) (a cos d tau + b sin d tau), by adaptive quadrature."""
    co = oracle.gp_coeffs([comp_type], theta)
    w = 2 * np.pi * freq_uhz
    total = 0.0
    for a, b, c, d in zip(co["a_comp"], co["b_comp"], co["c_comp"], co["d_comp"]):
        def k(tau):
            return np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau))
        # the integrand has decayed to exp(-50) at tau = 50 / c: finite-interval oscillatory rule (QAWO)
        val, _ = quad(k, 0, 50.0 / c, weight="cos", wvar=w, limit=4000, epsabs=1e-13 * abs(a) / c, epsrel=1e-13)
        total += val
    return np.sqrt(2 / np.pi) * total


@pytest.mark.parametrize("cls,ctype,theta", [
    (Granulation, 0, [368.6, 68.27]),
    (Granulation, 0, [150.0, 15.0]),
    (OscillationBump, 1, [249.7, 5.19, 105.8]),
    (OscillationBump, 1, [50.0, 1.2, 180.0]),
])
def test_sho_psd_is_the_fourier_transform_of_the_kernel(cls, ctype, theta):
    """component.get_psd = 2 sqrt(2 pi) x celerite's S(w); S(w) must be the Fourier transform of the very
    kernel whose likelihood the engine evaluates (coefficients a, b, c, d from the oracle)."""
    values = {n: [False, None, "uniform", 0.0, 1e4] for n in cls.parameter_names}
    comp = cls({"params": {"values": values}})
    freq = np.array([0.0, 3.0, 20.0, 60.0, 105.8, 180.0, 283.0])
    got = comp.get_psd(np.array(theta), freq, 1000, 283.0)
    want = np.array([_kernel_ft(ctype, theta, f) for f in freq]) * 2 * np.sqrt(2 * np.pi)
    np.testing.assert_allclose(got, want, rtol=2e-9, atol=0)


def test_white_noise_psd_level():
    comp = WhiteNoise({"params": {"values": {"jitter": [False, None, "uniform", 0.1, 400]}}})
    freq = np.linspace(0, 283.0, 17)
    got = comp.get_psd(np.array([88.7]), freq, 1292, 283.0)
    assert got.shape == freq.shape
    np.testing.assert_array_equal(got, np.full(17, 88.7 ** 2 / 283.0))
