"""Generate tests/golden/psd_golden.json by RUNNING the reference's own GPModel.get_psd.

Run in the authoring container only (it reads /root/reference):

    python tests/golden/make_psd_golden.py

gptransits/component.py and gp.py are imported unmodified from /root/reference.  celerite is
absent from this image, so `celerite.terms` is a stub whose SHOTerm.get_psd is celerite's published
power spectrum (Foreman-Mackey et al. 2017, eq. 20; celerite/terms.py SHOTerm.get_psd [recalled]):

    S(w) = sqrt(2/pi) S0 w0^4 / ((w^2 - w0^2)^2 + w0^2 w^2 / Q^2)

Everything else -- the frequency grid (gp.py:70-86), the physical -> (log S0, log Q, log w0)
transforms and kernel construction (component.py:83-108,128-145), the 2 sqrt(2 pi) normalisation
(component.py:59-63) and the white-noise level jitter^2 / nyquist (component.py:181-194) -- is the
reference's own code.  The committed JSON is what tests/test_psd.py reads.
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import REF, import_reference, load_example  # noqa: E402

OUT = Path(__file__).resolve().parent / "psd_golden.json"


class SHOTerm:
    def __init__(self, log_S0, log_Q, log_omega0):
        self.log_S0, self.log_Q, self.log_omega0 = float(log_S0), float(log_Q), float(log_omega0)

    def freeze_parameter(self, name):
        assert name == "log_Q"

    def get_psd(self, w):
        S0, Q, w0 = np.exp(self.log_S0), np.exp(self.log_Q), np.exp(self.log_omega0)
        w2 = np.asarray(w) ** 2
        return np.sqrt(2 / np.pi) * S0 * w0 ** 4 / ((w2 - w0 ** 2) ** 2 + w0 ** 2 * w2 / Q ** 2)


def main():
    comp, gp, _ = import_reference()
    terms = sys.modules["celerite.terms"]
    terms.SHOTerm = SHOTerm
    out = {"source": "Fill4/gptransits gp.py GPModel.get_psd executed from /root/reference with a stub SHOTerm"}
    for name in ("sim-2452144", "kic-012008916"):
        ns = load_example(name)
        np.random.seed(7)
        gpm = gp.GPModel(ns["gp"])
        theta = gpm.sample_prior(2)
        t = np.loadtxt(REF / "examples" / name / f"{name}.lc", unpack=True)[0]
        cases = []
        for th in theta:
            for kw in ({}, {"min_freq": 5.0, "nyquist_mult": 2}):
                freq, psd = gpm.get_psd(th, t, **kw)
                idx = np.unique(np.linspace(0, freq.size - 1, 40).astype(int))
                cases.append({"theta": th.tolist(), "kwargs": kw, "nfreq": int(freq.size),
                              "freq_first_last": [float(freq[0]), float(freq[-1])], "idx": idx.tolist(),
                              "freq": freq[idx].tolist(), "psd": psd[:, idx].tolist()})
        out[name] = cases
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
