"""Generate tests/golden/reference_golden.json by RUNNING the reference's own modules.

Run in the authoring container only (it reads /root/reference):

    python tests/golden/make_golden.py

gptransits/component.py, gp.py and transit.py are imported unmodified from /root/reference.
Their third-party imports that are absent here (celerite, batman) are replaced by empty stub
modules: nothing generated below calls into them -- only the reference's own Python arithmetic
(prior construction and log-pdfs via scipy.stats, prior sampling, the physical -> celerite
parameter transforms, the theta slicing, and the transit parameter scatter) is exercised.
The committed JSON is what the tests read; /root/reference is not needed at test time.
"""
import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "reference_golden.json"


def import_reference():
    for name in ("celerite", "celerite.terms", "batman"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["celerite"].terms = sys.modules["celerite.terms"]

    class _TransitParams:
        pass
    sys.modules["batman"].TransitParams = _TransitParams
    pkg = types.ModuleType("gptransits")
    pkg.__path__ = [str(REF / "gptransits")]

    class Settings:  # examples do `from gptransits import Settings`
        pass
    pkg.Settings = Settings
    sys.modules["gptransits"] = pkg
    comp = importlib.import_module("gptransits.component")
    gp = importlib.import_module("gptransits.gp")
    tr = importlib.import_module("gptransits.transit")
    return comp, gp, tr


def load_example(name):
    ns = {}
    exec((REF / "examples" / name / "config.py").read_text(), ns)
    return ns


def main():
    comp, gp, tr = import_reference()
    out = {"source": "Fill4/gptransits component.py/gp.py/transit.py executed from /root/reference",
           "numpy": np.__version__}
    import scipy
    out["scipy"] = scipy.__version__

    for name in ("sim-2452144", "kic-012008916"):
        ns = load_example(name)
        case = {}
        np.random.seed(42)
        gpm = gp.GPModel(ns["gp"])
        theta_gp = gpm.sample_prior(8)
        case["theta_gp"] = theta_gp.tolist()
        case["gp_names"] = list(map(str, gpm.get_parameters_names()))
        case["gp_lnprior"] = [float(gpm.lnprior(t)) for t in theta_gp]
        case["gp_celerite_params"] = [gpm.get_parameters_celerite(t).tolist() for t in theta_gp]
        # out-of-support and edge points
        edge = theta_gp[0].copy()
        edge[1] = 0.5
        case["theta_gp_outside"] = edge.tolist()
        case["gp_lnprior_outside"] = float(gpm.lnprior(edge))
        if "transit" in ns:
            cfg = dict(ns["transit"])
            vals = dict(cfg["params"]["values"])
            vals["cosi"] = vals.pop("cosib")          # the key the code looks up (transit.py:16,53)
            cfg["params"] = dict(cfg["params"], values=vals)
            trm = tr.BatmanModel(cfg)
            theta_tr = trm.sample_prior(8)
            case["theta_tr"] = theta_tr.tolist()
            case["tr_names"] = list(map(str, trm.get_parameters_names()))
            case["tr_lnprior"] = [float(trm.lnprior(t)) for t in theta_tr]
            case["tr_mask"] = trm.mask.astype(int).tolist()
            scat = []
            for t in theta_tr:
                trm.update_params(t)
                b = trm.batpars
                scat.append([b.per, b.t0, b.rp, b.a, b.inc, b.ecc, b.w, b.u[0], b.u[1]])
            case["tr_batman_params"] = scat
            case["tr_limb_dark"] = trm.batpars.limb_dark
        out[name] = case

    # normal + reciprocal priors on a synthetic component config (reciprocal needs the inverted
    # guard at component.py:38-40 to be bypassed, so the frozen distributions are built directly)
    from scipy.stats import norm, reciprocal, uniform
    pri = {"uniform": (2.0, 7.5), "normal": (1.5, 0.25), "reciprocal": (0.001, 0.1)}
    xs = [0.0005, 0.001, 0.0123, 0.1, 0.2, 1.4, 2.0, 3.3, 7.5, 7.6]
    out["priors"] = {
        "args": pri, "x": xs,
        "uniform": [float(uniform(2.0, 7.5 - 2.0).logpdf(x)) for x in xs],
        "normal": [float(norm(1.5, 0.25).logpdf(x)) for x in xs],
        "reciprocal": [float(reciprocal(0.001, 0.1).logpdf(x)) for x in xs],
    }
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT)

    # the two bundled example light curves (BASELINE configs 1 and 2) as arrays: data, not code
    arrs = {}
    for key, name in (("kic", "kic-012008916"), ("sim", "sim-2452144")):
        t, f, e = np.loadtxt(REF / "examples" / name / f"{name}.lc", unpack=True)
        arrs[f"{key}_time"], arrs[f"{key}_flux"], arrs[f"{key}_flux_err"] = t, f, e
    np.savez_compressed(OUT.parent / "example_lightcurves.npz", **arrs)
    print("wrote example_lightcurves.npz")


if __name__ == "__main__":
    main()
