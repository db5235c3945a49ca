"""BASELINE configs 1 and 2 end to end: Model(lc, config).run() on the GPU at the reference's own
chain length, timed, and the posterior summary compared with the CPU oracle sampler (different
seed, shorter chain) -- medians and 68 % widths must agree within Monte-Carlo error."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import KIC_CONFIG, SIM_CONFIG  # noqa: E402
from gptransits_b200 import LightCurve, Model  # noqa: E402


def summary(chain, burn):
    flat = chain[:, burn:, :].reshape(-1, chain.shape[2])
    return np.median(flat, axis=0), np.percentile(flat, 84.15, axis=0) - np.percentile(flat, 15.85, axis=0)


def run(name, cfg, t, f, e, nsteps, oracle_steps, settings):
    cfg = dict(cfg, settings=dict(cfg["settings"], num_steps=nsteps, save=False, **settings))
    lc = LightCurve(t, f * 1e6, None if e is None else e * 1e6)
    with tempfile.TemporaryDirectory() as d:
        m = Model(lc, cfg, wdir=d, quiet=True)
        m._ensure_engine()
        t0 = time.perf_counter()
        m.run()
        wall = time.perf_counter() - t0
        W, n, ndim = m.chain.shape
        med, wid = summary(m.chain, n // 2)
        init = m.chain[:, 0, :]
        yerr = m.lc.flux_err if cfg["settings"]["include_errors"] else None
        t1 = time.perf_counter()
        ochain, _, onacc = oracle.sample(m.spec, t, m._y(), yerr, init, oracle_steps, seed=777, nthreads=os.cpu_count())
        cpu_wall = time.perf_counter() - t1
        omed, owid = summary(ochain, oracle_steps // 2)
        out = {"example": name, "walkers": W, "steps": n, "ndim": ndim, "gpu_wall_s": wall,
               "gpu_evals_per_s": W * n / wall, "acceptance": float(m.acceptance_fraction.mean()),
               "cpu_oracle_steps": oracle_steps, "cpu_oracle_wall_s": cpu_wall, "cpu_threads": os.cpu_count(),
               "cpu_evals_per_s": W * oracle_steps / cpu_wall,
               "median_gpu": med.tolist(), "median_cpu": omed.tolist(), "width68_gpu": wid.tolist(),
               "width68_cpu": owid.tolist(),
               "max_abs_median_diff_in_sigma": float(np.max(np.abs(med - omed) / (0.5 * (wid + owid)))),
               "max_width_ratio_dev": float(np.max(np.abs(wid / owid - 1)))}
        m.close()
    return out


def main():
    with np.load(os.path.join(ROOT, "tests", "golden", "example_lightcurves.npz")) as z:
        kic = (z["kic_time"], z["kic_flux"], z["kic_flux_err"])
        sim = (z["sim_time"], z["sim_flux"], z["sim_flux_err"])
    res = [run("kic-012008916 (config 1: GP only, 24 walkers, 20000 steps as shipped)", KIC_CONFIG, *kic, 20000, 20000,
               {"baseline": "remove"}),
           run("sim-2452144 (config 2: GP + transit, 64 walkers)", SIM_CONFIG, sim[0], sim[1], None, 6000, 6000,
               {"baseline": "remove", "inc_mapping": "cosi", "num_walkers": 64})]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
