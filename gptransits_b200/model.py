"""Model: the drop-in for gptransits.Model(lc, config).run()/statistics()/analysis().

Mirrors gptransits/model.py:34-575.  Construction, config/settings handling, light-curve scaling
and the analysis glue stay host Python; `run()` hands the whole stretch-move loop and every
posterior evaluation to the CUDA engine (no per-step round trip), and `lnlike_func` evaluates
single parameter vectors on the GPU for callers that drive their own sampler.

Departures from the reference are the defect fixes of SURVEY.md appendix A (each opt-in where it
would change results): see DESIGN.md "Host surface".
"""
import logging
import pickle
from pathlib import Path

import numpy as np

from .chainstore import ChainStore
from .config import load_config
from .convergence import gelman_brooks, gelman_rubin, geweke
from .engine import Engine
from .gp import GPModel
from .lightcurve import LightCurve, as_lightcurve
from .settings import DEFAULT_SETTINGS
from .stats import hpd, mapv, mode
from .transit import BatmanModel

__all__ = ["Model", "build_spec"]

log = logging.getLogger(__name__)


def build_spec(gp_model, mean_model, settings=None):
    """Flat model description for Engine.set_model (and for the test oracle)."""
    settings = settings or {}
    comp_type, rows = ([], [])
    if gp_model is not None:
        comp_type, rows = gp_model.spec_entries()
        rows = list(rows)
    spec = {"comp_type": np.asarray(comp_type, dtype=np.int32), "has_transit": 0,
            "tr_fixed": np.zeros(9), "tr_free": np.zeros(9, dtype=np.uint8), "supersample": 1, "exp_time": 0.0}
    if mean_model is not None:
        fixed, free, trows = mean_model.spec_entries()
        rows += trows
        spec.update(has_transit=1, tr_fixed=fixed, tr_free=free, supersample=mean_model.supersample_factor,
                    exp_time=mean_model.exp_time)
    spec["prior_kind"] = np.array([r[0] for r in rows], dtype=np.int32)
    spec["prior_a"] = np.array([r[1] for r in rows], dtype=float)
    spec["prior_b"] = np.array([r[2] for r in rows], dtype=float)
    spec["ellip_mode"] = 1 if settings.get("ellip", "batman") == "exact" else 0
    spec["inc_mode"] = 1 if settings.get("inc_mapping", "reference") == "cosi" else 0
    return spec


class Model:
    def __init__(self, lc, config=None, wdir=None, quiet=False, logfile=None, engine=None):
        # ---- logging (model.py:64-84; FileHandler kwarg fixed, appendix A13)
        log.handlers = []
        log.propagate = False
        level = logging.WARN if quiet else logging.INFO
        log.setLevel(level)
        handler = logging.StreamHandler() if logfile is None else logging.FileHandler(logfile, mode="w")
        handler.setLevel(level)
        handler.setFormatter(logging.Formatter("[%(asctime)s] %(name)s %(levelname)s: %(message)s"))
        log.addHandler(handler)

        self.wdir = Path(".") if wdir is None else Path(wdir)
        self.datadir = self.wdir / "data"
        self.figdir = self.wdir / "figures"

        self.config = load_config(config, self.wdir)
        self.settings = dict(DEFAULT_SETTINGS)
        self.settings.update(self.config.get("settings", {}) or {})
        # model.py:128 -- controls the initial prior draws
        np.random.seed(self.settings["seed"])

        if isinstance(lc, (str, Path)):
            self.load_lc_from_file(Path(lc))
        else:
            self.lc = as_lightcurve(lc)

        self.mean_model = None
        self.gp_model = None
        self.engine = engine
        self._owns_engine = engine is None
        self.chain = None
        self.log_prob = None
        self.acceptance_fraction = None
        self.stats = None
        self.stats_flag = False
        self.setup_models()

    # ------------------------------------------------------------------ data
    def load_lc_from_file(self, lc_file):
        log.info(f"Reading lightcurve from: {lc_file.name}")
        data = np.loadtxt(lc_file, unpack=True)
        if data.shape[0] >= 3:
            time, flux, flux_err = data[0], data[1], data[2]
        elif data.shape[0] == 2:
            if self.settings["include_errors"]:
                raise ValueError("Need a flux_err column when include_errors is True")
            time, flux, flux_err = data[0], data[1], None
        else:
            raise ValueError(f"Couldn't load data from the lc_file: {lc_file}")
        # model.py:164 -- ppm scaling without removing the unit baseline (opt-in removal below)
        self.lc = LightCurve(np.copy(time), np.copy(flux) * 1e6, None if flux_err is None else flux_err * 1e6,
                             meta={"lc_file": lc_file})

    def _y(self):
        if self.settings.get("baseline", "reference") == "remove":
            return self.lc.flux - 1e6
        return self.lc.flux

    # ------------------------------------------------------------------ models
    def setup_models(self):
        time = self.lc.time
        if "transit" in self.config and self.config["transit"]:
            self.mean_model = BatmanModel(self.config["transit"][0])
            self.mean_model.init_model(time, time[1] - time[0], 6)   # model.py:174
        if "gp" in self.config and self.config["gp"]:
            self.gp_model = GPModel(self.config["gp"])
            self.gp_model.sample_prior()  # the reference draws once to build its first kernel (model.py:179)
        if self.mean_model is None and self.gp_model is None:
            raise ValueError("Need to define at least a gp or transit model")
        kinds = [k for k, v in (("GP", self.gp_model), ("transit", self.mean_model)) if v is not None]
        log.info("Model with " + " and ".join(kinds))
        self.ndim = (self.gp_model.ndim if self.gp_model else 0) + (self.mean_model.ndim if self.mean_model else 0)
        self.spec = build_spec(self.gp_model, self.mean_model, self.settings)

    def _ensure_engine(self):
        if self.engine is None:
            self.engine = Engine(self.settings.get("device", 0))
            self._owns_engine = True
        if not getattr(self, "_uploaded", False):
            self.engine.set_model(self.spec)
            yerr = self.lc.flux_err if (self.settings["include_errors"] and self.lc.flux_err is not None) else None
            self.engine.upload_star(0, self.lc.time, self._y(), yerr)
            if self.mean_model is not None:
                self.mean_model.bind(self.engine, 0, self.spec["inc_mode"])
            self._uploaded = True
        return self.engine

    # ------------------------------------------------------------------ posterior seam
    def lnlike_func(self, params):
        """log-posterior of one parameter vector (full_lnlike / gp_lnlike / transit_lnlike)."""
        lp = self._ensure_engine().log_prob(np.asarray(params, dtype=float)[None, :])
        return float(lp[0])

    full_lnlike = gp_lnlike = transit_lnlike = lnlike_func

    def log_prob_batch(self, theta, return_parts=False):
        return self._ensure_engine().log_prob(theta, return_parts=return_parts)

    def predict(self, params, x=None, return_var=True):
        """GP conditional mean (and variance) of the residual `flux - transit(params)` at times x
        (default: the sample times) -- what plot.py:250-260 gets from celerite's GP.predict."""
        x = self.lc.time if x is None else np.asarray(x, dtype=float)
        return self._ensure_engine().gp_predict(np.asarray(params, dtype=float), x, return_var=return_var)

    # ------------------------------------------------------------------ sampling
    def sample_initial(self, nwalkers):
        """Prior draws in the reference's order (model.py:257-266): GP block then transit block."""
        blocks = []
        if self.gp_model is not None:
            blocks.append(self.gp_model.sample_prior(num=nwalkers))
        if self.mean_model is not None and self.mean_model.ndim:
            blocks.append(self.mean_model.sample_prior(num=nwalkers))
        return np.hstack(blocks)

    def _state_file(self):
        return self.datadir / "chain_state.npz"

    def _store(self):
        """Append-only checkpoints under <datadir>/chain_state/ (a legacy chain_state.npz is read as block 0)."""
        return ChainStore(self.datadir / "chain_state", legacy=self._state_file())

    def run(self, reset=False, num_threads=1):
        eng = self._ensure_engine()
        nwalkers = self.settings.get("num_walkers") or 4 * self.ndim      # model.py:230
        nsteps = int(self.settings["num_steps"])
        save = bool(self.settings["save"])
        done, chain0, logp0, init = 0, None, None, None
        store = self._store() if save else None
        if save:
            self.datadir.mkdir(parents=True, exist_ok=True)
            if store.exists():
                if reset:
                    log.info("Resetting the backend sampler")
                    store.reset()
                else:
                    st = store.load()
                    done, chain0, logp0 = st["iteration"], st["chain"], st["log_prob"]
                    if done >= nsteps or st["converged"]:
                        log.warning("Skipping run. Backend number of iterations greater than settings"
                                    if done >= nsteps else "Skipping run. The stored run stopped on its rhat_stop criterion")
                        self.chain, self.log_prob = chain0, logp0
                        return
                    if chain0 is not None:
                        init = chain0[:, -1, :]
                        nwalkers = chain0.shape[0]
        if init is None:
            init = self.sample_initial(nwalkers)
        remaining = nsteps - done
        log.info(f"Running MCMC on GPU {eng.device} for {remaining} iterations with {nwalkers} walkers")
        block = int(self.settings.get("checkpoint_every") or remaining) if save else remaining
        block = max(1, min(block, remaining))
        chains, logps = ([chain0] if chain0 is not None else []), ([logp0] if logp0 is not None else [])
        self.block_rhat = []
        self.block_geweke = []
        eng.sampler_begin(init, seed=self.settings["seed"], step0=done)
        try:
            while remaining > 0:
                n = min(block, remaining)
                eng.sampler_run(n, keep=True)
                c, l = eng.sampler_read(0, n)
                chains.append(c[0])
                logps.append(l[0])
                remaining -= n
                done += n
                converged = False
                if n >= 4:
                    # Gelman-Rubin V/W of this block, computed on the device-resident chain
                    rhat = eng.sampler_rhat(0, n)[0]
                    self.block_rhat.append(rhat)
                    log.info(f"iteration {done}: max Gelman-Rubin V/W over the last {n} steps = {rhat.max():.4f}")
                    if n >= 40:
                        # Geweke z-scores of the block (the burn-in criterion of model.py:387-402), also on the device
                        _, zs = eng.sampler_geweke(0, n)
                        frac = float(np.mean(np.nanmax(zs[0][0], axis=1) > 2))
                        self.block_geweke.append(frac)
                        log.info(f"iteration {done}: {100 * frac:.0f}% of the walkers have a Geweke z-score above 2 in this block")
                    stop = self.settings.get("rhat_stop")
                    if stop and rhat.max() < stop:
                        log.info(f"stopping early: V/W below {stop}")
                        remaining = 0
                        converged = True
                if save:
                    # the block goes to disk on the store's thread while the next one is sampled
                    store.append(done - n, c[0], l[0], done, converged)
            _, _, nacc = eng.sampler_state()
        finally:
            eng.sampler_end()
            if store is not None:
                store.close()
        self.chain = np.concatenate(chains, axis=1)          # (W, nsteps, ndim)   model.py:283
        self.log_prob = np.concatenate(logps, axis=0)        # (nsteps, W)         model.py:284
        self.acceptance_fraction = nacc[0] / max(1, self.chain.shape[1] - (chain0.shape[1] if chain0 is not None else 0))
        log.info(f"Mean acceptance fraction: {self.acceptance_fraction.mean():.3f}")
        if save:
            log.info("Saving chain and log_prob")
            with open(self.datadir / "chain.pk", "wb") as f:
                pickle.dump(self.chain, f, protocol=-1)
            with open(self.datadir / "posterior.pk", "wb") as f:
                pickle.dump(self.log_prob, f, protocol=-1)

    # ------------------------------------------------------------------ analysis
    def parameter_latex_names(self):
        names = []
        if self.gp_model is not None:
            names.append(self.gp_model.get_parameters_latex())
        if self.mean_model is not None:
            names.append(self.mean_model.get_parameters_latex())
        return np.hstack(names)

    def statistics(self, fout=None):
        log.info("Running statistics")
        if self.chain is not None and self.log_prob is not None:
            chain, log_prob = self.chain.copy(), self.log_prob.copy()
        else:
            if not self.datadir.is_dir():
                raise FileNotFoundError(f"Directory has no target data to analyse: {self.datadir}")
            with open(self.datadir / "chain.pk", "rb") as f:
                chain = pickle.load(f)
            with open(self.datadir / "posterior.pk", "rb") as f:
                log_prob = pickle.load(f)
        names = self.parameter_latex_names()
        out = {}

        # burn-in from the Geweke z-scores (model.py:387-402)
        starts, z = geweke(chain)
        over = np.max(z, axis=2) > 2
        n_over = over.sum(axis=1)
        k = max(1, int(np.where(n_over == n_over.min())[0][0]))
        walkers = ~over[k]
        geweke_flag = False
        if over[k].sum() / chain.shape[0] > 0.4:
            walkers[:] = True
            geweke_flag = True
        red_chain = chain[walkers, starts[k]:, :]
        red_logp = log_prob[starts[k]:, walkers]
        out.update(geweke_flag=geweke_flag, num_walkers=int(walkers.sum()),
                   percentage_walkers=starts[k] / chain.shape[1])

        # drop walkers stuck at low posterior (model.py:410-413)
        keep = (np.median(red_logp, axis=0) - np.median(red_logp)) > -np.std(red_logp)
        red_chain, red_logp = red_chain[keep], red_logp[:, keep]

        r_hat, _, _ = gelman_rubin(red_chain)
        r_mvar = gelman_brooks(red_chain)
        out["multi_rhat"] = r_mvar.max()
        out["rhat"] = r_hat

        params = {"names": names}
        flat = red_chain.reshape(-1, red_chain.shape[2])
        params["median"] = np.median(flat, axis=0)
        params["hpd_down"], params["hpd_up"] = hpd(red_chain, level=0.683)
        lo99, hi99 = hpd(red_chain, level=0.99)
        params["hpd_99_interval"] = hi99 - lo99
        params["mapv"] = mapv(red_chain, red_logp)
        params["modes"] = mode(chain)
        out["params"] = params

        self.stats = {"params": params, "chain": chain, "log_prob": log_prob, "reduced_chain": red_chain,
                      "reduced_log_prob": red_logp, "diagnostics": out}
        self.stats_flag = True
        if fout is not None:
            with open(fout, "a+") as f:
                f.write(" ".join(f"{v:.10f}" for v in params["median"]) + "\n")
        return self.stats

    def analysis(self, fout=None, plot=True):
        """statistics() + figures.  The figure code of the reference (plot.py) is presentation and out
        of scope; when matplotlib is importable a trace plot is written."""
        if not self.stats_flag:
            self.statistics(fout=fout)
        if plot:
            try:
                import matplotlib
                matplotlib.use("Agg")
                import matplotlib.pyplot as plt
            except ImportError:
                log.info("matplotlib not available: skipping figures")
                return self.stats
            self.figdir.mkdir(parents=True, exist_ok=True)
            chain = self.stats["chain"]
            fig, axes = plt.subplots(chain.shape[2], 1, figsize=(8, 1.5 * chain.shape[2]), sharex=True)
            for d, ax in enumerate(np.atleast_1d(axes)):
                ax.plot(chain[:, ::10, d].T, lw=0.3, alpha=0.5)
                ax.set_ylabel(self.stats["params"]["names"][d])
            fig.savefig(self.figdir / "trace_plot.pdf")
            plt.close(fig)
        return self.stats

    def close(self):
        if self._owns_engine and self.engine is not None:
            self.engine.close()
            self.engine = None
