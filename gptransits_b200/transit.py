"""BatmanModel: the nine batman transit parameters, which are free/fixed, and their priors.

Mirror of gptransits/transit.py:13-127.  The flux itself (batman's _rsky + _quadratic_ld with 6x
supersampling) is computed by the CUDA transit kernel; `compute()` here calls the engine.
"""
import numpy as np

from .component import ConfigError, make_prior

PARAMETER_NAMES = ["P", "t0", "Rrat", "aR", "cosi", "e", "w", "u1", "u2"]


class BatmanModel:
    npars = 9

    def __init__(self, config):
        self.parameter_names = np.array(PARAMETER_NAMES)
        self.parameter_latex_names = np.array(["Period", "Epoch", r"$R_p / R_{\star}$", r"a / $R_{\star}$",
                                               r"$\cos (i)$", r"$e$", r"$w$", r"$u_1$", r"$u_2$"])
        self.parameter_units = np.array(["days", "days", "", "", "", "", "", "", ""])
        self.mask = np.full(self.npars, True)
        self.name = config.get("name", "BatmanTransit")
        if "params" not in config:
            raise ConfigError("Can't define transit model. No parameters in the configuration")
        values = dict(config["params"]["values"])
        if "cosib" in values and "cosi" not in values:  # spelling used by the bundled example
            values["cosi"] = values.pop("cosib")
        if self.npars != len(values.keys()):
            raise ConfigError(f"Transit model needs to have {self.npars} parameters")
        for key, attr in (("latex_names", "parameter_latex_names"), ("units", "parameter_units")):
            if key in config["params"]:
                if len(config["params"][key]) != self.npars:
                    raise ConfigError(f"Transit parameters {key} must have size {self.npars}")
                setattr(self, attr, np.array(config["params"][key]))
        self.prior, self.prior_table = [], []
        self.pars = np.zeros(self.npars)
        for i, pname in enumerate(self.parameter_names):
            entry = values[pname]
            if entry[0]:
                self.mask[i] = False
                self.pars[i] = entry[1]
            else:
                frozen, row = make_prior(entry, pname)
                self.prior.append(frozen)
                self.prior_table.append(row)
        self.prior = np.asarray(self.prior)
        self.ndim = int(np.sum(self.mask))
        # u2 fixed to 0 is the reference's "linear" law; identical flux from the quadratic law
        self.limb_dark = "linear" if (not self.mask[-1] and self.pars[-1] == 0.0) else "quadratic"
        self.supersample_factor = 1
        self.exp_time = 0.0
        self._engine = None

    def get_parameters_names(self):
        return self.parameter_names[self.mask]

    def get_parameters_latex(self):
        return self.parameter_latex_names[self.mask]

    def get_parameters_units(self):
        return self.parameter_units[self.mask]

    def init_model(self, time, cadence, supersample_factor):
        # the reference draws one prior sample here (transit.py:89); keep the RNG in step with it
        if self.prior.size:
            self.update_params(self.sample_prior()[0].T)
        self.supersample_factor = int(supersample_factor)
        self.exp_time = cadence - (cadence / supersample_factor)

    def sample_prior(self, num=1):
        return np.hstack([p.rvs([num, 1]) for p in self.prior])

    def update_params(self, params):
        self.pars[self.mask] = params

    def batman_params(self, params, inc_mode=0):
        """Nine values in batman's semantics (per,t0,rp,a,inc[deg],ecc,w[deg],u1,u2)."""
        self.update_params(params)
        out = self.pars.copy()
        if inc_mode == 1:
            out[4] = np.degrees(np.arccos(out[4]))
        return out

    def lnprior(self, params):
        return np.sum([self.prior[i].logpdf(params[i]) for i in range(self.prior.size)])

    def bind(self, engine, star=0, inc_mode=0):
        self._engine, self._star, self._inc_mode = engine, star, inc_mode

    def compute(self, params, time=None):
        """(flux - 1) * 1e6 at the bound star's sample times, on the GPU."""
        if self._engine is None:
            raise RuntimeError("BatmanModel is not bound to an engine")
        flux = self._engine.transit_flux(self.batman_params(params, self._inc_mode)[None, :], star=self._star)[0]
        return (flux - 1) * 1e6

    def spec_entries(self):
        return self.pars.copy(), self.mask.astype(np.uint8), list(self.prior_table)
