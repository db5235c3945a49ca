"""GP kernel components: granulation, oscillation bump, white noise.

Host-side mirror of the reference's component classes (gptransits/component.py:8-194): same
class names, config format, parameter names/latex/units, prior construction and sampling.  The
arithmetic the reference delegates to celerite (term coefficients, Cholesky, log-likelihood)
runs on the GPU; these classes only describe the model and hand the engine a prior table.
Deliberate fixes of reference defects are listed in DESIGN.md.
"""
import numpy as np
from scipy.stats import norm, reciprocal, uniform

PRIOR_KIND = {"uniform": 0, "normal": 1, "reciprocal": 2}
COMP_GRANULATION, COMP_OSCILLATION_BUMP, COMP_WHITE_NOISE = 0, 1, 2


class ConfigError(ValueError):
    """Raised where the reference prints a message and calls sys.exit()."""


def make_prior(entry, what):
    """entry = [fixed, value, dist, arg1, arg2] -> (scipy frozen dist, (kind, a, b))."""
    dist, a, b = entry[2], entry[3], entry[4]
    if dist == "uniform":
        return uniform(a, b - a), (PRIOR_KIND[dist], float(a), float(b))
    if dist == "normal":
        return norm(a, b), (PRIOR_KIND[dist], float(a), float(b))
    if dist == "reciprocal":
        if a == 0:
            raise ConfigError("Reciprocal prior cannot have a starting interval value of 0.0")
        return reciprocal(a, b), (PRIOR_KIND[dist], float(a), float(b))
    raise ConfigError(f"Parameter prior distribution chosen not recognized: {dist} ({what})")


class Component:
    npars = 0
    type_id = -1
    parameter_names = []
    parameter_latex_names = []
    parameter_units = []
    default_name = "Component"

    def __init__(self, config):
        self.name = config.get("name", self.default_name)
        params = config["params"]
        values = params["values"]
        if self.npars != len(values.keys()):
            raise ConfigError(f"Component needs to have {self.npars} parameters")
        self.parameter_names = list(self.parameter_names)
        self.parameter_latex_names = list(params.get("latex_names", self.parameter_latex_names))
        self.parameter_units = list(params.get("units", self.parameter_units))
        if len(self.parameter_latex_names) != self.npars:
            raise ConfigError(f"Component latex names must have size {self.npars}")
        if len(self.parameter_units) != self.npars:
            raise ConfigError(f"Component units must have size {self.npars}")
        self.prior, self.prior_table = [], []
        for pname in self.parameter_names:
            if values[pname][0]:
                raise ConfigError("Fixing parameter values is not yet supported")
            frozen, row = make_prior(values[pname], pname)
            self.prior.append(frozen)
            self.prior_table.append(row)
        self.prior = np.asarray(self.prior)

    def sample_prior(self, num=1):
        # one column per parameter, all walkers at once: same draw order as the reference
        return np.hstack([p.rvs([num, 1]) for p in self.prior])

    def lnprior(self, params):
        return np.sum([self.prior[i].logpdf(params[i]) for i in range(self.prior.size)])

    def get_parameters_celerite(self, params):
        raise NotImplementedError

    def sho_parameters(self, params):
        """(S0, Q, omega0) of the SHO term, or None for white noise."""
        raise NotImplementedError

    def get_psd(self, params, freq, N, nyquist):
        # celerite SHOTerm.get_psd(2 pi f) * 2 sqrt(2 pi)   (component.py:59-63)
        S0, Q, w0 = self.sho_parameters(params)
        w = 2 * np.pi * np.asarray(freq, dtype=float)
        psd = np.sqrt(2 / np.pi) * S0 * w0 ** 4 / ((w * w - w0 * w0) ** 2 + (w0 * w / Q) ** 2)
        return psd * 2 * np.sqrt(2 * np.pi)


class Granulation(Component):
    npars = 2
    type_id = COMP_GRANULATION
    parameter_names = ["a_gran", "b_gran"]
    parameter_latex_names = [r"$a_{gran}$", r"$b_{gran}$"]
    parameter_units = ["ppm", r"$\mu$Hz"]
    default_name = "Granulation"

    def sho_parameters(self, params):
        a, b = params
        return (np.sqrt(2) / (2 * np.pi)) * (a ** 2 / b), 1.0 / np.sqrt(2.0), b * (2 * np.pi)

    def get_parameters_celerite(self, params):
        S, _, w = self.sho_parameters(params)
        return np.log([S, w])


class OscillationBump(Component):
    npars = 3
    type_id = COMP_OSCILLATION_BUMP
    parameter_names = ["P_g", "Q", "nu_max"]
    parameter_latex_names = [r"$P_{g}$", r"$Q_{bump}$", r"$\nu_{max}$"]
    parameter_units = ["ppm", "", r"$\mu$Hz"]
    default_name = "OscillationBump"

    def sho_parameters(self, params):
        P_g, Q, numax = params
        return P_g / (4 * Q ** 2), Q, numax * (2 * np.pi)

    def get_parameters_celerite(self, params):
        return np.log(self.sho_parameters(params))


class WhiteNoise(Component):
    npars = 1
    type_id = COMP_WHITE_NOISE
    parameter_names = ["jitter"]
    parameter_latex_names = [r"White Noise"]
    parameter_units = ["ppm"]
    default_name = "WhiteNoise"

    def sho_parameters(self, params):
        return None

    def get_parameters_celerite(self, params):
        jitter, = params
        return np.log([jitter])

    def get_psd(self, params, freq, N, nyquist):
        jitter, = params
        return np.full(np.size(freq), jitter ** 2 / nyquist)


COMPONENT_TYPES = {"granulation": Granulation, "oscillation_bump": OscillationBump, "white_noise": WhiteNoise}
