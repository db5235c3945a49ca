"""GPModel: the list of kernel components and the layout of the GP part of theta.

Mirror of gptransits/gp.py:9-86.  `spec_entries()` is what the engine needs: component type
ids and the (kind, a, b) prior rows in theta order.
"""
import numpy as np

from .component import COMPONENT_TYPES, ConfigError


class GPModel:
    def __init__(self, config):
        self.component_vector = []
        for entry in config:
            try:
                ctype = entry["type"]
            except KeyError:
                raise ConfigError("Need component type to define component")
            if ctype not in COMPONENT_TYPES:
                raise ConfigError(f"Component type defined not recognized: {ctype}")
            self.component_vector.append(COMPONENT_TYPES[ctype](entry))
        self.ndim = int(np.sum([c.npars for c in self.component_vector])) if self.component_vector else 0

    def _slices(self):
        i = 0
        for c in self.component_vector:
            yield c, slice(i, i + c.npars)
            i += c.npars

    def get_component_names(self):
        return np.array([c.name for c in self.component_vector])

    def get_parameters_celerite(self, params):
        params = np.asarray(params, dtype=float)
        return np.concatenate([c.get_parameters_celerite(params[s]) for c, s in self._slices()])

    def get_parameters_names(self):
        return np.hstack([c.parameter_names for c in self.component_vector])

    def get_parameters_latex(self):
        return np.hstack([c.parameter_latex_names for c in self.component_vector])

    def get_parameters_units(self):
        return np.hstack([c.parameter_units for c in self.component_vector])

    def lnprior(self, params):
        return sum(c.lnprior(params[s]) for c, s in self._slices())

    def sample_prior(self, num=1):
        return np.hstack([c.sample_prior(num) for c in self.component_vector])

    def get_psd(self, params, time, min_freq=0.0, nyquist_mult=1):
        # frequency grid of gp.py:70-86 (muHz; time in days)
        days_to_microsec = (24 * 3600) / 1e6
        cadence = (time[1] - time[0]) * days_to_microsec
        nyquist = 1 / (2 * cadence)
        f_sampling = 1 / ((time[-1] - time[0]) * days_to_microsec)
        top = nyquist * nyquist_mult
        freq = np.linspace(min_freq, top, int((top - min_freq) / f_sampling) + 1)
        psd = [c.get_psd(np.asarray(params)[s], freq, np.size(time), top) for c, s in self._slices()]
        return freq, np.asarray(psd)

    # ---- engine description
    def spec_entries(self):
        comp_type = [c.type_id for c in self.component_vector]
        rows = [row for c in self.component_vector for row in c.prior_table]
        return comp_type, rows
