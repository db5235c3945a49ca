"""gptransits_b200 -- B200-native likelihood + sampling engine behind the gptransits surface.

    from gptransits_b200 import Model
    model = Model("star.lc", "config.py")
    model.run()
    model.analysis()

The numerical path (priors, celerite semiseparable Cholesky, batman transit flux, emcee stretch
move) runs in hand-written sm_100a CUDA kernels behind the C ABI of include/gptransits_b200.h.
There is no CPU fallback: creating an Engine without the built library or without a GPU raises.
"""
from .settings import Settings
from .component import Granulation, OscillationBump, WhiteNoise, ConfigError
from .gp import GPModel
from .transit import BatmanModel
from .lightcurve import LightCurve
from .engine import Engine
from ._lib import GptError
from .model import Model, build_spec

__version__ = "0.1"
__all__ = ["Model", "Settings", "Engine", "GPModel", "BatmanModel", "Granulation", "OscillationBump",
           "WhiteNoise", "LightCurve", "build_spec", "GptError", "ConfigError"]
