"""Run settings (gptransits/settings.py:6-14 and the dict at model.py:41) plus this build's opt-ins."""
import os

__all__ = ["Settings", "DEFAULT_SETTINGS"]

DEFAULT_SETTINGS = {
    "save": True,
    "include_errors": True,
    "seed": 42,
    "num_threads": os.cpu_count(),
    "num_steps": 25000,
    # --- additions (documented in DESIGN.md) ---
    "num_walkers": None,       # default 4*ndim as model.py:230
    "baseline": "reference",   # "reference": y = flux*1e6 (model.py:164); "remove": y = (flux-1)*1e6
    "inc_mapping": "reference",  # "reference": inc[deg] := cosi value (transit.py:103); "cosi": arccos
    "ellip": "batman",         # "batman": Hastings K/E like the reference; "exact"
    "device": 0,
    "checkpoint_every": 1000,
    "rhat_stop": None,         # stop when the per-block Gelman-Rubin V/W (device-side) falls below this
}


class Settings:
    save = True
    include_errors = True
    seed = 42
    num_threads = os.cpu_count()
    num_steps = 25000

    def as_dict(self):
        return {k: getattr(self, k) for k in dir(self)
                if not k.startswith("_") and not callable(getattr(self, k))}
