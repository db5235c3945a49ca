"""Minimal light-curve container with the four members the reference reads from
lightkurve.LightCurve (model.py:134,164-165: time, flux, flux_err, meta)."""
import numpy as np


class LightCurve:
    def __init__(self, time, flux, flux_err=None, meta=None):
        self.time = np.asarray(time, dtype=float)
        self.flux = np.asarray(flux, dtype=float)
        self.flux_err = None if flux_err is None else np.asarray(flux_err, dtype=float)
        self.meta = dict(meta or {})
        if self.time.ndim != 1 or self.time.shape != self.flux.shape:
            raise ValueError("time and flux must be 1-D arrays of the same length")


def as_lightcurve(obj):
    """Accept this class or anything lightkurve-like (attributes time/flux/flux_err)."""
    if isinstance(obj, LightCurve):
        return obj
    if all(hasattr(obj, k) for k in ("time", "flux")):
        def arr(v):
            v = getattr(v, "value", v)
            return None if v is None else np.asarray(v, dtype=float)
        return LightCurve(arr(obj.time), arr(obj.flux), arr(getattr(obj, "flux_err", None)),
                          getattr(obj, "meta", None))
    raise TypeError("LC needs to be a lightcurve filepath or LightCurve object")
