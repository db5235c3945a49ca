"""Configuration loading: dict, JSON file, or the `.py` module form the reference's examples use.

The reference parses the path with json.load (model.py:104-105) although its template and both
bundled examples are Python modules; all three forms are accepted here.
"""
import json
import types
from pathlib import Path

from .settings import Settings


def _exec_py(path):
    import sys
    # the example configs do `from gptransits import Settings`
    shim = types.ModuleType("gptransits")
    shim.Settings = Settings
    saved = sys.modules.get("gptransits")
    sys.modules["gptransits"] = shim
    try:
        ns = {}
        exec(compile(Path(path).read_text(), str(path), "exec"), ns)
    finally:
        if saved is None:
            sys.modules.pop("gptransits", None)
        else:
            sys.modules["gptransits"] = saved
    cfg = {}
    for key in ("settings", "gp", "transit"):
        if key in ns:
            cfg[key] = ns[key]
    return cfg


def load_config(config, wdir):
    if config is None or isinstance(config, (str, Path)):
        path = Path(wdir) / "config.json" if config is None else Path(config)
        if not path.is_file():
            raise FileNotFoundError(f"Config file {path} does not exist in the working directory")
        cfg = _exec_py(path) if path.suffix == ".py" else json.loads(path.read_text())
    elif isinstance(config, dict):
        cfg = dict(config)
    else:
        raise TypeError("Config needs to be a filepath or a configuration dictionary")
    s = cfg.get("settings")
    if s is not None and not isinstance(s, dict):
        cfg["settings"] = s.as_dict() if hasattr(s, "as_dict") else {
            k: getattr(s, k) for k in dir(s) if not k.startswith("_") and not callable(getattr(s, k))}
    t = cfg.get("transit")
    if isinstance(t, dict):   # model.py:173 indexes [0]; the example gives a dict
        cfg["transit"] = [t]
    return cfg
