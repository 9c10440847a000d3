"""Loader for the reference's python-dict configs (`da_configs/**`).  Only the two mmcv.Config features
they use are implemented: `_base_ = [relative paths]` with recursive dict merge,
and `--cfg-options a.b=c` overrides (tools/DA_train.py:56-65,187-189)."""
import ast
import copy
import os


class ConfigDict(dict):
    """dict with attribute access (mmcv.utils.ConfigDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{name}'")

    def __setattr__(self, name, value):
        self[name] = value


def _to_cfgdict(v):
    if isinstance(v, dict):
        return ConfigDict({k: _to_cfgdict(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_to_cfgdict(x) for x in v)
    return v


def _merge(base, child):
    """Recursive dict merge; the child wins.  `_delete_=True` replaces the base dict."""
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
            out[k] = copy.deepcopy(v)
    return out


def _load_file(path):
    path = os.path.abspath(path)
    with open(path, "r", encoding="utf-8") as f:
        src = f.read()
    scope = {"__file__": path}
    exec(compile(src, path, "exec"), scope)
    cfg = {k: v for k, v in scope.items() if not k.startswith("__") and not callable(v) and not isinstance(v, type(os))}
    bases = cfg.pop("_base_", None)
    if bases:
        if isinstance(bases, str):
            bases = [bases]
        merged = {}
        for b in bases:
            merged = _merge(merged, _load_file(os.path.join(os.path.dirname(path), b)))
        cfg = _merge(merged, cfg)
    return cfg


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg_dict", _to_cfgdict(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename):
        return Config(_load_file(filename), filename=filename)

    def merge_from_dict(self, options):
        """options: {'a.b.c': value} as produced by --cfg-options."""
        for full_key, v in options.items():
            d = self._cfg_dict
            keys = full_key.split(".")
            for k in keys[:-1]:
                d = d.setdefault(k, ConfigDict())
            d[keys[-1]] = _to_cfgdict(v)

    @staticmethod
    def parse_cfg_options(pairs):
        out = {}
        for kv in pairs or []:
            k, v = kv.split("=", 1)
            try:
                out[k] = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                out[k] = v
        return out

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def __repr__(self):
        return f"Config (path: {self.filename}): {dict(self._cfg_dict)!r}"
