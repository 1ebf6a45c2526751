"""Loads the reference's python config files (configs/cruw_pose/*.py) unchanged.

Mirrors det3d/torchie/utils/config.py:12-100 (`Config.fromfile`, attribute-access `ConfigDict`) without the
`addict` dependency; `munch` and `det3d.utils.config_tool` (imported by every cruw_pose config, unused by the
model path) are stubbed when absent.
"""
import importlib.util
import os
import sys
import types


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'%s' object has no attribute '%s'" % (self.__class__.__name__, name))

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


class Config(object):
    def __init__(self, cfg_dict=None, filename=None):
        cfg_dict = dict() if cfg_dict is None else cfg_dict
        if not isinstance(cfg_dict, dict):
            raise TypeError("cfg_dict must be a dict, but got {}".format(type(cfg_dict)))
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict))
        object.__setattr__(self, "_filename", filename)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError('file "{}" does not exist'.format(filename))
        if not filename.endswith(".py"):
            raise IOError("Only py type are supported here")
        stubs = {}
        if "munch" not in sys.modules and importlib.util.find_spec("munch") is None:
            m = types.ModuleType("munch")
            m.DefaultMunch = type("DefaultMunch", (ConfigDict,), {"fromDict": staticmethod(lambda d: _wrap(d))})
            stubs["munch"] = m
        try:
            importlib.import_module("det3d.utils.config_tool")
        except Exception:
            for n in ("det3d", "det3d.utils", "det3d.utils.config_tool"):
                if n not in sys.modules:
                    stubs[n] = types.ModuleType(n)
                    stubs[n].__path__ = []
            tool = stubs.get("det3d.utils.config_tool") or sys.modules["det3d.utils.config_tool"]
            tool.get_downsample_factor = lambda model_config: 1
        sys.modules.update(stubs)
        try:
            spec_ = importlib.util.spec_from_file_location("_rtpose_cfg_%d" % abs(hash(filename)), filename)
            mod = importlib.util.module_from_spec(spec_)
            spec_.loader.exec_module(mod)
        finally:
            for n in stubs:
                sys.modules.pop(n, None)
        cfg = {k: v for k, v in vars(mod).items() if not k.startswith("__") and not isinstance(v, types.ModuleType)
               and not callable(v)}
        return Config(cfg, filename=filename)

    @property
    def filename(self):
        return self._filename

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __repr__(self):
        return "Config (path: {}): {}".format(self._filename, dict.__repr__(self._cfg_dict))
